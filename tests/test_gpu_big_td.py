"""GPU parity at BASELINE-scale, timing-driven, heterogeneous fabric (-m gpu).

BASELINE.json configs[1..3] name VTR circuits (stereovision0, bgm, LU32PEEng on k6_frac_N10_mem32K) that are neither in the
reference nor on this box.  Their size-matched stand-ins — 11 k and 32 k LUT generated netlists with hard
multiplier blocks, packed, placed and routed timing-driven by the UNMODIFIED reference on tests/fixtures/k6_N10_het.xml
(tests/golden/big/make_big.sh) — are routed here with the device STA in the loop (nothing crosses PCIe between iterations) and
held to the same bar as every other fixture (tests/parity_bar.py): legal within the reference's 50 iterations, <= 1.5x its
iteration count, wirelength and critical path delay within the stated tolerance of the reference's own result, check_route on
the device, sink delays against a from-scratch Elmore recomputation."""
import json
import os
import time

import pytest

from parallel_eda_b200 import check_route, pfio, router
import parity_bar

pytestmark = pytest.mark.gpu
B = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "big")


@pytest.mark.parametrize("name", ["sv0_w220", "bgm_w260"])
def test_timing_driven_route_at_baseline_scale(name):
    p = pfio.read_problem(os.path.join(B, name + ".pfp.xz"))
    g = pfio.read_timing_graph(os.path.join(B, name + ".pftg.xz"))
    ref = json.load(open(os.path.join(B, name + ".json")))["reference"]
    assert int(p.opts["timing_analysis_enabled"]) == 1 and int(p.opts["max_router_iterations"]) == 50
    t = time.perf_counter()
    r = router.try_timing_driven_route(p, timing_graph=g)
    dt = time.perf_counter() - t
    cpd = float(r.iter_stats["crit_path_delay"][-1])
    print("%s: %d iterations (reference %d), wirelength x%.3f, cpd %.2f ns (reference %.2f), %.2f s incl. upload (reference router: %.0f s)" % (
        name, r.iterations, ref["iterations"], r.total_wirelength / ref["total_wirelength"], cpd, ref["final_crit_path_delay_ns"], dt,
        ref["route_time_s_build_container"]))

    class G:        # the reference's result as far as parity_bar needs it
        iterations = ref["iterations"]; total_wirelength = ref["total_wirelength"]
    parity_bar.check("big_td_device_sta", name, r, G, weighted=(cpd, ref["final_crit_path_delay_ns"]), wl_tol=parity_bar.BIG_WL_TOL)
    R = router.Router(p)
    rep = R.check_route(r)
    R.close()
    assert rep["ok"] == 1 and rep["overused_nodes"] == 0 and rep["wirelength"] == r.total_wirelength
    check_route.check_route(p, r, check_delays=True)       # incremental Elmore delays vs from scratch, every sink
