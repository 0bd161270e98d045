"""Full-size parity (-m gpu): a BASELINE configs[4]-density problem routed on the B200 and, IN THE SAME SESSION, by the
reference's own serial router (oracle/_ref/vpr_ref inject, the unmodified try_timing_driven_route on the same flat problem;
the bit-exact C restatement where that binary is absent).  No constant from another machine: legality by the device
check_route, wirelength and iteration count against the CPU result just produced."""
import os
import re
import subprocess
import time

import pytest

from parallel_eda_b200 import check_route, pathfinder, pfio, router
import parity_bar

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_configs4_density_200x200_against_the_reference_router(tmp_path):
    kw = dict(nx=200, ny=200, W=100, num_nets=50000)            # a quarter of configs[4], same density (1.25 nets per CLB)
    p = router.generate_grid_problem(**kw)
    nets, g = router.generate_grid_nets(**kw)
    R = router.Router(nets, generated=g)                          # graph built on the device
    t = time.perf_counter()
    rep = pathfinder.run(R)
    dt = time.perf_counter() - t
    res = R.result()
    res.success = int(rep.success)
    assert rep.success
    dev = R.check_route(res)
    R.close()
    assert dev["ok"] == 1 and dev["overused_nodes"] == 0 and dev["wirelength"] == res.total_wirelength
    assert check_route.check_route_fast(p, res)["overused"] == 0
    # the reference, now
    prob, out = str(tmp_path / "p.pfp"), str(tmp_path / "ref.pfr")
    pfio.write_problem(prob, p)
    ref = os.path.join(ROOT, "oracle", "_ref", "vpr_ref")
    if os.path.exists(ref):
        cmd, who = [ref, "inject", prob, "--result", out], "vpr_ref (unmodified reference)"
    else:
        cli = os.path.join(ROOT, "oracle", "_build", "pf_oracle_cli")
        if not os.path.exists(cli):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        cmd, who = [cli, prob, "--result", out], "pf_oracle_cli (bit-exact restatement)"
    t = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    cpu_s = time.perf_counter() - t
    assert r.returncode == 0, r.stderr[-2000:]
    o = pfio.read_result(out)
    m = re.search(r"route_time_s=([0-9.]+)", r.stderr)
    print("200x200 / 50 k nets: B200 %d iterations, %.3f s, wirelength %d | %s %d iterations, route %.1f s (wall %.1f s), wirelength %d (x%.4f)" % (
        rep.iterations, dt, res.total_wirelength, who, o.iterations, float(m.group(1)) if m else -1.0, cpu_s, o.total_wirelength,
        res.total_wirelength / o.total_wirelength))
    assert o.success == 1

    class G:
        iterations = int(o.iterations); total_wirelength = int(o.total_wirelength)
    res.iterations = rep.iterations
    parity_bar.check("fullsize_200x200_nt", "grid200_50k", res, G, wl_tol=parity_bar.BIG_WL_TOL)
