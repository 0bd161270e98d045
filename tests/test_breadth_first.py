"""Breadth-first router mode (reference route_breadth_first.c, `--router_algorithm breadth_first`):
the device code on the CPU emulator against the goldens of the unmodified reference (tests/golden/*_bf.*).  Route
trees depend on the order equal-cost labels are settled in, so — as for the timing-driven router — parity is a legal
routing (independent check_route), occupancy recomputed from the traces equal to the reported one, and total
wirelength / iteration count close to the reference's."""
import ctypes
import os

import pytest

from parallel_eda_b200 import check_route, pfio, router

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,slots", [("toy_w64", 1), ("toy_w64", 16)])
def test_breadth_first_on_the_emulator(name, slots, emu_lib):
    p = pfio.read_problem(os.path.join(G, name + "_bf.pfp.xz"))
    g = pfio.read_result(os.path.join(G, name + "_bf.pfr.xz"))
    assert int(p.opts["router_algorithm"]) == 1
    cfg = router.default_config(router.load_library(emu_lib), num_slots=slots, big_slots=2)
    r = router.try_timing_driven_route(p, cfg, lib_path=emu_lib)
    assert r.success == 1
    m = check_route.check_route(p, r, check_delays=False)
    assert m["overused"] == 0 and m["wirelength"] == r.total_wirelength
    print("%s slots %d: %d iterations (reference %d), wirelength x%.3f" % (name, slots, r.iterations, g.iterations, r.total_wirelength / g.total_wirelength))
    assert r.total_wirelength <= 1.08 * g.total_wirelength and r.iterations <= 2 * g.iterations + 2


def test_breadth_first_on_the_heterogeneous_fabric(emu_lib):
    """het_w70 (height-2 hard blocks).  The maze wave floods every net's bounding box, which the fiber emulator pays for
    dearly (a full run: 12 iterations against the reference's 11, wirelength x1.010, 54 s), so the suite routes the first
    PathFinder iteration and checks every route tree (connectivity, switches, sinks, occupancy) without asking for a
    congestion-free result."""
    p = pfio.read_problem(os.path.join(G, "het_w70_bf.pfp.xz"))
    p.opts["max_router_iterations"] = 1
    lib = router.load_library(emu_lib)
    lib.pfb_emu_bucket_refills.restype = ctypes.c_longlong
    before = lib.pfb_emu_bucket_refills()
    r = router.try_timing_driven_route(p, router.default_config(lib, num_slots=8, big_slots=2, far_cap=512), lib_path=emu_lib)
    assert r.iterations == 1
    # the flooding waves outgrow a 512-entry far list, are retried in the big slots and run on the cost buckets there
    # (pf_device.cuh, frontier; BK = 1)
    assert lib.pfb_emu_bucket_refills() > before
    m = check_route.check_route(p, r, check_delays=False, require_legal=False)
    assert m["wirelength"] == r.total_wirelength and m["overused"] == int(r.iter_stats["overused_nodes"][-1])


def test_breadth_first_refuses_nets_that_connect_twice_to_one_sink(emu_lib, oracle_cli, tmp_path):
    """The reference handles a second connection to the same SINK with heap surgery (route_breadth_first.c:208-256,
    invalidate_heap_entries) and its own check_route then rejects the routing ("check_sink: node ... does not connect to any
    terminal", observed on this very circuit), so the mode is not restated: device router and oracle refuse such a problem
    with PF_EINVAL and say why — they must not return a routing, crash, or call it an internal error.  The timing-driven
    router routes the same problem (tests/test_emu_router.py::test_nets_that_connect_twice_to_one_sink)."""
    import subprocess
    p = pfio.read_problem(os.path.join(G, "heq_w70.pfp.xz"))
    p.opts["router_algorithm"] = 1
    p.opts["timing_analysis_enabled"] = 0
    cfg = router.default_config(router.load_library(emu_lib), num_slots=4, big_slots=1)
    with pytest.raises(router.RouterError, match="connects twice to one SINK") as e:
        router.try_timing_driven_route(p, cfg, lib_path=emu_lib)
    assert e.value.code == -4
    pp = str(tmp_path / "p.pfp")
    pfio.write_problem(pp, p)
    r = subprocess.run([oracle_cli, pp], capture_output=True, text=True)
    assert r.returncode != 0 and "rc=-4" in (r.stdout + r.stderr)
