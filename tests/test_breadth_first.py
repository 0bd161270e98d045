"""Breadth-first router mode (reference route_breadth_first.c, `--router_algorithm breadth_first`):
the device code on the CPU emulator against the goldens of the unmodified reference (tests/golden/*_bf.*).  Route
trees depend on the order equal-cost labels are settled in, so — as for the timing-driven router — parity is a legal
routing (independent check_route), occupancy recomputed from the traces equal to the reported one, and total
wirelength / iteration count close to the reference's."""
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
