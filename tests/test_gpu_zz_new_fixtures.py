"""GPU parity on the fixtures added last (-m gpu): two wire types, pass-transistor switches.  Sorted after the other GPU
files on purpose: everything here passes on the CPU warp emulator in the same configuration, but had no B200 run yet when
it was written (the round's GPU budget was spent), and the driver runs the suite with -x."""
import os
import sys

import numpy as np
import pytest

from parallel_eda_b200 import check_route, pfio, router
import parity_bar

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_two_wire_types_and_unbuffered_switches_one_warp():
    """The two fixtures that reach code no other does — mix_w70: eight rr_indexed_data rows (length-1 + length-4 wires) in the
    lookahead; toy_w64 with pass-transistor wire switches: the unbuffered-ancestor branch of the incremental Elmore update
    (route_tree_timing.c:393-417).  One warp, the serial reference's policy, the reference's criticalities replayed; the same
    configuration passes on the CPU warp emulator (tests/test_emu_router.py), and one warp is deterministic."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_oracle_golden import unbuffered_toy
    cases = [(pfio.read_problem(os.path.join(G, "mix_w70.pfp.xz")), pfio.read_result(os.path.join(G, "mix_w70.pfr.xz"))),
             (unbuffered_toy(os.path.join(G, "toy_w64.pfp.xz"), True), pfio.read_result(os.path.join(G, "toy_w64_unbuf_td.pfr.xz")))]
    for p, g in cases:
        cfg = router.default_config(num_slots=1, big_slots=1, pop_slack=0.0, max_batch=1, reroute_all_iters=-1)
        r = router.try_timing_driven_route(p, cfg, sta=router.replay_sta(g))
        assert r.success == 1
        assert check_route.check_route(p, r, check_delays=True)["overused"] == 0
        assert r.total_wirelength <= parity_bar.ONE_WARP_WL * g.total_wirelength
        w = g.iter_crit[-1]
        assert float((w * r.net_delay).sum()) <= parity_bar.TD_TOL * float((w * g.net_delay).sum())


@pytest.mark.parametrize("name", ["het_w70", "mix_w70", "heq_w70"])
def test_new_fixtures_full_concurrency_timing_driven(name):
    """Default configuration (all warps, in-flight bound), timing-driven with the reference's criticalities replayed, on the
    heterogeneous fabric (het), two wire types (mix) and nets that connect twice to one SINK (heq): legal within the reference's
    own iteration budget, every sink delay equal to the from-scratch Elmore recomputation (1e-4), iterations / wirelength /
    criticality-weighted delay inside tests/parity_bar.py."""
    p = pfio.read_problem(os.path.join(G, name + ".pfp.xz"))
    g = pfio.read_result(os.path.join(G, name + ".pfr.xz"))
    w = g.iter_crit[-1]
    r = parity_bar.check_runs("concurrent_td_replay", name, lambda: router.try_timing_driven_route(p, router.default_config(), sta=router.replay_sta(g)), g,
                              weighted=lambda r: (float((w * r.net_delay).sum()), float((w * g.net_delay).sum())))
    assert check_route.check_route(p, r, check_delays=True)["overused"] == 0
