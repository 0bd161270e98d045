"""GPU parity tests (-m gpu): the CUDA router through the C-ABI versus the CPU oracle / the reference's golden
routings.  Route trees are integer rr-node lists whose shape depends on float-cost ties and on which nets are
in flight together, so — as BASELINE.json's north_star states — parity is: a LEGAL routing (independent
check_route), incremental Elmore delays equal to a from-scratch recomputation (rel. 1e-4, the reference's own
ERROR_TOL), occupancy recomputed from the traces bit-equal to the device's, and total wirelength /
criticality-weighted delay within the stated tolerance of the reference's routing of the same input:
the bar itself — tolerances, iteration budget — is tests/parity_bar.py, one place for every GPU test.
"""
import os
import sys

import numpy as np
import pytest

from parallel_eda_b200 import check_route, pfio, router
import parity_bar

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


TIGHT = ("toy_w64", "hub_w90")   # routed one or two tracks above their minimum channel width


def _load(name, timing):
    p = pfio.read_problem(os.path.join(G, name + ".pfp.xz"))
    p.opts["timing_analysis_enabled"] = 1 if timing else 0
    g = pfio.read_result(os.path.join(G, name + (".pfr.xz" if timing else "_nt.pfr.xz")))
    return p, g


def test_native_library_is_the_one_running():
    lib = router.load_library()
    assert lib.pf_backend_name() == b"cuda:sm_100a" and lib.pf_device_count() >= 1


def test_single_warp_is_bit_identical_to_the_emulated_device_code():
    """One warp is deterministic, and the sm_100a build must compute exactly what the same source computes on
    the CPU warp emulator (tests/test_emu_router.py pins the same constants): identical routing, hence identical
    magic cookie (route_common.c:224-254).  Catches any GPU-only arithmetic or memory-ordering difference."""
    import json
    sw = json.load(open(os.path.join(G, "single_warp_toy.json")))
    SINGLE_WARP_TOY = (sw["serial_num"], sw["total_wirelength"], sw["iterations"])
    p, g = _load("toy_w64", False)
    r = router.try_timing_driven_route(p, router.default_config(num_slots=1, big_slots=1, pop_slack=0.0, max_batch=1))
    assert (r.serial_num, r.total_wirelength, r.iterations) == SINGLE_WARP_TOY


@pytest.mark.parametrize("name", ["toy_w64", "mid_w200"])
def test_single_warp_serial_order_matches_reference(name):
    p, g = _load(name, False)
    r = router.try_timing_driven_route(p, router.default_config(num_slots=1, big_slots=1, pop_slack=0.0, max_batch=1))
    assert r.success == 1
    check_route.check_route(p, r)
    parity_bar.record("one_warp_nt", fixture=name, iterations=int(r.iterations), ref_iterations=int(g.iterations), wl_ratio=r.total_wirelength / g.total_wirelength)
    assert r.total_wirelength <= parity_bar.ONE_WARP_WL * g.total_wirelength
    assert r.iterations <= int(parity_bar.ITER_FACTOR * g.iterations) + 1


@pytest.mark.parametrize("name", ["toy_w64", "mid_w200", "hub_w90"])
def test_concurrent_routing_timing_off(name):
    p, g = _load(name, False)
    r = parity_bar.check_runs("concurrent_nt", name, lambda: router.try_timing_driven_route(p, router.default_config()), g)
    m = check_route.check_route(p, r)
    assert m["overused"] == 0


@pytest.mark.parametrize("name", ["toy_w64", "mid_w200", "hub_w90"])
def test_concurrent_routing_timing_driven(name):
    """Timing-driven mode with the reference's own per-iteration criticalities replayed as the STA."""
    p, g = _load(name, True)
    w = g.iter_crit[-1]
    r = parity_bar.check_runs("concurrent_td_replay", name, lambda: router.try_timing_driven_route(p, router.default_config(), sta=router.replay_sta(g)), g,
                              weighted=lambda r: (float((w * r.net_delay).sum()), float((w * g.net_delay).sum())))
    check_route.check_route(p, r)


def test_step_api_matches_reference_call_sequence():
    """The reference's loop written out with the step functions (route_timing.c:152-310)."""
    p, g = _load("mid_w200", False)
    R = router.Router(p)
    o = p.opts
    pres = float(o["first_iter_pres_fac"])
    for it in range(1, int(o["max_router_iterations"]) + 1):
        st = R.route_iteration(pres)
        assert st.nets_routed > 0
        if it == 1:
            wl, avail = R.total_wirelength()
            assert 0 < wl < 0.85 * avail
        R.reserve_locally_used_opins(pres, it != 1)
        pres, acc = (float(o["initial_pres_fac"]), 0.0) if it == 1 else (pres * float(o["pres_fac_mult"]), float(o["acc_fac"]))
        if R.pathfinder_update_cost(acc) == 0:
            break
    res = R.result()
    res.success = 1
    check_route.check_route(p, res)
    # reset really forgets: a second run from scratch behaves like a first
    R.reset()
    st = R.route_iteration(float(o["first_iter_pres_fac"]))
    assert st.nets_routed == len(p.routed_nets())
    R.close()


def test_large_grid_properties():
    """BASELINE configs[4] density on a 100x100 slice: size-independent properties — legality, occupancy
    recomputed from the traces bit-equal to the device's, every sink reached, sampled edge adjacency."""
    p = router.generate_grid_problem(nx=100, ny=100, W=100, num_nets=12500)
    r = router.try_timing_driven_route(p, router.default_config())
    assert r.success == 1
    m = check_route.check_route_fast(p, r)
    assert m["overused"] == 0 and m["sinks"] == 3 * p.num_nets


def test_generated_grid_full_check_against_oracle(tmp_path):
    """A generated 30x30 problem: full check_route incl. Elmore, and wirelength against the CPU oracle's routing."""
    import subprocess
    p = router.generate_grid_problem(nx=30, ny=30, W=60, num_nets=1500, window=8, seed=3)
    r = router.try_timing_driven_route(p, router.default_config())
    assert r.success == 1
    check_route.check_route(p, r)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "oracle", "_build", "pf_oracle_cli")
    if not os.path.exists(cli):
        subprocess.run(["make", "-s", "-C", os.path.join(root, "oracle")], check=True)
    prob, out = str(tmp_path / "g.pfp"), str(tmp_path / "g.pfr")
    pfio.write_problem(prob, p)
    subprocess.run([cli, prob, "--result", out], check=True, capture_output=True)
    o = pfio.read_result(out)
    assert o.success == 1 and r.total_wirelength <= parity_bar.WL_TOL * o.total_wirelength


def test_overflow_retry_and_small_scratch():
    p, g = _load("toy_w64", False)
    r = router.try_timing_driven_route(p, router.default_config(label_log2=7, far_cap=64, tree_cap=64, big_slots=4))
    assert r.success == 1
    check_route.check_route(p, r)


def test_malformed_graph_is_rejected_with_the_checker_message():
    """The node / edge range checks run inside the parallel upload passes; a bad array must still come back as
    PF_EINVAL with pf_problem_check's message, and the library must stay usable afterwards."""
    p, _ = _load("toy_w64", False)
    keep = int(p.edge_to[5])
    p.edge_to[5] = p.num_nodes + 7
    with pytest.raises(router.RouterError) as e:
        router.try_timing_driven_route(p)
    assert e.value.code == -4 and "invalid problem" in str(e.value)
    p.edge_to[5] = keep
    keep = int(p.capacity[3])
    p.capacity[3] = -1
    with pytest.raises(router.RouterError) as e:
        router.try_timing_driven_route(p)
    assert e.value.code == -4
    p.capacity[3] = keep
    assert router.try_timing_driven_route(p).success == 1


def test_device_check_route_agrees_with_the_python_checker():
    """pf_check_route on the sm_100a build: golden routings of the reference pass, the router's own result passes,
    a corrupted one is rejected with the offending net."""
    p, g = _load("mid_w200", False)
    R = router.Router(p)
    rep = R.check_route(g)
    assert rep["ok"] == 1 and rep["wirelength"] == g.total_wirelength and rep["overused_nodes"] == 0
    r = router.try_timing_driven_route(p)
    rep = R.check_route(r)
    assert rep["ok"] == 1 and rep["wirelength"] == r.total_wirelength == check_route.check_route(p, r)["wirelength"]
    import copy
    bad = copy.deepcopy(r)
    i = int(p.routed_nets()[7]); a = int(r.trace_ptr[i])
    bad.trace_node[a + 1] = bad.trace_node[a]
    rep = R.check_route(bad)
    assert rep["ok"] == 0 and rep["first_bad_net"] == i and rep["first_bad_code"] == 5
    R.close()


def test_stand_alone_cli(tmp_path):
    """python -m parallel_eda_b200 route / check over the flat containers (timing-driven with the device STA)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "duo.pfr")
    r = subprocess.run([sys.executable, "-m", "parallel_eda_b200", "route", os.path.join(G, "duo_w80.pfp.xz"), "--timing-graph",
                        os.path.join(G, "duo_w80.pftg.xz"), "--result", out, "--check"], cwd=root, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["success"] == 1 and rep["check_route"]["ok"] == 1 and rep["check_route"]["overused_nodes"] == 0
    r = subprocess.run([sys.executable, "-m", "parallel_eda_b200", "check", os.path.join(G, "duo_w80.pfp.xz"), out], cwd=root, capture_output=True, text=True)
    assert r.returncode == 0 and json.loads(r.stdout.strip().splitlines()[-1])["wirelength"] == rep["wirelength"]


def test_heterogeneous_fabric_single_warp_equals_the_emulated_device_code():
    """het_w70 (tests/fixtures/k6_N10_het.xml): height-2 hard blocks, i.e. SOURCE / SINK / pin rr nodes spanning two
    tiles.  One warp with the serial reference's policy is deterministic; the sm_100a build must reproduce the routing the
    same source produced on the CPU warp emulator (tests/test_emu_router.py pins the same constants), which is legal, has
    correct Elmore delays and lies within 3 % of the reference's wirelength for this problem."""
    import json
    pin = json.load(open(os.path.join(G, "single_warp_het.json")))["serial_policy"]
    p, g = _load("het_w70", False)
    r = router.try_timing_driven_route(p, router.default_config(num_slots=1, big_slots=1, pop_slack=0.0, max_batch=1, reroute_all_iters=-1))
    assert r.success == 1
    assert check_route.check_route(p, r)["overused"] == 0
    assert r.total_wirelength <= 1.03 * g.total_wirelength
    assert (r.serial_num, r.total_wirelength, r.iterations) == (pin["serial_num"], pin["total_wirelength"], pin["iterations"])



def test_device_built_graph_equals_the_uploaded_one():
    """pf_router_create_generated on the B200: the closed-form generator kernels (pf_gen_device.cuh) produce the very node
    records, edge words and ptc numbers pf_gen.cpp builds on the host and uploads (hashes computed on the device), at a size
    with a million rr nodes; a routing on the generated graph is legal by the independent checker (which reads the host copy)."""
    kw = dict(nx=100, ny=100, W=100, num_nets=12500)
    p = router.generate_grid_problem(**kw)
    nets, g = router.generate_grid_nets(**kw)
    A = router.Router(p)
    B = router.Router(nets, generated=g)
    ha, hb = A.graph_hash(), B.graph_hash()
    assert ha == hb and ha[3] == p.num_edges
    A.close()
    from parallel_eda_b200 import pathfinder
    rep = pathfinder.run(B)
    res = B.result()
    res.success = int(rep.success)
    assert rep.success
    m = check_route.check_route_fast(p, res)
    assert m["overused"] == 0 and m["sinks"] == 3 * p.num_nets
    B.reset()
    assert pathfinder.run(B).success
    B.close()
