"""check_route on the device (pf_check_route; reference route/check_route.c:27-155), here through the CPU emulator
backend: the reference's own golden routings must pass, every seeded violation must be caught — and the verdicts
must agree with the independent Python checker (parallel_eda_b200/check_route.py)."""
import copy
import os

import numpy as np
import pytest

from parallel_eda_b200 import check_route, pfio, router

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def toy(emu_lib):
    p = pfio.read_problem(os.path.join(G, "toy_w64.pfp.xz"))
    g = pfio.read_result(os.path.join(G, "toy_w64.pfr.xz"))
    R = router.Router(p, router.default_config(router.load_library(emu_lib), num_slots=1, big_slots=1), lib_path=emu_lib)
    yield p, g, R
    R.close()


def test_reference_golden_routings_pass(toy, emu_lib):
    p, g, R = toy
    rep = R.check_route(g)
    assert rep["ok"] == 1 and rep["bad_nets"] == 0 and rep["overused_nodes"] == 0 and rep["occupancy_mismatch"] == 0
    assert rep["wirelength"] == g.total_wirelength == check_route.check_route(p, g)["wirelength"]
    assert rep["reserved_opins"] == int(np.asarray(p.opin_group_count).sum())
    for name in ("toy_w64_nt", "toy_w64_bf"):
        assert R.check_route(pfio.read_result(os.path.join(G, name + ".pfr.xz")))["ok"] == 1


def _net_with(p, g, min_len=6):
    for i in p.routed_nets():
        if g.trace_ptr[i + 1] - g.trace_ptr[i] >= min_len:
            return int(i), int(g.trace_ptr[i]), int(g.trace_ptr[i + 1])
    raise AssertionError


@pytest.mark.parametrize("fault,code", [("source", 2), ("edge", 5), ("switch", 5), ("sink_switch", 6), ("drop_sink", 3), ("wrong_sink", 7), ("range", 8)])
def test_seeded_violations_are_caught(toy, fault, code):
    p, g, R = toy
    bad = copy.deepcopy(g)
    i, a, b = _net_with(p, g)
    if fault == "source":
        bad.trace_node[a] = bad.trace_node[a + 1]
    elif fault == "edge":
        bad.trace_node[a + 2] = bad.trace_node[a]                  # not a neighbour of element a+1
    elif fault == "switch":
        bad.trace_switch[a + 1] = (int(bad.trace_switch[a + 1]) + 1) % len(p.switches)
    elif fault == "sink_switch":
        k = a + int(np.nonzero(p.type[g.trace_node[a:b]] == pfio.SINK)[0][0])
        bad.trace_switch[k] = 0
    elif fault == "drop_sink":
        keep = np.ones(len(bad.trace_node), bool); keep[b - 1] = False
        bad.trace_node, bad.trace_switch = bad.trace_node[keep], bad.trace_switch[keep]
        bad.trace_ptr = bad.trace_ptr.copy(); bad.trace_ptr[i + 1:] -= 1
    elif fault == "wrong_sink":
        other = [int(s) for s in p.net_terminals[p.net_ptr[i - 1] + 1:p.net_ptr[i]]] if i > 0 else []
        sinks = np.nonzero(p.type == pfio.SINK)[0]
        mine = set(int(s) for s in p.net_terminals[p.net_ptr[i] + 1:p.net_ptr[i + 1]])
        bad.trace_node[b - 1] = next(int(s) for s in sinks if int(s) not in mine)
    elif fault == "range":
        bad.trace_node[a + 1] = p.num_nodes + 5
    rep = R.check_route(bad)
    assert rep["ok"] == 0 and rep["bad_nets"] >= 1 and rep["first_bad_net"] == i
    assert rep["first_bad_code"] in ((code, 5) if fault in ("wrong_sink",) else (code,))
    if fault != "range":
        with pytest.raises(check_route.RouteCheckError):
            check_route.check_route(p, bad, check_delays=False)


def test_occupancy_tampering_is_caught(toy):
    p, g, R = toy
    bad = copy.deepcopy(g)
    v = int(np.nonzero(p.type == pfio.CHANX)[0][0])
    bad.occ = bad.occ.copy(); bad.occ[v] += 1
    rep = R.check_route(bad)
    assert rep["ok"] == 0 and rep["bad_nets"] == 0 and rep["occupancy_mismatch"] >= 1


def test_random_mutations_agree_with_the_python_checker(toy):
    """Differential test: 40 seeded single-element mutations of the reference's golden routing (node replaced by a
    neighbour-ish id, switch changed, element removed); pf_check_route and the independent Python checker must
    give the same accept / reject verdict on every one."""
    p, g, R = toy
    rng = np.random.default_rng(7)
    routed = p.routed_nets()
    rejected = 0
    for _ in range(40):
        bad = copy.deepcopy(g)
        i = int(rng.choice(routed))
        a, b = int(g.trace_ptr[i]), int(g.trace_ptr[i + 1])
        k = int(rng.integers(a, b))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            bad.trace_node[k] = int(np.clip(int(bad.trace_node[k]) + int(rng.integers(-3, 4)), 0, p.num_nodes - 1))
        elif kind == 1:
            bad.trace_switch[k] = int(rng.integers(-1, len(p.switches)))
        else:
            keep = np.ones(len(bad.trace_node), bool); keep[k] = False
            bad.trace_node, bad.trace_switch = bad.trace_node[keep], bad.trace_switch[keep]
            bad.trace_ptr = bad.trace_ptr.copy(); bad.trace_ptr[i + 1:] -= 1
        rep = R.check_route(bad)
        try:
            check_route.check_route(p, bad, check_delays=False)
            py_ok = True
        except check_route.RouteCheckError:
            py_ok = False
        assert bool(rep["ok"]) == py_ok, (i, k, kind, rep)
        rejected += not py_ok
    assert rejected >= 20          # most mutations break the routing; the rest are no-ops (same value drawn)


def test_inconsistent_trace_offsets_are_refused_before_the_kernel_runs(toy):
    """pf_check_route takes ANY result (a parsed .route file, another router's dump): offsets that are not monotone from 0
    would send the kernel outside trace_node — PF_EINVAL instead."""
    p, g, R = toy
    for edit in ("swap", "start"):
        bad = copy.deepcopy(g)
        bad.trace_ptr = bad.trace_ptr.copy()
        if edit == "swap":
            i = int(p.routed_nets()[5])
            bad.trace_ptr[i + 1] = bad.trace_ptr[-1] + 1000         # beyond the arrays, later offsets come back down
        else:
            bad.trace_ptr[0] = 3
        with pytest.raises(router.RouterError, match="trace_ptr") as e:
            R.check_route(bad)
        assert e.value.code == -4
    assert R.check_route(g)["ok"] == 1                               # the router handle is still usable


@pytest.mark.parametrize("name", ["het_w70", "heq_w70", "mix_w70", "hub_w90", "duo_w80"])
def test_reference_routings_of_every_fixture_pass(name, emu_lib):
    """The reference's own check_route accepted these routings inside the flow that produced the goldens; the device checker
    and the Python checker must accept them too — tall blocks (het), two pins of a net on one SINK (heq: the SINK is matched
    once per pin), two wire types (mix), an 84-sink net (hub), two clock domains (duo) — and agree on the wirelength."""
    p = pfio.read_problem(os.path.join(G, name + ".pfp.xz"))
    g = pfio.read_result(os.path.join(G, name + ".pfr.xz"))
    R = router.Router(p, router.default_config(router.load_library(emu_lib), num_slots=1, big_slots=1), lib_path=emu_lib)
    try:
        rep = R.check_route(g)
    finally:
        R.close()
    assert rep["ok"] == 1 and rep["bad_nets"] == 0 and rep["overused_nodes"] == 0 and rep["occupancy_mismatch"] == 0
    assert rep["wirelength"] == g.total_wirelength == check_route.check_route(p, g, check_delays=True)["wirelength"]
