"""Pins the oracle's static timing analysis (oracle/pf_oracle.c: pf_oracle_sta) to the UNMODIFIED reference.

duo_w80 has two netlist clocks plus the virtual I/O clock: 3 domains, 7 analysed domain pairs, 2 DO_NOT_ANALYSE.
tests/golden/*.pftg.xz is the reference's own timing graph (tnode[] / tedge / levels / constraints, exported by
oracle/ref_build/harness.cxx) and *.pfsta.xz every (net_delay in, timing_criticality out, critical path delay) of the
do_timing_analysis calls the reference made while routing that fixture timing-driven (tests/golden/make_golden.sh).
The restatement must reproduce every criticality and critical path delay BIT FOR BIT."""
import ctypes as C
import os

import numpy as np
import pytest

from parallel_eda_b200 import pfio

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class _TG(C.Structure):
    _fields_ = [("num_tnodes", C.c_int32), ("num_tedges", C.c_int32), ("edge_ptr", C.c_void_p), ("edge_to", C.c_void_p),
                ("edge_Tdel", C.c_void_p), ("type", C.c_void_p), ("clock_domain", C.c_void_p), ("clock_delay", C.c_void_p),
                ("num_levels", C.c_int32), ("level_ptr", C.c_void_p), ("level_nodes", C.c_void_p), ("num_domains", C.c_int32),
                ("constraint", C.c_void_p), ("num_nets", C.c_int32), ("net_driver", C.c_void_p),
                ("num_overrides", C.c_int32), ("override_domain", C.c_void_p), ("override_tnode", C.c_void_p), ("override_constraint", C.c_void_p)]


def c_timing_graph(g: pfio.TimingGraph):
    keep = [np.ascontiguousarray(a) for a in (g.edge_ptr, g.edge_to, g.edge_Tdel, g.type, g.clock_domain, g.clock_delay,
                                               g.level_ptr, g.level_nodes, g.constraint, g.net_driver,
                                               g.override_domain, g.override_tnode, g.override_constraint)]
    p = [a.ctypes.data if a.size else None for a in keep]
    t = _TG(g.num_tnodes, len(g.edge_to), p[0], p[1], p[2], p[3], p[4], p[5], g.num_levels, p[6], p[7], int(g.constraint.shape[0]),
            p[8], len(g.net_driver), p[9], len(g.override_tnode), p[10], p[11], p[12])
    return t, keep


@pytest.mark.parametrize("name", ["toy_w64", "hub_w90", "mid_w200", "duo_w80", "het_w70"])
def test_oracle_sta_reproduces_reference_bit_for_bit(name, oracle_lib):
    lib = C.CDLL(oracle_lib)
    lib.pf_oracle_sta.argtypes = [C.POINTER(_TG), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    lib.pf_timing_graph_check.argtypes = [C.POINTER(_TG), C.c_void_p, C.c_char_p, C.c_int]
    p = pfio.read_problem(os.path.join(G, name + ".pfp.xz"))
    g = pfio.read_timing_graph(os.path.join(G, name + ".pftg.xz"))
    v = pfio.read_sta_vectors(os.path.join(G, name + ".pfsta.xz"))
    tg, keep = c_timing_graph(g)
    net_ptr = np.ascontiguousarray(p.net_ptr, dtype=np.int32)
    msg = C.create_string_buffer(256)
    assert lib.pf_timing_graph_check(C.byref(tg), net_ptr.ctypes.data, msg, 256) == 0, msg.value
    assert v.net_delay.shape[1] == p.num_terminals and len(g.net_driver) == p.num_nets
    calls = range(v.net_delay.shape[0]) if name != "mid_w200" else (0, 7, v.net_delay.shape[0] - 1)
    for k in calls:
        crit = np.zeros(p.num_terminals, np.float32)
        cpd = C.c_float(0)
        d = np.ascontiguousarray(v.net_delay[k])
        assert lib.pf_oracle_sta(C.byref(tg), net_ptr.ctypes.data, d.ctypes.data, crit.ctypes.data, C.byref(cpd)) == 0
        assert np.array_equal(crit.view(np.uint32), v.crit[k].view(np.uint32)), "call %d: %d criticalities differ" % (
            k, int((crit.view(np.uint32) != v.crit[k].view(np.uint32)).sum()))
        assert np.float32(cpd.value).view(np.uint32) == v.cpd[k].view(np.uint32)
    # and the analysis is what the golden routing's per-iteration criticalities came from
    gold = pfio.read_result(os.path.join(G, name + ".pfr.xz"))
    routed = np.repeat(p.net_is_global == 0, np.diff(p.net_ptr))
    assert np.array_equal(gold.iter_crit[1][routed], v.crit[0][routed])


@pytest.mark.parametrize("name", ["toy_w64", "hub_w90", "mid_w200", "duo_w80", "het_w70"])
def test_device_sta_code_on_the_emulator_is_bit_identical(name, emu_lib):
    """The device analysis (pf_sta_device.cuh behind pf_sta_analyze), compiled for the CPU emulator backend:
    level-synchronous pull over in-edges instead of the reference's push along out-edges — same floats."""
    from parallel_eda_b200 import router
    p = pfio.read_problem(os.path.join(G, name + ".pfp.xz"))
    g = pfio.read_timing_graph(os.path.join(G, name + ".pftg.xz"))
    v = pfio.read_sta_vectors(os.path.join(G, name + ".pfsta.xz"))
    s = router.Sta(g, p, router.default_config(router.load_library(emu_lib)), lib_path=emu_lib)
    for k in ((0, v.net_delay.shape[0] - 1) if name == "mid_w200" else range(v.net_delay.shape[0])):
        crit, cpd = s.analyze(v.net_delay[k])
        assert np.array_equal(crit.view(np.uint32), v.crit[k].view(np.uint32))
        assert np.float32(cpd).view(np.uint32) == v.cpd[k].view(np.uint32)
    s.close()


def test_route_with_device_sta_on_the_emulator(emu_lib):
    """pf_try_timing_driven_route_sta: the router with the analysis in the loop (no host callback)."""
    from parallel_eda_b200 import router, check_route
    p = pfio.read_problem(os.path.join(G, "toy_w64.pfp.xz"))
    p.opts["timing_analysis_enabled"] = 1
    p.opts["max_router_iterations"] = 150
    g = pfio.read_timing_graph(os.path.join(G, "toy_w64.pftg.xz"))
    gold = pfio.read_result(os.path.join(G, "toy_w64.pfr.xz"))
    r = router.try_timing_driven_route(p, router.default_config(router.load_library(emu_lib), num_slots=8, big_slots=1), lib_path=emu_lib,
                                       timing_graph=g)
    assert r.success == 1
    check_route.check_route(p, r)
    cpd = float(r.iter_stats["crit_path_delay"][-2])            # the last analysis ran before the final iteration
    ref = float(gold.iter_stats["crit_path_delay"][-2])
    assert abs(cpd - ref) <= 0.10 * ref and r.total_wirelength <= 1.12 * gold.total_wirelength


def test_step_loop_with_device_sta_on_the_emulator(emu_lib):
    """pathfinder.route(dsta=...): the analysis writes the router's criticality vector in place between iterations."""
    from parallel_eda_b200 import router, pathfinder, check_route
    p = pfio.read_problem(os.path.join(G, "toy_w64.pfp.xz"))
    p.opts["timing_analysis_enabled"] = 1
    p.opts["max_router_iterations"] = 150
    g = pfio.read_timing_graph(os.path.join(G, "toy_w64.pftg.xz"))
    cfg = router.default_config(router.load_library(emu_lib), num_slots=8, big_slots=1)
    R = router.Router(p, cfg, lib_path=emu_lib)
    S = router.Sta(g, p, cfg, lib_path=emu_lib)
    rep = pathfinder.route(R, dsta=S)
    assert rep.success
    res = R.result()
    res.success = 1
    check_route.check_route(p, res)
    # same routing as the C loop with the device analysis (both deterministic on the emulator)
    r2 = router.try_timing_driven_route(p, cfg, lib_path=emu_lib, timing_graph=g)
    assert (res.serial_num, res.total_wirelength) == (r2.serial_num, r2.total_wirelength) and rep.iterations == r2.iterations
    S.close(); R.close()


def test_device_sta_equals_oracle_on_random_delays(oracle_lib, emu_lib):
    """Beyond the reference's own vectors: random net delays (0 … 3x the golden ones, some exactly 0) on the
    three-domain fixture — the device code must equal the oracle (itself pinned to the reference) bit for bit."""
    from parallel_eda_b200 import router
    lib = C.CDLL(oracle_lib)
    lib.pf_oracle_sta.argtypes = [C.POINTER(_TG), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    p = pfio.read_problem(os.path.join(G, "duo_w80.pfp.xz"))
    g = pfio.read_timing_graph(os.path.join(G, "duo_w80.pftg.xz"))
    v = pfio.read_sta_vectors(os.path.join(G, "duo_w80.pfsta.xz"))
    tg, keep = c_timing_graph(g)
    net_ptr = np.ascontiguousarray(p.net_ptr, dtype=np.int32)
    s = router.Sta(g, p, router.default_config(router.load_library(emu_lib)), lib_path=emu_lib)
    rng = np.random.default_rng(3)
    for k in range(6):
        d = (v.net_delay[k % v.net_delay.shape[0]] * rng.uniform(0.0, 3.0, p.num_terminals)).astype(np.float32)
        d[rng.random(p.num_terminals) < 0.05] = 0.0
        want = np.zeros(p.num_terminals, np.float32); cpd = C.c_float(0)
        assert lib.pf_oracle_sta(C.byref(tg), net_ptr.ctypes.data, d.ctypes.data, want.ctypes.data, C.byref(cpd)) == 0
        got, got_cpd = s.analyze(d)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert np.float32(got_cpd).view(np.uint32) == np.float32(cpd.value).view(np.uint32)
    s.close()


FINAL = ["toy_w64", "het_w70", "duo_w80"]      # duo: 3 clock domains, 7 analysed pairs — the least slack over the pairs is kept


def _final_golden(name):
    v = pfio.read_sta_vectors(os.path.join(G, name + "_final.pfsta.xz"))
    s = pfio.read_sta_vectors(os.path.join(G, name + "_final.pfsta.slack.xz"))
    assert v.net_delay.shape[0] == 1 and np.array_equal(v.net_delay, s.net_delay)
    return v.net_delay[0], s.crit[0], v.crit[0], v.cpd[0]


@pytest.mark.parametrize("name", FINAL)
def test_final_analysis_oracle_and_device_code_equal_the_reference(name, oracle_lib, emu_lib):
    """The analysis routing_stats runs on the finished routing (base/stats.c:155-164: do_timing_analysis with
    is_final_analysis = TRUE — real required times, slacks kept): the reference's own slacks, criticalities and critical path
    delay of that call (hook PF_DUMP_STA_FINAL, tests/golden/make_golden.sh), bit for bit from the oracle restatement and from
    the device code on the emulator."""
    from parallel_eda_b200 import router
    p = pfio.read_problem(os.path.join(G, name + ".pfp.xz"))
    g = pfio.read_timing_graph(os.path.join(G, name + ".pftg.xz"))
    delay, slack_ref, crit_ref, cpd_ref = _final_golden(name)
    assert (slack_ref < 1e29).any() and (slack_ref[np.asarray(p.net_ptr[:-1])] > 1e29).all()      # driver slots stay HUGE_POSITIVE_FLOAT
    lib = C.CDLL(oracle_lib)
    lib.pf_oracle_sta_final.argtypes = [C.POINTER(_TG), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    tg, keep = c_timing_graph(g)
    net_ptr = np.ascontiguousarray(p.net_ptr, dtype=np.int32)
    slack = np.zeros(p.num_terminals, np.float32); crit = np.zeros(p.num_terminals, np.float32); cpd = C.c_float(0)
    d = np.ascontiguousarray(delay)
    assert lib.pf_oracle_sta_final(C.byref(tg), net_ptr.ctypes.data, d.ctypes.data, slack.ctypes.data, crit.ctypes.data, C.byref(cpd)) == 0
    assert np.array_equal(slack.view(np.uint32), slack_ref.view(np.uint32)), int((slack.view(np.uint32) != slack_ref.view(np.uint32)).sum())
    assert np.array_equal(crit.view(np.uint32), crit_ref.view(np.uint32))
    assert np.float32(cpd.value).view(np.uint32) == cpd_ref.view(np.uint32)
    s = router.Sta(g, p, router.default_config(router.load_library(emu_lib)), lib_path=emu_lib)
    slack_d, crit_d, cpd_d = s.analyze_final(delay)
    assert np.array_equal(slack_d.view(np.uint32), slack_ref.view(np.uint32))
    assert np.array_equal(crit_d.view(np.uint32), crit_ref.view(np.uint32))
    assert np.float32(cpd_d).view(np.uint32) == cpd_ref.view(np.uint32)
    # the relaxed analysis right after it is unaffected by the final one (flags reset), and differs where slacks were negative
    crit_relaxed, _ = s.analyze(delay)
    v = pfio.read_sta_vectors(os.path.join(G, name + ".pfsta.xz"))
    c0, _ = s.analyze(v.net_delay[0])
    assert np.array_equal(c0.view(np.uint32), v.crit[0].view(np.uint32))
    assert crit_relaxed.max() <= 1.0 + 1e-6
    s.close()


def test_clock_to_flipflop_override_constraints(oracle_lib, emu_lib, tmp_path):
    """An SDC file with override constraints from a clock to single flip-flops / pads (set_max_delay, set_false_path: the
    reference's g_sdc->cf_constraints, honoured at the sinks of the backward sweep, timing/path_delay.c:2753-2768): the exporter
    resolves them to (source domain, sink tnode) pairs in the timing graph (tests/golden/duo_ovr.sdc on the duo circuit: 7
    pairs).  Every analysis the reference made while routing with that SDC file (19 calls) and its analysis of the finished
    routing, bit for bit from the oracle and from the device code on the emulator — and NOT reproduced when the overrides are
    dropped, so the fixture does exercise them."""
    import dataclasses
    from parallel_eda_b200 import router
    p = pfio.read_problem(os.path.join(G, "duo_w80.pfp.xz"))          # the SDC file does not change the routing problem
    g = pfio.read_timing_graph(os.path.join(G, "duo_ovr_w80.pftg.xz"))
    v = pfio.read_sta_vectors(os.path.join(G, "duo_ovr_w80.pfsta.xz"))
    assert len(g.override_tnode) == 7 and (g.override_constraint < 0).any() and (g.override_constraint > 0).any()
    lib = C.CDLL(oracle_lib)
    lib.pf_oracle_sta.argtypes = [C.POINTER(_TG), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    lib.pf_oracle_sta_final.argtypes = [C.POINTER(_TG), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    lib.pf_timing_graph_check.argtypes = [C.POINTER(_TG), C.c_void_p, C.c_char_p, C.c_int]
    tg, keep = c_timing_graph(g)
    net_ptr = np.ascontiguousarray(p.net_ptr, dtype=np.int32)
    msg = C.create_string_buffer(256)
    assert lib.pf_timing_graph_check(C.byref(tg), net_ptr.ctypes.data, msg, 256) == 0, msg.value
    s = router.Sta(g, p, router.default_config(router.load_library(emu_lib)), lib_path=emu_lib)
    for k in range(v.net_delay.shape[0]):
        crit = np.zeros(p.num_terminals, np.float32); cpd = C.c_float(0)
        d = np.ascontiguousarray(v.net_delay[k])
        assert lib.pf_oracle_sta(C.byref(tg), net_ptr.ctypes.data, d.ctypes.data, crit.ctypes.data, C.byref(cpd)) == 0
        assert np.array_equal(crit.view(np.uint32), v.crit[k].view(np.uint32)), k
        assert np.float32(cpd.value).view(np.uint32) == v.cpd[k].view(np.uint32)
        c2, cpd2 = s.analyze(v.net_delay[k])
        assert np.array_equal(c2.view(np.uint32), v.crit[k].view(np.uint32)), k
        assert np.float32(cpd2).view(np.uint32) == v.cpd[k].view(np.uint32)
    delay, slack_ref, crit_ref, cpd_ref = _final_golden("duo_ovr_w80")
    slack_d, crit_d, cpd_d = s.analyze_final(delay)
    assert np.array_equal(slack_d.view(np.uint32), slack_ref.view(np.uint32)) and np.array_equal(crit_d.view(np.uint32), crit_ref.view(np.uint32))
    assert np.float32(cpd_d).view(np.uint32) == cpd_ref.view(np.uint32)
    s.close()
    # without the overrides the same inputs give other criticalities
    plain = dataclasses.replace(g, override_domain=g.override_domain[:0], override_tnode=g.override_tnode[:0], override_constraint=g.override_constraint[:0])
    tg0, keep0 = c_timing_graph(plain)
    crit = np.zeros(p.num_terminals, np.float32); cpd = C.c_float(0)
    d = np.ascontiguousarray(v.net_delay[0])
    assert lib.pf_oracle_sta(C.byref(tg0), net_ptr.ctypes.data, d.ctypes.data, crit.ctypes.data, C.byref(cpd)) == 0
    assert not np.array_equal(crit.view(np.uint32), v.crit[0].view(np.uint32))
    # container round trip (the override arrays ride at the end of the PFTIMG01 file; older files simply have none) and the checker
    out = str(tmp_path / "g.pftg")
    pfio.write_timing_graph(out, g)
    g2 = pfio.read_timing_graph(out)
    assert np.array_equal(g2.override_tnode, g.override_tnode) and np.array_equal(g2.override_constraint, g.override_constraint)
    bad = dataclasses.replace(g, override_tnode=g.override_tnode[::-1].copy())
    tgb, keepb = c_timing_graph(bad)
    assert lib.pf_timing_graph_check(C.byref(tgb), net_ptr.ctypes.data, msg, 256) != 0 and b"override" in msg.value
