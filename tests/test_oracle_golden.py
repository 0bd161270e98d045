"""Pins the CPU oracle (oracle/pf_oracle.c) to the UNMODIFIED reference router.

The golden .pfr files were written by oracle/_ref/vpr_ref — the reference's own
try_timing_driven_route (vpr/SRC/route/route_timing.c:85) compiled from /root/reference — see
tests/golden/make_golden.sh.  The oracle must reproduce them bit for bit: identical s_trace lists
of every net, iteration count, "magic cookie" (route_common.c:224-254), per-sink Elmore delays,
final occupancy and per-iteration overuse, in timing-driven mode (criticalities replayed from the
reference's own STA) and with timing analysis off.
"""
import os
import subprocess

import numpy as np
import pytest

from parallel_eda_b200 import pfio

CASES = [
    ("toy_w64", "toy_w64.pfr", True),        # 294 nets, 6x6, W=64, timing-driven: 21 iterations
    ("toy_w64", "toy_w64_nt.pfr", False),    # timing analysis off: 14 iterations
    ("mid_w200", "mid_w200_nt.pfr", False),  # 3852 nets, 21x21, W=200, timing off: 12 iterations
    ("mid_w200", "mid_w200.pfr", True),      # timing-driven: 21 iterations, 80,871 net routes
    ("hub_w90", "hub_w90.pfr", True),        # 865 nets, 10x10, W=90, one routed net with 84 sinks: exercises the
    ("hub_w90", "hub_w90_nt.pfr", False),    # high-fanout window of mark_node_expansion_by_bin (route_timing.c:867)
    ("duo_w80", "duo_w80.pfr", True),        # 494 nets, W=80, TWO netlist clocks (+ the virtual I/O clock): 21 iterations
    ("duo_w80", "duo_w80_nt.pfr", False),    # timing off: 11 iterations
    # heterogeneous fabric (tests/fixtures/k6_N10_het.xml): a column of height-2 hard multiplier blocks every 5 columns,
    # i.e. SOURCE / SINK / pin rr nodes that span two tiles, CLB columns interrupted, 441 nets, 8x8
    ("het_w70", "het_w70.pfr", True),        # timing-driven: 18 iterations
    ("het_w70", "het_w70_nt.pfr", False),    # timing off: 11 iterations
    # two wire types (tests/fixtures/k6_N10_mix.xml: 30 % length-1 + 70 % length-4): eight rr_indexed_data rows, the
    # lookahead mixes the inv_length / T_linear / C_load of a row and of its orthogonal row (route_timing.c:693-746)
    ("mix_w70", "mix_w70.pfr", True),        # 345 nets, 22 iterations
    ("mix_w60", "mix_w60.pfr", True),        # near the minimum width: 34 iterations
    # heq: 12 nets that connect TWICE to one SINK rr node (equivalent input pins of a hard block fed by one signal): mark_ends
    # counts the target twice (route_common.c:764-778) and the second search must not stop at the first arrival.  The reference
    # routes it and then fails its own DEBUG delay cross-check on exactly these nets (route_timing.c:246) — the golden was
    # dumped at the moment of success; the oracle reproduces the reference's delays here too, bit for bit
    ("heq_w70", "heq_w70.pfr", True),        # 442 nets, 25 iterations
    ("het_w60", "het_w60.pfr", True),        # one track too few: the reference gives up after max_router_iterations = 50
]                                            # with 1 overused node (success = 0); the oracle must fail the same way


@pytest.mark.parametrize("prob,gold,timing", CASES)
def test_oracle_reproduces_reference_bit_for_bit(prob, gold, timing, oracle_cli, unxz, tmp_path):
    out = str(tmp_path / "o.pfr")
    cmd = [oracle_cli, unxz(prob + ".pfp"), "--result", out]
    if timing:
        cmd += ["--crit", unxz(gold)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    g = pfio.read_result(unxz(gold))
    o = pfio.read_result(out)
    assert (o.success, o.iterations) == (g.success, g.iterations)
    assert o.serial_num == g.serial_num
    assert o.total_wirelength == g.total_wirelength
    assert np.array_equal(o.trace_ptr, g.trace_ptr)
    assert np.array_equal(o.trace_node, g.trace_node)
    assert np.array_equal(o.trace_switch, g.trace_switch)
    assert np.array_equal(o.net_delay.view(np.uint32), g.net_delay.view(np.uint32))   # bit-exact floats
    assert np.array_equal(o.occ, g.occ)
    assert list(o.iter_stats["overused_nodes"]) == list(g.iter_stats["overused_nodes"])


@pytest.mark.parametrize("name", ["toy_w64", "duo_w80", "hub_w90", "het_w70"])
def test_oracle_breadth_first_reproduces_reference_bit_for_bit(name, oracle_cli, unxz, tmp_path):
    """--router_algorithm breadth_first (reference route_breadth_first.c): the golden *_bf.pfr was written by the
    unmodified reference; *_bf.pfp is the problem it saw (same rr graph as the timing-driven fixture, but
    demand-only base costs — rr_graph_indexed_data.c — and VPR's breadth-first option defaults: first_iter_pres_fac
    0, acc_fac 0.2)."""
    p = pfio.read_problem(unxz(name + "_bf.pfp"))
    assert int(p.opts["router_algorithm"]) == 1 and int(p.opts["timing_analysis_enabled"]) == 0
    out = str(tmp_path / "o.pfr")
    r = subprocess.run([oracle_cli, unxz(name + "_bf.pfp"), "--result", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    g = pfio.read_result(unxz(name + "_bf.pfr"))
    o = pfio.read_result(out)
    assert (o.success, o.iterations, o.serial_num, o.total_wirelength) == (g.success, g.iterations, g.serial_num, g.total_wirelength)
    assert np.array_equal(o.trace_ptr, g.trace_ptr) and np.array_equal(o.trace_node, g.trace_node)
    assert np.array_equal(o.trace_switch, g.trace_switch) and np.array_equal(o.occ, g.occ)
    assert list(o.iter_stats["overused_nodes"]) == list(g.iter_stats["overused_nodes"])


def test_reference_binary_agrees_when_present(ref_bin, oracle_cli, unxz, tmp_path):
    """Where oracle/_ref was built, re-run the real reference on the flat problem (inject mode)."""
    out_r, out_o = str(tmp_path / "r.pfr"), str(tmp_path / "o.pfr")
    prob = unxz("toy_w64.pfp")
    subprocess.run([ref_bin, "inject", prob, "--result", out_r], check=True, capture_output=True)
    subprocess.run([oracle_cli, prob, "--result", out_o], check=True, capture_output=True)
    r, o = pfio.read_result(out_r), pfio.read_result(out_o)
    assert r.serial_num == o.serial_num and np.array_equal(r.trace_node, o.trace_node)
    assert np.array_equal(r.net_delay.view(np.uint32), o.net_delay.view(np.uint32))


@pytest.mark.parametrize("name", ["toy_w64", "duo_w80", "hub_w90", "mid_w200", "het_w70", "mix_w70", "heq_w70"])
def test_oracle_router_and_sta_in_closed_loop_reproduce_the_reference_run(name, oracle_cli, unxz, tmp_path):
    """No replay: the oracle router with the oracle's own static timing analysis between iterations (--timing-graph)
    must reproduce the reference's WHOLE timing-driven run — iteration count, every trace, the cookie, bit-exact sink
    delays and the criticalities of every iteration."""
    out = str(tmp_path / "o.pfr")
    r = subprocess.run([oracle_cli, unxz(name + ".pfp"), "--timing-graph", unxz(name + ".pftg"), "--result", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    p = pfio.read_problem(unxz(name + ".pfp"))
    g = pfio.read_result(unxz(name + ".pfr"))
    o = pfio.read_result(out)
    assert (o.success, o.iterations, o.serial_num, o.total_wirelength) == (g.success, g.iterations, g.serial_num, g.total_wirelength)
    assert np.array_equal(o.trace_ptr, g.trace_ptr) and np.array_equal(o.trace_node, g.trace_node) and np.array_equal(o.trace_switch, g.trace_switch)
    assert np.array_equal(o.net_delay.view(np.uint32), g.net_delay.view(np.uint32)) and np.array_equal(o.occ, g.occ)
    routed = np.repeat(p.net_is_global == 0, np.diff(p.net_ptr))
    assert np.array_equal(o.iter_crit.view(np.uint32)[:, routed], g.iter_crit.view(np.uint32)[:, routed])


def test_closed_loop_with_clock_to_flipflop_override_constraints(oracle_cli, unxz, tmp_path):
    """The duo circuit routed by the reference with an SDC file that overrides the constraint from a clock to single
    flip-flops (tests/golden/duo_ovr.sdc: set_max_delay / set_false_path, g_sdc->cf_constraints): other criticalities, hence
    another routing than duo_w80 (20 iterations instead of 21, another cookie).  The routing problem is the same file; the
    overrides travel in the timing graph — and the oracle router with its own analysis in the loop reproduces that WHOLE run."""
    out = str(tmp_path / "o.pfr")
    r = subprocess.run([oracle_cli, unxz("duo_w80.pfp"), "--timing-graph", unxz("duo_ovr_w80.pftg"), "--result", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    p = pfio.read_problem(unxz("duo_w80.pfp"))
    g = pfio.read_result(unxz("duo_ovr_w80.pfr"))
    plain = pfio.read_result(unxz("duo_w80.pfr"))
    o = pfio.read_result(out)
    assert (g.iterations, g.serial_num) != (plain.iterations, plain.serial_num)
    assert (o.success, o.iterations, o.serial_num, o.total_wirelength) == (g.success, g.iterations, g.serial_num, g.total_wirelength)
    assert np.array_equal(o.trace_node, g.trace_node) and np.array_equal(o.trace_switch, g.trace_switch)
    assert np.array_equal(o.net_delay.view(np.uint32), g.net_delay.view(np.uint32)) and np.array_equal(o.occ, g.occ)
    routed = np.repeat(p.net_is_global == 0, np.diff(p.net_ptr))
    assert np.array_equal(o.iter_crit.view(np.uint32)[:, routed], g.iter_crit.view(np.uint32)[:, routed])


def unbuffered_toy(unxz_or_path, timing):
    """toy_w64 with switch 0 (every wire-to-wire and OPIN-to-wire edge) turned into a PASS TRANSISTOR: buffered = 0,
    R = 400, Tdel = 10 ps.  All fixture architectures are unidirectional, whose mux switches are buffered, and the
    reference's bidirectional rr-graph builder overflows a 2-entry stack array as soon as a wire switch is unbuffered
    (rr_graph2.c:1436-1445: `switch_types[used]` with used == 2; ASan trace in DESIGN.md §5) — so the unbuffered branch
    of the Elmore code (update_unbuffered_ancestors_C_downstream, route_tree_timing.c:393-417, and the R_upstream /
    C_downstream bookkeeping of :216-390) is reached by editing the switch table of a flat problem and handing it to the
    unmodified reference ROUTER (inject mode), which is how the *_unbuf goldens were made."""
    p = pfio.read_problem(unxz_or_path)
    p.switches = p.switches.copy()
    p.switches["buffered"][0] = 0
    p.switches["R"][0] = 400.0
    p.switches["Tdel"][0] = 1e-11
    p.opts = p.opts.copy()
    p.opts["timing_analysis_enabled"] = 1 if timing else 0
    return p


@pytest.mark.parametrize("timing", [False, True])
def test_oracle_unbuffered_switches_bit_for_bit(timing, oracle_cli, unxz, tmp_path):
    prob, out = str(tmp_path / "u.pfp"), str(tmp_path / "o.pfr")
    pfio.write_problem(prob, unbuffered_toy(unxz("toy_w64.pfp"), timing))
    gold = unxz("toy_w64_unbuf_td.pfr" if timing else "toy_w64_unbuf_nt.pfr")
    cmd = [oracle_cli, prob, "--result", out] + (["--crit", unxz("toy_w64.pfr")] if timing else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    g, o = pfio.read_result(gold), pfio.read_result(out)
    assert (o.success, o.iterations, o.serial_num, o.total_wirelength) == (g.success, g.iterations, g.serial_num, g.total_wirelength)
    assert np.array_equal(o.trace_node, g.trace_node) and np.array_equal(o.trace_switch, g.trace_switch)
    assert np.array_equal(o.net_delay.view(np.uint32), g.net_delay.view(np.uint32)) and np.array_equal(o.occ, g.occ)
    # the edit changed the delays (so the branch really ran): same problem with buffered switches gives other numbers
    b = pfio.read_result(unxz("toy_w64.pfr" if timing else "toy_w64_nt.pfr"))
    assert not np.array_equal(b.net_delay, g.net_delay)


def test_reference_binary_agrees_on_unbuffered_switches_when_present(ref_bin, unxz, tmp_path):
    """Regenerates the timing-driven *_unbuf golden with the real reference (criticalities of toy_w64.pfr replayed)."""
    prob, out = str(tmp_path / "u.pfp"), str(tmp_path / "r.pfr")
    pfio.write_problem(prob, unbuffered_toy(unxz("toy_w64.pfp"), True))
    subprocess.run([ref_bin, "inject", prob, "--crit", unxz("toy_w64.pfr"), "--result", out], check=True, capture_output=True)
    g, r = pfio.read_result(unxz("toy_w64_unbuf_td.pfr")), pfio.read_result(out)
    assert r.serial_num == g.serial_num and np.array_equal(r.trace_node, g.trace_node)
    assert np.array_equal(r.net_delay.view(np.uint32), g.net_delay.view(np.uint32))


def test_parallel_cpu_router_one_thread_is_the_serial_oracle_and_many_threads_stay_legal(oracle_cli, unxz, tmp_path):
    """oracle/pf_oracle_par.c — the multi-threaded CPU baseline bench.py reports next to the serial reference (threads own
    their search state and share the occupancy arrays).  With one thread it IS the serial restatement: the reference's
    golden routing, bit for bit.  With four threads the schedule decides which net sees which occupancy, so the bar is the
    GPU's: a legal routing (independent check_route incl. from-scratch Elmore delays) and a wirelength close to the serial one
    on a generated fabric with sparse conflicts (tight circuit fixtures do not converge when every net is re-routed every
    iteration by several threads at once — DESIGN.md §4.5)."""
    from parallel_eda_b200 import check_route, router
    par = os.path.join(os.path.dirname(oracle_cli), "pf_oracle_par_cli")
    out = str(tmp_path / "p1.pfr")
    r = subprocess.run([par, unxz("toy_w64.pfp"), "--threads", "1", "--result", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    g, o = pfio.read_result(unxz("toy_w64_nt.pfr")), pfio.read_result(out)
    assert (o.success, o.iterations, o.serial_num, o.total_wirelength) == (g.success, g.iterations, g.serial_num, g.total_wirelength)
    assert np.array_equal(o.trace_node, g.trace_node) and np.array_equal(o.occ, g.occ)
    assert np.array_equal(o.net_delay.view(np.uint32), g.net_delay.view(np.uint32))

    p = router.generate_grid_problem(nx=40, ny=40, W=60, num_nets=2500, sinks_per_net=3, seed=8)
    pp, s_out, p_out = str(tmp_path / "g.pfp"), str(tmp_path / "s.pfr"), str(tmp_path / "p4.pfr")
    pfio.write_problem(pp, p)
    assert subprocess.run([oracle_cli, pp, "--result", s_out], capture_output=True).returncode == 0
    s = pfio.read_result(s_out)
    seen = []
    for attempt in range(3):       # the schedule is not deterministic (13-19 iterations over a dozen runs, serial: 11)
        r = subprocess.run([par, pp, "--threads", "4", "--result", p_out], capture_output=True, text=True)
        seen.append(r.stderr.strip().split("\n")[-1])
        if r.returncode != 0:
            continue
        q = pfio.read_result(p_out)
        assert check_route.check_route(p, q, check_delays=True)["overused"] == 0      # whatever converged must be legal
        if abs(q.total_wirelength - s.total_wirelength) <= 0.02 * s.total_wirelength:
            break
    else:
        raise AssertionError("no 4-thread run converged close to the serial routing: %s" % seen)
    # --congested-only: the device router's re-route policy on the CPU threads; with it even the tight circuit fixtures
    # converge under 8 threads (all-nets: 50 iterations are not enough on hub_w90)
    hp = pfio.read_problem(unxz("hub_w90.pfp"))
    hp.opts["timing_analysis_enabled"] = 0
    for attempt in range(3):
        r = subprocess.run([par, unxz("hub_w90.pfp"), "--threads", "8", "--congested-only", "--result", p_out], capture_output=True, text=True)
        if r.returncode == 0:
            break
    assert r.returncode == 0 and "policy=congested-only" in r.stderr, r.stderr[-800:]
    q = pfio.read_result(p_out)
    assert check_route.check_route(hp, q, check_delays=True)["overused"] == 0
    assert q.total_wirelength <= 1.08 * pfio.read_result(unxz("hub_w90_nt.pfr")).total_wirelength
