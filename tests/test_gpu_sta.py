"""GPU tests (-m gpu) of the device static timing analysis: bit-exact criticalities against the
golden vectors of the UNMODIFIED reference's do_timing_analysis, and the router with the analysis in the loop."""
import os
import time

import numpy as np
import pytest

from parallel_eda_b200 import check_route, pfio, router
import parity_bar

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["toy_w64", "hub_w90", "mid_w200", "duo_w80"])
def test_device_sta_bit_identical_to_reference(name):
    p = pfio.read_problem(os.path.join(G, name + ".pfp.xz"))
    g = pfio.read_timing_graph(os.path.join(G, name + ".pftg.xz"))
    v = pfio.read_sta_vectors(os.path.join(G, name + ".pfsta.xz"))
    s = router.Sta(g, p)
    for k in range(v.net_delay.shape[0]):
        crit, cpd = s.analyze(v.net_delay[k])
        assert np.array_equal(crit.view(np.uint32), v.crit[k].view(np.uint32)), "call %d" % k
        assert np.float32(cpd).view(np.uint32) == v.cpd[k].view(np.uint32)
    s.close()


@pytest.mark.parametrize("name", ["toy_w64", "het_w70", "duo_w80"])
def test_final_analysis_bit_identical_to_reference(name):
    """pf_sta_analyze_final on the B200: the slacks, criticalities and critical path delay of the reference's analysis of the
    finished routing (routing_stats, base/stats.c:155-164; goldens from the hook PF_DUMP_STA_FINAL) bit for bit, and the relaxed
    analysis of the router loop unchanged right after it."""
    p = pfio.read_problem(os.path.join(G, name + ".pfp.xz"))
    g = pfio.read_timing_graph(os.path.join(G, name + ".pftg.xz"))
    v = pfio.read_sta_vectors(os.path.join(G, name + "_final.pfsta.xz"))
    sl = pfio.read_sta_vectors(os.path.join(G, name + "_final.pfsta.slack.xz"))
    s = router.Sta(g, p)
    slack, crit, cpd = s.analyze_final(v.net_delay[0])
    assert np.array_equal(slack.view(np.uint32), sl.crit[0].view(np.uint32))
    assert np.array_equal(crit.view(np.uint32), v.crit[0].view(np.uint32))
    assert np.float32(cpd).view(np.uint32) == v.cpd[0].view(np.uint32)
    loop = pfio.read_sta_vectors(os.path.join(G, name + ".pfsta.xz"))
    c0, _ = s.analyze(loop.net_delay[0])
    assert np.array_equal(c0.view(np.uint32), loop.crit[0].view(np.uint32))
    s.close()


def test_override_constraints_bit_identical_to_reference():
    """Clock-to-flipflop override constraints of an SDC file (tests/golden/duo_ovr.sdc; g_sdc->cf_constraints, honoured at the
    sinks of the backward sweep, timing/path_delay.c:2753-2768) on the B200: all 19 analyses of the reference's routing run with
    that file and its analysis of the finished routing, bit for bit."""
    p = pfio.read_problem(os.path.join(G, "duo_w80.pfp.xz"))
    g = pfio.read_timing_graph(os.path.join(G, "duo_ovr_w80.pftg.xz"))
    v = pfio.read_sta_vectors(os.path.join(G, "duo_ovr_w80.pfsta.xz"))
    assert len(g.override_tnode) == 7
    s = router.Sta(g, p)
    for k in range(v.net_delay.shape[0]):
        crit, cpd = s.analyze(v.net_delay[k])
        assert np.array_equal(crit.view(np.uint32), v.crit[k].view(np.uint32)), "call %d" % k
        assert np.float32(cpd).view(np.uint32) == v.cpd[k].view(np.uint32)
    f = pfio.read_sta_vectors(os.path.join(G, "duo_ovr_w80_final.pfsta.xz"))
    sl = pfio.read_sta_vectors(os.path.join(G, "duo_ovr_w80_final.pfsta.slack.xz"))
    slack, crit, cpd = s.analyze_final(f.net_delay[0])
    assert np.array_equal(slack.view(np.uint32), sl.crit[0].view(np.uint32)) and np.array_equal(crit.view(np.uint32), f.crit[0].view(np.uint32))
    assert np.float32(cpd).view(np.uint32) == f.cpd[0].view(np.uint32)
    s.close()


@pytest.mark.parametrize("name", ["mid_w200", "hub_w90", "duo_w80"])
def test_route_with_device_sta(name):
    p = pfio.read_problem(os.path.join(G, name + ".pfp.xz"))
    p.opts["timing_analysis_enabled"] = 1
    g = pfio.read_timing_graph(os.path.join(G, name + ".pftg.xz"))
    gold = pfio.read_result(os.path.join(G, name + ".pfr.xz"))
    # critical path delay of the analysis of the FINAL routing against the reference's last analysis (same circuit, same
    # placement; the reference analyses before its last iteration, pf_route_run also after it)
    ref = float(gold.iter_stats["crit_path_delay"][-2])
    r = parity_bar.check_runs("closed_loop_device_sta", name, lambda: router.try_timing_driven_route(p, timing_graph=g), gold,
                              weighted=lambda r: (float(r.iter_stats["crit_path_delay"][-1]), ref))
    m = check_route.check_route(p, r)
    assert m["overused"] == 0
