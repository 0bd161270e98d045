"""The independent legality / delay checker accepts the reference's own routings and rejects broken ones."""
import os

import numpy as np
import pytest

from parallel_eda_b200 import check_route, pfio

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("gold", ["toy_w64.pfr.xz", "toy_w64_nt.pfr.xz"])
def test_reference_routing_passes(gold):
    p = pfio.read_problem(os.path.join(G, "toy_w64.pfp.xz"))
    r = pfio.read_result(os.path.join(G, gold))
    m = check_route.check_route(p, r)
    assert m["overused"] == 0 and m["wirelength"] == r.total_wirelength


def test_corruptions_are_caught():
    p = pfio.read_problem(os.path.join(G, "toy_w64.pfp.xz"))
    r = pfio.read_result(os.path.join(G, "toy_w64_nt.pfr.xz"))
    inet = int(p.routed_nets()[3])
    a = int(r.trace_ptr[inet])
    bad = pfio.Result(**{**r.__dict__, "trace_node": r.trace_node.copy()})
    bad.trace_node[a + 1] = bad.trace_node[a + 1] + 1          # break adjacency
    with pytest.raises(check_route.RouteCheckError):
        check_route.check_route(p, bad)
    bad2 = pfio.Result(**{**r.__dict__, "net_delay": r.net_delay * 1.01})   # 1 % off in every delay
    with pytest.raises(check_route.RouteCheckError):
        check_route.check_route(p, bad2)
    bad3 = pfio.Result(**{**r.__dict__, "occ": r.occ.copy()})
    bad3.occ[r.trace_node[a + 2]] += 1                          # occupancy not explained by the traces
    with pytest.raises(check_route.RouteCheckError):
        check_route.check_route(p, bad3)
