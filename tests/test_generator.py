"""The native grid generator (pf_gen.cpp): structural invariants, determinism, and that the reference-side
consumers accept what it produces."""
import os
import subprocess

import numpy as np
import pytest

from parallel_eda_b200 import check_route, pfio, router


@pytest.fixture(scope="module")
def small(cuda_lib):
    return router.generate_grid_problem(lib_path=cuda_lib, nx=12, ny=10, W=40, num_nets=400, sinks_per_net=3, window=5, seed=7)


def test_counts_match_the_reference_grid_convention(cuda_lib):
    # same grid/W as the reference's own dump of the 21x21, W=200 fixture: identical node census
    p = router.generate_grid_problem(lib_path=cuda_lib, nx=21, ny=21, W=200, num_nets=100)
    g = pfio.read_problem(os.path.join(os.path.dirname(__file__), "golden", "mid_w200.pfp.xz"))
    for t in range(6):
        assert int((p.type == t).sum()) == int((g.type == t).sum())
    assert p.num_nodes == g.num_nodes == 84615
    dg, dr = np.diff(p.row_ptr), np.diff(g.row_ptr)
    for t in (pfio.CHANX, pfio.CHANY):           # same degree profile (connection + switch boxes)
        assert abs(dg[p.type == t].mean() - dr[g.type == t].mean()) < 2.0
    assert (dg[p.type == pfio.OPIN] == 20).all() and (dg[p.type == pfio.IPIN] == 1).all()


def test_structure(small):
    p = small
    assert p.row_ptr[0] == 0 and p.row_ptr[-1] == p.num_edges and (np.diff(p.row_ptr) >= 0).all()
    assert p.edge_to.min() >= 0 and p.edge_to.max() < p.num_nodes
    src = np.repeat(np.arange(p.num_nodes), np.diff(p.row_ptr))
    ts, tt = p.type[src], p.type[p.edge_to]
    legal = {(pfio.SOURCE, pfio.OPIN), (pfio.OPIN, pfio.CHANX), (pfio.OPIN, pfio.CHANY), (pfio.CHANX, pfio.CHANX),
             (pfio.CHANX, pfio.CHANY), (pfio.CHANY, pfio.CHANX), (pfio.CHANY, pfio.CHANY), (pfio.CHANX, pfio.IPIN),
             (pfio.CHANY, pfio.IPIN), (pfio.IPIN, pfio.SINK)}
    assert set(zip(ts.tolist(), tt.tolist())) <= legal
    # unidirectional wires are driven only at their start: a wire-to-wire edge lands where the target begins
    ww = np.isin(ts, (pfio.CHANX, pfio.CHANY)) & np.isin(tt, (pfio.CHANX, pfio.CHANY))
    a, b = src[ww], p.edge_to[ww]
    touch = (p.xlow[a] <= p.xhigh[b] + 1) & (p.xlow[b] <= p.xhigh[a] + 1) & (p.ylow[a] <= p.yhigh[b] + 1) & (p.ylow[b] <= p.yhigh[a] + 1)
    assert touch.all()
    # nets: SOURCE + 3 distinct SINK tiles, every CLB output drives at most one net
    assert (np.diff(p.net_ptr) == 4).all()
    srcs = p.net_terminals[p.net_ptr[:-1]]
    assert (p.type[srcs] == pfio.SOURCE).all() and len(np.unique(srcs)) == p.num_nets
    sinks = p.net_terminals.reshape(-1, 4)[:, 1:]
    assert (p.type[sinks] == pfio.SINK).all() and all(len(set(r)) == 3 for r in sinks.tolist())
    assert np.bincount(sinks.ravel(), minlength=p.num_nodes).max() <= 40


def test_deterministic(cuda_lib, small):
    q = router.generate_grid_problem(lib_path=cuda_lib, nx=12, ny=10, W=40, num_nets=400, sinks_per_net=3, window=5, seed=7)
    assert np.array_equal(q.edge_to, small.edge_to) and np.array_equal(q.net_terminals, small.net_terminals)


def test_oracle_routes_it_and_reference_agrees(small, oracle_cli, tmp_path):
    prob, out = str(tmp_path / "g.pfp"), str(tmp_path / "g.pfr")
    pfio.write_problem(prob, small)
    subprocess.run([oracle_cli, prob, "--result", out], check=True, capture_output=True)
    r = pfio.read_result(out)
    assert r.success == 1
    check_route.check_route(small, r)
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "vpr_ref")
    if os.path.exists(ref):       # the real reference on the same generated problem: identical routing
        out2 = str(tmp_path / "r.pfr")
        subprocess.run([ref, "inject", prob, "--result", out2], check=True, capture_output=True)
        assert pfio.read_result(out2).serial_num == r.serial_num
