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


@pytest.mark.parametrize("kw", [dict(nx=6, ny=5, W=20, num_nets=40, window=3), dict(nx=13, ny=9, W=36, num_nets=200, window=5, seed=7),
                                dict(nx=24, ny=24, W=100, num_nets=600, window=8), dict(nx=7, ny=12, W=8, L=3, num_nets=60, window=4),
                                dict(nx=9, ny=9, W=14, L=5, num_nets=60, window=4, io_capacity=3)])
def test_device_built_graph_is_bit_identical_to_the_host_generator(kw, emu_lib):
    """pf_router_create_generated builds the rr graph ON the device from the closed forms of
    pf_gen_device.cuh (node numbering, wire spans, per-row edge order).  It must be the very graph pf_gen.cpp builds with its
    lookup tables and uploads: equal hashes of the 32-byte node records, of the packed edge words and of the ptc numbers, the
    same nets — and therefore the same routing.  (Emulator backend: the same device functions, run in plain loops.)"""
    lib = router.load_library(emu_lib)
    p = router.generate_grid_problem(lib_path=emu_lib, **kw)
    nets, g = router.generate_grid_nets(lib_path=emu_lib, **kw)
    assert nets.gen_num_nodes == p.num_nodes and len(nets.xlow) == 0 and len(nets.edge_to) == 0
    assert np.array_equal(nets.net_terminals, p.net_terminals) and np.array_equal(nets.net_bb, p.net_bb) and np.array_equal(nets.net_ptr, p.net_ptr)
    cfg = router.default_config(lib, num_slots=1, big_slots=1, pop_slack=0.0, max_batch=1)
    A = router.Router(p, cfg, lib_path=emu_lib)
    B = router.Router(nets, cfg, lib_path=emu_lib, generated=g)
    assert A.graph_hash() == B.graph_hash() and A.graph_hash()[3] == p.num_edges
    if kw["nx"] <= 13:                       # one warp is deterministic: identical graphs give identical routings
        from parallel_eda_b200 import pathfinder
        ra, rb = pathfinder.run(A), pathfinder.run(B)
        assert ra.success == rb.success and ra.iterations == rb.iterations and ra.overused == rb.overused   # (W = 8 is unroutable: equally so)
        xa, xb = A.result(), B.result()
        assert xa.serial_num == xb.serial_num and xa.total_wirelength == xb.total_wirelength
        B.reset()                            # reset of a generated router: a kernel, no host arrays to re-upload
        rc = pathfinder.run(B)
        assert rc.success == ra.success and B.result().serial_num == xa.serial_num
    A.close(); B.close()
