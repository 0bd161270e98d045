"""Device static timing analysis vs the CPU restatement of the reference's, on the golden fixtures.
usage: python tools/sta_bench.py [name ...]   (needs a GPU; oracle/_build/libpf_oracle.so for the CPU side)"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from parallel_eda_b200 import pfio, router
from test_sta_golden import _TG, c_timing_graph
G = os.path.join(ROOT, "tests", "golden")
lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libpf_oracle.so"))
lib.pf_oracle_sta.argtypes = [C.POINTER(_TG), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
import dataclasses


def replicate(p, g, delay, K):
    """K disjoint copies of a timing graph (and of the nets that feed it): a wide graph of the same depth."""
    N, E, T, n = g.num_tnodes, len(g.edge_to), p.num_terminals, p.num_nets
    rep = lambda a, step: np.concatenate([a + k * step for k in range(K)])
    lv_nodes = np.concatenate([np.concatenate([g.level_nodes[g.level_ptr[l]:g.level_ptr[l + 1]] + k * N for k in range(K)]) for l in range(g.num_levels)])
    g2 = pfio.TimingGraph(np.concatenate([[0], rep(g.edge_ptr[1:], E)]).astype(np.int32), rep(g.edge_to, N).astype(np.int32), np.tile(g.edge_Tdel, K),
                          np.tile(g.type, K), np.tile(g.clock_domain, K), np.tile(g.clock_delay, K), (g.level_ptr * K).astype(np.int32),
                          lv_nodes.astype(np.int32), g.constraint, np.where(np.tile(g.net_driver, K) >= 0, rep(g.net_driver, N), -1).astype(np.int32))
    p2 = dataclasses.replace(p, net_ptr=np.concatenate([[0], rep(p.net_ptr[1:], T)]).astype(np.int32), net_terminals=np.tile(p.net_terminals, K),
                             net_is_global=np.tile(p.net_is_global, K), net_bb=np.tile(p.net_bb, K))
    return p2, g2, np.tile(delay, K)


cases = [(nm, 1) for nm in (sys.argv[1:] or ["toy_w64", "hub_w90", "mid_w200"])] + [("mid_w200", 32)]
for name, K in cases:
    p = pfio.read_problem(os.path.join(G, name + ".pfp.xz")); g = pfio.read_timing_graph(os.path.join(G, name + ".pftg.xz"))
    v = pfio.read_sta_vectors(os.path.join(G, name + ".pfsta.xz"))
    if K > 1:
        p, g, dd = replicate(p, g, v.net_delay[-1], K)
        v = pfio.StaVectors(dd[None, :], dd[None, :], v.cpd[-1:])
        name = "%s x%d" % (name, K)
    tg, keep = c_timing_graph(g); net_ptr = np.ascontiguousarray(p.net_ptr, dtype=np.int32)
    d = np.ascontiguousarray(v.net_delay[-1]); crit = np.zeros(p.num_terminals, np.float32); cpd = C.c_float(0)
    n = 20
    t = time.perf_counter()
    for _ in range(n): lib.pf_oracle_sta(C.byref(tg), net_ptr.ctypes.data, d.ctypes.data, crit.ctypes.data, C.byref(cpd))
    cpu = (time.perf_counter() - t) / n
    s = router.Sta(g, p)
    s.analyze(d)
    t = time.perf_counter()
    for _ in range(n): c2, _ = s.analyze(d)
    gpu = (time.perf_counter() - t) / n
    assert np.array_equal(c2.view(np.uint32), crit.view(np.uint32))
    print("%s: %d tnodes, %d levels: CPU restatement %.3f ms, device (host buffers, incl. copies) %.3f ms" % (name, g.num_tnodes, g.num_levels, cpu * 1e3, gpu * 1e3))
