"""Independent legality and delay checker for a routing result (host logic, numpy/Python).

Mirrors what the reference runs after every successful route:
  * check_route        (reference vpr/SRC/route/check_route.c:27-155): each net's traceback starts at
    its SOURCE, every segment ends at one of its SINKs, consecutive elements are joined by a real
    rr edge carrying the recorded switch, later segments branch off a node already in the net, every
    sink is reached the right number of times; occupancy recomputed from the traces
    (recompute_occupancy_from_scratch, check_route.c:535-597) matches and respects capacity.
  * timing_driven_check_net_delays (route/route_timing.c:964-1004, tolerance ERROR_TOL 1e-4):
    Elmore delay of every sink recomputed from scratch from the traceback alone
    (timing/net_delay.c:181 load_net_delay_from_routing) versus the router's incremental value.
It never looks at how the result was produced, so it judges the CUDA router, the CPU oracle and
the reference's own golden results alike.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict

import numpy as np

from . import pfio


class RouteCheckError(AssertionError):
    pass


def _edge_lookup(p: pfio.Problem):
    row, to, sw = p.row_ptr, p.edge_to, p.edge_sw

    def has_edge(a: int, b: int, s: int) -> bool:
        lo, hi = int(row[a]), int(row[a + 1])
        seg = to[lo:hi]
        hit = np.nonzero(seg == b)[0]
        return any(int(sw[lo + k]) == s for k in hit)

    return has_edge


def net_segments(p: pfio.Problem, nodes: np.ndarray):
    """Split one net's traceback into segments (each ends at a SINK)."""
    segs, cur = [], []
    for k, v in enumerate(nodes):
        cur.append(k)
        if p.type[v] == pfio.SINK:
            segs.append(cur)
            cur = []
    if cur:
        raise RouteCheckError("traceback does not end in a SINK")
    return segs


def recompute_occupancy(p: pfio.Problem, r: pfio.Result) -> np.ndarray:
    occ = np.zeros(p.num_nodes, dtype=np.int64)
    for inet in p.routed_nets():
        nodes, _ = r.net_trace(int(inet))
        first = True
        for seg in net_segments(p, nodes) if len(nodes) else []:
            idx = seg if first else seg[1:]      # the join node is not counted again (route_common.c:570-575)
            np.add.at(occ, nodes[idx], 1)
            first = False
    return occ


def elmore_from_trace(p: pfio.Problem, nodes: np.ndarray, switches: np.ndarray) -> Dict[int, list]:
    """Per-sink Elmore delays recomputed from the traceback: {sink rr node: [delays in trace order]}."""
    # build the RC tree
    children = defaultdict(list)   # tree id -> [(child tree id, switch)]
    tnode = []                     # tree id -> rr node
    where = {}                     # rr node -> latest tree id (join lookups)
    segs = net_segments(p, nodes)
    sink_ids = []
    for si, seg in enumerate(segs):
        start = 0
        if si == 0:
            tid = len(tnode); tnode.append(int(nodes[seg[0]])); where[int(nodes[seg[0]])] = tid
            prev_tid, prev_k = tid, seg[0]
            start = 1
        else:
            j = int(nodes[seg[0]])
            if j not in where:
                raise RouteCheckError("segment joins at node %d which is not in the net" % j)
            prev_tid, prev_k = where[j], seg[0]
            start = 1
        for k in seg[start:]:
            tid = len(tnode); tnode.append(int(nodes[k]))
            children[prev_tid].append((tid, int(switches[prev_k])))
            where[int(nodes[k])] = tid
            prev_tid, prev_k = tid, k
        sink_ids.append(prev_tid)
    n = len(tnode)
    # downstream capacitance (post-order) — buffered switches isolate (net_delay.c load_rc_tree_C)
    C_down = [0.0] * n
    order = list(range(n))          # parents are created before children
    for tid in reversed(order):
        c = float(p.C[tnode[tid]])
        for ch, s in children[tid]:
            if not p.switches["buffered"][s]:
                c += C_down[ch]
        C_down[tid] = np.float32(c)
    # arrival times (pre-order) — net_delay.c load_rc_tree_T
    T = [0.0] * n
    T[0] = np.float32(0.5 * float(p.R[tnode[0]]) * float(C_down[0]))
    for tid in order:
        for ch, s in children[tid]:
            t = float(T[tid]) + float(p.switches["R"][s]) * float(C_down[ch]) + float(p.switches["Tdel"][s])
            t += 0.5 * float(C_down[ch]) * float(p.R[tnode[ch]])
            T[ch] = t
    out = defaultdict(list)
    for tid in sink_ids:
        out[tnode[tid]].append(float(T[tid]))
    return out


def check_route(p: pfio.Problem, r: pfio.Result, check_delays: bool = True, require_legal: bool = True,
                delay_tol: float = 1e-4) -> dict:
    has_edge = _edge_lookup(p)
    total_wl = 0
    max_delay = 0.0
    for inet in p.routed_nets():
        inet = int(inet)
        t0, t1 = int(p.net_ptr[inet]), int(p.net_ptr[inet + 1])
        terms = p.net_terminals[t0:t1]
        nodes, sws = r.net_trace(inet)
        if t1 - t0 - 1 == 0:
            continue
        if len(nodes) == 0:
            raise RouteCheckError("net %d has no traceback" % inet)
        if nodes[0] != terms[0]:
            raise RouteCheckError("net %d does not start at its SOURCE" % inet)
        segs = net_segments(p, nodes)
        seen = set()
        reached = defaultdict(int)
        for si, seg in enumerate(segs):
            if si > 0 and int(nodes[seg[0]]) not in seen:
                raise RouteCheckError("net %d segment %d starts at node %d outside the net" % (inet, si, nodes[seg[0]]))
            for a, b in zip(seg[:-1], seg[1:]):
                if not has_edge(int(nodes[a]), int(nodes[b]), int(sws[a])):
                    raise RouteCheckError("net %d: no edge %d -> %d with switch %d" % (inet, nodes[a], nodes[b], sws[a]))
            if sws[seg[-1]] != pfio.OPEN:
                raise RouteCheckError("net %d: SINK element carries a switch" % inet)
            for k in (seg if si == 0 else seg[1:]):
                v = int(nodes[k])
                ty = p.type[v]
                if ty in (pfio.CHANX, pfio.CHANY):
                    total_wl += 1 + int(p.xhigh[v]) - int(p.xlow[v]) + int(p.yhigh[v]) - int(p.ylow[v])
                seen.add(v)
            reached[int(nodes[seg[-1]])] += 1
        want = defaultdict(int)
        for s in terms[1:]:
            want[int(s)] += 1
        if dict(want) != dict(reached):
            raise RouteCheckError("net %d: sinks reached %s, wanted %s" % (inet, dict(reached), dict(want)))
        if check_delays:
            fresh = elmore_from_trace(p, nodes, sws)
            for k, s in enumerate(terms[1:], start=1):
                got = float(r.net_delay[t0 + k])
                cands = fresh[int(s)]
                if not any(abs(1.0 - got / c) <= delay_tol if c != 0.0 else abs(got) <= delay_tol for c in cands):
                    raise RouteCheckError("net %d pin %d: incremental delay %g, from-scratch %s" % (inet, k, got, cands))
                max_delay = max(max_delay, got)
    occ = recompute_occupancy(p, r)
    extra = r.occ.astype(np.int64) - occ
    if (extra < 0).any():
        raise RouteCheckError("reported occupancy below the traces at %d nodes" % int((extra < 0).sum()))
    reserved = int(np.asarray(p.opin_group_count).sum())
    if int(extra.sum()) != reserved:
        raise RouteCheckError("occupancy not explained by traces + locally used OPINs: %d vs %d" % (int(extra.sum()), reserved))
    if (extra[p.type != pfio.OPIN] != 0).any():
        raise RouteCheckError("non-OPIN occupancy differs from the traces")
    overused = int((r.occ > p.capacity).sum())
    if require_legal and overused:
        raise RouteCheckError("%d rr nodes over capacity" % overused)
    return {"wirelength": total_wl, "overused": overused, "max_net_delay": max_delay}


def check_route_fast(p: pfio.Problem, r: pfio.Result, sample: int = 200000) -> dict:
    """Vectorised subset of check_route for results with millions of trace elements: occupancy recomputed from
    the traces must equal the reported one (plus reserved OPINs), every sink must be reached, and a random sample
    of consecutive trace pairs must be real rr edges carrying the recorded switch."""
    tn = r.trace_node
    is_sink = p.type[tn] == pfio.SINK
    prev_sink = np.zeros(len(tn), bool)
    prev_sink[1:] = is_sink[:-1]
    starts = np.zeros(len(tn), bool)
    first = r.trace_ptr[:-1]
    starts[first[first < len(tn)]] = True
    join = prev_sink & ~starts
    occ = np.bincount(tn[~join], minlength=p.num_nodes)
    extra = r.occ.astype(np.int64) - occ
    if (extra < 0).any() or int(extra.sum()) != int(np.asarray(p.opin_group_count).sum()):
        raise RouteCheckError("occupancy not explained by the traces")
    nsink = int(is_sink.sum())
    nterm = int((np.diff(p.net_ptr) - 1)[p.net_is_global == 0].sum())
    if nsink != nterm:
        raise RouteCheckError("%d sinks reached, %d wanted" % (nsink, nterm))
    rng = np.random.default_rng(0)
    idx = rng.integers(0, max(len(tn) - 1, 1), size=min(sample, max(len(tn) - 1, 1)))
    idx = idx[~is_sink[idx]]
    a, b = tn[idx], tn[idx + 1]
    ok = np.zeros(len(idx), bool)
    for k in range(int(np.diff(p.row_ptr).max())):
        e = p.row_ptr[a] + k
        valid = e < p.row_ptr[a + 1]
        e = np.minimum(e, p.num_edges - 1)
        ok |= valid & (p.edge_to[e] == b) & (p.edge_sw[e] == r.trace_switch[idx])
    if not ok.all():
        raise RouteCheckError("%d sampled trace pairs are not rr edges" % int((~ok).sum()))
    return {"overused": int((r.occ > p.capacity).sum()), "sinks": nsink, "sampled_edges": int(len(idx))}
