"""Vectorised legality check for big results (occupancy recompute + capacity + sink coverage)."""
import numpy as np
from parallel_eda_b200 import pfio
def check_fast(p, r):
    tn=r.trace_node; ty=p.type[tn]
    is_sink=(ty==pfio.SINK)
    # join elements: element right after a SINK within the same net
    prev_sink=np.zeros(len(tn),bool); prev_sink[1:]=is_sink[:-1]
    starts=np.zeros(len(tn),bool); starts[r.trace_ptr[:-1][r.trace_ptr[:-1]<len(tn)]]=True
    join=prev_sink & ~starts
    occ=np.bincount(tn[~join],minlength=p.num_nodes)
    assert np.array_equal(occ, r.occ), 'occ mismatch %d'%int((occ!=r.occ).sum())
    over=int((occ>p.capacity).sum())
    # sinks reached
    want=np.sort(np.concatenate([p.net_terminals[p.net_ptr[i]+1:p.net_ptr[i+1]] for i in p.routed_nets()])) if p.num_nets<50000 else None
    nsink=int(is_sink.sum()); nterm=int((np.diff(p.net_ptr)-1)[p.net_is_global==0].sum())
    assert nsink==nterm, (nsink,nterm)
    # adjacency of consecutive non-join pairs: sample
    rng=np.random.default_rng(0); idx=rng.integers(0,len(tn)-1,size=min(200000,len(tn)-1))
    idx=idx[~is_sink[idx]]
    a=tn[idx]; b=tn[idx+1]
    ok=np.zeros(len(idx),bool)
    maxdeg=int(np.diff(p.row_ptr).max())
    for k in range(maxdeg):
        e=p.row_ptr[a]+k; valid=e<p.row_ptr[a+1]
        ok|= valid & (p.edge_to[np.minimum(e,p.num_edges-1)]==b) & (p.edge_sw[np.minimum(e,p.num_edges-1)]==r.trace_switch[idx])
    assert ok.all(), 'adjacency failed %d'%int((~ok).sum())
    return {'overused':over,'sinks':nsink,'sampled_edges':len(idx)}
