"""The GPU-vs-reference parity bar in ONE place.

Route trees are integer rr-node lists whose shape depends on float-cost ties and on which nets are in flight together, so
identity with the reference's routing is judged on what the routing is FOR:
  * legality            independent check_route, occupancy recomputed from the traces bit-equal to the device's
  * delays              every incremental Elmore delay equal to a from-scratch recomputation (1e-4, the reference's ERROR_TOL)
  * iterations          within the reference's own budget (max_router_iterations is never raised) and <= ITER_FACTOR x its count
  * wirelength          <= WL_TOL x the reference's on the same input
  * timing              criticality-weighted sink delay (what the timing-driven cost minimises) / critical path <= TD_TOL x
One warp (the reference's serial order) is held to ONE_WARP_WL.  Every test that measures these appends the numbers to
gpurun_out/parity_measured.jsonl so the tolerances can be read against evidence, not set to whatever passes."""
import json
import os

ITER_FACTOR = 1.5
WL_TOL = 1.08
TD_TOL = 1.03
ONE_WARP_WL = 1.03

_LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_measured.jsonl")


def record(test, **kw):
    try:
        os.makedirs(os.path.dirname(_LOG), exist_ok=True)
        with open(_LOG, "a") as f:
            f.write(json.dumps(dict(test=test, **kw)) + "\n")
    except OSError:
        pass


def check(test, name, r, g, weighted=None, iters=True):
    """r: our pfio.Result, g: the reference's golden result; weighted: (ours, reference) criticality-weighted delay or cpd."""
    wl = r.total_wirelength / g.total_wirelength
    wd = (weighted[0] / weighted[1]) if weighted else None
    record(test, fixture=name, iterations=int(r.iterations), ref_iterations=int(g.iterations), wl_ratio=round(wl, 4),
           td_ratio=None if wd is None else round(wd, 4), success=int(r.success))
    assert r.success == 1, (name, "not legal within the reference's iteration budget", int(r.iterations))
    assert wl <= WL_TOL, (name, "wirelength ratio", wl)
    if wd is not None:
        assert wd <= TD_TOL, (name, "timing ratio", wd)
    if iters:
        assert r.iterations <= int(ITER_FACTOR * g.iterations) + 1, (name, int(r.iterations), int(g.iterations))
