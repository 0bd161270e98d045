"""The GPU-vs-reference parity bar in ONE place.

Route trees are integer rr-node lists whose shape depends on float-cost ties and on which nets are in flight together, so
identity with the reference's routing is judged on what the routing is FOR:
  * legality            independent check_route, occupancy recomputed from the traces bit-equal to the device's
  * delays              every incremental Elmore delay equal to a from-scratch recomputation (1e-4, the reference's ERROR_TOL)
  * iterations          within the reference's own budget (max_router_iterations is never raised) and <= ITER_FACTOR x its count
  * wirelength          <= WL_TOL x the reference's on the same input
  * timing              criticality-weighted sink delay (what the timing-driven cost minimises) / critical path <= TD_TOL x
One warp (the reference's serial order) is deterministic and held to ONE_WARP_WL.
With all warps the schedule is not deterministic — which of two nets in flight commits first depends on timing — and on a circuit
of a few hundred nets one run differs from the next by more than the systematic distance to the reference (B200, 6 runs per
fixture, profiles/r02z_repeat.txt: het_w70 wirelength x1.022 .. x1.085, heq_w70 x1.004 .. x1.057, weighted delay x1.000 .. x1.023).
The small fixtures are therefore routed RUNS times (0.1 - 0.5 s each): EVERY run must be legal inside the reference's iteration
budget and inside the *_RUN caps, and the MEDIAN run is held to the tolerances; circuits from 10 k nets up, where the spread is
a few 0.1 %, are routed once and held to BIG_WL_TOL.  Every test that measures these appends the numbers to
gpurun_out/parity_measured.jsonl so the tolerances can be read against evidence, not set to whatever passes."""
import json
import os

ITER_FACTOR = 1.5
WL_TOL = 1.08          # median run (single run for check())
TD_TOL = 1.03
WL_RUN_TOL = 1.12      # any single run of a small fixture
TD_RUN_TOL = 1.05
BIG_WL_TOL = 1.05      # >= 10 k nets, one run (measured x1.031 / x1.038 on the 11 k / 32 k LUT circuits, x1.001 on 50 k 4-pin nets)
ONE_WARP_WL = 1.03
RUNS = 5

_LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_measured.jsonl")


def record(test, **kw):
    try:
        os.makedirs(os.path.dirname(_LOG), exist_ok=True)
        with open(_LOG, "a") as f:
            f.write(json.dumps(dict(test=test, **kw)) + "\n")
    except OSError:
        pass


def check(test, name, r, g, weighted=None, iters=True, wl_tol=None):
    """r: our pfio.Result, g: the reference's golden result; weighted: (ours, reference) criticality-weighted delay or cpd."""
    wl_tol = WL_TOL if wl_tol is None else wl_tol
    wl = r.total_wirelength / g.total_wirelength
    wd = (weighted[0] / weighted[1]) if weighted else None
    record(test, fixture=name, iterations=int(r.iterations), ref_iterations=int(g.iterations), wl_ratio=round(wl, 4),
           td_ratio=None if wd is None else round(wd, 4), success=int(r.success))
    assert r.success == 1, (name, "not legal within the reference's iteration budget", int(r.iterations))
    assert wl <= wl_tol, (name, "wirelength ratio", wl)
    if wd is not None:
        assert wd <= TD_TOL, (name, "timing ratio", wd)
    if iters:
        assert r.iterations <= int(ITER_FACTOR * g.iterations) + 1, (name, int(r.iterations), int(g.iterations))


def check_runs(test, name, route, g, weighted=None, runs=RUNS):
    """route() -> Result, called `runs` times; weighted(r) -> (ours, reference) or None.  Returns the last result."""
    rows, r = [], None
    for _ in range(runs):
        r = route()
        wl = r.total_wirelength / g.total_wirelength
        wd = None
        if weighted is not None:
            a, b = weighted(r)
            wd = a / b
        record(test, fixture=name, iterations=int(r.iterations), ref_iterations=int(g.iterations), wl_ratio=round(wl, 4),
               td_ratio=None if wd is None else round(wd, 4), success=int(r.success), runs=runs)
        assert r.success == 1, (name, "not legal within the reference's iteration budget", int(r.iterations))
        assert wl <= WL_RUN_TOL, (name, "wirelength ratio of one run", wl)
        assert wd is None or wd <= TD_RUN_TOL, (name, "timing ratio of one run", wd)
        rows.append((wl, wd, int(r.iterations)))
    med = lambda xs: sorted(xs)[len(xs) // 2]
    wl_m, it_m = med([x[0] for x in rows]), med([x[2] for x in rows])
    print("%s %s: %d runs, wirelength x%s, iterations %s (reference %d)%s" % (
        test, name, runs, [round(x[0], 3) for x in rows], [x[2] for x in rows], int(g.iterations),
        "" if weighted is None else ", timing x%s" % [round(x[1], 3) for x in rows]))
    assert wl_m <= WL_TOL, (name, "median wirelength ratio", wl_m, rows)
    if weighted is not None:
        assert med([x[1] for x in rows]) <= TD_TOL, (name, "median timing ratio", rows)
    assert it_m <= int(ITER_FACTOR * g.iterations) + 1, (name, "median iterations", rows, int(g.iterations))
    return r
