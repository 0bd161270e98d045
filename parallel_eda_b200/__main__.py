"""Stand-alone router CLI over the flat containers (include/pf_file.h):

    python -m parallel_eda_b200 route PROBLEM.pfp[.xz] [--timing-graph G.pftg[.xz]] [--result OUT.pfr]
                                      [--max-iters K] [--device D] [--check] [--verbose]
    python -m parallel_eda_b200 check PROBLEM.pfp[.xz] RESULT.pfr[.xz]      (device check_route of any result)
    python -m parallel_eda_b200 info  PROBLEM.pfp[.xz]

`route` runs try_timing_driven_route / try_breadth_first_route (opts.router_algorithm in the problem) on one GPU; with
--timing-graph the static timing analysis between iterations runs on the device, otherwise a problem with
opts.timing_analysis_enabled is routed with all criticalities at their initial value.  There is no CPU path.
"""
from __future__ import annotations

import argparse
import json
import sys
import time

from . import pfio, router


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m parallel_eda_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    r = sub.add_parser("route")
    r.add_argument("problem"); r.add_argument("--timing-graph"); r.add_argument("--result")
    r.add_argument("--max-iters", type=int, default=0); r.add_argument("--device", type=int, default=0)
    r.add_argument("--check", action="store_true", help="run the device check_route on the result")
    r.add_argument("--verbose", action="store_true")
    c = sub.add_parser("check"); c.add_argument("problem"); c.add_argument("result"); c.add_argument("--device", type=int, default=0)
    i = sub.add_parser("info"); i.add_argument("problem")
    a = ap.parse_args(argv)

    p = pfio.read_problem(a.problem)
    if a.cmd == "info":
        print(json.dumps({"nx": p.nx, "ny": p.ny, "rr_nodes": p.num_nodes, "rr_edges": p.num_edges, "nets": p.num_nets,
                          "routed_nets": int(len(p.routed_nets())), "terminals": p.num_terminals,
                          "opts": {k: float(p.opts[k]) for k in p.opts.dtype.names}}))
        return 0
    if a.cmd == "check":
        res = pfio.read_result(a.result)
        R = router.Router(p, router.default_config(device=a.device))
        rep = R.check_route(res)
        R.close()
        print(json.dumps(rep))
        return 0 if rep["ok"] and rep["overused_nodes"] == 0 else 1
    if a.max_iters > 0:
        p.opts["max_router_iterations"] = a.max_iters
    tg = pfio.read_timing_graph(a.timing_graph) if a.timing_graph else None
    cfg = router.default_config(device=a.device, verbose=1 if a.verbose else 0)
    t = time.perf_counter()
    res = router.try_timing_driven_route(p, cfg, timing_graph=tg)
    out = {"success": int(res.success), "iterations": int(res.iterations), "wirelength": int(res.total_wirelength),
           "serial_num": int(res.serial_num), "seconds": round(time.perf_counter() - t, 4)}
    if a.check:
        R = router.Router(p, router.default_config(device=a.device))
        out["check_route"] = R.check_route(res)
        R.close()
    if a.result:
        pfio.write_result(a.result, res)
    print(json.dumps(out))
    return 0 if res.success else 1


if __name__ == "__main__":
    sys.exit(main())
