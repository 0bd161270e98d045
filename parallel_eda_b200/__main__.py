"""Stand-alone router CLI over the flat containers (include/pf_file.h):

    python -m parallel_eda_b200 route PROBLEM.pfp[.xz] [--timing-graph G.pftg[.xz]] [--result OUT.pfr]
                                      [--route-file OUT.route [--names N.pfn[.xz]]]
                                      [--max-iters K] [--device D] [--check] [--verbose]
    python -m parallel_eda_b200 check PROBLEM.pfp[.xz] RESULT.pfr[.xz] | ROUTING.route   (device check_route)
    python -m parallel_eda_b200 info  PROBLEM.pfp[.xz]
    python -m parallel_eda_b200 gen   OUT.pfp --nx 400 [--ny 400] --width 100 --nets 200000 [--sinks 3] [--seed 1]   (no GPU)
    python -m parallel_eda_b200 print-route PROBLEM.pfp[.xz] RESULT.pfr[.xz] OUT.route [--names N.pfn[.xz]]   (no GPU)
    python -m parallel_eda_b200 read-route  PROBLEM.pfp[.xz] IN.route OUT.pfr                                  (no GPU)

--route-file / print-route write VPR's .route text (print_route, reference route_common.c:1322; include/pf_text.h);
without --names (the reference-side export of net / block names) nets are called n<i> and the IO ring is assumed.

`route` runs try_timing_driven_route / try_breadth_first_route (opts.router_algorithm in the problem) on one GPU; with
--timing-graph the static timing analysis between iterations runs on the device, otherwise a problem with
opts.timing_analysis_enabled is routed with all criticalities at their initial value.  There is no CPU path.
"""
from __future__ import annotations

import argparse
import json
import sys
import time

from . import pfio, router, textio


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m parallel_eda_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    r = sub.add_parser("route")
    r.add_argument("problem"); r.add_argument("--timing-graph"); r.add_argument("--result")
    r.add_argument("--route-file"); r.add_argument("--names")
    r.add_argument("--max-iters", type=int, default=0); r.add_argument("--device", type=int, default=0)
    r.add_argument("--check", action="store_true", help="run the device check_route on the result")
    r.add_argument("--verbose", action="store_true")
    c = sub.add_parser("check"); c.add_argument("problem"); c.add_argument("result"); c.add_argument("--device", type=int, default=0)
    i = sub.add_parser("info"); i.add_argument("problem")
    w = sub.add_parser("print-route"); w.add_argument("problem"); w.add_argument("result"); w.add_argument("route_file"); w.add_argument("--names")
    g = sub.add_parser("read-route"); g.add_argument("problem"); g.add_argument("route_file"); g.add_argument("result")
    n = sub.add_parser("gen", help="uniform k6_N10-style grid + random nets (pf_gen_grid_problem, include/pf_gen.h)")
    n.add_argument("problem"); n.add_argument("--nx", type=int, required=True); n.add_argument("--ny", type=int, default=0)
    n.add_argument("--width", type=int, default=100); n.add_argument("--nets", type=int, required=True)
    n.add_argument("--sinks", type=int, default=3); n.add_argument("--seed", type=int, default=1)
    q = sub.add_parser("read-net", help="packed netlist (.net) -> block / net statistics (pf_net_read, include/pf_text.h; no GPU)")
    q.add_argument("net_file"); q.add_argument("--place", help="a .place file of the same circuit: also report the blocks it positions")
    a = ap.parse_args(argv)

    if a.cmd == "read-net":
        import collections
        t = time.perf_counter()
        nl = textio.read_netlist(a.net_file)
        fan = [int(x) - 1 for x in (nl.net_ptr[1:] - nl.net_ptr[:-1])]
        out = {"blocks": nl.num_blocks, "block_types": dict(collections.Counter(nl.block_types)), "nets": nl.num_nets,
               "global_nets": [nl.net_names[i] for i in range(nl.num_nets) if nl.net_is_global[i]],
               "sinks": int(sum(fan)), "max_fanout": max(fan) if fan else 0, "seconds": round(time.perf_counter() - t, 3)}
        if a.place:
            # read_place (base/read_place.c:15) on the blocks the netlist names: the grid size comes from the file's own header
            import re
            import numpy as np
            m = re.search(r"Array size:\s*(\d+)\s*x\s*(\d+)", open(a.place).read(4096))
            if not m:
                raise SystemExit("%s: no 'Array size:' header" % a.place)
            nx, ny = int(m.group(1)), int(m.group(2))
            names = textio.Names.build(nx, ny, nl.net_names, np.zeros((nx + 2) * (ny + 2), np.uint8), [(b, 0, 0, 0) for b in nl.block_names])
            out["placed_blocks"] = textio.read_place(a.place, names)
            out["grid"] = [nx, ny]
            # half-perimeter wirelength of the routed (non-global) nets from the placement, block coordinates only (place.c:2083)
            bx, by = names.block_x, names.block_y
            hp = 0
            for i in range(nl.num_nets):
                if nl.net_is_global[i]:
                    continue
                blk = nl.net_block[nl.net_ptr[i]:nl.net_ptr[i + 1]]
                hp += int(bx[blk].max() - bx[blk].min() + by[blk].max() - by[blk].min())
            out["half_perimeter_of_block_positions"] = hp
        print(json.dumps(out))
        return 0
    if a.cmd == "gen":
        t = time.perf_counter()
        p = router.generate_grid_problem(nx=a.nx, ny=a.ny or a.nx, W=a.width, num_nets=a.nets, sinks_per_net=a.sinks, seed=a.seed)
        pfio.write_problem(a.problem, p)
        print(json.dumps({"rr_nodes": p.num_nodes, "rr_edges": p.num_edges, "nets": p.num_nets, "terminals": p.num_terminals,
                          "seconds": round(time.perf_counter() - t, 2)}))
        return 0
    p = pfio.read_problem(a.problem)
    if a.cmd == "info":
        print(json.dumps({"nx": p.nx, "ny": p.ny, "rr_nodes": p.num_nodes, "rr_edges": p.num_edges, "nets": p.num_nets,
                          "routed_nets": int(len(p.routed_nets())), "terminals": p.num_terminals,
                          "opts": {k: float(p.opts[k]) for k in p.opts.dtype.names}}))
        return 0
    if a.cmd == "print-route":
        names = textio.read_names(a.names) if a.names else textio.synthetic_names(p)
        textio.check_names(names, p)
        textio.write_route(a.route_file, p, names, pfio.read_result(a.result))
        return 0
    if a.cmd == "read-route":
        res = textio.read_route(a.route_file, p)
        pfio.write_result(a.result, res)
        print(json.dumps({"nets": p.num_nets, "trace_elements": int(len(res.trace_node)), "wirelength": int(res.total_wirelength),
                          "serial_num": int(res.serial_num)}))
        return 0
    if a.cmd == "check":
        from_text = a.result.endswith(".route")
        res = textio.read_route(a.result, p) if from_text else pfio.read_result(a.result)
        if from_text:       # a .route file lists no occupancy: the traces' own; locally reserved OPINs are not in the file
            from . import check_route as _cr
            res.occ = _cr.recompute_occupancy(p, res)
        R = router.Router(p, router.default_config(device=a.device))
        rep = R.check_route(res)
        R.close()
        if from_text and rep["bad_nets"] == 0 and rep["reserved_opins"] == 0 and rep["occupancy_mismatch"] == 1:
            rep["ok"], rep["occupancy_mismatch"] = 1, 0
            rep["note"] = "locally reserved OPINs (reserve_locally_used_opins) are not part of a .route file"
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
    if a.route_file:
        names = textio.read_names(a.names) if a.names else textio.synthetic_names(p)
        textio.check_names(names, p)
        textio.write_route(a.route_file, p, names, res)
    print(json.dumps(out))
    return 0 if res.success else 1


if __name__ == "__main__":
    sys.exit(main())
