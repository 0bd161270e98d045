#!/usr/bin/env python3
"""Benchmark of the PathFinder --route hot path (BASELINE.json metric: nets routed/sec and total
route time vs the reference CPU router).

    python bench.py --gpus N --steps K --warmup W            # our CUDA router (torchrun for N > 1)
    python bench.py --impl reference --gpus N ...            # the reference's own CPU router

Workload (config.workload): BASELINE.json configs[4] — synthetic 400x400 CLB k6_N10-style grid, W=100,
200,000 random 4-pin nets, timing analysis off, VPR default router options — the largest configuration
that fits one GPU; configs[1..3] need VTR benchmark files that are neither in the reference nor on the box
(SURVEY.md §8c).  A step is one complete routing of the problem: reset congestion, then PathFinder
iterations until the routing is legal.  N > 1 shards the nets of the SAME problem over N GPUs (strong
scaling) with an NCCL all-gather of the ranks' occupancy event logs twice per iteration.

value   = nets routed (summed over iterations and ranks) / device time of the step, graph already in HBM
e2e     = the same through the public API with HOST buffers: flatten + H2D of the whole problem, route,
          D2H of traces / delays / occupancy, every step
roofline: the dominant kernel (pf_route_kernel) against the measured HBM copy bandwidth
cpu_baseline / --impl reference: the reference's own serial router (oracle/_ref/vpr_ref, compiled from the
          unmodified reference) — or, where that binary is absent, the bit-exact C restatement — on a
          bounded sample of the same problem, on this box's host cores.
cpu_parallel_baseline (extra): the same CPU algorithm on all host threads (oracle/pf_oracle_par.c, kind "port"), first
          iterations of the whole problem — the stand-in for the reference's parallel routers (TBB / MPI: not buildable).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "synthetic {nx}x{ny} CLB k6_N10-style grid, W={W}, {nets} random 4-pin nets, timing off (BASELINE configs[4])"


def workload_config(a, p):
    """`config` of the JSON line: the workload and nothing about the arm that routes it, so that both arms (this router,
    `--impl reference`) print the very same object; what is specific to an arm lives in `parallelism` / `cpu_baseline.sample`."""
    return {"workload": WORKLOAD.format(nx=a.grid, ny=a.grid, W=a.width, nets=a.nets),
            "rr_nodes": int(p.num_nodes), "rr_edges": int(p.num_edges), "sinks": int(p.num_terminals - p.num_nets),
            "router_opts": "VPR defaults: astar 1.2, pres_fac 0.5 x1.3, acc_fac 1, bb_factor 3, max 50 iterations",
            "l2": ("working set (node records %d MB + edges %d MB) far exceeds the 126 MB L2; no flush needed" if (p.num_nodes * 32 + p.num_edges * 4) >> 20 > 252
                   else "working set (node records %d MB + edges %d MB) is NOT larger than the L2: a non-default size, L2 not flushed between steps")
                  % (p.num_nodes * 32 >> 20, p.num_edges * 4 >> 20)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "sv0", "bgm"],
                    help="cfg4: BASELINE configs[4] (default, the bench line).  sv0 / bgm: the size-matched stand-ins for configs[1] / [2] "
                         "(tests/golden/big): timing-driven route on the heterogeneous fabric with the device STA in the loop, 1 GPU")
    ap.add_argument("--grid", type=int, default=400)
    ap.add_argument("--nets", type=int, default=200000)
    ap.add_argument("--width", type=int, default=100)
    ap.add_argument("--cpu-sample-nets", type=int, default=25000)
    ap.add_argument("--cpu-sample-iters", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-full-reference", action="store_true", help="--impl reference: skip the one complete routing by vpr_ref (≈ 2 min)")
    ap.add_argument("--upload-graph", action="store_true", help="build the rr graph on the host and upload it (round-1 path) instead of generating it on the device")
    ap.add_argument("--step-api", action="store_true", help="drive the iterations from Python through the step API and torch collectives "
                    "(the round-1 path) instead of pf_route_run with the in-library transport")
    ap.add_argument("--max-batch", type=int, default=0, help="tuning: labels settled per step (0 = library default)")
    ap.add_argument("--pop-slack", type=float, default=-1.0, help="tuning: delta bucket width (<0 = library default)")
    ap.add_argument("--inflight-div", type=int, default=0)
    ap.add_argument("--min-slots", type=int, default=0, help="tuning: lower bound on nets in flight (0 = library default)")
    ap.add_argument("--slots", type=int, default=0)
    ap.add_argument("--cfg", default="", help="tuning: extra pf_config fields, e.g. 'validate_commits=-1 ripple=-1'")
    ap.add_argument("--sync-rounds", type=int, default=2, help="multi-GPU: occupancy syncs per PathFinder iteration")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.stop = index, [], False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_sample(problem_path: str, nets: int, iters: int):
    """Times the reference's serial router on a bounded sample.  Returns (nets_per_s, kind, cores, sample, secs)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "vpr_ref")
    orc = os.path.join(ROOT, "oracle", "_build", "pf_oracle_cli")
    sample = "first %d nets (fanout order) x first %d PathFinder iterations of the same problem, 1 thread" % (nets, iters)
    if os.path.exists(ref):
        cmd, kind, tag = [ref, "inject", problem_path, "--max_iters", str(iters), "--limit_nets", str(nets)], "reference", "PF_REF"
    else:
        if not os.path.exists(orc):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        cmd, kind, tag = [orc, problem_path, "--max_iters", str(iters), "--limit_nets", str(nets)], "port", "PF_ORACLE"
    r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    m = re.search(tag + r" route success=\d+ iterations=(\d+).*?route_time_s=([0-9.]+)", r.stderr)
    if not m:
        raise RuntimeError("CPU baseline run failed: " + r.stderr[-500:])
    its, secs = int(m.group(1)), float(m.group(2))
    return nets * its / secs, kind, 1, sample, secs


def cpu_parallel_sample(problem_path: str, iters: int):
    """The multi-threaded CPU router (oracle/pf_oracle_par.c: the bit-exact serial restatement run by all host threads on
    shared occupancy — a stand-in for the reference's parallel routers, which need TBB / MPI / Boost and cannot be built):
    the first `iters` PathFinder iterations of the WHOLE problem.  Returns a cpu_baseline-shaped dict; never raises."""
    try:
        exe = os.path.join(ROOT, "oracle", "_build", "pf_oracle_par_cli")
        if not os.path.exists(exe):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        try:
            avail = len(os.sched_getaffinity(0))                   # the cores this process may run on (cpuset-aware)
        except Exception:
            avail = os.cpu_count() or 1
        threads = max(1, min(avail, 32))                           # 0.44 GB of search state per thread on the 18 M-node graph
        r = subprocess.run([exe, problem_path, "--threads", str(threads), "--max_iters", str(iters)], stdout=subprocess.DEVNULL,
                           stderr=subprocess.PIPE, text=True, timeout=900)
        m = re.search(r"PF_ORACLE_PAR threads=(\d+) .*?route_time_s=([0-9.]+) nets_routed=(\d+) nets_per_s=([0-9.]+)", r.stderr)
        if not m:
            return None
        return {"value": float(m.group(4)), "unit": "nets/s", "cores": int(m.group(1)), "kind": "port",
                "sample": "all nets x first %d PathFinder iterations of the same problem, %s host threads sharing the occupancy arrays "
                          "(oracle/pf_oracle_par.c)" % (iters, m.group(1)), "sample_seconds": float(m.group(2))}
    except Exception:
        return None


def cpu_reference_full(problem_path: str, timeout_s: float = 900.0):
    """ONE complete routing of the same problem by the reference's own serial router (oracle/_ref/vpr_ref inject): the
    like-for-like time-to-legal and the wirelength yardstick.  Returns a dict or None (binary absent / run failed / too slow)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "vpr_ref")
    if not os.path.exists(ref):
        return None
    try:
        with tempfile.TemporaryDirectory() as d:
            out = os.path.join(d, "ref.pfr")
            t0 = time.perf_counter()
            r = subprocess.run([ref, "inject", problem_path, "--result", out], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=timeout_s)
            wall = time.perf_counter() - t0
        m = re.search(r"PF_REF route success=(\d+) iterations=(\d+) route_time_s=([0-9.]+)", r.stderr)
        w = re.search(r"wirelength=(-?\d+)", r.stderr)
        if not m:
            return None
        return {"success": int(m.group(1)), "iterations": int(m.group(2)), "route_time_s": float(m.group(3)), "wall_s": wall,
                "wirelength": int(w.group(1)) if w else None, "cores": 1, "kind": "reference",
                "what": "one complete routing of the same problem by the unmodified reference router (every net re-routed in every iteration)"}
    except Exception:
        return None


def cpu_parallel_full(problem_path: str):
    """The multi-threaded CPU restatement with the DEVICE router's re-route policy (--congested-only) on all host threads,
    run to a legal routing: the like-for-like CPU time-to-legal for the same algorithmic policy."""
    try:
        exe = os.path.join(ROOT, "oracle", "_build", "pf_oracle_par_cli")
        if not os.path.exists(exe):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        try:
            avail = len(os.sched_getaffinity(0))
        except Exception:
            avail = os.cpu_count() or 1
        threads = max(1, min(avail, 32))
        r = subprocess.run([exe, problem_path, "--threads", str(threads), "--congested-only"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE,
                           text=True, timeout=900)
        m = re.search(r"PF_ORACLE_PAR threads=(\d+) success=(\d+) iterations=(\d+) wirelength=(-?\d+) route_time_s=([0-9.]+) nets_routed=(\d+) nets_per_s=([0-9.]+)", r.stderr)
        if not m:
            return None
        return {"value": float(m.group(7)), "unit": "nets/s", "cores": int(m.group(1)), "kind": "port", "success": int(m.group(2)),
                "iterations": int(m.group(3)), "wirelength": int(m.group(4)), "route_time_s": float(m.group(5)), "nets_routed": int(m.group(6)),
                "sample": "the WHOLE routing to legality, %s host threads sharing the occupancy arrays, congested-only re-route policy "
                          "(oracle/pf_oracle_par.c --congested-only)" % m.group(1)}
    except Exception:
        return None


BIG = {"sv0": ("sv0_w220", 220, "stand-in for BASELINE configs[1] (stereovision0): 11 k LUTs + 40 hard multipliers"),
       "bgm": ("bgm_w260", 260, "stand-in for BASELINE configs[2] (bgm): 32 k LUTs + 100 hard multipliers")}


def big_workload_name(a):
    name, w, what = BIG[a.workload]
    return "%s — %s, k6_N10_het fabric, W=%d, timing-driven, packed and placed by the reference (tests/golden/big)" % (name, what, w)


def run_big_reference(a):
    """The UNMODIFIED reference's own timing-driven route of the circuit (vpr_ref flow --route, its own STA in the loop), timed
    by its own 'Routing took' line on this box's CPU; one run per step."""
    import lzma
    import shutil
    name, w, _ = BIG[a.workload]
    c = a.workload
    ref = os.path.join(ROOT, "oracle", "_ref", "vpr_ref")
    big = os.path.join(ROOT, "tests", "golden", "big")
    summary = json.load(open(os.path.join(big, name + ".json")))
    if not os.path.exists(ref):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/vpr_ref was not built (needs /root/reference at build time)"}))
        return
    times, its = [], 0
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(os.path.join(ROOT, "tests", "fixtures", "k6_N10_het.xml"), d)
        shutil.copy(os.path.join(big, c + ".place"), d)
        for ext in ("blif", "net"):
            with lzma.open(os.path.join(big, "%s.%s.xz" % (c, ext))) as f, open(os.path.join(d, "%s.%s" % (c, ext)), "wb") as o:
                o.write(f.read())
        for s in range(a.warmup + a.steps):
            r = subprocess.run([ref, "flow", "k6_N10_het.xml", c, "--nodisp", "--route", "--route_chan_width", str(w)], cwd=d, capture_output=True, text=True)
            m = re.search(r"Routing took ([0-9.]+) seconds", r.stdout)
            mi = re.search(r"Successfully routed after (\d+) routing iterations", r.stdout)
            if not (m and mi):
                raise RuntimeError("reference flow failed: " + r.stdout[-800:])
            if s >= a.warmup:
                times.append(float(m.group(1))); its = int(mi.group(1))
    nets = summary["routed_nets"] * its
    v = nets / (sum(times) / len(times))
    print(json.dumps({"impl": "reference", "metric": "nets_routed_per_sec", "value": v, "unit": "nets/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                      "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": {"workload": big_workload_name(a)},
                      "route": {"iterations": its, "route_time_s": sum(times) / len(times), "wirelength": summary["reference"]["total_wirelength"],
                                "crit_path_delay_ns": summary["reference"]["final_crit_path_delay_ns"]},
                      "cpu_baseline": {"value": v, "unit": "nets/s", "cores": 1, "kind": "reference", "sample": "the whole timing-driven routing, 1 thread"},
                      "e2e": {"value": v, "unit": "nets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_big_ours(a):
    """Timing-driven route of a BASELINE-scale circuit on one GPU: pf_route_run with the device STA in the loop (value: router and
    timing graph resident, device-timed), and end to end through pf_try_timing_driven_route_sta with host buffers."""
    import torch
    from parallel_eda_b200 import pathfinder, pfio, router
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the router has no CPU path")
    name, w, _ = BIG[a.workload]
    big = os.path.join(ROOT, "tests", "golden", "big")
    summary = json.load(open(os.path.join(big, name + ".json")))
    p = pfio.read_problem(os.path.join(big, name + ".pfp.xz"))
    g = pfio.read_timing_graph(os.path.join(big, name + ".pftg.xz"))
    cfg = router.default_config(device=0)
    R = router.Router(p, cfg)
    S = router.Sta(g, p, cfg)
    reps, times = [], []
    with ClockSampler(0) as clk:
        for s in range(a.warmup + a.steps):
            R.reset()
            torch.cuda.synchronize()
            R.timer_start()
            rep = pathfinder.run(R, dsta=S)
            ms = R.timer_stop()
            if s == a.warmup:
                R.timing(reset=True)
            if s >= a.warmup:
                reps.append(rep); times.append(ms)
    tm = R.timing(reset=True)
    res = R.result()
    chk = R.check_route(res)
    assert all(r.success for r in reps) and chk["ok"] == 1 and chk["overused_nodes"] == 0
    t0 = time.perf_counter()
    r2 = router.try_timing_driven_route(p, cfg, timing_graph=g)          # host buffers in, host buffers out
    e2e_s = time.perf_counter() - t0
    nets = sum(r.nets_routed for r in reps)
    total_ms = sum(times)
    visits = sum(r.edge_visits for r in reps); pops = sum(r.heap_pops for r in reps); pushes = sum(r.heap_pushes for r in reps)
    peak, peak_src = measured_peak()
    achieved = (36.0 * visits + 28.0 * pops + 20.0 * pushes) / (tm.route_kernel_ms * 1e-3) / 1e9 if tm.route_kernel_ms > 0 else 0.0
    ref = summary["reference"]
    print(json.dumps({
        "metric": "nets_routed_per_sec", "value": nets / (total_ms * 1e-3), "unit": "nets/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": big_workload_name(a), "rr_nodes": p.num_nodes, "rr_edges": p.num_edges, "nets": summary["routed_nets"],
                   "l2": "graph (%d MB) fits the 126 MB L2; every step starts from a reset congestion state" % ((p.num_nodes * 32 + p.num_edges * 4) >> 20)},
        "route": {"iterations": [r.iterations for r in reps], "route_time_s": total_ms * 1e-3 / a.steps, "wirelength": int(res.total_wirelength),
                  "crit_path_delay_ns": reps[-1].crit_path_delay[-1], "reference_iterations": ref["iterations"], "reference_wirelength": ref["total_wirelength"],
                  "reference_crit_path_delay_ns": ref["final_crit_path_delay_ns"], "reference_route_time_s_build_container": ref["route_time_s_build_container"],
                  "device_check_route": chk},
        "roofline": {"bound": "hbm", "kernel": "pf_route_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "kernel_ms_per_step": tm.route_kernel_ms / a.steps},
        "gpu_launches": int(tm.route_launches + tm.update_launches + tm.aux_launches), "clocks": clk.summary(),
        "e2e": {"value": sum(int(x) for x in r2.iter_stats["nets_routed"]) / e2e_s, "unit": "nets/s", "s_per_step": e2e_s,
                "h2d_bytes_per_step": int(p.num_nodes * 32 + p.num_edges * 4), "d2h_bytes_per_step": int(res.trace_node.nbytes + res.occ.nbytes)}}))
    S.close(); R.close()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if a.workload != "cfg4":
        return run_big_reference(a)
    from parallel_eda_b200 import pfio, router
    p = router.generate_grid_problem(nx=a.grid, ny=a.grid, W=a.width, num_nets=a.nets)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "bench.pfp")
        pfio.write_problem(path, p)
        vals, total = [], 0.0
        for s in range(a.warmup + a.steps):
            v, kind, cores, sample, secs = cpu_reference_sample(path, a.cpu_sample_nets, a.cpu_sample_iters)
            if s >= a.warmup:
                vals.append(v); total += secs
        par = cpu_parallel_sample(path, a.cpu_sample_iters)      # extra: the same CPU algorithm on all host threads
        par_full = cpu_parallel_full(path)                        # extra: the device router's policy on all host threads, to legality
        full = None if a.no_full_reference else cpu_reference_full(path)   # extra: ONE complete routing by the unmodified reference
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "nets_routed_per_sec", "value": v, "unit": "nets/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total / max(a.steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, p),
        "parallelism": "the reference's serial router, 1 host thread (bounded sample per step: cpu_baseline.sample)",
        "cpu_baseline": {"value": v, "unit": "nets/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "nets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if par:
        line["cpu_parallel_baseline"] = par
    if par_full:
        line["cpu_parallel_congested_only_full"] = par_full
    if full:
        line["reference_full_run"] = full
        line["reference_full_run"]["nets_per_s"] = a.nets * full["iterations"] / max(full["route_time_s"], 1e-9)
    print(json.dumps(line))


def run_ours(a):
    if a.workload != "cfg4":
        if int(os.environ.get("RANK", "0")) == 0:
            run_big_ours(a)
        return
    import numpy as np
    import torch
    from parallel_eda_b200 import distributed, pathfinder, pfio, router
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the router has no CPU path")
    backend = router.load_library().pf_backend_name()
    if not backend.startswith(b"cuda:"):
        raise SystemExit("bench.py measures the CUDA library only (loaded: %s)" % backend.decode())
    comm = distributed.init_from_env("nccl")
    rank = comm.rank if comm else 0
    world = comm.world if comm else 1
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    p = router.generate_grid_problem(nx=a.grid, ny=a.grid, W=a.width, num_nets=a.nets)     # full host copy: CPU legs and the independent checks
    # what the router itself gets: the nets (host buffers) and the generator's parameters — the rr graph is built ON the device
    # (pf_router_create_generated; --upload-graph restores the 1.4 GB host -> device path of round 1)
    nets, gen = (p, None) if a.upload_graph else router.generate_grid_nets(nx=a.grid, ny=a.grid, W=a.width, num_nets=a.nets)
    cfg = router.default_config(device=local, rank=rank, nranks=world, max_batch=a.max_batch, pop_slack=a.pop_slack,
                                inflight_div=a.inflight_div, num_slots=a.slots, min_slots=a.min_slots,
                                **{k: int(v) for k, v in (kv.split("=") for kv in a.cfg.split())})
    R = comm.create_router(nets, cfg, generated=gen) if comm else router.Router(nets, cfg, generated=gen)   # N > 1: incl. the one-off pf_comm_export / pf_comm_init
    R.timing(reset=True)

    def route(Rx):
        if a.step_api:
            return pathfinder.route(Rx, comm=comm, sync_rounds=a.sync_rounds)
        return pathfinder.run(Rx, comm=comm)       # pf_route_run: the loop, and for N > 1 the occupancy exchange, inside the library

    def one_step():
        R.reset()
        if comm:
            comm.barrier()
        torch.cuda.synchronize()
        R.timer_start()
        rep = route(R)
        ms = R.timer_stop()
        torch.cuda.synchronize()
        if comm:
            comm.barrier()
            ms = comm.all_reduce_max(ms)
        wl, _ = R.total_wirelength()               # after the timed region: reported, not measured
        if comm:
            wl = int(comm.all_reduce_scalar(wl))
            if not a.step_api:                     # per-rank counters of pf_route_run -> job totals (outside the timed region)
                rep.per_rank_nets = rep.nets_routed
                rep.nets_routed = int(comm.all_reduce_scalar(rep.nets_routed))
        rep.wirelength = int(wl)
        return rep, ms

    for _ in range(a.warmup):
        one_step()
    R.timing(reset=True)
    reps, times = [], []
    with ClockSampler(local) as clk:
        for _ in range(a.steps):
            rep, ms = one_step()
            reps.append(rep); times.append(ms)
    tm = R.timing(reset=True)
    total_ms = sum(times)
    nets_routed = sum(r.nets_routed for r in reps)
    value = nets_routed / (total_ms * 1e-3)
    assert all(r.success for r in reps), "routing did not converge"

    # roofline of the dominant kernel: algorithmic bytes / CUDA-event kernel time
    visits = sum(r.edge_visits for r in reps); pops = sum(r.heap_pops for r in reps); pushes = sum(r.heap_pushes for r in reps)
    alg_bytes = 36.0 * visits + 28.0 * pops + 20.0 * pushes                       # this rank's launches
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (tm.route_kernel_ms * 1e-3) / 1e9 if tm.route_kernel_ms > 0 else 0.0
    launches = int(tm.route_launches + tm.update_launches + tm.aux_launches)
    # DRAM traffic of the dominant launch from the committed `ncu --set full` capture (iteration-1 launch of this very
    # workload on 1 GPU: profiles/r01_ncu_full_pf_route_kernel.json, dram__bytes_read.sum + dram__bytes_write.sum);
    # that launch's algorithmic bytes are 3.75e9 (7.78e7 visits, 5.12e6 pops, 4.00e7 label writes); 3.2e9 of the traffic
    # are deliberate L2 prefetches of edge rows (10.4e9 without them, 2.5 % slower)
    # dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch: read from the summary of the newest committed
    # `ncu --set full` capture of THIS workload (tools/ncu_traffic.py writes profiles/*_traffic.json next to the raw csv);
    # null when no capture of this configuration exists
    traffic, traffic_detail = None, None
    if world == 1:
        import glob
        for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
            try:
                t = json.load(open(f))
                if (t.get("grid"), t.get("nets"), t.get("width")) == (a.grid, a.nets, a.width):
                    traffic_detail = dict(t, source=os.path.relpath(f, ROOT))
                    traffic = float(t["dram_bytes_per_launch"])
                    break
            except Exception:
                pass

    # end to end through the public API with host buffers (N GPUs)
    e2e = None
    if not a.no_e2e:
        e2e_phases = []

        def e2e_step():
            if comm:
                comm.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            # H2D of the problem: the nets; the rr graph is generated on every rank's device (--upload-graph: flatten + 1.4 GB H2D,
            # N GPUs: rank 0 uploads, the others receive over NVLink)
            R2 = comm.create_router(nets, cfg, generated=gen) if comm else router.Router(nets, cfg, generated=gen)
            t1 = time.perf_counter()
            rep = route(R2)
            if comm and not a.step_api:
                rep.nets_routed = int(comm.all_reduce_scalar(rep.nets_routed))
            t2 = time.perf_counter()
            res = R2.result()                          # D2H of traces, delays, occupancy
            t3 = time.perf_counter()
            t = R2.timing()
            R2.close()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            e2e_phases.append((t1 - t0, t2 - t1, t3 - t2, dt - (t3 - t0)))
            if comm:
                dt = comm.all_reduce_max(dt)
            return rep, dt, t.h2d_bytes, t.d2h_bytes, res
        e2e_step()
        acc_n = acc_t = 0.0
        hb = db = 0
        _res = None
        for _ in range(max(1, min(a.steps, 3))):
            _res = None                                # a caller consumes one result before it asks for the next (pf_result_free): the
            rep, dt, hb, db, _res = e2e_step()         # library then hands the same pinned arrays out again instead of fresh memory
            acc_n += rep.nets_routed; acc_t += dt
        # independent look at the routing the public API handed back (outside the timed region): occupancy recomputed
        # from the traces equals the reported one, every sink is reached, sampled trace pairs are real rr edges
        check = None
        if world == 1:
            from parallel_eda_b200 import check_route
            check = check_route.check_route_fast(p, _res)
            check["wirelength"] = int(_res.total_wirelength)
            # and the full check_route (every net, every edge, every sink) on the device, from the traces alone
            check["device_check_route"] = R.check_route(_res)
            assert check["device_check_route"]["ok"] == 1 and check["device_check_route"]["overused_nodes"] == 0
        e2e = {"value": acc_n / acc_t, "unit": "nets/s", "h2d_bytes_per_step": int(hb), "d2h_bytes_per_step": int(db),
               "s_per_step": acc_t / max(1, min(a.steps, 3)),
               "phases_s": dict(zip(("create_upload", "route", "result_download", "destroy"), [round(x, 4) for x in e2e_phases[-1]])),
               "graph": "uploaded from host arrays" if a.upload_graph else "generated on the device from the fabric parameters (pf_router_create_generated); H2D = nets, boxes, tables",
               "result_check": check}

    cpu, cpu_par, cpu_par_full = None, None, None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "bench.pfp")
            pfio.write_problem(path, p)
            v, kind, cores, sample, secs = cpu_reference_sample(path, a.cpu_sample_nets, a.cpu_sample_iters)
            cpu = {"value": v, "unit": "nets/s", "cores": cores, "kind": kind, "sample": sample, "sample_seconds": secs}
            cpu_par = cpu_parallel_sample(path, a.cpu_sample_iters)
            cpu_par_full = cpu_parallel_full(path)

    if rank == 0:
        out = {
            "metric": "nets_routed_per_sec", "value": value, "unit": "nets/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, p),
            "parallelism": ("nets sharded over %d GPU(s) in spatial stripes (stripe-interior nets, then cut-crossing nets); occupancy event logs exchanged 2x per iteration "
                                       "%s" % (world, "by torch.distributed all-gather (step API)" if a.step_api else "device-side over NVLink peer memory (pf_comm_exchange), one host sync per iteration")) if world > 1 else "1 GPU, one host sync per PathFinder iteration (pf_route_run)",
            "route": {"iterations": [r.iterations for r in reps], "nets_routed_per_step": nets_routed / a.steps,
                      "route_time_s": total_ms * 1e-3 / a.steps, "legal": True,
                      "wirelength": [r.wirelength for r in reps], "overused_per_iteration": reps[-1].overused},
            "roofline": {"bound": "hbm", "kernel": "pf_route_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_detail": traffic_detail, "peak_source": peak_src,
                         "algorithmic_bytes": "36 B/edge visit + 28 B/pop + 20 B/label write (SURVEY.md §8d)",
                         "kernel_ms_per_step": tm.route_kernel_ms / a.steps, "kernel_launches_per_step": tm.route_launches / a.steps},
            "gpu_launches": launches,
            "clocks": clk.summary(),
        }
        if e2e:
            out["e2e"] = e2e
        if cpu:
            out["cpu_baseline"] = cpu
        if cpu_par:
            out["cpu_parallel_baseline"] = cpu_par     # extra to the contract: the same CPU algorithm on all host threads
        if cpu_par_full:
            out["cpu_parallel_congested_only_full"] = cpu_par_full   # the device router's policy on all host threads, to legality
        print(json.dumps(out))
    R.close()
    if comm:
        import torch.distributed as dist
        dist.destroy_process_group()



def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
