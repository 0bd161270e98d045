"""Randomised runs of the DEVICE code on the warp emulator (no GPU): small generated fabrics, timing off / timing-driven (all criticalities 1) /
breadth-first, 1..32 warps; every result goes through the independent checker (legality, occupancy from the traces, Elmore delays
from scratch); where the device code gives up, the serial oracle is asked whether the problem was routable at all.
usage: python tools/fuzz_emu.py [seed] [seconds]"""
import sys, os, time, random, subprocess, tempfile
sys.path.insert(0, "/root/repo")
os.environ["PF_ALLOW_EMULATOR"] = "1"
import numpy as np
from parallel_eda_b200 import router, pfio, check_route
emu = "/root/repo/tests/emu/_build/libpf_router_emu.so"
lib = router.load_library(emu)
orc = "/root/repo/oracle/_build/pf_oracle_cli"
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
t_end = time.time() + float(sys.argv[2]) if len(sys.argv) > 2 else time.time() + 300
n = bad = 0
while time.time() < t_end:
    nx = rng.randint(6, 20); ny = rng.randint(6, 20); W = rng.choice([8, 12, 16, 20, 30]); nets = rng.randint(30, 60 + nx * ny); sinks = rng.randint(1, 7)
    seed = rng.randint(1, 10**6); bf = rng.random() < 0.25
    if bf: nets = min(nets, 50); sinks = min(sinks, 3); W = max(W, 16)
    else: nets = min(nets, 160)
    try:
        p = router.generate_grid_problem(lib_path=emu, nx=nx, ny=ny, W=W, num_nets=nets, sinks_per_net=sinks, seed=seed, window=rng.choice([3, 5, 8, 30]))
    except Exception as e:
        print("gen failed", nx, ny, W, nets, sinks, seed, e); continue
    td = (not bf) and rng.random() < 0.5
    if bf:
        p.opts["router_algorithm"] = 1; p.opts["first_iter_pres_fac"] = 0.0
    if td:
        p.opts["timing_analysis_enabled"] = 1      # no analysis callback: every criticality stays at 1 (max_criticality 0.99 applies)
    slots = rng.choice([1, 2, 8, 32])
    cfg = router.default_config(lib, num_slots=slots, big_slots=rng.choice([1, 2]))
    tag = (nx, ny, W, nets, sinks, seed, "bf" if bf else ("td" if td else "nt"), slots); print("start", tag, flush=True); t0 = time.time()
    try:
        r = router.try_timing_driven_route(p, cfg, lib_path=emu)
    except router.RouterError as e:
        print("ROUTER ERROR", tag, e); bad += 1; continue
    n += 1; print("run", n, tag, "success", r.success, "iters", r.iterations, "%.1fs" % (time.time() - t0), flush=True)
    try:
        m = check_route.check_route(p, r, check_delays=not bf, require_legal=bool(r.success))
    except AssertionError as e:
        print("CHECK FAILED", tag, str(e)[:300]); bad += 1; continue
    if r.success and m["overused"] != 0:
        print("SUCCESS BUT OVERUSED", tag, m); bad += 1
    if not r.success:
        # does the serial reference restatement succeed where we do not?
        with tempfile.TemporaryDirectory() as d:
            pp = os.path.join(d, "p.pfp"); pfio.write_problem(pp, p)
            o = subprocess.run([orc, pp], capture_output=True, text=True)
            ok = "success=1" in o.stderr
        print("no success", tag, "iterations", r.iterations, "oracle success" if ok else "oracle fails too")
print("runs", n, "anomalies", bad)
