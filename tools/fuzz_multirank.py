"""Randomised multi-rank runs of the device code on the warp emulator (no GPU): 2..6 emulated ranks under torchrun (gloo only
bootstraps the transport), random generated fabrics, the in-library exchange protocol over shared memory.  Every run checks
that all ranks end with the same occupancy, that every routed net was routed by exactly one rank, and puts the union of the
ranks' traces through the independent checker.  Round 2: 105 runs, no replica divergence, no unrouted or doubly routed net,
no deadlock (one run ended with exit code 1 after rank 0 had printed a consistent result; 18 repeats of the same inputs were clean).
usage: python tools/fuzz_multirank.py [seed] [seconds]"""
import json
import os
import random
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch
from parallel_eda_b200 import pfio, router, pathfinder, distributed, check_route
comm = distributed.init_from_env("gloo")
lib = os.path.join(%(root)r, "tests", "emu", "_build", "libpf_router_emu.so")
L = router.load_library(lib)
nx, ny, W, nets, sinks, seed, window = [int(x) for x in sys.argv[1:8]]
p = router.generate_grid_problem(lib_path=lib, nx=nx, ny=ny, W=W, num_nets=nets, sinks_per_net=sinks, seed=seed, window=window)
p.opts["max_router_iterations"] = 12          # unroutable draws cost 10 s per iteration on the emulator
cfg = router.default_config(L, num_slots=4, big_slots=2, rank=comm.rank, nranks=comm.world)
R = comm.create_router(p, cfg, lib_path=lib)
rep = pathfinder.run(R, comm=comm)
res = R.result()
occ = torch.from_numpy(res.occ.astype(np.int64)); ref = occ.clone(); torch.distributed.broadcast(ref, 0)
assert torch.equal(occ, ref), "occupancy replicas differ"
own = [i for i in p.routed_nets() if res.trace_ptr[i + 1] > res.trace_ptr[i]]
parts = [None] * comm.world
torch.distributed.all_gather_object(parts, (own, [res.net_trace(int(i)) for i in own]))
if comm.rank == 0:
    seen = {}
    for o, tr in parts:
        for i, t in zip(o, tr):
            assert i not in seen, "net routed by two ranks"
            seen[int(i)] = t
    assert sorted(seen) == [int(i) for i in p.routed_nets()], "some net routed by nobody"
    tp = [0]; tn = []; ts = []
    for i in range(p.num_nets):
        if i in seen: tn.append(seen[i][0]); ts.append(seen[i][1])
        tp.append(tp[-1] + (len(seen[i][0]) if i in seen else 0))
    full = pfio.Result(int(rep.success), rep.iterations, 0, 0, np.array(tp, np.int32), np.concatenate(tn), np.concatenate(ts), res.net_delay, res.occ, res.iter_stats)
    m = check_route.check_route(p, full, check_delays=False, require_legal=bool(rep.success))
    print(json.dumps({"ranks": comm.world, "success": bool(rep.success), "iters": int(rep.iterations), "overused": int(m["overused"]), "routed_by": [len(o) for o, _ in parts]}), flush=True)
'''


def main():
    rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    t_end = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 300)
    env = dict(os.environ, PF_ALLOW_EMULATOR="1")
    k = bad = 0
    with tempfile.TemporaryDirectory() as d:
        w = os.path.join(d, "worker.py")
        open(w, "w").write(WORKER % {"root": ROOT})
        while time.time() < t_end:
            n = rng.choice([2, 3, 4, 5, 6])
            args = [rng.randint(8, 40), rng.randint(6, 16), rng.choice([16, 20, 30]), rng.randint(40, 220), rng.randint(1, 4), rng.randint(1, 10**6), rng.choice([3, 6, 12])]
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % n, "--master-addr", "127.0.0.1",
                   "--master-port", str(30000 + k % 90), w] + [str(a) for a in args]
            k += 1
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
            except subprocess.TimeoutExpired:
                print("TIMEOUT", n, args, flush=True); bad += 1
                continue
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode != 0 or not line:
                bad += 1
                print("FAIL", n, args, "rc", r.returncode, line[-1] if line else "", flush=True)
                print("\n".join([l for l in (r.stdout + r.stderr).splitlines() if "rror" in l or "assert" in l.lower()][:15]), flush=True)
            else:
                print(n, args, line[-1], flush=True)
    print("runs", k, "failures", bad)


if __name__ == "__main__":
    main()
