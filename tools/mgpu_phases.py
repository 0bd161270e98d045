"""Where a multi-GPU PathFinder iteration spends its time (wall clock around each C-ABI / NCCL call).
launch: python -m torch.distributed.run --nproc-per-node N tools/mgpu_phases.py [grid] [nets]"""
import sys, os, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parallel_eda_b200 import router, distributed
comm = distributed.init_from_env()
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 400
nets = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
p = router.generate_grid_problem(nx=grid, ny=grid, W=100, num_nets=nets)
local = int(os.environ.get("LOCAL_RANK", "0"))
R = router.Router(p, router.default_config(device=local, rank=comm.rank, nranks=comm.world))
o = p.opts
for rep in range(2):
    if rep: R.reset()
    T = collections.OrderedDict()
    def timed(name, fn):
        t = time.perf_counter(); v = fn(); T[name] = T.get(name, 0.0) + time.perf_counter() - t; return v
    pres = float(o["first_iter_pres_fac"]); rows = []
    comm.barrier(); t_all = time.perf_counter()
    for it in range(1, 51):
        T.clear()
        timed("begin", lambda: R.iteration_begin(None))
        n = 0
        for part in range(2):
            st = timed("route%d" % part, lambda: R.iteration_route_part(pres, part, 2)); n += st.nets_routed
            timed("sync%d" % part, lambda: comm.sync_occupancy(R))
        timed("opins", lambda: R.reserve_locally_used_opins(pres, it != 1))
        new_pres, acc = (float(o["initial_pres_fac"]), 0.0) if it == 1 else (pres * float(o["pres_fac_mult"]), float(o["acc_fac"]))
        over = timed("update", lambda: R.pathfinder_update_cost(acc))
        pres = new_pres
        rows.append((it, n, over, dict(T)))
        if over == 0: break
    total = time.perf_counter() - t_all
    if rep:
        for rk in range(comm.world):
            comm.barrier()
            if rk == comm.rank:
                print("rank %d: %d iterations, %.1f ms" % (rk, len(rows), total * 1e3))
                for it, n, over, t in rows:
                    print("  it %2d nets %6d over %6d | " % (it, n, over) + " ".join("%s %.2f" % (k, v * 1e3) for k, v in t.items()), flush=True)
