"""Where a multi-GPU PathFinder iteration spends its DEVICE time, per rank and per phase (CUDA events on the router's stream):
begin | route part 0 | exchange 0 | route part 1 | exchange 1 | OPIN reservation + cost update + select — and the gap in front
of every iteration (the one read of the control block, the host's decisions, the launches).  An exchange column contains the
wait for the slowest peer: load imbalance shows up there, not in the route columns.

launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/mgpu_phases.py [grid] [nets] 2> phases.txt"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from parallel_eda_b200 import distributed, pathfinder, router  # noqa: E402

comm = distributed.init_from_env()
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 400
nets = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
local = int(os.environ.get("LOCAL_RANK", "0"))
p, gen = router.generate_grid_nets(nx=grid, ny=grid, W=100, num_nets=nets)     # the rr graph is built on every rank's device
cfg = router.default_config(device=local, rank=comm.rank if comm else 0, nranks=comm.world if comm else 1)
R = comm.create_router(p, cfg, generated=gen) if comm else router.Router(p, cfg, generated=gen)
for rep in range(3):
    R.reset()
    if comm:
        comm.barrier()
    torch.cuda.synchronize()
    if rep == 2:
        os.environ["PF_PHASES"] = "1"
    R.timer_start()
    t0 = time.perf_counter()
    r = pathfinder.run(R, comm=comm)
    ms = R.timer_stop()
    wall = (time.perf_counter() - t0) * 1e3
    if rep == 2:
        t = R.timing(reset=True)
        sys.stderr.write("PF_PHASES rank %d summary: %d iterations, step %.3f ms on the device (%.3f ms wall), route kernels %.3f ms in %d launches, "
                         "other kernels %.3f ms, nets routed by this rank per iteration %s\n" % (
                             cfg.rank, r.iterations, ms, wall, t.route_kernel_ms, t.route_launches, t.update_kernel_ms + t.aux_kernel_ms, r.per_iter_nets))
    else:
        R.timing(reset=True)
R.close()
