"""Phase timing of the host-buffer call (pf_try_timing_driven_route, verbose): create / route / result.
usage: python tools/e2e_phases.py [grid] [nets] [repeats] [gen]      gen: the rr graph is built on the device (bench.py's e2e path)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parallel_eda_b200 import router
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 400
nets = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
if len(sys.argv) > 4:
    from parallel_eda_b200 import pathfinder
    nets_only, gen = router.generate_grid_nets(nx=grid, ny=grid, W=100, num_nets=nets)
    for k in range(reps):
        t0 = time.perf_counter()
        R = router.Router(nets_only, router.default_config(verbose=1 if k else 0), generated=gen); t1 = time.perf_counter()
        rep = pathfinder.run(R); t2 = time.perf_counter()
        res = R.result(); t3 = time.perf_counter()
        R.close()
        print("call %d: create %.1f ms, route %.1f ms, result %.1f ms, close %.1f ms, iterations %d" % (
            k, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (time.perf_counter() - t3) * 1e3, rep.iterations), flush=True)
    sys.exit(0)
p = router.generate_grid_problem(nx=grid, ny=grid, W=100, num_nets=nets)
for k in range(reps):
    t0 = time.perf_counter()
    r = router.try_timing_driven_route(p, router.default_config(verbose=1 if k else 0))
    print("call %d: %.3f s, success %d, iterations %d" % (k, time.perf_counter() - t0, r.success, r.iterations), flush=True)
