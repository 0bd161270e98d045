"""Phase timing of the host-buffer call (pf_try_timing_driven_route, verbose): create / route / result.
usage: python tools/e2e_phases.py [grid] [nets] [repeats]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parallel_eda_b200 import router
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 400
nets = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
p = router.generate_grid_problem(nx=grid, ny=grid, W=100, num_nets=nets)
for k in range(reps):
    t0 = time.perf_counter()
    r = router.try_timing_driven_route(p, router.default_config(verbose=1 if k else 0))
    print("call %d: %.3f s, success %d, iterations %d" % (k, time.perf_counter() - t0, r.success, r.iterations), flush=True)
