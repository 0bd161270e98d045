"""Small driver for ncu captures: one problem, a few PathFinder iterations through the step API.
usage: python tools/prof_run.py [grid] [nets] [iterations]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parallel_eda_b200 import router
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 100
nets = int(sys.argv[2]) if len(sys.argv) > 2 else grid * grid * 5 // 4
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
p = router.generate_grid_problem(nx=grid, ny=grid, W=100, num_nets=nets)
R = router.Router(p, router.default_config(verbose=1))
o = p.opts
pres = float(o["first_iter_pres_fac"])
for it in range(1, iters + 1):
    t0 = R.timing(reset=True)
    st = R.route_iteration(pres)
    tk = R.timing(reset=True)
    pres, acc = (float(o["initial_pres_fac"]), 0.0) if it == 1 else (pres * float(o["pres_fac_mult"]), float(o["acc_fac"]))
    over = R.pathfinder_update_cost(acc)
    print("iter", it, "nets", st.nets_routed, "overused", over, "pops", st.heap_pops, "pushes", st.heap_pushes, "visits", st.edge_visits, "route_kernel_ms %.2f" % tk.route_kernel_ms, "launches", tk.route_launches, flush=True)
    if over == 0: break
t = R.timing()
print("route kernel ms", t.route_kernel_ms, "launches", t.route_launches)
