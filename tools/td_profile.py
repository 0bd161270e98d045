"""Per-iteration timing of a timing-driven fixture with the device STA in the loop (step API).
usage: python tools/td_profile.py [name]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parallel_eda_b200 import pfio, router
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
name = sys.argv[1] if len(sys.argv) > 1 else "mid_w200"
import json
extra = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
quiet = len(sys.argv) > 3
p = pfio.read_problem(os.path.join(G, name + ".pfp.xz")); p.opts["timing_analysis_enabled"] = 1; p.opts["max_router_iterations"] = 150
g = pfio.read_timing_graph(os.path.join(G, name + ".pftg.xz"))
for rep in range(2):
    R = router.Router(p, router.default_config(**extra)); S = router.Sta(g, p); o = p.opts
    pres = float(o["first_iter_pres_fac"]); rows = []
    t_all = time.perf_counter()
    for it in range(1, 151):
        R.timing(reset=True)
        t0 = time.perf_counter(); st = R.route_iteration(pres); t1 = time.perf_counter()
        tk = R.timing(reset=True)
        R.reserve_locally_used_opins(pres, it != 1)
        pres, acc = (float(o["initial_pres_fac"]), 0.0) if it == 1 else (pres * float(o["pres_fac_mult"]), float(o["acc_fac"]))
        over = R.pathfinder_update_cost(acc); t2 = time.perf_counter()
        if over == 0: rows.append((it, st.nets_routed, over, t1 - t0, tk.route_kernel_ms, tk.route_launches, 0.0)); break
        cpd = S.analyze_device(R.comm_net_delay_ptr(), R.comm_crit_ptr()); t3 = time.perf_counter()
        rows.append((it, st.nets_routed, over, t1 - t0, tk.route_kernel_ms, tk.route_launches, t3 - t2))
    total = time.perf_counter() - t_all
    if rep:
        print("%s %s: %d iterations, %.1f ms, slowest iteration %.1f ms" % (name, extra, len(rows), total * 1e3, max(r[3] for r in rows) * 1e3))
        for r in ([] if quiet else rows): print("  it %3d nets %5d over %5d | route call %.2f ms (kernel %.2f ms, %d launches) sta %.2f ms" % (r[0], r[1], r[2], r[3] * 1e3, r[4], r[5], r[6] * 1e3))
    S.close(); R.close()
