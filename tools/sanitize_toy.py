"""Small end-to-end exercise for compute-sanitizer: timing-driven with the device STA, breadth-first, check_route."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parallel_eda_b200 import pfio, router
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
p = pfio.read_problem(os.path.join(G, "toy_w64.pfp.xz")); p.opts["timing_analysis_enabled"] = 1; p.opts["max_router_iterations"] = 150
g = pfio.read_timing_graph(os.path.join(G, "toy_w64.pftg.xz"))
r = router.try_timing_driven_route(p, router.default_config(num_slots=64), timing_graph=g)
R = router.Router(p, router.default_config(num_slots=64))
print("timing-driven + device STA:", r.success, r.iterations, R.check_route(r)["ok"])
R.close()
b = pfio.read_problem(os.path.join(G, "toy_w64_bf.pfp.xz"))
rb = router.try_timing_driven_route(b, router.default_config(num_slots=64))
print("breadth-first:", rb.success, rb.iterations)
