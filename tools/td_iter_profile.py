"""Per-iteration profile of a BASELINE-scale timing-driven fixture (tests/golden/big) through pf_route_run with the device STA:
device time of every phase (PF_PHASES), nets / pops / edge visits / label writes per iteration, nets moved to the big slots.
usage: python tools/td_iter_profile.py sv0_w220|bgm_w260 [key=value pf_config fields ...]   (stderr carries the tables)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parallel_eda_b200 import pathfinder, pfio, router  # noqa: E402

B = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "big")
name = sys.argv[1] if len(sys.argv) > 1 else "sv0_w220"
kw = {k: (float(v) if "." in v else int(v)) for k, v in (s.split("=") for s in sys.argv[2:])}
p = pfio.read_problem(os.path.join(B, name + ".pfp.xz"))
g = pfio.read_timing_graph(os.path.join(B, name + ".pftg.xz"))
cfg = router.default_config(verbose=1, **kw)
R = router.Router(p, cfg)
S = router.Sta(g, p, cfg)
os.environ["PF_PHASES"] = "1"
R.timer_start()
t0 = time.perf_counter()
rep = pathfinder.run(R, dsta=S)
ms = R.timer_stop()
t = R.timing()
sys.stderr.write("%s %s: success %s, %d iterations, %.1f ms on the device (%.1f ms wall), route kernels %.1f ms in %d launches, other kernels %.1f ms\n" % (
    name, kw, rep.success, rep.iterations, ms, (time.perf_counter() - t0) * 1e3, t.route_kernel_ms, t.route_launches, t.update_kernel_ms + t.aux_kernel_ms))
ok, it, st = True, rep.iterations, None
sys.stderr.write("nets per iteration   %s\noverused             %s\n" % (rep.per_iter_nets, rep.overused))
S.close(); R.close()
