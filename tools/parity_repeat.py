"""Run-to-run spread of the concurrent router on the parity fixtures (the GPU schedule is not deterministic: which of two nets
in flight commits first depends on timing).  usage: python tools/parity_repeat.py REPS fixture[,fixture..] ["k=v k=v" ...]
Each config string is one variant of router.default_config; prints one line per run and a min / median / max summary."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parallel_eda_b200 import pfio, router

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
reps = int(sys.argv[1])
names = sys.argv[2].split(",")
variants = sys.argv[3:] or [""]
for name in names:
    p = pfio.read_problem(os.path.join(G, name + ".pfp.xz"))
    g = pfio.read_result(os.path.join(G, name + ".pfr.xz"))
    w = g.iter_crit[-1]
    for v in variants:
        kw = {k: (float(x) if "." in x else int(x)) for k, x in (kv.split("=") for kv in v.split())}
        rows = []
        for _ in range(reps):
            r = router.try_timing_driven_route(p, router.default_config(**kw), sta=router.replay_sta(g))
            rows.append((int(r.success), int(r.iterations), r.total_wirelength / g.total_wirelength,
                         float((w * r.net_delay).sum()) / float((w * g.net_delay).sum())))
        a = np.array(rows)
        print("%-9s [%s] ref %d it | ok %d/%d | it %s | wl min %.3f med %.3f max %.3f | td min %.3f med %.3f max %.3f" % (
            name, v, g.iterations, int(a[:, 0].sum()), reps, sorted(int(x) for x in a[:, 1]),
            a[:, 2].min(), np.median(a[:, 2]), a[:, 2].max(), a[:, 3].min(), np.median(a[:, 3]), a[:, 3].max()), flush=True)
