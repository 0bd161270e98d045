"""Throughput of the native .route writer / reader (include/pf_text.h) on a generated fabric — no GPU needed.

    python tools/text_bench.py 400 200000      # BASELINE configs[4]: 224 MB file, DESIGN.md §4.9 numbers

Generates the grid, lets the CPU oracle (test infrastructure) route ONE PathFinder iteration to have realistic traces
(legality does not matter for the text), then times pf_route_write / pf_route_read three times.  PF_TEXT_THREADS=1 gives
the single-thread numbers.  For the reference's own print_route on the same routing:
    oracle/_ref/vpr_ref inject /tmp/tb.pfp --max_iters 1 --route-file /tmp/ref.route     (prints its print_route time)
"""
import sys, time, subprocess, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parallel_eda_b200 import pfio, router, textio
nx, nets = int(sys.argv[1]), int(sys.argv[2])
t=time.time(); p = router.generate_grid_problem(nx=nx, ny=nx, W=100, num_nets=nets, sinks_per_net=3, seed=1); print("gen %.1fs N=%d"%(time.time()-t,p.num_nodes))
t=time.time(); pfio.write_problem("/tmp/tb.pfp", p); print("write problem %.1fs"%(time.time()-t))
t=time.time(); subprocess.run([os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_build", "pf_oracle_cli"),"/tmp/tb.pfp","--result","/tmp/tb.pfr","--max_iters","1"],capture_output=True); print("oracle 1 iter %.1fs"%(time.time()-t))
r = pfio.read_result("/tmp/tb.pfr")
n = textio.synthetic_names(p)
for k in range(3):
    t=time.time(); textio.write_route("/tmp/tb.route", p, n, r); tw=time.time()-t
    sz=os.path.getsize("/tmp/tb.route")
    t=time.time(); q = textio.read_route("/tmp/tb.route", p); tr=time.time()-t
    print("route file %.1f MB, %d trace elements: write %.3fs (%.0f MB/s)  read %.3fs (%.0f MB/s)"%(sz/1e6,len(r.trace_node),tw,sz/1e6/tw,tr,sz/1e6/tr))
import numpy as np
print(np.array_equal(q.trace_node,r.trace_node), np.array_equal(q.trace_switch,r.trace_switch))
