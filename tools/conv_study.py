"""Convergence study on the CPU warp emulator (tests/emu): iterations / wirelength / weighted delay of the device
router against the reference's golden routing, per fixture, for a given pf_config.

    python tools/conv_study.py [--fixtures toy_w64,hub_w90] [--mode td|nt|both] [--reps 1] key=value ...

key=value pairs are pf_config fields (num_slots=64 inflight_div=16 ...).  max_router_iterations stays at the
reference's 50 unless --max-iters is given."""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallel_eda_b200 import pfio, router  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fixtures", default="toy_w64,hub_w90,het_w70,mix_w70,heq_w70")
    ap.add_argument("--mode", default="both")
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--max-iters", type=int, default=0)
    ap.add_argument("--lib", default="")
    ap.add_argument("kv", nargs="*")
    a = ap.parse_args()
    lib = a.lib
    if not lib:
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
        lib = os.path.join(ROOT, "tests", "emu", "_build", "libpf_router_emu.so")
        os.environ.setdefault("PF_ALLOW_EMULATOR", "1")
    kw = {}
    for s in a.kv:
        k, v = s.split("=")
        kw[k] = float(v) if "." in v else int(v)
    kw.setdefault("num_slots", 64)
    L = router.load_library(lib)
    tot_it = tot_ref = 0
    for name in a.fixtures.split(","):
        for timing in ([1, 0] if a.mode == "both" else [1] if a.mode == "td" else [0]):
            gp = os.path.join(G, name + (".pfr.xz" if timing else "_nt.pfr.xz"))
            if not os.path.exists(gp):
                continue
            p = pfio.read_problem(os.path.join(G, name + ".pfp.xz"))
            p.opts["timing_analysis_enabled"] = timing
            if a.max_iters:
                p.opts["max_router_iterations"] = a.max_iters
            g = pfio.read_result(gp)
            for rep in range(a.reps):
                cfg = router.default_config(L, **kw)
                t0 = time.time()
                r = router.try_timing_driven_route(p, cfg, sta=router.replay_sta(g) if timing else None, lib_path=lib)
                w = g.iter_crit[-1] if timing else None
                wd = float((w * r.net_delay).sum()) / float((w * g.net_delay).sum()) if timing else float("nan")
                over = [int(x) for x in r.iter_stats["overused_nodes"]]
                nets = [int(x) for x in r.iter_stats["nets_routed"]]
                print("%-9s %s ok=%d it=%3d (ref %2d) wl=%.3f wd=%.3f nets=%d  over=%s  %.1fs" % (
                    name, "td" if timing else "nt", r.success, r.iterations, g.iterations, r.total_wirelength / g.total_wirelength,
                    wd, sum(nets), over[:4] + ["..."] + over[-6:], time.time() - t0), flush=True)
                tot_it += r.iterations if r.success else 999
                tot_ref += g.iterations
    print("total iterations %d (reference %d)" % (tot_it, tot_ref))


if __name__ == "__main__":
    main()
