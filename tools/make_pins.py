"""Regenerates the one-warp pins (tests/golden/single_warp_*.json): with one warp the device code is deterministic, and the
sm_100a build must reproduce, cookie for cookie, what the same source computes on the CPU warp emulator.  Run after any
change to the device code or to the re-route policy; the GPU tests compare against these files."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallel_eda_b200 import pfio, router  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
LIB = os.path.join(ROOT, "tests", "emu", "_build", "libpf_router_emu.so")
L = router.load_library(LIB)


def one(name, timing=0, **kw):
    p = pfio.read_problem(os.path.join(G, name + ".pfp.xz"))
    if not name.endswith("_bf"):
        p.opts["timing_analysis_enabled"] = timing
    r = router.try_timing_driven_route(p, router.default_config(L, num_slots=1, big_slots=1, pop_slack=0.0, max_batch=1, **kw), lib_path=LIB)
    assert r.success == 1
    return {"serial_num": int(r.serial_num), "total_wirelength": int(r.total_wirelength), "iterations": int(r.iterations)}


toy = one("toy_w64")
toy["what"] = "toy_w64, timing off, one warp, pop_slack 0, max_batch 1: the deterministic routing of the device code (emulated on CPU == sm_100a build)"
json.dump(toy, open(os.path.join(G, "single_warp_toy.json"), "w"))
het = {"serial_policy": one("het_w70", reroute_all_iters=-1), "default_policy": one("het_w70"),
       "note": "het_w70, timing off, one warp (num_slots=1, big_slots=1, pop_slack=0, max_batch=1): deterministic; written by the emulated device code (tools/make_pins.py), the GPU must reproduce it"}
json.dump(het, open(os.path.join(G, "single_warp_het.json"), "w"), indent=1)
bf = one("toy_w64_bf")
bf["note"] = "toy_w64_bf.pfp routed by the device code on the CPU warp emulator with one slot (tools/make_pins.py; tests/test_gpu_breadth_first.py)"
json.dump(bf, open(os.path.join(G, "single_warp_toy_bf.json"), "w"), indent=1)
print(toy, het, bf)
