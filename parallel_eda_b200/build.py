"""Builds libpf_router.so (CUDA, sm_100a) in-tree.  nvcc cross-compiles without a GPU.

    python -m parallel_eda_b200.build          # or: from parallel_eda_b200.build import build; build()

Flags: -gencode arch=compute_100a,code=sm_100a (B200 only, no fallback arch), -lineinfo so ncu's
source page maps to pf_device.cuh, -fmad=false so the cost arithmetic rounds like the CPU
reference (the kernels are memory-latency bound; FMA contraction buys nothing here).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpf_router.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build(force: bool = False, verbose: bool = False, variant: str = "") -> str:
    """variant "" is the product library libpf_router.so.  A non-empty variant (or PF_LIB_VARIANT in the environment) builds
    libpf_router_<variant>.so in its own object directory with PF_EXTRA_NVCC_FLAGS added — side-by-side experiment builds
    that tools/ab_bench.sh compares on the GPU (the Python mirror loads $PF_ROUTER_LIB instead of the product library)."""
    variant = variant or os.environ.get("PF_LIB_VARIANT", "")
    lib_path = LIB if not variant else os.path.join(HERE, "libpf_router_%s.so" % variant)
    srcs = [os.path.join(CSRC, f) for f in ("pf_kernels.cu", "pf_router.cpp", "pf_sta.cpp", "pf_check.cpp", "pf_gen.cpp", "pf_file.c", "pf_text.c")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("pf_device.cuh", "pf_sta_device.cuh", "pf_gen_device.cuh", "pf_backend.h", "pf_layout.h", "pf_host.h")] + [
        os.path.join(ROOT, "include", f) for f in ("pf_router.h", "pf_types.h", "pf_file.h", "pf_gen.h", "pf_text.h")]
    if not force and not variant and _newer(lib_path, deps):
        return lib_path
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PF_EXTRA_NVCC_FLAGS", "").split()     # experiments, e.g. -DPF_DEUNROLL_COLD=1 (pf_device.cuh)
    bdir = os.path.join(HERE, "_build" + ("_" + variant if variant else ""))
    os.makedirs(bdir, exist_ok=True)
    objs = []
    for s in srcs:
        o = os.path.join(bdir, os.path.basename(s) + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-x", "cu"] if s.endswith(".cu") else []) + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % s)
        if s.endswith(".cu"):
            with open(os.path.join(bdir, "ptxas_pf_kernels.txt"), "w") as f:
                f.write(r.stderr)
        objs.append(o)
    # link with the host compiler: `nvcc -shared` would add a default-arch (sm_52) device-link stub
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
    cmd = ["g++", "-shared", "-o", lib_path] + objs + ["-L" + cuda_lib, "-Wl,-rpath," + cuda_lib, "-lcudart", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
