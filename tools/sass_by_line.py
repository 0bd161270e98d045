"""Static code-size profile of a kernel: SASS instructions per source line and per enclosing device function (no GPU).

    python tools/sass_by_line.py [kernel-substring]        # default: pf_route_kernelILi1E (the strict variant bench.py runs)

Uses cuobjdump -xelf + nvdisasm -g on the in-tree libpf_router.so (built with -lineinfo).  Instructions that nvdisasm
attributes to CUDA header intrinsics (shuffles, votes, atomics) are charged to the device function of the last repository
line seen before them, which is where they were inlined.  Why it matters: ncu shows 23 % of the route kernel's stall
samples as stall_no_inst (profiles/r01c_ncu_source_hotspots.txt) — instruction fetch — and the kernel is ~120 KB of SASS.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else "pf_route_kernelILi1E"
    lib = os.path.join(ROOT, "parallel_eda_b200", "libpf_router.so")
    src = open(os.path.join(ROOT, "parallel_eda_b200", "csrc", "pf_device.cuh")).read().split("\n")
    # enclosing function of every line of pf_device.cuh
    func_of, cur = {}, "?"
    for i, l in enumerate(src, 1):
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?PF_DEV\s+[\w\s\*&:<>]+?\b(pf_\w+)\s*\(", l)
        if m:
            cur = m.group(1)
        func_of[i] = cur
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True, check=True).stdout
    by_line, by_func = collections.Counter(), collections.Counter()
    fn, line, ctx = None, None, "?"
    for l in sass.split("\n"):
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            f, n = os.path.basename(m.group(1)), int(m.group(2))
            line = (f, n)
            if f == "pf_device.cuh" and n > 150:        # above: one-line wrappers of intrinsics, inlined everywhere
                ctx = func_of.get(n, "?")
            elif f == "pf_kernels.cu":
                ctx = "pf_kernels.cu (kernel prologue)"
            continue
        if fn and want in fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            by_line[line] += 1
            by_func[ctx] += 1
    # mnemonics: what the kernel does to memory (load widths, atomics, prefetches), warp collectives, fp64
    mnem = collections.Counter()
    fn = None
    for l in sass.split("\n"):
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            fn = m.group(1)
            continue
        if fn and want in fn:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", l)
            if m:
                mnem[m.group(1)] += 1
    tot = sum(by_line.values())
    print("%s: %d SASS instructions = %.0f KB" % (want, tot, tot * 16 / 1024))
    print("\nby enclosing device function (intrinsics charged to the function they were inlined into):")
    for f, c in by_func.most_common(25):
        print("  %5d %5.1f%%  %s" % (c, 100.0 * c / tot, f))
    keep = re.compile(r"^(LDG|STG|LDS|STS|ATOM|RED|REDUX|CCTL|SHFL|VOTE|MATCH|WARPSYNC|BAR|D[A-Z]+$|DSETP|F2F|MUFU|IMAD\.MOV|LDL|STL)")
    print("\nmemory / collective / fp64 mnemonics (count in the kernel's SASS):")
    for k, c in sorted(mnem.items(), key=lambda kv: -kv[1]):
        if keep.match(k):
            print("  %5d  %s" % (c, k))
    print("\nby source line:")
    for (f, n), c in by_line.most_common(40):
        text = src[n - 1].strip()[:110] if f == "pf_device.cuh" and n - 1 < len(src) else ""
        print("  %5d %5.1f%%  %s:%d  %s" % (c, 100.0 * c / tot, f, n, text))


if __name__ == "__main__":
    main()
