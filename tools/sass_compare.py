"""Which kernels changed between two builds of libpf_router.so: per-function hash of the SASS mnemonics (cuobjdump -sass; no GPU).
usage: python tools/sass_compare.py old/libpf_router.so new/libpf_router.so      prints every pf_route_kernel variant and every function that differs"""
import subprocess, sys, hashlib, re
def funcs(so):
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    cur = None; d = {}
    for l in out.split("\n"):
        m = re.search(r"Function : (\S+)", l)
        if m: cur = m.group(1); d[cur] = []; continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
            d[cur].append(re.sub(r"/\*[0-9a-fx]+\*/", "", l).strip())
    return d
a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
for k in sorted(set(a) | set(b)):
    ha = hashlib.md5("\n".join(a.get(k, [])).encode()).hexdigest()[:10] if k in a else "-"
    hb = hashlib.md5("\n".join(b.get(k, [])).encode()).hexdigest()[:10] if k in b else "-"
    if ha != hb or "route_kernel" in k:
        print("%-70s %6d %s | %6d %s %s" % (k[:70], len(a.get(k, [])), ha, len(b.get(k, [])), hb, "SAME" if ha == hb else "DIFF"))
