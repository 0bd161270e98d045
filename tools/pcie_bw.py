"""Host<->device copy bandwidth of this box (pinned memory), the ceiling of the e2e upload phase."""
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    print("%s pinned 1 GiB: %.1f GB/s" % (name, n / best / 1e9))
# 4 MiB pieces issued back to back, as the router's upload does
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        for o in range(0, n, 4 << 20): d[o:o + (4 << 20)].copy_(h[o:o + (4 << 20)], non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
print("h2d pinned 4 MiB pieces: %.1f GB/s" % (n / best / 1e9))

# the same from a write-combined pinned buffer (what the router stages its upload in)
import ctypes
rt = ctypes.CDLL("libcudart.so")
ptr = ctypes.c_void_p()
assert rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(n), ctypes.c_uint(4)) == 0
ctypes.memset(ptr, 1, n)
rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
best = 1e9
for _ in range(5):
    torch.cuda.synchronize(); t = time.perf_counter()
    assert rt.cudaMemcpy(ctypes.c_void_p(d.data_ptr()), ptr, n, 1) == 0
    torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
print("h2d write-combined 1 GiB: %.1f GB/s" % (n / best / 1e9))
