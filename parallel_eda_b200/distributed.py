"""torch.distributed plumbing for the multi-GPU router: one process per GPU, NCCL over NVLink.

Only the collectives live here; the routing itself is in libpf_router.so.  On a CPU box the same
code runs over gloo with host tensors (tests/test_multi_rank_gloo.py), which is how the N>1 host
logic is covered without GPUs.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class Comm:
    def __init__(self, device: torch.device):
        self.device = device
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self._counts = None
        self._events = None

    def all_reduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()
        return t

    def create_router(self, problem, config=None, lib_path=None, generated=None):
        """A Router on every rank with ONE upload over PCIe: rank 0 packs the graph and uploads it, the other
        ranks are created with defer_graph and receive the packed node records / edge words / ptc numbers by an
        NCCL broadcast over NVLink (with N processes packing and uploading at once the host memory system is the
        bottleneck: 107 ms per rank at N = 4 against 45 ms alone)."""
        from . import router as _router
        lib = _router.load_library(lib_path)
        cfg = _router.Config.from_buffer_copy(config if config is not None else _router.default_config(lib))
        cfg.rank, cfg.nranks = self.rank, self.world
        # a generated fabric is built on every rank's own device (pf_router_create_generated): nothing to broadcast
        cfg.defer_graph = 0 if (self.rank == 0 or generated is not None) else 1
        if os.environ.get("PF_COMM_DEBUG"):
            cfg.verbose = 1
        import time
        R, err = None, None
        t0 = time.perf_counter()
        try:
            R = _router.Router(problem, cfg, lib_path=lib_path, generated=generated)
        except _router.RouterError as e:
            err = e
        t1 = time.perf_counter()
        if self.all_reduce_scalar(1 if err else 0) > 0:          # nobody waits in a broadcast for a rank that failed
            if R is not None:
                R.close()
            raise err or _router.RouterError(-5, "another rank could not create its router")
        for ptr, nbytes in (R.comm_graph_buffers() if generated is None else []):
            t = wrap_device_bytes(ptr, nbytes, self.device)
            dist.broadcast(t, src=0)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        if self.rank != 0 and generated is None:
            R.comm_graph_ready()
        self.connect(R)
        if os.environ.get("PF_COMM_DEBUG"):
            print("rank %d create_router: create %.1f ms, status + broadcast %.1f ms" % (self.rank, (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3), flush=True)
        return R

    def connect(self, r) -> None:
        """Bootstrap of the transport inside the library (pf_comm_export / pf_comm_init): all-gather the ranks' 128-byte
        handles once; afterwards the occupancy exchange is device-side (peer memory over NVLink) and this class is no
        longer on the path."""
        import time
        t0 = time.perf_counter()
        mine = torch.frombuffer(bytearray(r.comm_export()), dtype=torch.uint8).to(self.device)
        every = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(every, mine)
        blob = bytes(every.cpu().numpy().tobytes())
        t1 = time.perf_counter()
        r.comm_init(blob)
        t2 = time.perf_counter()
        dist.barrier()                                    # every rank has mapped every region before anybody publishes
        if os.environ.get("PF_COMM_DEBUG"):
            print("rank %d connect: export + all-gather %.2f ms, pf_comm_init %.2f ms, barrier %.2f ms" % (
                self.rank, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t2) * 1e3), flush=True)

    def sync_occupancy(self, r) -> int:
        """All ranks end up with the same rr-node occupancy: all-gather every rank's event log of the last
        route part (router.comm_events: 4 bytes per changed rr node) and replay the others' on this rank.
        Returns the number of events received."""
        ptr, n = r.comm_events()
        mine = torch.tensor([n], dtype=torch.int64, device=self.device)
        if self._counts is None:
            self._counts = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(self._counts, mine)
        counts = self._counts.tolist()
        m = max(counts)
        if m == 0:
            return 0
        send = wrap_device_ints(ptr, m, self.device)          # entries past n are padding, never replayed
        if self._events is None or self._events.numel() < self.world * m:
            self._events = torch.empty(self.world * (m + (m >> 2) + 1024), dtype=torch.int32, device=self.device)
        recv = self._events[: self.world * m]
        dist.all_gather_into_tensor(recv, send)
        if recv.is_cuda:
            torch.cuda.current_stream(recv.device).synchronize()
        got = 0
        for k in range(self.world):
            if k != self.rank and counts[k] > 0:
                r.comm_apply_events(recv.data_ptr() + 4 * k * m, counts[k])
                got += counts[k]
        return got

    def all_reduce_scalar(self, v, op=dist.ReduceOp.SUM):
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=op)
        return t.item()

    def all_reduce_max(self, v: float) -> float:
        return self.all_reduce_scalar(v, dist.ReduceOp.MAX)

    def barrier(self):
        dist.barrier()

    def release(self, lib_path=None) -> None:
        """After the last router of the job: free the exchange region and the peer mappings the library keeps between
        routers (pf_comm_release_cache).  Collective — no rank may still be inside a routing."""
        from . import router as _router
        dist.barrier()
        _router.load_library(lib_path).pf_comm_release_cache()
        dist.barrier()


def init_from_env(backend: str | None = None) -> Comm | None:
    """Join the process group torchrun described in RANK / WORLD_SIZE / MASTER_*; None if single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    if backend is None:
        backend = "nccl" if use_cuda else "gloo"
    if use_cuda:
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend)
    return Comm(torch.device("cuda", local) if use_cuda else torch.device("cpu"))


def wrap_device_bytes(ptr: int, n: int, device: torch.device) -> torch.Tensor:
    """A uint8 tensor view of n bytes at a raw device pointer owned by the router (no copy)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    if device.type == "cuda":
        return torch.as_tensor(h, device=device)
    import ctypes
    import numpy as np
    return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(ptr)))


def wrap_device_ints(ptr: int, n: int, device: torch.device) -> torch.Tensor:
    """An int32 tensor view of n words at a raw device pointer owned by the router (no copy)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}
    if device.type == "cuda":
        return torch.as_tensor(h, device=device)
    import ctypes
    import numpy as np
    return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_int32 * n).from_address(ptr)))


def wrap_device_floats(ptr: int, n: int, device: torch.device) -> torch.Tensor:
    """A float32 tensor view of n floats at a raw device pointer owned by the router (no copy)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}
    if device.type == "cuda":
        return torch.as_tensor(h, device=device)
    import ctypes
    import numpy as np
    arr = np.ctypeslib.as_array((ctypes.c_float * n).from_address(ptr))
    return torch.from_numpy(arr)
