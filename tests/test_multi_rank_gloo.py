"""N>1 host logic on CPU: two processes over gloo, each routing its shard of the nets with the emulated device
code, exchanging their occupancy event logs after each route part exactly as the NCCL path does on GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch
from parallel_eda_b200 import pfio, router, pathfinder, distributed, check_route
comm = distributed.init_from_env("gloo")
p = pfio.read_problem(os.path.join(%(root)r, "tests", "golden", "toy_w64.pfp.xz")); p.opts["timing_analysis_enabled"] = 0
lib = %(emu)r
cfg = router.default_config(router.load_library(lib), num_slots=2, big_slots=1, rank=comm.rank, nranks=comm.world)
R = comm.create_router(p, cfg, lib_path=lib)     # rank 0 packs the graph, rank 1 receives it by broadcast
rep = pathfinder.route(R, comm=comm)
res = R.result()
# every rank holds the full occupancy; traces only of its own nets
occ = torch.from_numpy(res.occ.astype(np.int64)); ref = occ.clone(); torch.distributed.broadcast(ref, 0)
own = [i for i in p.routed_nets() if res.trace_ptr[i + 1] > res.trace_ptr[i]]
parts = [None] * comm.world
torch.distributed.all_gather_object(parts, (own, [res.net_trace(int(i)) for i in own]))
if comm.rank == 0:
    seen = {}
    for o, tr in parts:
        for i, t in zip(o, tr): assert i not in seen; seen[int(i)] = t
    assert sorted(seen) == [int(i) for i in p.routed_nets()]
    tp = [0]; tn = []; ts = []
    for i in range(p.num_nets):
        if i in seen: tn.append(seen[i][0]); ts.append(seen[i][1])
        tp.append(tp[-1] + (len(seen[i][0]) if i in seen else 0))
    full = pfio.Result(int(rep.success), rep.iterations, 0, 0, np.array(tp, np.int32), np.concatenate(tn), np.concatenate(ts), res.net_delay, res.occ, res.iter_stats)
    m = check_route.check_route(p, full, check_delays=False)
    print(json.dumps({"success": rep.success, "iters": rep.iterations, "occ_equal": bool(torch.equal(occ, ref)), "overused": m["overused"], "nets": rep.nets_routed}))
else:
    assert torch.equal(occ, ref)
'''


def test_two_ranks_gloo(emu_lib, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "emu": emu_lib})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["success"] and out["occ_equal"] and out["overused"] == 0


WORKER_TD = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch
from parallel_eda_b200 import pfio, router, pathfinder, distributed
comm = distributed.init_from_env("gloo")
G = os.path.join(%(root)r, "tests", "golden")
p = pfio.read_problem(os.path.join(G, "duo_w80.pfp.xz")); p.opts["timing_analysis_enabled"] = 1; p.opts["max_router_iterations"] = 150
g = pfio.read_timing_graph(os.path.join(G, "duo_w80.pftg.xz"))
lib = %(emu)r
cfg = router.default_config(router.load_library(lib), num_slots=4, big_slots=1)
R = comm.create_router(p, cfg, lib_path=lib)
S = router.Sta(g, p, cfg, lib_path=lib)
delay = distributed.wrap_device_floats(R.comm_net_delay_ptr(), p.num_terminals, comm.device)
rep = pathfinder.route(R, comm=comm, dsta=S, delay_buf=delay)
res = R.result()
occ = torch.from_numpy(res.occ.astype(np.int64)); ref = occ.clone(); torch.distributed.broadcast(ref, 0)
crit = distributed.wrap_device_floats(R.comm_crit_ptr(), p.num_terminals, comm.device).clone(); cref = crit.clone(); torch.distributed.broadcast(cref, 0)
if comm.rank == 0:
    print(json.dumps({"success": bool(rep.success), "iters": rep.iterations, "occ_equal": bool(torch.equal(occ, ref)), "crit_equal": bool(torch.equal(crit, cref)),
                      "overused": int((res.occ > p.capacity).sum())}))
else:
    assert torch.equal(occ, ref) and torch.equal(crit, cref)
'''


def test_two_ranks_timing_driven_with_device_sta(emu_lib, tmp_path):
    """Timing-driven on two ranks: the ranks' sink delays are summed (all-reduce) into the router's delay vector,
    the device analysis then runs identically on every rank and writes the criticality vector in place."""
    script = tmp_path / "worker_td.py"
    script.write_text(WORKER_TD % {"root": ROOT, "emu": emu_lib})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    import json
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["success"] and out["occ_equal"] and out["crit_equal"] and out["overused"] == 0


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_stripe_sharding_properties(nranks, emu_lib):
    """The sharding every rank computes at create (no communication): all ranks agree, every routed net has one
    owner, and stripe-interior nets of different ranks are at least one maximum wire length apart in x — they can
    never touch the same rr node, which is what lets all ranks route them at once."""
    import numpy as np
    from parallel_eda_b200 import pfio, router
    lib = router.load_library(emu_lib)
    p = router.generate_grid_problem(lib_path=emu_lib, nx=48 if nranks < 8 else 112, ny=12, W=20, num_nets=600 if nranks < 8 else 1400, window=6, seed=5)
    owners = []
    for rank in range(nranks):
        R = router.Router(p, router.default_config(lib, num_slots=2, big_slots=1, rank=rank, nranks=nranks), lib_path=emu_lib)
        owners.append(R.comm_net_classes())
        R.close()
    owner, cut = owners[0]
    for o, c in owners[1:]:
        assert np.array_equal(o, owner) and np.array_equal(c, cut)
    routed = p.net_is_global == 0
    assert owner[routed].min() >= 0 and owner[routed].max() == nranks - 1 and (cut[routed] == 0).sum() > 0
    lmax = int(round(1.0 / float(p.indexed["inv_length"][4:].min())))
    bb = p.net_bb.reshape(-1, 4)
    interior = routed & (cut == 0)
    for a in range(nranks):
        for b in range(a + 1, nranks):
            xa_max = bb[interior & (owner == a), 1].max(initial=-10**9)
            xb_min = bb[interior & (owner == b), 0].min(initial=10**9)
            assert xb_min - xa_max >= lmax, (a, b, xa_max, xb_min, lmax)
    # cut nets belong to the rank that owns the cut they cross: never to rank 0
    assert not ((cut == 1) & routed & (owner == 0)).any()


WORKER_BF = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch
from parallel_eda_b200 import pfio, router, pathfinder, distributed
comm = distributed.init_from_env("gloo")
p = pfio.read_problem(os.path.join(%(root)r, "tests", "golden", "toy_w64_bf.pfp.xz"))
lib = %(emu)r
R = comm.create_router(p, router.default_config(router.load_library(lib), num_slots=4, big_slots=1), lib_path=lib)
rep = pathfinder.route(R, comm=comm)
res = R.result()
occ = torch.from_numpy(res.occ.astype(np.int64)); ref = occ.clone(); torch.distributed.broadcast(ref, 0)
if comm.rank == 0:
    print(json.dumps({"success": bool(rep.success), "occ_equal": bool(torch.equal(occ, ref)), "overused": int((res.occ > p.capacity).sum())}))
else:
    assert torch.equal(occ, ref)
'''


def test_two_ranks_breadth_first(emu_lib, tmp_path):
    """The breadth-first router mode through the same two-part iteration and event-log sync."""
    script = tmp_path / "worker_bf.py"
    script.write_text(WORKER_BF % {"root": ROOT, "emu": emu_lib})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29535", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    import json
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["success"] and out["occ_equal"] and out["overused"] == 0


WORKER_NATIVE = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch
from parallel_eda_b200 import pfio, router, pathfinder, distributed, check_route
comm = distributed.init_from_env("gloo")
G = os.path.join(%(root)r, "tests", "golden")
td = %(td)d
lib = %(emu)r
L = router.load_library(lib)
# three routers in a row on the same process group (timing-driven: two): the library keeps the exchange region and the peers'
# mappings between routers.  A smaller problem first, so that the second router finds the cached region too small and a new
# one is allocated while the old one is retired; the last router re-uses region and mappings; sequence numbers continue.
names = ["duo_w80"] * 2 if td else ["toy_w64", "hub_w90", "hub_w90"]
for k, name in enumerate(names):
    p = pfio.read_problem(os.path.join(G, name + ".pfp.xz")); p.opts["timing_analysis_enabled"] = td
    if td: p.opts["max_router_iterations"] = 100
    cfg = router.default_config(L, num_slots=4, big_slots=2, rank=comm.rank, nranks=comm.world)
    R = comm.create_router(p, cfg, lib_path=lib)      # includes comm.connect(R): pf_comm_export / all-gather / pf_comm_init
    S = router.Sta(pfio.read_timing_graph(os.path.join(G, name + ".pftg.xz")), p, cfg, lib_path=lib) if td else None
    rep = pathfinder.run(R, comm=comm, dsta=S)         # pf_route_run: exchange over (emulated) peer memory, no torch collective
    assert rep.success, (k, name)
    res = R.result()
    if k + 1 < len(names):
        if S is not None: S.close()
        R.close()
import ctypes
st = (ctypes.c_int64 * 5)(); L.pf_debug_comm_cache(st); st = list(st)      # bytes, regions allocated, re-used, peers mapped, mappings re-used
assert st[1] == (1 if td else 2) and st[2] == 1 and st[3] == st[1] and st[4] == 1, st
occ = torch.from_numpy(res.occ.astype(np.int64)); ref = occ.clone(); torch.distributed.broadcast(ref, 0)
nd = torch.from_numpy(res.net_delay.copy()); ndref = nd.clone(); torch.distributed.broadcast(ndref, 0)
own = [i for i in p.routed_nets() if res.trace_ptr[i + 1] > res.trace_ptr[i]]
parts = [None] * comm.world
torch.distributed.all_gather_object(parts, (own, [res.net_trace(int(i)) for i in own]))
if comm.rank == 0:
    seen = {}
    for o, tr in parts:
        for i, t in zip(o, tr): assert i not in seen; seen[int(i)] = t
    assert sorted(seen) == [int(i) for i in p.routed_nets()]
    tp = [0]; tn = []; ts = []
    for i in range(p.num_nets):
        if i in seen: tn.append(seen[i][0]); ts.append(seen[i][1])
        tp.append(tp[-1] + (len(seen[i][0]) if i in seen else 0))
    full = pfio.Result(int(rep.success), rep.iterations, 0, 0, np.array(tp, np.int32), np.concatenate(tn), np.concatenate(ts), res.net_delay, res.occ, res.iter_stats)
    # legality of the union of the ranks' traces AND every sink delay (also of nets the OTHER rank routed, and of nets that were
    # not re-routed in the last iterations) against a from-scratch Elmore recomputation
    m = check_route.check_route(p, full, check_delays=bool(td))
    print(json.dumps({"success": bool(rep.success), "iters": rep.iterations, "occ_equal": bool(torch.equal(occ, ref)), "delay_equal": bool(torch.equal(nd, ndref)),
                      "overused": m["overused"], "routed_by": [len(o) for o, _ in parts]}))
else:
    assert torch.equal(occ, ref) and torch.equal(nd, ndref)
'''


@pytest.mark.parametrize("td", [0, 1])
def test_two_ranks_native_transport(td, emu_lib, tmp_path):
    """pf_comm_export / pf_comm_init / pf_route_run: the transport inside the library.  On the emulator the ranks'
    exchange regions are POSIX shared memory and the device-side protocol (payload, release of the sequence number,
    acquire-polling consumers, double buffering by parity) is the one the CUDA kernels run over NVLink.  Timing-driven:
    the ranks publish the sink delays of their nets and gather the others' before the device analysis."""
    script = tmp_path / "worker_native.py"
    script.write_text(WORKER_NATIVE % {"root": ROOT, "emu": emu_lib, "td": td})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29537 + td), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    import json
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["success"] and out["occ_equal"] and out["delay_equal"] and out["overused"] == 0
    assert min(out["routed_by"]) > 0, out          # both ranks actually routed nets


WORKER_GRID8 = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch
from parallel_eda_b200 import router, pathfinder, distributed
comm = distributed.init_from_env("gloo")
lib = %(emu)r
L = router.load_library(lib)
p, gen = router.generate_grid_nets(nx=64, ny=64, W=100, num_nets=4000)       # the rr graph is generated by every rank (pf_router_create_generated)
cfg = router.default_config(L, num_slots=8, big_slots=2, rank=comm.rank, nranks=comm.world)
R = comm.create_router(p, cfg, generated=gen, lib_path=lib)
rep = pathfinder.run(R, comm=comm)
res = R.result()
occ = torch.from_numpy(res.occ.astype(np.int64)); ref = occ.clone(); torch.distributed.broadcast(ref, 0)
mine = torch.tensor([int(rep.nets_routed)], dtype=torch.int64); every = [torch.zeros(1, dtype=torch.int64) for _ in range(comm.world)]
torch.distributed.all_gather(every, mine)
assert torch.equal(occ, ref)
if comm.rank == 0:
    full = router.generate_grid_problem(lib_path=lib, nx=64, ny=64, W=100, num_nets=4000)
    print(json.dumps({"success": bool(rep.success), "iters": int(rep.iterations), "overused": int((res.occ > full.capacity).sum()),
                      "routed_by": [int(x) for x in every]}))
'''


def test_eight_ranks_native_transport_on_a_generated_grid(emu_lib, tmp_path):
    """The shape of the driver's 8-GPU run on the emulator: eight ranks, the graph generated on every rank, stripe sharding with
    the partition computed on a helper thread, the device-side exchange protocol between eight exchange regions (each rank
    polls seven peers), one host read per iteration.  Every rank ends with the same legal occupancy."""
    script = tmp_path / "worker_grid8.py"
    script.write_text(WORKER_GRID8 % {"root": ROOT, "emu": emu_lib})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=8", "--master-addr", "127.0.0.1",
           "--master-port", "29547", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=dict(os.environ, PF_ALLOW_EMULATOR="1"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    import json
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["success"] and out["overused"] == 0 and out["iters"] <= 50
    assert len(out["routed_by"]) == 8 and sum(1 for x in out["routed_by"] if x > 0) >= 7, out
