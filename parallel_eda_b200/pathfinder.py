"""The PathFinder outer loop on top of the C-ABI step functions, one process per GPU.

``route`` mirrors try_timing_driven_route (reference vpr/SRC/route/route_timing.c:85-343): route
every net, reserve locally used OPINs, test feasibility, raise pres_fac, update costs, run the host
STA.  With ``world_size > 1`` it is the data-parallel scheme of the reference's own MPI router
(parallel_route/mpi_route_load_balanced_nonblocking_send_recv_encoded.cxx): nets are sharded over
ranks, every rank keeps a full graph + congestion replica, and the occupancy changes are exchanged
across ranks — there a dense `MPI_Allreduce` (spatial.cxx:3371-3383), here an NCCL all-gather of each
rank's event log (4 bytes per changed rr node) over NVLink, twice per iteration (after the
stripe-interior nets and after the nets that cross a stripe cut), replayed with atomics on the node
records (pf_comm_events / pf_comm_apply_events).

``comm`` is anything with ``sync_occupancy(router)`` and ``all_reduce_*``; parallel_eda_b200.distributed wraps
torch.distributed (NCCL on GPUs; gloo on CPU tensors for the host-logic tests).
"""
from __future__ import annotations

import dataclasses
import time
from typing import Callable, List, Optional

import numpy as np

from . import pfio, router

HUGE_POSITIVE_FLOAT = 1.0e30
FIRST_ITER_WIRELENGTH_LIMIT = 0.85


@dataclasses.dataclass
class RouteReport:
    success: bool
    iterations: int
    nets_routed: int                 # sum over iterations and ranks of nets (re-)routed
    overused: List[int]
    per_iter_nets: List[int]
    heap_pops: int
    heap_pushes: int
    edge_visits: int
    wall_s: float


def run(r: router.Router, comm=None, sta: Optional[Callable] = None, dsta=None) -> RouteReport:
    """The product path: the whole loop inside the library (pf_route_run) — one host-device synchronisation per
    iteration, and with several ranks (``comm.connect(r)`` done) a device-side occupancy exchange over NVLink peer memory
    after every route part.  ``comm`` is only used to add up the ranks' counters afterwards."""
    t0 = time.perf_counter()
    ok, it, st = r.run(sta=sta, dsta=dsta)
    nets = [int(x) for x in st["nets_routed"]]
    rep = RouteReport(ok, it, sum(nets), [int(x) for x in st["overused_nodes"]], nets, int(st["heap_pops"].sum()),
                      int(st["heap_pushes"].sum()), int(st["edge_visits"].sum()), time.perf_counter() - t0)
    rep.crit_path_delay = [float(x) for x in st["crit_path_delay"]]
    return rep


def route(r: router.Router, comm=None, sta: Optional[Callable] = None, delay_buf=None, dsta=None,
          max_iters: Optional[int] = None, sync_rounds: int = 2) -> RouteReport:
    """Iterate until legal.  ``delay_buf``: tensor aliasing the router's device net_delay vector (needed when
    several ranks route timing-driven: the ranks' sink delays are summed into it before the analysis);
    ``sta``: host analysis callback; ``dsta``: a router.Sta — the analysis runs on the device in place."""
    o = r.problem.opts
    n_iter = int(max_iters or o["max_router_iterations"])
    pres_fac = float(o["first_iter_pres_fac"])
    crit = None
    overused, per_iter = [], []
    pops = pushes = visits = total_nets = 0
    t0 = time.perf_counter()
    success = False
    it = 0
    for it in range(1, n_iter + 1):
        if comm is None:
            st = r.route_iteration(pres_fac, crit)
            nets = st.nets_routed
            pops += st.heap_pops; pushes += st.heap_pushes; visits += st.edge_visits
        else:
            # two sub-rounds with an occupancy sync after each (pf_router_create): first every rank routes the nets
            # whose bounding boxes lie inside its own stripe of the grid — those cannot touch another stripe's
            # nets, so nobody works on a stale view — then the ranks that own a cut route the nets reaching across
            # it, with all interior routes visible.  (sync_rounds != 2: plain equal slices.)
            r.iteration_begin(crit)
            nets = 0
            for part in range(sync_rounds):
                st = r.iteration_route_part(pres_fac, part, sync_rounds)
                nets += st.nets_routed
                pops += st.heap_pops; pushes += st.heap_pushes; visits += st.edge_visits
                comm.sync_occupancy(r)                   # all-gather of the event logs over NVLink / NVSwitch
            nets = int(comm.all_reduce_scalar(nets))
        total_nets += nets
        per_iter.append(nets)
        breadth_first = int(o["router_algorithm"]) == 1       # try_breadth_first_route, route_breadth_first.c:23-91
        if it == 1 and not breadth_first:
            wl, avail = r.total_wirelength()
            if comm is not None:
                wl = int(comm.all_reduce_scalar(wl))
            if wl / max(avail, 1) > FIRST_ITER_WIRELENGTH_LIMIT:     # route_timing.c:189-225
                overused.append(-1)
                break
        if it == 1:
            new_pres, acc_fac = float(o["initial_pres_fac"]), (float(o["acc_fac"]) if breadth_first else 0.0)
        else:
            # single precision like the reference's `pres_fac *= router_opts.pres_fac_mult` (route_timing.c:288): a double
            # product rounded once differs from it by an ulp after a few iterations, and with it the routing
            new_pres = float(min(np.float32(pres_fac) * np.float32(o["pres_fac_mult"]), np.float32(HUGE_POSITIVE_FLOAT / 1e5)))
            acc_fac = float(o["acc_fac"])
        r.reserve_locally_used_opins(pres_fac, it != 1)   # every rank, on the same synced occupancy
        over = r.pathfinder_update_cost(acc_fac)
        pres_fac = new_pres
        overused.append(over)
        if over == 0:
            success = True
            break
        if dsta is not None and int(o["timing_analysis_enabled"]):
            # device STA (router.Sta): reads the router's delay vector, writes its criticality vector in place
            if comm is not None and delay_buf is not None:
                _assemble_delays(r, comm, delay_buf)
            dsta.analyze_device(r.comm_net_delay_ptr(), r.comm_crit_ptr())
            crit = None
        elif sta is not None and int(o["timing_analysis_enabled"]):
            if comm is not None and delay_buf is not None:
                _assemble_delays(r, comm, delay_buf)     # assemble every rank's sink delays
            crit, _cpd = sta(it, r.net_delay())
            crit = np.ascontiguousarray(crit, dtype=np.float32)
    if success and comm is not None and delay_buf is not None and int(o["timing_analysis_enabled"]):
        _assemble_delays(r, comm, delay_buf)             # the result carries every net's delays on every rank
    return RouteReport(success, it, total_nets, overused, per_iter, pops, pushes, visits, time.perf_counter() - t0)


def _assemble_delays(r, comm, delay_buf):
    """Step-API transport (torch collectives): every rank keeps the delays of the nets IT routes up to date — also of
    nets it did not re-route this iteration; the entries of other ranks' nets are whatever the last assembly left there.
    Sum only the owned entries."""
    import torch
    mask = getattr(r, "_own_mask", None)
    if mask is None:
        owner, _cut = r.comm_net_classes()
        p = r.problem
        per_term = np.repeat(owner, np.diff(p.net_ptr))
        mask = torch.from_numpy((per_term == r.config.rank).astype(np.float32)).to(delay_buf.device)
        r._own_mask = mask
    tmp = delay_buf * mask
    comm.all_reduce_sum_(tmp)
    delay_buf.copy_(tmp)
