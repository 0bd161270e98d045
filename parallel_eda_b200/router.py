"""Host-side Python mirror of the router boundary: ctypes binding of include/pf_router.h.

The names follow the reference's own interface for this path — ``try_timing_driven_route``
(reference vpr/SRC/route/route_timing.c:85), ``pathfinder_update_cost`` / ``feasible_routing``
(route_common.c:581,509), ``reserve_locally_used_opins`` (route_common.c:1435) — so the parity
tests read like the reference's call sequence.  All compute happens in libpf_router.so (CUDA,
sm_100a); this module only marshals numpy arrays across the C-ABI.  There is no CPU fallback: if
the library is missing or no B200 is visible the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional

import numpy as np

from . import pfio

_HERE = os.path.dirname(os.path.abspath(__file__))
# PF_ROUTER_LIB: an experiment build (parallel_eda_b200/build.py variants, tools/ab_bench.sh) instead of the product library
DEFAULT_LIB = os.environ.get("PF_ROUTER_LIB") or os.path.join(_HERE, "libpf_router.so")

PF_OK = 0
ERRORS = {-1: "PF_EIO", -2: "PF_EFORMAT", -3: "PF_ENOMEM", -4: "PF_EINVAL", -5: "PF_ECUDA", -6: "PF_EOVERFLOW",
          -7: "PF_EUNROUTABLE"}


class RouterError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "PF_E?"), code, msg))
        self.code = code


class _Opts(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("first_iter_pres_fac", "initial_pres_fac", "pres_fac_mult", "acc_fac",
                                         "bend_cost", "astar_fac", "max_criticality", "criticality_exp")] + \
               [(n, C.c_int32) for n in ("max_router_iterations", "timing_analysis_enabled", "bb_factor", "router_algorithm")]


class _Problem(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("num_nodes", C.c_int32), ("num_edges", C.c_int32),
                ("xlow", C.c_void_p), ("ylow", C.c_void_p), ("xhigh", C.c_void_p), ("yhigh", C.c_void_p),
                ("ptc_num", C.c_void_p), ("cost_index", C.c_void_p), ("capacity", C.c_void_p),
                ("type", C.c_void_p), ("direction", C.c_void_p), ("R", C.c_void_p), ("C", C.c_void_p),
                ("row_ptr", C.c_void_p), ("edge_to", C.c_void_p), ("edge_sw", C.c_void_p),
                ("num_switches", C.c_int32), ("switches", C.c_void_p),
                ("num_indexed", C.c_int32), ("indexed", C.c_void_p),
                ("num_nets", C.c_int32), ("num_terminals", C.c_int32),
                ("net_ptr", C.c_void_p), ("net_terminals", C.c_void_p), ("net_is_global", C.c_void_p),
                ("net_bb", C.c_void_p),
                ("num_opin_groups", C.c_int32), ("opin_group_source", C.c_void_p), ("opin_group_count", C.c_void_p),
                ("opts", _Opts)]


class IterStats(C.Structure):
    _fields_ = [("overused_nodes", C.c_int32), ("nets_routed", C.c_int32), ("heap_pushes", C.c_int64),
                ("heap_pops", C.c_int64), ("edge_visits", C.c_int64), ("pres_fac", C.c_float),
                ("crit_path_delay", C.c_float)]


class _Result(C.Structure):
    _fields_ = [("success", C.c_int32), ("iterations", C.c_int32), ("serial_num", C.c_int32),
                ("total_wirelength", C.c_int32), ("num_nets", C.c_int32),
                ("trace_ptr", C.POINTER(C.c_int32)), ("trace_node", C.POINTER(C.c_int32)),
                ("trace_switch", C.POINTER(C.c_int16)), ("num_terminals", C.c_int32),
                ("net_delay", C.POINTER(C.c_float)), ("num_nodes", C.c_int32), ("occ", C.POINTER(C.c_int32)),
                ("num_iter_stats", C.c_int32), ("iter_stats", C.POINTER(IterStats)),
                ("num_crit_iters", C.c_int32), ("iter_crit", C.POINTER(C.c_float))]


class Config(C.Structure):
    """pf_config (include/pf_router.h).  Zero fields mean "auto"."""
    _fields_ = [(n, C.c_int32) for n in ("device", "rank", "nranks", "num_slots", "warps_per_block", "label_log2", "label2_log2",
                                         "tree_cap", "far_cap", "sink_cap", "big_slots", "big_label_log2",
                                         "big_tree_cap", "big_far_cap", "max_batch")] + \
               [("pop_slack", C.c_float), ("win_rel", C.c_float), ("win_abs", C.c_float), ("verbose", C.c_int32),
                ("reroute_all_iters", C.c_int32), ("inflight_div", C.c_int32), ("min_slots", C.c_int32),
                ("stall_iters", C.c_int32), ("history_window", C.c_int32), ("keep_newcomer", C.c_int32),
                ("defer_graph", C.c_int32), ("validate_commits", C.c_int32), ("ripple", C.c_int32), ("ripple_max_nets", C.c_int32), ("polish", C.c_int32), ("lazy_seed_min", C.c_int32)]


class _TimingGraph(C.Structure):
    """pf_timing_graph (include/pf_types.h)."""
    _fields_ = [("num_tnodes", C.c_int32), ("num_tedges", C.c_int32), ("edge_ptr", C.c_void_p), ("edge_to", C.c_void_p),
                ("edge_Tdel", C.c_void_p), ("type", C.c_void_p), ("clock_domain", C.c_void_p), ("clock_delay", C.c_void_p),
                ("num_levels", C.c_int32), ("level_ptr", C.c_void_p), ("level_nodes", C.c_void_p), ("num_domains", C.c_int32),
                ("constraint", C.c_void_p), ("num_nets", C.c_int32), ("net_driver", C.c_void_p),
                ("num_overrides", C.c_int32), ("override_domain", C.c_void_p), ("override_tnode", C.c_void_p), ("override_constraint", C.c_void_p)]


class _TimingGraphHolder:
    def __init__(self, g: "pfio.TimingGraph"):
        self.keep = [np.ascontiguousarray(a, dtype=dt) for a, dt in (
            (g.edge_ptr, np.int32), (g.edge_to, np.int32), (g.edge_Tdel, np.float32), (g.type, np.uint8), (g.clock_domain, np.int32),
            (g.clock_delay, np.float32), (g.level_ptr, np.int32), (g.level_nodes, np.int32), (g.constraint, np.float32),
            (g.net_driver, np.int32), (g.override_domain, np.int32), (g.override_tnode, np.int32), (g.override_constraint, np.float32))]
        q = [a.ctypes.data if a.size else None for a in self.keep]
        self.c = _TimingGraph(g.num_tnodes, len(g.edge_to), q[0], q[1], q[2], q[3], q[4], q[5], g.num_levels, q[6], q[7],
                              int(g.constraint.shape[0]), q[8], len(g.net_driver), q[9], len(g.override_tnode), q[10], q[11], q[12])


class CheckReport(C.Structure):
    """pf_check_report (include/pf_router.h)."""
    _fields_ = [("ok", C.c_int32), ("bad_nets", C.c_int32), ("first_bad_net", C.c_int32), ("first_bad_code", C.c_int32),
                ("occupancy_mismatch", C.c_int32), ("overused_nodes", C.c_int32), ("wirelength", C.c_int64), ("reserved_opins", C.c_int64)]


class Timing(C.Structure):
    _fields_ = [("route_kernel_ms", C.c_double), ("update_kernel_ms", C.c_double), ("aux_kernel_ms", C.c_double),
                ("route_launches", C.c_int64), ("update_launches", C.c_int64), ("aux_launches", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


class GenParams(C.Structure):
    """pf_gen_params (include/pf_gen.h)."""
    _fields_ = [(n, C.c_int32) for n in ("nx", "ny", "W", "L", "num_nets", "sinks_per_net", "window")] + \
               [("seed", C.c_uint32), ("fc_in", C.c_float), ("fc_out", C.c_float), ("io_capacity", C.c_int32),
                ("bb_factor", C.c_int32), ("reserved", C.c_int32 * 4)]


STA_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float))

_libs = {}


def load_library(path: Optional[str] = None) -> C.CDLL:
    from_env = path is None and bool(os.environ.get("PF_ROUTER_LIB"))
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError("%s not found: build it with `python -m parallel_eda_b200.build` "
                           "(the router has no CPU fallback)" % path)
    lib = C.CDLL(path)
    lib.pf_last_error.restype = C.c_char_p
    lib.pf_backend_name.restype = C.c_char_p
    if from_env and not lib.pf_backend_name().startswith(b"cuda:") and os.environ.get("PF_ALLOW_EMULATOR") != "1":
        # PF_ROUTER_LIB is for experiment builds of the CUDA library; the warp emulator (tests/emu) is test infrastructure and
        # must never become a silent CPU path of the product
        raise RuntimeError("PF_ROUTER_LIB=%s is not a CUDA build (%s); set PF_ALLOW_EMULATOR=1 to load test infrastructure on purpose"
                           % (path, lib.pf_backend_name().decode()))
    lib.pf_config_default.argtypes = [C.POINTER(Config)]
    lib.pf_router_create.argtypes = [C.POINTER(_Problem), C.POINTER(Config), C.POINTER(C.c_void_p)]
    lib.pf_router_create_generated.argtypes = [C.POINTER(GenParams), C.POINTER(_Problem), C.POINTER(Config), C.POINTER(C.c_void_p)]
    lib.pf_gen_grid_nets.argtypes = [C.POINTER(GenParams), C.POINTER(_Problem)]
    lib.pf_debug_graph_hash.argtypes = [C.c_void_p, C.POINTER(C.c_uint64 * 3), C.POINTER(C.c_int64)]
    lib.pf_router_destroy.argtypes = [C.c_void_p]
    lib.pf_router_destroy.restype = None
    lib.pf_router_reset.argtypes = [C.c_void_p]
    lib.pf_route_iteration.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.POINTER(IterStats)]
    lib.pf_iteration_begin.argtypes = [C.c_void_p, C.c_void_p]
    lib.pf_iteration_route_part.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.POINTER(IterStats)]
    lib.pf_reserve_opins.argtypes = [C.c_void_p, C.c_float, C.c_int]
    lib.pf_update_costs.argtypes = [C.c_void_p, C.c_float, C.POINTER(C.c_int)]
    lib.pf_total_wirelength.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.pf_get_net_delay.argtypes = [C.c_void_p, C.c_void_p]
    lib.pf_get_result.argtypes = [C.c_void_p, C.POINTER(_Result)]
    lib.pf_get_timing.argtypes = [C.c_void_p, C.POINTER(Timing), C.c_int]
    lib.pf_timer_start.argtypes = [C.c_void_p]
    lib.pf_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.pf_check_route.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CheckReport)]
    lib.pf_sta_create.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Config), C.POINTER(C.c_void_p)]
    lib.pf_sta_destroy.argtypes = [C.c_void_p]
    lib.pf_sta_destroy.restype = None
    lib.pf_sta_analyze.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    lib.pf_sta_analyze_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    lib.pf_try_timing_driven_route_sta.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Config), C.c_void_p]
    lib.pf_comm_net_classes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pf_comm_graph_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p * 3), C.POINTER(C.c_int64 * 3)]
    lib.pf_comm_graph_ready.argtypes = [C.c_void_p]
    lib.pf_comm_events.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    lib.pf_comm_apply_events.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.pf_comm_net_delay_ptr.argtypes = [C.c_void_p]
    lib.pf_comm_net_delay_ptr.restype = C.c_void_p
    lib.pf_comm_crit_ptr.argtypes = [C.c_void_p]
    lib.pf_comm_crit_ptr.restype = C.c_void_p
    lib.pf_comm_export.argtypes = [C.c_void_p, C.c_void_p]
    lib.pf_comm_init.argtypes = [C.c_void_p, C.c_void_p]
    lib.pf_comm_exchange.argtypes = [C.c_void_p]
    lib.pf_comm_gather_delays.argtypes = [C.c_void_p]
    lib.pf_comm_abort.argtypes = [C.c_void_p]
    lib.pf_comm_release_cache.argtypes = []
    lib.pf_route_run.argtypes = [C.c_void_p, C.c_void_p, STA_FN, C.c_void_p, C.POINTER(IterStats), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.pf_try_timing_driven_route.argtypes = [C.POINTER(_Problem), C.POINTER(Config), STA_FN, C.c_void_p,
                                               C.POINTER(_Result)]
    lib.pf_result_free.argtypes = [C.POINTER(_Result)]
    lib.pf_result_free.restype = None
    lib.pf_gen_params_default.argtypes = [C.POINTER(GenParams)]
    lib.pf_gen_params_default.restype = None
    lib.pf_gen_grid_problem.argtypes = [C.POINTER(GenParams), C.POINTER(_Problem)]
    lib.pf_problem_free.argtypes = [C.POINTER(_Problem)]
    lib.pf_problem_free.restype = None
    lib.pf_problem_write.argtypes = [C.c_char_p, C.POINTER(_Problem)]
    _libs[path] = lib
    return lib


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


class _ProblemHolder:
    """Keeps contiguous numpy arrays alive while a C pf_problem points into them."""

    def __init__(self, p: pfio.Problem):
        self.arrays = {}
        cp = _Problem()
        # a nets-only problem (generate_grid_nets) has no node arrays: its node count is the generator's closed form
        cp.nx, cp.ny, cp.num_nodes, cp.num_edges = p.nx, p.ny, getattr(p, "gen_num_nodes", None) or p.num_nodes, p.num_edges
        for name, dt, _ in pfio._PROB_FIELDS:
            a = np.ascontiguousarray(getattr(p, name), dtype=np.dtype(dt))
            self.arrays[name] = a
            setattr(cp, name, _ptr(a))
        cp.num_switches, cp.num_indexed = len(p.switches), len(p.indexed)
        cp.num_nets, cp.num_terminals = p.num_nets, p.num_terminals
        cp.num_opin_groups = len(p.opin_group_source)
        o = np.asarray(p.opts, dtype=pfio.OPTS_DT)
        for f in _Opts._fields_:
            setattr(cp.opts, f[0], o[f[0]].item())
        self.c = cp


class _ResultOwner:
    """Keeps the C pf_result alive while numpy arrays view its buffers; frees it with the last view."""

    def __init__(self, lib, cr):
        self.lib, self.cr = lib, cr

    def __del__(self):
        try:
            self.lib.pf_result_free(C.byref(self.cr))
        except Exception:
            pass


def _view(owner, ptr, n, dtype):
    """numpy view of n elements at a ctypes pointer (no copy); the array keeps `owner` alive."""
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    ct = {np.dtype(np.int32): C.c_int32, np.dtype(np.int16): C.c_int16, np.dtype(np.float32): C.c_float}[np.dtype(dtype)]
    buf = (ct * n).from_address(C.addressof(ptr.contents))
    buf._owner = owner
    return np.frombuffer(buf, dtype=dtype)


def _take_result(lib, cr: _Result, T: int, with_stats: bool = True) -> pfio.Result:
    owner = _ResultOwner(lib, cr)
    n = cr.num_nets
    tp = _view(owner, cr.trace_ptr, n + 1, np.int32)
    nt = int(tp[n])
    tn = _view(owner, cr.trace_node, nt, np.int32)
    ts = _view(owner, cr.trace_switch, nt, np.int16)
    nd = _view(owner, cr.net_delay, cr.num_terminals, np.float32)
    occ = _view(owner, cr.occ, cr.num_nodes, np.int32)
    stats = np.zeros(cr.num_iter_stats if with_stats else 0, dtype=pfio.ITER_STATS_DT)
    for i in range(len(stats)):
        s = cr.iter_stats[i]
        stats[i] = (s.overused_nodes, s.nets_routed, s.heap_pushes, s.heap_pops, s.edge_visits, s.pres_fac,
                    s.crit_path_delay)
    crit = None
    if cr.num_crit_iters and cr.iter_crit:
        crit = _view(owner, cr.iter_crit, cr.num_crit_iters * T, np.float32).reshape(cr.num_crit_iters, T)
    return pfio.Result(cr.success, cr.iterations, cr.serial_num, cr.total_wirelength, tp, tn, ts, nd, occ, stats, crit)


class Router:
    """A device-resident routing problem (pf_router handle)."""

    def __init__(self, problem: pfio.Problem, config: Optional[Config] = None, lib_path: Optional[str] = None,
                 generated: "Optional[GenParams]" = None):
        """``generated``: the rr graph is built on the device from these generator parameters (``problem`` then comes from
        generate_grid_nets and carries no node / edge arrays) — pf_router_create_generated."""
        self.lib = load_library(lib_path)
        self.problem = problem
        self._holder = _ProblemHolder(problem)
        if config is None:
            config = default_config(self.lib)
        self.config = config
        h = C.c_void_p()
        if generated is not None:
            self._gen = generated
            rc = self.lib.pf_router_create_generated(C.byref(generated), C.byref(self._holder.c), C.byref(config), C.byref(h))
        else:
            rc = self.lib.pf_router_create(C.byref(self._holder.c), C.byref(config), C.byref(h))
        if rc != PF_OK:
            raise RouterError(rc, self.lib.pf_last_error().decode())
        self._h = h

    def _ck(self, rc: int):
        if rc != PF_OK:
            raise RouterError(rc, self.lib.pf_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.pf_router_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._ck(self.lib.pf_router_reset(self._h))

    def graph_hash(self):
        """(hash of node records, hash of edge words, hash of ptc numbers, number of edges) of the device graph."""
        h = (C.c_uint64 * 3)(); ne = C.c_int64(0)
        self._ck(self.lib.pf_debug_graph_hash(self._h, C.byref(h), C.byref(ne)))
        return int(h[0]), int(h[1]), int(h[2]), int(ne.value)

    def route_iteration(self, pres_fac: float, crit: Optional[np.ndarray] = None) -> IterStats:
        st = IterStats()
        cp = None
        if crit is not None:
            crit = np.ascontiguousarray(crit, dtype=np.float32)
            assert crit.shape[0] == self.problem.num_terminals
            cp = _ptr(crit)
        self._ck(self.lib.pf_route_iteration(self._h, pres_fac, cp, C.byref(st)))
        return st

    def iteration_begin(self, crit: Optional[np.ndarray] = None):
        cp = None
        if crit is not None:
            crit = np.ascontiguousarray(crit, dtype=np.float32)
            cp = _ptr(crit)
        self._ck(self.lib.pf_iteration_begin(self._h, cp))

    def iteration_route_part(self, pres_fac: float, part: int, nparts: int) -> IterStats:
        st = IterStats()
        self._ck(self.lib.pf_iteration_route_part(self._h, pres_fac, part, nparts, C.byref(st)))
        return st

    def reserve_locally_used_opins(self, pres_fac: float, rip_up_local_opins: bool):
        self._ck(self.lib.pf_reserve_opins(self._h, pres_fac, 1 if rip_up_local_opins else 0))

    def pathfinder_update_cost(self, acc_fac: float) -> int:
        """pathfinder_update_cost + feasible_routing in one pass; returns the overused-node count."""
        over = C.c_int(0)
        self._ck(self.lib.pf_update_costs(self._h, acc_fac, C.byref(over)))
        return over.value

    def total_wirelength(self):
        wl, av = C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.pf_total_wirelength(self._h, C.byref(wl), C.byref(av)))
        return wl.value, av.value

    def net_delay(self) -> np.ndarray:
        out = np.zeros(max(self.problem.num_terminals, 1), dtype=np.float32)
        self._ck(self.lib.pf_get_net_delay(self._h, _ptr(out)))
        return out[:self.problem.num_terminals]

    def result(self) -> pfio.Result:
        cr = _Result()
        self._ck(self.lib.pf_get_result(self._h, C.byref(cr)))
        return _take_result(self.lib, cr, self.problem.num_terminals, with_stats=False)

    def timing(self, reset: bool = False) -> Timing:
        t = Timing()
        self._ck(self.lib.pf_get_timing(self._h, C.byref(t), 1 if reset else 0))
        return t

    def timer_start(self):
        self._ck(self.lib.pf_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double(0.0)
        self._ck(self.lib.pf_timer_stop(self._h, C.byref(ms)))
        return ms.value

    # multi-GPU iteration boundary (device pointers, e.g. torch tensors' data_ptr())
    def check_route(self, result: pfio.Result) -> dict:
        """check_route (route/check_route.c:27) on the device for any routing of this problem: the router's own, the
        oracle's or a golden of the reference.  Returns the pf_check_report as a dict."""
        keep = [np.ascontiguousarray(result.trace_ptr, dtype=np.int32), np.ascontiguousarray(result.trace_node, dtype=np.int32),
                np.ascontiguousarray(result.trace_switch, dtype=np.int16), np.ascontiguousarray(result.occ, dtype=np.int32)]
        cr = _Result()
        cr.num_nets = len(keep[0]) - 1
        cr.trace_ptr = keep[0].ctypes.data_as(C.POINTER(C.c_int32)); cr.trace_node = keep[1].ctypes.data_as(C.POINTER(C.c_int32))
        cr.trace_switch = keep[2].ctypes.data_as(C.POINTER(C.c_int16))
        cr.num_nodes = len(keep[3]); cr.occ = keep[3].ctypes.data_as(C.POINTER(C.c_int32))
        rep = CheckReport()
        self._ck(self.lib.pf_check_route(self._h, C.byref(cr), C.byref(rep)))
        return {f: int(getattr(rep, f)) for f, _ in CheckReport._fields_}

    def comm_net_classes(self):
        """(owner[num_nets] int32, is_cut[num_nets] uint8): the stripe sharding decided at create."""
        n = self.problem.num_nets
        owner = np.zeros(n, np.int32); cut = np.zeros(n, np.uint8)
        self._ck(self.lib.pf_comm_net_classes(self._h, owner.ctypes.data, cut.ctypes.data))
        return owner, cut

    def comm_graph_buffers(self):
        """[(device pointer, bytes)] of the packed graph: node records, edge words, ptc numbers."""
        ptrs = (C.c_void_p * 3)(); nb = (C.c_int64 * 3)()
        self._ck(self.lib.pf_comm_graph_buffers(self._h, C.byref(ptrs), C.byref(nb)))
        return [(int(ptrs[k] or 0), int(nb[k])) for k in range(3)]

    def comm_graph_ready(self):
        self._ck(self.lib.pf_comm_graph_ready(self._h))

    def comm_events(self):
        """(device pointer, count) of this rank's occupancy event log of the last route part."""
        ptr = C.c_void_p(); n = C.c_int64(0)
        self._ck(self.lib.pf_comm_events(self._h, C.byref(ptr), C.byref(n)))
        return int(ptr.value or 0), int(n.value)

    def comm_apply_events(self, dev_ptr: int, count: int):
        self._ck(self.lib.pf_comm_apply_events(self._h, C.c_void_p(dev_ptr), count))

    def comm_net_delay_ptr(self) -> int:
        return int(self.lib.pf_comm_net_delay_ptr(self._h))

    # the transport inside the library (peer memory over NVLink; include/pf_router.h)
    COMM_HANDLE_BYTES = 128

    def comm_export(self) -> bytes:
        buf = (C.c_ubyte * self.COMM_HANDLE_BYTES)()
        self._ck(self.lib.pf_comm_export(self._h, buf))
        return bytes(buf)

    def comm_init(self, all_handles: bytes):
        assert len(all_handles) == self.COMM_HANDLE_BYTES * self.config.nranks
        buf = (C.c_ubyte * len(all_handles)).from_buffer_copy(all_handles)
        self._ck(self.lib.pf_comm_init(self._h, buf))

    def comm_exchange(self):
        self._ck(self.lib.pf_comm_exchange(self._h))

    def comm_gather_delays(self):
        self._ck(self.lib.pf_comm_gather_delays(self._h))

    def comm_abort(self):
        self.lib.pf_comm_abort(self._h)

    def run(self, sta: "Optional[StaFn]" = None, dsta: "Optional[Sta]" = None):
        """pf_route_run: the whole PathFinder loop on this router (one GPU, or one rank of several after comm_init)
        with one host-device synchronisation per iteration.  Returns (success, iterations, stats[iterations])."""
        T = self.problem.num_terminals
        cap = max(int(self.problem.opts["max_router_iterations"]), 1)
        stats = (IterStats * cap)()
        it, ok = C.c_int(0), C.c_int(0)

        def _cb(_user, iters_done, nd, crit, cpd):
            delays = np.ctypeslib.as_array(nd, shape=(max(T, 1),))[:T]
            c, d = sta(iters_done, delays)
            np.ctypeslib.as_array(crit, shape=(max(T, 1),))[:T] = np.asarray(c, dtype=np.float32)
            cpd[0] = float(d)

        cb = STA_FN(_cb) if sta is not None else C.cast(None, STA_FN)
        self._ck(self.lib.pf_route_run(self._h, dsta._h if dsta is not None else None, cb, None, stats, cap, C.byref(it), C.byref(ok)))
        n = min(it.value, cap)
        out = np.zeros(n, dtype=pfio.ITER_STATS_DT)
        for i in range(n):
            for f, _ in IterStats._fields_:
                out[f][i] = getattr(stats[i], f)
        return bool(ok.value), int(it.value), out

    def comm_crit_ptr(self) -> int:
        return int(self.lib.pf_comm_crit_ptr(self._h))


class Sta:
    """Device-resident timing graph: the static timing analysis the reference runs on the host between router
    iterations (load_timing_graph_net_delays + do_timing_analysis, route_timing.c:295-309), on the GPU."""

    def __init__(self, graph: "pfio.TimingGraph", problem: pfio.Problem, config: Optional[Config] = None, lib_path: Optional[str] = None):
        self.lib = load_library(lib_path)
        self._g = _TimingGraphHolder(graph)
        self._p = _ProblemHolder(problem)
        self.num_terminals = problem.num_terminals
        cfg = config if config is not None else default_config(self.lib)
        h = C.c_void_p()
        rc = self.lib.pf_sta_create(C.byref(self._g.c), C.byref(self._p.c), C.byref(cfg), C.byref(h))
        if rc != PF_OK:
            raise RouterError(rc, self.lib.pf_last_error().decode())
        self._h = h

    def analyze(self, net_delay: np.ndarray):
        """net_delay[num_terminals] -> (timing_criticality[num_terminals], critical path delay in ns)."""
        d = np.ascontiguousarray(net_delay, dtype=np.float32)
        crit = np.zeros(self.num_terminals, np.float32)
        cpd = C.c_float(0)
        rc = self.lib.pf_sta_analyze(self._h, d.ctypes.data, crit.ctypes.data, C.byref(cpd))
        if rc != PF_OK:
            raise RouterError(rc, self.lib.pf_last_error().decode())
        return crit, float(cpd.value)

    def analyze_final(self, net_delay: np.ndarray):
        """The analysis of the finished routing (do_timing_analysis(.., is_final_analysis = TRUE), reference base/stats.c:155-164):
        net_delay[num_terminals] -> (slack[num_terminals] with 1e30 where no analysed path passes, timing_criticality, cpd in ns)."""
        d = np.ascontiguousarray(net_delay, dtype=np.float32)
        slack = np.zeros(self.num_terminals, np.float32)
        crit = np.zeros(self.num_terminals, np.float32)
        cpd = C.c_float(0)
        self.lib.pf_sta_analyze_final.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        rc = self.lib.pf_sta_analyze_final(self._h, d.ctypes.data, slack.ctypes.data, crit.ctypes.data, C.byref(cpd))
        if rc != PF_OK:
            raise RouterError(rc, self.lib.pf_last_error().decode())
        return slack, crit, float(cpd.value)

    def analyze_device(self, dev_net_delay: int, dev_crit: int) -> float:
        """Device pointers in and out (e.g. Router.comm_net_delay_ptr() / comm_crit_ptr()); returns the cpd in ns."""
        cpd = C.c_float(0)
        rc = self.lib.pf_sta_analyze_device(self._h, C.c_void_p(dev_net_delay), C.c_void_p(dev_crit), C.byref(cpd))
        if rc != PF_OK:
            raise RouterError(rc, self.lib.pf_last_error().decode())
        return float(cpd.value)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.pf_sta_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_config(lib=None, **kw) -> Config:
    lib = lib or load_library()
    c = Config()
    lib.pf_config_default(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


StaFn = Callable[[int, np.ndarray], "tuple[np.ndarray, float]"]


def try_timing_driven_route(problem: pfio.Problem, config: Optional[Config] = None, sta: Optional[StaFn] = None,
                            lib_path: Optional[str] = None, timing_graph: "Optional[pfio.TimingGraph]" = None) -> pfio.Result:
    """The whole PathFinder loop on one GPU (reference route_timing.c:85-343).

    ``sta(iters_done, net_delay) -> (timing_criticality[num_terminals], crit_path_delay)`` stands in for the
    reference's load_timing_graph_net_delays + do_timing_analysis (route_timing.c:295-309); with ``timing_graph``
    the analysis runs on the device instead (pf_try_timing_driven_route_sta) and nothing is copied per iteration."""
    lib = load_library(lib_path)
    holder = _ProblemHolder(problem)
    if config is None:
        config = default_config(lib)
    T = problem.num_terminals
    if timing_graph is not None:
        tg = _TimingGraphHolder(timing_graph)
        cr = _Result()
        rc = lib.pf_try_timing_driven_route_sta(C.byref(holder.c), C.byref(tg.c), C.byref(config), C.byref(cr))
        if rc != PF_OK:
            raise RouterError(rc, lib.pf_last_error().decode())
        return _take_result(lib, cr, T)

    def _cb(_user, iters_done, nd, crit, cpd):
        delays = np.ctypeslib.as_array(nd, shape=(max(T, 1),))[:T]
        c, d = sta(iters_done, delays)
        np.ctypeslib.as_array(crit, shape=(max(T, 1),))[:T] = np.asarray(c, dtype=np.float32)
        cpd[0] = float(d)

    cb = STA_FN(_cb) if sta is not None else C.cast(None, STA_FN)
    cr = _Result()
    rc = lib.pf_try_timing_driven_route(C.byref(holder.c), C.byref(config), cb, None, C.byref(cr))
    if rc != PF_OK:
        raise RouterError(rc, lib.pf_last_error().decode())
    return _take_result(lib, cr, T)


def replay_sta(golden: pfio.Result) -> StaFn:
    """STA stand-in that replays criticalities recorded from a reference run (iter_crit row k is what
    the reference's STA produced after k iterations)."""

    def fn(iters_done: int, _delay: np.ndarray):
        k = min(iters_done, golden.iter_crit.shape[0] - 1)
        cpd = float(golden.iter_stats["crit_path_delay"][iters_done - 1]) if 1 <= iters_done <= len(golden.iter_stats) else 0.0
        return golden.iter_crit[k], cpd

    return fn


def generate_grid_nets(lib_path: Optional[str] = None, **kw):
    """(nets-only Problem, GenParams) of the same synthetic problem: nets, boxes, tables and options on the host, NO rr graph
    arrays — ``Router(nets, cfg, generated=params)`` builds the graph on the device (pf_router_create_generated)."""
    return generate_grid_problem(lib_path, nets_only=True, **kw)


def generate_grid_problem(lib_path: Optional[str] = None, nets_only: bool = False, **kw):
    """Synthetic k6_N10-style grid + random nets (pf_gen_grid_problem).  Keyword arguments override
    pf_gen_params_default (nx, ny, W, L, num_nets, sinks_per_net, window, seed, ...)."""
    lib = load_library(lib_path)
    g = GenParams()
    lib.pf_gen_params_default(C.byref(g))
    for k, v in kw.items():
        setattr(g, k, v)
    cp = _Problem()
    rc = (lib.pf_gen_grid_nets if nets_only else lib.pf_gen_grid_problem)(C.byref(g), C.byref(cp))
    if rc != PF_OK:
        raise RouterError(rc, "pf_gen_grid_problem failed")
    counts = {"N": cp.num_nodes, "N1": cp.num_nodes + 1, "E": cp.num_edges, "S": cp.num_switches, "I": cp.num_indexed,
              "n": cp.num_nets, "n1": cp.num_nets + 1, "n4": 4 * cp.num_nets, "T": cp.num_terminals,
              "G": cp.num_opin_groups}
    arrs = {}
    for name, dt, c in pfio._PROB_FIELDS:
        dt = np.dtype(dt)
        nbytes = counts[c] * dt.itemsize
        ptr = getattr(cp, name)
        if not ptr:
            nbytes = 0                        # nets-only problem: the rr graph arrays are NULL
        buf = (C.c_char * nbytes).from_address(ptr) if nbytes else b""
        arrs[name] = np.frombuffer(bytes(buf) if nbytes else b"", dtype=dt).copy()
    arrs["net_bb"] = arrs["net_bb"].reshape(cp.num_nets, 4)
    opts = np.zeros((), dtype=pfio.OPTS_DT)
    for f in _Opts._fields_:
        opts[f[0]] = getattr(cp.opts, f[0])
    p = pfio.Problem(nx=cp.nx, ny=cp.ny, opts=opts, **arrs)
    closed_form_nodes = int(cp.num_nodes)
    lib.pf_problem_free(C.byref(cp))
    if nets_only:
        p.gen_num_nodes = closed_form_nodes   # node count of the graph the device generator will build
        return p, g
    return p
