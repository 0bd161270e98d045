"""numpy reader/writer for the flat routing containers defined in include/pf_file.h.

Host plumbing only (no compute): the same byte layout as parallel_eda_b200/csrc/pf_file.c, so
fixtures written by the reference-side exporter (oracle/ref_build/harness.cxx), by the native
generator and by Python are interchangeable.  Field names follow include/pf_types.h, which in
turn cites the reference globals each array flattens (vpr/SRC/base/globals.c:48-97).
"""
from __future__ import annotations

import dataclasses
import lzma
import struct
from typing import Optional

import numpy as np

PROB_MAGIC = b"PFPROB01"
RSLT_MAGIC = b"PFRSLT01"

SWITCH_DT = np.dtype([("buffered", "<i4"), ("R", "<f4"), ("Cin", "<f4"), ("Cout", "<f4"), ("Tdel", "<f4")])
INDEXED_DT = np.dtype([
    ("base_cost", "<f4"), ("saved_base_cost", "<f4"), ("ortho_cost_index", "<i4"), ("seg_index", "<i4"),
    ("inv_length", "<f4"), ("T_linear", "<f4"), ("T_quadratic", "<f4"), ("C_load", "<f4")])
OPTS_DT = np.dtype([
    ("first_iter_pres_fac", "<f4"), ("initial_pres_fac", "<f4"), ("pres_fac_mult", "<f4"), ("acc_fac", "<f4"),
    ("bend_cost", "<f4"), ("astar_fac", "<f4"), ("max_criticality", "<f4"), ("criticality_exp", "<f4"),
    ("max_router_iterations", "<i4"), ("timing_analysis_enabled", "<i4"), ("bb_factor", "<i4"), ("router_algorithm", "<i4")])
ITER_STATS_DT = np.dtype([
    ("overused_nodes", "<i4"), ("nets_routed", "<i4"), ("heap_pushes", "<i8"), ("heap_pops", "<i8"),
    ("edge_visits", "<i8"), ("pres_fac", "<f4"), ("crit_path_delay", "<f4")])
assert SWITCH_DT.itemsize == 20 and INDEXED_DT.itemsize == 32 and OPTS_DT.itemsize == 48 and ITER_STATS_DT.itemsize == 40

SOURCE, SINK, IPIN, OPIN, CHANX, CHANY = range(6)
OPEN = -1


def default_opts(timing: bool = False) -> np.ndarray:
    """VPR defaults for the timing-driven router (reference base/SetupVPR.c:330-605)."""
    o = np.zeros((), dtype=OPTS_DT)
    o["first_iter_pres_fac"] = 0.5
    o["initial_pres_fac"] = 0.5
    o["pres_fac_mult"] = 1.3
    o["acc_fac"] = 1.0
    o["bend_cost"] = 0.0
    o["astar_fac"] = 1.2
    o["max_criticality"] = 0.99
    o["criticality_exp"] = 1.0
    o["max_router_iterations"] = 50
    o["timing_analysis_enabled"] = 1 if timing else 0
    o["bb_factor"] = 3
    return o


@dataclasses.dataclass
class Problem:
    nx: int
    ny: int
    xlow: np.ndarray
    ylow: np.ndarray
    xhigh: np.ndarray
    yhigh: np.ndarray
    ptc_num: np.ndarray
    cost_index: np.ndarray
    capacity: np.ndarray
    type: np.ndarray
    direction: np.ndarray
    R: np.ndarray
    C: np.ndarray
    row_ptr: np.ndarray
    edge_to: np.ndarray
    edge_sw: np.ndarray
    switches: np.ndarray
    indexed: np.ndarray
    net_ptr: np.ndarray
    net_terminals: np.ndarray
    net_is_global: np.ndarray
    net_bb: np.ndarray
    opin_group_source: np.ndarray
    opin_group_count: np.ndarray
    opts: np.ndarray

    @property
    def num_nodes(self) -> int:
        return int(self.xlow.shape[0])

    @property
    def num_edges(self) -> int:
        return int(self.edge_to.shape[0])

    @property
    def num_nets(self) -> int:
        return int(self.net_is_global.shape[0])

    @property
    def num_terminals(self) -> int:
        return int(self.net_terminals.shape[0])

    def routed_nets(self) -> np.ndarray:
        return np.nonzero(self.net_is_global == 0)[0]


_PROB_FIELDS = [
    ("xlow", "<i2", "N"), ("ylow", "<i2", "N"), ("xhigh", "<i2", "N"), ("yhigh", "<i2", "N"),
    ("ptc_num", "<i2", "N"), ("cost_index", "<i2", "N"), ("capacity", "<i2", "N"), ("type", "u1", "N"),
    ("direction", "u1", "N"), ("R", "<f4", "N"), ("C", "<f4", "N"), ("row_ptr", "<i4", "N1"),
    ("edge_to", "<i4", "E"), ("edge_sw", "<i2", "E"), ("switches", SWITCH_DT, "S"), ("indexed", INDEXED_DT, "I"),
    ("net_ptr", "<i4", "n1"), ("net_terminals", "<i4", "T"), ("net_is_global", "u1", "n"), ("net_bb", "<i4", "n4"),
    ("opin_group_source", "<i4", "G"), ("opin_group_count", "<i4", "G")]


def _open(path: str):
    """Fixtures are committed xz-compressed; both forms read the same."""
    return lzma.open(path, "rb") if path.endswith(".xz") else open(path, "rb")


def read_problem(path: str) -> Problem:
    with _open(path) as f:
        if f.read(8) != PROB_MAGIC:
            raise ValueError("%s: not a PFPROB01 file" % path)
        hdr = struct.unpack("<16i", f.read(64))
        nx, ny, N, E, S, I, n, T, G, optsz = hdr[:10]
        if optsz != OPTS_DT.itemsize:
            raise ValueError("router opts size mismatch")
        opts = np.frombuffer(f.read(OPTS_DT.itemsize), dtype=OPTS_DT)[0].copy()
        counts = {"N": N, "N1": N + 1, "E": E, "S": S, "I": I, "n": n, "n1": n + 1, "n4": 4 * n, "T": T, "G": G}
        arrs = {}
        for name, dt, c in _PROB_FIELDS:
            dt = np.dtype(dt)
            nbytes = counts[c] * dt.itemsize
            buf = f.read(nbytes)
            if len(buf) != nbytes:
                raise ValueError("%s: truncated at %s" % (path, name))
            arrs[name] = np.frombuffer(buf, dtype=dt).copy()
        arrs["net_bb"] = arrs["net_bb"].reshape(n, 4)
    return Problem(nx=nx, ny=ny, opts=opts, **arrs)


def write_problem(path: str, p: Problem) -> None:
    N, E, n = p.num_nodes, p.num_edges, p.num_nets
    hdr = [p.nx, p.ny, N, E, len(p.switches), len(p.indexed), n, p.num_terminals, len(p.opin_group_source),
           OPTS_DT.itemsize] + [0] * 6
    with open(path, "wb") as f:
        f.write(PROB_MAGIC)
        f.write(struct.pack("<16i", *hdr))
        f.write(np.asarray(p.opts, dtype=OPTS_DT).tobytes())
        for name, dt, _ in _PROB_FIELDS:
            f.write(np.ascontiguousarray(getattr(p, name), dtype=np.dtype(dt)).tobytes())


@dataclasses.dataclass
class Result:
    success: int
    iterations: int
    serial_num: int
    total_wirelength: int
    trace_ptr: np.ndarray
    trace_node: np.ndarray
    trace_switch: np.ndarray
    net_delay: np.ndarray
    occ: np.ndarray
    iter_stats: np.ndarray
    iter_crit: Optional[np.ndarray] = None  # [iters][num_terminals]

    def net_trace(self, inet: int):
        a, b = int(self.trace_ptr[inet]), int(self.trace_ptr[inet + 1])
        return self.trace_node[a:b], self.trace_switch[a:b]


def read_result(path: str) -> Result:
    with _open(path) as f:
        if f.read(8) != RSLT_MAGIC:
            raise ValueError("%s: not a PFRSLT01 file" % path)
        hdr = struct.unpack("<16i", f.read(64))
        success, iters, cookie, wl, n, ntrace, T, N, nst, ncrit, stsz = hdr[:11]
        if stsz != ITER_STATS_DT.itemsize:
            raise ValueError("iter stats size mismatch")

        def rd(dt, count):
            dt = np.dtype(dt)
            buf = f.read(count * dt.itemsize)
            if len(buf) != count * dt.itemsize:
                raise ValueError("%s: truncated" % path)
            return np.frombuffer(buf, dtype=dt).copy()

        trace_ptr = rd("<i4", n + 1)
        trace_node = rd("<i4", ntrace)
        trace_switch = rd("<i2", ntrace)
        net_delay = rd("<f4", T)
        occ = rd("<i4", N)
        stats = rd(ITER_STATS_DT, nst)
        crit = rd("<f4", ncrit * T).reshape(ncrit, T) if ncrit else None
    return Result(success, iters, cookie, wl, trace_ptr, trace_node, trace_switch, net_delay, occ, stats, crit)


def write_result(path: str, r: Result) -> None:
    n = len(r.trace_ptr) - 1
    ncrit = 0 if r.iter_crit is None else int(r.iter_crit.shape[0])
    hdr = [r.success, r.iterations, r.serial_num, r.total_wirelength, n, len(r.trace_node), len(r.net_delay),
           len(r.occ), len(r.iter_stats), ncrit, ITER_STATS_DT.itemsize] + [0] * 5
    with open(path, "wb") as f:
        f.write(RSLT_MAGIC)
        f.write(struct.pack("<16i", *hdr))
        f.write(np.ascontiguousarray(r.trace_ptr, dtype="<i4").tobytes())
        f.write(np.ascontiguousarray(r.trace_node, dtype="<i4").tobytes())
        f.write(np.ascontiguousarray(r.trace_switch, dtype="<i2").tobytes())
        f.write(np.ascontiguousarray(r.net_delay, dtype="<f4").tobytes())
        f.write(np.ascontiguousarray(r.occ, dtype="<i4").tobytes())
        f.write(np.ascontiguousarray(r.iter_stats, dtype=ITER_STATS_DT).tobytes())
        if ncrit:
            f.write(np.ascontiguousarray(r.iter_crit, dtype="<f4").tobytes())


# ---------------------------------------------------------------------------------------------------------------
# timing graph / STA golden vectors (include/pf_types.h: pf_timing_graph, pf_sta_vectors)
TIMG_MAGIC = b"PFTIMG01"
STAV_MAGIC = b"PFSTAV01"


@dataclasses.dataclass
class TimingGraph:
    """Flat image of the reference's timing graph (tnode[] / tedge, timing/path_delay.c:328)."""
    edge_ptr: np.ndarray       # [num_tnodes+1] int32
    edge_to: np.ndarray        # [num_tedges] int32
    edge_Tdel: np.ndarray      # [num_tedges] float32
    type: np.ndarray           # [num_tnodes] uint8 (e_tnode_type)
    clock_domain: np.ndarray   # [num_tnodes] int32
    clock_delay: np.ndarray    # [num_tnodes] float32
    level_ptr: np.ndarray      # [num_levels+1] int32
    level_nodes: np.ndarray    # [num_tnodes] int32
    constraint: np.ndarray     # [num_domains, num_domains] float32
    net_driver: np.ndarray     # [num_nets] int32
    # clock-to-flipflop override constraints (pf_types.h): sorted by (tnode, source domain); empty = none
    override_domain: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0, np.int32))
    override_tnode: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0, np.int32))
    override_constraint: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0, np.float32))

    @property
    def num_tnodes(self) -> int:
        return len(self.type)

    @property
    def num_levels(self) -> int:
        return len(self.level_ptr) - 1


@dataclasses.dataclass
class StaVectors:
    net_delay: np.ndarray      # [calls, num_terminals]
    crit: np.ndarray           # [calls, num_terminals]
    cpd: np.ndarray            # [calls] ns


def _rd(f, path, dt, count):
    dt = np.dtype(dt)
    buf = f.read(count * dt.itemsize)
    if len(buf) != count * dt.itemsize:
        raise ValueError("%s: truncated" % path)
    return np.frombuffer(buf, dtype=dt).copy()


def read_timing_graph(path: str) -> TimingGraph:
    with _open(path) as f:
        if f.read(8) != TIMG_MAGIC:
            raise ValueError("%s: not a PFTIMG01 file" % path)
        n, e, lv, c, nets, novr = struct.unpack("<16i", f.read(64))[:6]
        return TimingGraph(_rd(f, path, "<i4", n + 1), _rd(f, path, "<i4", e), _rd(f, path, "<f4", e), _rd(f, path, "u1", n),
                           _rd(f, path, "<i4", n), _rd(f, path, "<f4", n), _rd(f, path, "<i4", lv + 1), _rd(f, path, "<i4", n),
                           _rd(f, path, "<f4", c * c).reshape(c, c), _rd(f, path, "<i4", nets),
                           _rd(f, path, "<i4", novr), _rd(f, path, "<i4", novr), _rd(f, path, "<f4", novr))


def write_timing_graph(path: str, g: TimingGraph) -> None:
    c = int(g.constraint.shape[0])
    hdr = [g.num_tnodes, len(g.edge_to), g.num_levels, c, len(g.net_driver), len(g.override_tnode)] + [0] * 10
    with open(path, "wb") as f:
        f.write(TIMG_MAGIC)
        f.write(struct.pack("<16i", *hdr))
        for a, dt in ((g.edge_ptr, "<i4"), (g.edge_to, "<i4"), (g.edge_Tdel, "<f4"), (g.type, "u1"), (g.clock_domain, "<i4"),
                      (g.clock_delay, "<f4"), (g.level_ptr, "<i4"), (g.level_nodes, "<i4"), (g.constraint, "<f4"), (g.net_driver, "<i4"),
                      (g.override_domain, "<i4"), (g.override_tnode, "<i4"), (g.override_constraint, "<f4")):
            f.write(np.ascontiguousarray(a, dtype=dt).tobytes())


def read_sta_vectors(path: str) -> StaVectors:
    with _open(path) as f:
        if f.read(8) != STAV_MAGIC:
            raise ValueError("%s: not a PFSTAV01 file" % path)
        T, K = struct.unpack("<16i", f.read(64))[:2]
        return StaVectors(_rd(f, path, "<f4", K * T).reshape(K, T), _rd(f, path, "<f4", K * T).reshape(K, T), _rd(f, path, "<f4", K))
