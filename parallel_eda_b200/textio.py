"""VPR's .route / .place text files through the C-ABI of include/pf_text.h.

Python mirror of the native writers / readers in parallel_eda_b200/csrc/pf_text.c:

    write_route   print_route  reference vpr/SRC/route/route_common.c:1322-1417   (byte-identical output)
    read_route    -            a .route file back into trace arrays (VPR 7 cannot do this)
    write_place   print_place  reference vpr/SRC/base/read_place.c:266-293        (byte-identical output)
    read_place    read_place   reference vpr/SRC/base/read_place.c:15-139         (hashed block lookup)

`Names` is the numpy image of pf_names (container PFNAME01): net and block names, IO tiles, block locations and
the pins of global nets — everything the text files need that the flat pf_problem does not carry.
Host plumbing only; the formatting and parsing run in libpf_router.so.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import struct
from typing import List, Optional, Tuple

import numpy as np

from . import pfio
from . import router as _rt

NAME_MAGIC = b"PFNAME01"


class _Names(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("num_nets", C.c_int32),
                ("net_name_ptr", C.c_void_p), ("net_name_chars", C.c_void_p), ("tile_is_io", C.c_void_p),
                ("num_blocks", C.c_int32),
                ("block_name_ptr", C.c_void_p), ("block_name_chars", C.c_void_p),
                ("block_x", C.c_void_p), ("block_y", C.c_void_p), ("block_z", C.c_void_p),
                ("gpin_ptr", C.c_void_p), ("gpin_block", C.c_void_p), ("gpin_class", C.c_void_p)]


@dataclasses.dataclass
class Names:
    nx: int
    ny: int
    net_name_ptr: np.ndarray      # int32 [num_nets+1]
    net_name_chars: np.ndarray    # uint8
    tile_is_io: np.ndarray        # uint8 [(nx+2)*(ny+2)], index x*(ny+2)+y
    block_name_ptr: np.ndarray    # int32 [num_blocks+1]
    block_name_chars: np.ndarray  # uint8
    block_x: np.ndarray
    block_y: np.ndarray
    block_z: np.ndarray
    gpin_ptr: np.ndarray          # int32 [num_nets+1]
    gpin_block: np.ndarray
    gpin_class: np.ndarray

    @property
    def num_nets(self) -> int:
        return len(self.net_name_ptr) - 1

    @property
    def num_blocks(self) -> int:
        return len(self.block_name_ptr) - 1

    def net_name(self, i: int) -> str:
        return bytes(self.net_name_chars[self.net_name_ptr[i]:self.net_name_ptr[i + 1]]).decode()

    def block_name(self, i: int) -> str:
        return bytes(self.block_name_chars[self.block_name_ptr[i]:self.block_name_ptr[i + 1]]).decode()

    @staticmethod
    def build(nx: int, ny: int, net_names: List[str], tile_is_io: np.ndarray,
              blocks: List[Tuple[str, int, int, int]] = (), global_pins: Optional[dict] = None) -> "Names":
        """From Python lists: blocks = [(name, x, y, z)], global_pins = {inet: [(block, pin_class), ...]}."""
        def pack(strings):
            ptr = np.zeros(len(strings) + 1, dtype=np.int32)
            enc = [s.encode() for s in strings]
            ptr[1:] = np.cumsum([len(e) for e in enc], dtype=np.int64)
            return ptr, np.frombuffer(b"".join(enc), dtype=np.uint8).copy()
        nptr, nch = pack(list(net_names))
        bptr, bch = pack([b[0] for b in blocks])
        gptr = np.zeros(len(net_names) + 1, dtype=np.int32)
        gb, gc = [], []
        for i in range(len(net_names)):
            for blk, cls in (global_pins or {}).get(i, []):
                gb.append(blk)
                gc.append(cls)
            gptr[i + 1] = len(gb)
        return Names(nx, ny, nptr, nch, np.ascontiguousarray(tile_is_io, dtype=np.uint8).reshape(-1), bptr, bch,
                     np.array([b[1] for b in blocks], dtype=np.int32), np.array([b[2] for b in blocks], dtype=np.int32),
                     np.array([b[3] for b in blocks], dtype=np.int32), gptr, np.array(gb, dtype=np.int32),
                     np.array(gc, dtype=np.int32))


_FIELDS = (("net_name_ptr", "<i4"), ("net_name_chars", "u1"), ("tile_is_io", "u1"), ("block_name_ptr", "<i4"),
           ("block_name_chars", "u1"), ("block_x", "<i4"), ("block_y", "<i4"), ("block_z", "<i4"),
           ("gpin_ptr", "<i4"), ("gpin_block", "<i4"), ("gpin_class", "<i4"))


def read_names(path: str) -> Names:
    with pfio._open(path) as f:
        if f.read(8) != NAME_MAGIC:
            raise ValueError("%s: not a PFNAME01 file" % path)
        nx, ny, nnets, nblocks, nchars, bchars, gpins = struct.unpack("<16i", f.read(64))[:7]
        counts = (nnets + 1, nchars, (nx + 2) * (ny + 2), nblocks + 1, bchars, nblocks, nblocks, nblocks, nnets + 1,
                  gpins, gpins)
        arrays = []
        for (name, dt), cnt in zip(_FIELDS, counts):
            dt = np.dtype(dt)
            buf = f.read(cnt * dt.itemsize)
            if len(buf) != cnt * dt.itemsize:
                raise ValueError("%s: truncated at %s" % (path, name))
            arrays.append(np.frombuffer(buf, dtype=dt).copy())
    return Names(nx, ny, *arrays)


def write_names(path: str, n: Names) -> None:
    hdr = [n.nx, n.ny, n.num_nets, n.num_blocks, len(n.net_name_chars), len(n.block_name_chars), len(n.gpin_block)]
    with open(path, "wb") as f:
        f.write(NAME_MAGIC)
        f.write(struct.pack("<16i", *(hdr + [0] * 9)))
        for name, dt in _FIELDS:
            f.write(np.ascontiguousarray(getattr(n, name), dtype=np.dtype(dt)).tobytes())


class _NamesHolder:
    """Keeps contiguous numpy arrays alive while a C pf_names points into them."""

    def __init__(self, n: Names):
        self.arrays = {}
        c = _Names()
        c.nx, c.ny, c.num_nets, c.num_blocks = n.nx, n.ny, n.num_nets, n.num_blocks
        for name, dt in _FIELDS:
            a = np.ascontiguousarray(getattr(n, name), dtype=np.dtype(dt))
            if a.size == 0:
                a = np.zeros(1, dtype=np.dtype(dt))     # never hand C a NULL for an empty array
            self.arrays[name] = a
            setattr(c, name, a.ctypes.data)
        self.c = c


def _lib():
    lib = _rt.load_library()
    if not getattr(lib, "_pf_text_bound", False):
        lib.pf_text_error.restype = C.c_char_p
        lib.pf_route_write.argtypes = [C.c_char_p, C.POINTER(_rt._Problem), C.POINTER(_Names), C.POINTER(_rt._Result)]
        lib.pf_route_read.argtypes = [C.c_char_p, C.POINTER(_rt._Problem), C.POINTER(_rt._Result)]
        lib.pf_place_write.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(_Names)]
        lib.pf_place_read.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(_Names), C.POINTER(C.c_int)]
        lib.pf_names_check.argtypes = [C.POINTER(_Names), C.POINTER(_rt._Problem), C.c_char_p, C.c_int]
        lib.pf_names_synthetic.argtypes = [C.POINTER(_rt._Problem), C.POINTER(_Names)]
        lib.pf_names_free.argtypes = [C.POINTER(_Names)]
        lib.pf_names_free.restype = None
        lib.pf_names_write.argtypes = [C.c_char_p, C.POINTER(_Names)]
        lib.pf_names_read.argtypes = [C.c_char_p, C.POINTER(_Names)]
        lib._pf_text_bound = True
    return lib


def _raise(lib, rc: int):
    raise _rt.RouterError(rc, (lib.pf_text_error() or b"").decode())


def check_names(n: Names, p: Optional[pfio.Problem] = None) -> None:
    lib = _lib()
    msg = C.create_string_buffer(256)
    ph = _rt._ProblemHolder(p) if p is not None else None
    rc = lib.pf_names_check(C.byref(_NamesHolder(n).c), C.byref(ph.c) if ph else None, msg, 256)
    if rc != 0:
        raise _rt.RouterError(rc, msg.value.decode())


def _names_from_c(c: _Names) -> Names:
    def arr(ptr, count, dt):
        if count <= 0 or not ptr:
            return np.zeros(0, dtype=dt)
        return np.frombuffer((C.c_char * (count * np.dtype(dt).itemsize)).from_address(ptr), dtype=dt).copy()
    nn, nb = c.num_nets, c.num_blocks
    nptr = arr(c.net_name_ptr, nn + 1, np.int32)
    bptr = arr(c.block_name_ptr, nb + 1, np.int32)
    gptr = arr(c.gpin_ptr, nn + 1, np.int32)
    return Names(c.nx, c.ny, nptr, arr(c.net_name_chars, int(nptr[nn]), np.uint8),
                 arr(c.tile_is_io, (c.nx + 2) * (c.ny + 2), np.uint8), bptr,
                 arr(c.block_name_chars, int(bptr[nb]) if nb else 0, np.uint8),
                 arr(c.block_x, nb, np.int32), arr(c.block_y, nb, np.int32), arr(c.block_z, nb, np.int32), gptr,
                 arr(c.gpin_block, int(gptr[nn]), np.int32), arr(c.gpin_class, int(gptr[nn]), np.int32))


def synthetic_names(p: pfio.Problem) -> Names:
    """pf_names_synthetic: nets "n<i>", the IO ring of a VPR grid, no blocks (for generated fabrics)."""
    lib = _lib()
    ph = _rt._ProblemHolder(p)
    c = _Names()
    rc = lib.pf_names_synthetic(C.byref(ph.c), C.byref(c))
    if rc != 0:
        _raise(lib, rc)
    try:
        return _names_from_c(c)
    finally:
        lib.pf_names_free(C.byref(c))


def write_route(path: str, p: pfio.Problem, n: Names, r: pfio.Result) -> None:
    """print_route (reference route_common.c:1322): the .route file of routing r."""
    lib = _lib()
    ph, nh = _rt._ProblemHolder(p), _NamesHolder(n)
    tp = np.ascontiguousarray(r.trace_ptr, dtype=np.int32)
    tn = np.ascontiguousarray(r.trace_node, dtype=np.int32)
    ts = np.ascontiguousarray(r.trace_switch, dtype=np.int16)
    cr = _rt._Result()
    cr.num_nets = len(tp) - 1
    cr.trace_ptr = tp.ctypes.data_as(C.POINTER(C.c_int32))
    cr.trace_node = tn.ctypes.data_as(C.POINTER(C.c_int32))
    cr.trace_switch = ts.ctypes.data_as(C.POINTER(C.c_int16))
    rc = lib.pf_route_write(path.encode(), C.byref(ph.c), C.byref(nh.c), C.byref(cr))
    if rc != 0:
        _raise(lib, rc)


def read_route(path: str, p: pfio.Problem) -> pfio.Result:
    """A .route file as trace arrays (trace_ptr / trace_node / trace_switch, wirelength, serial number); the
    remaining Result fields are empty.  Raises RouterError(PF_EFORMAT) when the file does not fit the problem."""
    lib = _lib()
    ph = _rt._ProblemHolder(p)
    cr = _rt._Result()
    rc = lib.pf_route_read(path.encode(), C.byref(ph.c), C.byref(cr))
    if rc != 0:
        _raise(lib, rc)
    return _rt._take_result(lib, cr, p.num_terminals, with_stats=False)


def write_place(path: str, net_file: str, arch_file: str, n: Names) -> None:
    """print_place (reference read_place.c:266)."""
    lib = _lib()
    rc = lib.pf_place_write(path.encode(), net_file.encode(), arch_file.encode(), C.byref(_NamesHolder(n).c))
    if rc != 0:
        _raise(lib, rc)


def read_place(path: str, n: Names, net_file: Optional[str] = None, arch_file: Optional[str] = None) -> int:
    """read_place (reference read_place.c:15): sets n.block_x/y/z in place, returns the number of placed blocks."""
    lib = _lib()
    nh = _NamesHolder(n)
    placed = C.c_int(0)
    rc = lib.pf_place_read(path.encode(), net_file.encode() if net_file else None,
                           arch_file.encode() if arch_file else None, C.byref(nh.c), C.byref(placed))
    if rc != 0:
        _raise(lib, rc)
    for k in ("block_x", "block_y", "block_z"):
        setattr(n, k, nh.arrays[k][:n.num_blocks].copy())
    return placed.value


# ---- the packed netlist (.net): pf_net_read, include/pf_text.h
class _Netlist(C.Structure):
    _fields_ = [("num_blocks", C.c_int32),
                ("block_name_ptr", C.POINTER(C.c_int32)), ("block_name_chars", C.POINTER(C.c_char)),
                ("block_type_ptr", C.POINTER(C.c_int32)), ("block_type_chars", C.POINTER(C.c_char)),
                ("block_pin_ptr", C.POINTER(C.c_int32)), ("block_pin_net", C.POINTER(C.c_int32)),
                ("block_pin_kind", C.POINTER(C.c_uint8)),
                ("num_nets", C.c_int32),
                ("net_name_ptr", C.POINTER(C.c_int32)), ("net_name_chars", C.POINTER(C.c_char)),
                ("net_ptr", C.POINTER(C.c_int32)), ("net_block", C.POINTER(C.c_int32)), ("net_block_pin", C.POINTER(C.c_int32)),
                ("net_is_global", C.POINTER(C.c_uint8))]


@dataclasses.dataclass
class Netlist:
    """block[] and clb_net[] as the reference's read_netlist (base/read_netlist.c:74) leaves them for place and route."""
    block_names: List[str]
    block_types: List[str]           # block[i].type->name
    block_pin_ptr: np.ndarray        # int32[num_blocks + 1]: pins of one instance of the type
    block_pin_net: np.ndarray        # int32: block[i].nets[pin], -1 = OPEN
    block_pin_kind: np.ndarray       # uint8: 0 input, 1 output, 2 clock
    net_names: List[str]
    net_ptr: np.ndarray              # int32[num_nets + 1]; terminal 0 of a net is its driver
    net_block: np.ndarray            # clb_net[i].node_block[]
    net_block_pin: np.ndarray        # clb_net[i].node_block_pin[]
    net_is_global: np.ndarray        # uint8

    @property
    def num_blocks(self) -> int:
        return len(self.block_names)

    @property
    def num_nets(self) -> int:
        return len(self.net_names)


def read_netlist(path: str) -> Netlist:
    """read_netlist (reference base/read_netlist.c:74-244, load_external_nets_and_cb :836-984) without the architecture."""
    lib = _lib()
    if not getattr(lib, "_pf_net_bound", False):
        lib.pf_net_read.argtypes = [C.c_char_p, C.POINTER(_Netlist)]
        lib.pf_netlist_free.argtypes = [C.POINTER(_Netlist)]
        lib.pf_netlist_free.restype = None
        lib._pf_net_bound = True
    c = _Netlist()
    rc = lib.pf_net_read(path.encode(), C.byref(c))
    if rc != 0:
        _raise(lib, rc)
    try:
        def arr(ptr, count, dt):
            return np.ctypeslib.as_array(ptr, shape=(max(count, 1),))[:count].astype(dt, copy=True) if count else np.zeros(0, dt)

        def strings(ptr, chars, count):
            off = arr(ptr, count + 1, np.int32)
            raw = C.string_at(chars, int(off[-1])) if count and off[-1] else b""
            return [raw[off[i]:off[i + 1]].decode() for i in range(count)]
        nb, nn = c.num_blocks, c.num_nets
        bpp = arr(c.block_pin_ptr, nb + 1, np.int32)
        npt = arr(c.net_ptr, nn + 1, np.int32)
        return Netlist(strings(c.block_name_ptr, c.block_name_chars, nb), strings(c.block_type_ptr, c.block_type_chars, nb),
                       bpp, arr(c.block_pin_net, int(bpp[-1]), np.int32), arr(c.block_pin_kind, int(bpp[-1]), np.uint8),
                       strings(c.net_name_ptr, c.net_name_chars, nn), npt, arr(c.net_block, int(npt[-1]), np.int32),
                       arr(c.net_block_pin, int(npt[-1]), np.int32), arr(c.net_is_global, nn, np.uint8))
    finally:
        lib.pf_netlist_free(C.byref(c))
