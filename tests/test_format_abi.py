"""Host plumbing on a CPU-only box: container round trips, and that the product library loads and
exports every entry point include/*.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from parallel_eda_b200 import pfio, router

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_problem_roundtrip_python_and_c(unxz, oracle_cli, tmp_path):
    src = unxz("toy_w64.pfp")
    p = pfio.read_problem(src)
    out = str(tmp_path / "rt.pfp")
    pfio.write_problem(out, p)
    assert open(out, "rb").read() == open(src, "rb").read()      # byte-identical container
    q = pfio.read_problem(os.path.join(ROOT, "tests", "golden", "toy_w64.pfp.xz"))
    assert q.num_nodes == p.num_nodes == 5436 and q.num_edges == 32380 and q.num_nets == 294


def test_result_roundtrip(unxz, tmp_path):
    src = unxz("toy_w64.pfr")
    r = pfio.read_result(src)
    out = str(tmp_path / "rt.pfr")
    pfio.write_result(out, r)
    assert open(out, "rb").read() == open(src, "rb").read()


def _declared_functions():
    names = []
    for h in ("pf_router.h", "pf_file.h", "pf_gen.h", "pf_text.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(pf_[a-z_0-9]+)\s*\(", text)
    return sorted(set(n for n in names if n not in ("pf_sta_fn",)))


def test_cuda_library_exports_every_declared_symbol(cuda_lib):
    lib = ctypes.CDLL(cuda_lib)
    names = _declared_functions()
    assert "pf_try_timing_driven_route" in names and "pf_route_iteration" in names and len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.pf_backend_name.restype = ctypes.c_char_p
    assert lib.pf_backend_name() == b"cuda:sm_100a"


def test_cuda_library_is_sm100a_only(cuda_lib):
    out = subprocess.run(["cuobjdump", "-lelf", cuda_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu(cuda_lib):
    """On a box without a CUDA device the product path must fail loudly, never route on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    p = pfio.read_problem(os.path.join(ROOT, "tests", "golden", "toy_w64.pfp.xz"))
    with pytest.raises(router.RouterError) as e:
        router.Router(p, lib_path=cuda_lib)
    assert e.value.code == -5 and "CUDA" in str(e.value)


def test_product_library_does_not_link_the_oracle_or_emulator(cuda_lib):
    syms = subprocess.run(["nm", "-D", cuda_lib], capture_output=True, text=True).stdout
    assert "pf_oracle" not in syms and "pf_emu" not in syms


def test_timing_graph_container_python_and_c(unxz, cuda_lib, tmp_path):
    """PFTIMG01 / PFSTAV01: the C reader/writer (pf_file.c, linked into the product library) and pfio agree, and
    pf_timing_graph_check rejects a corrupted graph with a message."""
    import ctypes as C
    from test_sta_golden import _TG
    lib = C.CDLL(cuda_lib)
    lib.pf_timing_graph_read.argtypes = [C.c_char_p, C.POINTER(_TG)]
    lib.pf_timing_graph_write.argtypes = [C.c_char_p, C.POINTER(_TG)]
    lib.pf_timing_graph_free.argtypes = [C.POINTER(_TG)]
    lib.pf_timing_graph_check.argtypes = [C.POINTER(_TG), C.c_void_p, C.c_char_p, C.c_int]
    src = unxz("toy_w64.pftg")
    g = _TG()
    assert lib.pf_timing_graph_read(src.encode(), C.byref(g)) == 0
    out = str(tmp_path / "copy.pftg")
    assert lib.pf_timing_graph_write(out.encode(), C.byref(g)) == 0
    assert open(out, "rb").read() == open(src, "rb").read()
    a = pfio.read_timing_graph(out)
    out2 = str(tmp_path / "copy2.pftg")
    pfio.write_timing_graph(out2, a)
    assert open(out2, "rb").read() == open(src, "rb").read()
    p = pfio.read_problem(unxz("toy_w64.pfp"))
    net_ptr = np.ascontiguousarray(p.net_ptr, dtype=np.int32)
    msg = C.create_string_buffer(256)
    assert lib.pf_timing_graph_check(C.byref(g), net_ptr.ctypes.data, msg, 256) == 0
    C.cast(g.edge_to, C.POINTER(C.c_int32))[3] = -5
    assert lib.pf_timing_graph_check(C.byref(g), net_ptr.ctypes.data, msg, 256) != 0 and b"tedge" in msg.value
    lib.pf_timing_graph_free(C.byref(g))
    v = pfio.read_sta_vectors(unxz("toy_w64.pfsta"))
    assert v.net_delay.shape == v.crit.shape == (20, p.num_terminals) and v.cpd.shape == (20,)
