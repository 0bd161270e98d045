"""Shared fixtures.  GPU tests are marked @pytest.mark.gpu; everything else runs on a CPU-only box."""
import lzma
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("%s failed:\n%s\n%s" % (" ".join(cmd), r.stdout[-2000:], r.stderr[-2000:]))


@pytest.fixture(scope="session")
def oracle_cli():
    """The CPU restatement of the reference router (test infrastructure), built on demand."""
    _run(["make", "-s"], os.path.join(ROOT, "oracle"))
    return os.path.join(ROOT, "oracle", "_build", "pf_oracle_cli")


@pytest.fixture(scope="session")
def oracle_lib():
    _run(["make", "-s"], os.path.join(ROOT, "oracle"))
    return os.path.join(ROOT, "oracle", "_build", "libpf_oracle.so")


@pytest.fixture(scope="session")
def emu_lib():
    """Device + host router sources compiled against the fiber warp emulator (tests/emu)."""
    _run(["make", "-s"], os.path.join(ROOT, "tests", "emu"))
    return os.path.join(ROOT, "tests", "emu", "_build", "libpf_router_emu.so")


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library.  Building needs nvcc only (cross-compiles without a GPU)."""
    from parallel_eda_b200.build import build
    return build()


@pytest.fixture(scope="session")
def ref_bin():
    """The UNMODIFIED reference router, if it was built here (oracle/_ref travels to the GPU box)."""
    p = os.path.join(ROOT, "oracle", "_ref", "vpr_ref")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/vpr_ref not built (needs /root/reference; run `make -C oracle ref`)")
    return p


@pytest.fixture(scope="session")
def unxz(tmp_path_factory):
    """Decompress a committed golden fixture for tools that read the raw container."""
    cache = {}
    d = tmp_path_factory.mktemp("golden")

    def get(name):
        if name not in cache:
            out = os.path.join(str(d), name)
            with lzma.open(os.path.join(GOLDEN, name + ".xz"), "rb") as f, open(out, "wb") as g:
                g.write(f.read())
            cache[name] = out
        return cache[name]

    return get
