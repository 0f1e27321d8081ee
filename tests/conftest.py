import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFDATA = "/root/reference/src/data"   # only present in the authoring container; never required


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _pkg():
    return importlib.import_module("rust-compress_b200")


@pytest.fixture(scope="session")
def rcz():
    return _pkg()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def gen():
    from tools import gen as g
    g.build()
    return g


@pytest.fixture(scope="session")
def emu_ctx(rcz):
    """Context on librcz_emu.so: the product sources compiled against the CPU SIMT emulation (tests only)."""
    build = importlib.import_module("rust-compress_b200.build")
    build.build_emu()
    return rcz.Context(emu=True)


@pytest.fixture(scope="session")
def gpu_ctx(rcz):
    """Context on the real librcz.so; fails loudly (no fallback) when the library or the GPU is missing."""
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    assert os.path.exists(importlib.import_module("rust-compress_b200._abi").LIB_PATH), "librcz.so missing: run __graft_entry__.build()"
    ctx = rcz.Context(device=0)
    ctx.set_stream(torch.cuda.current_stream())
    return ctx


def golden(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()
