"""The C-ABI library: loads, exports every symbol include/rcz.h declares, and refuses to run without a GPU."""
import ctypes as C
import importlib
import os
import re

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "rcz.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rcz_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    abi = importlib.import_module("rust-compress_b200._abi")
    assert _declared() == sorted(abi.SIGNATURES)


def test_rust_shim_declares_every_symbol():
    """rust/rcz_sys.rs (uncompiled: no rustc in this image) must carry one `extern "C"` declaration per function of rcz.h, with the
    same number of parameters."""
    rs = open(os.path.join(ROOT, "rust", "rcz_sys.rs")).read()
    ext = rs[rs.index('extern "C" {'): rs.index("\n}\n", rs.index('extern "C" {'))]
    decl = dict((m.group(1), m.group(2)) for m in re.finditer(r"pub fn (rcz_[a-z0-9_]+)\(([^)]*)\)", ext))
    assert sorted(decl) == _declared()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "rcz.h")).read(), flags=re.S)
    for name, args in decl.items():
        m = re.search(r"\b%s\s*\(([^)]*)\)" % name, hdr)
        c_args = [a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]
        rs_args = [a for a in args.split(",") if a.strip()]
        assert len(c_args) == len(rs_args), name


def test_librcz_exports_every_declared_symbol():
    build = importlib.import_module("rust-compress_b200.build")
    lib = C.CDLL(build.build())
    for name in _declared():
        assert hasattr(lib, name), "librcz.so does not export " + name


def test_product_library_is_cuda_only():
    """librcz.so must be the nvcc build (no emulation code) and must fail loudly without a device."""
    abi = importlib.import_module("rust-compress_b200._abi")
    lib = abi.load()
    assert b"sm_100a" in lib.rcz_build_info()
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert lib.rcz_ctx_create(0, 0, C.byref(h)) == abi.E_NO_DEVICE
        rcz = importlib.import_module("rust-compress_b200")
        with pytest.raises(rcz.RczError):
            rcz.Context(device=0)


def test_sass_has_tma_bulk_copy():
    """The LZ4 kernel stages its input window with cp.async.bulk: SASS must show UBLKCP (B200_PROFILING.md)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump") and not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    build = importlib.import_module("rust-compress_b200.build")
    sass = subprocess.run([exe, "-sass", build.build()], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in sass and "UBLKCP" in sass
