"""C++ host mirrors (rust-compress_b200/host/rcz_stream.hpp) driven by tests/host/test_host.cpp — the reference crate's own
unit tests restated on the Decoder<R>/Encoder<W> mirrors.  Linked against librcz_emu.so here (host logic, no GPU) and
against the real librcz.so under -m gpu."""
import importlib
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

PKG = os.path.join(ROOT, "rust-compress_b200")


def _build(libname, exe):
    src = os.path.join(ROOT, "tests", "host", "test_host.cpp")
    out = os.path.join(ROOT, "tests", "host", exe)
    deps = [src, os.path.join(PKG, "host", "rcz_stream.hpp"), os.path.join(ROOT, "include", "rcz.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"), src, "-o", out,
                               "-L", PKG, "-l:" + libname, "-Wl,-rpath," + PKG])
    return out


def _run(exe):
    r = subprocess.run([exe, GOLDEN], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout)
    assert r.returncode == 0, r.stdout
    assert "PASSED 0" in r.stdout and "FAIL " not in r.stdout


def test_host_mirrors_emu():
    build = importlib.import_module("rust-compress_b200.build")
    build.build_emu()
    _run(_build("librcz_emu.so", "test_host_emu"))


@pytest.mark.gpu
def test_host_mirrors_gpu():
    build = importlib.import_module("rust-compress_b200.build")
    build.build()
    _run(_build("librcz.so", "test_host_gpu"))
