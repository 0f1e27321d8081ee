"""The reference's test application (main.rs) on librcz: archive container (main.rs:166-171), pass chaining order
(main.rs:154-160, 172-178) and round trips of every pass and of stacks of passes."""
import importlib
import os
import subprocess

import pytest

from conftest import ROOT, golden

PKG = os.path.join(ROOT, "rust-compress_b200")
TXT = golden("ref_test.txt")


def _build(libname, exe):
    src = os.path.join(PKG, "host", "rcz_cli.cpp")
    out = os.path.join(ROOT, "tests", "host", exe)
    deps = [src, os.path.join(PKG, "host", "rcz_stream.hpp"), os.path.join(ROOT, "include", "rcz.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"), src, "-o", out,
                               "-L", PKG, "-l:" + libname, "-Wl,-rpath," + PKG])
    return out


def _run(exe, args, data):
    r = subprocess.run([exe] + args, input=data, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()
    return r.stdout


def _check(exe, gen, scale=1):
    data = TXT * 4 * scale + gen.one("hextext", 3, 9000 * scale) + gen.one("lzsyn", 4, 6000 * scale)
    for methods in (["dummy"], ["lz4"], ["ari", "mtf", "bwt"], ["bwt", "mtf", "ari"], ["lz4c", "dummy", "bwt"]):
        arch = _run(exe, ["-block16384"] + methods, data)
        hdr = bytes([0x72, 0x21, 0x63, 0x73, len(methods)]) + b"".join(bytes([len(m)]) + m.encode() for m in methods)   # "r!cs" LE, main.rs:166-171
        assert arch[: len(hdr)] == hdr, methods
        assert _run(exe, ["-d"], arch) == data, methods
    # chaining order: the LAST method sees the input first (main.rs:172-178), so [ari, mtf, bwt] is bwt -> mtf -> ari and compresses text
    a = _run(exe, ["-block16384", "ari", "mtf", "bwt"], data)
    b = _run(exe, ["-block16384", "bwt", "mtf", "ari"], data)
    assert len(a) < 0.6 * len(data) and len(a) < len(b)
    # `echo -n abracadabra | app bwt` (main.rs:6): container + bwt stream of SURVEY Appendix C
    assert _run(exe, ["-block1024", "bwt"], b"abracadabra") == bytes([0x72, 0x21, 0x63, 0x73, 1, 3]) + b"bwt" + bytes.fromhex(
        "00040000" "0b000000" "7264617263616161616262" "02000000")
    assert _run(exe, ["-d"], _run(exe, ["lz4c"], b"")) == b""
    bad = subprocess.run([exe, "-d"], input=b"nope-not-an-archive", stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert bad.returncode != 0 and b"not a rust-compress archive" in bad.stderr


def test_cli_emu(gen):
    build = importlib.import_module("rust-compress_b200.build")
    build.build_emu()
    _check(_build("librcz_emu.so", "rcz_cli_emu"), gen)


@pytest.mark.gpu
def test_cli_gpu(gen):
    build = importlib.import_module("rust-compress_b200.build")
    build.build()
    _check(_build("librcz.so", "rcz_cli_gpu"), gen, scale=40)
