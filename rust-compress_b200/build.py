"""Builds librcz.so (nvcc, sm_100a — the product) and, for tests only, librcz_emu.so (g++, CPU SIMT emulation).

    python rust-compress_b200/build.py            # product library
    python rust-compress_b200/build.py --emu      # test-only emulation library
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["rcz_ctx.cu", "lz4_decode.cu", "lz4_encode.cu", "bwt_decode.cu", "bwt_encode.cu", "flate_decode.cu", "ari.cu", "dc.cu", "rle.cu", "mtf.cu", "pipeline.cu"]
LIB = os.path.join(HERE, "librcz.so")
LIB_EMU = os.path.join(HERE, "librcz_emu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--cudart", "static",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _stale(target, extra=()):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "rcz.h"), __file__] + list(extra)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale(LIB):
        return LIB
    cmd = [NVCC] + NVCC_FLAGS + ["-o", LIB] + _sources()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build_ptxas.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed (see %s)" % log)
    return LIB


def build_emu(force=False):
    if not force and not _stale(LIB_EMU):
        return LIB_EMU
    objs = []
    bdir = os.path.join(HERE, "build_emu")
    os.makedirs(bdir, exist_ok=True)
    flags = ["-O2", "-g", "-std=c++17", "-fPIC", "-DRCZ_EMU", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function",
             "-fno-omit-frame-pointer", "-fno-strict-aliasing"]
    procs = []
    for s in _sources() + [os.path.join(CSRC, "emu_rt.cpp")]:
        o = os.path.join(bdir, os.path.basename(s) + ".o")
        objs.append(o)
        procs.append((s, subprocess.Popen(["g++", "-x", "c++"] + flags + ["-c", s, "-o", o])))
    for s, p in procs:
        if p.wait() != 0:
            raise RuntimeError("emu compile failed: " + s)
    subprocess.check_call(["g++", "-shared", "-o", LIB_EMU] + objs)
    return LIB_EMU


def build_cli(force=False):
    """rcz_cli: the reference's test application (main.rs) over the host mirrors, linked against the product library."""
    src = os.path.join(HERE, "host", "rcz_cli.cpp")
    out = os.path.join(HERE, "rcz_cli")
    deps = [src, os.path.join(HERE, "host", "rcz_stream.hpp"), os.path.join(HERE, "..", "include", "rcz.h"), LIB]
    if force or not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(HERE, "..", "include"), "-I", os.path.join(HERE, "host"), src, "-o", out,
                               "-L", HERE, "-l:librcz.so", "-Wl,-rpath," + HERE])
    return out


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu(force="--force" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
