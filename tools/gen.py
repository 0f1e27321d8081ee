"""Deterministic synthetic inputs for tests and bench.py.  Not part of the product."""
import ctypes as C
import ctypes.util
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "librczgen.so")
MASK = (1 << 64) - 1
SEED_BASE = 0xC0DEC0DE00000000


def build(force=False):
    src = os.path.join(HERE, "rczgen.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", _SO, src])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def unit_seed(config_id, unit_index):
    return (SEED_BASE ^ (config_id << 32) ^ unit_index) & MASK


def splitmix64(state):
    """One step; returns (new_state, output).  Python twin of rczgen.cpp::sm64."""
    state = (state + 0x9E3779B97F4A7C15) & MASK
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
    return state, z ^ (z >> 31)


KINDS = {"random": 0, "lzsyn": 1, "hextext": 2, "runs": 3}


def units(kind, seed0, unit, count, nthreads=None, out=None):
    """`count` units of `unit` bytes, unit i seeded with seed0 ^ i; returns a numpy uint8 array."""
    nthreads = nthreads or os.cpu_count() or 1
    if out is None:
        out = np.empty(unit * count + 64, dtype=np.uint8)[: unit * count]
    lib().gen_units(C.c_int(KINDS[kind]), C.c_uint64(seed0 & MASK), out.ctypes.data_as(C.c_void_p), C.c_size_t(unit), C.c_size_t(count),
                    C.c_int(nthreads))
    return out


def one(kind, seed, n):
    return units(kind, seed, n, 1, nthreads=1).tobytes()


_lz4 = None


def liblz4():
    global _lz4
    if _lz4 is None:
        name = ctypes.util.find_library("lz4") or "liblz4.so.1"
        _lz4 = C.CDLL(name)
        _lz4.LZ4_compress_default.restype = C.c_int
        _lz4.LZ4_decompress_safe.restype = C.c_int
        _lz4.LZ4_compress_HC.restype = C.c_int
    return _lz4


def lz4_compress(data, hc=False):
    data = bytes(data)
    cap = len(data) + len(data) // 255 + 32
    dst = C.create_string_buffer(cap)
    if hc:
        n = liblz4().LZ4_compress_HC(data, dst, len(data), cap, 9)
    else:
        n = liblz4().LZ4_compress_default(data, dst, len(data), cap)
    assert n > 0
    return dst.raw[:n]


def lz4_decompress(data, cap):
    dst = C.create_string_buffer(max(cap, 1))
    n = liblz4().LZ4_decompress_safe(bytes(data), dst, len(data), cap)
    return n, dst.raw[: max(n, 0)]


def lz4_compress_units(raw, unit, count, nthreads=None):
    """liblz4-compress `count` units in parallel -> (packed uint8 array, in_off u64, in_len u64); blocks 16-byte aligned."""
    nthreads = nthreads or os.cpu_count() or 1
    stride = unit + unit // 255 + 32
    tmp = np.empty(stride * count, dtype=np.uint8)
    lens = np.zeros(count, dtype=np.uint64)
    fn = C.cast(liblz4().LZ4_compress_default, C.c_void_p)
    lib().lz4_compress_units(fn, raw.ctypes.data_as(C.c_void_p), C.c_size_t(unit), C.c_size_t(count), tmp.ctypes.data_as(C.c_void_p),
                             C.c_size_t(stride), lens.ctypes.data_as(C.c_void_p), C.c_int(nthreads))
    assert (lens > 0).all()
    al = (lens + np.uint64(15)) & ~np.uint64(15)
    off = np.zeros(count, dtype=np.uint64)
    off[1:] = np.cumsum(al)[:-1]
    packed = np.zeros(int(al.sum()) + 64, dtype=np.uint8)
    for i in range(count):
        packed[int(off[i]): int(off[i]) + int(lens[i])] = tmp[i * stride: i * stride + int(lens[i])]
    return packed, off, lens
