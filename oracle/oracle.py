"""ctypes binding of the CPU parity oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY — see oracle/oracle.h.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

OK, E_INVALID_INPUT, E_UNEXPECTED_EOF, E_OVERLONG_RUN, E_MALFORMED, E_OUTPUT_FULL, E_ARG = 0, -1, -2, -3, -4, -5, -6


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_lz4_compression_bound.restype = C.c_int64
        _lib.orc_adler32.restype = C.c_uint32
    return _lib


def _u8(b):
    a = np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else np.ascontiguousarray(b, dtype=np.uint8)
    return a, a.ctypes.data_as(C.c_void_p)


def _simple(fn, data, cap):
    a, pa = _u8(data)
    out = np.empty(max(int(cap), 1), dtype=np.uint8)
    n = C.c_size_t(0)
    st = fn(pa, C.c_size_t(a.size), out.ctypes.data_as(C.c_void_p), C.c_size_t(int(cap)), C.byref(n))
    return st, out[: min(n.value, int(cap))].tobytes(), n.value


def rle_encode(data):
    st, out, n = _simple(lib().orc_rle_encode, data, 2 * len(data) + 16)
    assert st == OK
    return out


def rle_decode(data, cap=None):
    cap = cap if cap is not None else 1 << 26
    st, out, n = _simple(lib().orc_rle_decode, data, cap)
    return st, out


def lz4_decode_block(data, cap):
    st, out, n = _simple(lib().orc_lz4_decode_block, data, cap)
    return st, out


def lz4_encode_block(data):
    b = lib().orc_lz4_compression_bound(C.c_uint32(len(data)))
    st, out, n = _simple(lib().orc_lz4_encode_block, data, max(b, 0) + 16)
    assert st == OK
    return out


def lz4_compression_bound(n):
    b = lib().orc_lz4_compression_bound(C.c_uint32(n))
    return None if b < 0 else int(b)


def lz4_frame_decode(data, cap):
    a, pa = _u8(data)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    n, used = C.c_size_t(0), C.c_size_t(0)
    st = lib().orc_lz4_frame_decode(pa, C.c_size_t(a.size), out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n), C.byref(used))
    return st, out[: n.value].tobytes(), used.value


def bwt_suffixes(data):
    a, pa = _u8(data)
    sa = np.empty(max(a.size, 1), dtype=np.uint32)
    st = lib().orc_bwt_suffixes(pa, C.c_size_t(a.size), sa.ctypes.data_as(C.c_void_p))
    assert st == OK
    return sa[: a.size]


def bwt_encode(data):
    a, pa = _u8(data)
    out = np.empty(max(a.size, 1), dtype=np.uint8)
    origin = C.c_uint32(0)
    st = lib().orc_bwt_encode(pa, C.c_size_t(a.size), out.ctypes.data_as(C.c_void_p), C.byref(origin))
    return st, out[: a.size].tobytes(), origin.value


def bwt_inversion_table(l, origin):
    a, pa = _u8(l)
    t = np.empty(max(a.size, 1), dtype=np.uint32)
    st = lib().orc_bwt_inversion_table(pa, C.c_size_t(a.size), C.c_size_t(origin), t.ctypes.data_as(C.c_void_p))
    return st, t[: a.size]


def bwt_decode(l, origin):
    a, pa = _u8(l)
    out = np.empty(max(a.size, 1), dtype=np.uint8)
    n = C.c_size_t(0)
    st = lib().orc_bwt_decode(pa, C.c_size_t(a.size), C.c_size_t(origin), out.ctypes.data_as(C.c_void_p), C.byref(n))
    return st, out[: n.value].tobytes()


def bwt_stream_encode(data, block_size):
    a, pa = _u8(data)
    cap = a.size + 16 + 8 * (a.size // max(block_size, 1) + 2)
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    st = lib().orc_bwt_stream_encode(pa, C.c_size_t(a.size), C.c_uint32(block_size), out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n))
    return st, out[: n.value].tobytes()


def bwt_stream_decode(data, cap=None):
    cap = cap if cap is not None else len(data) + 16
    st, out, n = _simple(lib().orc_bwt_stream_decode, data, cap)
    return st, out


def mtf_encode(data):
    a, pa = _u8(data)
    out = np.empty(max(a.size, 1), dtype=np.uint8)
    lib().orc_mtf_encode(pa, C.c_size_t(a.size), out.ctypes.data_as(C.c_void_p))
    return out[: a.size].tobytes()


def mtf_decode(data):
    a, pa = _u8(data)
    out = np.empty(max(a.size, 1), dtype=np.uint8)
    lib().orc_mtf_decode(pa, C.c_size_t(a.size), out.ctypes.data_as(C.c_void_p))
    return out[: a.size].tobytes()


def dc_encode(data, with_ctx=False):
    """-> (status, init[256] u32, dist[ndist] u32[, (sym, rank, limit)])"""
    a, pa = _u8(data)
    n = a.size
    init = np.empty(256, dtype=np.uint32)
    dist = np.empty(max(n, 1), dtype=np.uint32)
    cs, cr, cl = np.empty(max(n, 1), np.uint8), np.empty(max(n, 1), np.uint8), np.empty(max(n, 1), np.uint32)
    nd = C.c_size_t(0)
    st = lib().orc_dc_encode(pa, C.c_size_t(n), init.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p), C.byref(nd),
                             cs.ctypes.data_as(C.c_void_p), cr.ctypes.data_as(C.c_void_p), cl.ctypes.data_as(C.c_void_p))
    k = nd.value
    if with_ctx:
        return st, init, dist[:k].copy(), (cs[:k].copy(), cr[:k].copy(), cl[:k].copy())
    return st, init, dist[:k].copy()


def dc_decode(n, init, dist, with_ctx=False):
    init = np.ascontiguousarray(init, dtype=np.uint32)
    dist = np.ascontiguousarray(dist, dtype=np.uint32)
    out = np.empty(max(n, 1), dtype=np.uint8)
    m = max(dist.size, 1)
    cs, cr, cl = np.empty(m + 1, np.uint8), np.empty(m + 1, np.uint8), np.empty(m + 1, np.uint32)
    used = C.c_size_t(0)
    st = lib().orc_dc_decode(C.c_size_t(n), init.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p), C.c_size_t(dist.size),
                             out.ctypes.data_as(C.c_void_p), C.byref(used),
                             cs.ctypes.data_as(C.c_void_p), cr.ctypes.data_as(C.c_void_p), cl.ctypes.data_as(C.c_void_p))
    k = used.value
    if with_ctx:
        return st, out[:n].tobytes(), k, (cs[:k].copy(), cr[:k].copy(), cl[:k].copy())
    return st, out[:n].tobytes(), k


def flate_decode(data, cap, blocks=False):
    a, pa = _u8(data)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    n, used, detail, nblk = C.c_size_t(0), C.c_size_t(0), C.c_int(0), C.c_size_t(0)
    sizes = np.zeros(1 << 16, dtype=np.uint32)
    st = lib().orc_flate_decode_blocks(pa, C.c_size_t(a.size), out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n), C.byref(used),
                                       C.byref(detail), sizes.ctypes.data_as(C.c_void_p), C.c_size_t(sizes.size), C.byref(nblk))
    res = (st, out[: n.value].tobytes(), used.value, detail.value)
    if blocks:
        return res + (sizes[: min(nblk.value, sizes.size)].copy(),)
    return res


def zlib_decode(data, cap):
    """zlib::Decoder read_to_end: returns (status, bytes, consumed, detail, adler32 of the output)."""
    a, pa = _u8(data)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    n, used, detail, ad = C.c_size_t(0), C.c_size_t(0), C.c_int(0), C.c_uint32(0)
    st = lib().orc_zlib_decode(pa, C.c_size_t(a.size), out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n), C.byref(used),
                               C.byref(detail), C.byref(ad))
    return st, out[: n.value].tobytes(), used.value, detail.value, ad.value


def ari_encode(data):
    st, out, n = _simple(lib().orc_ari_encode, data, 2 * len(data) + 64)
    assert st == OK
    return out


def ari_decode(data, cap):
    a, pa = _u8(data)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    n, cr, cf = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    st = lib().orc_ari_decode(pa, C.c_size_t(a.size), out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n), C.byref(cr), C.byref(cf))
    return st, out[: n.value].tobytes(), cr.value, cf.value


def adler32(data):
    a, pa = _u8(data)
    return int(lib().orc_adler32(pa, C.c_size_t(a.size)))


def _desc(arr):
    return np.ascontiguousarray(arr, dtype=np.uint64)


def lz4_decode_blocks_mt(in_buf, in_off, in_len, out_buf, out_off, out_cap, nthreads):
    """Batch CPU decode (bench cpu_baseline). in_buf/out_buf: numpy uint8 arrays."""
    nb = len(in_off)
    in_off, in_len, out_off, out_cap = map(_desc, (in_off, in_len, out_off, out_cap))
    out_len = np.zeros(nb, dtype=np.uint64)
    status = np.zeros(nb, dtype=np.int32)
    lib().orc_lz4_decode_blocks_mt(in_buf.ctypes.data_as(C.c_void_p), in_off.ctypes.data_as(C.c_void_p), in_len.ctypes.data_as(C.c_void_p),
                                   out_buf.ctypes.data_as(C.c_void_p), out_off.ctypes.data_as(C.c_void_p), out_cap.ctypes.data_as(C.c_void_p),
                                   out_len.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p), C.c_size_t(nb), C.c_int(nthreads))
    return out_len, status


def bwt_decode_blocks_mt(l_buf, off, n, origin, out_buf, nthreads):
    nb = len(off)
    off, n = map(_desc, (off, n))
    origin = np.ascontiguousarray(origin, dtype=np.uint32)
    out_len = np.zeros(nb, dtype=np.uint64)
    status = np.zeros(nb, dtype=np.int32)
    lib().orc_bwt_decode_blocks_mt(l_buf.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p),
                                   origin.ctypes.data_as(C.c_void_p), out_buf.ctypes.data_as(C.c_void_p),
                                   out_len.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p), C.c_size_t(nb), C.c_int(nthreads))
    return out_len, status


def bwt_encode_blocks_mt(in_buf, off, n, out_buf, nthreads):
    nb = len(off)
    off, n = map(_desc, (off, n))
    origin = np.zeros(nb, dtype=np.uint32)
    status = np.zeros(nb, dtype=np.int32)
    lib().orc_bwt_encode_blocks_mt(in_buf.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p),
                                   out_buf.ctypes.data_as(C.c_void_p), origin.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p),
                                   C.c_size_t(nb), C.c_int(nthreads))
    return origin, status


def flate_decode_streams_mt(in_buf, in_off, in_len, out_buf, out_off, out_cap, nthreads):
    nb = len(in_off)
    in_off, in_len, out_off, out_cap = map(_desc, (in_off, in_len, out_off, out_cap))
    out_len = np.zeros(nb, dtype=np.uint64)
    status = np.zeros(nb, dtype=np.int32)
    lib().orc_flate_decode_streams_mt(in_buf.ctypes.data_as(C.c_void_p), in_off.ctypes.data_as(C.c_void_p), in_len.ctypes.data_as(C.c_void_p),
                                      out_buf.ctypes.data_as(C.c_void_p), out_off.ctypes.data_as(C.c_void_p), out_cap.ctypes.data_as(C.c_void_p),
                                      out_len.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p), C.c_size_t(nb), C.c_int(nthreads))
    return out_len, status


def bda_encode_blocks_mt(in_buf, in_off, n, chunk, out_buf, out_off, out_cap, nthreads):
    """bwt -> dc -> ari composition (oracle/pipeline.cpp). Returns (out_len, origin, status)."""
    nb = len(in_off)
    in_off, n, out_off, out_cap = map(_desc, (in_off, n, out_off, out_cap))
    out_len = np.zeros(nb, dtype=np.uint64)
    origin = np.zeros(nb, dtype=np.uint32)
    status = np.zeros(nb, dtype=np.int32)
    lib().orc_bda_encode_blocks_mt(in_buf.ctypes.data_as(C.c_void_p), in_off.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p), C.c_uint32(chunk),
                                   out_buf.ctypes.data_as(C.c_void_p), out_off.ctypes.data_as(C.c_void_p), out_cap.ctypes.data_as(C.c_void_p),
                                   out_len.ctypes.data_as(C.c_void_p), origin.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p),
                                   C.c_size_t(nb), C.c_int(nthreads))
    return out_len, origin, status


def bda_decode_blocks_mt(in_buf, in_off, in_len, chunk, out_buf, out_off, n, nthreads):
    nb = len(in_off)
    in_off, in_len, out_off, n = map(_desc, (in_off, in_len, out_off, n))
    out_len = np.zeros(nb, dtype=np.uint64)
    status = np.zeros(nb, dtype=np.int32)
    lib().orc_bda_decode_blocks_mt(in_buf.ctypes.data_as(C.c_void_p), in_off.ctypes.data_as(C.c_void_p), in_len.ctypes.data_as(C.c_void_p), C.c_uint32(chunk),
                                   out_buf.ctypes.data_as(C.c_void_p), out_off.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p),
                                   out_len.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p), C.c_size_t(nb), C.c_int(nthreads))
    return out_len, status
