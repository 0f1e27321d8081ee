"""rust-compress_b200 — B200 (sm_100a) kernels for the inner loops of the Rust `compress` crate
(rusty-shell/rust-compress), behind the crate's Reader/Writer surface.

Import with ``importlib.import_module("rust-compress_b200")`` (the directory name is not a Python identifier).

Layers
  include/rcz.h                 the C ABI (extern "C", plain pointers) exported by librcz.so
  rust-compress_b200/csrc       hand-written CUDA kernels + the ABI implementation
  rust-compress_b200/_abi.py    ctypes binding
  this module                   `Context`: batched per-block calls on numpy (host) or torch (device) buffers
  rust-compress_b200/host/rcz_stream.hpp   C++ mirrors of the crate's Decoder<R>/Encoder<W> types (the Rust shim is rust/rcz_sys.rs)
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import (E_ARG, E_CUDA, E_INVALID_INPUT, E_MALFORMED, E_NO_DEVICE, E_OUTPUT_FULL, E_OVERLONG_RUN,  # noqa: F401
                   E_UNEXPECTED_EOF, E_UNSUPPORTED, MEM_DEVICE, MEM_DEVICE_ASYNC, MEM_HOST, OK, RczError)

try:  # torch is plumbing (device memory, streams); the ABI itself does not need it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    if _is_torch(x):
        assert x.is_contiguous()
        return x.data_ptr()
    if isinstance(x, int):
        return x
    raise TypeError(type(x))


def _kind_of(data, async_):
    """HOST for numpy / CPU tensors; DEVICE (or DEVICE_ASYNC when async_) for CUDA tensors."""
    if isinstance(data, np.ndarray):
        return MEM_DEVICE if async_ == "emu-device" else MEM_HOST
    if _is_torch(data):
        if not data.is_cuda:
            return MEM_HOST
        return MEM_DEVICE_ASYNC if async_ else MEM_DEVICE
    raise TypeError("buffers must be numpy arrays (host) or torch tensors")


def _u64(x):
    if _is_torch(x):
        x = x.cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.uint64)


class Context:
    """One rcz_ctx: one device, one stream.  `emu=True` selects the CPU emulation build (tests only)."""

    def __init__(self, device=0, stream=None, emu=False):
        self._lib = _abi.load(emu=emu)
        self.emu = emu
        h = C.c_void_p()
        st = self._lib.rcz_ctx_create(int(device), 0, C.byref(h))
        if st == E_NO_DEVICE:
            raise RczError(st, "no CUDA device visible; librcz has no CPU fallback")
        if st != OK:
            raise RczError(st, "rcz_ctx_create")
        self._h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rcz_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        s = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
        self._check(self._lib.rcz_ctx_set_stream(self._h, s), "set_stream")

    def sync(self):
        self._check(self._lib.rcz_ctx_sync(self._h), "sync")

    @property
    def launches(self):
        return int(self._lib.rcz_kernel_launches(self._h))

    def last_kernel_ms(self):
        return float(self._lib.rcz_last_kernel_ms(self._h))

    def last_stage_ms(self):
        """Per-kernel durations of the most recent multi-kernel batch call (lz4: parse, scan, materialise)."""
        import ctypes
        buf = (ctypes.c_float * 8)()
        n = self._lib.rcz_last_stage_ms(self._h, buf, 8)
        return [float(buf[i]) for i in range(n)]

    def _check(self, st, what):
        if st != OK:
            raise RczError(st, "%s: %s %s" % (what, self._lib.rcz_strerror(st).decode(), self._lib.rcz_last_error(self._h).decode()))

    # ---- result arrays -------------------------------------------------------------------------------------
    def _results(self, kind, n, like, specs):
        outs = []
        for dt in specs:
            if kind == MEM_DEVICE_ASYNC:
                tdt = {np.uint64: torch.int64, np.int32: torch.int32, np.uint32: torch.int32}[dt]
                outs.append(torch.zeros(n, dtype=tdt, device=like.device))
            else:
                outs.append(np.zeros(n, dtype=dt))
        return outs

    def _batch(self, fn, name, in_buf, in_off, in_len, out_buf, out_off, out_cap, extra_out=(), async_=False):
        kind = _kind_of(in_buf, async_)
        n = len(in_off)
        in_off, in_len, out_off, out_cap = map(_u64, (in_off, in_len, out_off, out_cap))
        out_len, status = self._results(kind, n, in_buf, [np.uint64, np.int32])
        extras = self._results(kind, n, in_buf, list(extra_out))
        return kind, n, (in_off, in_len, out_off, out_cap), out_len, status, extras

    # ---- lz4 -----------------------------------------------------------------------------------------------
    def lz4_decode_blocks(self, in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=False):
        """rcz_lz4_decode_blocks.  Returns (out_len, status) arrays."""
        kind, n, (io, il, oo, oc), out_len, status, _ = self._batch(None, "lz4", in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=async_)
        st = self._lib.rcz_lz4_decode_blocks(self._h, _ptr(in_buf), _ptr(io), _ptr(il), _ptr(out_buf), _ptr(oo), _ptr(oc),
                                             _ptr(out_len), _ptr(status), n, kind)
        self._check(st, "rcz_lz4_decode_blocks")
        return out_len, status

    def lz4_decode_blocks_gather(self, in_buf, in_off, in_len, out_buf, out_off, out_cap, peer_ptrs, async_=True):
        """rcz_lz4_decode_blocks_gather: decode + store every output chunk into the peers' buffers as well (device pointers as ints)."""
        kind, n, (io, il, oo, oc), out_len, status, _ = self._batch(None, "lz4", in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=async_)
        arr = (C.c_void_p * max(1, len(peer_ptrs)))(*[int(p) for p in peer_ptrs])
        st = self._lib.rcz_lz4_decode_blocks_gather(self._h, _ptr(in_buf), _ptr(io), _ptr(il), _ptr(out_buf), _ptr(oo), _ptr(oc),
                                                    _ptr(out_len), _ptr(status), n, kind, C.cast(arr, C.c_void_p), len(peer_ptrs))
        self._check(st, "rcz_lz4_decode_blocks_gather")
        return out_len, status

    def lz4_encode_blocks(self, in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=False):
        """rcz_lz4_encode_blocks (lz4.rs:226-310, bit-exact).  Returns (out_len, status) arrays."""
        kind, n, (io, il, oo, oc), out_len, status, _ = self._batch(None, "lz4e", in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=async_)
        st = self._lib.rcz_lz4_encode_blocks(self._h, _ptr(in_buf), _ptr(io), _ptr(il), _ptr(out_buf), _ptr(oo), _ptr(oc),
                                             _ptr(out_len), _ptr(status), n, kind)
        self._check(st, "rcz_lz4_encode_blocks")
        return out_len, status

    # ---- bwt -----------------------------------------------------------------------------------------------
    def bwt_decode_blocks(self, in_buf, in_off, n_arr, origin, out_buf, out_off, async_=False):
        kind = _kind_of(in_buf, async_)
        nb = len(in_off)
        io, na, oo = map(_u64, (in_off, n_arr, out_off))
        org = np.ascontiguousarray(origin.cpu().numpy() if _is_torch(origin) else origin, dtype=np.uint32)
        out_len, status = self._results(kind, nb, in_buf, [np.uint64, np.int32])
        st = self._lib.rcz_bwt_decode_blocks(self._h, _ptr(in_buf), _ptr(io), _ptr(na), _ptr(org), _ptr(out_buf), _ptr(oo),
                                             _ptr(out_len), _ptr(status), nb, kind)
        self._check(st, "rcz_bwt_decode_blocks")
        return out_len, status

    def bwt_encode_blocks(self, in_buf, in_off, n_arr, out_buf, out_off, async_=False):
        kind = _kind_of(in_buf, async_)
        nb = len(in_off)
        io, na, oo = map(_u64, (in_off, n_arr, out_off))
        origin, status = self._results(kind, nb, in_buf, [np.uint32, np.int32])
        st = self._lib.rcz_bwt_encode_blocks(self._h, _ptr(in_buf), _ptr(io), _ptr(na), _ptr(out_buf), _ptr(oo),
                                             _ptr(origin), _ptr(status), nb, kind)
        self._check(st, "rcz_bwt_encode_blocks")
        return origin, status

    # ---- flate ---------------------------------------------------------------------------------------------
    def flate_decode_streams(self, in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=False):
        kind, n, (io, il, oo, oc), out_len, status, (in_used, detail) = self._batch(
            None, "flate", in_buf, in_off, in_len, out_buf, out_off, out_cap, extra_out=(np.uint64, np.int32), async_=async_)
        st = self._lib.rcz_flate_decode_streams(self._h, _ptr(in_buf), _ptr(io), _ptr(il), _ptr(out_buf), _ptr(oo), _ptr(oc),
                                                _ptr(out_len), _ptr(in_used), _ptr(status), _ptr(detail), n, kind)
        self._check(st, "rcz_flate_decode_streams")
        return out_len, status, in_used, detail

    # ---- zlib (zlib.rs + checksum/adler.rs) -------------------------------------------------------------------
    def zlib_decode_streams(self, in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=False):
        """rcz_zlib_decode_streams.  Returns (out_len, status, in_used, detail, adler) arrays."""
        kind, n, (io, il, oo, oc), out_len, status, (in_used, detail, adler) = self._batch(
            None, "zlib", in_buf, in_off, in_len, out_buf, out_off, out_cap, extra_out=(np.uint64, np.int32, np.uint32), async_=async_)
        st = self._lib.rcz_zlib_decode_streams(self._h, _ptr(in_buf), _ptr(io), _ptr(il), _ptr(out_buf), _ptr(oo), _ptr(oc),
                                               _ptr(out_len), _ptr(in_used), _ptr(status), _ptr(detail), _ptr(adler), n, kind)
        self._check(st, "rcz_zlib_decode_streams")
        return out_len, status, in_used, detail, adler

    def adler32_streams(self, buf, off, length, async_=False):
        """rcz_adler32_streams: Adler-32 of every byte range."""
        kind = _kind_of(buf, async_)
        n = len(off)
        o, l = map(_u64, (off, length))
        (adler,) = self._results(kind, n, buf, [np.uint32])
        st = self._lib.rcz_adler32_streams(self._h, _ptr(buf), _ptr(o), _ptr(l), _ptr(adler), n, kind)
        self._check(st, "rcz_adler32_streams")
        return adler

    # ---- mtf (bwt/mtf.rs stream coder) -----------------------------------------------------------------------
    def mtf_encode_streams(self, in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=False):
        kind, n, (io, il, oo, oc), out_len, status, _ = self._batch(None, "mtf", in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=async_)
        st = self._lib.rcz_mtf_encode_streams(self._h, _ptr(in_buf), _ptr(io), _ptr(il), _ptr(out_buf), _ptr(oo), _ptr(oc),
                                              _ptr(out_len), _ptr(status), n, kind)
        self._check(st, "rcz_mtf_encode_streams")
        return out_len, status

    def mtf_decode_streams(self, in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=False):
        kind, n, (io, il, oo, oc), out_len, status, _ = self._batch(None, "mtf", in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=async_)
        st = self._lib.rcz_mtf_decode_streams(self._h, _ptr(in_buf), _ptr(io), _ptr(il), _ptr(out_buf), _ptr(oo), _ptr(oc),
                                              _ptr(out_len), _ptr(status), n, kind)
        self._check(st, "rcz_mtf_decode_streams")
        return out_len, status

    # ---- ari -----------------------------------------------------------------------------------------------
    def ari_encode_streams(self, in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=False):
        kind, n, (io, il, oo, oc), out_len, status, _ = self._batch(None, "ari", in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=async_)
        st = self._lib.rcz_ari_encode_streams(self._h, _ptr(in_buf), _ptr(io), _ptr(il), _ptr(out_buf), _ptr(oo), _ptr(oc),
                                              _ptr(out_len), _ptr(status), n, kind)
        self._check(st, "rcz_ari_encode_streams")
        return out_len, status

    def ari_decode_streams(self, in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=False):
        kind, n, (io, il, oo, oc), out_len, status, (in_used,) = self._batch(
            None, "ari", in_buf, in_off, in_len, out_buf, out_off, out_cap, extra_out=(np.uint64,), async_=async_)
        st = self._lib.rcz_ari_decode_streams(self._h, _ptr(in_buf), _ptr(io), _ptr(il), _ptr(out_buf), _ptr(oo), _ptr(oc),
                                              _ptr(out_len), _ptr(in_used), _ptr(status), n, kind)
        self._check(st, "rcz_ari_decode_streams")
        return out_len, status, in_used

    # ---- dc ------------------------------------------------------------------------------------------------
    def dc_encode_blocks(self, in_buf, in_off, n_arr, out_buf_u32, out_off, out_cap, async_=False):
        kind, n, (io, il, oo, oc), out_len, status, _ = self._batch(None, "dc", in_buf, in_off, n_arr, out_buf_u32, out_off, out_cap, async_=async_)
        st = self._lib.rcz_dc_encode_blocks(self._h, _ptr(in_buf), _ptr(io), _ptr(il), _ptr(out_buf_u32), _ptr(oo), _ptr(oc),
                                            _ptr(out_len), _ptr(status), n, kind)
        self._check(st, "rcz_dc_encode_blocks")
        return out_len, status

    def dc_decode_blocks(self, in_buf_u32, in_off, in_len, out_buf, out_off, n_arr, async_=False):
        kind = _kind_of(in_buf_u32, async_)
        nb = len(in_off)
        io, il, oo, na = map(_u64, (in_off, in_len, out_off, n_arr))
        (status,) = self._results(kind, nb, in_buf_u32, [np.int32])
        st = self._lib.rcz_dc_decode_blocks(self._h, _ptr(in_buf_u32), _ptr(io), _ptr(il), _ptr(out_buf), _ptr(oo), _ptr(na),
                                            _ptr(status), nb, kind)
        self._check(st, "rcz_dc_decode_blocks")
        return status

    # ---- bwt -> dc -> ari pipeline (BASELINE configs[4]) ------------------------------------------------------
    def bwt_dc_ari_encode_blocks(self, in_buf, in_off, n_arr, out_buf, out_off, out_cap, ari_chunk=65536, async_=False):
        """rcz_bwt_dc_ari_encode_blocks.  Returns (out_len, origin, status) arrays."""
        kind = _kind_of(in_buf, async_)
        nb = len(in_off)
        io, na, oo, oc = map(_u64, (in_off, n_arr, out_off, out_cap))
        out_len, origin, status = self._results(kind, nb, in_buf, [np.uint64, np.uint32, np.int32])
        st = self._lib.rcz_bwt_dc_ari_encode_blocks(self._h, _ptr(in_buf), _ptr(io), _ptr(na), _ptr(out_buf), _ptr(oo), _ptr(oc),
                                                    _ptr(out_len), _ptr(origin), _ptr(status), nb, int(ari_chunk), kind)
        self._check(st, "rcz_bwt_dc_ari_encode_blocks")
        return out_len, origin, status

    def bwt_dc_ari_decode_blocks(self, in_buf, in_off, in_len, out_buf, out_off, n_arr, ari_chunk=65536, async_=False):
        """rcz_bwt_dc_ari_decode_blocks.  Returns (out_len, status) arrays."""
        kind = _kind_of(in_buf, async_)
        nb = len(in_off)
        io, il, oo, na = map(_u64, (in_off, in_len, out_off, n_arr))
        out_len, status = self._results(kind, nb, in_buf, [np.uint64, np.int32])
        st = self._lib.rcz_bwt_dc_ari_decode_blocks(self._h, _ptr(in_buf), _ptr(io), _ptr(il), _ptr(out_buf), _ptr(oo), _ptr(na),
                                                    _ptr(out_len), _ptr(status), nb, int(ari_chunk), kind)
        self._check(st, "rcz_bwt_dc_ari_decode_blocks")
        return out_len, status

    # ---- rle -----------------------------------------------------------------------------------------------
    def rle_decode_streams(self, in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=False):
        kind, n, (io, il, oo, oc), out_len, status, _ = self._batch(None, "rle", in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=async_)
        st = self._lib.rcz_rle_decode_streams(self._h, _ptr(in_buf), _ptr(io), _ptr(il), _ptr(out_buf), _ptr(oo), _ptr(oc),
                                              _ptr(out_len), _ptr(status), n, kind)
        self._check(st, "rcz_rle_decode_streams")
        return out_len, status

    def rle_encode_streams(self, in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=False):
        kind, n, (io, il, oo, oc), out_len, status, _ = self._batch(None, "rle", in_buf, in_off, in_len, out_buf, out_off, out_cap, async_=async_)
        st = self._lib.rcz_rle_encode_streams(self._h, _ptr(in_buf), _ptr(io), _ptr(il), _ptr(out_buf), _ptr(oo), _ptr(oc),
                                              _ptr(out_len), _ptr(status), n, kind)
        self._check(st, "rcz_rle_encode_streams")
        return out_len, status


_default = {}


def default_context(emu=False, device=0):
    key = (bool(emu), device)
    if key not in _default:
        _default[key] = Context(device=device, emu=emu)
    return _default[key]
