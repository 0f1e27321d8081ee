"""ctypes binding of the C ABI declared in include/rcz.h.

The product path loads `librcz.so` (nvcc build for sm_100a) and fails loudly if it is missing or no CUDA
device is present: there is no CPU fallback.  `load(emu=True)` loads `librcz_emu.so`, the same sources
compiled against the CPU SIMT emulation; it exists for the CPU-only test suite and is never selected
implicitly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librcz.so")
LIB_EMU_PATH = os.path.join(HERE, "librcz_emu.so")

OK, E_INVALID_INPUT, E_UNEXPECTED_EOF, E_OVERLONG_RUN, E_MALFORMED, E_OUTPUT_FULL, E_ARG, E_CUDA, E_NO_DEVICE, E_UNSUPPORTED = \
    0, -1, -2, -3, -4, -5, -6, -7, -8, -9
MEM_HOST, MEM_DEVICE, MEM_DEVICE_ASYNC = 0, 1, 2

_P, _SZ, _I = C.c_void_p, C.c_size_t, C.c_int
# name -> (restype, argtypes); must list every function declared in include/rcz.h
_BATCH = [_P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _I]   # ctx,in,in_off,in_len,out,out_off,out_cap,out_len,status,n,kind
SIGNATURES = {
    "rcz_ctx_create": (_I, [_I, C.c_uint, C.POINTER(_P)]),
    "rcz_ctx_destroy": (_I, [_P]),
    "rcz_ctx_set_stream": (_I, [_P, _P]),
    "rcz_ctx_sync": (_I, [_P]),
    "rcz_strerror": (C.c_char_p, [_I]),
    "rcz_last_error": (C.c_char_p, [_P]),
    "rcz_kernel_launches": (C.c_uint64, [_P]),
    "rcz_last_kernel_ms": (C.c_float, [_P]),
    "rcz_last_stage_ms": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int]),
    "rcz_host_alloc": (_I, [C.POINTER(_P), _SZ]),
    "rcz_host_free": (_I, [_P]),
    "rcz_build_info": (C.c_char_p, []),
    "rcz_lz4_decode_blocks": (_I, _BATCH),
    "rcz_lz4_decode_blocks_gather": (_I, _BATCH + [_P, _I]),
    "rcz_lz4_encode_blocks": (_I, _BATCH),
    "rcz_lz4_compression_bound": (C.c_int64, [C.c_uint32]),
    "rcz_bwt_decode_blocks": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _I]),
    "rcz_bwt_encode_blocks": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _SZ, _I]),
    "rcz_flate_decode_streams": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _I]),
    "rcz_zlib_decode_streams": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _I]),
    "rcz_adler32_streams": (_I, [_P, _P, _P, _P, _P, _SZ, _I]),
    "rcz_ari_encode_streams": (_I, _BATCH),
    "rcz_ari_decode_streams": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _I]),
    "rcz_dc_encode_blocks": (_I, _BATCH),
    "rcz_dc_decode_blocks": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _SZ, _I]),
    "rcz_bwt_dc_ari_encode_blocks": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, C.c_uint32, _I]),
    "rcz_bwt_dc_ari_decode_blocks": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, C.c_uint32, _I]),
    "rcz_rle_decode_streams": (_I, _BATCH),
    "rcz_rle_encode_streams": (_I, _BATCH),
    "rcz_mtf_encode_streams": (_I, _BATCH),
    "rcz_mtf_decode_streams": (_I, _BATCH),
}

_libs = {}


class RczError(RuntimeError):
    def __init__(self, status, msg=""):
        self.status = status
        super().__init__("librcz status %d%s" % (status, (": " + msg) if msg else ""))


def load(emu=False):
    key = bool(emu)
    if key in _libs:
        return _libs[key]
    path = LIB_EMU_PATH if emu else LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            "%s not found: build it with `python rust-compress_b200/build.py%s` (or __graft_entry__.build()). "
            "librcz has no CPU fallback." % (path, " --emu" if emu else ""))
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here == missing export
        fn.restype, fn.argtypes = res, args
    _libs[key] = lib
    return lib
