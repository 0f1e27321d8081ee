// stubs.cu — entry points declared in include/rcz.h whose kernels are not built yet.
// Each returns RCZ_E_UNSUPPORTED (never a CPU fallback).  Removed one by one as the kernels land.
#include "rcz_internal.h"
#define RCZ_STUB_BATCH(name)                                                                                         \
    extern "C" int name(rcz_ctx*, const void*, const uint64_t*, const uint64_t*, void*, const uint64_t*,             \
                        const uint64_t*, uint64_t*, int32_t*, size_t, int) { return RCZ_E_UNSUPPORTED; }
#if !__has_include("bwt_encode.cu")
extern "C" int rcz_bwt_encode_blocks(rcz_ctx*, const void*, const uint64_t*, const uint64_t*, void*, const uint64_t*, uint32_t*,
                                     int32_t*, size_t, int) { return RCZ_E_UNSUPPORTED; }
#endif
#if !__has_include("flate_decode.cu")
extern "C" int rcz_flate_decode_streams(rcz_ctx*, const void*, const uint64_t*, const uint64_t*, void*, const uint64_t*,
                                        const uint64_t*, uint64_t*, uint64_t*, int32_t*, int32_t*, size_t, int) { return RCZ_E_UNSUPPORTED; }
#endif
#if !__has_include("ari.cu")
RCZ_STUB_BATCH(rcz_ari_encode_streams)
extern "C" int rcz_ari_decode_streams(rcz_ctx*, const void*, const uint64_t*, const uint64_t*, void*, const uint64_t*,
                                      const uint64_t*, uint64_t*, uint64_t*, int32_t*, size_t, int) { return RCZ_E_UNSUPPORTED; }
#endif
#if !__has_include("dc.cu")
extern "C" int rcz_dc_encode_blocks(rcz_ctx*, const void*, const uint64_t*, const uint64_t*, uint32_t*, const uint64_t*,
                                    const uint64_t*, uint64_t*, int32_t*, size_t, int) { return RCZ_E_UNSUPPORTED; }
extern "C" int rcz_dc_decode_blocks(rcz_ctx*, const uint32_t*, const uint64_t*, const uint64_t*, void*, const uint64_t*,
                                    const uint64_t*, int32_t*, size_t, int) { return RCZ_E_UNSUPPORTED; }
#endif
#if !__has_include("rle.cu")
RCZ_STUB_BATCH(rcz_rle_decode_streams)
RCZ_STUB_BATCH(rcz_rle_encode_streams)
#endif
