// rcz_ctx.cu — context management for librcz (see include/rcz.h).
#include "rcz_internal.h"

extern "C" int rcz_ctx_create(int device, unsigned flags, rcz_ctx** out) {
    (void)flags;
    if (!out) return RCZ_E_ARG;
    int ndev = 0;
    if (rt_device_count(&ndev) != 0 || ndev <= 0) return RCZ_E_NO_DEVICE;   // no CPU fallback, by design
    if (device < 0 || device >= ndev) return RCZ_E_ARG;
    rcz_ctx* c = new rcz_ctx();
    c->device = device;
    if (rt_set_device(device) != 0) { delete c; return RCZ_E_CUDA; }
    rt_sm_count(device, &c->sm_count);
    if (getenv("RCZ_L2_FETCH")) rt_l2_fetch_granularity(atoi(getenv("RCZ_L2_FETCH")));
    if (rt_stream_create(&c->stream) != 0) { delete c; return RCZ_E_CUDA; }
    c->own_stream = true;
    if (rt_event_create(&c->ev0) != 0 || rt_event_create(&c->ev1) != 0) { delete c; return RCZ_E_CUDA; }
    *out = c;
    return RCZ_OK;
}

extern "C" int rcz_ctx_destroy(rcz_ctx* c) {
    if (!c) return RCZ_OK;
    rt_set_device(c->device);
    rt_stream_sync(c->stream);
    for (auto& w : c->ws) if (w.p) rt_free(w.p);
    if (c->pinned) rt_host_free(c->pinned);
    rt_event_destroy(c->ev0); rt_event_destroy(c->ev1);
    for (auto e : c->stage_ev) if (e) rt_event_destroy(e);
    for (auto e : c->events) rt_event_destroy(e);
    for (int i = 0; i < 9; ++i) if (c->aux[i]) rt_stream_destroy(c->aux[i]);
    if (c->own_stream) rt_stream_destroy(c->stream);
    delete c;
    return RCZ_OK;
}

extern "C" int rcz_ctx_set_stream(rcz_ctx* c, void* s) {
    if (!c) return RCZ_E_ARG;
    rt_stream_sync(c->stream);
    if (c->own_stream) { rt_stream_destroy(c->stream); c->own_stream = false; }
    c->stream = (rt_stream_t)s;
    return RCZ_OK;
}

extern "C" int rcz_ctx_sync(rcz_ctx* c) {
    if (!c) return RCZ_E_ARG;
    RCZ_CK(c, rt_stream_sync(c->stream));
    return RCZ_OK;
}

extern "C" const char* rcz_strerror(int s) {
    switch (s) {
    case RCZ_OK: return "ok";
    case RCZ_E_INVALID_INPUT: return "invalid input";
    case RCZ_E_UNEXPECTED_EOF: return "unexpected end of file";
    case RCZ_E_OVERLONG_RUN: return "Overly long run";
    case RCZ_E_MALFORMED: return "malformed input (the reference implementation panics here)";
    case RCZ_E_OUTPUT_FULL: return "output buffer too small";
    case RCZ_E_ARG: return "bad argument";
    case RCZ_E_CUDA: return "CUDA error";
    case RCZ_E_NO_DEVICE: return "no CUDA device (librcz has no CPU fallback)";
    case RCZ_E_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
    }
}

extern "C" const char* rcz_last_error(rcz_ctx* c) { return c ? c->err : ""; }
extern "C" uint64_t rcz_kernel_launches(rcz_ctx* c) { return c ? c->launches : 0; }

extern "C" int rcz_last_stage_ms(rcz_ctx* c, float* ms, int cap) {
    if (!c || !ms || !c->ev_valid) return 0;
    int n = 0;
    for (int i = 0; i < c->nstage && i < cap; ++i) {
        float t = 0.f;
        if (rt_event_elapsed(&t, c->stage_ev[i], c->stage_ev[i + 1]) != 0) break;
        ms[n++] = t;
    }
    return n;
}

extern "C" float rcz_last_kernel_ms(rcz_ctx* c) {
    if (!c || !c->ev_valid) return -1.f;
    if (rt_stream_sync(c->stream) != 0) return -1.f;
    float ms = 0.f;
    if (rt_event_elapsed(&ms, c->ev0, c->ev1) != 0) return -1.f;
    return ms;
}

extern "C" int rcz_host_alloc(void** p, size_t bytes) { return rt_host_alloc(p, bytes) == 0 ? RCZ_OK : RCZ_E_CUDA; }
extern "C" int rcz_host_free(void* p) { return rt_host_free(p) == 0 ? RCZ_OK : RCZ_E_CUDA; }

extern "C" const char* rcz_build_info(void) {
#ifdef RCZ_EMU
    return "librcz_emu (CPU SIMT emulation, tests only)";
#else
    return "librcz sm_100a (nvcc " __DATE__ ")";
#endif
}

extern "C" int64_t rcz_lz4_compression_bound(uint32_t size) {   // lz4.rs:175-181
    if (size > 0x7e000000u) return -1;
    return (int64_t)size + (size / 255) + 16 + 4;
}
