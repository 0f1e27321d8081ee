// rcz_internal.h — context, workspace and batch staging shared by all librcz ops.
#pragma once
#include "../../include/rcz.h"
#include "rt.h"
#include <stdio.h>
#include <vector>
#include <algorithm>
#include <stdlib.h>

// WS_P*: intermediates owned by a composed call (pipeline.cu) that must survive the nested stage calls
enum { WS_IN = 0, WS_OUT = 1, WS_DESC = 2, WS_A = 3, WS_B = 4, WS_C = 5, WS_D = 6, WS_E = 7, WS_P0 = 8, WS_P1 = 9, WS_P2 = 10, WS_P3 = 11, WS_P4 = 12, WS_COUNT = 13 };
constexpr int RCZ_MAX_STAGES = 8;

struct rcz_ctx {
    int device = 0;
    int sm_count = 148;
    rt_stream_t stream = 0;
    bool own_stream = false;
    uint64_t launches = 0;
    rt_event_t ev0 = 0, ev1 = 0;
    bool ev_valid = false;
    rt_event_t stage_ev[RCZ_MAX_STAGES + 1] = {};   // stage boundaries of the most recent multi-kernel call
    int nstage = 0;
    int nest = 0;                     // > 0 while a composed call (pipeline.cu) drives the stage entry points: they leave the timers alone
    bool nest_may_sync = false;       // the composed call itself is not RCZ_MEM_DEVICE_ASYNC: a stage may wait for the stream (forward BWT: stop after the last live round)
    char err[256] = {0};
    struct { void* p; size_t cap; } ws[WS_COUNT] = {};
    void* pinned = nullptr;
    size_t pinned_cap = 0;
    rt_stream_t aux[9] = {};          // D2H stream + 8 kernel streams of the pipelined host-buffer path (created on first use)
    std::vector<rt_event_t> events;   // event pool of the pipelined path
    // lz4: window plan of the most recent call; a caller that decodes batches of the same geometry (same compressed lengths, same
    // chunking) gets it back without the host recomputing ~70 K tickets
    struct { std::vector<uint64_t> in_len; std::vector<size_t> cut, tk_off; std::vector<uint32_t> nw, wbase, tickets; size_t totwin = 0; } lz4_plan;
};

#define RCZ_CK(ctx, expr)                                                                              \
    do {                                                                                               \
        int e__ = (expr);                                                                              \
        if (e__ != 0) {                                                                                \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,     \
                     rt_error_string(e__));                                                            \
            return RCZ_E_CUDA;                                                                         \
        }                                                                                              \
    } while (0)

#define RCZ_KLAUNCH(ctx, kern, grid, block, smem, ...)                   \
    do {                                                                 \
        RCZ_LAUNCH(kern, grid, block, smem, (ctx)->stream, __VA_ARGS__); \
        (ctx)->launches++;                                               \
        RCZ_CK(ctx, rt_last_error());                                    \
    } while (0)

inline int ctx_aux_streams(rcz_ctx* c) {
    for (int i = 0; i < 9; ++i) if (!c->aux[i]) RCZ_CK(c, rt_stream_create(&c->aux[i]));
    return RCZ_OK;
}
inline int ctx_events(rcz_ctx* c, size_t n) {
    while (c->events.size() < n) { rt_event_t e; RCZ_CK(c, rt_event_create(&e)); c->events.push_back(e); }
    return RCZ_OK;
}

// grow-only device workspace slot
inline int ctx_ws(rcz_ctx* c, int slot, size_t bytes, void** out) {
    if (c->ws[slot].cap < bytes) {
        if (c->ws[slot].p) { RCZ_CK(c, rt_stream_sync(c->stream)); RCZ_CK(c, rt_free(c->ws[slot].p)); c->ws[slot].p = nullptr; c->ws[slot].cap = 0; }
        size_t cap = (bytes + (bytes >> 3) + 4095) & ~(size_t)4095;
        RCZ_CK(c, rt_malloc(&c->ws[slot].p, cap));
        c->ws[slot].cap = cap;
    }
    *out = c->ws[slot].p;
    return RCZ_OK;
}
inline int ctx_pinned(rcz_ctx* c, size_t bytes, void** out) {
    if (c->pinned_cap < bytes) {
        if (c->pinned) { RCZ_CK(c, rt_stream_sync(c->stream)); RCZ_CK(c, rt_host_free(c->pinned)); c->pinned = nullptr; c->pinned_cap = 0; }
        size_t cap = (bytes * 2 + 4095) & ~(size_t)4095;
        RCZ_CK(c, rt_host_alloc(&c->pinned, cap));
        c->pinned_cap = cap;
    }
    *out = c->pinned;
    return RCZ_OK;
}
inline int ctx_stage_mark(rcz_ctx* c, int i) {      // boundary i of the call's kernel sequence (0 = before the first kernel)
    if (i > RCZ_MAX_STAGES || c->nest) return RCZ_OK;
    if (!c->stage_ev[i]) RCZ_CK(c, rt_event_create(&c->stage_ev[i]));
    RCZ_CK(c, rt_event_record(c->stage_ev[i], c->stream));
    c->nstage = i;
    return RCZ_OK;
}
inline int ctx_timer_begin(rcz_ctx* c) { if (c->nest) return RCZ_OK; c->ev_valid = false; c->nstage = 0; RCZ_CK(c, rt_event_record(c->ev0, c->stream)); return RCZ_OK; }
inline int ctx_timer_end(rcz_ctx* c) { if (c->nest) return RCZ_OK; RCZ_CK(c, rt_event_record(c->ev1, c->stream)); c->ev_valid = true; return RCZ_OK; }

// ------------------------------------------------------------------------------------------------
// Descriptor staging.  Every batch op has a few per-unit input arrays (offsets, lengths, ...) — always HOST
// arrays, in every mem_kind, so that the host can size grids — and a few per-unit result arrays (out_len,
// status, ...).  Inputs are packed and uploaded with one H2D from pageable memory (the driver stages it
// before returning, so the call may return before the copy executes).  Results live in a device arena and
// come back with one D2H in HOST / DEVICE mode; in DEVICE_ASYNC mode the caller's device arrays are written
// directly and nothing is synchronised.
// ------------------------------------------------------------------------------------------------
struct DescStager {
    rcz_ctx* c; int kind; size_t n;
    struct In { const void* host; size_t bytes; size_t off; };
    struct Out { void* host; size_t bytes; size_t off; };
    std::vector<In> ins; std::vector<Out> outs;
    std::vector<uint8_t> pack;
    size_t in_bytes = 0, out_bytes = 0;
    uint8_t* dev = nullptr;
    DescStager(rcz_ctx* c_, int kind_, size_t n_) : c(c_), kind(kind_), n(n_) {}
    size_t add_in(const void* host, size_t bytes) { ins.push_back({host, bytes, in_bytes}); in_bytes += (bytes + 15) & ~(size_t)15; return ins.size() - 1; }
    size_t add_out(void* host, size_t bytes) { outs.push_back({host, bytes, out_bytes}); out_bytes += (bytes + 15) & ~(size_t)15; return outs.size() - 1; }
    int upload(int slot = WS_DESC) {
        void* d;
        int st = ctx_ws(c, slot, in_bytes + out_bytes + 256, &d); if (st) return st;
        dev = (uint8_t*)d;
        pack.assign(in_bytes + 16, 0);
        for (auto& i : ins) if (i.bytes) memcpy(pack.data() + i.off, i.host, i.bytes);
        RCZ_CK(c, rt_h2d(dev, pack.data(), in_bytes, c->stream));
        return RCZ_OK;
    }
    template <class T> const T* in_ptr(size_t idx) const { return (const T*)(dev + ins[idx].off); }
    template <class T> T* out_ptr(size_t idx) const {
        if (kind == RCZ_MEM_DEVICE_ASYNC) return (T*)outs[idx].host;
        return (T*)(dev + in_bytes + outs[idx].off);
    }
    // download results + synchronise (no-op in async mode)
    int download() {
        if (kind == RCZ_MEM_DEVICE_ASYNC) return RCZ_OK;
        pack.assign(out_bytes + 16, 0);
        RCZ_CK(c, rt_d2h(pack.data(), dev + in_bytes, out_bytes, c->stream));
        RCZ_CK(c, rt_stream_sync(c->stream));
        for (auto& o : outs) if (o.bytes && o.host) memcpy(o.host, pack.data() + o.off, o.bytes);
        return RCZ_OK;
    }
};

// Host-resident data staging: upload the byte span covered by (off[i], len[i]) and return a virtual device
// base such that base + off[i] addresses block i (offsets keep their low 8 bits => same 16-byte phase).
inline int stage_span_in(rcz_ctx* c, int slot, const void* host_base, const uint64_t* off, const uint64_t* len, size_t n,
                         size_t elem, const uint8_t** dev_base) {
    uint64_t lo = UINT64_MAX, hi = 0;
    for (size_t i = 0; i < n; ++i) { if (len[i] == 0) continue; lo = off[i] < lo ? off[i] : lo; hi = off[i] + len[i] > hi ? off[i] + len[i] : hi; }
    if (lo == UINT64_MAX) { lo = 0; hi = 0; }
    uint64_t lo_al = (lo * elem) & ~(uint64_t)255;
    void* d; int st = ctx_ws(c, slot, (size_t)(hi * elem - lo_al) + 512, &d); if (st) return st;
    // the units' bytes, not the span: runs of units that touch or lie within 64 KiB of one another go up in one copy, so that an
    // arena of generous per-unit capacities (containers at 6 x their size in the bench) does not drag its gaps over PCIe
    const uint64_t GAP = 65536 / elem;
    size_t i = 0;
    while (i < n) {
        if (len[i] == 0) { ++i; continue; }
        const uint64_t s0 = off[i]; uint64_t e0 = off[i] + len[i];
        size_t j = i + 1;
        while (j < n && (len[j] == 0 || (off[j] >= s0 && off[j] <= e0 + GAP))) { if (len[j]) e0 = off[j] + len[j] > e0 ? off[j] + len[j] : e0; ++j; }
        RCZ_CK(c, rt_h2d((uint8_t*)d + (s0 * elem - lo_al), (const uint8_t*)host_base + s0 * elem, (size_t)((e0 - s0) * elem), c->stream));
        i = j;
    }
    *dev_base = (const uint8_t*)d - lo_al;
    return RCZ_OK;
}
inline int stage_span_out(rcz_ctx* c, int slot, const uint64_t* off, const uint64_t* cap, size_t n, size_t elem, uint8_t** dev_base) {
    uint64_t lo = UINT64_MAX, hi = 0;
    for (size_t i = 0; i < n; ++i) { if (cap[i] == 0) continue; lo = off[i] < lo ? off[i] : lo; hi = off[i] + cap[i] > hi ? off[i] + cap[i] : hi; }
    if (lo == UINT64_MAX) { lo = 0; hi = 0; }
    uint64_t lo_al = (lo * elem) & ~(uint64_t)255;
    void* d; int st = ctx_ws(c, slot, (size_t)(hi * elem - lo_al) + 512, &d); if (st) return st;
    *dev_base = (uint8_t*)d - lo_al;
    return RCZ_OK;
}
// copy back out_len[i] elements of every unit (adjacent units are merged into one copy); call after download()
inline int unstage_span_out(rcz_ctx* c, void* host_base, const uint8_t* dev_base, const uint64_t* off, const uint64_t* len, size_t n,
                            size_t elem) {
    size_t i = 0;
    while (i < n) {
        if (len[i] == 0) { ++i; continue; }
        uint64_t s = off[i], e = off[i] + len[i];
        size_t j = i + 1;
        while (j < n && (len[j] == 0 || off[j] == e)) { e += len[j]; ++j; }
        RCZ_CK(c, rt_d2h((uint8_t*)host_base + s * elem, dev_base + s * elem, (size_t)((e - s) * elem), c->stream));
        i = j;
    }
    RCZ_CK(c, rt_stream_sync(c->stream));
    return RCZ_OK;
}
// ------------------------------------------------------------------------------------------------
// HOST-buffer batches in pipelined chunks.  A batch op's host path is "upload everything, run, download everything": three
// phases of which two are PCIe copies.  This helper cuts the batch into chunks of consecutive units (about `chunk_bytes` of
// max(input, output capacity) each, sized by the caller so that one chunk still fills the GPU) and runs the op's own
// RCZ_MEM_DEVICE form on every chunk — call(b0, nb, dev_in_base, dev_out_base), which returns when the chunk's kernels are done —
// while the next chunk's input goes up on one copy stream and the previous chunk's output comes down on another.  Device arenas
// mirror the host arenas (same offsets), as in stage_span_in / stage_span_out.  *handled = false (nothing done) when the batch is
// too small to cut or its units are not laid out in order (a chunk's span would drag in other chunks' bytes).
// out_elems(i) = elements of unit i to bring back (looked at after the chunk's call, i.e. with its results in the host arrays).
// ------------------------------------------------------------------------------------------------
template <class Call, class OutElems>
inline int host_chunked(rcz_ctx* c, size_t n, uint64_t chunk_bytes, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                        size_t in_elem, void* out_base, const uint64_t* out_off, const uint64_t* out_cap, size_t out_elem, Call call,
                        OutElems out_elems, bool* handled) {
    *handled = false;
    if (c->nest || n < 2) return RCZ_OK;
    if (const char* e = getenv("RCZ_HOST_CHUNK_BYTES")) chunk_bytes = strtoull(e, nullptr, 10);
    if (chunk_bytes == 0) return RCZ_OK;
    std::vector<size_t> cut{0};
    uint64_t ai = 0, ao = 0;
    for (size_t i = 0; i < n; ++i) {
        ai += in_len[i] * in_elem; ao += out_cap[i] * out_elem;
        if ((ai >= chunk_bytes || ao >= chunk_bytes) && i + 1 < n) { cut.push_back(i + 1); ai = ao = 0; }
    }
    cut.push_back(n);
    const size_t nch = cut.size() - 1;
    if (nch < 2) return RCZ_OK;
    // input span of every chunk; the chunks' spans together must not be (much) more than the batch's span
    std::vector<uint64_t> lo(nch, UINT64_MAX), hi(nch, 0);
    uint64_t glo = UINT64_MAX, ghi = 0, sum = 0;
    for (size_t k = 0; k < nch; ++k) {
        for (size_t i = cut[k]; i < cut[k + 1]; ++i) {
            if (!in_len[i]) continue;
            lo[k] = std::min(lo[k], in_off[i]); hi[k] = std::max(hi[k], in_off[i] + in_len[i]);
        }
        if (lo[k] == UINT64_MAX) { lo[k] = hi[k] = 0; continue; }
        glo = std::min(glo, lo[k]); ghi = std::max(ghi, hi[k]); sum += hi[k] - lo[k];
    }
    if (glo == UINT64_MAX) return RCZ_OK;
    if (sum > (ghi - glo) + (ghi - glo) / 4 + 4096) return RCZ_OK;
    int st = ctx_aux_streams(c); if (st) return st;
    st = ctx_events(c, nch); if (st) return st;
    const uint8_t* din; uint8_t* dout;
    {
        const uint64_t lo_al = (glo * in_elem) & ~(uint64_t)255;
        void* d; st = ctx_ws(c, WS_IN, (size_t)(ghi * in_elem - lo_al) + 512, &d); if (st) return st;
        din = (const uint8_t*)d - lo_al;
    }
    st = stage_span_out(c, WS_OUT, out_off, out_cap, n, out_elem, &dout); if (st) return st;
    *handled = true;
    const rt_stream_t up = c->aux[0], down = c->aux[1];
    RCZ_CK(c, rt_stream_sync(c->stream));                                      // (workspace (re)allocation, earlier calls)
    auto upload = [&](size_t k) -> int {
        if (hi[k] > lo[k]) RCZ_CK(c, rt_h2d((uint8_t*)din + lo[k] * in_elem, (const uint8_t*)in_base + lo[k] * in_elem, (size_t)((hi[k] - lo[k]) * in_elem), up));
        RCZ_CK(c, rt_event_record(c->events[k], up));
        return RCZ_OK;
    };
    st = upload(0); if (st) return st;
    for (size_t k = 0; k < nch; ++k) {
        if (k + 1 < nch) { st = upload(k + 1); if (st) { rt_stream_sync(up); rt_stream_sync(down); return st; } }
        RCZ_CK(c, rt_stream_wait_event(c->stream, c->events[k]));
        st = call(cut[k], cut[k + 1] - cut[k], din, dout);                     // returns with the chunk's kernels done and its results on the host
        if (st) { rt_stream_sync(up); rt_stream_sync(down); return st; }       // (no copy may still be touching the caller's buffers)
        size_t i = cut[k];
        const size_t e = cut[k + 1];
        while (i < e) {                                                        // adjacent units merged into one copy
            const uint64_t li = out_elems(i);
            if (li == 0) { ++i; continue; }
            const uint64_t s0 = out_off[i]; uint64_t t = s0 + li;
            size_t j = i + 1;
            while (j < e) { const uint64_t lj = out_elems(j); if (lj && out_off[j] != t) break; t += lj; ++j; }
            RCZ_CK(c, rt_d2h((uint8_t*)out_base + s0 * out_elem, dout + s0 * out_elem, (size_t)((t - s0) * out_elem), down));
            i = j;
        }
    }
    RCZ_CK(c, rt_stream_sync(down));
    return RCZ_OK;
}

// ---- entry points with DEVICE descriptor arrays, for composed calls (pipeline.cu); they only enqueue on c->stream
constexpr unsigned long long RCZ_STREAM_SKIP = ~0ull;     // in_len value of an unused stream slot: out_len = 0, status = OK
int rcz_ari_launch(rcz_ctx* c, bool decode, const uint8_t* din, const uint64_t* d_in_off, const uint64_t* d_in_len, uint8_t* dout,
                   const uint64_t* d_out_off, const uint64_t* d_out_cap, uint64_t* d_out_len, uint64_t* d_in_used, int32_t* d_status, size_t n);
int rcz_dc_decode_launch(rcz_ctx* c, const uint32_t* din, const uint64_t* d_in_off, const uint64_t* d_in_len, uint8_t* dout,
                         const uint64_t* d_out_off, const uint64_t* d_n, int32_t* d_status, size_t nblocks);
// inverse BWT with the origins in DEVICE memory (origin_dev != nullptr; origin_host is then ignored); mem_kind as in rcz.h
int rcz_bwt_decode_run(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* n_arr, const uint32_t* origin_host,
                       const uint32_t* origin_dev, void* out_base, const uint64_t* out_off, uint64_t* out_len, int32_t* status,
                       size_t nblocks, int mem_kind);
// Every (off[i], len[i]) pair of a batch call: len < 2^31 units (rcz.h) and off + len representable; `elem` = bytes per unit.
// A caller's "unbounded" sentinel such as UINT64_MAX would otherwise wrap the span arithmetic of stage_span_in / _out.
inline bool rcz_spans_ok(const uint64_t* off, const uint64_t* len, size_t n, uint64_t elem = 1) {
    for (size_t i = 0; i < n; ++i) {
        if (len[i] >= (1ull << 31)) return false;
        if (off[i] > (1ull << 62) / elem) return false;
    }
    return true;
}
inline bool rcz_bad_kind(int k) { return k != RCZ_MEM_HOST && k != RCZ_MEM_DEVICE && k != RCZ_MEM_DEVICE_ASYNC; }
