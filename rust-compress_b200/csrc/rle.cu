// rle.cu — K11: RLE decode / encode, one warp per stream.
//
// Decode replaces the 3-state machine of /root/reference/src/rle.rs:212-259 (`Clean -> Single(b) -> Run`): at a
// token start p, `in[p+1] == in[p]` opens a run `b b v...` (7-bit little-endian groups, MSB set on the LAST
// group, value = run - 2, rle.rs:140-149, 261-263), otherwise in[p] is a literal.  A warp looks at 32 positions at
// once: the ballot of "equals its successor" gives the length of the literal stretch (copied by 32 lanes) and the
// position of the next run (expanded by 32 lanes).  A 10th length byte is "Overly long run" (rle.rs:151-154); input
// that ends inside a run flushes the partial run (rle.rs:247-256).
// Encode replaces rle.rs:62-122 for one whole-buffer write + finish: run heads come from a ballot, encoded sizes
// from a warp-shuffle prefix scan, every head lane writes its own `b` or `b b varint(n-2)`.
#include "rcz_internal.h"
#include <algorithm>

namespace rlek {

constexpr int NT = 128;

__global__ void __launch_bounds__(NT)
rle_decode_kernel(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
                  uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ out_cap,
                  uint64_t* __restrict__ out_len, int32_t* __restrict__ status, unsigned nstreams) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned wpb = NT / 32;
    for (unsigned sidx = blockIdx.x * wpb + (threadIdx.x >> 5); sidx < nstreams; sidx += gridDim.x * wpb) {
        const uint8_t* in = in_base + in_off[sidx];
        const unsigned long long n = in_len[sidx];
        uint8_t* out = out_base + out_off[sidx];
        const unsigned long long cap = out_cap[sidx];
        unsigned long long p = 0, o = 0;
        int err = 0;
        bool overlong = false;
        while (p < n) {
            const unsigned long long i = p + lane;
            const unsigned c = i < n ? (unsigned)in[i] : 256u;
            const unsigned cn = i + 1 < n ? (unsigned)in[i + 1] : 257u;
            const unsigned eq = __ballot_sync(RCZ_FULL, c == cn);
            const unsigned long long left = n - p;
            const unsigned nlit = eq ? (unsigned)__ffs((int)eq) - 1u : (unsigned)(left < 32 ? left : 32);
            if (nlit > 0) {                                       // literal stretch (rle.rs:226-228)
                if (lane < nlit) { if (o + lane < cap) out[o + lane] = (uint8_t)c; else err = RCZ_E_OUTPUT_FULL; }
                o += nlit; p += nlit;
                continue;
            }
            // run at p: bytes p, p+1 are equal; length groups start at p+2 (rle.rs:231-239)
            const unsigned byte = __shfl_sync(RCZ_FULL, c, 0);
            const unsigned long long q = p + 2 + lane;
            const bool have = lane < 10 && q < n;
            const unsigned v = have ? (unsigned)in[q] : 0u;
            const unsigned fin = __ballot_sync(RCZ_FULL, have && (v & 0x80u));
            const unsigned avail = __popc(__ballot_sync(RCZ_FULL, have));      // groups present (<= 10)
            unsigned used;                                                       // groups consumed
            if (fin && (unsigned)__ffs((int)fin) <= 9u) used = (unsigned)__ffs((int)fin);
            else if (avail >= 10) { overlong = true; break; }                     // a 10th group is read before a terminator
            else used = avail;                                                   // input exhausted inside the run: flush partial
            unsigned long long part = (lane < used) ? ((unsigned long long)(v & 0x7fu) << (7 * lane)) : 0ull;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) part |= __shfl_xor_sync(RCZ_FULL, part, d);
            const unsigned long long reps = 2ull + part;                         // rle.rs:140-149
            const unsigned long long room = o < cap ? cap - o : 0ull;
            const unsigned long long w = reps < room ? reps : room;
            for (unsigned long long k = lane; k < w; k += 32) out[o + k] = (uint8_t)byte;
            if (w < reps) err = RCZ_E_OUTPUT_FULL;
            o = (o + reps < o) ? ~0ull : o + reps;
            p += 2 + used;
        }
        err = (int)__reduce_min_sync(RCZ_FULL, (unsigned)(err + 16)) - 16;      // any lane's error (codes are negative)
        if (overlong) err = RCZ_E_OVERLONG_RUN;                                 // reported at once by the reference (rle.rs:232)
        if (lane == 0) { out_len[sidx] = o; status[sidx] = err; }
    }
}

__device__ __forceinline__ unsigned enc_size(unsigned long long len) {
    if (len == 1) return 1;
    unsigned long long v = len - 2; unsigned k = 1;
    while (v >>= 7) ++k;
    return 2 + k;
}
__device__ __forceinline__ void emit_run(uint8_t* out, unsigned long long o, unsigned long long cap, unsigned byte, unsigned long long len, int& err) {
    if (len == 1) { if (o < cap) out[o] = (uint8_t)byte; else err = RCZ_E_OUTPUT_FULL; return; }   // rle.rs:97-98
    unsigned long long v = len - 2;                                                                // rle.rs:99-118
    if (o < cap) out[o] = (uint8_t)byte; else err = RCZ_E_OUTPUT_FULL;
    if (o + 1 < cap) out[o + 1] = (uint8_t)byte; else err = RCZ_E_OUTPUT_FULL;
    unsigned k = 2;
    for (;;) {
        unsigned g = (unsigned)(v & 0x7f); v >>= 7;
        if (v == 0) g |= 0x80u;
        if (o + k < cap) out[o + k] = (uint8_t)g; else err = RCZ_E_OUTPUT_FULL;
        ++k;
        if (g & 0x80u) break;
    }
}

__global__ void __launch_bounds__(NT)
rle_encode_kernel(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
                  uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ out_cap,
                  uint64_t* __restrict__ out_len, int32_t* __restrict__ status, unsigned nstreams) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned wpb = NT / 32;
    for (unsigned sidx = blockIdx.x * wpb + (threadIdx.x >> 5); sidx < nstreams; sidx += gridDim.x * wpb) {
        const uint8_t* in = in_base + in_off[sidx];
        const unsigned long long n = in_len[sidx];
        uint8_t* out = out_base + out_off[sidx];
        const unsigned long long cap = out_cap[sidx];
        unsigned long long o = 0, pend_len = 0;                   // pending run (continues from earlier chunks)
        unsigned pend_byte = 0;
        int err = 0;
        for (unsigned long long base = 0; base < n; base += 32) {
            const unsigned long long i = base + lane;
            const bool valid = i < n;
            const unsigned c = valid ? (unsigned)in[i] : 0u;
            const bool head = valid && (i == 0 || c != (unsigned)in[i - 1]);
            const unsigned hm = __ballot_sync(RCZ_FULL, head);
            const unsigned nvalid = __popc(__ballot_sync(RCZ_FULL, valid));
            if (hm == 0) { pend_len += nvalid; continue; }
            const unsigned first = (unsigned)__ffs((int)hm) - 1u;
            const unsigned last = 31u - (unsigned)__clz((int)hm);
            // the pending run ends at `first`
            unsigned long long plen = pend_len + first;
            unsigned psize = plen ? enc_size(plen) : 0;
            // complete runs inside the chunk: every head except the last one
            unsigned mylen = 0;
            if (head && lane != last) { const unsigned above = hm & ~((2u << lane) - 1u); mylen = (unsigned)__ffs((int)above) - 1u - lane; }
            const unsigned mysize = mylen ? enc_size(mylen) : 0;
            const unsigned incl = warp_incl_scan_add(mysize);
            const unsigned total = __shfl_sync(RCZ_FULL, incl, 31);
            if (lane == 0 && plen) emit_run(out, o, cap, pend_byte, plen, err);
            if (mylen) emit_run(out, o + psize + (incl - mysize), cap, c, mylen, err);
            o += psize + total;
            pend_byte = __shfl_sync(RCZ_FULL, c, (int)last);
            pend_len = nvalid - last;
        }
        if (lane == 0 && pend_len) emit_run(out, o, cap, pend_byte, pend_len, err);   // finish -> flush (rle.rs:62-66)
        if (pend_len) o += enc_size(pend_len);
        err = (int)__reduce_min_sync(RCZ_FULL, (unsigned)(err + 16)) - 16;
        if (lane == 0) { out_len[sidx] = o; status[sidx] = err; }
    }
}

}  // namespace rlek

static int rle_batch(rcz_ctx* c, bool decode, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                     const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t n, int mem_kind) {
    if (!c || rcz_bad_kind(mem_kind)) return RCZ_E_ARG;
    if (n == 0) return RCZ_OK;
    if (!in_base || !in_off || !in_len || !out_base || !out_off || !out_cap || !out_len || !status || n > 0x7fffffffu) return RCZ_E_ARG;
    if (!rcz_spans_ok(in_off, in_len, n) || !rcz_spans_ok(out_off, out_cap, n)) return RCZ_E_ARG;
    rt_set_device(c->device);
    DescStager ds(c, mem_kind, n);
    ds.add_in(in_off, n * 8); ds.add_in(in_len, n * 8); ds.add_in(out_off, n * 8); ds.add_in(out_cap, n * 8);
    ds.add_out(out_len, n * 8); ds.add_out(status, n * 4);
    int st = ds.upload(); if (st) return st;
    const uint8_t* din = (const uint8_t*)in_base; uint8_t* dout = (uint8_t*)out_base;
    if (mem_kind == RCZ_MEM_HOST) {
        st = stage_span_in(c, WS_IN, in_base, in_off, in_len, n, 1, &din); if (st) return st;
        st = stage_span_out(c, WS_OUT, out_off, out_cap, n, 1, &dout); if (st) return st;
    }
    const unsigned grid = (unsigned)std::min<size_t>((n + 3) / 4, (size_t)c->sm_count * 16);
    st = ctx_timer_begin(c); if (st) return st;
    if (decode)
        RCZ_KLAUNCH(c, rlek::rle_decode_kernel, grid, rlek::NT, 0, din, ds.in_ptr<uint64_t>(0), ds.in_ptr<uint64_t>(1), dout, ds.in_ptr<uint64_t>(2),
                    ds.in_ptr<uint64_t>(3), ds.out_ptr<uint64_t>(0), ds.out_ptr<int32_t>(1), (unsigned)n);
    else
        RCZ_KLAUNCH(c, rlek::rle_encode_kernel, grid, rlek::NT, 0, din, ds.in_ptr<uint64_t>(0), ds.in_ptr<uint64_t>(1), dout, ds.in_ptr<uint64_t>(2),
                    ds.in_ptr<uint64_t>(3), ds.out_ptr<uint64_t>(0), ds.out_ptr<int32_t>(1), (unsigned)n);
    st = ctx_timer_end(c); if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) {
        std::vector<uint64_t> clipped(n);
        for (size_t i = 0; i < n; ++i) clipped[i] = out_len[i] < out_cap[i] ? out_len[i] : out_cap[i];
        st = unstage_span_out(c, out_base, dout, out_off, clipped.data(), n, 1); if (st) return st;
    }
    return RCZ_OK;
}

extern "C" int rcz_rle_decode_streams(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                                      const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t n,
                                      int mem_kind) {
    return rle_batch(c, true, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, n, mem_kind);
}
extern "C" int rcz_rle_encode_streams(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                                      const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t n,
                                      int mem_kind) {
    return rle_batch(c, false, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, n, mem_kind);
}
