// ari.cu — K7/K8: adaptive order-0 range coder, one warp per stream.
//
// Replaces /root/reference/src/entropy/ari/table.rs:203-219 (`ByteEncoder::write` + `finish`) and :255-272
// (`ByteDecoder::read`), i.e. `RangeEncoder::process` (ari/mod.rs:117-150) driven by the 257-symbol frequency table
// `table::Model` (table.rs:20-122: counts start at 1, `add = (total>>10)+1`, halve-with-round-up at total >= 4096).
// The chain is strictly serial in the symbol index (adaptive model + carried low/hai), so parallelism is across
// streams; inside a warp the reference's O(257) linear sums (table.rs:100-117) become warp-shuffle reductions/scans
// over a frequency table distributed 9 bins per lane.
#include "rcz_internal.h"
#include <algorithm>

namespace arik {

constexpr int NT = 128;
constexpr unsigned SYMBOL_MASK = 0xFF000000u;     // ari/mod.rs:59
constexpr unsigned THRESHOLD = 1u << 14;          // ari/mod.rs:61
constexpr unsigned CUT = THRESHOLD >> 2;          // table.rs:195

// frequency table: bin b lives in lane b / 9, slot b % 9
struct Model {
    unsigned f[9];
    unsigned lsum;     // sum of this lane's bins
    unsigned total;    // warp-uniform
    __device__ __forceinline__ void init(unsigned lane) {
        lsum = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) { f[k] = (lane * 9 + k) < 257u ? 1u : 0u; lsum += f[k]; }
        total = 257;
    }
    // table.rs:100-103 get_range: lo = sum of bins below v, hi = lo + f[v]
    __device__ __forceinline__ void range_of(unsigned v, unsigned lane, unsigned& lo, unsigned& hi) const {
        const unsigned lv = v / 9, kv = v - lv * 9;
        unsigned part = 0, fv = 0;
        if (lane < lv) part = lsum;
        else if (lane == lv) {
#pragma unroll
            for (int k = 0; k < 9; ++k) { if ((unsigned)k < kv) part += f[k]; if ((unsigned)k == kv) fv = f[k]; }
        }
        lo = warp_reduce_add(part);
        hi = lo + __shfl_sync(RCZ_FULL, fv, (int)lv);
    }
    // table.rs:105-117 find_value: first v with cumulative(v+1) > offset
    __device__ __forceinline__ void find(unsigned offset, unsigned lane, unsigned& v, unsigned& lo, unsigned& hi) const {
        const unsigned incl = warp_incl_scan_add(lsum);
        const unsigned excl = incl - lsum;
        const unsigned hit = __ballot_sync(RCZ_FULL, offset < incl);
        const unsigned lv = (unsigned)__ffs((int)hit) - 1u;
        unsigned myv = 0, mylo = 0, myhi = 0;
        if (lane == lv) {
            unsigned c = excl; bool found = false;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const unsigned nxt = c + f[k];
                if (!found && offset < nxt) { found = true; myv = lane * 9 + k; mylo = c; myhi = nxt; }
                c = nxt;
            }
        }
        v = __shfl_sync(RCZ_FULL, myv, (int)lv);
        lo = __shfl_sync(RCZ_FULL, mylo, (int)lv);
        hi = __shfl_sync(RCZ_FULL, myhi, (int)lv);
    }
    // table.rs:69-91 update(value, 10, 1) + downscale
    __device__ __forceinline__ void update(unsigned v, unsigned lane) {
        const unsigned add = (total >> 10) + 1;
        const unsigned lv = v / 9, kv = v - lv * 9;
        if (lane == lv) {
#pragma unroll
            for (int k = 0; k < 9; ++k) if ((unsigned)k == kv) f[k] = (f[k] + add) & 0xffffu;   // Frequency = u16
            lsum += add;
        }
        total += add;
        if (total >= CUT) {
            lsum = 0;
#pragma unroll
            for (int k = 0; k < 9; ++k) { f[k] = (f[k] + 1) >> 1; lsum += f[k]; }
            total = warp_reduce_add(lsum);
        }
    }
};

// ari/mod.rs:117-150 process(): returns the number of bytes shifted out (0..4) in `nout`, bytes MSB-first in `bytes`
__device__ __forceinline__ void range_step(unsigned& low, unsigned& hai, unsigned total, unsigned from, unsigned to, unsigned& bytes, unsigned& nout) {
    const unsigned range = (hai - low) / total;
    unsigned lo = low + range * from, hi = low + range * to;
    bytes = 0; nout = 0;
    for (;;) {
        if ((lo ^ hi) & SYMBOL_MASK) {
            if (hi - lo > THRESHOLD) break;
            const unsigned lim = hi & SYMBOL_MASK;
            if (hi - lim >= lim - lo) lo = lim; else hi = lim - 1;
        }
        if (nout >= 4) break;                          // cannot happen (the reference would index out of bounds)
        bytes = (bytes << 8) | (lo >> 24);
        ++nout;
        lo <<= 8; hi <<= 8;
    }
    low = lo; hai = hi;
}

__global__ void __launch_bounds__(NT)
ari_encode_kernel(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
                  uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ out_cap,
                  uint64_t* __restrict__ out_len, int32_t* __restrict__ status, unsigned nstreams) {
    const unsigned lane = threadIdx.x & 31, wpb = NT / 32;
    for (unsigned sidx = blockIdx.x * wpb + (threadIdx.x >> 5); sidx < nstreams; sidx += gridDim.x * wpb) {
        const uint8_t* in = in_base + in_off[sidx];
        const unsigned long long n = in_len[sidx];
        uint8_t* out = out_base + out_off[sidx];
        const unsigned long long cap = out_cap[sidx];
        if (n == RCZ_STREAM_SKIP) { if (lane == 0) { out_len[sidx] = 0; status[sidx] = RCZ_OK; } continue; }   // unused slot of a composed call
        Model m; m.init(lane);
        unsigned low = 0, hai = 0xFFFFFFFFu;
        unsigned long long o = 0;
        bool full = false;
        for (unsigned long long i = 0; i <= n; i += 32) {
            const unsigned chunk = (i + lane < n) ? (unsigned)in[i + lane] : 256u;      // 32 symbols per load
            const unsigned cnt = (unsigned)((n - i) < 32 ? (n - i) + 1 : 32);           // +1: the terminator (table.rs:203-204)
            for (unsigned k = 0; k < cnt; ++k) {
                const unsigned v = __shfl_sync(RCZ_FULL, chunk, (int)k);
                unsigned lo, hi, bytes, nout;
                m.range_of(v, lane, lo, hi);
                range_step(low, hai, m.total, lo, hi, bytes, nout);
                if (lane < nout) {                                                       // ari/mod.rs:223-227
                    if (o + lane < cap) out[o + lane] = (uint8_t)(bytes >> (8 * (nout - 1 - lane))); else full = true;
                }
                o += nout;
                if (v != 256u) m.update(v, lane);                                        // table.rs:215; no update after the terminator
            }
            if (cnt < 32 || i + 32 > n) break;
        }
        if (lane < 4) { if (o + lane < cap) out[o + lane] = (uint8_t)(low >> (8 * (3 - lane))); else full = true; }   // ari/mod.rs:230-237
        o += 4;
        full = __any_sync(RCZ_FULL, full);
        if (lane == 0) { out_len[sidx] = o; status[sidx] = full ? RCZ_E_OUTPUT_FULL : RCZ_OK; }
    }
}

__global__ void __launch_bounds__(NT)
ari_decode_kernel(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
                  uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ out_cap,
                  uint64_t* __restrict__ out_len, uint64_t* __restrict__ in_used, int32_t* __restrict__ status, unsigned nstreams) {
    const unsigned lane = threadIdx.x & 31, wpb = NT / 32;
    for (unsigned sidx = blockIdx.x * wpb + (threadIdx.x >> 5); sidx < nstreams; sidx += gridDim.x * wpb) {
        const uint8_t* in = in_base + in_off[sidx];
        const unsigned long long n = in_len[sidx];
        uint8_t* out = out_base + out_off[sidx];
        const unsigned long long cap = out_cap[sidx];
        if (n == RCZ_STREAM_SKIP) { if (lane == 0) { out_len[sidx] = 0; status[sidx] = RCZ_OK; if (in_used) in_used[sidx] = 0; } continue; }
        Model m; m.init(lane);
        unsigned low = 0, hai = 0xFFFFFFFFu, code = 0, pending = 4;
        unsigned long long p = 0, o = 0;
        int err = 0;
        unsigned obuf = 0, ocnt = 0;                       // output bytes are produced one per step: lane k keeps byte k of a 32-byte group
        for (;;) {
            // feed(): ari/mod.rs:271-278 (.unwrap() => panic on a truncated stream)
            if (p + pending > n) { err = RCZ_E_MALFORMED; break; }
            for (unsigned k = 0; k < pending; ++k) code = (code << 8) + (unsigned)in[p + k];
            p += pending;
            const unsigned range = (hai - low) / m.total;  // ari/mod.rs:153-159 query()
            if (range == 0) { err = RCZ_E_MALFORMED; break; }
            const unsigned offset = (code - low) / range;
            if (offset >= m.total) { err = RCZ_E_MALFORMED; break; }   // table.rs:106 assert!
            unsigned v, lo, hi, bytes, nout;
            m.find(offset, lane, v, lo, hi);
            range_step(low, hai, m.total, lo, hi, bytes, nout);        // ari/mod.rs:199: re-run the step to learn the shift
            pending = nout;
            if (v == 256u) break;                                      // table.rs:263-266 terminator
            m.update(v, lane);
            if (o >= cap) { err = RCZ_E_OUTPUT_FULL; break; }
            if (lane == ocnt) obuf = v;
            ++ocnt; ++o;
            if (ocnt == 32) { out[o - 32 + lane] = (uint8_t)obuf; ocnt = 0; }
        }
        if (lane < ocnt) out[o - ocnt + lane] = (uint8_t)obuf;
        if (lane == 0) {
            out_len[sidx] = o; status[sidx] = err;
            if (in_used) in_used[sidx] = p + (err ? 0 : pending);      // incl. the bytes only finish() consumes (ari/mod.rs:289-292)
        }
    }
}

}  // namespace arik

// all descriptor / result arrays are DEVICE pointers (composed calls: pipeline.cu); only enqueues
int rcz_ari_launch(rcz_ctx* c, bool decode, const uint8_t* din, const uint64_t* d_in_off, const uint64_t* d_in_len, uint8_t* dout,
                   const uint64_t* d_out_off, const uint64_t* d_out_cap, uint64_t* d_out_len, uint64_t* d_in_used, int32_t* d_status, size_t n) {
    if (n == 0) return RCZ_OK;
    const unsigned grid = (unsigned)std::min<size_t>((n + 3) / 4, (size_t)c->sm_count * 16);
    if (decode)
        RCZ_KLAUNCH(c, arik::ari_decode_kernel, grid, arik::NT, 0, din, d_in_off, d_in_len, dout, d_out_off, d_out_cap, d_out_len, d_in_used, d_status, (unsigned)n);
    else
        RCZ_KLAUNCH(c, arik::ari_encode_kernel, grid, arik::NT, 0, din, d_in_off, d_in_len, dout, d_out_off, d_out_cap, d_out_len, d_status, (unsigned)n);
    return RCZ_OK;
}

static int ari_batch(rcz_ctx* c, bool decode, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                     const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, uint64_t* in_used, int32_t* status, size_t n,
                     int mem_kind) {
    if (!c || rcz_bad_kind(mem_kind)) return RCZ_E_ARG;
    if (n == 0) return RCZ_OK;
    if (!in_base || !in_off || !in_len || !out_base || !out_off || !out_cap || !out_len || !status || n > 0x7fffffffu) return RCZ_E_ARG;
    if (!rcz_spans_ok(in_off, in_len, n) || !rcz_spans_ok(out_off, out_cap, n)) return RCZ_E_ARG;
    rt_set_device(c->device);
    DescStager ds(c, mem_kind, n);
    ds.add_in(in_off, n * 8); ds.add_in(in_len, n * 8); ds.add_in(out_off, n * 8); ds.add_in(out_cap, n * 8);
    ds.add_out(out_len, n * 8); ds.add_out(status, n * 4);
    size_t o_used = 0;
    if (decode) o_used = ds.add_out(in_used, in_used ? n * 8 : 0);
    int st = ds.upload(); if (st) return st;
    const uint8_t* din = (const uint8_t*)in_base; uint8_t* dout = (uint8_t*)out_base;
    if (mem_kind == RCZ_MEM_HOST) {
        st = stage_span_in(c, WS_IN, in_base, in_off, in_len, n, 1, &din); if (st) return st;
        st = stage_span_out(c, WS_OUT, out_off, out_cap, n, 1, &dout); if (st) return st;
    }
    st = ctx_timer_begin(c); if (st) return st;
    st = rcz_ari_launch(c, decode, din, ds.in_ptr<uint64_t>(0), ds.in_ptr<uint64_t>(1), dout, ds.in_ptr<uint64_t>(2), ds.in_ptr<uint64_t>(3),
                        ds.out_ptr<uint64_t>(0), (decode && in_used) ? ds.out_ptr<uint64_t>(o_used) : (uint64_t*)nullptr, ds.out_ptr<int32_t>(1), n);
    if (st) return st;
    st = ctx_timer_end(c); if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) {
        std::vector<uint64_t> clipped(n);
        for (size_t i = 0; i < n; ++i) clipped[i] = out_len[i] < out_cap[i] ? out_len[i] : out_cap[i];
        st = unstage_span_out(c, out_base, dout, out_off, clipped.data(), n, 1); if (st) return st;
    }
    return RCZ_OK;
}

extern "C" int rcz_ari_encode_streams(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                                      const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t n,
                                      int mem_kind) {
    return ari_batch(c, false, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, nullptr, status, n, mem_kind);
}
extern "C" int rcz_ari_decode_streams(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                                      const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, uint64_t* in_used, int32_t* status,
                                      size_t n, int mem_kind) {
    return ari_batch(c, true, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, in_used, status, n, mem_kind);
}
