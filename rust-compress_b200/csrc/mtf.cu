// mtf.cu — move-to-front stream coder: `bwt::mtf::Encoder<W>::write` / `Decoder<R>::read`
// (/root/reference/src/bwt/mtf.rs:95-169) over `MTF::encode` / `MTF::decode` (mtf.rs:63-91) with the alphabetical start
// list (mtf.rs:56-60, 101-102, 138-139).
//
// The list is one dependency chain per stream, so a warp owns a stream and keeps the 256-entry list in REGISTERS: lane l
// holds ranks 8l..8l+7 as one 64-bit word.  Encode finds the symbol with a zero-byte test per lane + one ballot; decode
// fetches rank r from lane r/8 with one shuffle; "move to front" is a byte shift of every word below the rank with the
// carry byte handed down by a shuffle.  Input is read and output written 32 bytes per warp at a time (coalesced).
// Throughput scales with the number of streams (C5 runs one stream per 4 MiB block).
#include "rcz_internal.h"
#include <algorithm>

namespace mtfk {

constexpr int NT = 128, WPB = NT / 32;
constexpr unsigned long long ONES = 0x0101010101010101ull, HIGH = 0x8080808080808080ull;

// ranks 0..pos move up by one, `sym` goes to rank 0 (mtf.rs:68-78 / 84-89).  pos = 8 * lr + bi.
__device__ __forceinline__ void move_to_front(unsigned long long& L, unsigned lane, unsigned lr, unsigned bi, unsigned sym) {
    unsigned carry = __shfl_up_sync(RCZ_FULL, (unsigned)(L >> 56), 1);
    if (lane == 0) carry = sym;
    const unsigned long long shifted = (L << 8) | carry;
    if (lane < lr) L = shifted;
    else if (lane == lr) {
        const unsigned long long keep = bi == 7 ? 0ull : ~0ull << (8u * (bi + 1u));      // ranks above pos stay
        L = (L & keep) | (shifted & ~keep);
    }
}

template <bool DECODE>
__global__ void __launch_bounds__(NT)
mtf_kernel(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
           uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ out_cap,
           uint64_t* __restrict__ out_len, int32_t* __restrict__ status, unsigned n) {
    const unsigned lane = threadIdx.x & 31;
    for (unsigned s = blockIdx.x * WPB + (threadIdx.x >> 5); s < n; s += gridDim.x * WPB) {
        const uint8_t* in = in_base + in_off[s];
        uint8_t* out = out_base + out_off[s];
        const unsigned long long len = in_len[s], cap = out_cap[s];
        const unsigned long long m = len < cap ? len : cap;
        unsigned long long L = 0;                                              // reset_alphabetical (mtf.rs:56-60): rank i holds symbol i
#pragma unroll
        for (int k = 0; k < 8; ++k) L |= (unsigned long long)(8u * lane + (unsigned)k) << (8 * k);
        unsigned front = 0;                                                    // the symbol at rank 0 (warp-uniform)
        for (unsigned long long g = 0; g < m; g += 32) {
            const unsigned cnt = m - g < 32 ? (unsigned)(m - g) : 32u;
            const unsigned mine = lane < cnt ? in[g + lane] : 0u;
            unsigned res = 0;
            for (unsigned k = 0; k < cnt; ++k) {
                const unsigned x = __shfl_sync(RCZ_FULL, mine, (int)k);
                // rank 0 (the symbol that is already in front — every repeat in a BWT column): nothing moves, nothing to look up
                if (DECODE ? x == 0u : x == front) { if (lane == k) res = DECODE ? front : 0u; continue; }
                unsigned lr, bi, sym, rank;
                if (DECODE) {                                                  // mtf.rs:82-91
                    rank = x; lr = x >> 3; bi = x & 7u;
                    sym = __shfl_sync(RCZ_FULL, (unsigned)(L >> (8u * bi)) & 255u, (int)lr);
                } else {                                                       // mtf.rs:63-79: the one lane whose word holds the symbol
                    sym = x;
                    const unsigned long long z = L ^ (ONES * x);
                    const unsigned long long t = (z - ONES) & ~z & HIGH;       // lowest set bit marks the (only) zero byte
                    const unsigned hit = __ballot_sync(RCZ_FULL, t != 0);
                    lr = (unsigned)__ffs((int)hit) - 1u;
                    bi = __shfl_sync(RCZ_FULL, (unsigned)(__ffsll((long long)t) - 1) >> 3, (int)lr);
                    rank = 8u * lr + bi;
                }
                if (lane == k) res = DECODE ? sym : rank;
                move_to_front(L, lane, lr, bi, sym);
                front = sym;
            }
            if (lane < cnt) out[g + lane] = (uint8_t)res;
        }
        if (lane == 0) { out_len[s] = m; status[s] = len > cap ? RCZ_E_OUTPUT_FULL : RCZ_OK; }
    }
}

}  // namespace mtfk

static int mtf_batch(rcz_ctx* c, bool decode, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
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
    const unsigned grid = (unsigned)std::min<size_t>((n + mtfk::WPB - 1) / mtfk::WPB, (size_t)c->sm_count * 16);
    st = ctx_timer_begin(c); if (st) return st;
    if (decode)
        RCZ_KLAUNCH(c, mtfk::mtf_kernel<true>, grid, mtfk::NT, 0, din, ds.in_ptr<uint64_t>(0), ds.in_ptr<uint64_t>(1), dout, ds.in_ptr<uint64_t>(2),
                    ds.in_ptr<uint64_t>(3), ds.out_ptr<uint64_t>(0), ds.out_ptr<int32_t>(1), (unsigned)n);
    else
        RCZ_KLAUNCH(c, mtfk::mtf_kernel<false>, grid, mtfk::NT, 0, din, ds.in_ptr<uint64_t>(0), ds.in_ptr<uint64_t>(1), dout, ds.in_ptr<uint64_t>(2),
                    ds.in_ptr<uint64_t>(3), ds.out_ptr<uint64_t>(0), ds.out_ptr<int32_t>(1), (unsigned)n);
    st = ctx_timer_end(c); if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) { st = unstage_span_out(c, out_base, dout, out_off, out_len, n, 1); if (st) return st; }
    return RCZ_OK;
}

extern "C" int rcz_mtf_encode_streams(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                                      const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t n, int mem_kind) {
    return mtf_batch(c, false, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, n, mem_kind);
}
extern "C" int rcz_mtf_decode_streams(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                                      const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t n, int mem_kind) {
    return mtf_batch(c, true, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, n, mem_kind);
}
