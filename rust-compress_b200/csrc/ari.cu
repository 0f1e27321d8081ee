// ari.cu — K7/K8: adaptive order-0 range coder, one THREAD per stream.
//
// Replaces /root/reference/src/entropy/ari/table.rs:203-219 (`ByteEncoder::write` + `finish`) and :255-272
// (`ByteDecoder::read`), i.e. `RangeEncoder::process` (ari/mod.rs:117-150) driven by the 257-symbol frequency table
// `table::Model` (table.rs:20-122: counts start at 1, `add = (total>>10)+1`, halve-with-round-up at total >= 4096).
// The chain is strictly serial in the symbol index (adaptive model + carried low/hai), so parallelism is across
// streams.  A stream is one thread: its 257-bin model lives in shared memory (u16 counts, two per word, plus 17 group sums),
// which turns the reference's O(257) linear sums (table.rs:100-117) into <= 17 + 16 terms, and 32 streams share every issued
// instruction (the round-1 kernel gave a whole warp to one stream: same chain length, 1/32 of the throughput).
// Input and output run through per-thread register buffers (aligned 16-byte loads, aligned 4-byte stores).
#include "rcz_internal.h"
#include <algorithm>

namespace arik {

constexpr int NT = 128;                            // streams (threads) per CTA
constexpr unsigned SYMBOL_MASK = 0xFF000000u;     // ari/mod.rs:59
constexpr unsigned THRESHOLD = 1u << 14;          // ari/mod.rs:61
constexpr unsigned CUT = THRESHOLD >> 2;          // table.rs:195
constexpr int NBIN = 272;                         // 257 bins in 17 groups of 16 (bins 257.. stay 0)
constexpr int ROW = NBIN / 2 + 9;                 // 32-bit words per stream: 136 of bins (two u16 each) + 9 of group sums; odd => rows spread over all banks
static_assert(ROW % 2 == 1, "odd row stride");

// table::Model (table.rs:20-122) of ONE stream, owned by one thread, in shared memory: u16 counts packed two per word plus the
// sums of 17 groups of 16 bins, so that the reference's O(257) linear sums (table.rs:100-117) become <= 17 + 16 terms.
struct Model {
    unsigned* w;        // bins: w[0..136), group sums: w[136..145)
    unsigned total;
    __device__ __forceinline__ void init(unsigned* row) {
        w = row;
        for (int i = 0; i < 128; ++i) w[i] = 0x00010001u;                 // bins 0..255 = 1
        w[128] = 0x00000001u;                                             // bin 256 (terminator) = 1, bin 257 = 0
        for (int i = 129; i < 136; ++i) w[i] = 0;
        for (int g = 0; g < 8; ++g) w[136 + g] = 0x00100010u;             // groups 0..15: 16 each
        w[144] = 0x00000001u;                                             // group 16: 1
        total = 257;
    }
    // sum of the first k u16 values of the packed array a[0..nw): branch-free (every stream of the warp runs the same instructions)
    template <int NW>
    __device__ __forceinline__ unsigned prefix(const unsigned* a, unsigned k) const {
        unsigned acc = 0;
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            const unsigned x = a[j];
            const unsigned lo = x & 0xffffu, hi = x >> 16;
            acc += (k > 2u * j ? lo : 0u) + (k > 2u * j + 1u ? hi : 0u);
        }
        return acc;
    }
    // table.rs:100-103 get_range
    __device__ __forceinline__ void range_of(unsigned v, unsigned& lo, unsigned& hi) const {
        const unsigned g = v >> 4, k = v & 15u;
        lo = prefix<9>(w + 136, g) + prefix<8>(w + 8 * g, k);
        const unsigned x = w[v >> 1];
        hi = lo + ((v & 1u) ? x >> 16 : x & 0xffffu);
    }
    // table.rs:105-117 find_value: first v with cumulative(v + 1) > offset (offset < total)
    __device__ __forceinline__ void find(unsigned offset, unsigned& v, unsigned& lo, unsigned& hi) const {
        unsigned acc = 0, g = 0, base = 0;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const unsigned x = w[136 + j];
            unsigned inc = acc + (x & 0xffffu);
            bool c = inc <= offset; g += c; base = c ? inc : base; acc = inc;
            if (j < 8) { inc = acc + (x >> 16); c = inc <= offset; g += c; base = c ? inc : base; acc = inc; }
        }
        g = g > 16u ? 16u : g;                                            // (offset < total always lands in a group)
        const unsigned* a = w + 8 * g;
        unsigned k = 0, lo_ = base, hi_ = base;
        acc = base;
        bool open = true;                                                 // still looking for the bin
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned x = a[j];
            unsigned f = x & 0xffffu, inc = acc + f;
            bool hit = open && offset < inc;
            if (hit) { lo_ = acc; hi_ = inc; k = 2 * j; open = false; }
            acc = inc;
            f = x >> 16; inc = acc + f;
            hit = open && offset < inc;
            if (hit) { lo_ = acc; hi_ = inc; k = 2 * j + 1; open = false; }
            acc = inc;
        }
        v = 16u * g + k; lo = lo_; hi = hi_;
    }
    // table.rs:69-91 update(value, 10, 1) + downscale
    __device__ __forceinline__ void update(unsigned v) {
        const unsigned add = (total >> 10) + 1;
        const unsigned sh = (v & 1u) * 16u;
        unsigned x = w[v >> 1];
        x = (x & ~(0xffffu << sh)) | ((((x >> sh) + add) & 0xffffu) << sh);   // Frequency = u16
        w[v >> 1] = x;
        const unsigned g = v >> 4, gs = (g & 1u) * 16u;
        w[136 + (g >> 1)] += add << gs;                                   // group sums stay far below 2^16
        total += add;
        if (total >= CUT) {
            unsigned tot = 0;
            for (int g2 = 0; g2 < 17; ++g2) {
                unsigned gsum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    unsigned y = w[8 * g2 + j];
                    y = ((y + 0x00010001u) >> 1) & 0x7fff7fffu;           // (f + 1) >> 1 on both halves (counts < 2^15: no carry between them)
                    w[8 * g2 + j] = y;
                    gsum += (y & 0xffffu) + (y >> 16);
                }
                const unsigned s2 = (g2 & 1u) * 16u;
                w[136 + (g2 >> 1)] = (w[136 + (g2 >> 1)] & ~(0xffffu << s2)) | (gsum << s2);
                tot += gsum;
            }
            total = tot;
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

// sequential byte reader of one thread: aligned 16-byte loads, the next one in flight while the current one is consumed
struct ByteIn {
    // 16 bytes in (lo, hi), the 16 after them already loaded (nx): every lane of a warp runs dry at its own symbol, so a refill that
    // waits for its load stalls the whole warp on nearly every symbol (12.5 % of the decode kernel's stall samples before)
    const uint4* vp; const uint4* vend; unsigned long long lo, hi; uint4 nx; unsigned left;
    __device__ __forceinline__ void init(const uint8_t* p, unsigned long long n) {
        const unsigned mis = (unsigned)((uintptr_t)p & 15u);
        vp = reinterpret_cast<const uint4*>(p - mis);
        vend = reinterpret_cast<const uint4*>(p + ((n + mis + 15u) & ~15ull) - mis);      // reads stay inside the 16-byte-aligned span of the stream (rcz.h)
        nx = __ldg(vp); ++vp;
        load();
        for (unsigned i = 0; i < mis; ++i) shift();
        left = 16u - mis;
    }
    __device__ __forceinline__ void load() {
        lo = (unsigned long long)nx.x | ((unsigned long long)nx.y << 32); hi = (unsigned long long)nx.z | ((unsigned long long)nx.w << 32);
        nx = vp < vend ? __ldg(vp) : make_uint4(0, 0, 0, 0);
        ++vp;
    }
    __device__ __forceinline__ void shift() { lo = (lo >> 8) | (hi << 56); hi >>= 8; }
    // the caller guarantees that the byte exists
    __device__ __forceinline__ unsigned next() {
        if (left == 0) { load(); left = 16; }
        const unsigned b = (unsigned)lo & 255u;
        shift(); --left;
        return b;
    }
};

// sequential byte writer of one thread: head bytes up to the first 4-byte boundary one by one, then aligned 32-bit stores
struct ByteOut {
    uint8_t* out; unsigned long long cap, o; unsigned acc, cnt; bool full;
    __device__ __forceinline__ void init(uint8_t* p, unsigned long long c) { out = p; cap = c; o = 0; acc = 0; cnt = 0; full = false; }
    __device__ __forceinline__ void put(unsigned b) {
        if (o >= cap) { if (cnt) flush(); full = true; ++o; return; }             // out_len keeps counting: the size the stream would have needed
        if (((uintptr_t)(out + o) & 3u) != 0 && cnt == 0) { out[o++] = (uint8_t)b; return; }      // head (or an unaligned stream start)
        acc |= b << (8u * cnt); ++cnt; ++o;
        if (cnt == 4) { *reinterpret_cast<unsigned*>(out + o - 4) = acc; acc = 0; cnt = 0; }
    }
    __device__ __forceinline__ void flush() { for (unsigned i = 0; i < cnt; ++i) out[o - cnt + i] = (uint8_t)(acc >> (8u * i)); cnt = 0; acc = 0; }
};

__global__ void __launch_bounds__(NT)
ari_encode_kernel(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
                  uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ out_cap,
                  uint64_t* __restrict__ out_len, int32_t* __restrict__ status, unsigned nstreams) {
    RCZ_DYN_SMEM(raw);
    unsigned* row = reinterpret_cast<unsigned*>(raw) + threadIdx.x * ROW;
    for (unsigned sidx = blockIdx.x * NT + threadIdx.x; sidx < nstreams; sidx += gridDim.x * NT) {
        const unsigned long long n = in_len[sidx];
        if (n == RCZ_STREAM_SKIP) { out_len[sidx] = 0; status[sidx] = RCZ_OK; continue; }          // unused slot of a composed call
        Model m; m.init(row);
        ByteIn bi; if (n) bi.init(in_base + in_off[sidx], n);
        ByteOut bo; bo.init(out_base + out_off[sidx], out_cap[sidx]);
        unsigned low = 0, hai = 0xFFFFFFFFu;
        for (unsigned long long i = 0; i <= n; ++i) {
            const unsigned v = i < n ? bi.next() : 256u;                                          // the terminator follows the data (table.rs:203-204)
            unsigned lo, hi, bytes, nout;
            m.range_of(v, lo, hi);
            range_step(low, hai, m.total, lo, hi, bytes, nout);
            for (unsigned k = 0; k < nout; ++k) bo.put((bytes >> (8u * (nout - 1u - k))) & 255u);   // ari/mod.rs:223-227
            if (v != 256u) m.update(v);                                                            // table.rs:215; no update after the terminator
        }
        for (unsigned k = 0; k < 4; ++k) bo.put((low >> (8u * (3u - k))) & 255u);                  // ari/mod.rs:230-237
        bo.flush();
        out_len[sidx] = bo.o; status[sidx] = bo.full ? RCZ_E_OUTPUT_FULL : RCZ_OK;
    }
}

__global__ void __launch_bounds__(NT)
ari_decode_kernel(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
                  uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ out_cap,
                  uint64_t* __restrict__ out_len, uint64_t* __restrict__ in_used, int32_t* __restrict__ status, unsigned nstreams) {
    RCZ_DYN_SMEM(raw);
    unsigned* row = reinterpret_cast<unsigned*>(raw) + threadIdx.x * ROW;
    for (unsigned sidx = blockIdx.x * NT + threadIdx.x; sidx < nstreams; sidx += gridDim.x * NT) {
        const unsigned long long n = in_len[sidx];
        if (n == RCZ_STREAM_SKIP) { out_len[sidx] = 0; status[sidx] = RCZ_OK; if (in_used) in_used[sidx] = 0; continue; }
        Model m; m.init(row);
        ByteIn bi; if (n) bi.init(in_base + in_off[sidx], n);
        ByteOut bo; bo.init(out_base + out_off[sidx], out_cap[sidx]);
        unsigned low = 0, hai = 0xFFFFFFFFu, code = 0, pending = 4;
        unsigned long long p = 0;
        int err = 0;
        for (;;) {
            // feed(): ari/mod.rs:271-278 (.unwrap() => panic on a truncated stream)
            if (p + pending > n) { err = RCZ_E_MALFORMED; break; }
            for (unsigned k = 0; k < pending; ++k) code = (code << 8) + bi.next();
            p += pending;
            const unsigned range = (hai - low) / m.total;  // ari/mod.rs:153-159 query()
            if (range == 0) { err = RCZ_E_MALFORMED; break; }
            const unsigned offset = (code - low) / range;
            if (offset >= m.total) { err = RCZ_E_MALFORMED; break; }   // table.rs:106 assert!
            unsigned v, lo, hi, bytes, nout;
            m.find(offset, v, lo, hi);
            range_step(low, hai, m.total, lo, hi, bytes, nout);        // ari/mod.rs:199: re-run the step to learn the shift
            pending = nout;
            if (v == 256u) break;                                      // table.rs:263-266 terminator
            m.update(v);
            if (bo.o >= bo.cap) { err = RCZ_E_OUTPUT_FULL; break; }
            bo.put(v);
        }
        bo.flush();
        out_len[sidx] = bo.o; status[sidx] = err;
        if (in_used) in_used[sidx] = p + (err ? 0 : pending);          // incl. the bytes only finish() consumes (ari/mod.rs:289-292)
    }
}

constexpr size_t SMEM = (size_t)NT * ROW * 4;

}  // namespace arik

// all descriptor / result arrays are DEVICE pointers (composed calls: pipeline.cu); only enqueues
int rcz_ari_launch(rcz_ctx* c, bool decode, const uint8_t* din, const uint64_t* d_in_off, const uint64_t* d_in_len, uint8_t* dout,
                   const uint64_t* d_out_off, const uint64_t* d_out_cap, uint64_t* d_out_len, uint64_t* d_in_used, int32_t* d_status, size_t n) {
    if (n == 0) return RCZ_OK;
    RCZ_CK(c, RCZ_KERNEL_SMEM_OPTIN(arik::ari_decode_kernel, arik::SMEM));
    RCZ_CK(c, RCZ_KERNEL_SMEM_OPTIN(arik::ari_encode_kernel, arik::SMEM));
    const unsigned grid = (unsigned)std::min<size_t>((n + arik::NT - 1) / arik::NT, (size_t)c->sm_count * 3);
    if (decode)
        RCZ_KLAUNCH(c, arik::ari_decode_kernel, grid, arik::NT, arik::SMEM, din, d_in_off, d_in_len, dout, d_out_off, d_out_cap, d_out_len, d_in_used, d_status, (unsigned)n);
    else
        RCZ_KLAUNCH(c, arik::ari_encode_kernel, grid, arik::NT, arik::SMEM, din, d_in_off, d_in_len, dout, d_out_off, d_out_cap, d_out_len, d_status, (unsigned)n);
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
