// lz4_encode.cu — LZ4 block ENCODE, bit-exact with the reference's greedy single-probe compressor.
//
// Replaces /root/reference/src/lz4.rs:226-310 `BlockEncoder::encode` (+ `write_literals` :192-224, `seq_at` :185-190) behind
// `lz4::encode_block` (:616-627).  What the reference computes is defined by its probe sequence: every probed position p looks up
// table[hash(seq(p))] (2^17 entries, the most recent PROBED position with that hash, biased by UNINITHASH), overwrites it with p, and
// is a hit when the candidate lies < 64 KiB back and its four bytes are equal; misses advance by `step`, which grows with the
// distance from the last emitted sequence (:255-262); a hit found while step > 1 restores the table entry, rewinds to just after the
// previous probe and retries with step 1 (:264-269); a hit with step 1 emits literals + match, extended byte by byte up to len - 5.
// Which positions get probed depends on every earlier hit, so the chain is serial per block — but everything between two hits is
// predictable: while no hit occurs the next 32 probe positions (and their step / limit) follow from the state alone.
//
// Shape on the GPU: one warp per block.  The warp evaluates a WINDOW of the next 32 probes speculatively, one per lane (4-byte
// sequence, hash, table look-up from L2, candidate check), resolves equal hashes inside the window with a warp match (a lane's
// candidate is the nearest earlier lane with its hash, else the table), finds the first hit with a ballot, commits exactly the
// table writes the reference would have made up to that probe (one writer per hash), and either rewinds or emits the sequence with
// warp-wide literal copy and match extension.  The result is byte-identical to `encode_block`.
#include "rcz_internal.h"
#include <algorithm>

namespace lz4e {

constexpr unsigned HASH_LOG = 17, TABLE = 1u << HASH_LOG;       // lz4.rs:48-49
constexpr unsigned UNINIT = 0x88888888u;                        // lz4.rs:52
constexpr unsigned INCOMPRESSIBLE = 128;                        // lz4.rs:51
constexpr unsigned MAX_INPUT_SIZE = 0x7e000000u;                     // lz4.rs:53

// four bytes at an arbitrary address, little endian (seq_at, lz4.rs:185-190): two aligned words + funnel shift
__device__ __forceinline__ unsigned ld32u(const uint8_t* p) {
    const uintptr_t a = (uintptr_t)p, a0 = a & ~(uintptr_t)3;
    const unsigned sh = (unsigned)(a & 3) * 8u;
    const unsigned w0 = __ldg(reinterpret_cast<const unsigned*>(a0));
    if (sh == 0) return w0;
    const unsigned w1 = __ldg(reinterpret_cast<const unsigned*>(a0 + 4));
    return __funnelshift_r(w0, w1, sh);
}

// one miss: lz4.rs:256-261
__device__ __forceinline__ void miss(unsigned& pos, unsigned& step, unsigned& limit, unsigned anchor) {
    if (pos - anchor > limit) { limit <<= 1; step += 1 + (step >> 2); }
    pos += step;
}

// 255-continued length bytes of write_literals / the match length (lz4.rs:208-216, 293-304): rem = value - 15
__device__ __forceinline__ unsigned put_ext(uint8_t* out, unsigned dest, unsigned rem, unsigned lane) {
    const unsigned n255 = rem / 255u;
    for (unsigned t = lane; t < n255; t += 32) out[dest + t] = 255;
    if (lane == 0) out[dest + n255] = (uint8_t)(rem - n255 * 255u);
    return dest + n255 + 1;
}

__global__ void __launch_bounds__(32)
lz4_encode_kernel(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
                  uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ out_cap,
                  uint64_t* __restrict__ out_len, int32_t* __restrict__ status, unsigned* __restrict__ tables, unsigned nblocks) {
    const unsigned lane = threadIdx.x & 31;
    for (unsigned b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const unsigned long long n64 = in_len[b];
        if (n64 > MAX_INPUT_SIZE) { if (lane == 0) { out_len[b] = 0; status[b] = RCZ_OK; } continue; }       // lz4.rs:229-230: returns 0
        const unsigned n = (unsigned)n64;
        const unsigned long long bound = (unsigned long long)n + n / 255u + 16 + 4;                     // compression_bound, lz4.rs:175-181
        if (out_cap[b] < bound) { if (lane == 0) { out_len[b] = 0; status[b] = RCZ_E_OUTPUT_FULL; } continue; }
        const uint8_t* in = in_base + in_off[b];
        uint8_t* out = out_base + out_off[b];
        unsigned* table = tables + (size_t)blockIdx.x * TABLE;                                           // zeroed by the host side (lz4.rs:620)
        if (b != blockIdx.x) {                                                                           // a second block on this warp: clear again
            for (unsigned i = lane; i < TABLE; i += 32) table[i] = 0;
            __syncwarp();
        }
        unsigned pos = 0, anchor = 0, dest = 0, step = 1, limit = INCOMPRESSIBLE;
        for (;;) {
            // ---- the next 32 probes if none of them hits: position, step and limit of lane j after j misses
            unsigned p = pos, st = step, lm = limit;
            if (!(step == 1 && pos + 31u - anchor <= limit)) { for (unsigned t = 0; t < lane; ++t) miss(p, st, lm, anchor); }
            else p = pos + lane;
            const bool ended = (unsigned long long)p + 12 > n;                                           // lz4.rs:243: the block ends at this probe
            unsigned seq = 0, h = 0, r = 0;
            bool hit = false;
            if (!ended) {
                seq = ld32u(in + p);
                h = (seq * 2654435761u) >> (32 - HASH_LOG);                                             // lz4.rs:251
                r = __ldcg(table + h) + UNINIT;                                                         // lz4.rs:252 (wrapping)
            }
            // probes of this window with the same hash: the candidate of a lane is the nearest earlier one (its table write, lz4.rs:253)
            const unsigned grp = __match_any_sync(RCZ_FULL, ended ? (TABLE + lane) : h);
            const unsigned before = grp & ((1u << lane) - 1u);
            const int src = before ? 31 - __clz((int)before) : 0;
            const unsigned pprev = __shfl_sync(RCZ_FULL, p, src);
            if (!ended && before) r = pprev;
            if (!ended && ((p - r) >> 16) == 0) hit = ld32u(in + r) == seq;                             // lz4.rs:255
            const unsigned ev = __ballot_sync(RCZ_FULL, hit || ended);
            if (ev == 0) {
                // 32 misses: every hash keeps its last probe
                if ((grp >> lane) == 1u) table[h] = p - UNINIT;
                __syncwarp();
                unsigned p2 = p, s2 = st, l2 = lm;
                miss(p2, s2, l2, anchor);
                pos = __shfl_sync(RCZ_FULL, p2, 31); step = __shfl_sync(RCZ_FULL, s2, 31); limit = __shfl_sync(RCZ_FULL, l2, 31);
                continue;
            }
            const unsigned k = (unsigned)__ffs((int)ev) - 1u;
            const unsigned pk = __shfl_sync(RCZ_FULL, p, (int)k), rk = __shfl_sync(RCZ_FULL, r, (int)k);
            const unsigned stk = __shfl_sync(RCZ_FULL, st, (int)k), lmk = __shfl_sync(RCZ_FULL, lm, (int)k);
            const unsigned hitk_end = __shfl_sync(RCZ_FULL, ended ? 1u : 0u, (int)k);                    // the event is the end of the block, not a hit
            // table writes of probes 0..k: one writer per hash = its last probe at or before k; probe k itself leaves its own position,
            // or the restored candidate when it is a hit found while skipping (lz4.rs:264-266)
            if (!hitk_end || lane < k) {
                const unsigned upto = grp & ((2u << k) - 1u);
                if (lane <= k && !ended && (upto >> lane) == 1u) {
                    unsigned v = p - UNINIT;
                    if (lane == k && stk > 1) v = rk - UNINIT;
                    table[h] = v;
                }
            }
            __syncwarp();
            if (hitk_end) {                                                                              // lz4.rs:243-248: trailing literals, done
                const unsigned ln = n - anchor;
                const unsigned code = ln > 14 ? 15u : ln;
                if (lane == 0) out[dest] = (uint8_t)(code << 4);
                ++dest;
                if (code == 15) dest = put_ext(out, dest, ln - 15, lane);
                for (unsigned t = lane; t < ln; t += 32) out[dest + t] = in[anchor + t];
                dest += ln;
                break;
            }
            if (stk > 1) { pos = pk - (stk - 1); step = 1; limit = lmk; continue; }                      // lz4.rs:264-269 rewind
            // ---- emit: literals [anchor, pk), match of 4 + extension at distance pk - rk (lz4.rs:271-306)
            const unsigned ln = pk - anchor, back = pk - rk;
            unsigned mp = pk + 4, mr = rk + 4;
            for (;;) {                                                                                   // lz4.rs:281-284, 32 bytes per step
                const unsigned q = mp + lane;
                const bool same = q < n - 5 && in[q] == in[mr + lane];
                const unsigned neq = ~__ballot_sync(RCZ_FULL, same);
                if (neq) { const unsigned f = (unsigned)__ffs((int)neq) - 1u; mp += f; break; }
                mp += 32; mr += 32;
            }
            const unsigned ml = mp - (pk + 4);
            const unsigned code = ln > 14 ? 15u : ln;
            if (lane == 0) out[dest] = (uint8_t)((code << 4) + (ml > 14 ? 15u : ml));                    // lz4.rs:197-201
            ++dest;
            if (code == 15) dest = put_ext(out, dest, ln - 15, lane);
            for (unsigned t = lane; t < ln; t += 32) out[dest + t] = in[anchor + t];
            dest += ln;
            if (lane == 0) { out[dest] = (uint8_t)back; out[dest + 1] = (uint8_t)(back >> 8); }          // lz4.rs:289-291
            dest += 2;
            if (ml > 14) dest = put_ext(out, dest, ml - 15, lane);
            pos = mp; anchor = mp; step = 1; limit = INCOMPRESSIBLE;
        }
        __syncwarp();
        if (lane == 0) { out_len[b] = dest; status[b] = RCZ_OK; }
    }
}

}  // namespace lz4e

extern "C" int rcz_lz4_encode_blocks(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                                     const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t nblocks,
                                     int mem_kind) {
    using namespace lz4e;
    if (!c || rcz_bad_kind(mem_kind)) return RCZ_E_ARG;
    if (nblocks == 0) return RCZ_OK;
    if (!in_base || !in_off || !in_len || !out_base || !out_off || !out_cap || !out_len || !status || nblocks > 0x7fffffffu) return RCZ_E_ARG;
    if (!rcz_spans_ok(in_off, in_len, nblocks) || !rcz_spans_ok(out_off, out_cap, nblocks)) return RCZ_E_ARG;
    rt_set_device(c->device);
    DescStager ds(c, mem_kind, nblocks);
    ds.add_in(in_off, nblocks * 8); ds.add_in(in_len, nblocks * 8); ds.add_in(out_off, nblocks * 8); ds.add_in(out_cap, nblocks * 8);
    ds.add_out(out_len, nblocks * 8); ds.add_out(status, nblocks * 4);
    int st = ds.upload(); if (st) return st;
    const uint8_t* din = (const uint8_t*)in_base; uint8_t* dout = (uint8_t*)out_base;
    if (mem_kind == RCZ_MEM_HOST) {
        st = stage_span_in(c, WS_IN, in_base, in_off, in_len, nblocks, 1, &din); if (st) return st;
        st = stage_span_out(c, WS_OUT, out_off, out_cap, nblocks, 1, &dout); if (st) return st;
    }
    // one hash table (512 KiB) per resident warp
    const unsigned grid = (unsigned)std::min<size_t>(nblocks, (size_t)c->sm_count * 16);
    void* wt; st = ctx_ws(c, WS_A, (size_t)grid * TABLE * 4 + 256, &wt); if (st) return st;
    RCZ_CK(c, rt_memset(wt, 0, (size_t)grid * TABLE * 4, c->stream));
    st = ctx_timer_begin(c); if (st) return st;
    RCZ_KLAUNCH(c, lz4_encode_kernel, grid, 32, 0, din, ds.in_ptr<uint64_t>(0), ds.in_ptr<uint64_t>(1), dout, ds.in_ptr<uint64_t>(2), ds.in_ptr<uint64_t>(3),
                ds.out_ptr<uint64_t>(0), ds.out_ptr<int32_t>(1), (unsigned*)wt, (unsigned)nblocks);
    st = ctx_timer_end(c); if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) {
        std::vector<uint64_t> lens(nblocks);
        for (size_t i = 0; i < nblocks; ++i) lens[i] = status[i] == RCZ_OK ? out_len[i] : 0;
        st = unstage_span_out(c, out_base, dout, out_off, lens.data(), nblocks, 1); if (st) return st;
    }
    return RCZ_OK;
}
