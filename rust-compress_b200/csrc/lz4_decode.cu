// lz4_decode.cu — K1: LZ4 block decode, one CTA per block with intra-block parallelism.
//
// Replaces the serial loop `BlockDecoder::decode` (/root/reference/src/lz4.rs:64-162: token, literal run,
// u16 offset, match with the offset<4 DECR dance, byte-forward copy) — same bytes out, different shape:
//
//   per batch (one 8 KiB window of the compressed block, staged in shared memory by a TMA bulk copy):
//     P1  every window position speculatively parses "the token that would start here" -> nxt1[]
//     P2  pointer doubling (5 rounds) gives nxt32[]; one thread hops 32 tokens at a time, 32 chains are
//         then filled in parallel -> the true token list (sequence boundaries) of the window
//     P3  per-token field parse + block-wide warp-shuffle prefix scan of (literals+match) lengths
//         -> output offset of every sequence (literal/match boundary resolution)
//   per 16 KiB output tile:
//     P4  sequence heads are scattered into the tile, a per-warp max-scan labels every output byte with
//         its sequence; every byte becomes a literal value, a value gathered from already-final output,
//         or a 15-bit pointer to an earlier byte of the same tile
//     P5  in-tile pointers are resolved by pointer doubling in shared memory (<= log2(tile) rounds,
//         usually 1-2) — this is what makes overlapping / chained matches order-independent
//     P6  the tile is written to HBM with coalesced stores
//
// Offsets < match length (incl. 1..3, the reference's DECR path lz4.rs:100-102) are folded analytically:
// byte k of a match with offset `off` equals byte (k mod off) of the `off` bytes before the match.
// Malformed input (truncated fields, offset 0, offset before block start — the cases on which the
// reference panics) yields RCZ_E_MALFORMED for that block.
#include "rcz_internal.h"
#include <algorithm>

namespace lz4k {

constexpr int NT = 512;             // threads per CTA (2 CTAs / SM)
constexpr int CH = 8192;            // compressed window bytes in shared memory (= 16 * NT)
constexpr int SMAX = 1024;          // sequences per batch
constexpr int HOP = 8;              // tokens per walker hop (3 doubling rounds)
constexpr int NCHAIN = SMAX / HOP;
constexpr int TILE = 16384;         // output tile entries (pointer field is 15 bits)
constexpr int EXT_MAX = 8;          // 0xFF-continuation bytes parsed in-window before deferring to the slow path
constexpr unsigned SINK = CH;
constexpr unsigned NX_EXIT = 0xFFFE, NX_END = 0xFFFD, NX_SLOW = 0xFFFC, NX_BAD = 0xFFFB, NX_CONT = 0xFFFA;
static_assert(CH == 16 * NT, "P1 assigns 16 consecutive window bytes to every thread");
static_assert(NCHAIN <= NT, "one thread per chain");

struct __align__(16) SeqEnt { uint32_t out_start, L, lit_abs, off; };

struct Smem {
    __align__(16) uint8_t chunk[2][CH + 16];   // double-buffered compressed window (TMA destination)
    __align__(16) uint16_t nxt1[CH + 8];
    __align__(16) uint16_t pp[2][CH + 8];      // ping-pong doubling tables; reused as ent[TILE] in the copy phase
    SeqEnt seq[SMAX];
    uint16_t tok[SMAX];
    uint16_t W[NCHAIN + 8];
    unsigned scan[40];
    unsigned nchains, K, term_kind, term_idx, next_idx_last, blk;
    unsigned slow_next, slow_tot;
    int slow_err;
    rcz_mbar bar[2];
};
static_assert(2 * (CH + 8) >= TILE, "ent alias");

struct Tok { uint32_t L, M, off, lit_idx, next_idx, kind; };   // kind: 0 regular, else NX_END / NX_SLOW / NX_BAD

// Speculative parse of the token at window index j. `lim` = valid window bytes, `e` = window index of the
// end of the block's input (may lie far beyond the window).
__device__ __forceinline__ Tok parse_token(const uint8_t* c, uint32_t j, uint32_t lim, uint32_t e) {
    Tok t; t.M = 0; t.off = 0; t.kind = 0; t.next_idx = 0;
    const uint32_t tk = c[j];
    uint32_t L = tk >> 4, p = j + 1;
    t.L = 0; t.lit_idx = p;
    if (L == 15) {                                              // lz4.rs:112-122 length()
        int cnt = 0;
        for (;;) {
            if (p >= e) { t.kind = NX_BAD; return t; }
            if (p >= lim || cnt >= EXT_MAX) { t.kind = NX_SLOW; return t; }
            uint32_t b = c[p++]; L += b; ++cnt;
            if (b != 255) break;
        }
    }
    t.L = L; t.lit_idx = p;
    if (L > e - p) { t.kind = NX_BAD; return t; }               // literal run past the input end
    p += L;
    if (p == e) { t.kind = NX_END; t.next_idx = p; return t; }  // lz4.rs:87: last sequence has literals only
    if (e - p < 2) { t.kind = NX_BAD; return t; }               // truncated offset
    if (p + 2 > lim) { t.kind = NX_SLOW; return t; }
    t.off = (uint32_t)c[p] | ((uint32_t)c[p + 1] << 8);         // lz4.rs:91
    p += 2;
    uint32_t M = tk & 15;
    if (M == 15) {
        int cnt = 0;
        for (;;) {
            if (p >= e) { t.kind = NX_BAD; return t; }
            if (p >= lim || cnt >= EXT_MAX) { t.kind = NX_SLOW; return t; }
            uint32_t b = c[p++]; M += b; ++cnt;
            if (b != 255) break;
        }
    }
    t.M = M + 4;                                                // lz4.rs:99-106: always 4 + len bytes in total
    t.next_idx = p;
    if (p == e) t.kind = NX_END;                                // block ends right after a match: loop at lz4.rs:68 exits
    return t;
}

// next-token code of window position j whose token byte is tk (same result as parse_token, but the common
// case — neither length nibble is 15 — needs no further loads)
__device__ __forceinline__ unsigned next_code(const uint8_t* c, uint32_t tk, uint32_t j, uint32_t lim, uint32_t e) {
    const uint32_t L = tk >> 4, M = tk & 15;
    if (L != 15 && M != 15) {
        const uint32_t p = j + 1 + L;
        if (p > e) return NX_BAD;
        if (p == e) return NX_END;
        if (e - p < 2) return NX_BAD;
        if (p + 2 > lim) return NX_SLOW;
        if (p + 2 == e) return NX_END;
        return p + 2 < lim ? p + 2 : NX_EXIT;
    }
    const Tok t = parse_token(c, j, lim, e);
    return t.kind ? t.kind : (t.next_idx < lim ? t.next_idx : NX_EXIT);
}

// warp-parallel scan of a 0xFF-continued length starting at absolute position p (slow path only)
__device__ __forceinline__ bool ext_scan(const uint8_t* in, unsigned n, unsigned& p, unsigned long long& acc) {
    const unsigned lane = threadIdx.x & 31;
    for (;;) {
        unsigned idx = p + lane;
        unsigned b = idx < n ? in[idx] : 0u;
        unsigned m = __ballot_sync(RCZ_FULL, b != 255u);
        if (m == 0) { acc += 255ull * 32; p += 32; if (acc > 0xffffffffull) return false; continue; }
        unsigned f = (unsigned)__ffs((int)m) - 1;
        unsigned bf = __shfl_sync(RCZ_FULL, b, (int)f);
        if (p + f >= n) return false;                           // ran off the end of the input: bump() panic
        acc += 255ull * f + bf;
        p += f + 1;
        return true;
    }
}

__device__ __forceinline__ void unpack8(const uint4 q, unsigned* v) {
    v[0] = q.x & 0xffff; v[1] = q.x >> 16; v[2] = q.y & 0xffff; v[3] = q.y >> 16;
    v[4] = q.z & 0xffff; v[5] = q.z >> 16; v[6] = q.w & 0xffff; v[7] = q.w >> 16;
}
__device__ __forceinline__ uint4 pack8(const unsigned* v) {
    uint4 q;
    q.x = v[0] | (v[1] << 16); q.y = v[2] | (v[3] << 16); q.z = v[4] | (v[5] << 16); q.w = v[6] | (v[7] << 16);
    return q;
}

// issue the TMA bulk copy of the window that starts at compressed offset cpos into buffer `buf`
__device__ __forceinline__ void issue_window(Smem& sm, int buf, const uint8_t* in, unsigned n, unsigned cpos) {
    const uintptr_t a = (uintptr_t)(in + cpos);
    const unsigned lead = (unsigned)(a & 15);
    const unsigned avail = lead + (n - cpos);
    const unsigned lim = avail < (unsigned)CH ? avail : (unsigned)CH;
    const unsigned load_bytes = (lim + 15u) & ~15u;
    fence_proxy_async_smem();
    mbar_expect_tx(&sm.bar[buf], load_bytes);
    tma_load_1d(sm.chunk[buf], (const uint8_t*)(a - lead), load_bytes, &sm.bar[buf]);
}

__global__ void __launch_bounds__(NT, 2)
lz4_decode_kernel(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off,
                  const uint64_t* __restrict__ in_len, uint8_t* __restrict__ out_base,
                  const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ out_cap,
                  uint64_t* __restrict__ out_len, int32_t* __restrict__ status, unsigned nblocks,
                  unsigned* __restrict__ ticket) {
    RCZ_DYN_SMEM(raw);
    Smem& sm = *reinterpret_cast<Smem*>(raw);
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint16_t* const ent = &sm.pp[0][0];
    volatile uint16_t* const vent = ent;
    unsigned phase[2] = {0, 0};

    if (tid == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); mbar_fence_init(); }
    __syncthreads();

    for (;;) {
        if (tid == 0) sm.blk = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned b = sm.blk;
        if (b >= nblocks) break;

        const uint8_t* in = in_base + in_off[b];
        const unsigned n = (unsigned)in_len[b];
        uint8_t* out = out_base + out_off[b];
        const unsigned long long cap64 = out_cap[b];
        const unsigned cap = cap64 > 0x7fffffffull ? 0x7fffffffu : (unsigned)cap64;
        unsigned cpos = 0, opos = 0;
        int err = 0, buf = 0;
        bool loaded = false;                                     // window for `cpos` already in flight in chunk[buf]?

        while (cpos < n && !err) {
            // ---------------- P0: compressed window in shared memory (TMA bulk copy, 16-byte aligned source) ------
            const unsigned lead = (unsigned)((uintptr_t)(in + cpos) & 15);
            const unsigned avail = lead + (n - cpos);            // window index of the input end
            const unsigned lim = avail < (unsigned)CH ? avail : (unsigned)CH;
            const uint8_t* chunk = sm.chunk[buf];
            __syncthreads();                                     // everyone is done with ent / seq of the previous batch
            if (!loaded && tid == 0) issue_window(sm, buf, in, n, cpos);
            mbar_wait(&sm.bar[buf], phase[buf]);
            phase[buf] ^= 1;
            loaded = false;

            // ---------------- P1: speculative next-token table (16 consecutive positions per thread) ---------------
            {
                const unsigned j0 = tid * 16;
                const uint4 q = *reinterpret_cast<const uint4*>(chunk + j0);
                const unsigned w4[4] = {q.x, q.y, q.z, q.w};
                unsigned nx[16];
                // interior threads: the next-token index is arithmetic on the token byte plus at most one 0xFF-continuation
                // byte per length (one more shared-memory byte load each); anything longer, and every position near the window
                // or input edge, is decided exactly by next_code()
                const bool interior = (j0 >= lead) && (j0 + 16 + 20 < lim);
                const unsigned nb = chunk[j0 + 16];                  // first byte of the next thread's group (16 B of slack)
                unsigned slowmask = 0;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const unsigned j = j0 + k;
                    const unsigned tk = (w4[k >> 2] >> ((k & 3) * 8)) & 0xffu;
                    const unsigned e1 = k == 15 ? nb : ((w4[(k + 1) >> 2] >> (((k + 1) & 3) * 8)) & 0xffu);
                    const unsigned L = tk >> 4;
                    const bool l15 = L == 15u, m15 = (tk & 15u) == 15u;
                    const unsigned p = j + 1 + L + (l15 ? 1u + e1 : 0u);      // position of the 2-byte offset
                    bool slow = l15 && e1 == 255u;
                    unsigned nxt = p + 2;
                    if (m15) {
                        const unsigned ei = p + 2 < (unsigned)CH + 15u ? p + 2 : (unsigned)CH + 15u;
                        slow |= chunk[ei] == 255u;
                        nxt += 1;
                    }
                    slow |= l15 && nxt >= lim;                                // a long literal run may reach the window / input edge
                    nx[k] = nxt;
                    if (slow) slowmask |= 1u << k;
                }
                if (!interior) slowmask = 0xffffu;
                *reinterpret_cast<uint4*>(&sm.nxt1[j0]) = pack8(nx);
                *reinterpret_cast<uint4*>(&sm.nxt1[j0 + 8]) = pack8(nx + 8);
#pragma unroll
                for (int k = 0; k < 16; ++k) nx[k] = nx[k] < (unsigned)CH ? nx[k] : SINK;
                *reinterpret_cast<uint4*>(&sm.pp[0][j0]) = pack8(nx);
                *reinterpret_cast<uint4*>(&sm.pp[0][j0 + 8]) = pack8(nx + 8);
#pragma unroll 1
                while (slowmask) {                                   // window edges, lengths with two or more continuation bytes
                    const int k = __ffs((int)slowmask) - 1;
                    slowmask &= slowmask - 1;
                    const unsigned j = j0 + k;
                    const unsigned v = (j >= lead && j < lim) ? next_code(chunk, chunk[j], j, lim, avail) : NX_BAD;
                    sm.nxt1[j] = (uint16_t)v;
                    sm.pp[0][j] = (uint16_t)(v < (unsigned)CH ? v : SINK);
                }
                if (tid == 0) { sm.nxt1[CH] = (uint16_t)NX_BAD; sm.pp[0][CH] = (uint16_t)SINK; sm.pp[1][CH] = (uint16_t)SINK; }
            }
            __syncthreads();
            // ---------------- P2: pointer doubling nxt1 -> nxt8, hop, fill in ----------------------------------------
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
                const uint16_t* s = sm.pp[r & 1];
                uint16_t* d = sm.pp[(r & 1) ^ 1];
                const unsigned j0 = tid * 16;
                unsigned v[16];
                unpack8(*reinterpret_cast<const uint4*>(&s[j0]), v);
                unpack8(*reinterpret_cast<const uint4*>(&s[j0 + 8]), v + 8);
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = s[v[k]];
                *reinterpret_cast<uint4*>(&d[j0]) = pack8(v);
                *reinterpret_cast<uint4*>(&d[j0 + 8]) = pack8(v + 8);
                __syncthreads();
            }
            if (tid == 0) {
                const uint16_t* f8 = sm.pp[1];
                unsigned nch = 1, t = lead;
                sm.W[0] = (uint16_t)t;
                while (nch < (unsigned)NCHAIN) {
                    t = f8[t];
                    if (t == SINK) break;
                    sm.W[nch++] = (uint16_t)t;
                }
                sm.nchains = nch;
            }
            __syncthreads();
            {
                const unsigned nch = sm.nchains;
                if (tid < nch) {
                    unsigned t = sm.W[tid], cnt = 0;
                    const unsigned base = tid * HOP;
                    const bool last = (tid == nch - 1);
                    for (int i = 0; i < HOP; ++i) {
                        const unsigned nx = sm.nxt1[t];
                        if (nx == NX_SLOW || nx == NX_BAD) { sm.term_kind = nx; sm.term_idx = t; break; }
                        sm.tok[base + i] = (uint16_t)t; ++cnt;
                        if (nx >= (unsigned)CH) { sm.term_kind = nx; sm.term_idx = t; break; }
                        t = nx;
                        if (i == HOP - 1 && last) { sm.term_kind = NX_CONT; sm.term_idx = t; }
                    }
                    if (last) sm.K = base + cnt;
                }
            }
            __syncthreads();
            unsigned K = sm.K;
            const unsigned term_kind = sm.term_kind, term_idx = sm.term_idx;
            unsigned batch_end, next_cpos;

            if (K == 0) {
                if (term_kind == NX_BAD) { err = RCZ_E_MALFORMED; break; }
                // ------------ slow path: the first token of the window has fields that do not fit the window ---------
                if (warp == 0) {
                    int serr = 0;
                    unsigned p = cpos;
                    const unsigned tk = in[p]; ++p;
                    unsigned long long L = tk >> 4, M = 0;
                    unsigned off = 0, lit_abs = 0;
                    if (L == 15 && !ext_scan(in, n, p, L)) serr = RCZ_E_MALFORMED;
                    if (!serr) {
                        lit_abs = p;
                        if (L > (unsigned long long)(n - p)) serr = RCZ_E_MALFORMED;
                        else p += (unsigned)L;
                    }
                    if (!serr && p < n) {
                        if (n - p < 2) serr = RCZ_E_MALFORMED;
                        else {
                            off = (unsigned)in[p] | ((unsigned)in[p + 1] << 8); p += 2;
                            M = tk & 15;
                            if (M == 15 && !ext_scan(in, n, p, M)) serr = RCZ_E_MALFORMED;
                            M += 4;
                        }
                    }
                    if (!serr && M > 0 && (off == 0 || (unsigned long long)off > (unsigned long long)opos + L)) serr = RCZ_E_MALFORMED;
                    if (!serr && L + M > (unsigned long long)(cap - opos)) serr = RCZ_E_OUTPUT_FULL;
                    if (lane == 0) {
                        sm.slow_err = serr;
                        if (!serr) {
                            SeqEnt s; s.out_start = opos; s.L = (unsigned)L; s.lit_abs = lit_abs; s.off = off;
                            sm.seq[0] = s;
                            sm.slow_next = p; sm.slow_tot = (unsigned)(L + M);
                        }
                    }
                }
                __syncthreads();
                if (sm.slow_err) { err = sm.slow_err; break; }
                K = 1;
                batch_end = opos + sm.slow_tot;
                next_cpos = sm.slow_next;
            } else {
                // ------------ P3: fields of every real token + prefix scan of sequence lengths -------------------------
                Tok tk[2]; unsigned len[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const unsigned i = 2 * tid + q;
                    len[q] = 0;
                    if (i < K) { tk[q] = parse_token(chunk, sm.tok[i], lim, avail); len[q] = tk[q].L + tk[q].M; }
                }
                unsigned tot;
                const unsigned ex = block_excl_scan_add<NT>(len[0] + len[1], sm.scan, &tot);
                int bad = 0;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const unsigned i = 2 * tid + q;
                    if (i < K) {
                        SeqEnt s;
                        s.out_start = opos + ex + (q ? len[0] : 0u);
                        s.L = tk[q].L;
                        s.lit_abs = cpos + (tk[q].lit_idx - lead);
                        s.off = tk[q].off;
                        if (tk[q].M > 0 && (s.off == 0 || s.off > s.out_start + s.L)) bad = 1;   // lz4.rs:93 underflow / App. B #9
                        sm.seq[i] = s;
                        if (i == K - 1) sm.next_idx_last = tk[q].next_idx;
                    }
                }
                bad = __syncthreads_or(bad);
                if (bad) { err = RCZ_E_MALFORMED; break; }
                if (tot > cap - opos) { err = RCZ_E_OUTPUT_FULL; break; }
                batch_end = opos + tot;
                if (term_kind == NX_END) next_cpos = n;
                else if (term_kind == NX_EXIT) next_cpos = cpos + (sm.next_idx_last - lead);
                else next_cpos = cpos + (term_idx - lead);           // NX_CONT / NX_SLOW / NX_BAD: resume at that token
            }
            // the next window is known: start its TMA now, it lands while this batch's output is materialised
            if (next_cpos < n) {
                if (tid == 0) issue_window(sm, buf ^ 1, in, n, next_cpos);
                loaded = true;
            }

            // ---------------- P4-P6: materialise the batch's output, one tile at a time -----------------------------
            // Entry i of a tile corresponds to global address (out + tbase + i); tbase is chosen so that entry 0 is
            // 16-byte aligned in HBM, so full vectors leave the SM as 16-byte stores.
            for (unsigned t0 = opos; t0 < batch_end;) {
                const unsigned h = (unsigned)((uintptr_t)(out + t0) & 15);
                const unsigned room = (unsigned)TILE - h;
                const unsigned tlen = (batch_end - t0) < room ? (batch_end - t0) : room;
                const unsigned t1 = t0 + tlen;
                const unsigned tend = h + tlen;                          // entries [h, tend) are live
                const unsigned tend8 = (tend + 7u) & ~7u;
                const unsigned tbase = t0 - h;                           // may wrap below zero; only used as pos - tbase, pos >= t0
                // first / one-past-last sequence overlapping [t0, t1)
                unsigned ja, jb;
                {
                    unsigned lo = 0, hi = K;                              // ja = (#seq with out_start <= t0) - 1
                    while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (sm.seq[mid].out_start <= t0) lo = mid + 1; else hi = mid; }
                    ja = lo - 1;
                    lo = ja; hi = K;                                      // jb = #seq with out_start < t1
                    while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (sm.seq[mid].out_start < t1) lo = mid + 1; else hi = mid; }
                    jb = lo;
                }
                if (tid < 16) { if (tid < h) ent[tid] = 0; if (tend + tid < tend8 + 8) ent[tend + tid] = 0; }   // dead head / tail entries
                // ---- P4: one warp per sequence: literal bytes, bytes gathered from final output, or in-tile pointers
                int anyptr = 0;
                for (unsigned j = ja + warp; j < jb; j += NT / 32) {
                    const SeqEnt s = sm.seq[j];
                    const unsigned nstart = (j + 1 < K) ? sm.seq[j + 1].out_start : batch_end;
                    const unsigned ms = s.out_start + s.L;               // match part starts here
                    {   // literal run (lz4.rs:75-85)
                        const unsigned lo = s.out_start > t0 ? s.out_start : t0, hi = ms < t1 ? ms : t1;
#pragma unroll 1
                        for (unsigned pos = lo + lane; pos < hi; pos += 32) {
                            const unsigned la = s.lit_abs + (pos - s.out_start);
                            const unsigned wi = (la - cpos) + lead;
                            ent[pos - tbase] = wi < lim ? (uint16_t)chunk[wi] : (uint16_t)__ldg(in + la);
                        }
                    }
                    {   // match (lz4.rs:96-107, cp lz4.rs:131-140): byte k equals byte (k mod off) of the `off` bytes before it
                        const unsigned lo = ms > t0 ? ms : t0, hi = nstart < t1 ? nstart : t1;
                        const unsigned off = s.off;
#pragma unroll 1
                        for (unsigned pos = lo + lane; pos < hi; pos += 32) {
                            const unsigned k = pos - ms;
                            const unsigned srcp = k < off ? pos - off : (ms - off) + (k % off);
                            unsigned e;
                            if (srcp < t0) e = out[srcp];                // already final in HBM/L2
                            else { e = 0x8000u | (srcp - tbase); anyptr = 1; }
                            ent[pos - tbase] = (uint16_t)e;
                        }
                    }
                }
                int pend = __syncthreads_or(anyptr);
                while (pend) {                                           // P5: pointer doubling inside the tile
                    int p2 = 0;
                    for (unsigned i8 = tid * 8; i8 < tend8; i8 += NT * 8) {
                        const uint4 q = lds128_volatile(&ent[i8]);
                        if (((q.x | q.y | q.z | q.w) & 0x80008000u) == 0) continue;
                        unsigned e[8], f[8];
                        unpack8(q, e);
#pragma unroll
                        for (int u = 0; u < 8; ++u) f[u] = (e[u] & 0x8000u) ? (unsigned)vent[e[u] & 0x7fffu] : e[u];
#pragma unroll
                        for (int u = 0; u < 8; ++u) p2 |= (int)(f[u] >> 15);
                        sts128_volatile(&ent[i8], pack8(f));
                    }
                    pend = __syncthreads_or(p2);
                }
                // ---- P6: 16 entries -> one aligned 16-byte store (byte stores only for the ragged head / tail vector)
                for (unsigned vi = tid; vi * 16 < tend; vi += NT) {
                    const unsigned i = vi * 16;
                    unsigned v[16];
                    unpack8(*reinterpret_cast<const uint4*>(&ent[i]), v);
                    unpack8(*reinterpret_cast<const uint4*>(&ent[i + 8]), v + 8);
                    uint8_t* dst = out + t0 - h + i;
                    if (i >= h && i + 16 <= tend) {
                        uint4 q;
                        q.x = (v[0] & 255) | ((v[1] & 255) << 8) | ((v[2] & 255) << 16) | (v[3] << 24);
                        q.y = (v[4] & 255) | ((v[5] & 255) << 8) | ((v[6] & 255) << 16) | (v[7] << 24);
                        q.z = (v[8] & 255) | ((v[9] & 255) << 8) | ((v[10] & 255) << 16) | (v[11] << 24);
                        q.w = (v[12] & 255) | ((v[13] & 255) << 8) | ((v[14] & 255) << 16) | (v[15] << 24);
                        *reinterpret_cast<uint4*>(dst) = q;
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; ++k) if (i + k >= h && i + k < tend) dst[k] = (uint8_t)v[k];
                    }
                }
                __syncthreads();
                t0 += tlen;
            }
            opos = batch_end;
            cpos = next_cpos;
            buf ^= 1;
        }
        __syncthreads();
        if (loaded) { mbar_wait(&sm.bar[buf], phase[buf]); phase[buf] ^= 1; }   // drain a window issued before an error exit
        if (tid == 0) { out_len[b] = opos; status[b] = err; }
    }
}

}  // namespace lz4k

static int lz4_launch(rcz_ctx* c, const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len, uint8_t* out,
                      const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t nblocks) {
    void* tk;
    int st = ctx_ws(c, WS_A, 256, &tk); if (st) return st;
    RCZ_CK(c, rt_memset(tk, 0, 256, c->stream));
    const size_t smem = sizeof(lz4k::Smem);
    RCZ_CK(c, RCZ_KERNEL_SMEM_OPTIN(lz4k::lz4_decode_kernel, smem));
    size_t grid = nblocks < (size_t)2 * c->sm_count ? nblocks : (size_t)2 * c->sm_count;
    st = ctx_timer_begin(c); if (st) return st;
    RCZ_KLAUNCH(c, lz4k::lz4_decode_kernel, (unsigned)grid, lz4k::NT, smem, in, in_off, in_len, out, out_off, out_cap, out_len, status,
                (unsigned)nblocks, (unsigned*)tk);
    return ctx_timer_end(c);
}

// Host-resident buffers, pipelined: the batch is cut into chunks of consecutive blocks; chunk k's compressed bytes go up
// (stream 0) while chunk k-1 decodes (stream 1) and chunk k-2's output comes down (stream 2), so the call costs about
// max(H2D, kernel, D2H) instead of their sum.  Host buffers should be pinned (rcz_host_alloc) for the copies to overlap.
static int lz4_host_pipelined(rcz_ctx* c, DescStager& ds, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                              const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t nblocks) {
    const uint64_t CHUNK_BYTES = getenv("RCZ_LZ4_CHUNK_BYTES") ? strtoull(getenv("RCZ_LZ4_CHUNK_BYTES"), nullptr, 10) : (64ull << 20);   // decoded bytes per chunk
    std::vector<size_t> cut{0};
    uint64_t acc = 0;
    for (size_t i = 0; i < nblocks; ++i) { acc += out_cap[i]; if (acc >= CHUNK_BYTES && i + 1 < nblocks) { cut.push_back(i + 1); acc = 0; } }
    cut.push_back(nblocks);
    const size_t nchunks = cut.size() - 1;
    int st = ctx_aux_streams(c); if (st) return st;
    st = ctx_events(c, 3 * nchunks); if (st) return st;
    // device arenas laid out like the host arenas (same offsets), as in the single-shot path
    const uint8_t* din; uint8_t* dout;
    {   // stage_span_in without the copy: compute the span, allocate, copy chunk by chunk below
        uint64_t lo = UINT64_MAX, hi = 0;
        for (size_t i = 0; i < nblocks; ++i) { if (!in_len[i]) continue; lo = std::min(lo, in_off[i]); hi = std::max(hi, in_off[i] + in_len[i]); }
        if (lo == UINT64_MAX) { lo = 0; hi = 0; }
        const uint64_t lo_al = lo & ~(uint64_t)255;
        void* d; st = ctx_ws(c, WS_IN, (size_t)(hi - lo_al) + 512, &d); if (st) return st;
        din = (const uint8_t*)d - lo_al;
    }
    st = stage_span_out(c, WS_OUT, out_off, out_cap, nblocks, 1, &dout); if (st) return st;
    void* tk;
    st = ctx_ws(c, WS_A, 4 * nchunks + 256, &tk); if (st) return st;
    RCZ_CK(c, rt_memset(tk, 0, 4 * nchunks + 256, c->stream));
    const size_t smem = sizeof(lz4k::Smem);
    RCZ_CK(c, RCZ_KERNEL_SMEM_OPTIN(lz4k::lz4_decode_kernel, smem));
    uint64_t* d_len = ds.out_ptr<uint64_t>(0); int32_t* d_st = ds.out_ptr<int32_t>(1);
    for (size_t k = 0; k < nchunks; ++k) {
        const size_t b0 = cut[k], nb = cut[k + 1] - cut[k];
        uint64_t lo = UINT64_MAX, hi = 0;
        for (size_t i = b0; i < b0 + nb; ++i) { if (!in_len[i]) continue; lo = std::min(lo, in_off[i]); hi = std::max(hi, in_off[i] + in_len[i]); }
        if (lo != UINT64_MAX) RCZ_CK(c, rt_h2d((uint8_t*)din + lo, (const uint8_t*)in_base + lo, (size_t)(hi - lo), c->stream));
        RCZ_CK(c, rt_event_record(c->events[3 * k], c->stream));
        const rt_stream_t ks = c->aux[1 + (k & 7)];                // chunk kernels run concurrently: one chunk alone cannot fill the GPU
        RCZ_CK(c, rt_stream_wait_event(ks, c->events[3 * k]));
        const size_t grid = nb < (size_t)2 * c->sm_count ? nb : (size_t)2 * c->sm_count;
        RCZ_LAUNCH(lz4k::lz4_decode_kernel, (unsigned)grid, lz4k::NT, smem, ks, din, ds.in_ptr<uint64_t>(0) + b0, ds.in_ptr<uint64_t>(1) + b0, dout,
                   ds.in_ptr<uint64_t>(2) + b0, ds.in_ptr<uint64_t>(3) + b0, d_len + b0, d_st + b0, (unsigned)nb, (unsigned*)tk + k);
        c->launches++;
        RCZ_CK(c, rt_last_error());
        RCZ_CK(c, rt_event_record(c->events[3 * k + 1], ks));
    }
    // results and output come down on one D2H stream in chunk order; the (tiny) result copy targets a pinned scratch so that
    // it does not block the host, and the host only waits for it right before it sizes the chunk's output copies
    void* pin; st = ctx_pinned(c, nblocks * 12 + 64, &pin); if (st) return st;
    uint64_t* p_len = (uint64_t*)pin; int32_t* p_st = (int32_t*)((uint8_t*)pin + nblocks * 8);
    const rt_stream_t dsm = c->aux[0];
    size_t issued = 0;                                          // result copies issued so far
    auto issue_results = [&](size_t k) -> int {
        const size_t b0 = cut[k], nb = cut[k + 1] - cut[k];
        RCZ_CK(c, rt_stream_wait_event(dsm, c->events[3 * k + 1]));
        RCZ_CK(c, rt_d2h(p_len + b0, d_len + b0, nb * 8, dsm));
        RCZ_CK(c, rt_d2h(p_st + b0, d_st + b0, nb * 4, dsm));
        RCZ_CK(c, rt_event_record(c->events[3 * k + 2], dsm));
        return RCZ_OK;
    };
    for (size_t k = 0; k < nchunks; ++k) {
        while (issued < nchunks && issued <= k + 1) { st = issue_results(issued++); if (st) return st; }   // one chunk ahead
        RCZ_CK(c, rt_event_sync(c->events[3 * k + 2]));
        size_t i = cut[k];
        const size_t e = cut[k + 1];
        for (size_t q = i; q < e; ++q) { out_len[q] = p_len[q]; status[q] = p_st[q]; }
        while (i < e) {
            if (out_len[i] == 0) { ++i; continue; }
            const uint64_t s = out_off[i]; uint64_t t = s + out_len[i];
            size_t j = i + 1;
            while (j < e && (out_len[j] == 0 || out_off[j] == t)) { t += out_len[j]; ++j; }
            RCZ_CK(c, rt_d2h((uint8_t*)out_base + s, dout + s, (size_t)(t - s), dsm));
            i = j;
        }
    }
    RCZ_CK(c, rt_stream_sync(dsm));
    c->ev_valid = false;
    return RCZ_OK;
}

extern "C" int rcz_lz4_decode_blocks(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                                     void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                                     uint64_t* out_len, int32_t* status, size_t nblocks, int mem_kind) {
    if (!c || rcz_bad_kind(mem_kind)) return RCZ_E_ARG;
    if (nblocks == 0) return RCZ_OK;
    if (!in_base || !in_off || !in_len || !out_base || !out_off || !out_cap || !out_len || !status) return RCZ_E_ARG;
    if (nblocks > 0x7fffffffu) return RCZ_E_ARG;
    rt_set_device(c->device);
    for (size_t i = 0; i < nblocks; ++i) if (in_len[i] >= (1ull << 31)) return RCZ_E_ARG;
    DescStager ds(c, mem_kind, nblocks);
    ds.add_in(in_off, nblocks * 8); ds.add_in(in_len, nblocks * 8); ds.add_in(out_off, nblocks * 8); ds.add_in(out_cap, nblocks * 8);
    ds.add_out(out_len, nblocks * 8); ds.add_out(status, nblocks * 4);
    int st = ds.upload(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST && nblocks >= 8)
        return lz4_host_pipelined(c, ds, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, nblocks);
    const uint8_t* din = (const uint8_t*)in_base; uint8_t* dout = (uint8_t*)out_base;
    if (mem_kind == RCZ_MEM_HOST) {
        st = stage_span_in(c, WS_IN, in_base, in_off, in_len, nblocks, 1, &din); if (st) return st;
        st = stage_span_out(c, WS_OUT, out_off, out_cap, nblocks, 1, &dout); if (st) return st;
    }
    st = lz4_launch(c, din, ds.in_ptr<uint64_t>(0), ds.in_ptr<uint64_t>(1), dout, ds.in_ptr<uint64_t>(2), ds.in_ptr<uint64_t>(3),
                    ds.out_ptr<uint64_t>(0), ds.out_ptr<int32_t>(1), nblocks);
    if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) { st = unstage_span_out(c, out_base, dout, out_off, out_len, nblocks, 1); if (st) return st; }
    return RCZ_OK;
}
