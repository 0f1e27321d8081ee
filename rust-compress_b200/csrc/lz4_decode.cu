// lz4_decode.cu — K1: LZ4 block decode as three window-parallel kernels.
//
// Replaces the serial loop `BlockDecoder::decode` (/root/reference/src/lz4.rs:64-162: token, literal run,
// u16 offset, match with the offset<4 DECR dance, byte-forward copy) — same bytes out, different shape.
// A block is cut into 8 KiB windows of its COMPRESSED bytes; the windows of all blocks are the work units
// (69 K units for 256 x 4 MiB blocks), taken by persistent CTAs in (window index, block) order so that the
// only serial links — one per window — are spread over the whole launch.
//
//   lz4_parse_kernel  (unit = window; no dependence on earlier windows except ONE 64-bit chain word)
//     every window position speculatively parses "the token that would start here" (length extensions are
//     resolved with a 0xFF-run bitmask, so the cost does not depend on the data), then
//       level A  exit of each position from its 32-byte segment: 4 rounds of warp-shuffle pointer doubling
//       level B  exit from the 1 KiB group: reverse sweep over the 32 segments of a warp through shared memory
//     The window that knows its first true token (from the chain word of its predecessor, look-back) hops
//     <= 8 groups to the exit, publishes the successor's chain word, then the warps walk their own group's
//     true tokens, a block-wide warp-shuffle prefix scan of (literals + match) lengths places every sequence
//     in the window's output (literal/match boundary resolution), and the descriptors go to HBM.
//   lz4_scan_kernel   (one warp per block) exclusive scan of the windows' output sizes, block status with
//     the reference's error order (literal overrun, output full, offset underflow lz4.rs:93, ...).
//   lz4_mat_kernel    (one CTA per block, its windows in order; next window's bytes and descriptors prefetched by TMA)
//     the last 64 KiB of the block's output live in a ring in shared memory, so every match source is a
//     shared-memory read.  Output is produced in tiles of <= 8 KiB; one thread per PIECE (the part of a literal
//     run or match that falls into one aligned 16-byte chunk): pieces per sequence -> block prefix scan ->
//     piece -> sequence table.  A piece is fetched with 4-byte loads + funnel shifts (staged compressed window,
//     ring) and OR-ed into its zeroed chunk; per-chunk byte masks order the pieces whose source lies inside
//     the tile (no barrier), and whoever completes a chunk stores its 16 bytes to HBM.
//
// Offsets < match length (incl. 1..3, the reference's DECR path lz4.rs:100-102) are folded analytically:
// byte k of a match with offset `off` equals byte k - m * off for any whole number of periods m that stays inside the
// match's periodic region, so long overlapping matches read ~32 KiB back and form no chunk-to-chunk dependency chain.
// Malformed input (truncated fields, offset 0, offset before block start — the cases on which the
// reference panics) yields RCZ_E_MALFORMED for that block.
#include "rcz_internal.h"
#include <algorithm>

namespace lz4k {

constexpr int W = 8192;                    // compressed bytes per window
constexpr int HALO = 320;                  // bytes after the window that token fields may be read from
constexpr int VIS = W + HALO;
constexpr int NFW = (16 + VIS + 31) / 32 + 2;   // words of the 0xFF bitmask (bit i = byte i of the staged chunk, i.e. window position + lead)
constexpr int CHUNK_BYTES = 16 + VIS + 176;
constexpr int PT = 256;                    // parse: 8 warps, one 1 KiB group each
constexpr int SLOT = W / 3 + 3;            // sequence descriptors per window (a token is >= 3 bytes, except the last of a block)
constexpr unsigned TERM = 0x8000u, NONE = 0xFFFFu;
constexpr int MT = 512;                    // materialise: threads per CTA (2 CTAs / SM)
constexpr int TCAP = 8192;                 // materialise: at most this many output bytes per tile (one chunk per thread)
constexpr int NCH = TCAP / 16;
constexpr int RING = 73760;                // output ring in shared memory: 64 KiB of history + one tile + 16-byte chunk slack
constexpr int PEER_INFLIGHT = 5;           // fused gather: tiles whose bulk stores may still be reading the ring
constexpr int DCAP = 544;                  // sequence descriptors of a unit staged in shared memory (the rest is read from HBM)
static_assert(RING % 16 == 0 && RING > 65535 + TCAP + 16, "a tile must not overwrite history that its matches can reach");
static_assert((PEER_INFLIGHT + 2) * TCAP <= RING - TCAP, "a tile must not overwrite ring bytes that a bulk store may still read");

constexpr unsigned long long CH_VALID = 1ull << 63;
enum { ST_RUN = 0, ST_END = 1, ST_DEAD = 2 };
enum { TK_NONE = 0, TK_NORMAL = 1, TK_END = 2, TK_BAD = 3 };

struct __align__(16) SeqEnt { uint32_t start, L, lit_abs, off; };   // start: offset in the window's output; lit_abs: offset in the block's input
struct __align__(16) WinInfo { unsigned long long total; int minmargin; uint32_t nseq, flags, pad[3]; };
// WinInfo.flags: bit 0 = a match with offset 0; bits 8-9 = the last token is malformed: 1 before its literals are copied, 2 after

static_assert(W == 32 * 32 * (PT / 32), "one 1 KiB group per warp");

// ------------------------------------------------------------------------------------------------------
// spin helpers (cross-CTA look-back).  The emulator runs CTAs one after another in ticket order, so a flag
// that is not yet set there is a bug.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_vol_u64(const unsigned long long* p) { return *(volatile const unsigned long long*)p; }
__device__ __forceinline__ unsigned ld_vol_u32(const unsigned* p) { return *(volatile const unsigned*)p; }
__device__ __forceinline__ void spin_pause(unsigned& n) {
#ifdef RCZ_EMU
    if (++n > 1000u) { fprintf(stderr, "lz4: look-back flag never set (emulation)\n"); abort(); }
    emu::yield();
#else
    (void)n; rcz_backoff(40);
#endif
}

// warp-parallel scan of a 0xFF-continued length starting at block offset p (exact parse only).  acc saturates.
__device__ __forceinline__ bool ext_scan(const uint8_t* in, unsigned n, unsigned& p, unsigned long long& acc) {
    const unsigned lane = threadIdx.x & 31;
    for (;;) {
        const unsigned idx = p + lane;
        const unsigned b = idx < n ? in[idx] : 0u;
        const unsigned m = __ballot_sync(RCZ_FULL, b != 255u);
        if (m == 0) { acc += 255ull * 32; p += 32; if (acc > (1ull << 40)) acc = 1ull << 40; continue; }
        const unsigned f = (unsigned)__ffs((int)m) - 1;
        const unsigned bf = __shfl_sync(RCZ_FULL, b, (int)f);
        if (p + f >= n) return false;                           // ran off the end of the input: bump() panic
        acc += 255ull * f + bf;
        p += f + 1;
        return true;
    }
}

struct TokX { unsigned long long L, M; unsigned off, lit_abs, next, kind, stage; };
// Exact parse of the token at block offset tp (all lanes of a warp call it with the same arguments).  lz4.rs:68-107.
__device__ __forceinline__ TokX tok_exact(const uint8_t* in, unsigned n, unsigned tp) {
    TokX t; t.L = 0; t.M = 0; t.off = 0; t.lit_abs = tp + 1; t.next = n; t.kind = TK_BAD; t.stage = 1;
    unsigned p = tp;
    const unsigned tk = in[p]; ++p;
    unsigned long long L = tk >> 4;
    if (L == 15 && !ext_scan(in, n, p, L)) return t;           // lz4.rs:112-122 length(): bump() past the end
    t.lit_abs = p;
    if (L > (unsigned long long)(n - p)) return t;             // literal run past the input end
    t.L = L; p += (unsigned)L;
    if (p == n) { t.kind = TK_END; return t; }                 // lz4.rs:87: last sequence has literals only
    t.stage = 2;
    if (n - p < 2) return t;                                   // truncated offset
    t.off = (unsigned)in[p] | ((unsigned)in[p + 1] << 8); p += 2;   // lz4.rs:91
    unsigned long long M = tk & 15;
    if (M == 15 && !ext_scan(in, n, p, M)) return t;
    t.M = M + 4;                                               // lz4.rs:99-106: always 4 + len bytes in total
    t.next = p;
    t.kind = p == n ? TK_END : TK_NORMAL;                      // block ends right after a match: loop at lz4.rs:68 exits
    return t;
}

// TMA bulk copy of the visible part of window `base` of a block (16-byte aligned source; `lead` junk bytes in front)
__device__ __forceinline__ void issue_window(uint8_t* dst, rcz_mbar* bar, const uint8_t* in, unsigned n, unsigned base) {
    const uintptr_t a = (uintptr_t)(in + base);
    const unsigned lead = (unsigned)(a & 15);
    const unsigned rem = n - base;
    const unsigned lim = rem < (unsigned)VIS ? rem : (unsigned)VIS;
    const unsigned load_bytes = (lead + lim + 15u) & ~15u;
    fence_proxy_async_smem();
    mbar_expect_tx(bar, load_bytes);
    tma_load_1d(dst, (const uint8_t*)(a - lead), load_bytes, bar);
}

// ======================================================================================================
// parse
// ======================================================================================================
struct ParseSmem {
    __align__(16) uint8_t chunk[CHUNK_BYTES];
    __align__(16) uint16_t exitB[W];
    uint32_t ff[NFW + 2];
    uint32_t lead;
    uint16_t nfull[NFW + 2];
    uint32_t gentry[8], wcnt[8], wlen[8];
    int wmargin[8];
    uint32_t woff0[8];
    unsigned long long t_L, t_M;
    uint32_t t_pos, t_off, t_lit_abs, t_kind, t_stage;
    uint32_t ticket_lo, ticket_hi;
    rcz_mbar bar;
};

// number of consecutive 0xFF bytes starting at window position x (bits beyond the visible limit are 0)
__device__ __forceinline__ unsigned ff_run(const ParseSmem& sm, unsigned xw) {
    const unsigned x = xw + sm.lead;                                   // the mask is indexed by chunk byte
    const unsigned k = x >> 5, j = x & 31;
    unsigned r = (unsigned)__ffs((int)~(sm.ff[k] >> j)) - 1u;   // j == 0 and a full word: ffs(0) - 1 = 0xffffffff
    if (r >= 32u - j) {
        const unsigned f = sm.nfull[k + 1];
        r = (32u - j) + 32u * f + ((unsigned)__ffs((int)~sm.ff[k + 1 + f]) - 1u);
    }
    return r;
}

struct TokF { unsigned L, M, off, lit, next; };
// Token at window position p (p < lim) when all of its fields are visible and it does not end the block; false
// otherwise (such a token always leaves the window or ends the block, and is parsed exactly by the chain warp).
__device__ __noinline__ bool tok_gen(const ParseSmem& sm, const uint8_t* c, unsigned p, unsigned lim, unsigned e_rel, TokF& t) {
    const unsigned tk = c[p];
    unsigned L = tk >> 4, lit = p + 1;
    if (L == 15u) {
        const unsigned x = lit + ff_run(sm, lit);
        if (x >= lim) return false;
        L += 255u * (x - lit) + c[x]; lit = x + 1;
    }
    const unsigned q = lit + L;
    if (q + 2 > lim) return false;
    unsigned M = tk & 15u, nx = q + 2;
    if (M == 15u) {
        const unsigned x = nx + ff_run(sm, nx);
        if (x >= lim) return false;
        M += 255u * (x - nx) + c[x]; nx = x + 1;
    }
    if (nx >= e_rel) return false;
    t.L = L; t.M = M + 4; t.off = (unsigned)c[q] | ((unsigned)c[q + 1] << 8); t.lit = lit; t.next = nx;
    return true;
}

// The same answer without branches for the common case (no 0xFF continuation byte in either length): three byte loads and
// arithmetic.  limq = lim - (lim == e_rel).  Tokens with longer lengths (rare) take tok_gen.  Every load stays inside the
// staged chunk (q + 2 <= p + 274).
__device__ __forceinline__ bool tok_fast(const ParseSmem& sm, const uint8_t* c, unsigned p, unsigned lim, unsigned limq, unsigned e_rel, TokF& t) {
    const unsigned tk = c[p], b1 = c[p + 1];
    const unsigned L4 = tk >> 4, M4 = tk & 15u;
    const bool l15 = L4 == 15u, m15 = M4 == 15u;
    const unsigned L = L4 + (l15 ? b1 : 0u);
    const unsigned lit = p + 1u + (l15 ? 1u : 0u);
    const unsigned q = lit + L;
    const unsigned b2 = c[q + 2];
    const unsigned nx = q + 2u + (m15 ? 1u : 0u);
    t.L = L; t.M = M4 + 4u + (m15 ? b2 : 0u); t.off = (unsigned)c[q] | ((unsigned)c[q + 1] << 8); t.lit = lit; t.next = nx;
    bool ok = nx <= limq;
    if ((l15 && b1 == 255u) || (m15 && b2 == 255u)) {      // (its own TokF: t's address must not escape, or t lives in local memory)
        TokF g;
        ok = p < lim && tok_gen(sm, c, p, lim, e_rel, g);
        t.L = g.L; t.M = g.M; t.off = g.off; t.lit = g.lit; t.next = g.next;
    }
    return ok;
}

// Level A only needs where the token ends: the same decision as tok_fast without the fields (no TokF in the common path:
// its address would escape to tok_gen and put it in local memory).
__device__ __forceinline__ unsigned tok_next(const ParseSmem& sm, const uint8_t* c, const uint8_t* cp /* = c + p, with a compile-time offset */,
                                             unsigned p, unsigned lim, unsigned limq, unsigned e_rel) {
    const unsigned tk = cp[0], b1 = cp[1];
    const unsigned L4 = tk >> 4, M4 = tk & 15u;
    const bool l15 = L4 == 15u, m15 = M4 == 15u;
    const unsigned q = p + 1u + (l15 ? 1u + b1 : 0u) + L4;
    const unsigned b2 = c[q + 2];
    unsigned nx = q + 2u + (m15 ? 1u : 0u);
    bool ok = nx <= limq;
    if ((l15 && b1 == 255u) || (m15 && b2 == 255u)) { TokF t; ok = p < lim && tok_gen(sm, c, p, lim, e_rel, t); nx = t.next; }
    return ok ? nx : (TERM | p);
}

__global__ void __launch_bounds__(PT, 5)
lz4_parse_kernel(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
                 const uint32_t* __restrict__ wbase, const uint2* __restrict__ tickets, unsigned nticket,
                 unsigned long long* chain, WinInfo* __restrict__ winfo, SeqEnt* __restrict__ seqs, unsigned* ticket_ctr) {
    RCZ_DYN_SMEM(raw);
    ParseSmem& sm = *reinterpret_cast<ParseSmem*>(raw);
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned phase = 0;
    if (tid == 0) { mbar_init(&sm.bar, 1); mbar_fence_init(); }
    __syncthreads();

    for (;;) {
        __syncthreads();                                                  // everyone is done with the previous ticket's shared state
        if (tid == 0) sm.ticket_lo = atomicAdd(ticket_ctr, 1u);
        __syncthreads();
        const unsigned tkt = sm.ticket_lo;
        if (tkt >= nticket) break;
        const uint2 bw = tickets[tkt];
        const unsigned b = bw.x, w = bw.y;
        const uint8_t* in = in_base + in_off[b];
        const unsigned n = (unsigned)in_len[b];
        const unsigned base = w * (unsigned)W;
        const unsigned e_rel = n - base;                                  // > 0
        const unsigned lim = e_rel < (unsigned)VIS ? e_rel : (unsigned)VIS;
        const unsigned limq = lim - (lim == e_rel ? 1u : 0u);
        const unsigned widx = wbase[b] + w;
        const unsigned lead = (unsigned)((uintptr_t)(in + base) & 15);
        const uint8_t* c = sm.chunk + lead;
        if (tid == 0) { issue_window(sm.chunk, &sm.bar, in, n, base); sm.t_kind = TK_NONE; sm.lead = lead; }
        if (tid < 8) sm.gentry[tid] = NONE;
        mbar_wait(&sm.bar, phase); phase ^= 1;

        // ---- 0xFF bitmask of the visible bytes + runs of full words: 16 chunk bytes per lane (one aligned 16-byte load, a
        //      per-byte compare and a multiply that gathers the four result bits), two lanes make a mask word
        for (unsigned it = warp; it * 512u < 32u * ((unsigned)NFW + 2u); it += PT / 32) {
            const unsigned i0 = it * 512u + lane * 16u;                   // first chunk byte of this lane
            unsigned m16 = 0;
            if (i0 < (unsigned)CHUNK_BYTES) {
                const uint4 q = *reinterpret_cast<const uint4*>(sm.chunk + i0);
                const unsigned wv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned x = ~wv[k];                             // a 0xFF byte becomes 0x00
                    const unsigned t = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;   // top bit of every zero byte (no borrows)
                    m16 |= ((((t >> 7) * 0x00204081u) >> 21) & 0x0fu) << (4 * k);                // bits 0, 8, 16, 24 -> one nibble
                }
                const unsigned lo = lead, hi = lead + lim;                // valid chunk bytes [lo, hi)
                const unsigned a = lo > i0 ? (lo - i0 < 16u ? lo - i0 : 16u) : 0u, b = hi > i0 ? (hi - i0 < 16u ? hi - i0 : 16u) : 0u;
                m16 &= ((1u << b) - 1u) & ~((1u << a) - 1u);
            }
            unsigned v = m16 << (16u * (lane & 1u));
            v |= __shfl_xor_sync(RCZ_FULL, v, 1);
            const unsigned k = it * 16u + (lane >> 1);
            if (!(lane & 1u) && k < (unsigned)NFW + 2u) sm.ff[k] = v;
        }
        __syncthreads();
        if (warp == 0) {
            unsigned carry = 0;
            for (int blk = (NFW + 2 + 31) / 32 - 1; blk >= 0; --blk) {
                const unsigned kk = (unsigned)blk * 32 + lane;
                const bool full = kk < (unsigned)NFW + 2 && sm.ff[kk] == 0xffffffffu;
                const unsigned mm = __ballot_sync(RCZ_FULL, full);
                unsigned r = (unsigned)__ffs((int)~(mm >> lane)) - 1u;
                if (r >= 32u - lane) r = (32u - lane) + carry;
                if (kk < (unsigned)NFW + 2) sm.nfull[kk] = (uint16_t)r;
                carry = __shfl_sync(RCZ_FULL, r, 0);
            }
        }
        __syncthreads();

        // ---- levels A and B: this warp's 1 KiB group, segments in reverse order
        const unsigned gbase = warp * 1024u, gend = gbase + 1024u;
        unsigned aA[16];                                                  // exit of (segment s, this lane) from the segment, two per register
        const unsigned pbase = gbase + lane;                              // position of this lane in segment 0 of the group
        const uint8_t* const cpb = c + pbase;
        uint16_t* const ebb = sm.exitB + pbase;
        for (int s = 31; s >= 0; --s) {
            const unsigned se = gbase + (unsigned)s * 32u + 32u, p = pbase + (unsigned)s * 32u;
            unsigned a = tok_next(sm, c, cpb + s * 32, p, lim, limq, e_rel);
            // pointer doubling inside the segment; a token is >= 3 bytes, so 4 rounds always suffice (an early exit on a warp
            // vote was measured: slower)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const unsigned t2 = __shfl_sync(RCZ_FULL, a, (int)a);          // (the source lane is taken modulo 32: position a of this segment)
                if (a < se) a = t2;                                       // TERM codes are >= 0x8000: never inside the segment
            }
            if (s & 1) aA[s >> 1] = a << 16; else aA[s >> 1] |= a;
            unsigned bb = a;
            if (a < gend) bb = sm.exitB[a];                               // a later segment of this group: already final
            ebb[s * 32] = (uint16_t)bb;
            __syncwarp();
        }
        __syncthreads();

        // ---- chain: first true token of this window from the predecessor, hop the groups, tell the successor
        if (warp == 0) {
            unsigned long long word = CH_VALID;
            if (w > 0) {
                if (lane == 0) {
                    unsigned spins = 0;
                    while (!((word = ld_vol_u64(&chain[widx - 1])) & CH_VALID)) spin_pause(spins);
                }
                word = __shfl_sync(RCZ_FULL, word, 0);
            }
            const unsigned st = (unsigned)(word >> 60) & 3u, entry = (unsigned)word;
            if (st == ST_RUN && entry < base + (unsigned)W) {
                unsigned cur = entry - base;
                for (;;) {
                    if (lane == 0) sm.gentry[cur >> 10] = cur;
                    const unsigned x = sm.exitB[cur];
                    if (x & TERM) {
                        const unsigned pt = x & 0x7fffu;
                        const TokX t = tok_exact(in, n, base + pt);
                        if (lane == 0) {
                            sm.t_pos = pt; sm.t_kind = t.kind; sm.t_stage = t.stage; sm.t_L = t.L; sm.t_M = t.M; sm.t_off = t.off; sm.t_lit_abs = t.lit_abs;
                        }
                        if (t.kind == TK_NORMAL) word = CH_VALID | t.next;
                        else if (t.kind == TK_END) word = CH_VALID | ((unsigned long long)ST_END << 60) | n;
                        else word = CH_VALID | ((unsigned long long)ST_DEAD << 60);
                        break;
                    }
                    if (x >= (unsigned)W) { word = CH_VALID | (base + x); break; }
                    cur = x;
                }
            }
            if (lane == 0) { __threadfence(); *(volatile unsigned long long*)&chain[widx] = word; }
        }
        __syncthreads();

        // ---- true tokens of this warp's group: segment entries (uniform walk over the packed level-A exits) ...
        unsigned myEntry = NONE;
        {
            unsigned cur = sm.gentry[warp];
#pragma unroll
            for (int s = 0; s < 32; ++s) {
                if (cur != NONE && (cur >> 5) == (gbase >> 5) + (unsigned)s) {      // warp-uniform
                    if (lane == (unsigned)s) myEntry = cur;
                    const unsigned av = (s & 1) ? (aA[s >> 1] >> 16) : (aA[s >> 1] & 0xffffu);
                    const unsigned a = __shfl_sync(RCZ_FULL, av, (int)cur);
                    cur = a < gend ? a : NONE;
                }
            }
        }
        // ... then lane s walks segment s: pass 1 counts tokens and output bytes
        const unsigned segEnd = gbase + lane * 32u + 32u;
        const unsigned t_kind = sm.t_kind, t_pos = sm.t_pos;
        unsigned cnt = 0, len = 0;
        {
            unsigned p = myEntry;
            while (p < segEnd) {
                TokF t;
                if (!tok_fast(sm, c, p, lim, limq, e_rel, t)) { ++cnt; break; }     // the exactly parsed token: last of the window
                ++cnt; len += t.L + t.M; p = t.next;
            }
        }
        const unsigned icnt = warp_incl_scan_add(cnt), ilen = warp_incl_scan_add(len);
        if (lane == 31) { sm.wcnt[warp] = icnt; sm.wlen[warp] = ilen; }
        __syncthreads();
        unsigned idx = icnt - cnt, start = ilen - len, nseq = 0, fastlen = 0;
#pragma unroll
        for (int g = 0; g < PT / 32; ++g) {
            if ((unsigned)g < warp) { idx += sm.wcnt[g]; start += sm.wlen[g]; }
            nseq += sm.wcnt[g]; fastlen += sm.wlen[g];
        }
        // pass 2: descriptors
        int margin = 0x7fffffff; unsigned off0 = 0;
        {
            SeqEnt* D = seqs + (size_t)widx * SLOT;
            unsigned p = myEntry;
            while (p < segEnd) {
                TokF t; SeqEnt s;
                if (!tok_fast(sm, c, p, lim, limq, e_rel, t)) {
                    s.start = start; s.L = (uint32_t)sm.t_L; s.lit_abs = sm.t_lit_abs; s.off = sm.t_off;
                    if (sm.t_M > 0) {
                        const long long mg = (long long)start + (long long)sm.t_L - (long long)s.off;
                        if (mg < margin) margin = (int)(mg < -0x7fffffffll ? -0x7fffffffll : mg);
                        off0 |= s.off == 0;
                    }
                    D[idx] = s;
                    break;
                }
                s.start = start; s.L = t.L; s.lit_abs = base + t.lit; s.off = t.off;
                const int mg = (int)(start + t.L) - (int)t.off;
                margin = mg < margin ? mg : margin; off0 |= t.off == 0;
                D[idx] = s;
                ++idx; start += t.L + t.M; p = t.next;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const int m2 = __shfl_xor_sync(RCZ_FULL, margin, d);
            margin = m2 < margin ? m2 : margin;
            off0 |= __shfl_xor_sync(RCZ_FULL, off0, d);
        }
        if (lane == 0) { sm.wmargin[warp] = margin; sm.woff0[warp] = off0; }
        __syncthreads();
        if (tid == 0) {
            WinInfo wi;
            int mg = 0x7fffffff; unsigned o0 = 0;
            for (int g = 0; g < PT / 32; ++g) { mg = sm.wmargin[g] < mg ? sm.wmargin[g] : mg; o0 |= sm.woff0[g]; }
            wi.total = fastlen; wi.flags = o0 ? 1u : 0u;
            if (t_kind != TK_NONE) { wi.total += sm.t_L + sm.t_M; if (t_kind == TK_BAD) wi.flags |= sm.t_stage << 8; }
            wi.minmargin = mg; wi.nseq = nseq; wi.pad[0] = wi.pad[1] = wi.pad[2] = 0;
            winfo[widx] = wi;
        }
        (void)t_pos;
    }
}

// ======================================================================================================
// scan: output base of every window, block status in the reference's error order
// ======================================================================================================
__global__ void __launch_bounds__(256)
lz4_scan_kernel(const uint32_t* __restrict__ wbase, const uint32_t* __restrict__ nwin, const uint64_t* __restrict__ out_cap,
                const WinInfo* __restrict__ winfo, const SeqEnt* __restrict__ seqs, unsigned long long* __restrict__ obase,
                uint64_t* __restrict__ out_len, int32_t* __restrict__ status, int32_t* __restrict__ blkstate, unsigned nblocks) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned b = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    const unsigned nw = nwin[b], wb = wbase[b];
    const unsigned long long cap64 = out_cap[b];
    const unsigned long long cap = cap64 > 0x7fffffffull ? 0x7fffffffull : cap64;
    unsigned long long acc = 0;
    int st = RCZ_OK;
    for (unsigned w0 = 0; w0 < nw && st == RCZ_OK; w0 += 32) {
        const unsigned w = w0 + lane;
        WinInfo wi; wi.total = 0; wi.minmargin = 0x7fffffff; wi.nseq = 0; wi.flags = 0;
        if (w < nw) wi = winfo[wb + w];
        unsigned long long incl = wi.total;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(RCZ_FULL, incl, d);
            if (lane >= (unsigned)d) incl += t;
        }
        const unsigned long long ob = acc + incl - wi.total;
        if (w < nw) obase[wb + w] = ob;
        const bool fail = w < nw && ((wi.flags != 0) || (wi.minmargin != 0x7fffffff && (long long)ob + wi.minmargin < 0) || ob + wi.total > cap);
        const unsigned fm = __ballot_sync(RCZ_FULL, fail);
        if (fm) {
            // first failing window: its tokens in order, every lane a strided subset, the smallest failing token index wins
            const unsigned fl = (unsigned)__ffs((int)fm) - 1;
            const unsigned long long fob = __shfl_sync(RCZ_FULL, ob, (int)fl);
            const unsigned long long ftot = __shfl_sync(RCZ_FULL, wi.total, (int)fl);
            const unsigned fn = __shfl_sync(RCZ_FULL, wi.nseq, (int)fl), ffl = __shfl_sync(RCZ_FULL, wi.flags, (int)fl);
            const SeqEnt* D = seqs + (size_t)(wb + w0 + fl) * SLOT;
            const unsigned stage = (ffl >> 8) & 3u;
            unsigned best = 0xffffffffu; int code = RCZ_OK;
            for (unsigned j = lane; j < fn; j += 32) {
                const SeqEnt s = D[j];
                const unsigned long long nstart = j + 1 < fn ? (unsigned long long)D[j + 1].start : ftot;
                const unsigned long long o0 = fob + s.start, M = nstart - s.start - s.L;
                const bool last = j + 1 == fn;
                int cd = RCZ_OK;
                if (last && stage == 1) cd = RCZ_E_MALFORMED;
                else if (o0 + s.L > cap) cd = RCZ_E_OUTPUT_FULL;
                else if (last && stage == 2) cd = RCZ_E_MALFORMED;
                else if (M > 0 && (s.off == 0 || (unsigned long long)s.off > o0 + s.L)) cd = RCZ_E_MALFORMED;   // lz4.rs:93 underflow / App. B #9
                else if (o0 + s.L + M > cap) cd = RCZ_E_OUTPUT_FULL;
                if (cd != RCZ_OK) { best = j; code = cd; break; }
            }
            const unsigned bmin = __reduce_min_sync(RCZ_FULL, best);
            const unsigned src = __ballot_sync(RCZ_FULL, best == bmin && best != 0xffffffffu);
            st = src ? __shfl_sync(RCZ_FULL, code, __ffs((int)src) - 1) : RCZ_E_MALFORMED;
            const unsigned long long fs = bmin != 0xffffffffu ? fob + D[bmin].start : fob;
            acc = fs;
            break;
        }
        acc += __shfl_sync(RCZ_FULL, incl, 31);
    }
    // nothing of a failed block is materialised (lz4_mat_kernel skips it), so its length is reported as 0 in every mem_kind
    if (lane == 0) { out_len[b] = st == RCZ_OK ? acc : 0ull; status[b] = st; blkstate[b] = st; }
}

// ======================================================================================================
// materialise
// ======================================================================================================
struct __align__(16) UnitInfo { uint32_t nseq, tot, ob, pad; };
// Fused gather (multi-GPU): every finished tile of this GPU's output also goes to the same offset of the output buffers of up to
// 7 peers (pointers mapped over NVLink / NVSwitch), so the final gather of the decoded shards rides on the decode instead of
// following it as a separate collective.  The tile's whole 16-byte chunks leave the shared-memory output ring as TMA bulk
// stores (one per peer, issued by lanes 0..n-1 of warp 0: large NVLink writes, no store slots of the decoding warps); only the
// ragged first / last chunk of a unit is written with byte stores.  delta[p] = peer base - local base (a multiple of 16).
struct PeerOut { long long delta[7]; int n; };
struct MatSmem {
    __align__(16) uint8_t ring_[32 + RING + 32];     // 32-byte mirrors of the other end on both sides: source windows may overhang
    __align__(16) uint8_t win[2][CHUNK_BYTES];       // compressed window of the current / next unit (literal sources)
    __align__(16) SeqEnt ds[2][DCAP];                // sequence descriptors of the current / next unit
    __align__(16) uint4 lt[17];                      // lt[k]: low k bytes set
    UnitInfo ui[2];
    uint32_t cmask[NCH + 1];                         // per chunk of the tile: bit i = byte i is final
    uint16_t sbase[MT];                              // first piece of every sequence of the batch
    uint16_t pseq[3 * MT];                           // sequence (index in the batch) of every piece: <= 2 per sequence + one per chunk
    unsigned scan[40];
    uint32_t blk, ja_next;
    rcz_mbar barw[2], bard[2];
};

// literal bytes straight from the block's input (never written by these kernels): read-only path
__device__ __forceinline__ unsigned ld_gen_u32(uintptr_t a) { return __ldg(reinterpret_cast<const unsigned*>(a)); }
__device__ __forceinline__ unsigned ld_gen_u8(uintptr_t a) { return __ldg(reinterpret_cast<const uint8_t*>(a)); }

struct MatCtx {
    const SeqEnt* D; const SeqEnt* ds; const uint4* lt; uint32_t* cmask;
    const uint8_t* in; uint8_t* outb; uint8_t* ring; rcz_saddr rings, wins, cmasks;
    unsigned ob, uend, nseq, n, cbase, limw, hb, T0, T1, h, c00, nch, tend;
    const PeerOut* peer; int npeer;
};
__device__ __forceinline__ uint4 mat_desc(const MatCtx& k, unsigned j) {
    return j < (unsigned)DCAP ? *reinterpret_cast<const uint4*>(&k.ds[j]) : __ldg(reinterpret_cast<const uint4*>(&k.D[j]));
}
__device__ __forceinline__ unsigned mat_start(const MatCtx& k, unsigned j) {      // block offset of sequence j (uend past the last)
    if (j >= k.nseq) return k.uend;
    return k.ob + (j < (unsigned)DCAP ? k.ds[j].start : __ldg(&k.D[j].start));
}
// ring offset of block output offset x: (x + hb) mod RING, so that 16-byte chunks of the ring are 16-byte chunks of HBM
__device__ __forceinline__ unsigned ring_off(const MatCtx& k, unsigned x) { return (x + k.hb) % (unsigned)RING; }
__device__ __forceinline__ void ring_store(const MatCtx& k, unsigned cro, const uint4 v) {
    sts128_volatile(k.ring + cro, v);
    if (cro < 32u) sts128_volatile(k.ring + RING + cro, v);
    if (cro >= (unsigned)RING - 32u) sts128_volatile(k.ring + cro - RING, v);
}
__device__ __forceinline__ void ring_or(const MatCtx& k, unsigned wro, unsigned v) {      // wro: ring offset of a 32-bit word whose bytes start out zero
    atomicOr(reinterpret_cast<uint32_t*>(k.ring + wro), v);
    if (wro < 32u) atomicOr(reinterpret_cast<uint32_t*>(k.ring + RING + wro), v);
    if (wro >= (unsigned)RING - 32u) atomicOr(reinterpret_cast<uint32_t*>(k.ring + wro - RING), v);
}
// byte mask (bit i = byte i of its 16-byte chunk) of the tile offsets [a, b) that fall into chunk c
__device__ __forceinline__ unsigned chunk_bits(unsigned a, unsigned b, unsigned c) {
    const unsigned lo = a > 16u * c ? a - 16u * c : 0u, hi = b < 16u * c + 16u ? b - 16u * c : 16u;
    return ((1u << hi) - 1u) & ~((1u << lo) - 1u);
}

// bytes [lo, hi) of a 16-byte chunk to HBM: the ragged first / last chunk of a unit (out of line: the sixteen byte predicates
// must not be hoisted into every piece)
__device__ __noinline__ void store_chunk_bytes(uint8_t* g, uint4 v, unsigned lo, unsigned hi) {
    const unsigned wv[4] = {v.x, v.y, v.z, v.w};
    for (unsigned i = lo; i < hi; ++i) g[i] = (uint8_t)(wv[i >> 2] >> ((i & 3u) * 8u));
}

// Byte-by-byte fetch of a piece (rare, out of line): literals at the very edge of the input (kind 1) and the head of an
// overlapping match with offset < 16 (kind 3), where byte i of the piece is seed byte ((pos - ms) + i) mod off.
__device__ __noinline__ uint4 piece_bytes(rcz_saddr rings, unsigned hb, unsigned kind, unsigned pos, unsigned ms, unsigned doff, uintptr_t gsrc, unsigned d0, unsigned len) {
    unsigned long long tl = 0, th = 0;
    unsigned sidx = 0, sro = 0;
    if (kind == 3) { sidx = (pos - ms) % doff; sro = (ms - doff + hb) % (unsigned)RING; }      // ring offset of the seed
    for (unsigned i = 0; i < len; ++i) {
        unsigned v;
        if (kind == 3) {
            unsigned a = sro + sidx; a = a >= (unsigned)RING ? a - RING : a;
            v = lds8_volatile(rings + a);
            sidx = sidx + 1 == doff ? 0 : sidx + 1;
        } else v = ld_gen_u8(gsrc + i);
        const unsigned bi = d0 + i;
        if (bi < 8u) tl |= (unsigned long long)v << (8u * bi); else th |= (unsigned long long)v << (8u * (bi - 8u));
    }
    uint4 t; t.x = (unsigned)tl; t.y = (unsigned)(tl >> 32); t.z = (unsigned)th; t.w = (unsigned)(th >> 32);
    return t;
}

// source distance for byte kk >= off of an overlapping match: the largest whole number of periods that stays inside the
// periodic region and within 32 KiB (out of line: two integer divisions that the common path must not pay for)
__device__ __noinline__ unsigned far_back(unsigned kk, unsigned off) {
    const unsigned mmax = kk / off + 1u, reach = 32768u / off;
    return (mmax < reach ? mmax : (reach ? reach : 1u)) * off;
}

// One piece = the part of a literal run or of a match that falls into one 16-byte chunk of the output.  Piece p of the batch
// belongs to the sequence i with sbase[i] <= p < sbase[i + 1]; its rank there says which part and which chunk.  The piece is
// fetched with 4-byte loads + funnel shifts (staged compressed window, ring) and OR-ed into the zeroed chunk; per-chunk byte
// masks order pieces whose source lies inside the tile, and whoever completes a chunk stores its 16 bytes to HBM.
// All lanes of the warp call it together; lanes with active == false only take part in the votes.
template <bool DEP, bool LIT, bool PEERS>
__device__ __forceinline__ void mat_piece(const MatCtx& k, const uint16_t* sbase, const uint16_t* pseq, unsigned j0, unsigned p, bool active) {
    constexpr unsigned NOCHK = 0xffffu;
    const unsigned T0 = k.T0, T1 = k.T1, c00 = k.c00;
    unsigned len = 0, d0 = 0, cc = 0, kind = 0, pos = 0, ms = 0, doff = 0, dlit = 0, dS = 0;
    unsigned cA = NOCHK, cB = NOCHK, mA = 0, mB = 0;
    rcz_saddr srcs = 0;
    bool done = !active;
    // sequence of the piece: written by the sequence's thread when the pieces were counted
    const unsigned si = active ? (unsigned)pseq[p] : 0u;
    if (active) {
        const unsigned j = j0 + si, r = p - sbase[si];
        const uint4 q = mat_desc(k, j);
        dS = k.ob + q.x; dlit = q.z; doff = q.w; ms = dS + q.y;
        const unsigned send = mat_start(k, j + 1);
        const unsigned la = dS > T0 ? dS : T0, lb = ms < T1 ? ms : T1;          // literal run inside the tile (lz4.rs:75-85)
        unsigned nl = 0, ca = 0;
        if (lb > la) { ca = (la - c00) >> 4; nl = ((lb - 1u - c00) >> 4) - ca + 1u; }
        unsigned a, b;
        if (LIT && r < nl) { cc = ca + r; a = la; b = lb; }
        else {                                                                  // match inside the tile (lz4.rs:96-107, cp lz4.rs:131-140)
            a = ms > T0 ? ms : T0; b = send < T1 ? send : T1;
            cc = ((a - c00) >> 4) + (LIT ? r - nl : r); kind = 2;
        }
        const unsigned cs = c00 + 16u * cc;                                     // block offset of the chunk's byte 0 (modular)
        if (cc && a < cs) a = cs;                                               // cc == 0: cs may lie below zero, a >= T0 > cs anyway
        if (b - cs > 16u) b = cs + 16u;
        pos = a; len = b - a; d0 = a - cs;
        unsigned x0 = 0, x1 = 0;                                                // source bytes inside the tile, as tile offsets [x0, x1)
        if (kind == 0) {
            const unsigned lw = dlit + (pos - dS) - k.cbase;
            if (lw + len <= k.limw) srcs = k.wins + lw; else kind = 1;
        } else {
            const unsigned kk = pos - ms;
            // byte pos of a match equals byte pos - m * off for every m that stays inside [ms - off, pos): take the source a
            // whole number of periods back, as far as the ring safely reaches, so that long overlapping matches form no
            // chunk-to-chunk chain and never look behind the ring
            unsigned back = doff;
            if (kk >= doff) back = far_back(kk, doff);                          // overlapping match (rare): kept out of line, it divides
            if (len > back) {                                                   // fewer than 16 bytes back: offset < 16 at the head of an overlapping match
                kind = 3;                                                       // periodic: seed [ms - off, ms)
                if (DEP && ms > T0) { x0 = (ms - doff > T0 ? ms - doff : T0) - c00; x1 = ms - c00; }
            } else {
                const unsigned sp = pos - back;                                 // block offset of the first source byte
                kind = 0;
                srcs = k.rings + ring_off(k, sp);
                if (DEP && sp + len > T0) { x0 = (sp > T0 ? sp : T0) - c00; x1 = sp + len - c00; }
            }
        }
        if (DEP && x1 > x0) {
            cA = x0 >> 4; mA = chunk_bits(x0, x1, cA);
            const unsigned c2 = (x1 - 1u) >> 4;
            if (c2 != cA) { cB = c2; mB = chunk_bits(x0, x1, c2); }
        }
    }
    while (__any_sync(RCZ_FULL, !done)) {
        bool progressed = false;
        bool ready = true;
        if (DEP && !done) {
            if (cA != NOCHK) ready = (lds32_volatile(k.cmasks + 4u * cA) & mA) == mA;          // (shared-memory loads, not generic ones)
            if (cB != NOCHK) ready = ready && (lds32_volatile(k.cmasks + 4u * cB) & mB) == mB;
        }
        if (!done) {
            if (ready) {
                uint4 t;
                if (kind == 0) {
                    const rcz_saddr a = srcs - d0, a0 = a & ~(rcz_saddr)3;
                    const unsigned sh = (unsigned)(a & 3) * 8u;
                    const unsigned w0 = lds32_volatile(a0), w1 = lds32_volatile(a0 + 4), w2 = lds32_volatile(a0 + 8), w3 = lds32_volatile(a0 + 12), w4 = lds32_volatile(a0 + 16);
                    t.x = __funnelshift_r(w0, w1, sh); t.y = __funnelshift_r(w1, w2, sh); t.z = __funnelshift_r(w2, w3, sh); t.w = __funnelshift_r(w3, w4, sh);
                } else {
                    const unsigned la = dlit + (pos - dS);
                    const uintptr_t gsrc = (uintptr_t)(k.in + la);
                    if (kind == 1 && la >= 20u && la + 20u <= k.n) {            // literals beyond the staged window: straight from the input
                        const uintptr_t a = gsrc - d0, a0 = a & ~(uintptr_t)3;
                        const unsigned sh = (unsigned)(a & 3) * 8u;
                        const unsigned w0 = ld_gen_u32(a0), w1 = ld_gen_u32(a0 + 4), w2 = ld_gen_u32(a0 + 8), w3 = ld_gen_u32(a0 + 12), w4 = ld_gen_u32(a0 + 16);
                        t.x = __funnelshift_r(w0, w1, sh); t.y = __funnelshift_r(w1, w2, sh); t.z = __funnelshift_r(w2, w3, sh); t.w = __funnelshift_r(w3, w4, sh);
                    } else t = piece_bytes(k.rings, k.hb, kind, pos, ms, doff, gsrc, d0, len);      // byte by byte (rare): input edges, periodic matches
                }
                const uint4 mlo = k.lt[d0], mhi = k.lt[d0 + len];
                const unsigned cro = ring_off(k, c00 + 16u * cc);
                const unsigned vx = t.x & mhi.x & ~mlo.x, vy = t.y & mhi.y & ~mlo.y, vz = t.z & mhi.z & ~mlo.z, vw = t.w & mhi.w & ~mlo.w;
                uint32_t* const cw = reinterpret_cast<uint32_t*>(k.ring + cro);
                if (cro >= 32u && cro < (unsigned)RING - 32u) {                 // (all but the two chunks at either end of the ring: no mirror)
                    if (mhi.x & ~mlo.x) atomicOr(cw, vx);
                    if (mhi.y & ~mlo.y) atomicOr(cw + 1, vy);
                    if (mhi.z & ~mlo.z) atomicOr(cw + 2, vz);
                    if (mhi.w & ~mlo.w) atomicOr(cw + 3, vw);
                } else {
                    if (mhi.x & ~mlo.x) ring_or(k, cro, vx);
                    if (mhi.y & ~mlo.y) ring_or(k, cro + 4, vy);
                    if (mhi.z & ~mlo.z) ring_or(k, cro + 8, vz);
                    if (mhi.w & ~mlo.w) ring_or(k, cro + 12, vw);
                }
                __threadfence_block();
                const unsigned bits = ((1u << len) - 1u) << d0;
                const unsigned old = atomicOr(&k.cmask[cc], bits);
                const unsigned hi = cc + 1u == k.nch ? k.tend - 16u * cc : 16u;
                const unsigned full = (1u << hi) - 1u;
                if ((old | bits) == full) {                                     // this piece completed the chunk: its 16 bytes go to HBM
                    __threadfence_block();
                    const uint4 v = lds128_volatile(k.ring + cro);
                    uint8_t* g = k.outb + T0 - k.h + 16u * cc;                  // pointer arithmetic: T0 - h alone may wrap below zero
                    const unsigned lo = cc ? 0u : k.h;
                    if (lo == 0 && hi == 16u) {
                        *reinterpret_cast<uint4*>(g) = v;                           // (peers: with the tile's bulk store)
                    } else {
                        store_chunk_bytes(g, v, lo, hi);                        // first / last chunk of a unit (rare, out of line)
                        if (PEERS) for (int pp = 0; pp < k.npeer; ++pp) store_chunk_bytes(g + k.peer->delta[pp], v, lo, hi);
                    }
                }
                done = true; progressed = true;
            }
        }
#ifdef RCZ_EMU
        if (!__any_sync(RCZ_FULL, progressed)) emu::yield();
#else
        (void)progressed;                                                       // no back-off: few warps ever wait here, and a sleep costs more than the producer takes
#endif
    }
}

// Unit w of the block: its output size / position and the TMA copies of its compressed window and its descriptors.
__device__ __forceinline__ void mat_prefetch(MatSmem& sm, unsigned w, unsigned widx, const WinInfo* winfo, const unsigned long long* obase,
                                             const SeqEnt* seqs, const uint8_t* in, unsigned n) {
    const unsigned buf = w & 1;
    const WinInfo wi = winfo[widx];
    UnitInfo u; u.nseq = wi.nseq; u.tot = (unsigned)wi.total; u.ob = (unsigned)obase[widx]; u.pad = 0;
    sm.ui[buf] = u;
    if (u.tot > 0) {
        issue_window(sm.win[buf], &sm.barw[buf], in, n, w * (unsigned)W);
        const unsigned nd = u.nseq < (unsigned)DCAP ? u.nseq : (unsigned)DCAP;
        mbar_expect_tx(&sm.bard[buf], nd * 16u);
        tma_load_1d(sm.ds[buf], seqs + (size_t)widx * SLOT, nd * 16u, &sm.bard[buf]);
    }
}

template <bool PEERS>
__global__ void __launch_bounds__(MT, 2)
lz4_mat_kernel(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
               uint8_t* out_base, const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ wbase, const uint32_t* __restrict__ nwin,
               unsigned nblocks, const WinInfo* __restrict__ winfo, const SeqEnt* __restrict__ seqs,
               const unsigned long long* __restrict__ obase, const int32_t* __restrict__ blkstate, unsigned* ticket_ctr,
               const __grid_constant__ PeerOut peers) {
    RCZ_DYN_SMEM(raw);
    MatSmem& sm = *reinterpret_cast<MatSmem*>(raw);
    const unsigned tid = threadIdx.x;
    unsigned phw[2] = {0, 0};
    if (tid == 0) { mbar_init(&sm.barw[0], 1); mbar_init(&sm.barw[1], 1); mbar_init(&sm.bard[0], 1); mbar_init(&sm.bard[1], 1); mbar_fence_init(); }
    if (tid < 17) {
        uint4 m;
        m.x = tid >= 4 ? 0xffffffffu : ((1u << (8 * tid)) - 1u);
        m.y = tid >= 8 ? 0xffffffffu : tid <= 4 ? 0u : ((1u << (8 * (tid - 4))) - 1u);
        m.z = tid >= 12 ? 0xffffffffu : tid <= 8 ? 0u : ((1u << (8 * (tid - 8))) - 1u);
        m.w = tid >= 16 ? 0xffffffffu : tid <= 12 ? 0u : ((1u << (8 * (tid - 12))) - 1u);
        sm.lt[tid] = m;
    }
    __syncthreads();

    for (;;) {
        __syncthreads();                                                  // everyone is done with the previous block's shared state
        if (tid == 0) sm.blk = atomicAdd(ticket_ctr, 1u);
        __syncthreads();
        const unsigned b = sm.blk;
        if (b >= nblocks) break;
        const unsigned nw = nwin[b], wb = wbase[b];
        if (blkstate[b] != RCZ_OK || nw == 0) continue;                   // nothing of a failed block is materialised (uniform)
        MatCtx k;
        k.in = in_base + in_off[b];
        k.n = (unsigned)in_len[b];
        k.outb = out_base + out_off[b];
        k.hb = (unsigned)((uintptr_t)k.outb & 15);
        k.ring = sm.ring_ + 32;
        k.rings = saddr_of(k.ring);
        k.lt = sm.lt; k.cmask = sm.cmask; k.cmasks = saddr_of(sm.cmask);
        k.peer = &peers; k.npeer = PEERS ? peers.n : 0;
        if (tid == 0) { fence_proxy_async_smem(); mat_prefetch(sm, 0, wb, winfo, obase, seqs, k.in, k.n); }

        for (unsigned w = 0; w < nw; ++w) {
            const unsigned buf = w & 1;
            __syncthreads();                                              // unit w's info is visible; unit w-1 is complete, its buffers are free
            if (tid == 0 && w + 1 < nw) { fence_proxy_async_smem(); mat_prefetch(sm, w + 1, wb + w + 1, winfo, obase, seqs, k.in, k.n); }
            const UnitInfo u = sm.ui[buf];
            if (u.tot == 0) continue;                                     // uniform
            mbar_wait(&sm.barw[buf], phw[buf]); mbar_wait(&sm.bard[buf], phw[buf]); phw[buf] ^= 1;
            k.nseq = u.nseq; k.ob = u.ob; k.uend = u.ob + u.tot;
            k.cbase = w * (unsigned)W;
            const unsigned e_rel = k.n - k.cbase;
            k.limw = e_rel < (unsigned)VIS ? e_rel : (unsigned)VIS;
            k.wins = saddr_of(sm.win[buf] + ((uintptr_t)(k.in + k.cbase) & 15));        // window position 0
            k.D = seqs + (size_t)(wb + w) * SLOT;
            k.ds = sm.ds[buf];
            // tiles of equal size (a multiple of 16, at most TCAP), cut in HBM-aligned coordinates
            const unsigned h0 = (k.ob + k.hb) & 15u;
            const unsigned span = h0 + u.tot;
            const unsigned ntile = (span + TCAP - 1) / TCAP;
            const unsigned ts = (((span + ntile - 1) / ntile) + 15u) & ~15u;
            unsigned ja = 0;
            for (unsigned t = 0; t < ntile; ++t) {
                k.h = t ? 0u : h0;
                k.T0 = t ? k.ob - h0 + t * ts : k.ob;
                const unsigned e1 = k.ob - h0 + (t + 1) * ts;
                k.T1 = e1 < k.uend ? e1 : k.uend;
                k.c00 = k.T0 - k.h;
                k.tend = k.h + (k.T1 - k.T0);
                k.nch = (k.tend + 15u) >> 4;
                // ---- the tile's chunks start out zero (the bytes below T0 in the first one belong to the previous tile and stay)
                if (tid < k.nch) {
                    const unsigned cro = ring_off(k, k.c00 + 16u * tid);
                    uint4 v = make_uint4(0, 0, 0, 0);
                    unsigned pre = 0;
                    if (tid == 0 && k.h) {
                        const uint4 o = lds128_volatile(k.ring + cro), m = sm.lt[k.h];
                        v.x = o.x & m.x; v.y = o.y & m.y; v.z = o.z & m.z; v.w = o.w & m.w;
                        pre = (1u << k.h) - 1u;
                    }
                    ring_store(k, cro, v);
                    sm.cmask[tid] = pre;
                }
                if (tid == 0 && t == 0) sm.ja_next = 0;                   // (later tiles: ja_next already holds ja; rewriting it would race with the slower threads' read below)
                __syncthreads();
                // ---- sequences in batches of MT: pieces per sequence, prefix scan, then one piece per thread and round
                for (unsigned j0 = ja;; j0 += MT) {
                    const unsigned j = j0 + tid;
                    unsigned cnt = 0; int more = 0;
                    if (j < k.nseq) {
                        const uint4 q = mat_desc(k, j);
                        const unsigned s = k.ob + q.x, ms = s + q.y;
                        if (s < k.T1) {
                            more = 1;
                            const unsigned send = mat_start(k, j + 1);
                            const unsigned la = s > k.T0 ? s : k.T0, lb = ms < k.T1 ? ms : k.T1;
                            if (lb > la) cnt = ((lb - 1u - k.c00) >> 4) - ((la - k.c00) >> 4) + 1u;
                            const unsigned ma = ms > k.T0 ? ms : k.T0, mb = send < k.T1 ? send : k.T1;
                            if (mb > ma) {
                                const unsigned nm = ((mb - 1u - k.c00) >> 4) - ((ma - k.c00) >> 4) + 1u;
                                cnt += nm;
                            }
                            if (send >= k.T1) atomicMax(&sm.ja_next, j);       // the sequence that holds the next tile's first byte
                        }
                    }
                    // block prefix scan of the piece counts with ONE barrier: warp-shuffle scan, warp totals through shared memory,
                    // every warp adds up the totals before it
                    const unsigned incl = warp_incl_scan_add(cnt);
                    if ((tid & 31u) == 31u) sm.scan[tid >> 5] = incl;
                    __syncthreads();
                    // (lane g of every warp holds warp g's total; one more shuffle scan gives the offsets)
                    const unsigned wt = (tid & 31u) < MT / 32 ? sm.scan[tid & 31u] : 0u;
                    const unsigned wi = warp_incl_scan_add(wt);
                    const unsigned P = __shfl_sync(RCZ_FULL, wi, MT / 32 - 1);
                    const unsigned base = incl - cnt + __shfl_sync(RCZ_FULL, wi - wt, (int)(tid >> 5));
                    sm.sbase[tid] = (uint16_t)base;
                    for (unsigned i = 0; i < cnt; ++i) sm.pseq[base + i] = (uint16_t)tid;      // piece -> sequence (a few per thread; a long match has one per chunk)
                    __syncthreads();
                    for (unsigned p0 = 0; p0 < P; p0 += MT) {
                        if (p0 + (tid & ~31u) < P) mat_piece<true, true, PEERS>(k, sm.sbase, sm.pseq, j0, p0 + tid, p0 + tid < P);    // warp-uniform
                    }
                    if (PEERS) fence_proxy_async_smem();                  // this thread's ring writes, for the bulk stores below
                    if (!__syncthreads_or(more && j0 + MT < k.nseq)) break;
                }
                if (PEERS && tid < (unsigned)k.npeer) {
                    // the tile is complete in the ring: its whole chunks go to peer `tid` in one bulk store (two where the ring wraps)
                    const unsigned c_lo = k.h ? 1u : 0u, c_hi = (k.tend & 15u) ? k.nch - 1u : k.nch;
                    if (c_hi > c_lo) {
                        const unsigned r0 = ring_off(k, k.c00 + 16u * c_lo), bytes = 16u * (c_hi - c_lo);
                        uint8_t* g = k.outb + k.T0 - k.h + 16u * c_lo + peers.delta[tid];
                        const unsigned first = bytes < (unsigned)RING - r0 ? bytes : (unsigned)RING - r0;
                        tma_store_1d(g, k.ring + r0, first);
                        if (bytes > first) tma_store_1d(g + first, k.ring, bytes - first);
                    }
                    bulk_commit();
                    // at most PEER_INFLIGHT tiles still being read from the ring: PEER_INFLIGHT + 2 tiles back from the tile
                    // that is zeroed next stays below the 64 KiB + one tile that the ring keeps
                    bulk_wait_read<PEER_INFLIGHT>();
                }
                ja = sm.ja_next;                                          // (every thread reads it before the next tile's barrier lets thread 0 overwrite it)
            }
        }
        if (PEERS && tid < (unsigned)k.npeer) bulk_wait_read0();          // the next block starts the ring over (barrier at the loop head)
    }
}

}  // namespace lz4k

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
struct Lz4Plan {
    std::vector<uint32_t> nw, wbase;            // per block: windows, first compact window index
    std::vector<uint32_t> tickets;              // (block, window) pairs in (window, block) order — per range, see lz4_plan_range
    size_t totwin = 0;
};
// windows of blocks [b0, b0 + nb) in (window index, block) order: every block's chain advances one link per level
static void lz4_plan_range(const Lz4Plan& pl, size_t b0, size_t nb, std::vector<uint32_t>& tk) {
    std::vector<uint32_t> order(nb);
    uint32_t maxw = 0;
    for (size_t i = 0; i < nb; ++i) maxw = std::max(maxw, pl.nw[b0 + i]);
    std::vector<uint32_t> cnt(maxw + 2, 0);
    for (size_t i = 0; i < nb; ++i) cnt[pl.nw[b0 + i]]++;
    // blocks by descending window count (counting sort, stable)
    std::vector<uint32_t> pos(maxw + 2, 0);
    for (int v = (int)maxw - 1; v >= 0; --v) pos[v] = pos[v + 1] + cnt[v + 1];
    for (size_t i = 0; i < nb; ++i) order[pos[pl.nw[b0 + i]]++] = (uint32_t)i;
    size_t active = nb;
    for (uint32_t w = 0; w < maxw; ++w) {
        while (active > 0 && pl.nw[b0 + order[active - 1]] <= w) --active;
        for (size_t i = 0; i < active; ++i) { tk.push_back(order[i]); tk.push_back(w); }
    }
}

struct Lz4Dev {
    lz4k::PeerOut peers;
    const uint32_t *wbase, *nw; const uint2* tickets;
    unsigned long long* chain; unsigned* done; unsigned* ctr; lz4k::WinInfo* winfo; unsigned long long* obase; int32_t* blkstate; lz4k::SeqEnt* seqs;
};

// the three kernels for blocks [b0, b0 + nb) on `stream`; descriptor pointers are already offset to b0, tickets to the range
static int lz4_enqueue(rcz_ctx* c, rt_stream_t stream, const Lz4Dev& d, size_t range_idx, const uint8_t* din, const uint64_t* in_off, const uint64_t* in_len,
                       uint8_t* dout, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t nb,
                       const uint2* tk, size_t ntk) {
    using namespace lz4k;
    unsigned* ctr = d.ctr + 2 * range_idx;
    const bool timed = stream == c->stream;                                  // single-shot path: per-kernel event marks
    if (timed) { int st = ctx_stage_mark(c, 0); if (st) return st; }
    if (ntk) {
        const size_t g1 = std::min<size_t>(ntk, (size_t)c->sm_count * 5);
        RCZ_LAUNCH(lz4_parse_kernel, (unsigned)g1, PT, sizeof(ParseSmem), stream, din, in_off, in_len, d.wbase, tk, (unsigned)ntk, d.chain, d.winfo, d.seqs, ctr);
        c->launches++;
        RCZ_CK(c, rt_last_error());
    }
    if (timed) { int st = ctx_stage_mark(c, 1); if (st) return st; }
    RCZ_LAUNCH(lz4_scan_kernel, (unsigned)((nb + 7) / 8), 256, 0, stream, d.wbase, d.nw, out_cap, d.winfo, d.seqs, d.obase, out_len, status, d.blkstate, (unsigned)nb);
    c->launches++;
    RCZ_CK(c, rt_last_error());
    if (timed) { int st = ctx_stage_mark(c, 2); if (st) return st; }
    if (ntk) {
        const size_t g2 = std::min<size_t>(nb, (size_t)c->sm_count * 2);
        if (d.peers.n)
            RCZ_LAUNCH(lz4_mat_kernel<true>, (unsigned)g2, MT, sizeof(MatSmem), stream, din, in_off, in_len, dout, out_off, d.wbase, d.nw, (unsigned)nb, d.winfo, d.seqs,
                       d.obase, d.blkstate, ctr + 1, d.peers);
        else
            RCZ_LAUNCH(lz4_mat_kernel<false>, (unsigned)g2, MT, sizeof(MatSmem), stream, din, in_off, in_len, dout, out_off, d.wbase, d.nw, (unsigned)nb, d.winfo, d.seqs,
                       d.obase, d.blkstate, ctr + 1, d.peers);
        c->launches++;
        RCZ_CK(c, rt_last_error());
    }
    if (timed) { int st = ctx_stage_mark(c, 3); if (st) return st; }
    return RCZ_OK;
}

// plan + workspaces + uploads shared by the single-shot and the pipelined path.  `cut` = block ranges (one per launch group).
struct Lz4Job {
    Lz4Plan pl;
    std::vector<size_t> cut, tk_off;            // tk_off[k]: first ticket (pair index) of range k
    Lz4Dev dev;
};
static int lz4_prepare(rcz_ctx* c, DescStager& ds, const uint64_t* in_len, size_t nblocks, const std::vector<size_t>& cut, Lz4Job& job) {
    using namespace lz4k;
    Lz4Plan& pl = job.pl;
    job.cut = cut;
    const size_t nr = cut.size() - 1;
    auto& pc = c->lz4_plan;
    if (pc.in_len.size() == nblocks && pc.cut == cut && memcmp(pc.in_len.data(), in_len, nblocks * 8) == 0) {
        pl.nw = pc.nw; pl.wbase = pc.wbase; pl.tickets = pc.tickets; pl.totwin = pc.totwin; job.tk_off = pc.tk_off;
    } else {
        pl.nw.resize(nblocks); pl.wbase.resize(nblocks + 1);
        size_t tot0 = 0;
        for (size_t i = 0; i < nblocks; ++i) { pl.wbase[i] = (uint32_t)tot0; pl.nw[i] = (uint32_t)((in_len[i] + W - 1) / W); tot0 += pl.nw[i]; }
        pl.wbase[nblocks] = (uint32_t)tot0; pl.totwin = tot0;
        if (tot0 >= 0x7fffffffull) return RCZ_E_ARG;
        pl.tickets.reserve(2 * tot0);
        for (size_t k = 0; k < nr; ++k) { job.tk_off.push_back(pl.tickets.size() / 2); lz4_plan_range(pl, cut[k], cut[k + 1] - cut[k], pl.tickets); }
        job.tk_off.push_back(pl.tickets.size() / 2);
        pc.in_len.assign(in_len, in_len + nblocks); pc.cut = cut; pc.tk_off = job.tk_off;
        pc.nw = pl.nw; pc.wbase = pl.wbase; pc.tickets = pl.tickets; pc.totwin = pl.totwin;
    }
    const size_t tot = pl.totwin;
    const size_t i_wb = ds.add_in(pl.wbase.data(), (nblocks + 1) * 4), i_nw = ds.add_in(pl.nw.data(), nblocks * 4);
    const size_t i_tk = ds.add_in(pl.tickets.data(), pl.tickets.size() * 4);
    int st = ds.upload(); if (st) return st;
    // control words (zeroed every call): chain, done, ticket counters
    const size_t ctl_bytes = tot * 12 + nr * 8 + 64;
    void* ctl; st = ctx_ws(c, WS_A, ctl_bytes, &ctl); if (st) return st;
    RCZ_CK(c, rt_memset(ctl, 0, ctl_bytes, c->stream));
    void* wsb; st = ctx_ws(c, WS_B, tot * (sizeof(WinInfo) + 8) + nblocks * 4 + 64, &wsb); if (st) return st;
    void* wsc; st = ctx_ws(c, WS_C, tot * (size_t)SLOT * sizeof(SeqEnt) + 64, &wsc); if (st) return st;
    Lz4Dev& d = job.dev;
    memset(&d.peers, 0, sizeof d.peers);
    d.wbase = ds.in_ptr<uint32_t>(i_wb); d.nw = ds.in_ptr<uint32_t>(i_nw); d.tickets = (const uint2*)ds.in_ptr<uint32_t>(i_tk);
    d.chain = (unsigned long long*)ctl; d.done = (unsigned*)((uint8_t*)ctl + tot * 8); d.ctr = (unsigned*)((uint8_t*)ctl + tot * 12 + ((-(long)(tot * 12)) & 7));
    d.winfo = (WinInfo*)wsb; d.obase = (unsigned long long*)((uint8_t*)wsb + tot * sizeof(WinInfo)); d.blkstate = (int32_t*)((uint8_t*)wsb + tot * (sizeof(WinInfo) + 8));
    d.seqs = (SeqEnt*)wsc;
    RCZ_CK(c, RCZ_KERNEL_SMEM_OPTIN(lz4_parse_kernel, sizeof(ParseSmem)));
    RCZ_CK(c, RCZ_KERNEL_SMEM_OPTIN(lz4_mat_kernel<false>, sizeof(MatSmem)));
    RCZ_CK(c, RCZ_KERNEL_SMEM_OPTIN(lz4_mat_kernel<true>, sizeof(MatSmem)));
    return RCZ_OK;
}
// kernels of range k of a prepared job (descriptor arrays in `ds` slots 0..3 = in_off, in_len, out_off, out_cap)
static int lz4_enqueue_range(rcz_ctx* c, rt_stream_t stream, const Lz4Job& job, const DescStager& ds, size_t k, const uint8_t* din, uint8_t* dout,
                             uint64_t* d_len, int32_t* d_st) {
    const size_t b0 = job.cut[k], nb = job.cut[k + 1] - b0;
    Lz4Dev d = job.dev;
    d.wbase += b0; d.nw += b0; d.blkstate += b0;
    return lz4_enqueue(c, stream, d, k, din, ds.in_ptr<uint64_t>(0) + b0, ds.in_ptr<uint64_t>(1) + b0, dout, ds.in_ptr<uint64_t>(2) + b0,
                       ds.in_ptr<uint64_t>(3) + b0, d_len + b0, d_st + b0, nb, job.dev.tickets + job.tk_off[k], job.tk_off[k + 1] - job.tk_off[k]);
}

// Host-resident buffers, pipelined: the batch is cut into chunks of consecutive blocks; chunk k's compressed bytes go up
// (stream 0) while chunk k-1 decodes (stream 1) and chunk k-2's output comes down (stream 2), so the call costs about
// max(H2D, kernel, D2H) instead of their sum.  Host buffers should be pinned (rcz_host_alloc) for the copies to overlap.
static int lz4_host_pipelined(rcz_ctx* c, DescStager& ds, const Lz4Job& job, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                              const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t nblocks) {
    const std::vector<size_t>& cut = job.cut;
    const size_t nchunks = cut.size() - 1;
    int st = ctx_aux_streams(c); if (st) return st;
    st = ctx_events(c, 3 * nchunks + 1); if (st) return st;
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
    uint64_t* d_len = ds.out_ptr<uint64_t>(0); int32_t* d_st = ds.out_ptr<int32_t>(1);
    // descriptors, plan and the zeroed control words were enqueued on c->stream: the kernel streams wait for them
    RCZ_CK(c, rt_event_record(c->events[3 * nchunks], c->stream));
    for (size_t k = 0; k < nchunks; ++k) {
        const size_t b0 = cut[k], nb = cut[k + 1] - cut[k];
        uint64_t lo = UINT64_MAX, hi = 0;
        for (size_t i = b0; i < b0 + nb; ++i) { if (!in_len[i]) continue; lo = std::min(lo, in_off[i]); hi = std::max(hi, in_off[i] + in_len[i]); }
        if (lo != UINT64_MAX) RCZ_CK(c, rt_h2d((uint8_t*)din + lo, (const uint8_t*)in_base + lo, (size_t)(hi - lo), c->stream));
        RCZ_CK(c, rt_event_record(c->events[3 * k], c->stream));
        const rt_stream_t ks = c->aux[1 + (k & 7)];                // chunk kernels run concurrently: one chunk alone cannot fill the GPU
        RCZ_CK(c, rt_stream_wait_event(ks, c->events[3 * k]));
        st = lz4_enqueue_range(c, ks, job, ds, k, din, dout, d_len, d_st); if (st) return st;
        RCZ_CK(c, rt_event_record(c->events[3 * k + 1], ks));
    }
    // output comes down on one D2H stream in chunk order, each chunk behind its kernels' event and with NO host involvement: the
    // whole out_cap span of every block is copied (adjacent blocks merged into one copy), so nothing has to wait for the lengths;
    // the results (out_len, status) follow in one small copy at the end.  Bytes of a block's region beyond out_len[i] are
    // unspecified afterwards (rcz.h).
    void* pin; st = ctx_pinned(c, nblocks * 12 + 64, &pin); if (st) return st;
    uint64_t* p_len = (uint64_t*)pin; int32_t* p_st = (int32_t*)((uint8_t*)pin + nblocks * 8);
    const rt_stream_t dsm = c->aux[0];
    for (size_t k = 0; k < nchunks; ++k) {
        RCZ_CK(c, rt_stream_wait_event(dsm, c->events[3 * k + 1]));
        size_t i = cut[k];
        const size_t e = cut[k + 1];
        // a chunk whose capacity is out of proportion to its compressed bytes (small blocks in a frame that declares 4 MiB ones) is not
        // worth copying whole: for it the host waits for the lengths and copies what was decoded
        uint64_t csum = 0, isum = 0;
        for (size_t q = i; q < e; ++q) { csum += out_cap[q]; isum += in_len[q]; }
        if (csum > 8 * isum + (1u << 20)) {
            RCZ_CK(c, rt_d2h(p_len + i, d_len + i, (e - i) * 8, dsm));
            RCZ_CK(c, rt_d2h(p_st + i, d_st + i, (e - i) * 4, dsm));
            RCZ_CK(c, rt_stream_sync(dsm));
            while (i < e) {
                if (p_len[i] == 0 || p_st[i] != RCZ_OK) { ++i; continue; }
                const uint64_t s0 = out_off[i]; uint64_t t = s0 + p_len[i];
                size_t j = i + 1;
                while (j < e && p_st[j] == RCZ_OK && (p_len[j] == 0 || out_off[j] == t)) { t += p_len[j]; ++j; }
                RCZ_CK(c, rt_d2h((uint8_t*)out_base + s0, dout + s0, (size_t)(t - s0), dsm));
                i = j;
            }
            continue;
        }
        while (i < e) {
            if (out_cap[i] == 0) { ++i; continue; }
            const uint64_t s0 = out_off[i]; uint64_t t = s0 + out_cap[i];
            size_t j = i + 1;
            while (j < e && (out_cap[j] == 0 || out_off[j] == t)) { t += out_cap[j]; ++j; }
            RCZ_CK(c, rt_d2h((uint8_t*)out_base + s0, dout + s0, (size_t)(t - s0), dsm));
            i = j;
        }
    }
    RCZ_CK(c, rt_d2h(p_len, d_len, nblocks * 8, dsm));
    RCZ_CK(c, rt_d2h(p_st, d_st, nblocks * 4, dsm));
    RCZ_CK(c, rt_stream_sync(dsm));
    for (int i = 1; i < 9; ++i) RCZ_CK(c, rt_stream_sync(c->aux[i]));
    for (size_t q = 0; q < nblocks; ++q) { out_len[q] = p_len[q]; status[q] = p_st[q]; }
    c->ev_valid = false;
    return RCZ_OK;
}

static int lz4_decode_impl(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                           void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                           uint64_t* out_len, int32_t* status, size_t nblocks, int mem_kind, void* const* peer_out_base, int npeers);

extern "C" int rcz_lz4_decode_blocks(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                                     void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                                     uint64_t* out_len, int32_t* status, size_t nblocks, int mem_kind) {
    return lz4_decode_impl(c, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, nblocks, mem_kind, nullptr, 0);
}

extern "C" int rcz_lz4_decode_blocks_gather(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                                            void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                                            uint64_t* out_len, int32_t* status, size_t nblocks, int mem_kind,
                                            void* const* peer_out_base, int npeers) {
    if (npeers < 0 || npeers > 7 || (npeers && !peer_out_base) || mem_kind == RCZ_MEM_HOST) return RCZ_E_ARG;
    return lz4_decode_impl(c, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, nblocks, mem_kind, peer_out_base, npeers);
}

static int lz4_decode_impl(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                           void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                           uint64_t* out_len, int32_t* status, size_t nblocks, int mem_kind, void* const* peer_out_base, int npeers) {
    if (!c || rcz_bad_kind(mem_kind)) return RCZ_E_ARG;
    if (nblocks == 0) return RCZ_OK;
    if (!in_base || !in_off || !in_len || !out_base || !out_off || !out_cap || !out_len || !status) return RCZ_E_ARG;
    if (nblocks > 0x7fffffffu) return RCZ_E_ARG;
    rt_set_device(c->device);
    if (!rcz_spans_ok(in_off, in_len, nblocks) || !rcz_spans_ok(out_off, out_cap, nblocks)) return RCZ_E_ARG;
    DescStager ds(c, mem_kind, nblocks);
    ds.add_in(in_off, nblocks * 8); ds.add_in(in_len, nblocks * 8); ds.add_in(out_off, nblocks * 8); ds.add_in(out_cap, nblocks * 8);
    ds.add_out(out_len, nblocks * 8); ds.add_out(status, nblocks * 4);
    const bool pipelined = mem_kind == RCZ_MEM_HOST && nblocks >= 8;
    std::vector<size_t> cut{0};
    if (pipelined) {
        const uint64_t CHUNK_BYTES = getenv("RCZ_LZ4_CHUNK_BYTES") ? strtoull(getenv("RCZ_LZ4_CHUNK_BYTES"), nullptr, 10) : (64ull << 20);   // decoded bytes per chunk
        uint64_t acc = 0;
        for (size_t i = 0; i < nblocks; ++i) { acc += out_cap[i]; if (acc >= CHUNK_BYTES && i + 1 < nblocks) { cut.push_back(i + 1); acc = 0; } }
    }
    cut.push_back(nblocks);
    Lz4Job job;
    int st = lz4_prepare(c, ds, in_len, nblocks, cut, job); if (st) return st;
    job.dev.peers.n = npeers;
    for (int p = 0; p < npeers; ++p) {
        if (!peer_out_base[p]) return RCZ_E_ARG;
        job.dev.peers.delta[p] = (long long)((const uint8_t*)peer_out_base[p] - (const uint8_t*)out_base);
        if (job.dev.peers.delta[p] & 15) return RCZ_E_ARG;                  // bulk stores: peer chunks as aligned as the local ones
    }
    if (pipelined) return lz4_host_pipelined(c, ds, job, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, nblocks);
    const uint8_t* din = (const uint8_t*)in_base; uint8_t* dout = (uint8_t*)out_base;
    if (mem_kind == RCZ_MEM_HOST) {
        st = stage_span_in(c, WS_IN, in_base, in_off, in_len, nblocks, 1, &din); if (st) return st;
        st = stage_span_out(c, WS_OUT, out_off, out_cap, nblocks, 1, &dout); if (st) return st;
    }
    st = ctx_timer_begin(c); if (st) return st;
    st = lz4_enqueue_range(c, c->stream, job, ds, 0, din, dout, ds.out_ptr<uint64_t>(0), ds.out_ptr<int32_t>(1)); if (st) return st;
    st = ctx_timer_end(c); if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) {
        std::vector<uint64_t> ok_len(out_len, out_len + nblocks);
        for (size_t i = 0; i < nblocks; ++i) if (status[i] != RCZ_OK) ok_len[i] = 0;     // nothing of a failed block is materialised
        st = unstage_span_out(c, out_base, dout, out_off, ok_len.data(), nblocks, 1); if (st) return st;
    }
    return RCZ_OK;
}
