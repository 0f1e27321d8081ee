// bwt_encode.cu — K2 (suffix sort) + K3 (L-column gather) for batches of independent BWT blocks.
//
// Replaces /root/reference/src/bwt/mod.rs:136-166 `compute_suffixes` (bucket by first byte, then
// `sort_by(|a,b| input[a..].cmp(&input[b..]))` — quadratic and worse on repetitive input) and :193-203
// `TransformIterator` (L[i] = in[SA[i]-1], in[n-1] at SA[i]==0 whose i is `origin`), as called by the stream
// encoder at bwt/mod.rs:470-475.  The suffix array is unique, so any correct sort is bit-exact; the order is the
// reference's slice order: a suffix that is a proper prefix of another sorts first (implicit end marker below 0x00).
//
// Shape on the GPU — prefix doubling (Larsson-Sadakane) over LSD radix sorts with DISCARDING, all blocks of a batch in lock step:
//   round 0   key = first 5 symbols, 9 bits each (byte+1; 0 = past the end)                                   -> 6 radix passes over n keys
//   round r   only the suffixes whose group (equal key so far) still has more than one member are sorted again:
//             key = (group head << b) | (i+h < n ? rank[i+h]+1 : 0),  h = 5 * 2^(r-1)                         -> 6 passes over the ACTIVE keys
//   after every round the sorted active keys are cut into groups again: a suffix takes the place `group head + offset inside its
//   old group`, its rank becomes the head of its new (finer) group, and a group of one is resolved for good.  Random data (BASELINE
//   config 3) is resolved after round 0 except for a handful of suffixes; hexdump text needs round 1 in full and a few per cent of
//   round 2.  (Round 1 of this file sorted all n keys in every round: 8 + 6 passes per round.)
// One radix pass = per-tile digit histogram, per-block exclusive scan, stable 256-way partition per tile (ballot-based equal-key
// ranking inside each warp, done once and kept in registers; shared-memory staging so every (tile, digit) run leaves as one contiguous
// store).  No host synchronisation is needed: finished blocks are skipped on the device (state[]), so the whole round schedule can be
// enqueued blind (DEVICE_ASYNC); the synchronous modes read one counter per round to stop early.
#include "rcz_internal.h"
#include <algorithm>

namespace bwte {

// elements per radix tile.  Measured on 256 x 4 MiB random blocks (ms per GiB): 8192 (2 CTAs / SM: 128 registers, 108 KB) 139, 4096 (3) 124,
// 3072 (4) 127, 2048 (4 CTAs / SM: 64 registers, 36 KB) 118, 1024 (5) 145 — the partition kernel is latency-bound, more resident CTAs
// hide more of it until the per-tile bookkeeping (256-entry scans, the tile histograms) takes over.
constexpr int TB = 2048;
constexpr int NT = 256;                   // threads per tile CTA
constexpr int NW = NT / 32;
constexpr int WSPAN = TB / NW;            // consecutive elements per warp
constexpr int SPAN = 1024;                // rank-assignment granularity (one warp)
constexpr unsigned NONE = 0xFFFFFFFFu;
constexpr int NSYM0 = 5;                 // symbols in a round-0 key (45 bits: 6 radix passes)
constexpr unsigned ACTIVE = 0;

struct Blk {
    unsigned long long in_off, out_off;   // bytes
    unsigned long long e0;                // element offset into K/V/R
    unsigned n, tile0, ntiles, span0, nspans, skip;
};

__device__ __forceinline__ unsigned find_blk_tile(const Blk* blks, unsigned nblocks, unsigned tile) {
    unsigned lo = 0, hi = nblocks;
    while (hi - lo > 1) { const unsigned mid = (lo + hi) >> 1; if (blks[mid].tile0 <= tile) lo = mid; else hi = mid; }
    return lo;
}
__device__ __forceinline__ unsigned find_blk_span(const Blk* blks, unsigned nblocks, unsigned span) {
    unsigned lo = 0, hi = nblocks;
    while (hi - lo > 1) { const unsigned mid = (lo + hi) >> 1; if (blks[mid].span0 <= span) lo = mid; else hi = mid; }
    return lo;
}

// ------------------------------------------------------------------------------------------ round-0 keys
__global__ void __launch_bounds__(NT)
init_keys_kernel(const uint8_t* __restrict__ in_base, const Blk* __restrict__ blks, unsigned nblocks, unsigned long long* __restrict__ K,
                 unsigned* __restrict__ V) {
    __shared__ uint8_t s[TB + 8];
    const unsigned tile = blockIdx.x, tid = threadIdx.x;
    const Blk bk = blks[find_blk_tile(blks, nblocks, tile)];
    if (bk.skip) return;
    const uint8_t* in = in_base + bk.in_off;
    const unsigned lo = (tile - bk.tile0) * TB, hi = min(bk.n, lo + TB);
    for (unsigned j = tid; j < TB + 8; j += NT) s[j] = lo + j < bk.n ? in[lo + j] : 0;
    __syncthreads();
    for (unsigned j = tid; lo + j < hi; j += NT) {
        const unsigned i = lo + j;
        unsigned long long key = 0;
#pragma unroll
        for (int k = 0; k < NSYM0; ++k) key = (key << 9) | (i + k < bk.n ? (unsigned long long)s[j + k] + 1ull : 0ull);
        K[bk.e0 + i] = key;
        V[bk.e0 + i] = i;
    }
}

// ------------------------------------------------------------------------------------------ radix pass: histogram
__global__ void __launch_bounds__(NT)
hist_kernel(const Blk* __restrict__ blks, unsigned nblocks, const unsigned* __restrict__ state, const unsigned* __restrict__ cnt,
            const unsigned long long* __restrict__ K, unsigned shift, unsigned* __restrict__ tile_hist, const unsigned* __restrict__ guard) {
    if (guard && *guard == 0) return;                                          // the previous round left no unresolved block: nothing to do
    __shared__ unsigned hsm[256];
    const unsigned tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const unsigned b = find_blk_tile(blks, nblocks, tile);
    const Blk bk = blks[b];
    if (bk.skip || state[b] != ACTIVE) return;
    const unsigned nn = cnt[b];                                                // keys being sorted in this round (n in round 0, the active ones later)
    const unsigned lo = (tile - bk.tile0) * TB, hi = min(nn, lo + TB);
    if (lo >= nn) return;
    hsm[tid] = 0;
    __syncthreads();
    const unsigned long long* k = K + bk.e0;
    for (unsigned base = lo; base < hi; base += NT) {
        const unsigned i = base + tid;
        const bool valid = i < hi;
        const unsigned d = valid ? (unsigned)(k[i] >> shift) & 255u : 0u;
        const unsigned m = warp_match_u8(d, valid);
        if (valid && (m & ((1u << lane) - 1u)) == 0) atomicAdd(&hsm[d], (unsigned)__popc(m));
    }
    __syncthreads();
    tile_hist[(size_t)tile * 256 + tid] = hsm[tid];
}

// tile_hist[tile][d] <- number of d's in earlier tiles of the block; cbase[blk][d] <- number of digits < d
__global__ void __launch_bounds__(256)
scan_kernel(const Blk* __restrict__ blks, const unsigned* __restrict__ state, const unsigned* __restrict__ cnt, unsigned* __restrict__ tile_hist,
            unsigned* __restrict__ cbase, const unsigned* __restrict__ guard) {
    if (guard && *guard == 0) return;                                          // the previous round left no unresolved block: nothing to do
    __shared__ unsigned scratch[40];
    const unsigned b = blockIdx.x, c = threadIdx.x;
    const Blk bk = blks[b];
    if (bk.skip || state[b] != ACTIVE) return;
    unsigned run = 0;
    const unsigned nt = (cnt[b] + TB - 1) / TB;                                // tiles that hold keys in this round
    for (unsigned t0 = 0; t0 < nt; t0 += 16) {                                // 16 independent loads in flight, then the running sums
        unsigned hc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) hc[k] = t0 + k < nt ? tile_hist[(size_t)(bk.tile0 + t0 + k) * 256 + c] : 0u;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (t0 + k < nt) { tile_hist[(size_t)(bk.tile0 + t0 + k) * 256 + c] = run; run += hc[k]; }
    }
    unsigned total;
    const unsigned ex = block_excl_scan_add<256>(run, scratch, &total);
    cbase[(size_t)b * 256 + c] = ex;
}

// ------------------------------------------------------------------------------------------ radix pass: stable partition
struct ScatSmem {
    unsigned long long skey[TB];
    unsigned sval[TB];
    unsigned wcnt[NW][256];
    unsigned symbase[256];
    unsigned gbase[256];
    unsigned scratch[40];
};

__global__ void __launch_bounds__(NT, 4)
scatter_kernel(const Blk* __restrict__ blks, unsigned nblocks, const unsigned* __restrict__ state, const unsigned* __restrict__ cnt,
               const unsigned long long* __restrict__ Kin, const unsigned* __restrict__ Vin, unsigned long long* __restrict__ Kout, unsigned* __restrict__ Vout, unsigned shift,
               const unsigned* __restrict__ tile_hist, const unsigned* __restrict__ cbase, const unsigned* __restrict__ guard) {
    if (guard && *guard == 0) return;                                          // the previous round left no unresolved block: nothing to do
    RCZ_DYN_SMEM(raw);
    ScatSmem& sm = *reinterpret_cast<ScatSmem*>(raw);
    const unsigned tile = blockIdx.x, tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const unsigned b = find_blk_tile(blks, nblocks, tile);
    const Blk bk = blks[b];
    if (bk.skip || state[b] != ACTIVE) return;
    const unsigned nn = cnt[b];
    const unsigned lo = (tile - bk.tile0) * TB;
    if (lo >= nn) return;
    const unsigned hi = min(nn, lo + TB), tlen = hi - lo;
    const unsigned long long* kin = Kin + bk.e0;
    const unsigned* vin = Vin + bk.e0;

    for (unsigned i = tid; i < NW * 256; i += NT) (&sm.wcnt[0][0])[i] = 0;
    __syncthreads();
    // per-warp digit counts, and for every key its rank among the equal digits of the warp's WSPAN-key span (count before this group of
    // 32 + rank inside the group): the 8 ballots are done ONCE per 32 keys and the ranks kept in registers, two per register
    unsigned rb[WSPAN / 64];
#pragma unroll
    for (unsigned g = 0; g < WSPAN / 32; ++g) {
        const unsigned i = lo + w * WSPAN + g * 32 + lane;
        const bool valid = i < hi;
        const unsigned d = valid ? (unsigned)(kin[i] >> shift) & 255u : 0u;
        const unsigned m = warp_match_u8(d, valid);
        const unsigned r = __popc(m & ((1u << lane) - 1u));
        unsigned old = 0;
        if (valid && r == 0) { old = sm.wcnt[w][d]; sm.wcnt[w][d] = old + (unsigned)__popc(m); }
        old = __shfl_sync(RCZ_FULL, old, m ? __ffs((int)m) - 1 : 0);
        const unsigned v = old + r;                                            // < 1024 + 32
        if (g & 1) rb[g >> 1] |= v << 16; else rb[g >> 1] = v;
        __syncwarp();
    }
    __syncthreads();
    unsigned tot = 0;                                                          // exclusive scan over warps, then digits
#pragma unroll
    for (int k = 0; k < NW; ++k) { const unsigned x = sm.wcnt[k][tid]; sm.wcnt[k][tid] = tot; tot += x; }
    unsigned dummy;
    const unsigned sb = block_excl_scan_add<NT>(tot, sm.scratch, &dummy);
    sm.symbase[tid] = sb;
    sm.gbase[tid] = cbase[(size_t)b * 256 + tid] + tile_hist[(size_t)tile * 256 + tid] - sb;   // dst of sorted index j: gbase[d] + j
    __syncthreads();
#pragma unroll
    for (unsigned g = 0; g < WSPAN / 32; ++g) {                                // place (keys come back from L1 / L2; no ballots left)
        const unsigned i = lo + w * WSPAN + g * 32 + lane;
        if (i < hi) {
            const unsigned long long key = kin[i];
            const unsigned d = (unsigned)(key >> shift) & 255u;
            const unsigned v = (g & 1) ? rb[g >> 1] >> 16 : rb[g >> 1] & 0xffffu;
            const unsigned p = sm.symbase[d] + sm.wcnt[w][d] + v;
            sm.skey[p] = key; sm.sval[p] = vin[i];
        }
    }
    __syncthreads();
    unsigned long long* kout = Kout + bk.e0;
    unsigned* vout = Vout + bk.e0;
    for (unsigned j = tid; j < tlen; j += NT) {
        const unsigned long long key = sm.skey[j];
        const unsigned dst = sm.gbase[(unsigned)(key >> shift) & 255u] + j;
        kout[dst] = key;
        vout[dst] = sm.sval[j];
    }
}

// ------------------------------------------------------------------------------------------ regrouping after a sort
// The sorted (active) keys of a block, cnt[b] of them, one warp per span of 1024: where subgroups (equal key) and groups (equal
// high part = old group head) begin.
__global__ void __launch_bounds__(128)
heads_kernel(const Blk* __restrict__ blks, unsigned nblocks, unsigned nspans, const unsigned* __restrict__ state, const unsigned* __restrict__ cnt,
             const unsigned long long* __restrict__ K, unsigned bits, unsigned* __restrict__ span_last, unsigned* __restrict__ span_glast, const unsigned* __restrict__ guard) {
    if (guard && *guard == 0) return;                                          // the previous round left no unresolved block: nothing to do
    const unsigned lane = threadIdx.x & 31;
    const unsigned span = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (span >= nspans) return;
    const unsigned b = find_blk_span(blks, nblocks, span);
    const Blk bk = blks[b];
    if (bk.skip || state[b] != ACTIVE) return;
    const unsigned nn = cnt[b];
    const unsigned lo = (span - bk.span0) * SPAN, hi = min(nn, lo + SPAN);
    const unsigned long long* k = K + bk.e0;
    unsigned last = NONE, glast = NONE;
    for (unsigned base = lo; base < hi; base += 32) {
        const unsigned j = base + lane;
        const unsigned long long kj = j < hi ? k[j] : 0ull, kp = (j < hi && j > 0) ? k[j - 1] : 0ull;
        const unsigned m = __ballot_sync(RCZ_FULL, j < hi && (j == 0 || kj != kp));
        const unsigned gm = __ballot_sync(RCZ_FULL, j < hi && (j == 0 || (kj >> bits) != (kp >> bits)));
        if (m) last = base + 31u - (unsigned)__clz((int)m);
        if (gm) glast = base + 31u - (unsigned)__clz((int)gm);
    }
    if (lane == 0) { span_last[span] = last; span_glast[span] = glast; }
}

// per block: span_last / span_glast <- the last (sub)group head BEFORE the span (position 0 is always a head)
__global__ void __launch_bounds__(256)
carry_kernel(const Blk* __restrict__ blks, const unsigned* __restrict__ state, unsigned* __restrict__ span_last, unsigned* __restrict__ span_glast, const unsigned* __restrict__ guard) {
    if (guard && *guard == 0) return;                                          // the previous round left no unresolved block: nothing to do
    __shared__ unsigned smax[2][256];
    const unsigned b = blockIdx.x, t = threadIdx.x;
    const Blk bk = blks[b];
    if (bk.skip || state[b] != ACTIVE) return;
    const unsigned per = (bk.nspans + 255) / 256;
    const unsigned s0 = min(bk.nspans, t * per), s1 = min(bk.nspans, s0 + per);
    unsigned mx = 0, gmx = 0;                                                  // positions are stored +1 so that 0 == none
    for (unsigned s = s0; s < s1; ++s) {
        const unsigned l = span_last[bk.span0 + s], g = span_glast[bk.span0 + s];
        if (l != NONE) mx = l + 1;
        if (g != NONE) gmx = g + 1;
    }
    smax[0][t] = mx; smax[1][t] = gmx;
    __syncthreads();
    if (t < 2) { unsigned run = 0; for (unsigned i = 0; i < 256; ++i) { const unsigned m = smax[t][i]; smax[t][i] = run; if (m) run = m; } }
    __syncthreads();
    unsigned run = smax[0][t], grun = smax[1][t];
    for (unsigned s = s0; s < s1; ++s) {
        const unsigned l = span_last[bk.span0 + s], g = span_glast[bk.span0 + s];
        span_last[bk.span0 + s] = run ? run - 1 : 0u;
        span_glast[bk.span0 + s] = grun ? grun - 1 : 0u;
        if (l != NONE) run = l + 1;
        if (g != NONE) grun = g + 1;
    }
}

// Every sorted active key p (old group head g in its high bits): new place in the suffix array = g + (p - first p of the group),
// new rank = g + (head of p's subgroup - first p of the group); a subgroup of one is resolved (flag 0).
__global__ void __launch_bounds__(128)
update_kernel(const Blk* __restrict__ blks, unsigned nblocks, unsigned nspans, const unsigned* __restrict__ state, const unsigned* __restrict__ cnt,
              const unsigned long long* __restrict__ K, const unsigned* __restrict__ V, unsigned bits, const unsigned* __restrict__ span_carry,
              const unsigned* __restrict__ span_gcarry, unsigned* __restrict__ SA, unsigned* __restrict__ R, uint8_t* __restrict__ aflag, const unsigned* __restrict__ guard) {
    if (guard && *guard == 0) return;                                          // the previous round left no unresolved block: nothing to do
    const unsigned lane = threadIdx.x & 31;
    const unsigned span = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (span >= nspans) return;
    const unsigned b = find_blk_span(blks, nblocks, span);
    const Blk bk = blks[b];
    if (bk.skip || state[b] != ACTIVE) return;
    const unsigned nn = cnt[b];
    const unsigned lo = (span - bk.span0) * SPAN, hi = min(nn, lo + SPAN);
    const unsigned long long* k = K + bk.e0;
    const unsigned* v = V + bk.e0;
    unsigned* sa = SA + bk.e0;
    unsigned* r = R + bk.e0;
    uint8_t* af = aflag + bk.e0;
    unsigned carry = span_carry[span], gcarry = span_gcarry[span];
    for (unsigned base = lo; base < hi; base += 32) {
        const unsigned j = base + lane;
        const bool in = j < hi;
        const unsigned long long kj = in ? k[j] : 0ull, kp = (in && j > 0) ? k[j - 1] : 0ull;
        const unsigned long long kn = (in && j + 1 < nn) ? k[j + 1] : ~0ull;
        const bool head = in && (j == 0 || kj != kp);
        const bool ghead = in && (j == 0 || (kj >> bits) != (kp >> bits));
        const unsigned m = __ballot_sync(RCZ_FULL, head), gm = __ballot_sync(RCZ_FULL, ghead);
        const unsigned upto = 0xFFFFFFFFu >> (31 - lane);
        const unsigned hp = (m & upto) ? base + 31u - (unsigned)__clz((int)(m & upto)) : carry;
        const unsigned gp = (gm & upto) ? base + 31u - (unsigned)__clz((int)(gm & upto)) : gcarry;
        if (in) {
            const unsigned g = (unsigned)(kj >> bits);
            const unsigned pos = g + (j - gp), val = v[j];
            sa[pos] = val;
            r[val] = g + (hp - gp);
            af[pos] = (head && (j + 1 >= nn || kn != kj)) ? 0 : 1;            // alone in its subgroup: resolved
        }
        if (m) carry = base + 31u - (unsigned)__clz((int)m);
        if (gm) gcarry = base + 31u - (unsigned)__clz((int)gm);
    }
}

// active suffixes per span of the suffix array (all n positions)
__global__ void __launch_bounds__(128)
count_active_kernel(const Blk* __restrict__ blks, unsigned nblocks, unsigned nspans, const unsigned* __restrict__ state,
                    const uint8_t* __restrict__ aflag, unsigned* __restrict__ span_act, const unsigned* __restrict__ guard) {
    if (guard && *guard == 0) return;                                          // the previous round left no unresolved block: nothing to do
    const unsigned lane = threadIdx.x & 31;
    const unsigned span = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (span >= nspans) return;
    const unsigned b = find_blk_span(blks, nblocks, span);
    const Blk bk = blks[b];
    if (bk.skip || state[b] != ACTIVE) return;
    const unsigned lo = (span - bk.span0) * SPAN, hi = min(bk.n, lo + SPAN);
    const uint8_t* af = aflag + bk.e0;
    unsigned c = 0;
    for (unsigned j = lo + lane; j < hi; j += 32) c += af[j] ? 1u : 0u;
    c = warp_reduce_add(c);
    if (lane == 0) span_act[span] = c;
}

// per block: span_act <- active suffixes before the span; cnt_next[b] = their total; none left -> the block is finished in this round
__global__ void __launch_bounds__(256)
active_scan_kernel(const Blk* __restrict__ blks, unsigned* __restrict__ state, unsigned* __restrict__ span_act, unsigned* __restrict__ cnt_next,
                   unsigned round, unsigned* __restrict__ nactive, const unsigned* __restrict__ guard) {
    if (guard && *guard == 0) return;                                          // the previous round left no unresolved block: nothing to do
    __shared__ unsigned ssum[256];
    const unsigned b = blockIdx.x, t = threadIdx.x;
    const Blk bk = blks[b];
    if (bk.skip || state[b] != ACTIVE) return;
    const unsigned per = (bk.nspans + 255) / 256;
    const unsigned s0 = min(bk.nspans, t * per), s1 = min(bk.nspans, s0 + per);
    unsigned sum = 0;
    for (unsigned s = s0; s < s1; ++s) sum += span_act[bk.span0 + s];
    ssum[t] = sum;
    __syncthreads();
    if (t == 0) {
        unsigned run = 0;
        for (unsigned i = 0; i < 256; ++i) { const unsigned x = ssum[i]; ssum[i] = run; run += x; }
        cnt_next[b] = run;
        if (run == 0) state[b] = round + 1; else atomicAdd(nactive, 1u);
    }
    __syncthreads();
    unsigned run = ssum[t];
    for (unsigned s = s0; s < s1; ++s) { const unsigned x = span_act[bk.span0 + s]; span_act[bk.span0 + s] = run; run += x; }
}

// next round's keys: the active suffixes in suffix-array order, key = (rank (= group head) << bits) | (rank of the suffix h further + 1)
__global__ void __launch_bounds__(128)
compact_kernel(const Blk* __restrict__ blks, unsigned nblocks, unsigned nspans, const unsigned* __restrict__ state, const uint8_t* __restrict__ aflag,
               const unsigned* __restrict__ span_abase, const unsigned* __restrict__ SA, const unsigned* __restrict__ R, unsigned h, unsigned bits,
               unsigned long long* __restrict__ K, unsigned* __restrict__ V, const unsigned* __restrict__ guard) {
    if (guard && *guard == 0) return;                                          // the previous round left no unresolved block: nothing to do
    const unsigned lane = threadIdx.x & 31;
    const unsigned span = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (span >= nspans) return;
    const unsigned b = find_blk_span(blks, nblocks, span);
    const Blk bk = blks[b];
    if (bk.skip || state[b] != ACTIVE) return;
    const unsigned lo = (span - bk.span0) * SPAN, hi = min(bk.n, lo + SPAN);
    const uint8_t* af = aflag + bk.e0;
    const unsigned* sa = SA + bk.e0;
    const unsigned* r = R + bk.e0;
    unsigned long long* k = K + bk.e0;
    unsigned* v = V + bk.e0;
    unsigned p0 = span_abase[span];
    for (unsigned base = lo; base < hi; base += 32) {
        const unsigned j = base + lane;
        const bool act = j < hi && af[j] != 0;
        const unsigned m = __ballot_sync(RCZ_FULL, act);
        if (act) {
            const unsigned p = p0 + __popc(m & ((1u << lane) - 1u));
            const unsigned i = sa[j];
            const unsigned long long r2 = (i + h < bk.n && i + h >= i) ? (unsigned long long)r[i + h] + 1ull : 0ull;
            k[p] = ((unsigned long long)r[i] << bits) | r2;
            v[p] = i;
        }
        p0 += __popc(m);
    }
}

// L column + origin for the blocks that finished in this round (bwt/mod.rs:193-203)
__global__ void __launch_bounds__(NT)
gather_kernel(const uint8_t* __restrict__ in_base, const Blk* __restrict__ blks, unsigned nblocks, const unsigned* __restrict__ state,
              unsigned round, const unsigned* __restrict__ V, uint8_t* __restrict__ out_base, uint32_t* __restrict__ origin,
              int32_t* __restrict__ status, const unsigned* __restrict__ guard) {
    if (guard && *guard == 0) return;                                          // the previous round left no unresolved block: nothing to do
    const unsigned tile = blockIdx.x, tid = threadIdx.x;
    const unsigned b = find_blk_tile(blks, nblocks, tile);
    const Blk bk = blks[b];
    if (bk.skip || state[b] != round + 1) return;
    const uint8_t* in = in_base + bk.in_off;
    uint8_t* out = out_base + bk.out_off;
    const unsigned* v = V + bk.e0;
    const unsigned lo = (tile - bk.tile0) * TB, hi = min(bk.n, lo + TB);
    for (unsigned j = lo + tid; j < hi; j += NT) {
        const unsigned p = v[j];
        if (p == 0) { out[j] = in[bk.n - 1]; origin[b] = j; status[b] = RCZ_OK; }
        else out[j] = in[p - 1];
    }
}

__global__ void init_state_kernel(const Blk* __restrict__ blks, unsigned nblocks, unsigned* __restrict__ state, unsigned* __restrict__ nactive,
                                  unsigned nrounds, uint32_t* __restrict__ origin, int32_t* __restrict__ status,
                                  const int32_t* __restrict__ host_status, unsigned* __restrict__ cnt0) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nrounds) nactive[i] = 0;
    if (i < nblocks) {
        state[i] = ACTIVE;
        cnt0[i] = blks[i].n;
        origin[i] = 0;
        status[i] = blks[i].skip ? host_status[i] : RCZ_E_CUDA;               // overwritten by gather_kernel when the block finishes
    }
}

}  // namespace bwte

extern "C" int rcz_bwt_encode_blocks(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* n_arr, void* out_base,
                                     const uint64_t* out_off, uint32_t* origin, int32_t* status, size_t nblocks, int mem_kind) {
    using namespace bwte;
    if (!c || rcz_bad_kind(mem_kind)) return RCZ_E_ARG;
    if (nblocks == 0) return RCZ_OK;
    if (!in_base || !in_off || !n_arr || !out_base || !out_off || !origin || !status || nblocks > 0x3fffffu) return RCZ_E_ARG;
    for (size_t i = 0; i < nblocks; ++i) if (in_off[i] > (1ull << 62) || out_off[i] > (1ull << 62)) return RCZ_E_ARG;
    rt_set_device(c->device);
    if (mem_kind == RCZ_MEM_HOST && rcz_spans_ok(in_off, n_arr, nblocks)) {    // big host batches: pipelined chunks of 256 MiB
        bool handled = false;
        const int st = host_chunked(c, nblocks, 256ull << 20, in_base, in_off, n_arr, 1, out_base, out_off, n_arr, 1,
            [&](size_t b0, size_t nb, const uint8_t* din, uint8_t* dout) {
                return rcz_bwt_encode_blocks(c, din, in_off + b0, n_arr + b0, dout, out_off + b0, origin + b0, status + b0, nb, RCZ_MEM_DEVICE);
            },
            [&](size_t i) { return status[i] == RCZ_OK ? n_arr[i] : 0; }, &handled);
        if (st || handled) return st;
    }

    // ---- geometry; big batches are cut into groups so that the sort workspace (28 B / symbol) stays bounded
    constexpr unsigned long long GROUP_ELEMS = 1ull << 28;
    std::vector<Blk> blks(nblocks);
    std::vector<int32_t> hstatus(nblocks, 0);
    struct Group { size_t b0, b1; unsigned long long elems; unsigned ntiles, nspans, nmax; };
    std::vector<Group> groups;
    Group cur{0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < nblocks; ++i) {
        Blk& b = blks[i];
        memset(&b, 0, sizeof b);
        b.in_off = in_off[i]; b.out_off = out_off[i];
        const unsigned long long n = n_arr[i];
        const bool bad = n == 0 || n >= (1ull << 31);
        if (!bad && cur.b1 > cur.b0 && (cur.elems + n > GROUP_ELEMS || cur.ntiles > 0x3fffffffu - (unsigned)(n / TB + 1))) {
            groups.push_back(cur);
            cur = Group{i, i, 0, 0, 0, 0};
        }
        b.e0 = cur.elems; b.tile0 = cur.ntiles; b.span0 = cur.nspans;
        if (bad) { b.skip = 1; hstatus[i] = n == 0 ? RCZ_E_MALFORMED : RCZ_E_ARG; }   // n == 0: get_origin().unwrap() on None (bwt/mod.rs:186-188)
        else {
            b.n = (unsigned)n;
            b.ntiles = (b.n + TB - 1) / TB; b.nspans = (b.n + SPAN - 1) / SPAN;
            cur.elems += (n + 63) & ~63ull; cur.ntiles += b.ntiles; cur.nspans += b.nspans;
            cur.nmax = std::max(cur.nmax, b.n);
        }
        cur.b1 = i + 1;
    }
    groups.push_back(cur);
    unsigned long long max_elems = 0; unsigned max_tiles = 0, max_spans = 0; size_t max_blocks = 0;
    for (auto& g : groups) {
        max_elems = std::max(max_elems, g.elems); max_tiles = std::max(max_tiles, g.ntiles); max_spans = std::max(max_spans, g.nspans);
        max_blocks = std::max(max_blocks, g.b1 - g.b0);
    }

    DescStager ds(c, mem_kind, nblocks);
    const size_t i_blk = ds.add_in(blks.data(), nblocks * sizeof(Blk));
    const size_t i_hst = ds.add_in(hstatus.data(), nblocks * 4);
    const size_t o_org = ds.add_out(origin, nblocks * 4);
    const size_t o_st = ds.add_out(status, nblocks * 4);
    int st = ds.upload(); if (st) return st;
    const uint8_t* din = (const uint8_t*)in_base; uint8_t* dout = (uint8_t*)out_base;
    if (mem_kind == RCZ_MEM_HOST) {
        st = stage_span_in(c, WS_IN, in_base, in_off, n_arr, nblocks, 1, &din); if (st) return st;
        st = stage_span_out(c, WS_OUT, out_off, n_arr, nblocks, 1, &dout); if (st) return st;
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    void *wK, *wV, *wR, *wM;
    st = ctx_ws(c, WS_A, al((size_t)max_elems * 8) * 2 + 256, &wK); if (st) return st;
    st = ctx_ws(c, WS_B, al((size_t)max_elems * 4) * 2 + 256, &wV); if (st) return st;
    st = ctx_ws(c, WS_C, al((size_t)max_elems * 4) * 2 + al((size_t)max_elems) + 256, &wR); if (st) return st;   // ranks | suffix array | active flags
    const size_t sz_hist = al((size_t)max_tiles * 256 * 4), sz_cb = al(max_blocks * 256 * 4), sz_sp = al((size_t)max_spans * 4), sz_state = al(max_blocks * 4);
    st = ctx_ws(c, WS_D, sz_hist + sz_cb + 3 * sz_sp + 3 * sz_state + 1024, &wM); if (st) return st;
    unsigned long long* K[2] = {(unsigned long long*)wK, (unsigned long long*)((uint8_t*)wK + al((size_t)max_elems * 8))};
    unsigned* V[2] = {(unsigned*)wV, (unsigned*)((uint8_t*)wV + al((size_t)max_elems * 4))};
    unsigned* R = (unsigned*)wR;
    unsigned* SA = (unsigned*)((uint8_t*)wR + al((size_t)max_elems * 4));
    uint8_t* aflag = (uint8_t*)wR + 2 * al((size_t)max_elems * 4);
    uint8_t* m = (uint8_t*)wM;
    unsigned* tile_hist = (unsigned*)m; m += sz_hist;
    unsigned* cbase = (unsigned*)m; m += sz_cb;
    unsigned* span_last = (unsigned*)m; m += sz_sp;
    unsigned* span_glast = (unsigned*)m; m += sz_sp;
    unsigned* span_act = (unsigned*)m; m += sz_sp;
    unsigned* state = (unsigned*)m; m += sz_state;
    unsigned* cntbuf[2] = {(unsigned*)m, (unsigned*)(m + sz_state)}; m += 2 * sz_state;
    unsigned* nactive = (unsigned*)m;                                          // one counter per round (<= 64)

    RCZ_CK(c, RCZ_KERNEL_SMEM_OPTIN(scatter_kernel, sizeof(ScatSmem)));
    st = ctx_timer_begin(c); if (st) return st;
    for (auto& g : groups) {
        const unsigned nb = (unsigned)(g.b1 - g.b0);
        const Blk* dblk = ds.in_ptr<Blk>(i_blk) + g.b0;
        uint32_t* d_org = ds.out_ptr<uint32_t>(o_org) + g.b0;
        int32_t* d_st = ds.out_ptr<int32_t>(o_st) + g.b0;
        unsigned bits = 1; while ((1ull << bits) <= g.nmax) ++bits;            // rank2 takes values 0..n
        unsigned nrounds = 1; while ((unsigned long long)NSYM0 << (nrounds - 1) < g.nmax) ++nrounds;   // after round r ranks cover NSYM0 * 2^r symbols
        if (nrounds > 60) nrounds = 60;
        RCZ_KLAUNCH(c, init_state_kernel, (std::max(nb, 64u) + 255) / 256, 256, 0, dblk, nb, state, nactive, 64u, d_org, d_st,
                    ds.in_ptr<int32_t>(i_hst) + g.b0, cntbuf[0]);
        if (g.ntiles == 0) continue;
        const unsigned sgrid = (g.nspans + 3) / 4;
        for (unsigned round = 0; round < nrounds; ++round) {
            const unsigned* cnt = cntbuf[round & 1];
            const unsigned* guard = round ? nactive + (round - 1) : (const unsigned*)nullptr;   // blocks still unresolved after the previous round
            unsigned* cnt_next = cntbuf[(round & 1) ^ 1];
            unsigned npass, gbits;
            if (round == 0) {
                RCZ_KLAUNCH(c, init_keys_kernel, g.ntiles, NT, 0, din, dblk, nb, K[0], V[0]);
                npass = (9 * NSYM0 + 7) / 8; npass += npass & 1;              // even: the sorted data ends up in buffer 0
                gbits = 63;                                                    // round 0: one group (the whole block), head 0
            } else {
                RCZ_KLAUNCH(c, compact_kernel, sgrid, 128, 0, dblk, nb, g.nspans, state, aflag, span_act, SA, R, (unsigned)((unsigned)NSYM0 << (round - 1)), bits, K[0], V[0], guard);
                npass = (2 * bits + 7) / 8; npass += npass & 1;
                gbits = bits;
            }
            for (unsigned p = 0; p < npass; ++p) {
                const unsigned a = p & 1;
                RCZ_KLAUNCH(c, hist_kernel, g.ntiles, NT, 0, dblk, nb, state, cnt, K[a], p * 8, tile_hist, guard);
                RCZ_KLAUNCH(c, scan_kernel, nb, 256, 0, dblk, state, cnt, tile_hist, cbase, guard);
                RCZ_KLAUNCH(c, scatter_kernel, g.ntiles, NT, sizeof(ScatSmem), dblk, nb, state, cnt, K[a], V[a], K[a ^ 1], V[a ^ 1], p * 8, tile_hist, cbase, guard);
            }
            RCZ_KLAUNCH(c, heads_kernel, sgrid, 128, 0, dblk, nb, g.nspans, state, cnt, K[0], gbits, span_last, span_glast, guard);
            RCZ_KLAUNCH(c, carry_kernel, nb, 256, 0, dblk, state, span_last, span_glast, guard);
            RCZ_KLAUNCH(c, update_kernel, sgrid, 128, 0, dblk, nb, g.nspans, state, cnt, K[0], V[0], gbits, span_last, span_glast, SA, R, aflag, guard);
            RCZ_KLAUNCH(c, count_active_kernel, sgrid, 128, 0, dblk, nb, g.nspans, state, aflag, span_act, guard);
            RCZ_KLAUNCH(c, active_scan_kernel, nb, 256, 0, dblk, state, span_act, cnt_next, round, nactive + round, guard);
            RCZ_KLAUNCH(c, gather_kernel, g.ntiles, NT, 0, din, dblk, nb, state, round, SA, dout, d_org, d_st, guard);
            if (round + 1 == nrounds) break;
            if (mem_kind != RCZ_MEM_DEVICE_ASYNC || (c->nest && c->nest_may_sync)) {   // stop as soon as every block is finished
                unsigned left = 0;
                RCZ_CK(c, rt_d2h(&left, nactive + round, 4, c->stream));
                RCZ_CK(c, rt_stream_sync(c->stream));
                if (left == 0) break;
            }
        }
    }
    st = ctx_timer_end(c); if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) {
        std::vector<uint64_t> lens(nblocks);
        for (size_t i = 0; i < nblocks; ++i) lens[i] = status[i] == RCZ_OK ? n_arr[i] : 0;
        st = unstage_span_out(c, out_base, dout, out_off, lens.data(), nblocks, 1); if (st) return st;
    }
    return RCZ_OK;
}
