// dc.cu — K9/K10: distance coding of BWT blocks.
//
// Encode replaces /root/reference/src/bwt/dc.rs:110-149 `encode` + :88-104 `EncodeIterator` (+ MTF::encode,
// bwt/mtf.rs:63-79).  The reference walks the block left to right with a move-to-front list; what it emits is,
// for every run end p (p == n-1 or in[p+1] != in[p]) in position order,
//        d[p] = q - p - r - 1,   q = next position > p holding in[p] (n if none),
//                                r = number of distinct symbols in (p, q)   (== the MTF rank, dc.rs:131-136,
//                                                                             and the final sweep dc.rs:139-144)
// and init[c] = first position of c (n when absent).  That form has no carried list, so it is computed right to
// left from a per-symbol "next occurrence" table nxt[256]: r = #{s != in[p] : nxt[s] < q}.  Blocks are cut into
// 16 KiB segments (one warp each); a first pass records the first occurrence of every symbol per segment and the
// number of run ends, a per-block suffix-min / prefix-sum turns those into each segment's starting nxt[] table
// and output offset, and the last pass emits the distances.
//
// Decode replaces dc.rs:162-233 `decode` fed by `decode_simple`'s closure (dc.rs:236-252).  It is inherently
// serial in the run index (every distance re-sorts the symbol list): one dependency chain per block, and what counts is
// the length of that chain per run end.  One warp decodes one block with the whole (next position, symbol) list in
// registers, 1 / 2 / 4 / 8 consecutive ranks per lane by alphabet size (dc_decode_list<R>), so the slide of dc.rs:215-218
// costs register moves + one shuffle at any rank and there is no slow path for deep re-entries; a scalar loop with the
// top four ranks in registers was measured 3x slower, and a version that kept ranks >= 32 in shared memory measured the
// same as this one (~450 cycles per run end on a B200: ~110 dependent SASS instructions, one warp per scheduler).
#include "rcz_internal.h"
#include <algorithm>

namespace dck {

constexpr unsigned SEG = 16384;          // bytes per encode segment
constexpr int NT = 128;                  // 4 warps per CTA
constexpr int WPB = NT / 32;

struct Blk {
    unsigned long long in_off;           // bytes
    unsigned long long out_off, out_cap; // u32 elements
    unsigned n, seg0, nseg, skip;
};

// ---------------------------------------------------------------------------------------------- encode, pass 1
// first[seg][c] = first position of c inside the segment (n if none); nruns[seg] = run ends inside the segment
__global__ void __launch_bounds__(NT)
dc_first_kernel(const uint8_t* __restrict__ in_base, const Blk* __restrict__ blks, const unsigned* __restrict__ seg2blk,
                unsigned nsegs, unsigned* __restrict__ first, unsigned* __restrict__ nruns) {
    __shared__ unsigned sfirst[WPB][256];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned seg = blockIdx.x * WPB + w;
    if (seg >= nsegs) return;
    const Blk bk = blks[seg2blk[seg]];
    const uint8_t* in = in_base + bk.in_off;
    const unsigned n = bk.n, lo = (seg - bk.seg0) * SEG, hi = min(n, lo + SEG);
    for (unsigned i = lane; i < 256; i += 32) sfirst[w][i] = n;
    __syncwarp();
    unsigned runs = 0;
    for (unsigned p = lo + lane; p < hi; p += 32) {
        const unsigned c = in[p];
        if (p == lo || in[p - 1] != c) atomicMin(&sfirst[w][c], p);          // only run heads can be first occurrences
        if (p + 1 == n || in[p + 1] != c) ++runs;
    }
    runs = warp_reduce_add(runs);
    __syncwarp();
    for (unsigned i = lane; i < 256; i += 32) first[(size_t)seg * 256 + i] = sfirst[w][i];
    if (lane == 0) nruns[seg] = runs;
}

// ---------------------------------------------------------------------------------------------- encode, pass 2
// per block: first[seg][c] <- first occurrence of c at or after the END of seg (suffix min over later segments),
// nruns[seg] <- number of run ends before seg; writes init[256] (dc.rs:153-159) and out_len.
__global__ void __launch_bounds__(256)
dc_scan_kernel(const Blk* __restrict__ blks, unsigned* __restrict__ first, unsigned* __restrict__ nruns, uint32_t* __restrict__ out_base,
               uint64_t* __restrict__ out_len, int32_t* __restrict__ status, const int32_t* __restrict__ host_status) {
    const unsigned b = blockIdx.x, c = threadIdx.x;
    const Blk bk = blks[b];
    if (bk.skip) { if (c == 0) { out_len[b] = 0; status[b] = host_status[b]; } return; }
    unsigned run = bk.n;
    for (unsigned s = bk.nseg; s-- > 0;) {
        const size_t idx = (size_t)(bk.seg0 + s) * 256 + c;
        const unsigned f = first[idx];
        first[idx] = run;
        run = min(run, f);
    }
    __shared__ unsigned total;
    if (c == 0) {
        unsigned acc = 0;
        for (unsigned s = 0; s < bk.nseg; ++s) { const unsigned r = nruns[bk.seg0 + s]; nruns[bk.seg0 + s] = acc; acc += r; }
        total = acc;
    }
    __syncthreads();
    const bool fits = 256ull + total <= bk.out_cap;
    if (fits) out_base[bk.out_off + c] = run;                                // init[c]: first position or n
    if (c == 0) { out_len[b] = 256ull + total; status[b] = fits ? RCZ_OK : RCZ_E_OUTPUT_FULL; }
}

// ---------------------------------------------------------------------------------------------- encode, pass 3
__global__ void __launch_bounds__(NT)
dc_emit_kernel(const uint8_t* __restrict__ in_base, const Blk* __restrict__ blks, const unsigned* __restrict__ seg2blk, unsigned nsegs,
               const unsigned* __restrict__ first, const unsigned* __restrict__ nruns, uint32_t* __restrict__ out_base,
               const int32_t* __restrict__ status) {
    __align__(16) __shared__ unsigned snxt[WPB][256];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned seg = blockIdx.x * WPB + w;
    if (seg >= nsegs) return;
    const unsigned bi = seg2blk[seg];
    const Blk bk = blks[bi];
    if (bk.skip || status[bi] != RCZ_OK) return;
    const uint8_t* in = in_base + bk.in_off;
    const unsigned n = bk.n, lo = (seg - bk.seg0) * SEG, hi = min(n, lo + SEG);
    unsigned* nxt = snxt[w];
    for (unsigned i = lane; i < 256; i += 32) nxt[i] = first[(size_t)seg * 256 + i];
    __syncwarp();
    // distances of this segment's run ends go to dst[0 .. runs_here); we walk right to left, so count first
    uint32_t* dst = out_base + bk.out_off + 256 + nruns[seg];
    unsigned runs_here = 0;
    for (unsigned p = lo + lane; p < hi; p += 32) runs_here += (p + 1 == n || in[p + 1] != in[p]) ? 1u : 0u;
    runs_here = warp_reduce_add(runs_here);
    unsigned idx = runs_here;                                                 // index after the next run end to emit
    for (unsigned base = (hi - 1) & ~31u;; base -= 32) {                      // 32 positions per step, high to low
        const unsigned p = base + lane;
        const bool valid = p >= lo && p < hi;
        const unsigned c = valid ? (unsigned)in[p] : 0u;
        const unsigned cn = (valid && p + 1 < n) ? (unsigned)in[p + 1] : 256u;
        unsigned m = __ballot_sync(RCZ_FULL, valid && c != cn);
        while (m) {
            const int l = 31 - __clz((int)m);
            m &= ~(1u << l);
            const unsigned pp = base + (unsigned)l;
            const unsigned cc = __shfl_sync(RCZ_FULL, c, l), cnn = __shfl_sync(RCZ_FULL, cn, l);
            if (lane == 0 && cnn < 256u) nxt[cnn] = pp + 1;                   // the run to the right starts at pp+1
            __syncwarp();
            const unsigned q = nxt[cc];
            const uint4 a = *reinterpret_cast<const uint4*>(&nxt[lane * 8]);
            const uint4 b = *reinterpret_cast<const uint4*>(&nxt[lane * 8 + 4]);
            unsigned r = (a.x < q) + (a.y < q) + (a.z < q) + (a.w < q) + (b.x < q) + (b.y < q) + (b.z < q) + (b.w < q);
            r = __reduce_add_sync(RCZ_FULL, r);                               // nxt[cc] == q is never < q
            --idx;
            if (lane == 0) dst[idx] = q - pp - r - 1;                         // dc.rs:136 / :143
            __syncwarp();
        }
        if (base <= lo) break;
    }
}

// ---------------------------------------------------------------------------------------------- decode
struct DecSmem { unsigned nx[WPB][256]; uint8_t sy[WPB][256]; };              // the (next position, symbol) list of every warp's block, by rank

// bytes [k0, stop) step 32 of a run longer than one warp store (rare in BWT columns of text; out of line so that the decode loop's
// common path has no taken branch)
__device__ __noinline__ void dc_fill_long(uint8_t* __restrict__ out, unsigned k0, unsigned stop, unsigned sym) {
    for (unsigned k = k0; k < stop; k += 32) out[k] = (uint8_t)sym;
}

// dc.rs:199-229.  Every distance re-sorts the (next position, symbol) list, so the loop is one dependency chain per block and what
// counts is the latency from one list state to the next, not the number of instructions (with one chain per block the GPU is
// almost empty).  The whole list lives in registers, R consecutive ranks per lane (lane l holds ranks R*l .. R*l + R-1, 32-bit
// positions, n < 2^31), whatever the alphabet: the slide of dc.rs:215-218 is a register move inside the lane plus ONE shuffle for
// the entry that crosses lanes, the rank comes from R independent ballots (one per register level), and there is no slow path
// for deep re-entries (a BWT column of hexdump text has 96 symbols and re-enters at rank >= 32
// in 18 % of the steps, at rank 16 on average).  The loop body is straight-line: the error tests only set a flag (and clamp, so
// that nothing goes out of bounds) and are looked at when the loop ends.  The distances arrive 32 at a time, the next group in
// flight while one is used.
template <int R>
__device__ __forceinline__ int dc_decode_list(const unsigned* tnx, const uint8_t* tsy, unsigned A, unsigned N, const uint32_t* __restrict__ dist,
                                              unsigned long long ndist, uint8_t* __restrict__ out, unsigned lane) {
    unsigned nx[R], sy[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const unsigned q = R * lane + i;
        nx[i] = q < A ? tnx[q] : 0xFFFFFFFFu;                                 // (sentinel: never passed, future + q < 2^32)
        sy[i] = q < A ? (unsigned)tsy[q] : 0u;
    }
    unsigned i = 0, di = 0;
    const unsigned ND = ndist > 0x80000000ull ? 0x80000000u : (unsigned)ndist;   // (a block has at most n < 2^31 run ends)
    unsigned dreg = lane < ND ? __ldg(dist + lane) : 0u;                      // distances [32g, 32g + 32) of the current group, one per lane
    unsigned dnxt = 32u + lane < ND ? __ldg(dist + 32 + lane) : 0u;           // the next group
    const uint32_t* dp = dist + 64 + lane;                                    // this lane's distance of the group after that
    bool bad = false;
    while (i < N && di < ND && !bad) {
        // ranks 0 and 1: the current symbol and the end of its run
        const unsigned sym = __shfl_sync(RCZ_FULL, sy[0], 0);
        unsigned stop = R > 1 ? __shfl_sync(RCZ_FULL, nx[R > 1 ? 1 : 0], 0) : __shfl_sync(RCZ_FULL, nx[0], 1);
        const unsigned cross_nx = __shfl_down_sync(RCZ_FULL, nx[0], 1), cross_sy = __shfl_down_sync(RCZ_FULL, sy[0], 1);   // rank R*(l+1), for the slide
        unsigned d = __shfl_sync(RCZ_FULL, dreg, (int)(di & 31u));
        bad = stop > N;                                                       // output[i] index panic
        stop = stop > N ? N : stop;
        bad |= d > N - stop;                                                  // dc.rs:213 assert!(future <= n)
        d = d > N - stop ? N - stop : d;
        // the run [i, stop): almost always a few bytes (one predicated store); the rest of a long one out of line
        const unsigned run = stop > i ? stop - i : 0u;
        if (lane < run) out[i + lane] = (uint8_t)sym;
        if (run > 32u) dc_fill_long(out, i + 32u + lane, stop, sym);
        i = stop > i ? stop : i;
        ++di;
        if ((di & 31u) == 0) {                                                // next group of 32 distances (short: stays predicated)
            dreg = dnxt;
            dnxt = di + 32u + lane < ND ? __ldg(dp) : 0u;
            dp += 32;
        }
        const unsigned future = stop + d;
        // rank = 1 + #{leading q in [1, A) : future + q > next(list[q])}  (dc.rs:215-218; the reference stops at the first rank that
        // fails the test: LEADING hits, not all hits).  Rank 0 (the symbol itself) counts as passed.  One ballot per register level,
        // all independent; level r's first failing lane f gives the candidate rank R*f + r, the smallest candidate is the rank
        // (no failing entry at all: A == 32 R, rank = A).
        unsigned rank = 32u * R;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const unsigned q = R * lane + r;
            const unsigned pass = __ballot_sync(RCZ_FULL, q == 0 || future + q > nx[r]);
            const unsigned f = (unsigned)__popc(pass & ~(pass + 1u));         // trailing ones: lanes before the first failing one
            const unsigned cand = pass == RCZ_FULL ? 32u * R : R * f + r;
            rank = cand < rank ? cand : rank;
        }
        // list[q] = list[q + 1] for q < rank - 1; list[rank - 1] = (future + rank - 1, sym)
        const unsigned at = rank - 1u;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const unsigned q = R * lane + r;
            const unsigned un = r + 1 < R ? nx[r + 1 < R ? r + 1 : r] : cross_nx, us = r + 1 < R ? sy[r + 1 < R ? r + 1 : r] : cross_sy;
            nx[r] = q < at ? un : q == at ? future + at : nx[r];
            sy[r] = q < at ? us : q == at ? sym : sy[r];
        }
    }
    // a flagged step is where the reference panics (its list update is then meaningless and is not looked at); running out of
    // distances before the block is full is dc.rs:245-246
    int err = 0;
    if (bad) err = RCZ_E_MALFORMED;
    else if (i < N) {                                                         // out of distances: the reference still looks at the next run first
        const unsigned stop = R > 1 ? __shfl_sync(RCZ_FULL, nx[R > 1 ? 1 : 0], 0) : __shfl_sync(RCZ_FULL, nx[0], 1);
        const unsigned sym = __shfl_sync(RCZ_FULL, sy[0], 0);
        if (stop > N) err = RCZ_E_MALFORMED;
        else { for (unsigned k = i + lane; k < stop; k += 32) out[k] = (uint8_t)sym; err = RCZ_E_UNEXPECTED_EOF; }
    }
    if (!err) {                                                               // dc.rs:230-231 assert_eq!
        bool off = false;
#pragma unroll
        for (int r = 0; r < R; ++r) off |= R * lane + r < A && (nx[r] < N || nx[r] >= N + A);
        if (__any_sync(RCZ_FULL, off) || i != N) err = RCZ_E_MALFORMED;
    }
    return err;
}

// The same loop for lists whose positions are pairwise distinct (every well-formed stream; ties can only come from a damaged
// init[] table and take dc_decode_list above).  A list without ties is strictly increasing in the rank and stays so, which makes
// the test of dc.rs:215-218 monotone (passed ... passed, failed ... failed) and the whole update LOCAL to an entry and its
// upper neighbour: with c[q] = "rank q is passed" = future + q > next(old[q])  (c[0] and c[1] always hold: the popped symbol's own
// run ends where rank 1 begins),
//        new[q] = old[q + 1]             if c[q + 1]
//               = (future + q, sym)      else if c[q]
//               = old[q]                 otherwise
// — no ballot, no rank.  What is left of the step-to-step dependency is the front of the list, kept in uniform registers:
//        sym' = symbol(old[1]),     new[1] = old[2] if c[2] else (future + 1, sym),
// and old[2] is broadcast one step ahead, so a step's critical path is two integer instructions instead of the
// shuffle -> compare -> ballot -> count -> shuffle -> select chain; what is left is the issue rate of a single warp, so the loop
// is written for instruction count: an entry is ONE 32-bit key (position << 8 | symbol, hence n + 256 < 2^24 — every BWT block this
// library accepts), "passed" is key < (future + q) << 8, and an entry costs one add, one compare and two selects per step.
template <int R>
__device__ __forceinline__ int dc_decode_sorted(const unsigned* tnx, const uint8_t* tsy, unsigned A, unsigned N, const uint32_t* __restrict__ dist,
                                                unsigned long long ndist, uint8_t* __restrict__ out, unsigned lane) {
    unsigned key[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const unsigned q = R * lane + r;
        key[r] = q < A ? (tnx[q] << 8) | (unsigned)tsy[q] : 0xFFFFFFFFu;      // (sentinel: never passed)
    }
    // ranks 0, 1, 2 as every lane sees them
    unsigned sym = __shfl_sync(RCZ_FULL, key[0], 0) & 0xFFu;
    unsigned e1 = R > 1 ? __shfl_sync(RCZ_FULL, key[R > 1 ? 1 : 0], 0) : __shfl_sync(RCZ_FULL, key[0], 1);
    unsigned e2 = R > 2 ? __shfl_sync(RCZ_FULL, key[R > 2 ? 2 : 0], 0) : __shfl_sync(RCZ_FULL, key[0], 2 / R);
    unsigned i = 0;
    const unsigned ND = ndist > 0x80000000ull ? 0x80000000u : (unsigned)ndist;   // (a block has at most n < 2^31 run ends)
    const unsigned last = ND ? ND - 1u : 0u;
    // distances in groups of 32, one per lane, in two registers used in turn: while one group is consumed the group after the next is
    // already on its way into the other register, which nothing reads for a whole group (with one "next" register rotated into a
    // "current" one, ptxas copies the loaded value right behind the load and the lone warp waits out the HBM latency per group)
    unsigned dA = ND ? __ldg(dist + (lane < ND ? lane : last)) : 0u;
    unsigned dB = ND ? __ldg(dist + (32u + lane < ND ? 32u + lane : last)) : 0u;
    uint8_t* op = out + lane;                                                 // this lane's byte of the current run: out + i + lane
#ifndef RCZ_EMU
    asm volatile("" : "+l"(op));                                              // keep the pointer in registers (rebuilt from the parameter bank it costs a
#endif                                                                        //  constant load per step, and a lone warp waits for every one of them)
    bool bad = false;
    // A single warp issues in order, so a result that is waited for stalls everything behind it (ncu: ~1 cycle per issued instruction,
    // 5-13 per dependent or predicate-dependent one, ~30 per taken branch).  Every shuffle is issued as early as its operand exists and
    // its result used as late as the step allows; the step's only taken branch is the loop's own: the refill sits between two groups,
    // and a run longer than one warp store leaves the inner loop to be filled.
    unsigned base = 0;                                                        // step index = base + k, k inside the current group
    unsigned run = 0, rsym = 0;
    auto run_group = [&](const unsigned dcur) {
        const unsigned cnt = ND - base < 32u ? ND - base : 32u;
        unsigned k = 0;
        unsigned d = __shfl_sync(RCZ_FULL, dcur, 0);                          // the distance of the current step; the later ones are fetched one step ahead
        for (;;) {
            bool go = k < cnt && i < N && !bad;
            while (go) {
                const unsigned stop = e1 >> 8;
                const unsigned future = stop + d;                             // (a flagged step ends the loop before the list is looked at again)
                unsigned cross = __shfl_down_sync(RCZ_FULL, key[0], 1);       // rank R*(l+1); nothing lies above the last lane's entries
                // the run [i, stop): almost always a few bytes, one predicated store — first thing in the step, its pointer moves on
                // last (the add has to wait until the store has read its address).  stop only grows.
                const unsigned stop_c = stop > N ? N : stop;
                run = stop_c - i;
                rsym = sym;
                if (lane < run) *op = (uint8_t)sym;
                const unsigned fk = (future + R * lane) << 8;                 // rank q is passed  <=>  key[q] < (future + q) << 8
                bool c[R];
#pragma unroll
                for (int r = 0; r < R; ++r) c[r] = key[r] < fk + ((unsigned)r << 8);
#pragma unroll
                for (int r = 0; r + 1 < R; ++r)                               // entries whose upper neighbour is in this lane
                    key[r] = c[r + 1 < R ? r + 1 : r] ? key[r + 1 < R ? r + 1 : r] : c[r] ? (fk + ((unsigned)r << 8)) | sym : key[r];
                unsigned e2_new = 0;
                if (R > 1) e2_new = R > 2 ? __shfl_sync(RCZ_FULL, key[R > 2 ? 2 : 0], 0) : __shfl_sync(RCZ_FULL, key[0], 1);   // the new rank 2 is final
                ++k;
                const unsigned d_ahead = __shfl_sync(RCZ_FULL, dcur, (int)(k & 31u));
                bad = d > N || future > N;                                    // output[i] index panic (stop > n); dc.rs:213 assert!(future <= n)
                {                                                             // the lane's last entry: its upper neighbour is the next lane's first
                    cross = lane == 31u ? 0xFFFFFFFFu : cross;
                    key[R - 1] = cross < fk + ((unsigned)R << 8) ? cross : c[R - 1] ? (fk + ((unsigned)(R - 1) << 8)) | sym : key[R - 1];
                }
                if (R == 1) e2_new = __shfl_sync(RCZ_FULL, key[0], 2);
                // the front of the new list: rank 0 is the old rank 1; rank 1 is the old rank 2 if that one is passed, else the
                // re-entered symbol; rank 2 came by broadcast
                const unsigned e1_new = e2 < (future + 2u) << 8 ? e2 : ((future + 1u) << 8) | sym;
                sym = e1 & 0xFFu;
                e1 = e1_new; e2 = e2_new;
                d = d_ahead;
                op += run;
                i = stop_c;
                go = k < cnt && i < N && !bad && run <= 32u;
            }
            if (run <= 32u) break;
            dc_fill_long(out, i - run + 32u + lane, i, rsym);                 // the rest of a long run, then on with the group
            run = 0;
        }
        base += k;
    };
    for (;;) {
        run_group(dA);
        if (bad || i >= N || base >= ND) break;
        { const unsigned gi = base + 32u + lane; dA = __ldg(dist + (gi < ND ? gi : last)); }
        run_group(dB);
        if (bad || i >= N || base >= ND) break;
        { const unsigned gi = base + 32u + lane; dB = __ldg(dist + (gi < ND ? gi : last)); }
    }
    int err = 0;
    if (bad) err = RCZ_E_MALFORMED;
    else if (i < N) {                                                         // out of distances: the reference still looks at the next run first
        const unsigned stop = e1 >> 8;
        if (stop > N) err = RCZ_E_MALFORMED;
        else { for (unsigned k = i + lane; k < stop; k += 32) out[k] = (uint8_t)sym; err = RCZ_E_UNEXPECTED_EOF; }
    }
    if (!err) {                                                               // dc.rs:230-231 assert_eq!
        bool off = false;
#pragma unroll
        for (int r = 0; r < R; ++r) off |= R * lane + r < A && ((key[r] >> 8) < N || (key[r] >> 8) >= N + A);
        if (__any_sync(RCZ_FULL, off) || i != N) err = RCZ_E_MALFORMED;
    }
    return err;
}

__global__ void __launch_bounds__(NT)
dc_decode_kernel(const uint32_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
                 uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ n_arr,
                 int32_t* __restrict__ status, unsigned nblocks) {
    __shared__ DecSmem sm;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (unsigned b = blockIdx.x * WPB + w; b < nblocks; b += gridDim.x * WPB) {
        const uint32_t* in = in_base + in_off[b];
        const unsigned long long len = in_len[b], n = n_arr[b];
        uint8_t* out = out_base + out_off[b];
        if (len < 256 || n >= (1ull << 31)) { if (lane == 0) status[b] = len < 256 ? RCZ_E_UNEXPECTED_EOF : RCZ_E_ARG; continue; }
        const uint32_t* dist = in + 256;
        const unsigned long long ndist = len - 256;
        // ---- dc.rs:169-179: present symbols ordered by first position (stable in the symbol value)
        unsigned init[8]; unsigned rk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { init[k] = in[lane * 8 + k]; rk[k] = 0; }
        unsigned present = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) present += init[k] < n ? 1u : 0u;
        const unsigned A = __reduce_add_sync(RCZ_FULL, present);
        // dc.rs:230 looks at all 256 entries when the loop is over: a symbol that never occurs must still carry a "next position" in
        // [n, n + alphabet) (the encoder writes n), else the block is malformed
        bool stray = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) stray |= init[k] >= n && init[k] >= n + A;
        const bool absent_bad = __any_sync(RCZ_FULL, stray);
        for (int src = 0; src < 32; ++src) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned v = __shfl_sync(RCZ_FULL, init[j], src);
                const unsigned t = (unsigned)src * 8 + j;
                if (v < n) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) rk[k] += (v < init[k] || (v == init[k] && t < lane * 8 + k)) ? 1u : 0u;
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (init[k] < n) { sm.nx[w][rk[k]] = init[k]; sm.sy[w][rk[k]] = (uint8_t)(lane * 8 + k); }   // ranks are a permutation of 0..A-1
        __syncwarp();
        if (A <= 1) {                                                         // dc.rs:180-187
            const unsigned sym = A ? (unsigned)sm.sy[w][0] : 0u;
            for (unsigned long long i = lane; i < n; i += 32) out[i] = (uint8_t)sym;
            if (lane == 0) status[b] = RCZ_OK;
            __syncwarp();
            continue;
        }
        // ---- dc.rs:199-229: see dc_decode_list below; R = list entries per lane, the smallest that holds the block's alphabet
        int err;
        __syncwarp();
        bool tie = false;                                                     // two symbols with one first position: only in a damaged init[] table
        for (unsigned q = lane; q + 1 < A; q += 32) tie |= sm.nx[w][q] == sm.nx[w][q + 1];
        if (!__any_sync(RCZ_FULL, tie) && n + 256 < (1ull << 24)) {
            if (A <= 32) err = dc_decode_sorted<1>(sm.nx[w], sm.sy[w], A, (unsigned)n, dist, ndist, out, lane);
            else if (A <= 64) err = dc_decode_sorted<2>(sm.nx[w], sm.sy[w], A, (unsigned)n, dist, ndist, out, lane);
            else if (A <= 128) err = dc_decode_sorted<4>(sm.nx[w], sm.sy[w], A, (unsigned)n, dist, ndist, out, lane);
            else err = dc_decode_sorted<8>(sm.nx[w], sm.sy[w], A, (unsigned)n, dist, ndist, out, lane);
        } else if (A <= 32) err = dc_decode_list<1>(sm.nx[w], sm.sy[w], A, (unsigned)n, dist, ndist, out, lane);
        else if (A <= 64) err = dc_decode_list<2>(sm.nx[w], sm.sy[w], A, (unsigned)n, dist, ndist, out, lane);
        else if (A <= 128) err = dc_decode_list<4>(sm.nx[w], sm.sy[w], A, (unsigned)n, dist, ndist, out, lane);
        else err = dc_decode_list<8>(sm.nx[w], sm.sy[w], A, (unsigned)n, dist, ndist, out, lane);
        if (!err && absent_bad) err = RCZ_E_MALFORMED;
        __syncwarp();
        if (lane == 0) status[b] = err;
    }
}

}  // namespace dck

extern "C" int rcz_dc_encode_blocks(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* n_arr, uint32_t* out_base,
                                    const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t nblocks,
                                    int mem_kind) {
    using namespace dck;
    if (!c || rcz_bad_kind(mem_kind)) return RCZ_E_ARG;
    if (nblocks == 0) return RCZ_OK;
    if (!in_base || !in_off || !n_arr || !out_base || !out_off || !out_cap || !out_len || !status || nblocks > 0x3fffffu) return RCZ_E_ARG;
    if (!rcz_spans_ok(in_off, n_arr, nblocks) || !rcz_spans_ok(out_off, out_cap, nblocks, 4)) return RCZ_E_ARG;
    rt_set_device(c->device);
    std::vector<Blk> blks(nblocks);
    std::vector<int32_t> hstatus(nblocks, 0);
    std::vector<unsigned> seg2blk;
    for (size_t i = 0; i < nblocks; ++i) {
        Blk& b = blks[i];
        memset(&b, 0, sizeof b);
        b.in_off = in_off[i]; b.out_off = out_off[i]; b.out_cap = out_cap[i];
        b.seg0 = (unsigned)seg2blk.size();
        if (n_arr[i] >= (1ull << 31)) { b.skip = 1; hstatus[i] = RCZ_E_ARG; continue; }
        b.n = (unsigned)n_arr[i];
        b.nseg = (b.n + SEG - 1) / SEG;
        for (unsigned s = 0; s < b.nseg; ++s) seg2blk.push_back((unsigned)i);
        if (seg2blk.size() > 0x3fffffffu) return RCZ_E_ARG;
    }
    const unsigned nsegs = (unsigned)seg2blk.size();
    DescStager ds(c, mem_kind, nblocks);
    const size_t i_blk = ds.add_in(blks.data(), nblocks * sizeof(Blk));
    const size_t i_s2b = ds.add_in(seg2blk.data(), (size_t)nsegs * 4);
    const size_t i_hst = ds.add_in(hstatus.data(), nblocks * 4);
    const size_t o_len = ds.add_out(out_len, nblocks * 8);
    const size_t o_st = ds.add_out(status, nblocks * 4);
    int st = ds.upload(); if (st) return st;
    const uint8_t* din = (const uint8_t*)in_base; uint8_t* dout = (uint8_t*)out_base;
    if (mem_kind == RCZ_MEM_HOST) {
        st = stage_span_in(c, WS_IN, in_base, in_off, n_arr, nblocks, 1, &din); if (st) return st;
        st = stage_span_out(c, WS_OUT, out_off, out_cap, nblocks, 4, &dout); if (st) return st;
    }
    void* wf;
    const size_t sz_first = ((size_t)nsegs * 256 * 4 + 255) & ~(size_t)255;
    st = ctx_ws(c, WS_A, sz_first + (size_t)nsegs * 4 + 256, &wf); if (st) return st;
    unsigned* first = (unsigned*)wf;
    unsigned* nruns = (unsigned*)((uint8_t*)wf + sz_first);
    const Blk* dblk = ds.in_ptr<Blk>(i_blk);
    const unsigned* ds2b = ds.in_ptr<unsigned>(i_s2b);
    st = ctx_timer_begin(c); if (st) return st;
    const unsigned sgrid = (nsegs + WPB - 1) / WPB;
    if (nsegs) RCZ_KLAUNCH(c, dc_first_kernel, sgrid, NT, 0, din, dblk, ds2b, nsegs, first, nruns);
    RCZ_KLAUNCH(c, dc_scan_kernel, (unsigned)nblocks, 256, 0, dblk, first, nruns, (uint32_t*)dout, ds.out_ptr<uint64_t>(o_len), ds.out_ptr<int32_t>(o_st),
                ds.in_ptr<int32_t>(i_hst));
    if (nsegs) RCZ_KLAUNCH(c, dc_emit_kernel, sgrid, NT, 0, din, dblk, ds2b, nsegs, first, nruns, (uint32_t*)dout, ds.out_ptr<int32_t>(o_st));
    st = ctx_timer_end(c); if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) {
        std::vector<uint64_t> clipped(nblocks);
        for (size_t i = 0; i < nblocks; ++i) clipped[i] = status[i] == RCZ_OK ? out_len[i] : 0;
        st = unstage_span_out(c, out_base, dout, out_off, clipped.data(), nblocks, 4); if (st) return st;
    }
    return RCZ_OK;
}

int rcz_dc_decode_launch(rcz_ctx* c, const uint32_t* din, const uint64_t* d_in_off, const uint64_t* d_in_len, uint8_t* dout,
                         const uint64_t* d_out_off, const uint64_t* d_n, int32_t* d_status, size_t nblocks) {
    if (nblocks == 0) return RCZ_OK;
    const unsigned grid = (unsigned)std::min<size_t>((nblocks + dck::WPB - 1) / dck::WPB, (size_t)c->sm_count * 16);
    RCZ_KLAUNCH(c, dck::dc_decode_kernel, grid, dck::NT, 0, din, d_in_off, d_in_len, dout, d_out_off, d_n, d_status, (unsigned)nblocks);
    return RCZ_OK;
}

extern "C" int rcz_dc_decode_blocks(rcz_ctx* c, const uint32_t* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                                    const uint64_t* out_off, const uint64_t* n_arr, int32_t* status, size_t nblocks, int mem_kind) {
    if (!c || rcz_bad_kind(mem_kind)) return RCZ_E_ARG;
    if (nblocks == 0) return RCZ_OK;
    if (!in_base || !in_off || !in_len || !out_base || !out_off || !n_arr || !status || nblocks > 0x7fffffffu) return RCZ_E_ARG;
    if (!rcz_spans_ok(in_off, in_len, nblocks, 4) || !rcz_spans_ok(out_off, n_arr, nblocks)) return RCZ_E_ARG;
    rt_set_device(c->device);
    DescStager ds(c, mem_kind, nblocks);
    ds.add_in(in_off, nblocks * 8); ds.add_in(in_len, nblocks * 8); ds.add_in(out_off, nblocks * 8); ds.add_in(n_arr, nblocks * 8);
    ds.add_out(status, nblocks * 4);
    int st = ds.upload(); if (st) return st;
    const uint8_t* din = (const uint8_t*)in_base; uint8_t* dout = (uint8_t*)out_base;
    if (mem_kind == RCZ_MEM_HOST) {
        st = stage_span_in(c, WS_IN, in_base, in_off, in_len, nblocks, 4, &din); if (st) return st;
        st = stage_span_out(c, WS_OUT, out_off, n_arr, nblocks, 1, &dout); if (st) return st;
    }
    st = ctx_timer_begin(c); if (st) return st;
    st = rcz_dc_decode_launch(c, (const uint32_t*)din, ds.in_ptr<uint64_t>(0), ds.in_ptr<uint64_t>(1), dout, ds.in_ptr<uint64_t>(2), ds.in_ptr<uint64_t>(3),
                              ds.out_ptr<int32_t>(0), nblocks);
    if (st) return st;
    st = ctx_timer_end(c); if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) {
        std::vector<uint64_t> lens(nblocks);
        for (size_t i = 0; i < nblocks; ++i) lens[i] = n_arr[i] < (1ull << 31) ? n_arr[i] : 0;
        st = unstage_span_out(c, out_base, dout, out_off, lens.data(), nblocks, 1); if (st) return st;
    }
    return RCZ_OK;
}
