// bwt_decode.cu — K4 (link-table build) + K5 (inverse-BWT walk) for batches of independent BWT blocks.
//
// Replaces /root/reference/src/bwt/mod.rs:223-239 `compute_inversion_table` and :243-282 `InverseIterator`
// (one dependent `cur = table[cur]-1` hop per output byte, bwt/mod.rs:270) as driven by the stream decoder at
// bwt/mod.rs:388-393.  Same bytes out; the serial chain is cut into thousands of sub-chains per block:
//
//   A  ibwt_hist     per 16 Ki-symbol tile: 256-bin histogram of L (Radix::gather, bwt/mod.rs:95-99)
//   B  ibwt_scan     per block: exclusive scan over symbols and tiles (Radix::accumulate, :102-109)
//   C  ibwt_scatter  per tile: stable 256-way partition (ballot-based equal-key ranking inside each warp, smem-staged so
//                    that every (tile, symbol) run leaves the SM as one contiguous store) ->
//                    P[slot] = (position-in-L << 8) | symbol, with the `origin`-first rule of :230-236
//                    (origin's entry is the END marker; the reference stores index+1 and 0)
//   D  ibwt_walk     sampled list ranking, pass 1: every 2^k-th row starts a sub-chain that hops
//                    cur = P[cur]>>8 until it meets the next sampled row (or END), packing the emitted bytes
//                    16 at a time into a chain-local scratch slot; over-long chains are split on the fly
//                    (continuation slots come from an atomic ticket) so no chain is walked twice
//   E  ibwt_heads / ibwt_headrank   the chain descriptors are ranked by sampled list ranking again (every 32nd chain is a head: one
//                    thread per head walks to the next head; heads are ranked per block by Wyllie pointer jumping in shared memory)
//   F  ibwt_place    one thread per head walks its chains once more and copies their bytes to their final positions
//
// The walk starts at `origin`, emits F[cur] (== L[table[cur]-1], bwt/mod.rs:270-279) per hop, and ends at the END
// entry — exactly the reference's iterator, so a block that is not a valid BWT yields the same (shorter) output.
#include "rcz_internal.h"
#include <algorithm>

namespace ibwt {

constexpr int TB = 16384;                 // symbols per tile (kernels A, C)
constexpr int NT_TILE = 512;
constexpr unsigned END24 = 0xFFFFFFu;     // "next" field of the origin entry
constexpr unsigned SUCC_END = 0xFFFFFFFFu;
constexpr unsigned OFF_INVALID = 0xFFFFFFFFu;
constexpr unsigned MAX_N = 0xFFFFFEu;     // positions must fit the 24-bit field
constexpr int RANK_NT = 1024;

struct Blk {
    unsigned long long in_off, out_off, scratch_off;   // bytes
    unsigned n, origin;
    unsigned p_off;        // element offset of this block's P table
    unsigned tile0, ntiles;
    unsigned stride_log2;  // sampled rows are multiples of 1 << stride_log2
    unsigned K;            // number of sampled rows; chain K is the origin chain
    unsigned cap;          // bytes per chain slot (2 * stride)
    unsigned chain0;       // global id of this block's first chain descriptor
    unsigned max_chains;
    unsigned work0;        // first index of this block's initial work items (K + 1 of them)
    unsigned pf_off, pf_elems;   // link table of the block `prefetch distance` ahead (element offset, element count; 0 = none)
    unsigned head0;        // first slot of this block's head nodes (ranking of the chain descriptors)
    unsigned skip;         // block rejected on the host (status already set)
    unsigned bad;          // origin supplied on the device (composed calls) and out of range: walked with origin 0, reported MALFORMED
};

// ------------------------------------------------------------------------------------------ A: histogram
__global__ void __launch_bounds__(NT_TILE)
ibwt_hist_kernel(const uint8_t* __restrict__ in_base, const Blk* __restrict__ blks, const unsigned* __restrict__ tile2blk,
                 unsigned* __restrict__ tile_hist) {
    __shared__ unsigned h[NT_TILE / 32][256];
    const unsigned tile = blockIdx.x, tid = threadIdx.x, w = tid >> 5;
    const Blk bk = blks[tile2blk[tile]];
    const unsigned t = tile - bk.tile0;
    const uint8_t* L = in_base + bk.in_off;
    const unsigned lo = t * TB, hi = min(bk.n, lo + TB);
    for (unsigned i = tid; i < (NT_TILE / 32) * 256; i += NT_TILE) (&h[0][0])[i] = 0;
    __syncthreads();
    for (unsigned i = lo + tid; i < hi; i += NT_TILE) atomicAdd(&h[w][L[i]], 1u);
    __syncthreads();
    if (tid < 256) {
        unsigned s = 0;
#pragma unroll
        for (int k = 0; k < NT_TILE / 32; ++k) s += h[k][tid];
        tile_hist[(size_t)tile * 256 + tid] = s;
    }
}

// ------------------------------------------------------------------------------------------ B: scan
// tile_hist[tile][c] becomes the number of c's in earlier tiles of the block; cbase[blk][c] = #symbols < c.
__global__ void __launch_bounds__(256)
ibwt_scan_kernel(const Blk* __restrict__ blks, unsigned* __restrict__ tile_hist, unsigned* __restrict__ cbase) {
    __shared__ unsigned scratch[40];
    const unsigned b = blockIdx.x, c = threadIdx.x;
    const Blk bk = blks[b];
    unsigned run = 0;
    if (!bk.skip)
        for (unsigned t0 = 0; t0 < bk.ntiles; t0 += 16) {                 // 16 independent loads in flight, then the running sums
            unsigned hc[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) hc[k] = t0 + k < bk.ntiles ? tile_hist[(size_t)(bk.tile0 + t0 + k) * 256 + c] : 0u;
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (t0 + k < bk.ntiles) { tile_hist[(size_t)(bk.tile0 + t0 + k) * 256 + c] = run; run += hc[k]; }
        }
    unsigned total;
    const unsigned ex = block_excl_scan_add<256>(run, scratch, &total);
    cbase[(size_t)b * 256 + c] = ex;
}

// ------------------------------------------------------------------------------------------ C: stable partition
struct ScatterSmem {
    unsigned sorted[TB];
    unsigned wcnt[NT_TILE / 32][256];
    unsigned symbase[256];
    unsigned gbase[256];
    unsigned scratch[40];
};

__global__ void __launch_bounds__(NT_TILE, 2)
ibwt_scatter_kernel(const uint8_t* __restrict__ in_base, const Blk* __restrict__ blks, const unsigned* __restrict__ tile2blk,
                    const unsigned* __restrict__ tile_hist, const unsigned* __restrict__ cbase, unsigned* __restrict__ P_base) {
    RCZ_DYN_SMEM(raw);
    ScatterSmem& sm = *reinterpret_cast<ScatterSmem*>(raw);
    const unsigned tile = blockIdx.x, tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const unsigned bi = tile2blk[tile];
    const Blk bk = blks[bi];
    const unsigned t = tile - bk.tile0;
    const uint8_t* L = in_base + bk.in_off;
    unsigned* P = P_base + bk.p_off;
    const unsigned lo = t * TB, hi = min(bk.n, lo + TB), tlen = hi - lo;
    constexpr unsigned WSPAN = TB / (NT_TILE / 32);   // 1024 consecutive symbols per warp

    for (unsigned i = tid; i < (NT_TILE / 32) * 256; i += NT_TILE) (&sm.wcnt[0][0])[i] = 0;
    __syncthreads();
    // pass A: per-warp symbol counts, and for every symbol its rank among the equal symbols of the warp's 1024-symbol span
    // (count before this group of 32 + rank inside the group: 8 ballots, done ONCE and kept in registers, two ranks per register)
    unsigned rb[WSPAN / 64];
#pragma unroll
    for (unsigned g = 0; g < WSPAN / 32; ++g) {
        const unsigned i = lo + w * WSPAN + g * 32 + lane;
        const bool valid = i < hi;
        const unsigned s = valid ? (unsigned)L[i] : 0u;
        const unsigned m = warp_match_u8(s, valid);
        const unsigned r = __popc(m & ((1u << lane) - 1u));
        unsigned old = 0;
        if (valid && r == 0) { old = sm.wcnt[w][s]; sm.wcnt[w][s] = old + (unsigned)__popc(m); }   // one add per distinct symbol
        old = __shfl_sync(RCZ_FULL, old, m ? __ffs((int)m) - 1 : 0);
        const unsigned v = old + r;                                           // < 1024 + 32
        if (g & 1) rb[g >> 1] |= v << 16; else rb[g >> 1] = v;
        __syncwarp();
    }
    __syncthreads();
    // pass B: exclusive scan over warps (per symbol), then over symbols
    unsigned tot = 0;
    if (tid < 256) {
#pragma unroll
        for (int k = 0; k < NT_TILE / 32; ++k) { const unsigned x = sm.wcnt[k][tid]; sm.wcnt[k][tid] = tot; tot += x; }
    }
    unsigned dummy;
    const unsigned sb = block_excl_scan_add<NT_TILE>(tot, sm.scratch, &dummy);
    if (tid < 256) {
        sm.symbase[tid] = sb;
        // destination of local sorted index j with symbol c:  gbase[c] + j
        sm.gbase[tid] = cbase[(size_t)bi * 256 + tid] + tile_hist[(size_t)tile * 256 + tid] - sb;
    }
    __syncthreads();
    // pass C: place (the symbols come back from L1; no ballots, no warp-level ordering left)
#pragma unroll
    for (unsigned g = 0; g < WSPAN / 32; ++g) {
        const unsigned i = lo + w * WSPAN + g * 32 + lane;
        if (i < hi) {
            const unsigned s = (unsigned)L[i];
            const unsigned v = (g & 1) ? rb[g >> 1] >> 16 : rb[g >> 1] & 0xffffu;
            sm.sorted[sm.symbase[s] + sm.wcnt[w][s] + v] = (i << 8) | s;
        }
    }
    __syncthreads();
    // pass D: write out; runs of equal symbols are contiguous in P. origin-first rule of bwt/mod.rs:230-236.
    const unsigned origin = bk.origin;
    const unsigned c0 = L[origin];
    const unsigned slot0 = cbase[(size_t)bi * 256 + c0];
    for (unsigned j = tid; j < tlen; j += NT_TILE) {
        const unsigned v = sm.sorted[j];
        const unsigned s = v & 255u, pos = v >> 8;
        unsigned slot = sm.gbase[s] + j;
        unsigned val = v;
        if (s == c0) {
            if (pos == origin) { slot = slot0; val = (END24 << 8) | s; }
            else if (pos < origin) slot += 1;
        }
        P[slot] = val;
    }
}

// ------------------------------------------------------------------------------------------ D: walk
struct Desc { unsigned len, succ; };

__global__ void __launch_bounds__(256)
ibwt_walk_kernel(const Blk* __restrict__ blks, unsigned nblocks, unsigned total_work, const unsigned* __restrict__ P_base,
                 uint8_t* __restrict__ scratch_base, Desc* __restrict__ desc_base, unsigned* __restrict__ chain_ctr,
                 unsigned* __restrict__ queue) {
    const unsigned lane = threadIdx.x & 31;
    bool active = false, done = false;
    // per-chain state
    const unsigned* P = nullptr; uint8_t* slotp = nullptr; Desc* desc = nullptr; unsigned* ctr = nullptr;
    unsigned cur = 0, count = 0, chain = 0, mask = 0, slog = 0, cap = 0, hops = 0, nmax = 0, maxch = 0;
    unsigned long long scratch_off = 0;
    unsigned b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    // block of the lane's previous work item: consecutive tickets almost always stay inside it, so the block table is searched
    // (and the dozen per-block values reloaded) only when a ticket leaves [w_lo, w_hi)
    unsigned w_lo = 1, w_hi = 0, bK = 0, borigin = 0, pf_off = 0, pf_elems = 0;

    for (;;) {
        // ---- refill idle lanes with new chains (warp-aggregated ticket)
        const unsigned need = __ballot_sync(RCZ_FULL, !active && !done);
        if (need) {
            unsigned base = 0;
            const unsigned leader = (unsigned)__ffs((int)need) - 1;
            if (lane == leader) base = atomicAdd(queue, (unsigned)__popc(need));
            base = __shfl_sync(RCZ_FULL, base, (int)leader);
            if (!active && !done) {
                const unsigned wi = base + __popc(need & ((1u << lane) - 1u));
                if (wi >= total_work) done = true;
                else {
                    if (wi < w_lo || wi >= w_hi) {
                        unsigned lo = 0, hi = nblocks;            // last block with work0 <= wi
                        while (hi - lo > 1) { const unsigned mid = (lo + hi) >> 1; if (blks[mid].work0 <= wi) lo = mid; else hi = mid; }
                        const Blk bk = blks[lo];
                        w_lo = bk.work0; w_hi = bk.work0 + bk.K + 1; bK = bk.K; borigin = bk.origin;
                        P = P_base + bk.p_off;
                        desc = desc_base + bk.chain0;
                        ctr = chain_ctr + lo;
                        scratch_off = bk.scratch_off;
                        slog = bk.stride_log2; mask = (1u << slog) - 1u; cap = bk.cap; nmax = bk.n; maxch = bk.max_chains;
                        pf_off = bk.pf_off; pf_elems = bk.pf_elems;
                    }
                    chain = wi - w_lo;
                    cur = chain < bK ? (chain << slog) : borigin;
                    slotp = scratch_base + scratch_off + (size_t)chain * cap;
                    count = 0; hops = 0;
                    active = true;
                    // (optional) every chain start pulls its share of a LATER block's link table into L2
                    if (pf_elems) {
                        const unsigned per = ((pf_elems + bK) / (bK + 1) + 3u) & ~3u;                   // elements per chain, 16-byte multiple
                        const unsigned long long e0 = (unsigned long long)chain * per;
                        if (e0 < pf_elems) {
                            const unsigned cnt = (unsigned)min((unsigned long long)per, (unsigned long long)pf_elems - e0) & ~3u;
                            if (cnt) prefetch_l2(P_base + pf_off + e0, cnt * 4u);
                        }
                    }
                    if (chain < bK && cur == borigin) {            // no row links to `origin`: this sampled chain is unreachable,
                        Desc d; d.len = 0; d.succ = SUCC_END;      // and the origin chain (id K) walks the same rows
                        desc[chain] = d;
                        active = false;
                    }
                }
            }
        }
        if (__all_sync(RCZ_FULL, done)) break;
        // ---- a burst of hops between refill checks
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            if (active) {
                const unsigned e = __ldcg(P + cur);             // L2 only: an L1 miss would pull the whole 128-byte line for one 4-byte entry
                const unsigned byte = e & 255u, nxt = e >> 8;
                // append to the 16-byte register buffer
                const unsigned k = count & 15u;
                const unsigned sh = (k & 3u) * 8u;
                if (k < 4) b0 = (k == 0 ? 0u : b0) | (byte << sh);
                else if (k < 8) b1 = (k == 4 ? 0u : b1) | (byte << sh);
                else if (k < 12) b2 = (k == 8 ? 0u : b2) | (byte << sh);
                else b3 = (k == 12 ? 0u : b3) | (byte << sh);
                ++count; ++hops;
                if ((count & 15u) == 0) __stcs(reinterpret_cast<uint4*>(slotp + count - 16), make_uint4(b0, b1, b2, b3));   // streaming: keep L2 for the tables
                const bool at_end = (nxt == END24) || hops > nmax;
                const bool at_sample = !at_end && (nxt & mask) == 0;
                if (at_end || at_sample) {
                    if (count & 15u) {
                        if ((count & 15u) <= 4) { b1 = 0; b2 = 0; b3 = 0; } else if ((count & 15u) <= 8) { b2 = 0; b3 = 0; } else if ((count & 15u) <= 12) b3 = 0;
                        __stcs(reinterpret_cast<uint4*>(slotp + (count & ~15u)), make_uint4(b0, b1, b2, b3));
                    }
                    Desc d; d.len = count; d.succ = at_end ? SUCC_END : (nxt >> slog);
                    desc[chain] = d;
                    active = false;
                } else {
                    cur = nxt;
                    if (count == cap) {                       // slot full: continue in a fresh chain slot
                        const unsigned nc = atomicAdd(ctr, 1u);
                        Desc d; d.len = count; d.succ = nc < maxch ? nc : SUCC_END;
                        desc[chain] = d;
                        if (nc >= maxch) active = false;      // cannot happen (see max_chains); never write out of bounds
                        chain = nc;
                        slotp = scratch_base + scratch_off + (size_t)chain * cap;
                        count = 0;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ E/F: rank the chains, move their bytes
// Sampled list ranking once more, one level up: every 2^hlog-th chain descriptor (and the origin chain) is a HEAD.
//   ibwt_heads_kernel    one thread per head, all blocks at once: walk the chain list to the next head summing lengths
//   ibwt_headrank_kernel one CTA per block: the <= nch / 2^hlog + 1 heads are ranked by Wyllie pointer jumping in shared memory
//                        ((distance to END << 32) | next head) -> output offset of every head
//   ibwt_place_kernel    one thread per head again: walk the same chains and copy their bytes from the chain slots to their final
//                        offsets (aligned 16-byte loads; head bytes, funnel-shifted 32-bit words, tail bytes)
// The two walks are O(nch) dependent 8-byte loads spread over ~10 K threads per block — latency is hidden by parallelism instead of
// being paid by one CTA per block (the round-1 kernel: 1.5 ms per wave of blocks once chains became short and numerous).
constexpr unsigned RANK_MAX_HEADS = 12288;                    // 2 x 8 B each in shared memory (read copy, write copy)

struct HeadGeom { unsigned nch, nreg, H, K, mask, hlog; };
__device__ __forceinline__ HeadGeom head_geom(const Blk& bk, const unsigned* chain_ctr, unsigned b, unsigned hlog) {
    HeadGeom g;
    g.nch = chain_ctr[b];
    if (g.nch > bk.max_chains) g.nch = bk.max_chains;
    g.nreg = (g.nch + (1u << hlog) - 1) >> hlog;              // heads at chain ids 0, 2^hlog, ...
    g.H = g.nreg + 1;                                         // + the origin chain (id K); SENT = H
    g.K = bk.K; g.mask = (1u << hlog) - 1u; g.hlog = hlog;
    return g;
}

__global__ void __launch_bounds__(256)
ibwt_heads_kernel(const Blk* __restrict__ blks, const Desc* __restrict__ desc_base, const unsigned* __restrict__ chain_ctr,
                  unsigned long long* __restrict__ node_base, unsigned hlog) {
    const unsigned b = blockIdx.y;
    const Blk bk = blks[b];
    if (bk.skip) return;
    const HeadGeom g = head_geom(bk, chain_ctr, b, hlog);
    const Desc* desc = desc_base + bk.chain0;
    unsigned long long* node = node_base + bk.head0;
    for (unsigned h = blockIdx.x * blockDim.x + threadIdx.x; h < g.H; h += gridDim.x * blockDim.x) {
        unsigned cur = h < g.nreg ? h << hlog : g.K;
        unsigned long long acc = 0;
        unsigned nxt_head = h;                                          // self-loop unless the walk reaches a head or END
        if (cur < g.nch) {
            for (unsigned steps = 0; steps <= g.nch; ++steps) {
                const Desc d = desc[cur];
                acc += d.len;
                const unsigned s = d.succ;
                if (s == SUCC_END) { nxt_head = g.H; break; }
                if (s >= g.nch) break;                                  // dangling link: never reaches END
                if (s == g.K) { nxt_head = g.nreg; break; }
                if ((s & g.mask) == 0) { nxt_head = s >> hlog; break; }
                cur = s;
            }
        }
        node[h] = (acc << 32) | nxt_head;
    }
}

__global__ void __launch_bounds__(RANK_NT, 1)
ibwt_headrank_kernel(const Blk* __restrict__ blks, const unsigned* __restrict__ chain_ctr, unsigned long long* __restrict__ node_base,
                     uint64_t* __restrict__ out_len, int32_t* __restrict__ status, unsigned hlog) {
    RCZ_DYN_SMEM(raw);
    unsigned long long* node = reinterpret_cast<unsigned long long*>(raw);
    const unsigned b = blockIdx.x, tid = threadIdx.x;
    const Blk bk = blks[b];
    if (bk.skip) return;
    const HeadGeom g = head_geom(bk, chain_ctr, b, hlog);
    unsigned long long* gnode = node_base + bk.head0;
    const unsigned H = g.H, SENT = g.H;
    // two copies of the node array: a round reads one and writes the other, so the jumping is race-free (one barrier per round)
    unsigned long long* nodeB = node + (H + 2);
    for (unsigned i = tid; i < H; i += RANK_NT) node[i] = gnode[i];
    if (tid == 0) { node[SENT] = SENT; nodeB[SENT] = SENT; }
    __syncthreads();
    // ---- Wyllie over the heads
    unsigned rounds = 2;
    for (unsigned v = H; v; v >>= 1) ++rounds;
    unsigned long long* src = node; unsigned long long* dst = nodeB;
    for (unsigned r = 0; r < rounds; ++r) {
        int pend = 0;
        for (unsigned i = tid; i < H; i += RANK_NT) {
            unsigned long long a = src[i];
            const unsigned s = (unsigned)a;
            if (s != SENT && s != i) {
                const unsigned long long q = src[s];
                a = (((a >> 32) + (q >> 32)) << 32) | (unsigned)q;
                pend = 1;
            }
            dst[i] = a;
        }
        unsigned long long* t2 = src; src = dst; dst = t2;
        if (!__syncthreads_or(pend)) break;
    }
    node = src;                                                          // the copy the last round wrote
    const unsigned long long org = node[g.nreg];
    const unsigned total = (unsigned)org == SENT ? (unsigned)(org >> 32) : 0u;   // the origin chain always reaches END
    // head h -> output offset of its first chain (OFF_INVALID when it never reaches END: its chains are not part of the output)
    for (unsigned i = tid; i < H; i += RANK_NT) {
        const unsigned long long a = node[i];
        gnode[i] = (unsigned)a == SENT ? (unsigned long long)(total - (unsigned)(a >> 32)) : (unsigned long long)OFF_INVALID;
    }
    if (tid == 0) { out_len[b] = bk.bad ? 0 : total; status[b] = bk.bad ? RCZ_E_MALFORMED : RCZ_OK; }
}

// len bytes from the 16-byte-aligned chain slot to an arbitrarily aligned destination
__device__ __forceinline__ void copy_chain(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, unsigned len) {
    for (unsigned base = 0; base < len; base += 16) {
        const uint4 q = __ldcs(reinterpret_cast<const uint4*>(src + base));      // read once: streaming
        const unsigned m = len - base < 16u ? len - base : 16u;
        const unsigned wv[5] = {q.x, q.y, q.z, q.w, 0u};
        uint8_t* d = dst + base;
        const unsigned head = (unsigned)((4u - ((uintptr_t)d & 3u)) & 3u);
        if (m >= 4u + head) {
            unsigned done = 0;
            for (; done < head; ++done) d[done] = (uint8_t)(q.x >> (8u * done));
            const unsigned sh = 8u * head;                                       // words from byte `head` on: (w[i] >> sh) | (w[i+1] << (32 - sh))
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (done + 4u <= m) { *reinterpret_cast<unsigned*>(d + done) = __funnelshift_r(wv[i], wv[i + 1], sh); done += 4; }
            for (; done < m; ++done) d[done] = (uint8_t)(wv[done >> 2] >> (8u * (done & 3u)));
        } else {
            for (unsigned k = 0; k < m; ++k) d[k] = (uint8_t)(wv[k >> 2] >> (8u * (k & 3u)));
        }
    }
}

// Persistent CTAs take (block, 256 heads) tickets in block-major order, so that only a few blocks are in flight at a time: a block's
// descriptors and chain slots (read in list order = random order) then stay in L2 while its heads are processed, and every 64-byte
// line that DRAM delivers serves all the chains in it instead of one.
__global__ void __launch_bounds__(256)
ibwt_place_kernel(const Blk* __restrict__ blks, unsigned nblocks, unsigned tiles_per_block, const uint8_t* __restrict__ scratch_base,
                  const Desc* __restrict__ desc_base, const unsigned* __restrict__ chain_ctr, const unsigned long long* __restrict__ node_base,
                  uint8_t* __restrict__ out_base, unsigned hlog, unsigned* __restrict__ ticket) {
    __shared__ unsigned s_t;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_t = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned t = s_t;
        if (t >= nblocks * tiles_per_block) break;
        const unsigned b = t / tiles_per_block, tile = t - b * tiles_per_block;
        const Blk bk = blks[b];
        if (bk.skip) continue;
        const HeadGeom g = head_geom(bk, chain_ctr, b, hlog);
        const unsigned h = tile * 256u + threadIdx.x;
        if (h >= g.H) continue;
        const Desc* desc = desc_base + bk.chain0;
        const unsigned long long* node = node_base + bk.head0;
        const uint8_t* scratch = scratch_base + bk.scratch_off;
        uint8_t* out = out_base + bk.out_off;
        unsigned off = (unsigned)node[h];
        if (off == OFF_INVALID) continue;
        unsigned cur = h < g.nreg ? h << hlog : g.K;
        if (cur >= g.nch) continue;
        if (h < g.nreg && cur == g.K) continue;                         // the origin chain is walked once, as head `nreg`
        for (unsigned steps = 0; steps <= g.nch; ++steps) {
            const Desc d = desc[cur];
            if (off + d.len > bk.n) break;                              // (a well-formed block never gets here)
            copy_chain(scratch + (size_t)cur * bk.cap, out + off, d.len);
            off += d.len;
            const unsigned s = d.succ;
            if (s == SUCC_END || s >= g.nch || s == g.K || (s & g.mask) == 0) break;
            cur = s;
        }
    }
}

// origin_dev (optional): the origins live in device memory (bwt -> dc -> ari pipeline); they are patched into the block table here,
// before any other kernel of the group reads it
__global__ void ibwt_init_kernel(Blk* __restrict__ blks, unsigned nblocks, unsigned* __restrict__ chain_ctr, unsigned* __restrict__ queue,
                                 uint64_t* __restrict__ out_len, int32_t* __restrict__ status, const int32_t* __restrict__ host_status,
                                 const uint32_t* __restrict__ origin_dev) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { queue[0] = 0; queue[1] = 0; }      // work tickets of the walk and of the placement
    if (i < nblocks) {
        if (origin_dev && !blks[i].skip) {
            const unsigned o = origin_dev[i];
            const bool ok = o < blks[i].n;                                   // bwt/mod.rs:230 index panic otherwise
            blks[i].origin = ok ? o : 0u;
            blks[i].bad = ok ? 0u : 1u;
        }
        chain_ctr[i] = blks[i].K + 1;
        if (blks[i].skip) { out_len[i] = 0; status[i] = host_status[i]; }
    }
}

}  // namespace ibwt

extern "C" int rcz_bwt_decode_blocks(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* n_arr,
                                     const uint32_t* origin, void* out_base, const uint64_t* out_off,
                                     uint64_t* out_len, int32_t* status, size_t nblocks, int mem_kind) {
    if (nblocks && !origin) return RCZ_E_ARG;
    if (mem_kind == RCZ_MEM_HOST && c && nblocks && in_base && in_off && n_arr && out_base && out_off && out_len && status &&
        rcz_spans_ok(in_off, n_arr, nblocks) && rcz_spans_ok(out_off, n_arr, nblocks)) {   // big host batches: pipelined chunks of 128 MiB (sweep: profiles/r2_host_chunk_sweep.txt)
        rt_set_device(c->device);
        bool handled = false;
        const int st = host_chunked(c, nblocks, 128ull << 20, in_base, in_off, n_arr, 1, out_base, out_off, n_arr, 1,
            [&](size_t b0, size_t nb, const uint8_t* din, uint8_t* dout) {
                return rcz_bwt_decode_run(c, din, in_off + b0, n_arr + b0, origin + b0, nullptr, dout, out_off + b0, out_len + b0, status + b0, nb, RCZ_MEM_DEVICE);
            },
            [&](size_t i) { return status[i] == RCZ_OK ? out_len[i] : 0; }, &handled);
        if (st || handled) return st;
    }
    return rcz_bwt_decode_run(c, in_base, in_off, n_arr, origin, nullptr, out_base, out_off, out_len, status, nblocks, mem_kind);
}

int rcz_bwt_decode_run(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* n_arr, const uint32_t* origin,
                       const uint32_t* origin_dev, void* out_base, const uint64_t* out_off, uint64_t* out_len, int32_t* status,
                       size_t nblocks, int mem_kind) {
    using namespace ibwt;
    if (!c || rcz_bad_kind(mem_kind)) return RCZ_E_ARG;
    if (nblocks == 0) return RCZ_OK;
    if (!in_base || !in_off || !n_arr || (!origin && !origin_dev) || !out_base || !out_off || !out_len || !status) return RCZ_E_ARG;
    if (nblocks > 0x3fffffu) return RCZ_E_ARG;
    for (size_t i = 0; i < nblocks; ++i) if (in_off[i] > (1ull << 62) || out_off[i] > (1ull << 62)) return RCZ_E_ARG;
    rt_set_device(c->device);

    // ---- host-side geometry.  Blocks are processed in GROUPS of <= group_syms symbols (default 2^30: one group for 256 blocks of
    // 4 MiB), one kernel sequence per group; all offsets in Blk are group-relative and the workspaces are reused by every group.
    // L2 residency of the link tables does not come from small groups (a group of a few blocks leaves the GPU idle at every
    // kernel boundary and pays a chain tail per group: measured 2-6x slower, profiles/r2_ibwt_sweeps.txt) but from SHORT chains:
    // rows are sampled every 2^slog = 16, so a chain lives for ~16 hops and the work queue (block-major) keeps the live chains
    // inside the last one or two blocks' tables (L2 hit rate 69 % against 32 % with a stride of 64, profiles/r2_ibwt_walk_kernel.txt).
    // Ceiling for this access pattern, tools/micro/gather_bench2.cu: 285 G hops/s while the table fits L2.  RCZ_IBWT_* are tuning
    // overrides used by the sweeps.
    const unsigned tune_slog = getenv("RCZ_IBWT_SLOG") ? (unsigned)atoi(getenv("RCZ_IBWT_SLOG")) : 4u;
    const unsigned tune_place = getenv("RCZ_IBWT_PLACE_CTAS") ? (unsigned)atoi(getenv("RCZ_IBWT_PLACE_CTAS")) : 8u;
    const unsigned tune_ctas = getenv("RCZ_IBWT_WALK_CTAS") ? (unsigned)atoi(getenv("RCZ_IBWT_WALK_CTAS")) : 4u;
    const unsigned long long group_syms = getenv("RCZ_IBWT_GROUP_SYMS") ? strtoull(getenv("RCZ_IBWT_GROUP_SYMS"), nullptr, 10) : (1ull << 30);
    struct Group { size_t b0, b1; unsigned tile0_abs, ntiles; unsigned long long p_elems, scratch_bytes, chains, work, heads; };
    std::vector<Blk> blks(nblocks);
    std::vector<int32_t> hstatus(nblocks, 0);
    std::vector<unsigned> tile2blk;
    std::vector<Group> groups;
    Group cur{0, 0, 0, 0, 0, 0, 0, 0, 0};
    unsigned hlog = 5, max_mc = 0;
    for (size_t i = 0; i < nblocks; ++i) {
        const unsigned long long n = n_arr[i];
        if (cur.b1 > cur.b0 && cur.p_elems && cur.p_elems + n > group_syms) {
            groups.push_back(cur);
            cur = Group{i, i, (unsigned)tile2blk.size(), 0, 0, 0, 0, 0, 0};
        }
        Blk& b = blks[i];
        memset(&b, 0, sizeof b);
        b.in_off = in_off[i]; b.out_off = out_off[i];
        b.tile0 = cur.ntiles;
        b.chain0 = (unsigned)cur.chains; b.work0 = (unsigned)cur.work;
        b.p_off = (unsigned)cur.p_elems; b.scratch_off = cur.scratch_bytes;
        cur.b1 = i + 1;
        if (n == 0 || (!origin_dev && origin[i] >= n)) { b.skip = 1; hstatus[i] = RCZ_E_MALFORMED; continue; }   // bwt/mod.rs:230 index panic
        if (n > MAX_N) { b.skip = 1; hstatus[i] = RCZ_E_UNSUPPORTED; continue; }
        b.n = (unsigned)n; b.origin = origin_dev ? 0u : origin[i];
        unsigned slog = tune_slog;
        while ((n >> slog) > (1u << 19)) ++slog;              // keep <= 512 Ki sampled rows per block
        b.stride_log2 = slog;
        b.K = (unsigned)((n + (1ull << slog) - 1) >> slog);
        b.cap = std::max(32u, 4u << slog);                    // bytes per chain slot: P(chain longer than 4 strides) = e^-4
        b.max_chains = b.K + 1 + (unsigned)(n / b.cap) + 1;   // every row is walked at most once => <= n/cap continuations
        while (((b.max_chains >> hlog) + 3) > RANK_MAX_HEADS) ++hlog;   // one head sampling for the whole call
        max_mc = std::max(max_mc, b.max_chains);
        b.ntiles = (unsigned)((n + TB - 1) / TB);
        for (unsigned t = 0; t < b.ntiles; ++t) tile2blk.push_back((unsigned)(i - cur.b0));
        cur.ntiles += b.ntiles;
        cur.p_elems += (n + 63) & ~63ull;
        cur.scratch_bytes += (unsigned long long)b.max_chains * b.cap;
        cur.chains += b.max_chains;
        cur.work += b.K + 1;
    }
    groups.push_back(cur);
    // head nodes of the chain ranking: (max_chains >> hlog) + 3 slots per block, group-relative like the other offsets
    for (auto& g : groups) {
        unsigned long long h0 = 0;
        for (size_t i = g.b0; i < g.b1; ++i) { blks[i].head0 = (unsigned)h0; if (!blks[i].skip) h0 += (blks[i].max_chains >> hlog) + 3; }
        g.heads = h0;
    }
    // prefetch distance (blocks ahead) of the walk; 0 = off
    const unsigned pf_dist = getenv("RCZ_IBWT_PREFETCH") ? (unsigned)atoi(getenv("RCZ_IBWT_PREFETCH")) : 0u;
    if (pf_dist)
        for (auto& g : groups)
            for (size_t i = g.b0; i + pf_dist < g.b1; ++i) {
                const Blk& t = blks[i + pf_dist];
                if (!blks[i].skip && !t.skip) { blks[i].pf_off = t.p_off; blks[i].pf_elems = (t.n + 3u) & ~3u; }
            }
    unsigned long long max_p = 0, max_scratch = 0, max_chains = 0, max_heads = 0; unsigned max_tiles = 0; size_t max_nb = 0;
    for (auto& g : groups) {
        max_p = std::max(max_p, g.p_elems); max_scratch = std::max(max_scratch, g.scratch_bytes); max_chains = std::max(max_chains, g.chains);
        max_tiles = std::max(max_tiles, g.ntiles); max_nb = std::max(max_nb, g.b1 - g.b0); max_heads = std::max(max_heads, g.heads);
    }

    DescStager ds(c, mem_kind, nblocks);
    const size_t i_blk = ds.add_in(blks.data(), nblocks * sizeof(Blk));
    const size_t i_t2b = ds.add_in(tile2blk.data(), tile2blk.size() * 4);
    const size_t i_hst = ds.add_in(hstatus.data(), nblocks * 4);
    const size_t o_len = ds.add_out(out_len, nblocks * 8);
    const size_t o_st = ds.add_out(status, nblocks * 4);
    int st = ds.upload(); if (st) return st;

    const uint8_t* din = (const uint8_t*)in_base; uint8_t* dout = (uint8_t*)out_base;
    if (mem_kind == RCZ_MEM_HOST) {
        st = stage_span_in(c, WS_IN, in_base, in_off, n_arr, nblocks, 1, &din); if (st) return st;
        st = stage_span_out(c, WS_OUT, out_off, n_arr, nblocks, 1, &dout); if (st) return st;
    }
    void *wP, *wS, *wM;
    st = ctx_ws(c, WS_A, (size_t)max_p * 4 + 256, &wP); if (st) return st;
    st = ctx_ws(c, WS_B, (size_t)max_scratch + 256, &wS); if (st) return st;
    // misc: tile_hist | cbase | desc | head nodes | chain_ctr | queue
    const size_t sz_hist = (size_t)max_tiles * 256 * 4, sz_cb = max_nb * 256 * 4, sz_desc = (size_t)max_chains * 8, sz_coff = (size_t)max_heads * 8,
                 sz_ctr = max_nb * 4;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    st = ctx_ws(c, WS_C, al(sz_hist) + al(sz_cb) + al(sz_desc) + al(sz_coff) + al(sz_ctr) + 512, &wM); if (st) return st;
    uint8_t* m = (uint8_t*)wM;
    unsigned* tile_hist = (unsigned*)m; m += al(sz_hist);
    unsigned* cbase = (unsigned*)m; m += al(sz_cb);
    Desc* desc = (Desc*)m; m += al(sz_desc);
    unsigned long long* nodes = (unsigned long long*)m; m += al(sz_coff);
    unsigned* chain_ctr = (unsigned*)m; m += al(sz_ctr);
    unsigned* queue = (unsigned*)m;

    RCZ_CK(c, RCZ_KERNEL_SMEM_OPTIN(ibwt_scatter_kernel, sizeof(ScatterSmem)));
    const size_t rank_smem = ((size_t)(max_mc >> hlog) + 5) * 8 * 2;     // two copies of the head nodes
    RCZ_CK(c, RCZ_KERNEL_SMEM_OPTIN(ibwt_headrank_kernel, rank_smem));
    st = ctx_timer_begin(c); if (st) return st;
    // rcz_last_stage_ms: {partition (hist, scan, scatter), walk, rank + compact} of the FIRST group
    bool mark = true;
    for (auto& g : groups) {
        const unsigned nb = (unsigned)(g.b1 - g.b0);
        if (nb == 0) continue;
        if (mark) { st = ctx_stage_mark(c, 0); if (st) return st; }
        Blk* dblk = const_cast<Blk*>(ds.in_ptr<Blk>(i_blk)) + g.b0;
        const unsigned* dt2b = ds.in_ptr<unsigned>(i_t2b) + g.tile0_abs;
        uint64_t* d_len = ds.out_ptr<uint64_t>(o_len) + g.b0;
        int32_t* d_st = ds.out_ptr<int32_t>(o_st) + g.b0;
        RCZ_KLAUNCH(c, ibwt_init_kernel, (nb + 255) / 256, 256, 0, dblk, nb, chain_ctr, queue, d_len, d_st, ds.in_ptr<int32_t>(i_hst) + g.b0,
                    origin_dev ? origin_dev + g.b0 : (const uint32_t*)nullptr);
        if (!g.ntiles) continue;
        RCZ_KLAUNCH(c, ibwt_hist_kernel, g.ntiles, NT_TILE, 0, din, dblk, dt2b, tile_hist);
        RCZ_KLAUNCH(c, ibwt_scan_kernel, nb, 256, 0, dblk, tile_hist, cbase);
        RCZ_KLAUNCH(c, ibwt_scatter_kernel, g.ntiles, NT_TILE, sizeof(ScatterSmem), din, dblk, dt2b, tile_hist, cbase, (unsigned*)wP);
        if (mark) { st = ctx_stage_mark(c, 1); if (st) return st; }
        const unsigned walk_grid = (unsigned)std::min<unsigned long long>((g.work + 255) / 256, (unsigned long long)c->sm_count * tune_ctas);
        RCZ_KLAUNCH(c, ibwt_walk_kernel, walk_grid, 256, 0, dblk, nb, (unsigned)g.work, (const unsigned*)wP, (uint8_t*)wS, desc, chain_ctr, queue);
        if (mark) { st = ctx_stage_mark(c, 2); if (st) return st; }
        const unsigned hx = (unsigned)std::min<unsigned long long>(((max_mc >> hlog) + 3 + 255) / 256, std::max<unsigned long long>(1, (unsigned long long)c->sm_count * 16 / nb));
        RCZ_KLAUNCH(c, ibwt_heads_kernel, dim3(hx, nb), 256, 0, dblk, desc, chain_ctr, nodes, hlog);
        RCZ_KLAUNCH(c, ibwt_headrank_kernel, nb, RANK_NT, rank_smem, dblk, chain_ctr, nodes, d_len, d_st, hlog);
        const unsigned tiles = ((max_mc >> hlog) + 3 + 255) / 256;
        const unsigned place_grid = std::min<unsigned>(nb * tiles, (unsigned)c->sm_count * tune_place);
        RCZ_KLAUNCH(c, ibwt_place_kernel, place_grid, 256, 0, dblk, nb, tiles, (const uint8_t*)wS, desc, chain_ctr, nodes, dout, hlog, queue + 1);
        if (mark) { st = ctx_stage_mark(c, 3); if (st) return st; mark = false; }
    }
    st = ctx_timer_end(c); if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) { st = unstage_span_out(c, out_base, dout, out_off, out_len, nblocks, 1); if (st) return st; }
    return RCZ_OK;
}
