// simt.h — one source, two builds: nvcc (sm_100a, the product) and -DRCZ_EMU (g++, CPU fibers, tests only).
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef RCZ_EMU
#include "simt_emu.h"
#define __grid_constant__
#define RCZ_DYN_SMEM(name) unsigned char* name = emu::S().dyn_smem
#define RCZ_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (smem), [=]() { kern(__VA_ARGS__); })
#define RCZ_KERNEL_SMEM_OPTIN(kern, bytes) 0
#else
#include <cuda_runtime.h>
#define RCZ_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#define RCZ_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define RCZ_KERNEL_SMEM_OPTIN(kern, bytes) \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#endif

#define RCZ_FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------
// TMA 1-D bulk copy (cp.async.bulk -> SASS UBLKCP) + mbarrier, with an emulation stand-in.
// dst (shared) and src (global) must be 16-byte aligned, bytes a multiple of 16.
// ---------------------------------------------------------------------------------------------
#ifdef RCZ_EMU
struct rcz_mbar { uint64_t v; };
__device__ inline void mbar_init(rcz_mbar* b, unsigned) { b->v = 0; }
__device__ inline void mbar_fence_init() {}
__device__ inline void mbar_expect_tx(rcz_mbar*, unsigned) {}
// emulation: the copy completes immediately and flips the barrier's phase; waiters yield until then
__device__ inline void tma_load_1d(void* dst, const void* src, unsigned bytes, rcz_mbar* b) { memcpy(dst, src, bytes); b->v++; }
__device__ inline void mbar_wait(rcz_mbar* b, unsigned parity) { while ((b->v & 1) == parity) emu::yield(); }
__device__ inline void fence_proxy_async_smem() {}
// emulation: the bulk store completes immediately
__device__ inline void tma_store_1d(void* gdst, const void* ssrc, unsigned bytes) { memcpy(gdst, ssrc, bytes); }
__device__ inline void bulk_commit() {}
__device__ inline void bulk_wait_read0() {}
template <int N> __device__ inline void bulk_wait_read() {}
__device__ inline void bulk_wait0() {}
__device__ inline void prefetch_l2(const void*, unsigned) {}
__device__ inline void rcz_backoff(unsigned) {}
__device__ inline void rcz_bar_sync(unsigned id, unsigned count) { rcz_named_bar_emu(id, count); }
__device__ inline uint4 lds128_volatile(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ inline void sts128_volatile(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
// shared-memory addresses as values (32-bit shared-window addresses on the device, plain pointers in the emulation)
typedef uintptr_t rcz_saddr;
__device__ inline rcz_saddr saddr_of(const void* p) { return (uintptr_t)p; }
__device__ inline unsigned lds32_volatile(rcz_saddr a) { return *reinterpret_cast<const volatile unsigned*>(a); }
__device__ inline unsigned lds8_volatile(rcz_saddr a) { return *reinterpret_cast<const volatile uint8_t*>(a); }
__device__ inline unsigned long long lds64_volatile(rcz_saddr a) { return *reinterpret_cast<const volatile unsigned long long*>(a); }
__device__ inline void sts8_volatile(rcz_saddr a, unsigned v) { *reinterpret_cast<volatile uint8_t*>(a) = (uint8_t)v; }
#else
struct rcz_mbar { unsigned long long v; };
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(rcz_mbar* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(rcz_mbar* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, rcz_mbar* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(rcz_mbar* b, unsigned phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(b)),
        "r"(phase)
        : "memory");
}
// generic-proxy accesses to shared memory -> visible/ordered w.r.t. the async proxy (TMA) before re-filling a buffer
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// TMA 1-D bulk store shared -> global (SASS UBLKCP); both addresses 16-byte aligned, bytes a multiple of 16.
// Writers of the shared source must have executed fence_proxy_async_smem() + a barrier before the issue.
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores of this thread have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all but the N most recent bulk groups of this thread have finished reading shared memory
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// all committed bulk stores of this thread are complete (writes performed)
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// polling back-off: frees the issue slots of a warp that waits on a flag in shared memory
__device__ __forceinline__ void rcz_backoff(unsigned ns) { __nanosleep(ns); }
// named barrier: `count` threads (a multiple of 32) of the CTA meet on barrier `id` (1..15; 0 is __syncthreads)
__device__ __forceinline__ void rcz_bar_sync(unsigned id, unsigned count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// 16-byte shared-memory accesses that the compiler may neither cache nor reorder (racy-by-design pointer doubling)
__device__ __forceinline__ uint4 lds128_volatile(const void* p) {
    uint4 r;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(p)) : "memory");
    return r;
}
__device__ __forceinline__ void sts128_volatile(void* p, uint4 v) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
typedef unsigned rcz_saddr;
__device__ __forceinline__ rcz_saddr saddr_of(const void* p) { return smem_u32(p); }
__device__ __forceinline__ unsigned lds32_volatile(rcz_saddr a) { unsigned r; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(r) : "r"(a) : "memory"); return r; }
__device__ __forceinline__ unsigned lds8_volatile(rcz_saddr a) { unsigned r; asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(r) : "r"(a) : "memory"); return r; }
__device__ __forceinline__ unsigned long long lds64_volatile(rcz_saddr a) { unsigned long long r; asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(r) : "r"(a) : "memory"); return r; }
__device__ __forceinline__ void sts8_volatile(rcz_saddr a, unsigned v) { asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
#endif

// ---------------------------------------------------------------------------------------------
// lanes of the warp that hold the same 8-bit key as this lane (among the lanes with valid == true).  Eight ballots instead
// of __match_any_sync: constant cost, where MATCH.ANY serialises over the distinct values (~28 of 32 for high-entropy bytes).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned warp_match_u8(unsigned key, bool valid) {
    unsigned m = __ballot_sync(RCZ_FULL, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const unsigned bit = (key >> b) & 1u;
        const unsigned bal = __ballot_sync(RCZ_FULL, bit);
        m &= bit ? bal : ~bal;
    }
    return m;
}

// ---------------------------------------------------------------------------------------------
// warp / block scans (warp-shuffle prefix scans; the block level goes through one smem array)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned warp_incl_scan_add(unsigned v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned t = __shfl_up_sync(RCZ_FULL, v, d);
        if (lane >= (unsigned)d) v += t;
    }
    return v;
}
__device__ __forceinline__ unsigned warp_incl_scan_max(unsigned v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned t = __shfl_up_sync(RCZ_FULL, v, d);
        if (lane >= (unsigned)d) v = max(v, t);
    }
    return v;
}
__device__ __forceinline__ unsigned warp_reduce_add(unsigned v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(RCZ_FULL, v, d);
    return v;
}
// Exclusive block scan of one value per thread. scratch: >= 33 unsigned in shared memory.
// All threads of the block must call; contains two __syncthreads. Returns exclusive prefix; *total = block sum.
template <int NT>
__device__ __forceinline__ unsigned block_excl_scan_add(unsigned v, unsigned* scratch, unsigned* total) {
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned incl = warp_incl_scan_add(v);
    if (lane == 31) scratch[w] = incl;
    __syncthreads();
    if (w == 0) {
        unsigned x = lane < NT / 32 ? scratch[lane] : 0;
        unsigned s = warp_incl_scan_add(x);
        scratch[lane] = s - x;
        if (lane == 31) scratch[32] = s;
    }
    __syncthreads();
    unsigned r = incl - v + scratch[w];
    *total = scratch[32];
    return r;
}
