// simt_emu.h — CPU emulation of the CUDA SIMT execution model (TEST HARNESS ONLY).
//
// Compiling the kernels with -DRCZ_EMU (g++ -x c++) turns every CTA into a set of cooperatively
// scheduled fibers so that the exact kernel source (barriers, warp shuffles/ballots, shared memory,
// atomics) can be exercised on a machine without a GPU.  It backs tests/ (`librcz_emu.so`); the
// product library `librcz.so` is always the nvcc build and never contains this file.
#pragma once
#ifndef RCZ_EMU
#error "simt_emu.h is only for the -DRCZ_EMU test build"
#endif
#include <sys/mman.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };

namespace emu {

extern "C" void rcz_emu_switch(void** from_sp, void* to_sp);

struct Fiber { void* sp; char* stack; bool done; unsigned tid; };
struct WarpState {
    unsigned count = 0, gen = 0;
    uint64_t vals[2][32];
    unsigned ballot[2] = {0, 0};
};
struct State {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    std::function<void()> body;
    void* main_sp = nullptr;
    Fiber* cur = nullptr;
    unsigned nthreads = 0;
    unsigned bar_count = 0, bar_gen = 0;
    int bar_or[2] = {0, 0}, bar_cnt[2] = {0, 0};
    unsigned nb_count[16] = {0}, nb_gen[16] = {0};   // named barriers (bar.sync id, count)
    unsigned char* dyn_smem = nullptr;
    size_t dyn_smem_cap = 0;
    char* stacks = nullptr;
    size_t stack_bytes = 256 * 1024, nstacks = 0;
    uint64_t launches = 0;
};
inline State& S() { static State s; return s; }

}  // namespace emu

// CUDA builtin variables
inline uint3_emu threadIdx, blockIdx;
inline dim3 blockDim, gridDim;
inline constexpr int warpSize = 32;

namespace emu {

inline void yield() { State& s = S(); rcz_emu_switch(&s.cur->sp, s.main_sp); }

inline void fiber_entry() {
    State& s = S();
    s.body();
    s.cur->done = true;
    rcz_emu_switch(&s.cur->sp, s.main_sp);
    abort();
}

inline void run_block(unsigned nthreads) {
    State& s = S();
    if (s.nstacks < nthreads) {
        if (s.stacks) munmap(s.stacks, s.nstacks * s.stack_bytes);
        s.nstacks = std::max<size_t>(nthreads, 1024);
        s.stacks = (char*)mmap(nullptr, s.nstacks * s.stack_bytes, PROT_READ | PROT_WRITE,
                               MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (s.stacks == MAP_FAILED) { perror("mmap"); abort(); }
    }
    s.nthreads = nthreads;
    s.bar_count = 0; s.bar_gen = 0; s.bar_or[0] = s.bar_or[1] = 0; s.bar_cnt[0] = s.bar_cnt[1] = 0;
    for (int i = 0; i < 16; ++i) { s.nb_count[i] = 0; s.nb_gen[i] = 0; }
    s.fibers.assign(nthreads, Fiber{});
    s.warps.assign((nthreads + 31) / 32, WarpState{});
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber& f = s.fibers[t];
        f.stack = s.stacks + (size_t)t * s.stack_bytes;
        f.tid = t; f.done = false;
        uintptr_t top = ((uintptr_t)f.stack + s.stack_bytes) & ~(uintptr_t)15;
        void** p = (void**)(top - 16);
        p[0] = (void*)&fiber_entry;               // return address consumed by `ret` in rcz_emu_switch
        for (int i = 1; i <= 6; ++i) p[-i] = nullptr;   // rbp rbx r12 r13 r14 r15
        f.sp = (void*)(p - 6);
    }
    // Scheduling order of the fibers inside one sweep: RCZ_EMU_ORDER=0 ascending (default), 1 descending, 2 pseudo-random
    // per sweep.  Kernels must not depend on it; the tests run the dependency-sensitive ones under all three.
    static const int order_mode = getenv("RCZ_EMU_ORDER") ? atoi(getenv("RCZ_EMU_ORDER")) : 0;
    static uint64_t lcg = 0x9E3779B97F4A7C15ull;
    unsigned remaining = nthreads;
    while (remaining) {
        unsigned progressed = 0;
        unsigned mul = 1, add = 0;
        if (order_mode == 2) {                                   // t -> (t * mul + add) mod 2^k is a permutation of a power-of-two range for odd mul
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            mul = (unsigned)(lcg >> 33) | 1u; add = (unsigned)(lcg >> 13);
        }
        unsigned pow2 = 1; while (pow2 < nthreads) pow2 <<= 1;
        for (unsigned tt = 0; tt < pow2; ++tt) {
            unsigned t = order_mode == 1 ? pow2 - 1 - tt : order_mode == 2 ? ((tt * mul + add) & (pow2 - 1)) : tt;
            if (t >= nthreads) continue;
            Fiber& f = s.fibers[t];
            if (f.done) continue;
            s.cur = &f;
            threadIdx.x = t % blockDim.x;
            threadIdx.y = (t / blockDim.x) % blockDim.y;
            threadIdx.z = t / (blockDim.x * blockDim.y);
            rcz_emu_switch(&s.main_sp, f.sp);
            if (f.done) { --remaining; ++progressed; }
        }
        (void)progressed;
    }
}

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F&& f) {
    State& s = S();
    s.launches++;
    if (smem > s.dyn_smem_cap) {
        free(s.dyn_smem);
        s.dyn_smem_cap = std::max<size_t>(smem, 256 * 1024);
        s.dyn_smem = (unsigned char*)aligned_alloc(1024, s.dyn_smem_cap);
    }
    gridDim = grid; blockDim = block;
    s.body = std::function<void()>(f);
    unsigned nthreads = block.x * block.y * block.z;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                run_block(nthreads);
            }
}

inline unsigned lin_tid() { return S().cur->tid; }
inline unsigned lane() { return lin_tid() & 31; }
inline WarpState& warp() { return S().warps[lin_tid() >> 5]; }
inline unsigned warp_width() {
    State& s = S();
    unsigned w = lin_tid() >> 5;
    unsigned n = s.nthreads - w * 32;
    return n < 32 ? n : 32;
}
// warp-level rendezvous of the lanes in `mask`; returns the generation index the op used
inline unsigned warp_arrive_wait(unsigned mask) {
    WarpState& w = warp();
    unsigned want = __builtin_popcount(mask);
    unsigned ww = warp_width();
    if (want > ww) want = ww;
    unsigned g = w.gen;
    if (++w.count == want) { w.count = 0; w.ballot[(g + 1) & 1] = 0; ++w.gen; }
    else while (w.gen == g) yield();
    return g;
}

}  // namespace emu

// ---------------- CUDA keyword shims ----------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __constant__ static const

// ---------------- barriers ----------------
inline void __syncthreads() {
    emu::State& s = emu::S();
    unsigned g = s.bar_gen;
    if (++s.bar_count == s.nthreads) { s.bar_count = 0; s.bar_or[(g + 1) & 1] = 0; s.bar_cnt[(g + 1) & 1] = 0; ++s.bar_gen; }
    else while (s.bar_gen == g) emu::yield();
}
// bar.sync id, count: rendezvous of `count` threads of the CTA on barrier `id`
inline void rcz_named_bar_emu(unsigned id, unsigned count) {
    emu::State& s = emu::S();
    unsigned g = s.nb_gen[id];
    if (++s.nb_count[id] == count) { s.nb_count[id] = 0; ++s.nb_gen[id]; }
    else while (s.nb_gen[id] == g) emu::yield();
}
inline int __syncthreads_or(int p) {
    emu::State& s = emu::S();
    unsigned g = s.bar_gen;
    s.bar_or[g & 1] |= (p != 0);
    __syncthreads();
    return s.bar_or[g & 1];
}
inline int __syncthreads_count(int p) {
    emu::State& s = emu::S();
    unsigned g = s.bar_gen;
    s.bar_cnt[g & 1] += (p != 0);
    __syncthreads();
    return s.bar_cnt[g & 1];
}
inline int __syncthreads_and(int p) { return __syncthreads_count(p) == (int)emu::S().nthreads; }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_arrive_wait(mask); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline unsigned __activemask() { unsigned w = emu::warp_width(); return w == 32 ? 0xffffffffu : ((1u << w) - 1); }

// ---------------- warp collectives ----------------
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    static_assert(sizeof(T) <= 8, "shfl");
    emu::WarpState& w = emu::warp();
    unsigned l = emu::lane(), g = w.gen;
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    w.vals[g & 1][l] = raw;
    emu::warp_arrive_wait(mask);
    unsigned base = l & ~(unsigned)(width - 1);
    unsigned s = base + ((unsigned)src & (unsigned)(width - 1));
    T r; memcpy(&r, &w.vals[g & 1][s], sizeof(T));
    return r;
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    unsigned l = emu::lane();
    unsigned base = l & ~(unsigned)(width - 1);
    T r = __shfl_sync(mask, v, (int)(l >= base + d ? l - d : l), 32);
    return (l >= base + d) ? r : v;
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    unsigned l = emu::lane();
    unsigned base = l & ~(unsigned)(width - 1);
    bool ok = l + d < base + (unsigned)width && l + d < emu::warp_width();
    T r = __shfl_sync(mask, v, (int)(ok ? l + d : l), 32);
    return ok ? r : v;
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    (void)width;
    unsigned l = emu::lane();
    return __shfl_sync(mask, v, (int)(l ^ (unsigned)x), 32);
}
inline unsigned __ballot_sync(unsigned mask, int p) {
    emu::WarpState& w = emu::warp();
    unsigned g = w.gen;
    if (p) w.ballot[g & 1] |= 1u << emu::lane();
    emu::warp_arrive_wait(mask);
    return w.ballot[g & 1] & mask;
}
inline int __any_sync(unsigned mask, int p) { return __ballot_sync(mask, p) != 0; }
inline int __all_sync(unsigned mask, int p) { return __ballot_sync(mask, !p) == 0; }
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) {
    emu::WarpState& w = emu::warp();
    unsigned l = emu::lane(), g = w.gen;
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    w.vals[g & 1][l] = raw;
    emu::warp_arrive_wait(mask);
    unsigned m = 0, ww = emu::warp_width();
    for (unsigned i = 0; i < ww; ++i) if ((mask >> i & 1) && w.vals[g & 1][i] == raw) m |= 1u << i;
    return m;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    emu::WarpState& w = emu::warp();
    unsigned l = emu::lane(), g = w.gen;
    w.vals[g & 1][l] = v;
    emu::warp_arrive_wait(mask);
    unsigned s = 0, ww = emu::warp_width();
    for (unsigned i = 0; i < ww; ++i) if (mask >> i & 1) s += (unsigned)w.vals[g & 1][i];
    return s;
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    emu::WarpState& w = emu::warp();
    unsigned l = emu::lane(), g = w.gen;
    w.vals[g & 1][l] = v;
    emu::warp_arrive_wait(mask);
    unsigned s = 0, ww = emu::warp_width();
    for (unsigned i = 0; i < ww; ++i) if (mask >> i & 1) s = std::max(s, (unsigned)w.vals[g & 1][i]);
    return s;
}
inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    emu::WarpState& w = emu::warp();
    unsigned l = emu::lane(), g = w.gen;
    w.vals[g & 1][l] = v;
    emu::warp_arrive_wait(mask);
    unsigned s = 0xffffffffu, ww = emu::warp_width();
    for (unsigned i = 0; i < ww; ++i) if (mask >> i & 1) s = std::min(s, (unsigned)w.vals[g & 1][i]);
    return s;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
    emu::WarpState& w = emu::warp();
    unsigned l = emu::lane(), g = w.gen;
    w.vals[g & 1][l] = v;
    emu::warp_arrive_wait(mask);
    unsigned s = 0, ww = emu::warp_width();
    for (unsigned i = 0; i < ww; ++i) if (mask >> i & 1) s |= (unsigned)w.vals[g & 1][i];
    return s;
}

// ---------------- atomics (fibers never preempt, so plain RMW is atomic) ----------------
template <class T, class U> inline T atomicAdd(T* p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> inline T atomicSub(T* p, U v) { T o = *p; *p = (T)(o - (T)v); return o; }
template <class T, class U> inline T atomicMax(T* p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> inline T atomicMin(T* p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> inline T atomicOr(T* p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> inline T atomicAnd(T* p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U> inline T atomicExch(T* p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U, class V> inline T atomicCAS(T* p, U cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }

// ---------------- integer intrinsics ----------------
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)(v >> (s & 31)); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)((v << (s & 31)) >> 32); }
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    uint64_t v = ((uint64_t)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; ++i) { unsigned s = (sel >> (4 * i)) & 7; r |= (unsigned)((v >> (8 * s)) & 0xff) << (8 * i); }
    return r;
}
inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) { for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 255u) * ((b >> (8 * i)) & 255u); return c; }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
template <class T> inline void __stcg(T* p, T v) { *p = v; }
using std::max;
using std::min;

struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
