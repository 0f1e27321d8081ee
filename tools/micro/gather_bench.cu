// gather_bench.cu — dependent random 4-byte gathers (the inverse-BWT access pattern) vs working-set size and threads in flight.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu ; run on the B200 box.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
// full-period LCG on 2^lg elements: a single cycle through the whole table, so walks never collapse onto a short cycle
__global__ void fill(unsigned* P, unsigned mask) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i <= mask; i += (size_t)gridDim.x * blockDim.x)
        P[i] = ((((unsigned)i * 1664525u + 1013904223u) & mask) << 6) | ((unsigned)i & 63u);
}
__global__ void walk(const unsigned* __restrict__ P, unsigned mask, unsigned hops, unsigned* out, int mode) {
    unsigned cur = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u & mask;
    unsigned acc = 0;
    for (unsigned i = 0; i < hops; ++i) {
        unsigned e = mode == 0 ? __ldg(P + cur) : __ldcg(P + cur);
        acc += e & 63u;
        cur = (e >> 6) & mask;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    const size_t maxn = 1u << 26;   // 256 MiB of u32
    unsigned *P, *out;
    cudaMalloc(&P, maxn * 4); cudaMalloc(&out, 1 << 24);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 2; ++mode)
        for (int lg = 21; lg <= 26; ++lg)                       // working set 8 MiB .. 256 MiB
            for (int ctas = 1; ctas <= 8; ctas *= 2) {
                unsigned mask = (1u << lg) - 1, hops = 2000;
                fill<<<1184, 256>>>(P, mask);
                int grid = 148 * ctas, block = 256;
                walk<<<grid, block>>>(P, mask, 200, out, mode);
                cudaEventRecord(a);
                walk<<<grid, block>>>(P, mask, hops, out, mode);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                printf("mode %s  set %4d MiB  threads %7d  %.1f Ghops/s\n", mode ? "ldcg" : "ldg ", (4 << lg) >> 20, grid * block, (double)grid * block * hops / ms / 1e6);
            }
    return 0;
}
