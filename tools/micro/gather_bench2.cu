// gather_bench2.cu — dependent random 4-byte gathers over a TRUE random single-cycle permutation (Sattolo), the access pattern of the
// inverse-BWT walk, vs working-set size and threads in flight.  gather_bench.cu used an LCG permutation, whose structure flatters the
// cache; this one is the honest ceiling for `ibwt_walk_kernel`.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench2 gather_bench2.cu ; run on the B200 box.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <cuda_runtime.h>
__global__ void walk(const unsigned* __restrict__ P, unsigned n, unsigned hops, unsigned* out) {
    unsigned cur = (unsigned)(((unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x) * 2654435761ull) % n);
    unsigned acc = 0;
    for (unsigned i = 0; i < hops; ++i) {
        const unsigned e = __ldcg(P + cur);
        acc += e & 255u;
        cur = e >> 8;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    const size_t maxn = 1u << 24;   // entries are (next << 8) | byte: next < 2^24 => up to 64 MiB per table; larger sets = several tables side by side
    std::vector<unsigned> h(maxn);
    unsigned *P, *out;
    cudaMalloc(&P, maxn * 4 * 4); cudaMalloc(&out, 1 << 24);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    std::mt19937_64 rng(12345);
    for (int lg = 21; lg <= 24; ++lg) {                       // working set 8 MiB .. 64 MiB
        const unsigned n = 1u << lg;
        std::vector<unsigned> perm(n);
        for (unsigned i = 0; i < n; ++i) perm[i] = i;
        for (unsigned i = n - 1; i > 0; --i) { unsigned j = (unsigned)(rng() % i); std::swap(perm[i], perm[j]); }   // Sattolo: one cycle
        for (unsigned i = 0; i < n; ++i) h[i] = (perm[i] << 8) | (i & 255u);
        cudaMemcpy(P, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
        for (int ctas = 2; ctas <= 8; ctas *= 2) {
            const unsigned hops = 2000;
            const int grid = 148 * ctas, block = 256;
            walk<<<grid, block>>>(P, n, 200, out);
            cudaEventRecord(a);
            walk<<<grid, block>>>(P, n, hops, out);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("random cycle  set %4d MiB  threads %7d  %.1f Ghops/s  (%.0f ns per hop per thread)\n", (4 << lg) >> 20, grid * block,
                   (double)grid * block * hops / ms / 1e6, ms * 1e6 / hops);
        }
    }
    return 0;
}
