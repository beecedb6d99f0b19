// Microbenchmark: dependent random 4-byte accesses (load or atomicMin with return) from many CTAs, each inside its
// own window of a private region -- how does latency under load depend on the number of distinct 2 MB pages in use?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mem ubench_mem.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int OP>
__global__ void __launch_bounds__(128) k(uint32_t *base, size_t region_words, uint32_t window_words, int iters, unsigned long long *sink)
{
    uint32_t *r = base + (size_t)blockIdx.x * region_words;
    uint32_t s = hash32(blockIdx.x * 1315423911u + threadIdx.x);
    uint32_t acc = 0;
    for (int i = 0; i < iters; i++) {
        // lanes of a warp hit neighbouring sectors (like a wavefront): one random 1 KB chunk per warp, lane picks a word in it
        uint32_t chunk = hash32(s + (threadIdx.x >> 5) * 7919u + i * 2654435761u + acc) % (window_words / 256);
        uint32_t w = chunk * 256 + ((threadIdx.x & 31) * 8 + (hash32(s + i) & 7));
        uint32_t v;
        if (OP == 0) v = __ldcg(r + w);
        else v = atomicMin(r + w, 0xFFFFFFF0u - (uint32_t)i);
        acc = __shfl_sync(0xFFFFFFFFu, v, 0) & 1u;  // dependent chain (value-dependent next address, warp-uniform part)
        s += v & 1u;
    }
    if (acc == 12345u) sink[0] = s;
}
int main(int argc, char **argv)
{
    int ctas = argc > 1 ? atoi(argv[1]) : 888;
    size_t region_mb = argc > 2 ? atoi(argv[2]) : 64;
    int iters = 2000;
    size_t region_words = region_mb * (1 << 20) / 4;
    uint32_t *buf; unsigned long long *sink;
    if (cudaMalloc(&buf, (size_t)ctas * region_words * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 0xFF, (size_t)ctas * region_words * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    printf("ctas %d region %zu MB total %.1f GB\n", ctas, region_mb, ctas * region_mb / 1024.0);
    for (int op = 0; op < 2; op++)
        for (size_t wkb : {64, 512, 2048, 8192, 32768, 65536}) {
            if (wkb * 1024 > region_mb << 20) continue;
            uint32_t ww = (uint32_t)(wkb * 1024 / 4);
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(a);
                if (op == 0) k<0><<<ctas, 128>>>(buf, region_words, ww, iters, sink); else k<1><<<ctas, 128>>>(buf, region_words, ww, iters, sink);
                cudaEventRecord(b); cudaEventSynchronize(b);
            }
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("%s window %6zu KB  pages~%6zu  %.3f ms  latency/op %.0f ns  %.2f G lane-ops/s\n", op ? "atomicMin" : "ldcg     ", wkb,
                   (size_t)ctas * ((wkb + 2047) / 2048), ms, ms * 1e6 / iters, (double)ctas * 128 * iters / ms / 1e6);
        }
    return 0;
}
