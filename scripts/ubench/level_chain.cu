// Microbenchmark of the per-level dependency chain of the single-query search kernels (tuning aid, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/level_chain scripts/ubench/level_chain.cu
// Prints cycles per iteration for: (a) block barrier only, (b) cluster barrier only, (c) RED + cluster barrier,
// (d) RED by thread t, barrier, ld.cg of the word thread t+1 reduced (the field hand-over of a level), one CTA,
// (e) the same across a cluster, (f) e + a remote shared-memory store (the broadcast), (g) st.cg + barrier + ld.cg (queue hand-over)
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void red_min(unsigned *p, unsigned v) { asm volatile("red.relaxed.gpu.global.min.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(unsigned *buf, long long *out, int iters, int stride)
{
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ unsigned s_word[8];
    const unsigned rank = cluster.block_rank(), nct = cluster.num_blocks();
    const unsigned ctid = rank * blockDim.x + threadIdx.x, NT = nct * blockDim.x;
    unsigned *remote = cluster.map_shared_rank(s_word, (rank + 1) % nct);
    unsigned acc = 0;
    cluster.sync();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        unsigned *mine = buf + (size_t)((ctid + (unsigned)it * 7919u) % NT) * stride;
        unsigned *next = buf + (size_t)(((ctid + 1) % NT + (unsigned)it * 7919u) % NT) * stride;
        if (MODE == 2 || MODE == 3 || MODE == 4 || MODE == 5) red_min(mine, 0x7FFFFFFFu - (unsigned)it);
        if (MODE == 6) __stcg(mine, (unsigned)it);
        if (MODE == 5 && threadIdx.x < 8) remote[rank] = (unsigned)it;
        if (MODE == 0 || MODE == 3 || MODE == 6) __syncthreads();
        else cluster.sync();
        if (MODE == 3 || MODE == 4 || MODE == 5 || MODE == 6) acc += __ldcg(next);
        if (MODE == 5) acc += s_word[(rank + nct - 1) % nct];
        // make the next iteration depend on what was loaded (as the search does: the loaded word decides what is pushed)
        if (acc == 0xDEADBEEFu) buf[0] = acc;
    }
    const long long t1 = clock64();
    if (ctid == 0) out[0] = (t1 - t0) / iters;
    if (acc == 0x12345678u) out[1] = acc;
}

template <int MODE>
static void run(const char *name, int cl, unsigned *buf, long long *out, int stride)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cl, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaMemset(buf, 0xFF, (size_t)8 * 256 * stride * 4 + 4096);
    const int iters = 2000;
    cudaLaunchKernelEx(&cfg, k<MODE>, buf, out, iters, stride);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-70s cluster %d stride %4d B: %6lld cycles/iter  (%s)\n", name, cl, stride * 4, h, cudaGetErrorString(e));
}

int main()
{
    unsigned *buf; long long *out;
    cudaMalloc(&buf, (size_t)8 * 256 * 1024 * 4 + 4096);
    cudaMalloc(&out, 64);
    for (int stride : {1, 32, 1024}) {
        run<0>("(a) __syncthreads only", 1, buf, out, stride);
        run<1>("(b) cluster.sync only", 8, buf, out, stride);
        run<2>("(c) RED + cluster.sync", 8, buf, out, stride);
        run<3>("(d) RED, __syncthreads, ld.cg of the neighbour's word (one CTA)", 1, buf, out, stride);
        run<4>("(e) RED, cluster.sync, ld.cg of the neighbour's word (cluster)", 8, buf, out, stride);
        run<5>("(f) e + remote shared-memory store / local read", 8, buf, out, stride);
        run<6>("(g) st.cg, __syncthreads, ld.cg (one CTA)", 1, buf, out, stride);
        run<1>("(b2) cluster.sync only, cluster of 2", 2, buf, out, stride);
        run<4>("(e2) RED, cluster.sync, ld.cg, cluster of 2", 2, buf, out, stride);
        run<4>("(e4) RED, cluster.sync, ld.cg, cluster of 4", 4, buf, out, stride);
    }
    return 0;
}
