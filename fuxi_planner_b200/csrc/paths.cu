// paths.cu -- compact (CSR) form of the batched search's path output (sm_100a).
//
// fx_search_batch writes turning points into a padded [Q][max_path][2] array (the layout the reference's callers
// index, one row per query).  A path on the headline workload has tens of points, so copying the padded array to the
// host moves ~30x more bytes than the paths themselves (r01: 67 MB per 8192-query batch, the whole gap between the
// device-timed and the end-to-end figure).  fx_paths_compact packs the rows into offsets[Q+1] + xy[total][2] on the
// device; the host entry points copy only `total` points and scatter them into the caller's rows.
#include "common.cuh"

#define PC_THREADS 1024

// offsets[q] = sum of the lengths of the paths before q (a path that is missing, FX_COST_*, or did not fit max_path
// counts 0 points); offsets[Q] = total.  One CTA: each thread sums a contiguous chunk, one block scan, second sweep.
__global__ void __launch_bounds__(PC_THREADS) k_path_offsets(const int32_t *__restrict__ path_len, int Q, int max_path,
                                                            long long *__restrict__ offsets)
{
    __shared__ long long s_sum[PC_THREADS];
    const int tid = threadIdx.x;
    const int per = (Q + PC_THREADS - 1) / PC_THREADS;
    const int q0 = tid * per, q1 = min(Q, q0 + per);
    long long acc = 0;
    for (int q = q0; q < q1; q++) {
        const int n = path_len[q];
        acc += (n > 0 && n <= max_path) ? n : 0;
    }
    s_sum[tid] = acc;
    __syncthreads();
    for (int off = 1; off < PC_THREADS; off <<= 1) {  // inclusive Hillis-Steele scan
        long long v = tid >= off ? s_sum[tid - off] : 0;
        __syncthreads();
        s_sum[tid] += v;
        __syncthreads();
    }
    long long run = s_sum[tid] - acc;
    for (int q = q0; q < q1; q++) {
        offsets[q] = run;
        const int n = path_len[q];
        run += (n > 0 && n <= max_path) ? n : 0;
    }
    if (tid == PC_THREADS - 1) offsets[Q] = s_sum[tid];
}

// one warp per query: copy its points (8 bytes each) behind offsets[q]; points beyond cap are dropped
__global__ void __launch_bounds__(256) k_path_gather(const int32_t *__restrict__ path_xy, const int32_t *__restrict__ path_len, int Q,
                                                    int max_path, const long long *__restrict__ offsets, int2 *__restrict__ out, long long cap)
{
    const int lane = threadIdx.x & 31;
    for (int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < Q; q += gridDim.x * (blockDim.x >> 5)) {
        const int n = path_len[q];
        if (n <= 0 || n > max_path) continue;
        const long long o = offsets[q];
        const int2 *src = reinterpret_cast<const int2 *>(path_xy) + (size_t)q * max_path;
        for (int i = lane; i < n; i += 32)
            if (o + i < cap) out[o + i] = src[i];
    }
}

extern "C" int fx_paths_compact(fx_context *ctx, const int32_t *path_xy, const int32_t *path_len, int Q, int max_path,
                                int64_t *offsets, int32_t *out_xy, int64_t cap, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (Q < 0 || max_path < 0 || cap < 0 || !offsets || (Q > 0 && (!path_xy || !path_len)) || (cap > 0 && !out_xy))
        return fx_set_err(ctx, FX_ERR_ARG, "fx_paths_compact: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(long long) == sizeof(int64_t), "offsets are 64-bit");
    k_path_offsets<<<1, PC_THREADS, 0, st>>>(path_len, Q, max_path, reinterpret_cast<long long *>(offsets));
    FX_LAUNCH_CHECK(ctx);
    if (Q > 0 && cap > 0) {
        int blocks = (Q + 7) / 8;
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        k_path_gather<<<blocks, 256, 0, st>>>(path_xy, path_len, Q, max_path, reinterpret_cast<const long long *>(offsets),
                                              reinterpret_cast<int2 *>(out_xy), (long long)cap);
        FX_LAUNCH_CHECK(ctx);
    }
    return FX_OK;
}
