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

// ---- jump-point form of a path -----------------------------------------------------------------------------------
// The reference returns the JUMP POINTS of its path (scripts/jps1.py:199-208: the came_from chain of A* over jump
// points), fx_search_batch the turning points.  Every jump point of a path lies on it, so the list the reference would
// return for the SAME cell path is recovered by walking each straight run and keeping the cells where jps1.jump
// (:95-164) would have stopped: the goal, a cell with a forced neighbour for the travel direction (:110-114, :134-138,
// :150-154 -- the same rule as fx_canon_succ), and on a diagonal run a cell from which one of the two straight
// sub-jumps finds a jump point (:116-118).  One warp per path, one lane per cell of a run.
__device__ __forceinline__ unsigned jp_moves_at(const uint8_t *__restrict__ grid, int W, int H, int x, int y)
{
    unsigned nb = 0;  // bit (a+1)*3 + (b+1): obstacle (== 1) or outside
#pragma unroll
    for (int a = -1; a <= 1; a++)
#pragma unroll
        for (int b = -1; b <= 1; b++) {
            const int xx = x + a, yy = y + b;
            bool blk = xx < 0 || xx >= W || yy < 0 || yy >= H;
            if (!blk) blk = __ldg(grid + (size_t)xx * H + yy) == 1;
            nb |= (blk ? 1u : 0u) << ((a + 1) * 3 + (b + 1));
        }
    unsigned m = 0;
#pragma unroll
    for (int d = 0; d < 8; d++) {
        const int a = fx_dx(d), b = fx_dy(d);
        bool ok = !((nb >> ((a + 1) * 3 + (b + 1))) & 1u);
        if (d >= 4) ok = ok && !(((nb >> ((a + 1) * 3 + 1)) & 1u) && ((nb >> (3 + (b + 1))) & 1u));
        m |= (ok ? 1u : 0u) << d;
    }
    return m;
}
// natural successors of a cell reached in direction d (jps1.py:59-92 without the forced ones)
__device__ __forceinline__ unsigned jp_natural(int d)
{
    if (d < 4) return 1u << d;
    const int dx = fx_dx(d), dy = fx_dy(d);
    return (1u << d) | (1u << fx_dir_of(dx, 0)) | (1u << fx_dir_of(0, dy));
}
__device__ __forceinline__ bool jp_forced(const uint8_t *__restrict__ grid, int W, int H, int x, int y, int d)
{
    return (fx_canon_succ((unsigned)d, jp_moves_at(grid, W, H, x, y)) & ~jp_natural(d)) != 0u;
}
// jps1.jump(c, straight direction e) != None: scanning from c along e, a forced-neighbour cell or the goal comes before
// an obstacle / the border
__device__ bool jp_subjump(const uint8_t *__restrict__ grid, int W, int H, int x, int y, int e, int gx, int gy)
{
    const int ex = fx_dx(e), ey = fx_dy(e);
    for (;;) {
        x += ex; y += ey;
        if (x < 0 || x >= W || y < 0 || y >= H || __ldg(grid + (size_t)x * H + y) == 1) return false;
        if (x == gx && y == gy) return true;
        if (jp_forced(grid, W, H, x, y, e)) return true;
    }
}

__global__ void __launch_bounds__(128) k_jump_points(const uint8_t *__restrict__ grid, int W, int H, const int32_t *__restrict__ path_xy,
                                                    const int32_t *__restrict__ path_len, int Q, int max_path, int32_t *__restrict__ out_xy,
                                                    int32_t *__restrict__ out_len, int max_out)
{
    const int lane = threadIdx.x & 31;
    for (int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < Q; q += gridDim.x * (blockDim.x >> 5)) {
        const int n = path_len[q];
        if (n <= 0 || n > max_path) { if (lane == 0) out_len[q] = n; continue; }
        const int32_t *p = path_xy + (size_t)q * max_path * 2;
        int32_t *o = out_xy + (size_t)q * max_out * 2;
        const int gx = p[2 * (n - 1)], gy = p[2 * (n - 1) + 1];
        int cnt = 1;
        if (lane == 0 && max_out > 0) { o[0] = p[0]; o[1] = p[1]; }
        for (int s = 0; s + 1 < n; s++) {
            const int ax = p[2 * s], ay = p[2 * s + 1], bx = p[2 * s + 2], by = p[2 * s + 3];
            const int L = max(abs(bx - ax), abs(by - ay));
            const int sx = (bx > ax) - (bx < ax), sy = (by > ay) - (by < ay);
            const int d = fx_dir_of(sx, sy);
            for (int i0 = 1; i0 <= L; i0 += 32) {
                const int i = i0 + lane;
                bool stop = false;
                if (i <= L) {
                    const int x = ax + i * sx, y = ay + i * sy;
                    stop = i == L || (x == gx && y == gy) || jp_forced(grid, W, H, x, y, d);
                    if (!stop && d >= 4)
                        stop = jp_subjump(grid, W, H, x, y, fx_dir_of(sx, 0), gx, gy) || jp_subjump(grid, W, H, x, y, fx_dir_of(0, sy), gx, gy);
                }
                const unsigned m = __ballot_sync(0xFFFFFFFFu, stop);
                if (stop) {
                    const int pos = cnt + __popc(m & ((1u << lane) - 1u));
                    if (pos < max_out) { o[2 * pos] = ax + i * sx; o[2 * pos + 1] = ay + i * sy; }
                }
                cnt += __popc(m);
            }
        }
        if (lane == 0) out_len[q] = cnt;
    }
}

extern "C" int fx_paths_jump_points(fx_context *ctx, const uint8_t *grid, int W, int H, const int32_t *path_xy, const int32_t *path_len,
                                    int Q, int max_path, int32_t *out_xy, int32_t *out_len, int max_out, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!grid || W <= 0 || H <= 0 || Q < 0 || max_path < 0 || max_out < 0 || (Q > 0 && (!path_xy || !path_len || !out_len)) ||
        (max_out > 0 && Q > 0 && !out_xy))
        return fx_set_err(ctx, FX_ERR_ARG, "fx_paths_jump_points: bad argument");
    if (Q == 0) return FX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    int blocks = (Q + 3) / 4;
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_jump_points<<<blocks, 128, 0, st>>>(grid, W, H, path_xy, path_len, Q, max_path, out_xy, out_len, max_out);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}

// host-buffer form for the drop-in: one path in, its jump points out.  h_grid == NULL reuses the grid the last
// fx_plan_host* call on this context uploaded (same W, H).
extern "C" int fx_jump_points_host(fx_context *ctx, const uint8_t *h_grid, int W, int H, const int32_t *h_path_xy, int n,
                                   int32_t *h_out_xy, int max_out, int32_t *h_out_len)
{
    if (!ctx) return FX_ERR_ARG;
    if (W <= 0 || H <= 0 || n <= 0 || !h_path_xy || !h_out_len || max_out <= 0 || !h_out_xy)
        return fx_set_err(ctx, FX_ERR_ARG, "fx_jump_points_host: bad argument");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->own_stream;
    const size_t cells = (size_t)W * H;
    int rc;
    if (h_grid) {
        if ((rc = fx_grow_bytes(ctx, (void **)&ctx->d_grid, &ctx->d_grid_cap, cells))) return rc;
        FX_CUDA(ctx, cudaMemcpyAsync(ctx->d_grid, h_grid, cells, cudaMemcpyHostToDevice, st));
    } else if (!ctx->d_grid || ctx->d_grid_cap < cells) {
        return fx_set_err(ctx, FX_ERR_ARG, "fx_jump_points_host: no grid on the device to reuse");
    }
    const size_t in_b = (size_t)n * 8, out_b = (size_t)max_out * 8;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->d_path, &ctx->d_path_cap, in_b + out_b + 64))) return rc;
    int32_t *d_in = ctx->d_path, *d_out = ctx->d_path + 2 * (size_t)n, *d_len = d_out + 2 * (size_t)max_out;
    FX_CUDA(ctx, cudaMemcpyAsync(d_in, h_path_xy, in_b, cudaMemcpyHostToDevice, st));
    FX_CUDA(ctx, cudaMemcpyAsync(d_len, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    if ((rc = fx_paths_jump_points(ctx, ctx->d_grid, W, H, d_in, d_len, 1, n, d_out, d_len + 1, max_out, (void *)st))) return rc;
    FX_CUDA(ctx, cudaMemcpyAsync(h_out_len, d_len + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    FX_CUDA(ctx, cudaStreamSynchronize(st));
    const int k = *h_out_len < max_out ? *h_out_len : max_out;
    if (k > 0) {
        FX_CUDA(ctx, cudaMemcpyAsync(h_out_xy, d_out, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
        FX_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return FX_OK;
}
