// assemble.cu -- planner-side grid assembly on the device: OccupancyGrid decode / encode, bounding box of
// the occupied cells (crop), rectangle paste (pad / shift / pre-map merge), goal relocation, and the fused host-buffer
// replan call.  Integer / byte work, HBM-bound (2 B per cell) or latency-bound (one row scan).
#include <math.h>

#include "common.cuh"

// ---- OccupancyGrid message <-> [x][y] array ----------------------------------------------------------------------
// Replaces map_callback, scripts/global_planner_st.py:15-20 (= global_planner_ccst.py:17-24):
//   np.array(data).reshape(height, width).T ; 100 -> 1 ; -1 -> 0 ; everything else unchanged
// fused with the paste `mapu0[d0:d0+w, d1:d1+h] = mapu[sx0:sx0+w, sy0:sy0+h]` of st:249-250 (source window = the crop
// of ccst:36-63 when there is one).  32x32 tiles through shared memory: reads run along x (the message's fast axis),
// writes along y (the array's fast axis).
__global__ void __launch_bounds__(256)
k_decode_paste(const int8_t *__restrict__ msg, int width, int height, int sx0, int sy0, int w, int h,
               uint8_t *__restrict__ dst, int dW, int dH, int px, int py)
{
    __shared__ uint8_t tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;    // window coordinates of the tile
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int wx = bx + tx, wy = by + ty + 8 * k;
        uint8_t v = 0;
        if (wx < w && wy < h) {
            const int mx = sx0 + wx, my = sy0 + wy;
            if (mx >= 0 && mx < width && my >= 0 && my < height) {
                const int8_t s = msg[(size_t)my * width + mx];
                v = s == 100 ? (uint8_t)1 : s == -1 ? (uint8_t)0 : (uint8_t)s;
            }
        }
        tile[ty + 8 * k][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int wx = bx + ty + 8 * k, wy = by + tx;
        if (wx < w && wy < h) {
            const int ox = px + wx, oy = py + wy;
            if (ox >= 0 && ox < dW && oy >= 0 && oy < dH) dst[(size_t)ox * dH + oy] = tile[tx][ty + 8 * k];
        }
    }
}

// publish_map, scripts/global_planner_st.py:102-115: 1 -> 100, data = mapu.T.flatten() (int8)
__global__ void __launch_bounds__(256)
k_encode(const uint8_t *__restrict__ grid, int W, int H, int8_t *__restrict__ msg)
{
    __shared__ uint8_t tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;  // bx: x of the grid, by: y of the grid
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = bx + ty + 8 * k, y = by + tx;
        tile[ty + 8 * k][tx] = (x < W && y < H) ? grid[(size_t)x * H + y] : (uint8_t)0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = bx + tx, y = by + ty + 8 * k;
        if (x < W && y < H) {
            const uint8_t v = tile[tx][ty + 8 * k];
            msg[(size_t)y * W + x] = v == 1 ? (int8_t)100 : (int8_t)v;
        }
    }
}

// rectangle paste between [x][y] byte arrays: dst[px + i][py + j] = src[sx0 + i][sy0 + j], overwrite (zeros included),
// exactly like the numpy slice assignments of st:210-224 (pre-map merge) and st:249-250 (pad / shift)
__global__ void __launch_bounds__(256)
k_paste(const uint8_t *__restrict__ src, int sW, int sH, int sx0, int sy0, int w, int h, uint8_t *__restrict__ dst, int dW, int dH,
        int px, int py)
{
    const size_t total = (size_t)w * h;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int wx = (int)(i / h), wy = (int)(i % h);
        const int sx = sx0 + wx, sy = sy0 + wy, ox = px + wx, oy = py + wy;
        if (sx >= 0 && sx < sW && sy >= 0 && sy < sH && ox >= 0 && ox < dW && oy >= 0 && oy < dH)
            dst[(size_t)ox * dH + oy] = src[(size_t)sx * sH + sy];
    }
}

// ---- bounding box of the non-zero cells ---------------------------------------------------------------------------
// Replaces the X.nonzero() / np.unique / min / max part of remove_zero_rowscols, scripts/global_planner_ccst.py:42-49.
// bbox = {min x, max x, min y, max y}; {INT_MAX, -1, INT_MAX, -1} when the grid is all zero.  `msg` form reads the
// OccupancyGrid message directly (x = index % width, y = index / width; value != 0 after the 100/-1 mapping <=> raw
// value not in {0, -1}).
__global__ void k_bbox_init(int32_t *bbox)
{
    bbox[0] = 0x7FFFFFFF; bbox[1] = -1; bbox[2] = 0x7FFFFFFF; bbox[3] = -1;
}
template <bool MSG>
__global__ void __launch_bounds__(256)
k_bbox(const uint8_t *__restrict__ a, int n_slow, int n_fast, int32_t *bbox)
{
    // element (s, f) at a[s * n_fast + f]; MSG: s = y, f = x (raw int8); else s = x, f = y
    int smin = 0x7FFFFFFF, smax = -1, fmin = 0x7FFFFFFF, fmax = -1;
    const size_t total = (size_t)n_slow * n_fast;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t v = a[i];
        const bool nz = MSG ? (v != 0 && v != 0xFF) : (v != 0);
        if (nz) {
            const int s = (int)(i / n_fast), f = (int)(i % n_fast);
            smin = min(smin, s); smax = max(smax, s); fmin = min(fmin, f); fmax = max(fmax, f);
        }
    }
    for (int o = 16; o; o >>= 1) {
        smin = min(smin, __shfl_xor_sync(0xFFFFFFFFu, smin, o));
        smax = max(smax, __shfl_xor_sync(0xFFFFFFFFu, smax, o));
        fmin = min(fmin, __shfl_xor_sync(0xFFFFFFFFu, fmin, o));
        fmax = max(fmax, __shfl_xor_sync(0xFFFFFFFFu, fmax, o));
    }
    if ((threadIdx.x & 31) == 0 && smax >= 0) {
        // bbox = {min x, max x, min y, max y}: the message's slow axis is y, the array's is x
        atomicMin(&bbox[MSG ? 2 : 0], smin);
        atomicMax(&bbox[MSG ? 3 : 1], smax);
        atomicMin(&bbox[MSG ? 0 : 2], fmin);
        atomicMax(&bbox[MSG ? 1 : 3], fmax);
    }
}

// ---- goal relocation ----------------------------------------------------------------------------------------------
// Replaces scripts/global_planner_st.py:268-275 (= global_planner_ccst.py:454-464).  goal_io = {gx, gy} in, out =
// {gx, gy, moved, end_occu}: when grid[gx][gy] == 1 the goal moves to the nearest cell == 0 of its x-row (by |dy|, the
// lower y on ties -- np.argmin takes the first minimum), or, when the row has none, of its y-column; moved = -1 when
// neither exists (the reference raises there).  end_occu: st = moved; ccst = any cell == 1 in
// [gx-ifa, gx+ifa) x [gy-ifa, gy+ifa) around the final goal (ccst:460-463; empty when a lower bound is negative,
// a negative lower bound wraps around in numpy, which leaves an empty slice at these sizes).
__global__ void __launch_bounds__(256)
k_relocate_goal(const uint8_t *__restrict__ grid, int W, int H, int32_t *goal_io, int ifa, int ccst)
{
    __shared__ unsigned best;
    const int gx = goal_io[0], gy = goal_io[1];
    int ox = gx, oy = gy, moved = 0;
    if (gx < 0 || gx >= W || gy < 0 || gy >= H) {
        if (threadIdx.x == 0) { goal_io[2] = -2; goal_io[3] = 0; }
        return;
    }
    if (grid[(size_t)gx * H + gy] == 1) {
        if (threadIdx.x == 0) best = 0xFFFFFFFFu;
        __syncthreads();
        unsigned mine = 0xFFFFFFFFu;
        for (int y = threadIdx.x; y < H; y += blockDim.x)
            if (grid[(size_t)gx * H + y] == 0) mine = min(mine, ((unsigned)abs(y - gy) << 16) | (unsigned)y);
        if (mine != 0xFFFFFFFFu) atomicMin(&best, mine);
        __syncthreads();
        unsigned b = best;
        __syncthreads();
        if (b != 0xFFFFFFFFu) {
            oy = (int)(b & 0xFFFFu);
            moved = 1;
        } else {
            mine = 0xFFFFFFFFu;
            for (int x = threadIdx.x; x < W; x += blockDim.x)
                if (grid[(size_t)x * H + gy] == 0) mine = min(mine, ((unsigned)abs(x - gx) << 16) | (unsigned)x);
            if (mine != 0xFFFFFFFFu) atomicMin(&best, mine);
            __syncthreads();
            b = best;
            if (b != 0xFFFFFFFFu) {
                ox = (int)(b & 0xFFFFu);
                moved = 1;
            } else
                moved = -1;
        }
    }
    int occ = moved == 1;
    if (ccst) {
        occ = 0;
        // Python slice semantics: a negative lower bound counts from the end (an empty slice unless the grid is tiny)
        int x0 = ox - ifa, y0 = oy - ifa;
        const int x1 = min(ox + ifa, W), y1 = min(oy + ifa, H);
        x0 = x0 < 0 ? max(x0 + W, 0) : x0;
        y0 = y0 < 0 ? max(y0 + H, 0) : y0;
        if (threadIdx.x == 0)
            for (int x = x0; x < x1; x++)
                for (int y = y0; y < y1; y++) occ |= grid[(size_t)x * H + y] == 1;
    }
    if (threadIdx.x == 0) {
        goal_io[0] = ox; goal_io[1] = oy; goal_io[2] = moved; goal_io[3] = occ;
    }
}

// ---- C ABI --------------------------------------------------------------------------------------------------------
extern "C" int fx_grid_decode(fx_context *ctx, const int8_t *msg, int width, int height, int sx0, int sy0, int w, int h,
                              uint8_t *dst, int dW, int dH, int px, int py, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!msg || !dst || width <= 0 || height <= 0 || w < 0 || h < 0 || dW <= 0 || dH <= 0)
        return fx_set_err(ctx, FX_ERR_ARG, "fx_grid_decode: bad argument");
    if (w == 0 || h == 0) return FX_OK;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    dim3 g((w + 31) / 32, (h + 31) / 32);
    k_decode_paste<<<g, 256, 0, (cudaStream_t)stream>>>(msg, width, height, sx0, sy0, w, h, dst, dW, dH, px, py);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}

extern "C" int fx_grid_encode(fx_context *ctx, const uint8_t *grid, int W, int H, int8_t *msg, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!grid || !msg || W <= 0 || H <= 0) return fx_set_err(ctx, FX_ERR_ARG, "fx_grid_encode: bad argument");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    dim3 g((W + 31) / 32, (H + 31) / 32);
    k_encode<<<g, 256, 0, (cudaStream_t)stream>>>(grid, W, H, msg);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}

// ---- map image <-> [x][y] array ---------------------------------------------------------
// Save block scripts/global_planner_st.py:368-372 (= global_planner_ccst.py:628-632): free (== 0) -> 255, anything
// else -> 0, then `.T[::-1]`: image row r, column c holds cell [x = c][y = H-1-r] (image is H rows x W columns).
// Load block st:176-182 (pre-map) / the fixture convention: cell = pixel > threshold ? 0 : 1, `img[::-1].T`.
// DIR 0: grid -> image, DIR 1: image -> grid.  32x32 tiles through shared memory, coalesced on both sides.
template <int DIR>
__global__ void __launch_bounds__(256)
k_grid_image(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int W, int H, int threshold)
{
    __shared__ uint8_t tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;  // bx: x of the grid / image column, by: y of the grid
    if (DIR == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = bx + ty + 8 * k, y = by + tx;
            tile[ty + 8 * k][tx] = (x < W && y < H) ? (src[(size_t)x * H + y] == 0 ? (uint8_t)255 : (uint8_t)0) : (uint8_t)0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = bx + tx, y = by + ty + 8 * k;
            if (x < W && y < H) dst[(size_t)(H - 1 - y) * W + x] = tile[tx][ty + 8 * k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = bx + tx, y = by + ty + 8 * k;
            tile[ty + 8 * k][tx] = (x < W && y < H) ? ((int)src[(size_t)(H - 1 - y) * W + x] > threshold ? (uint8_t)0 : (uint8_t)1) : (uint8_t)0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = bx + ty + 8 * k, y = by + tx;
            if (x < W && y < H) dst[(size_t)x * H + y] = tile[tx][ty + 8 * k];
        }
    }
}

extern "C" int fx_grid_to_image(fx_context *ctx, const uint8_t *grid, int W, int H, uint8_t *img, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!grid || !img || W <= 0 || H <= 0) return fx_set_err(ctx, FX_ERR_ARG, "fx_grid_to_image: bad argument");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    dim3 g((W + 31) / 32, (H + 31) / 32);
    k_grid_image<0><<<g, 256, 0, (cudaStream_t)stream>>>(grid, img, W, H, 0);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}

extern "C" int fx_image_to_grid(fx_context *ctx, const uint8_t *img, int W, int H, int threshold, uint8_t *grid, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!grid || !img || W <= 0 || H <= 0) return fx_set_err(ctx, FX_ERR_ARG, "fx_image_to_grid: bad argument");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    dim3 g((W + 31) / 32, (H + 31) / 32);
    k_grid_image<1><<<g, 256, 0, (cudaStream_t)stream>>>(img, grid, W, H, threshold);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}

extern "C" int fx_grid_paste(fx_context *ctx, const uint8_t *src, int sW, int sH, int sx0, int sy0, int w, int h, uint8_t *dst,
                             int dW, int dH, int px, int py, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!src || !dst || src == dst || sW <= 0 || sH <= 0 || dW <= 0 || dH <= 0 || w < 0 || h < 0)
        return fx_set_err(ctx, FX_ERR_ARG, "fx_grid_paste: bad argument");
    if (w == 0 || h == 0) return FX_OK;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t total = (size_t)w * h;
    const int blocks = (int)min((total + 255) / 256, (size_t)ctx->sm_count * 16);
    k_paste<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, sW, sH, sx0, sy0, w, h, dst, dW, dH, px, py);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}

extern "C" int fx_grid_bbox(fx_context *ctx, const void *a, int W, int H, int is_msg, int32_t *d_bbox4, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!a || !d_bbox4 || W <= 0 || H <= 0) return fx_set_err(ctx, FX_ERR_ARG, "fx_grid_bbox: bad argument");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    k_bbox_init<<<1, 1, 0, st>>>(d_bbox4);
    FX_LAUNCH_CHECK(ctx);
    const size_t total = (size_t)W * H;
    const int blocks = (int)min((total + 255) / 256, (size_t)ctx->sm_count * 16);
    if (is_msg)
        k_bbox<true><<<blocks, 256, 0, st>>>((const uint8_t *)a, H, W, d_bbox4);   // message: y slow, x fast
    else
        k_bbox<false><<<blocks, 256, 0, st>>>((const uint8_t *)a, W, H, d_bbox4);  // array: x slow, y fast
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}

extern "C" int fx_relocate_goal(fx_context *ctx, const uint8_t *grid, int W, int H, int32_t *d_goal4, int ifa, int ccst, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!grid || !d_goal4 || W <= 0 || H <= 0 || W > 65535 || H > 65535 || ifa < 0)
        return fx_set_err(ctx, FX_ERR_ARG, "fx_relocate_goal: bad argument");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    k_relocate_goal<<<1, 256, 0, (cudaStream_t)stream>>>(grid, W, H, d_goal4, ifa, ccst);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}
