// post.cu -- path post-processing of the ccmapping planner, one warp per path:
//   (1) near-vehicle point drop      scripts/global_planner_ccst.py:507-513
//   (2) greedy line-of-sight shortcutting around map_line_col   scripts/global_planner_ccst.py:258-283, 515-521
//   (3) cells -> world coordinates   scripts/global_planner_st.py:291-296 / global_planner_ccst.py:487-491
// All integer / IEEE double work with one rounding per operation (no FMA), so results are bit-identical to the
// reference's numpy arithmetic.
#include "common.cuh"

// map_line_col(p2 = b, p1 = a, mapu[min x:max x, min y:max y]) of the reference: True (clear) unless one of the
// samples -- one per integer x strictly between the end points, y = rint(slope * x) + int(y of the left end) in
// the frame of the bounding box -- is a cell == 1 INSIDE the half-open bounding box (the crop excludes the upper
// edges, so a sample that rounds onto y == max y is not seen; axis-aligned pairs have an empty crop).
__device__ __forceinline__ bool fx_line_clear(const uint8_t *__restrict__ grid, int W, int H, int ax, int ay, int bx, int by, int lane)
{
    const int x0 = min(ax, bx), y0 = min(ay, by);
    const int dx = abs(ax - bx), dy = abs(ay - by);
    if (dx < 2 || dy == 0) return true;
    const int ly = ((ax <= bx) ? ay : by) - y0;  // left end (the reference swaps so that p1 has the smaller x)
    const int ry = ((ax <= bx) ? by : ay) - y0;
    const double slope = __ddiv_rn((double)(ry - ly), (double)dx);
    bool hit = false;
    for (int x = 1 + lane; x < dx; x += 32) {
        const int y = (int)rint(__dmul_rn(slope, (double)x)) + ly;
        const int gx = x0 + x, gy = y0 + y;
        if (y >= 0 && y < dy && gx >= 0 && gx < W && gy >= 0 && gy < H) hit |= grid[(size_t)gx * H + gy] == 1;
    }
    return !__any_sync(0xFFFFFFFFu, hit);
}

struct PostArgs {
    int do_shortcut;
    int do_drop;
    double px, py, pz, radius;  // near-vehicle drop (world frame)
    double reso, ox, oy;        // world = (cell + off) * reso + origin
    int offx, offy;
};

__global__ void __launch_bounds__(128)
k_path_post(const uint8_t *__restrict__ grid, int W, int H, const int32_t *path_xy, const int32_t *path_len, int Q, int max_path,
            int32_t *out_xy, int32_t *out_len, double *out_world, PostArgs a)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q) return;
    const int32_t *in = path_xy + (size_t)q * max_path * 2;
    int32_t *out = out_xy + (size_t)q * max_path * 2;
    const int len0 = path_len[q];
    if (len0 <= 0) {
        if (lane == 0) out_len[q] = len0;
        return;
    }
    const int n = min(len0, max_path);
    // (1) copy / compact.  Point ii >= 1 is dropped when its world position (z = 0) is closer than `radius` to the
    // vehicle, provided the path has more than two points (ccst:507-513).
    const bool drop = a.do_drop && n > 2;
    int m = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        int cx = 0, cy = 0;
        bool keep = false;
        if (i < n) {
            cx = in[2 * i];
            cy = in[2 * i + 1];
            keep = true;
            if (drop && i >= 1) {
                const double wx = __dadd_rn(__dmul_rn((double)(cx + a.offx), a.reso), a.ox);
                const double wy = __dadd_rn(__dmul_rn((double)(cy + a.offy), a.reso), a.oy);
                const double ex = wx - a.px, ey = wy - a.py, ez = 0.0 - a.pz;
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
                keep = !(sqrt(d2) < a.radius);
            }
        }
        const unsigned kept = __ballot_sync(0xFFFFFFFFu, keep);
        __syncwarp();  // in may alias out: every lane has read its point before anyone writes
        if (keep) {
            const int p = m + __popc(kept & ((1u << lane) - 1u));
            out[2 * p] = cx;
            out[2 * p + 1] = cy;
        }
        m += __popc(kept);
    }
    __syncwarp();
    // (2) greedy shortcutting: `ii = 1; while ii < len-1: if clear(path[ii+1], path[ii-1]): delete ii else ii += 1`.
    // The list is out[0..top] (kept) followed by out[j..m-1] (not yet visited); deleting never moves the kept prefix.
    if (a.do_shortcut && m >= 3) {
        int top = 0, j = 1;
        int px = out[0], py = out[1];
        int mx = out[2], my = out[3];
        while (j < m - 1) {
            const int nx = out[2 * (j + 1)], ny = out[2 * (j + 1) + 1];
            if (!fx_line_clear(grid, W, H, px, py, nx, ny, lane)) {
                top++;
                if (lane == 0) {
                    out[2 * top] = mx;
                    out[2 * top + 1] = my;
                }
                px = mx;
                py = my;
            }
            mx = nx;
            my = ny;
            j++;
        }
        top++;
        if (lane == 0) {
            out[2 * top] = mx;
            out[2 * top + 1] = my;
        }
        m = top + 1;
        __syncwarp();
    }
    if (lane == 0) out_len[q] = m;  // a path longer than max_path is processed as the max_path points that were stored
    // (3) world coordinates of the surviving points, z = 0
    if (out_world) {
        double *w = out_world + (size_t)q * max_path * 3;
        for (int i = lane; i < m; i += 32) {
            w[3 * i] = __dadd_rn(__dmul_rn((double)(out[2 * i] + a.offx), a.reso), a.ox);
            w[3 * i + 1] = __dadd_rn(__dmul_rn((double)(out[2 * i + 1] + a.offy), a.reso), a.oy);
            w[3 * i + 2] = 0.0;
        }
    }
}

extern "C" int fx_path_post(fx_context *ctx, const uint8_t *grid, int W, int H, const int32_t *path_xy, const int32_t *path_len,
                            int Q, int max_path, int shortcut, const double *h_drop4, const double *h_world5,
                            int32_t *out_xy, int32_t *out_len, double *out_world, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (Q < 0 || max_path <= 0 || !path_xy || !path_len || !out_xy || !out_len || (shortcut && (!grid || W <= 0 || H <= 0)) ||
        ((h_drop4 || out_world) && !h_world5))
        return fx_set_err(ctx, FX_ERR_ARG, "fx_path_post: bad argument");
    if (Q == 0) return FX_OK;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    PostArgs a;
    memset(&a, 0, sizeof(a));
    a.do_shortcut = shortcut != 0;
    a.do_drop = h_drop4 != nullptr && h_drop4[3] > 0.0;
    if (h_drop4) { a.px = h_drop4[0]; a.py = h_drop4[1]; a.pz = h_drop4[2]; a.radius = h_drop4[3]; }
    if (h_world5) { a.reso = h_world5[0]; a.ox = h_world5[1]; a.oy = h_world5[2]; a.offx = (int)h_world5[3]; a.offy = (int)h_world5[4]; }
    const int wpb = 4;
    k_path_post<<<(Q + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(grid, W, H, path_xy, path_len, Q, max_path, out_xy, out_len,
                                                                         out_world, a);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}
