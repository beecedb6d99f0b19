// replan.cu -- one global replan in one call: everything the planner main loops do between receiving
// an OccupancyGrid and publishing a path, scripts/global_planner_st.py:226-298 and scripts/global_planner_ccst.py:36-63,
// 411-521, with one host->device copy (the raw message), one device->host copy (the result) and one synchronisation.
//   crop (ccst) -> index / pad / shift -> decode + paste -> inflate -> goal relocation -> search -> near-vehicle drop +
//   line-of-sight shortcutting (ccst) -> world coordinates
// The scalar index arithmetic runs on the host in IEEE double exactly as the reference's numpy expressions.
#include <math.h>

#include "common.cuh"

extern "C" int fx_grid_decode(fx_context *, const int8_t *, int, int, int, int, int, int, uint8_t *, int, int, int, int, void *);
extern "C" int fx_grid_paste(fx_context *, const uint8_t *, int, int, int, int, int, int, uint8_t *, int, int, int, int, void *);
extern "C" int fx_grid_bbox(fx_context *, const void *, int, int, int, int32_t *, void *);
extern "C" int fx_relocate_goal(fx_context *, const uint8_t *, int, int, int32_t *, int, int, void *);
extern "C" int fx_path_post(fx_context *, const uint8_t *, int, int, const int32_t *, const int32_t *, int, int, int, const double *,
                            const double *, int32_t *, int32_t *, double *, void *);

// Small planning grids (the reference's own maps): the message is read straight out of the mapped pinned staging buffer
// and EVERY cell of the planning grid is written in one pass -- window cells decoded (scripts/global_planner_st.py:15-25:
// 100 -> 1, -1 -> 0) or copied, the padding zeroed -- instead of H2D copy + memset + tiled decode / paste.
__global__ void __launch_bounds__(256) k_assemble_small(const uint8_t *__restrict__ src, int layout_msg, int sW, int sH, int wx0, int wy0, int ww, int wh,
                                                        uint8_t *__restrict__ dst, int W, int H, int px, int py)
{
    const int total = W * H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int x = i / H, y = i - x * H;
        const int wx = x - px, wy = y - py;
        uint8_t v = 0;
        if (wx >= 0 && wx < ww && wy >= 0 && wy < wh) {
            const int mx = wx0 + wx, my = wy0 + wy;
            if (mx >= 0 && mx < sW && my >= 0 && my < sH) {
                if (layout_msg) {
                    const int8_t m = (int8_t)src[(size_t)my * sW + mx];
                    v = m == 100 ? (uint8_t)1 : m == -1 ? (uint8_t)0 : (uint8_t)m;
                } else {
                    v = src[(size_t)mx * sH + my];
                }
            }
        }
        dst[i] = v;
    }
}

// `((xy - origin) / reso).astype(int)`: double arithmetic, truncation toward zero
static inline long long cell_of(double v, double o, double reso) { return (long long)((v - o) / reso); }

// Python slice [lo:hi] on an axis of length n -> [a, b) with b >= a
static inline void py_slice(long long lo, long long hi, int n, int *a, int *b)
{
    if (lo < 0) lo = lo + n < 0 ? 0 : lo + n;
    if (lo > n) lo = n;
    if (hi < 0) hi = hi + n < 0 ? 0 : hi + n;
    if (hi > n) hi = n;
    if (hi < lo) hi = lo;
    *a = (int)lo;
    *b = (int)hi;
}

extern "C" int fx_replan_host(fx_context *ctx, const void *h_map, int width, int height, const fx_replan_in *in, fx_replan_out *out,
                              int32_t *h_path_xy, double *h_path_world, int max_path)
{
    if (!ctx) return FX_ERR_ARG;
    if (!h_map || !in || !out || width <= 0 || height <= 0 || max_path < 2 || in->ifa < 1 || (in->hchoice != 1 && in->hchoice != 2) ||
        !(in->reso > 0.0) || (in->variant != 0 && in->variant != 1) || (in->layout != 0 && in->layout != 1))
        return fx_set_err(ctx, FX_ERR_ARG, "fx_replan_host: bad argument");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->own_stream;
    memset(out, 0, sizeof(*out));
    const int ifa = in->ifa, ccst = in->variant == 1;
    const size_t msg_bytes = (size_t)width * height;
    int rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->d_msg, &ctx->d_msg_cap, msg_bytes))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->d_rp, &ctx->d_rp_cap, 256))) return rc;
    const size_t path_ints = (size_t)max_path * 2, world_bytes = (size_t)max_path * 3 * sizeof(double);
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->d_path, &ctx->d_path_cap, 2 * path_ints * sizeof(int32_t)))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->d_pts, &ctx->d_pts_cap, world_bytes))) return rc;
    // pinned staging: [message | result block 256 B | path | world]
    const size_t off_rp = (msg_bytes + 255) / 256 * 256, off_p = off_rp + 256, off_w = off_p + path_ints * sizeof(int32_t);
    if ((rc = fx_grow_pinned(ctx, off_w + world_bytes))) return rc;
    char *pin = (char *)ctx->h_pin;
    // Small messages stay in the pinned staging buffer, which the device reads in place (unified addressing), and the
    // parameter block / path / world points are written by the kernels straight into it: no copy-engine transfers at all
    // on the critical path of a replan (each one is a few microseconds of engine hand-over).
    const bool mapped = msg_bytes <= ((size_t)256 << 10) && !(getenv("FUXI_B200_REPLAN_MAPPED") && getenv("FUXI_B200_REPLAN_MAPPED")[0] == '0');
    const void *d_msg_in = ctx->d_msg;
    if (mapped) { memcpy(pin, h_map, msg_bytes); d_msg_in = pin; }
    else if ((rc = fx_staged_copy_in(ctx, (uint8_t *)ctx->d_msg, (const uint8_t *)h_map, msg_bytes, st))) return rc;
    int32_t *d_rp = mapped ? (int32_t *)(pin + off_rp) : ctx->d_rp;
    int32_t *d_bbox = ctx->d_rp /* reduced with atomics: always device memory */, *d_goal = d_rp + 4, *d_start = d_rp + 8, *d_len = d_rp + 10, *d_cost = d_rp + 11, *d_plen = d_rp + 12;
    double *d_costf = (double *)(d_rp + 32);

    // the map as the reference sees it: array [x][y] of extent (map_c, map_r) at world origin map_o; in the message
    // layout x is the fast axis (index y*width + x), in the array layout y is (index x*height + y), W = width, H = height
    const int aW = width, aH = height;
    double ox = in->origin_x, oy = in->origin_y;
    long long map_c = aW, map_r = aH;  // the extents the reference carries (differ from the window after a crop)
    int wx0 = 0, wx1 = aW, wy0 = 0, wy1 = aH;  // source window actually pasted
    if (in->crop) {
        // remove_zero_rowscols, ccst:36-63: window [min(min x, start x) : max x) x [min(min y, start y) : max y) --
        // exclusive upper ends, i.e. the last occupied row and column are dropped like in the reference
        if ((rc = fx_grid_bbox(ctx, d_msg_in, aW, aH, in->layout == 0, d_bbox, (void *)st))) return rc;
        int32_t bb[4];
        FX_CUDA(ctx, cudaMemcpyAsync(pin + off_rp, d_bbox, 16, cudaMemcpyDeviceToHost, st));
        FX_CUDA(ctx, cudaStreamSynchronize(st));
        memcpy(bb, pin + off_rp, 16);
        if (bb[1] < 0) {  // empty map: the reference's main loop skips the iteration (mapu is 0)
            out->skipped = 2;
            return FX_OK;
        }
        const long long s0x = cell_of(in->start_x, ox, in->reso), s0y = cell_of(in->start_y, oy, in->reso);
        const long long r0 = bb[0] < s0x ? bb[0] : s0x, c0 = bb[2] < s0y ? bb[2] : s0y;
        map_c = bb[1] - r0;
        map_r = bb[3] - c0;
        ox = (double)r0 * in->reso + ox;
        oy = (double)c0 * in->reso + oy;
        py_slice(r0, bb[1], aW, &wx0, &wx1);
        py_slice(c0, bb[3], aH, &wy0, &wy1);
        if (!(map_c * map_r > 0)) {  // ccst:352 `planner.map_c*planner.map_r > 0`
            out->skipped = 2;
            return FX_OK;
        }
    }
    // index / pad / shift, st:226-250 (ccst:411-436)
    long long gx = cell_of(in->goal_x, ox, in->reso), gy = cell_of(in->goal_y, oy, in->reso);
    long long sx = cell_of(in->start_x, ox, in->reso), sy = cell_of(in->start_y, oy, in->reso);
    long long o2x = -2 * ifa, o2y = -2 * ifa;
    if (gx < 0 || sx < 0) o2x += gx < sx ? gx : sx;
    if (gy < 0 || sy < 0) o2y += gy < sy ? gy : sy;
    const long long dx = -o2x, dy = -o2y;  // map_d = abs(map_o2), map_o2 <= -2 ifa
    const double nox = (double)o2x * in->reso + ox, noy = (double)o2y * in->reso + oy;
    long long mc = map_c > gx ? map_c : gx; if (sx > mc) mc = sx; mc += dx;
    long long mr = map_r > gy ? map_r : gy; if (sy > mr) mr = sy; mr += dy;
    const long long Wl = mc + 4 * ifa, Hl = mr + 4 * ifa;
    if (Wl > 32767 || Hl > 32767 || Wl * Hl > (1LL << 31))
        return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_replan_host: planning grid %lld x %lld too large", Wl, Hl);
    const int W = (int)Wl, H = (int)Hl;
    const int off = ccst ? 0 : -1;  // st:266-267 `+ map_d - 1`, ccst:452-453 `+ map_d`
    sx += dx + off; sy += dy + off; gx += dx + off; gy += dy + off;
    out->W = W; out->H = H; out->paste_x = (int)dx; out->paste_y = (int)dy;
    out->origin_x = nox; out->origin_y = noy;
    out->start_x = (int)sx; out->start_y = (int)sy; out->goal_x = (int)gx; out->goal_y = (int)gy;

    const size_t cells = (size_t)W * H;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->d_grid, &ctx->d_grid_cap, cells))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->d_grid2, &ctx->d_grid2_cap, cells))) return rc;
    if (mapped && cells <= ((size_t)1 << 20)) {
        int blocks = (int)((cells + 255) / 256);
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        k_assemble_small<<<blocks, 256, 0, st>>>((const uint8_t *)d_msg_in, in->layout == 0, aW, aH, wx0, wy0, wx1 - wx0, wy1 - wy0, ctx->d_grid, W, H,
                                                 (int)dx, (int)dy);
        FX_LAUNCH_CHECK(ctx);
        rc = FX_OK;
    } else {
        FX_CUDA(ctx, cudaMemsetAsync(ctx->d_grid, 0, cells, st));
        if (in->layout == 0)
            rc = fx_grid_decode(ctx, (const int8_t *)d_msg_in, width, height, wx0, wy0, wx1 - wx0, wy1 - wy0, ctx->d_grid, W, H, (int)dx, (int)dy, (void *)st);
        else
            rc = fx_grid_paste(ctx, (const uint8_t *)d_msg_in, aW, aH, wx0, wy0, wx1 - wx0, wy1 - wy0, ctx->d_grid, W, H, (int)dx, (int)dy,
                               (void *)st);
    }
    if (rc) return rc;
    // inflation: st = 9-point stencil {-ifa, 0, ifa}^2 (st:256-262), ccst = dense square (ccst:442-448)
    if ((rc = fx_inflate(ctx, ctx->d_grid, ctx->d_grid2, W, H, ifa, ccst ? 1 : ifa, (void *)st))) return rc;
    ctx->rp_W = W; ctx->rp_H = H;
    // goal relocation + end_occu (st:268-275, ccst:454-464)
    int32_t *h_blk = (int32_t *)(pin + off_rp);
    memset(h_blk, 0, 256);
    h_blk[4] = (int)gx; h_blk[5] = (int)gy; h_blk[8] = (int)sx; h_blk[9] = (int)sy;
    h_blk[10] = FX_COST_UNREACHABLE; h_blk[11] = FX_COST_UNREACHABLE; h_blk[12] = FX_COST_UNREACHABLE;
    if (!mapped) FX_CUDA(ctx, cudaMemcpyAsync(ctx->d_rp, h_blk, 256, cudaMemcpyHostToDevice, st));
    if (gx < 0 || gx >= W || gy < 0 || gy >= H)
        return fx_set_err(ctx, FX_ERR_ARG, "fx_replan_host: goal cell (%lld, %lld) outside the %d x %d grid", gx, gy, W, H);
    if ((rc = fx_relocate_goal(ctx, ctx->d_grid2, W, H, d_goal, ifa, ccst, (void *)st))) return rc;
    // st:280 / ccst:466 `if map_start[0] > map_c or map_start[1] > map_r: wp = global_goal` (no search)
    const bool skip = sx > mc || sy > mr;
    int32_t *d_raw = ctx->d_path, *d_post = mapped ? (int32_t *)(pin + off_p) : ctx->d_path + path_ints;
    double *d_world = mapped ? (double *)(pin + off_w) : (double *)ctx->d_pts;
    double world5[5] = {in->reso, nox, noy, 1.0, ccst ? 0.0 : 1.0};  // st:291 `+ [1,1]`, ccst:487 `+ [1,0]`
    double drop4[4] = {in->drop_px, in->drop_py, in->drop_pz, in->drop_radius};
    if (!skip) {
        if ((rc = fx_search_batch(ctx, ctx->d_grid2, W, H, d_start, d_goal, 1, in->hchoice, d_cost, d_costf, d_raw, d_len, max_path, (void *)st)))
            return rc;
        if ((rc = fx_path_post(ctx, ctx->d_grid2, W, H, d_raw, d_len, 1, max_path, in->shortcut, in->drop_radius > 0.0 ? drop4 : nullptr,
                               world5, d_post, d_plen, d_world, (void *)st)))
            return rc;
    }
    if (!mapped) FX_CUDA(ctx, cudaMemcpyAsync(h_blk, ctx->d_rp, 256, cudaMemcpyDeviceToHost, st));
    if (!skip && !mapped) {
        FX_CUDA(ctx, cudaMemcpyAsync(pin + off_p, d_post, path_ints * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        FX_CUDA(ctx, cudaMemcpyAsync(pin + off_w, ctx->d_pts, world_bytes, cudaMemcpyDeviceToHost, st));
    }
    FX_CUDA(ctx, cudaStreamSynchronize(st));
    out->goal_x = h_blk[4]; out->goal_y = h_blk[5]; out->goal_moved = h_blk[6]; out->end_occu = h_blk[7];
    out->skipped = skip ? 1 : 0;
    out->raw_len = skip ? 0 : h_blk[10];
    out->cost_i = skip ? FX_COST_UNREACHABLE : h_blk[11];
    out->path_len = skip ? 0 : h_blk[12];
    memcpy(&out->cost_f, h_blk + 32, sizeof(double));
    if (!skip && out->path_len > 0) {
        const int n = out->path_len < max_path ? out->path_len : max_path;
        if (h_path_xy) memcpy(h_path_xy, pin + off_p, (size_t)n * 2 * sizeof(int32_t));
        if (h_path_world) memcpy(h_path_world, pin + off_w, (size_t)n * 3 * sizeof(double));
    }
    return FX_OK;
}

// The inflated planning grid of the last fx_replan_host call (uint8 [W][H]); h_out may be NULL to query the shape only.
extern "C" int fx_replan_grid_host(fx_context *ctx, uint8_t *h_out, size_t cap, int *W, int *H)
{
    if (!ctx) return FX_ERR_ARG;
    if (ctx->rp_W <= 0 || !ctx->d_grid2) return fx_set_err(ctx, FX_ERR_ARG, "fx_replan_grid_host: no replan has run on this context");
    if (W) *W = ctx->rp_W;
    if (H) *H = ctx->rp_H;
    if (!h_out) return FX_OK;
    const size_t cells = (size_t)ctx->rp_W * ctx->rp_H;
    if (cap < cells) return fx_set_err(ctx, FX_ERR_ARG, "fx_replan_grid_host: buffer too small (%zu < %zu)", cap, cells);
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    FX_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->d_grid2, cells, cudaMemcpyDeviceToHost, ctx->own_stream));
    FX_CUDA(ctx, cudaStreamSynchronize(ctx->own_stream));
    return FX_OK;
}
