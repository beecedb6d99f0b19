/*
 * fuxi_b200.h -- C ABI of libfuxi_b200.so: the B200 (sm_100a) implementation of FUXI's
 * global-planning hot path.  This is the drop-in boundary: plain pointers and sizes, no C++ or
 * torch types.  Every entry point cites the reference code it replaces (paths relative to the
 * chenhanpolyu/fuxi-planner repo).  The ctypes binding a maintainer adds is in INTEGRATION.md.
 *
 * Conventions
 *   - grids are uint8 [W][H], index x*H + y (y fastest) == the reference's matrix[x][y]
 *     (scripts/global_planner_st.py:16-18).  For the search, value 1 is an obstacle and every other
 *     value is free, exactly like `matrix[x][y] == 1` in scripts/jps1.py:20-29.
 *   - every function returns 0 on success or a negative FX_ERR_* code; fx_last_error(ctx) gives text.
 *     No C++ exception crosses the boundary.
 *   - pointers are DEVICE pointers unless the parameter name starts with h_ (host).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device-pointer
 *     entry points only enqueue work; they never synchronise the host.
 *   - a context is bound to one device and is not re-entrant: one caller thread at a time
 *     (the reference planners are single-threaded 80 Hz loops, global_planner_st.py:157).
 */
#ifndef FUXI_B200_H
#define FUXI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FX_OK                  0
#define FX_ERR_ARG            -1   /* bad argument */
#define FX_ERR_CUDA           -2   /* CUDA runtime error (text in fx_last_error) */
#define FX_ERR_NOMEM          -3   /* device or host allocation failed */
#define FX_ERR_UNSUPPORTED    -4   /* e.g. W or H > 65535 for the search */
#define FX_ERR_NCCL           -5

/* per-query values written to cost_i / path_len */
#define FX_COST_UNREACHABLE   -1   /* reference returns (0, t): scripts/jps1.py:230 */
#define FX_COST_START_OOB     -2   /* reference raises IndexError (start outside the array) */
#define FX_COST_OVERFLOW      -3   /* internal queue or 28-bit cost range exceeded; query not answered */

/* integer Euclidean metric (metric 2): straight : diagonal = 2378 : 3363, a convergent of sqrt(2)
 * (3363/2378 = sqrt(2) * (1 + 4.4e-8)); small enough that cost and arrival direction share one 32-bit word */
#define FX_EUCLID_WS 2378
#define FX_EUCLID_WD 3363

typedef struct fx_context fx_context;

/* ---- lifetime ------------------------------------------------------------------------------ */
int         fx_create(int device, fx_context **ctx);
int         fx_destroy(fx_context *ctx);
const char *fx_last_error(fx_context *ctx);     /* ctx may be NULL: last creation error */
int         fx_version(void);                   /* ABI version, currently 1 */
/* number of kernels this library launched through ctx since creation (bench.py: gpu_launches) */
int64_t     fx_launch_count(fx_context *ctx);
/* tuning: number of concurrent search slots (0 = auto: 8 per SM, bounded by half of the free memory) and the
 * half-width in cells of the band-limited passes of the search kernel (0 = default 17) */
int         fx_set_search_tuning(fx_context *ctx, int slots, int band0);
/* which form of the batched kernel fx_search_batch uses on maps too large for the shared-memory kernel: 0 = by batch size
 * (default: at most one query per SM -> latency form, else throughput form), 1 = always the throughput form (band kernel +
 * one forward search per query: every path is forward-canonical, which fx_paths_jump_points needs to reproduce the
 * reference's jump-point list), 2 = always the latency form (bidirectional: the goal side's half of a path is canonical
 * for the REVERSE direction).  Costs are identical in all forms. */
int         fx_set_search_form(fx_context *ctx, int form);

/* ---- (1) point cloud -> 2D occupancy grid ---------------------------------------------------
 * Replaces: the camera->earth transform + height filter of scripts/plc_point2_st.py:244-256
 * (= plc_point2_ccst.py:311-326) fused with the 3D->2D projection the reference delegates to
 * octomap_server / ccmapping (launch/map_st.launch:3, launch/map_ccst.launch:2; consumed at
 * scripts/global_planner_st.py:54-56).  Semantics (float32, one rounding per op, no FMA):
 *     e_k = ((a_k0*x + a_k1*y) + a_k2*z) + a_k3          h_affine3x4 = 12 floats, row major (HOST)
 *     fx = floorf((e_x - ox) / reso), fy likewise
 *     grid[fx*H + fy] = 1  iff  zmin < e_z <= zmax, 0 <= fx < W, 0 <= fy < H
 * pts: n points, stride_floats = 3 (packed xyz, the PointCloud2 layout of plc_point2_st.py:112-138)
 * or 4 (float4, w ignored; pointer must be 16-byte aligned).  clear_first != 0 zeroes grid first.
 */
int fx_project(fx_context *ctx, const float *pts, int64_t n, int stride_floats, const float *h_affine3x4,
               float zmin, float zmax, float ox, float oy, float reso, int W, int H,
               uint8_t *grid, int clear_first, void *stream);

/* ---- (2) obstacle inflation -----------------------------------------------------------------
 * Replaces: scripts/global_planner_st.py:256-262 (step == radius: 9-point stencil {-r,0,+r}^2) and
 * scripts/global_planner_ccst.py:442-448 (step == 1: dense (2r+1)^2 square).  Sources are cells > 0;
 * out = 1 where any source lies in the stencil, else 0.  Stencil targets outside the grid are dropped
 * (the reference pads the grid so that none exist, global_planner_st.py:230-250).  in != out.
 */
int fx_inflate(fx_context *ctx, const uint8_t *in, uint8_t *out, int W, int H, int radius, int step,
               void *stream);

/* Exact squared Euclidean distance (in cells^2) to the nearest cell > 0; INT32_MAX if the grid has
 * none.  Not in the reference (north-star addition); oracle = scipy.ndimage.distance_transform_edt.
 * W, H <= 65534 and (W-1)^2 + (H-1)^2 < 2^31 - 1 (the output is int32), else FX_ERR_UNSUPPORTED. */
int fx_edt(fx_context *ctx, const uint8_t *occ, int32_t *dist2, int W, int H, void *stream);
/* The two separable passes of fx_edt on their own (row-tiled multi-GPU mode, tiled.edt_tiled): fx_edt_rows writes
 * g[x][y] = distance along y to the nearest cell > 0 of row x (uint16, 0xFFFF = the row has none) -- local to an x-slab;
 * fx_edt_cols takes g for WHOLE columns (uint16 [W][Hb], any column block) and writes dist2 = min over x' of
 * (x-x')^2 + g(x',y)^2 (int32 [W][Hb], INT32_MAX when the grid has no occupied cell). */
int fx_edt_rows(fx_context *ctx, const uint8_t *occ, uint16_t *g, int W, int H, void *stream);
int fx_edt_cols(fx_context *ctx, const uint16_t *g, int32_t *dist2, int W, int H, void *stream);

/* ---- (3) shortest paths on the 8-connected grid ------------------------------------------------
 * Replaces: scripts/jps1.py:183-230 `method(matrix, start, goal, hchoice)` and everything it calls
 * (:3-179, :232-246), for Q independent (start, goal) queries per launch.  Edges are exactly
 * `not blocked(c, d)` of scripts/jps1.py:14-31 (target != 1; diagonal also needs not both orthogonal
 * cells == 1; the source cell is never tested; outside the array is blocked).
 *   metric 1: weights 10 / 14 (hchoice 1) -> cost_i[q] is the integer the reference prints (jps1.py:207)
 *   metric 2: Euclidean (hchoice 2)       -> cost_i[q] in units of 1/FX_EUCLID_WS cell (FX_EUCLID_WS/WD),
 *                                            cost_f[q] = straight + diagonal*sqrt(2) of the path found
 * starts_xy / goals_xy: int32 [Q][2].  cost_i: int32 [Q] (or FX_COST_*).  cost_f: double [Q] or NULL.
 * path_xy: int32 [Q][max_path][2] turning points, start first, goal last (consecutive points are
 * joined by a straight 8-direction run -- same contract as the reference's jump-point list), or NULL;
 * path_len: int32 [Q] number of turning points (FX_COST_* when there is no path), or NULL.  path_len[q] > max_path
 * means the path did not fit: the contents of that query's path buffer are then unspecified (the forms below
 * differ in which points they keep) -- call again with max_path >= path_len[q].
 * Three forms behind this one entry point, same results: maps up to 20 000 cells (every map the reference ships)
 * are searched by one CTA per query entirely in shared memory, one launch per batch; larger maps by the batched
 * wavefront kernel, with 512-thread CTAs when the batch is smaller than the machine (latency) and 128-thread CTAs,
 * eight per SM, otherwise (throughput).  W, H <= 32767.
 * The latency forms size their first pass from a table the context keeps (how far above the octile bound the optimum
 * lay on earlier queries with the same metric, by direction mix): it changes how many passes a query takes, never
 * its answer.
 */
int fx_search_batch(fx_context *ctx, const uint8_t *grid, int W, int H,
                    const int32_t *starts_xy, const int32_t *goals_xy, int Q, int metric,
                    int32_t *cost_i, double *cost_f, int32_t *path_xy, int32_t *path_len, int max_path,
                    void *stream);

/* Compact (CSR) form of the path output: offsets[q] = number of points of the paths before q (int64 [Q+1], offsets[Q] =
 * total), out_xy = the points of all paths back to back (int32 [cap][2]; points beyond cap are dropped -- compare
 * offsets[Q] with cap).  A query without a path (path_len <= 0) or whose path did not fit max_path contributes no point.
 * path_xy / path_len: exactly what fx_search_batch wrote.  The reference returns one Python list per call
 * (scripts/jps1.py:199-208); this is the batched equivalent without the padding. */
int fx_paths_compact(fx_context *ctx, const int32_t *path_xy, const int32_t *path_len, int Q, int max_path,
                     int64_t *offsets, int32_t *out_xy, int64_t cap, void *stream);

/* Jump-point form of the paths (opt-in): the reference returns the jump points of its path (scripts/jps1.py:199-208), the
 * search the turning points.  For each path this keeps, besides the turning points, every cell of its straight runs at
 * which jps1.jump (:95-164) would have stopped -- the goal, a cell with a forced neighbour for the travel direction, on a
 * diagonal run a cell whose straight sub-jump finds a jump point -- i.e. the list the reference would return for the same
 * cell path.  path_xy / path_len as fx_search_batch wrote them; out_xy int32 [Q][max_out][2], out_len int32 [Q] (may exceed
 * max_out: only max_out were stored; FX_COST_* is passed through).  The result is a list jps1.method could return (every
 * consecutive pair is what jps1.jump yields) when the path is forward-canonical: the shared-memory kernel and the
 * throughput form produce such paths, the bidirectional latency form does not (fx_set_search_form(ctx, 1) selects the
 * throughput form for any batch size). */
int fx_paths_jump_points(fx_context *ctx, const uint8_t *grid, int W, int H, const int32_t *path_xy, const int32_t *path_len,
                         int Q, int max_path, int32_t *out_xy, int32_t *out_len, int max_out, void *stream);
/* host-buffer form for one path (the drop-in jps1.method with POINTS = "jump"); h_grid == NULL reuses the grid the last
 * fx_plan_host* call on ctx uploaded */
int fx_jump_points_host(fx_context *ctx, const uint8_t *h_grid, int W, int H, const int32_t *h_path_xy, int n,
                        int32_t *h_out_xy, int max_out, int32_t *h_out_len);

/* Successor rule of the batched search, exposed for parity tests (pure host function, no GPU needed).
 * Replaces: scripts/jps1.py:49-93 `nodeNeighbours` (natural + forced neighbours of a cell given the direction it
 * was reached by), in single-step form.  code: 0..7 = arrival direction in the order (-1,0) (+1,0) (0,-1) (0,+1)
 * (-1,-1) (-1,+1) (+1,-1) (+1,+1), 8 = the start cell; moves: bit d set iff `not blocked(c, direction d)`
 * (scripts/jps1.py:14-31).  Returns the bit mask of directions the cell relaxes, or FX_ERR_ARG. */
int fx_canon_successors(int code, int moves);

/* Cost-from-source field: field[x*H+y] = cost of the cheapest legal path source -> (x,y), -1 if none
 * (int32, same units as cost_i).  The source may sit on an obstacle (jps1.py never tests it). */
int fx_field(fx_context *ctx, const uint8_t *grid, int W, int H, int sx, int sy, int metric,
             int32_t *field, void *stream);

/* Row-tiled variant for grids split into x-slabs across ranks: relaxes `field` (int32, -1 = unknown,
 * already holding seeds: the source and/or halo rows received from neighbours) to its local fixpoint.
 * grid/field cover slab rows [0, Wloc) where rows 0 and Wloc-1 may be ghost rows.  *d_changed (device
 * int) is set to 1 if any cell improved.  See fuxi_planner_b200/tiled.py for the exchange loop. */
int fx_field_relax(fx_context *ctx, const uint8_t *grid, int Wloc, int H, int metric,
                   int32_t *field, int32_t *d_changed, void *stream);

/* Halo merge of the row-tiled mode: dst[i] = min(dst[i], src[i]) over n int32 cells, -1 (unknown) counting as
 * +infinity; *d_changed (device int, may be NULL) is set to 1 if any dst cell improved.  dst is a boundary row of
 * this rank's slab, src the same row as received from the neighbouring rank (ncclSend/ncclRecv). */
int fx_halo_merge(fx_context *ctx, int32_t *dst, const int32_t *src, int64_t n, int32_t *d_changed, void *stream);

/* Synchronises and reports whether the last fx_field / fx_field_relax on ctx completed (FX_OK) or overflowed
 * an internal queue / the 31-bit cost range (FX_ERR_UNSUPPORTED).  Optional outputs: wavefront levels run and
 * cells settled. */
int fx_field_status(fx_context *ctx, int64_t *h_levels, int64_t *h_settled);

/* Search statistics of the last fx_search_batch / fx_plan_host on this ctx (host sync):
 * h_stats[0] = cells settled (popped and expanded), [1] = wavefront levels, [2] = search passes,
 * [3] = queries answered in the band-limited first pass alone. */
int fx_search_stats(fx_context *ctx, int64_t *h_stats4);
/* duration in ms of the last k_search_batch launch alone: CUDA events recorded on the launching stream immediately
 * before and after it (waits for that launch to finish).  bench.py: roofline of the dominant kernel. */
int fx_search_kernel_ms(fx_context *ctx, float *h_ms);
/* the same for both kernels of the batched search: h_ms2 = {k_band_bound (upper bounds), k_search_batch} */
int fx_search_timings(fx_context *ctx, float *h_ms2);
/* Where a host-buffer planning call spends its time (break-down of the single-replan latency): with the
 * environment variable FUXI_B200_TRACE set to 1 (also prints a line per call on stderr) or 2 (silent), fx_plan_host*
 * records, in microseconds, h_us6[0] host: filling the pinned staging buffer and issuing the uploads, [1] host: enqueueing
 * search and copies, [2] host: waiting for the device, [3] device (CUDA events): grid upload, [4] device: fx_search_batch
 * (legal-move mask + kernels), [5] device: path compaction + D2H.  FX_ERR_ARG if the last call was not traced. */
int fx_plan_host_stages(fx_context *ctx, double *h_us6);

/* ---- host-buffer convenience = what the Python drop-in `jps1.method` calls ---------------------
 * Same as fx_search_batch but all buffers are HOST memory; copies in, runs, copies out and
 * synchronises.  Uses pinned staging owned by ctx. */
int fx_plan_host(fx_context *ctx, const uint8_t *h_grid, int W, int H,
                 const int32_t *h_starts_xy, const int32_t *h_goals_xy, int Q, int metric,
                 int32_t *h_cost_i, double *h_cost_f, int32_t *h_path_xy, int32_t *h_path_len, int max_path);

/* Same with the paths in compact form: h_offsets int64 [Q+1], h_xy int32 [cap][2] (the first min(total, cap) points),
 * *h_total = offsets[Q] (may be NULL).  Only `total` points cross the bus (fx_plan_host does the same internally and
 * scatters them into the padded rows).  max_path still bounds a single path (longer ones report path_len > max_path
 * and contribute no point). */
int fx_plan_host_csr(fx_context *ctx, const uint8_t *h_grid, int W, int H,
                     const int32_t *h_starts_xy, const int32_t *h_goals_xy, int Q, int metric,
                     int32_t *h_cost_i, double *h_cost_f, int32_t *h_path_len, int max_path,
                     int64_t *h_offsets, int32_t *h_xy, int64_t cap, int64_t *h_total);
/* device->host bytes the last fx_plan_host / fx_plan_host_f64 / fx_plan_host_csr on ctx copied (bench.py: e2e) */
int64_t fx_last_d2h_bytes(fx_context *ctx);

/* Same, taking the matrix as the reference's callers hold it: float64 [W][H] (np.zeros, scripts/global_planner_st.py:248),
 * obstacle iff matrix[x][y] == 1.0 (scripts/jps1.py:20-29).  The `== 1 -> uint8` conversion runs on a few host threads
 * straight into the pinned staging buffer, overlapped with the H2D copy. */
int fx_plan_host_f64(fx_context *ctx, const double *h_matrix, int W, int H,
                     const int32_t *h_starts_xy, const int32_t *h_goals_xy, int Q, int metric,
                     int32_t *h_cost_i, double *h_cost_f, int32_t *h_path_xy, int32_t *h_path_len, int max_path);

/* host-buffer map pipeline: project + inflate in one call (HOST in, HOST out), for the ROS-side glue */
int fx_map_host(fx_context *ctx, const float *h_pts, int64_t n, int stride_floats, const float *h_affine3x4,
                float zmin, float zmax, float ox, float oy, float reso, int W, int H,
                int radius, int step, uint8_t *h_grid_out);

/* ---- (4) planner-side grid assembly on the device ---------------------------------
 * fx_grid_decode: nav_msgs/OccupancyGrid.data (int8, index y*width + x) -> uint8 array [x][y] with the value
 * mapping of map_callback, scripts/global_planner_st.py:15-20 (= global_planner_ccst.py:17-24): 100 -> 1, -1 -> 0,
 * everything else unchanged; fused with the slice paste of st:249-250: the source window
 * [sx0, sx0+w) x [sy0, sy0+h) of the message lands at dst[px.., py..] of a dW x dH array (cells outside either
 * array are skipped; dst cells outside the window are left as they are).
 * fx_grid_encode: the inverse for publishing (publish_map, st:102-115): 1 -> 100, data = mapu.T.flatten().
 * fx_grid_paste: dst[px+i][py+j] = src[sx0+i][sy0+j] between [x][y] arrays, overwriting -- the slice assignments of
 * the pre-map merge (st:210-224) and of the pad / shift step.
 * fx_grid_bbox: {min x, max x, min y, max y} of the non-zero cells (is_msg: of a raw message, value not in {0,-1});
 * {INT32_MAX, -1, INT32_MAX, -1} if there are none.  Replaces X.nonzero()/np.unique/min/max in
 * remove_zero_rowscols, scripts/global_planner_ccst.py:42-49.  d_bbox4: device int32[4].
 * fx_relocate_goal: d_goal4 = device int32[4] {gx, gy, -, -} -> {gx', gy', moved, end_occu}; replaces
 * st:268-275 / ccst:454-464: a goal on a cell == 1 moves to the nearest cell == 0 of its x-row (lower y on ties),
 * else of its y-column; moved = -1 if neither has one (the reference raises), -2 if the goal is outside the grid.
 * end_occu: st (ccst == 0) = moved; ccst = any cell == 1 in [gx-ifa, gx+ifa) x [gy-ifa, gy+ifa). */
int fx_grid_decode(fx_context *ctx, const int8_t *msg, int width, int height, int sx0, int sy0, int w, int h,
                   uint8_t *dst, int dW, int dH, int px, int py, void *stream);
int fx_grid_encode(fx_context *ctx, const uint8_t *grid, int W, int H, int8_t *msg, void *stream);
int fx_grid_paste(fx_context *ctx, const uint8_t *src, int sW, int sH, int sx0, int sy0, int w, int h,
                  uint8_t *dst, int dW, int dH, int px, int py, void *stream);
int fx_grid_bbox(fx_context *ctx, const void *a, int W, int H, int is_msg, int32_t *d_bbox4, void *stream);
int fx_relocate_goal(fx_context *ctx, const uint8_t *grid, int W, int H, int32_t *d_goal4, int ifa, int ccst, void *stream);
/* Map image <-> array.  img: uint8 'L' image, H rows x W columns, row major.
 * fx_grid_to_image: the save block scripts/global_planner_st.py:368-372 (= global_planner_ccst.py:628-632):
 *   pixel[H-1-y][x] = grid[x][y] == 0 ? 255 : 0   (mapsave.T[::-1]).
 * fx_image_to_grid: the pre-map loader st:176-182: grid[x][y] = img[H-1-y][x] > threshold ? 0 : 1  (img[::-1].T;
 *   threshold 200 there; threshold 0 inverts fx_grid_to_image = the fixture convention `a[::-1].T == 0`). */
int fx_grid_to_image(fx_context *ctx, const uint8_t *grid, int W, int H, uint8_t *img, void *stream);
int fx_image_to_grid(fx_context *ctx, const uint8_t *img, int W, int H, int threshold, uint8_t *grid, void *stream);

/* ---- (5) path post-processing -----------------------------------------------------
 * Per path q (path_xy int32 [Q][max_path][2], path_len int32 [Q]; same layout as fx_search_batch writes):
 *   h_drop4 = {px, py, pz, radius} (HOST, may be NULL; radius <= 0 = off): if the path has more than two points,
 *     drop every point ii >= 1 whose world position (z = 0) is closer than radius to (px, py, pz)
 *     -- scripts/global_planner_ccst.py:507-513 (radius 1.5 there);
 *   shortcut != 0: greedy line-of-sight shortcutting `ii = 1; while ii < len-1: if map_line_col(path[ii+1],
 *     path[ii-1], mapu[bounding box]) delete ii else ii += 1` -- ccst:258-283 + 515-521, same sampling
 *     (one sample per integer x strictly between the points, y = rint(slope*x), half-open bounding box), cell == 1;
 *   out_world (double [Q][max_path][3], may be NULL): (cell + off) * reso + origin, z = 0 -- st:291-296 /
 *     ccst:487-491; h_world5 = {reso, origin_x, origin_y, off_x, off_y} (HOST; off = (1,1) st, (1,0) ccst).
 * out_xy may alias path_xy.  out_len[q] = points kept (path_len[q] <= 0 is passed through).  Bit-exact. */
int fx_path_post(fx_context *ctx, const uint8_t *grid, int W, int H, const int32_t *path_xy, const int32_t *path_len,
                 int Q, int max_path, int shortcut, const double *h_drop4, const double *h_world5,
                 int32_t *out_xy, int32_t *out_len, double *out_world, void *stream);

/* ---- (6) one global replan in one call -----------------------------------------------------------
 * What the planner loops do per iteration between receiving the map and publishing the path:
 * scripts/global_planner_st.py:226-298 (variant 0) / scripts/global_planner_ccst.py:36-63 + 411-521 (variant 1):
 * [crop] -> index/pad/shift -> decode+paste -> inflate -> goal relocation -> search -> [near-vehicle drop,
 * shortcutting] -> world coordinates.  One H2D (the map), one D2H (the result), one synchronisation. */
typedef struct fx_replan_in {
    int32_t variant;    /* 0 = st (9-point stencil, index offset -1, path offset [1,1]); 1 = ccst (dense square, 0, [1,0]) */
    int32_t layout;     /* 0 = h_map is OccupancyGrid.data (int8, y*width + x); 1 = uint8 array [x][y], W = width, H = height */
    int32_t crop;       /* != 0: crop to the occupied bounding box first (remove_zero_rowscols, ccst:36-63) */
    int32_t ifa;        /* inflation radius in cells (st: 1, ccst: 2 in the reference) */
    int32_t hchoice;    /* 1: 10/14 integer metric, 2: Euclidean (the planners pass 2) */
    int32_t shortcut;   /* != 0: line-of-sight shortcutting of the path (ccst:515-521) */
    double origin_x, origin_y, reso;                 /* OccupancyGrid.info.origin / resolution */
    double start_x, start_y, goal_x, goal_y;         /* world coordinates of the vehicle and the goal */
    double drop_px, drop_py, drop_pz, drop_radius;   /* near-vehicle drop (ccst:507-513); radius <= 0 = off */
} fx_replan_in;
typedef struct fx_replan_out {
    int32_t W, H;                 /* padded planning grid */
    int32_t paste_x, paste_y;     /* the reference's map_d */
    int32_t start_x, start_y, goal_x, goal_y;   /* cells in the planning grid; goal after relocation */
    int32_t goal_moved, end_occu;
    int32_t skipped;              /* 1: start beyond the map (st:280, no search); 2: empty map (ccst main loop skips) */
    int32_t raw_len, path_len;    /* turning points found / points after post-processing (<= 0: FX_COST_*) */
    int32_t cost_i;
    double cost_f;                /* what jps1.method prints: gscore[goal] */
    double origin_x, origin_y;    /* world position of cell [0][0] of the planning grid (the shifted map_o) */
} fx_replan_out;
int fx_replan_host(fx_context *ctx, const void *h_map, int width, int height, const fx_replan_in *in, fx_replan_out *out,
                   int32_t *h_path_xy, double *h_path_world, int max_path);
/* the inflated planning grid of the last fx_replan_host (uint8 [W][H]); h_out may be NULL to query W, H only */
int fx_replan_grid_host(fx_context *ctx, uint8_t *h_out, size_t cap, int *W, int *H);

/* ---- (7) upstream cloud conditioning ------------------------------------------------
 * fx_cloud_filter: the PCL chain of src/chen_filter_rgb.cpp:52-71 in one call,
 *     PassThrough("z", pass_lo, pass_hi) -> VoxelGrid(leaf_x, leaf_y, leaf_z) -> RadiusOutlierRemoval(radius, min_neighbors)
 * (the reference: 0..4, 0.17/0.17/0.2, 0.35/13).  PCL is an un-vendored, unpinned dependency (CMakeLists.txt:10-18), so
 * parity is against its published algorithm as restated in oracle/fuxi_oracle.c (fxo_cloud_filter), bit-exact:
 *   PassThrough: keep a point iff x, y, z are finite and !(z < pass_lo || z > pass_hi);
 *   VoxelGrid: inv = 1.0f/leaf; min_b = floor(min_p*inv), div_b = max_b - min_b + 1 over the kept points;
 *     voxel = (int)(floorf(p*inv) - (float)min_b); idx = i + j*div_x + k*div_x*div_y; one output point per occupied
 *     voxel in ascending idx.  Centroid: PCL sums floats in the order its (unstable) std::sort leaves them, which is
 *     not reproducible; here coordinates accumulate as rint(p * 2^24) in int64 (order independent) and
 *     centroid = (float)((double)sum / n * 2^-24); r, g, b = floor(channel sum / n) (downsample_all_data);
 *   RadiusOutlierRemoval: k = #{q : ((cx-qx)^2 + (cy-qy)^2) + (cz-qz)^2 < (float)(radius*radius)} over the voxel
 *     centroids, float32, the point itself included; kept iff k > min_neighbors.
 * pts: n points of stride_floats floats each, x y z at 0 1 2; rgb_offset = float index of PCL's packed rgb word
 * (4 for PointXYZRGB's 32-byte point_step) or -1.  out: float [cap][4] = x, y, z, rgb word (16-byte aligned);
 * d_counts: device int64[4] = {points after PassThrough, voxels, points kept, status}; status 0 = ok, > 0 = voxel index
 * space needed when it exceeds the reserved one (nothing is written; fx_cloud_reserve and call again), -1 = coordinates
 * outside the int range.  PCL itself refuses a voxel index space above 2^31.  fx_cloud_filter_host: HOST buffers, grows
 * the reservation and retries by itself, h_counts like d_counts. */
typedef struct fx_cloud_params {
    int32_t stride_floats, rgb_offset;
    float pass_lo, pass_hi;
    float leaf_x, leaf_y, leaf_z;
    int32_t min_neighbors;
    double radius;
} fx_cloud_params;
int fx_cloud_reserve(fx_context *ctx, int64_t max_voxel_space);   /* default 2^28 */
/* fx_cloud_filter (11 launches), fx_edt (8) and fx_distance_filter are replayed as one CUDA graph from the third call with
 * identical arguments (same pointers, sizes, parameters) on; the graph is launched into `stream` like the kernels would be. */
int fx_cloud_filter(fx_context *ctx, const float *pts, int64_t n, const fx_cloud_params *h_params, float *out, int64_t cap,
                    int64_t *d_counts, void *stream);
int fx_cloud_filter_host(fx_context *ctx, const float *h_pts, int64_t n, const fx_cloud_params *h_params, float *h_out,
                         int64_t cap, int64_t *h_counts);
/* fx_distance_filter: convert_plc.distance_filter, scripts/plc_point2_st.py:139-148 (= plc_point2_ccst.py:178-193):
 * d = sqrt((x*x + y*y) + z*z) in float64, keep d < dis, order by (d, z, y, x) like np.lexsort (stable).
 * pts / out: double [n][3]; d_count: device int32 = rows written.  Bit-exact against the reference function. */
int fx_distance_filter(fx_context *ctx, const double *pts, int64_t n, double dis, double *out, int32_t *d_count, void *stream);
int fx_distance_filter_host(fx_context *ctx, const double *h_pts, int64_t n, double dis, double *h_out, int64_t *h_count);

/* fx_transform_filter: the cloud the node publishes, scripts/plc_point2_st.py:243-256 + 336-339 (= plc_point2_ccst.py:310-326,
 * 405-408), float64 like the reference:
 *   camera branch (h_R9 != NULL): b = (z_c + 0.12, -x_c, -y_c); e = R b + t with R = utils.body_to_earth_frame(r, p, y)
 *     (row major, utils.py:21-28) and t = local_pos1, each component ((R_k0 b0 + R_k1 b1) + R_k2 b2) + t_k with one rounding per
 *     operation; keep e_z > zmin (0.3 in the reference);
 *   octomap-centres branch (h_R9 == NULL, plc_point2_st.py:351-362): e = p;
 *   then q = e - c (c = local_pos); box > 0: keep |q_k| < box on every axis (:361, 4 in the reference); distance_filter(q, dis):
 *     keep ||q|| < dis, order by (||q||, z, y, x) like np.lexsort; out = q + c.
 * pts: n points of `stride` elements each (x y z first), float32 (is_f64 == 0: PointCloud2 data) or float64.  out: double [n][3];
 * d_count: device int32 = rows written.  The appended [n_dyn, 0, 0] row (:341) and the float32 packing (a18) stay on the host. */
int fx_transform_filter(fx_context *ctx, const void *pts, int64_t n, int stride, int is_f64, const double *h_R9,
                        const double *h_t3, const double *h_c3, double zmin, double box, double dis, double *out,
                        int32_t *d_count, void *stream);
int fx_transform_filter_host(fx_context *ctx, const void *h_pts, int64_t n, int stride, int is_f64, const double *h_R9,
                             const double *h_t3, const double *h_c3, double zmin, double box, double dis, double *h_out,
                             int64_t *h_count);

#ifdef __cplusplus
}
#endif
#endif /* FUXI_B200_H */
