// common.cuh -- shared declarations for libfuxi_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fuxi_b200.h"

#define FX_INF 0xFFFFFFFFu
#ifndef FX_SEARCH_THREADS
#define FX_SEARCH_THREADS 128
#endif
#ifndef FX_SEARCH_MINB
#define FX_SEARCH_MINB 8 /* resident search CTAs per SM the register budget is compiled for (64 registers) */
#endif
#ifndef FX_SEARCH_WIDE
#define FX_SEARCH_WIDE 512 /* threads per CTA of the latency form (batches of at most sm_count queries) */
#endif
#define FX_SMALL_CELLS 20000 /* maps up to this many cells are searched entirely in one SM's shared memory (small.cu) */
#define FX_DIRTY_SHIFT 5 /* one dirty flag per 32 field cells (one 128 B line) */

struct fx_context {
    int device;
    int sm_count;
    size_t l2_bytes;
    char err[512];
    int64_t launches;

    // tuning
    int cfg_slots;
    int cfg_band0;
    int cfg_cluster;     // latency form: batches of at most sm_count / 8 queries run one query per thread-block cluster (FUXI_B200_CLUSTER=0: off)
    int cfg_wide_below;  // batches of at most this many queries use the wide (latency) CTA form; -1 = sm_count

    // search scratch, one set per kernel form (each sized for its sW x sH, reallocated when the shape changes):
    //   scr[0] throughput form: one cost field per slot, sm_count * FX_SEARCH_MINB slots
    //   scr[1] latency form (batches of at most sm_count queries): two fields per slot (bidirectional pass), sm_count slots
    struct SearchScratch {
        int sW, sH, slots, qcap, path_cap, nfields;
        size_t cells, dirty_n;
        uint32_t *fields;   // [slots][nfields][cells] packed cost | arrival direction, FX_INF = unreached
        uint8_t *dirty;     // [slots][dirty_n] one flag per 32 words of the slot's fields
        uint32_t *queues;   // [slots][4][qcap] rotating Dial buckets of (packed xy, packed word)
        int32_t *tmp_path;  // [slots][2][path_cap][2] turning points of the two walks of the path extraction
    } scr[2];
    uint8_t *moves;     // [cells] legal-move mask per cell
    size_t moves_cap;
    unsigned long long *counters;  // [8]: 0 work counter, 1 settled, 2 levels, 3 passes, 4 band-only, 5 flags
    // band pass scratch (band.cu)
    uint32_t *q_order;   // [Q] query indices, largest estimated work first
    uint32_t *q_ubound;  // [Q] upper bound on the cost from the band pass, FX_INF = none
    size_t q_cap;
    uint32_t *bfields;   // [bslots][bcap] band-coordinate cost fields
    size_t bcap; int bslots;
    // full-field (cooperative) scratch
    uint32_t *fq;       // [4][fqcap]
    size_t fqcap;
    unsigned int *fstate;  // small device state block for the cooperative kernel
    uint2 *seeds;       // (cost, packed xy) seed list for fx_field_relax
    uint2 *seeds_sorted;
    unsigned int *seed_hist;
    size_t seed_cap, seed_hist_cap;
    // projection scratch: bit-packed grid for the large-grid scatter
    unsigned *proj_bits;
    int inflate_attr_set;
    size_t proj_bits_cap;
    // EDT scratch
    uint16_t *edt_g;
    uint16_t *edt_s, *edt_t;
    size_t edt_cap, edt_st_cap;
    int *edt_flag;
    // host-buffer entry points: device buffers + pinned staging
    uint8_t *d_grid, *d_grid2;  size_t d_grid_cap, d_grid2_cap;
    int32_t *d_q;               size_t d_q_cap;      // starts, goals
    int32_t *d_out_i;           size_t d_out_cap;    // cost_i + path_len
    double *d_out_f;
    int32_t *d_path;            size_t d_path_cap;
    int32_t *d_cpath;           size_t d_cpath_cap;  // compact paths (paths.cu)
    int64_t *d_coff;            size_t d_coff_cap;   // offsets[Q+1]
    int64_t last_d2h_bytes;                          // device->host bytes of the last fx_plan_host* call
    float *d_pts;               size_t d_pts_cap;
    void *h_pin;                size_t h_pin_cap;
    // fused replan (assemble.cu): raw OccupancyGrid message, small device block {bbox4, goal4, start2, path_len, cost_i, ...}
    int8_t *d_msg;              size_t d_msg_cap;
    int32_t *d_rp;              size_t d_rp_cap;
    int rp_W, rp_H;             // shape of the last planning grid (kept in d_grid2)
    // cloud conditioning (cloud.cu): voxel bitmap + ranks, per-voxel accumulators / centroids, keep bitmap, sort records
    unsigned *cl_bits, *cl_gpref, *cl_chunk, *cl_vidx, *cl_keep, *cl_gpref2, *cl_chunk2;
    unsigned long long *cl_acc;
    float4 *cl_vox;
    void *cl_state, *cl_out, *df_rec;
    void *tf_q; size_t tf_q_bytes;  // fx_transform_filter: UAV-relative float64 points
    size_t cl_bits_bytes, cl_gpref_bytes, cl_chunk_bytes, cl_vidx_bytes, cl_keep_bytes, cl_gpref2_bytes, cl_chunk2_bytes,
        cl_acc_bytes, cl_vox_bytes, cl_state_bytes, cl_out_bytes, df_rec_bytes;
    unsigned long long cl_cap_bits;  // voxel index space the bitmap is reserved for (fx_cloud_reserve)
    cudaStream_t own_stream;
    cudaEvent_t ev_search[2];  // around the last k_search_batch launch (fx_search_kernel_ms)
    cudaEvent_t ev_band[2];    // around the last k_band_bound launch (fx_search_timings)
    int ev_search_valid;
    int small_attr_set, cfg_small_off, lat_attr_set;
    // CUDA-graph cache of the multi-launch device entry points (fx_graph_run)
    struct GraphSlot {
        unsigned char key[192];
        size_t keylen;
        int sightings;          // consecutive calls with this key
        cudaGraphExec_t exec;   // instantiated on the second sighting
    } graphs[4];
    double last_stage_us[6];            // fx_plan_host_stages (valid when last_stage_valid)
    int last_stage_valid;
    const uint8_t *moves_prebuilt_for;  // fx_plan_host -> fx_search_batch: the legal-move mask of this grid is already enqueued
    int calib_W, calib_H, calib_metric;  // what the first-bound table of the latency forms was learnt on (search.cu)
    cudaStream_t cap_stream;    // capture happens here (the caller's stream may be the legacy default stream)
    int cfg_graphs;             // FUXI_B200_GRAPHS=0: plain launches
};
enum { FX_GRAPH_CLOUD = 0, FX_GRAPH_EDT = 1, FX_GRAPH_DFILTER = 2, FX_GRAPH_TFILTER = 3 };

int fx_set_err(fx_context *ctx, int code, const char *fmt, ...);
#define FX_CUDA(ctx, call)                                                                        \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return fx_set_err(ctx, e__ == cudaErrorMemoryAllocation ? FX_ERR_NOMEM : FX_ERR_CUDA, \
                              "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define FX_LAUNCH_CHECK(ctx)                 \
    do {                                     \
        (ctx)->launches++;                   \
        FX_CUDA(ctx, cudaGetLastError());    \
    } while (0)

// direction d in 0..7: 0:(-1,0) 1:(+1,0) 2:(0,-1) 3:(0,+1) 4:(-1,-1) 5:(-1,+1) 6:(+1,-1) 7:(+1,+1)
// (d < 4 straight, d >= 4 diagonal; the opposite of d is d^1 for straight, 11-d for diagonal)
// packed 2-bit tables of (delta + 1), d0 in the low bits
__host__ __device__ constexpr int fx_dx(int d) { return (int)((0xA058u >> (2 * d)) & 3u) - 1; }
__host__ __device__ constexpr int fx_dy(int d) { return (int)((0x8885u >> (2 * d)) & 3u) - 1; }
static_assert(fx_dx(0) == -1 && fx_dy(0) == 0 && fx_dx(1) == 1 && fx_dy(1) == 0, "dir table");
static_assert(fx_dx(2) == 0 && fx_dy(2) == -1 && fx_dx(3) == 0 && fx_dy(3) == 1, "dir table");
static_assert(fx_dx(4) == -1 && fx_dy(4) == -1 && fx_dx(5) == -1 && fx_dy(5) == 1, "dir table");
static_assert(fx_dx(6) == 1 && fx_dy(6) == -1 && fx_dx(7) == 1 && fx_dy(7) == 1, "dir table");

// edge weights: metric 1 = the reference's hchoice 1 (10 / 14, scripts/jps1.py:3-12,232-246); metric 2 = Euclidean in 2^-16 fixed point
template <int METRIC> struct Wt;
template <> struct Wt<1> { static constexpr uint32_t WS = 10, WD = 14; };
template <> struct Wt<2> { static constexpr uint32_t WS = FX_EUCLID_WS, WD = FX_EUCLID_WD; };

__device__ __forceinline__ uint32_t octile(int ax, int ay, uint32_t ws, uint32_t wdiff)
{
    // ax, ay >= 0.  ws*max + (wd-ws)*min; fits 32 bits for W,H <= 32767 (checked on the host side)
    int mx = max(ax, ay), mn = min(ax, ay);
    return ws * (uint32_t)mx + wdiff * (uint32_t)mn;
}

// ---- canonical successor generation (batched search) -----------------------------------------------------------
// Packed cost-field entry: (cost << 4) | arrival direction (0..7, FX_CODE_START for the source); FX_INF = unreached.
// Costs therefore live in 28 bits: 32767 * 3363 < 2^27 covers any monotone path on the largest grid, longer (maze)
// paths up to 2^28 / WS = 112 k straight steps; beyond that the query reports FX_COST_OVERFLOW.
#define FX_CODE_START 8u
#define FX_COST_MAX28 0x0FFFFFFEu
__device__ __forceinline__ uint32_t fx_pack(uint32_t g, unsigned code) { return (g << 4) | code; }
__host__ __device__ constexpr int fx_dir_of(int dx, int dy)
{
    return dy == 0 ? (dx < 0 ? 0 : 1) : dx == 0 ? (dy < 0 ? 2 : 3) : 4 + (dx > 0 ? 2 : 0) + (dy > 0 ? 1 : 0);
}
static_assert(fx_dir_of(-1, 0) == 0 && fx_dir_of(1, 0) == 1 && fx_dir_of(0, -1) == 2 && fx_dir_of(0, 1) == 3, "dir_of");
static_assert(fx_dir_of(-1, -1) == 4 && fx_dir_of(-1, 1) == 5 && fx_dir_of(1, -1) == 6 && fx_dir_of(1, 1) == 7, "dir_of");
// Successors of a cell whose legal-move mask is m and which was reached by a move in direction `code`: the natural
// and forced neighbours of scripts/jps1.py:49-93 (nodeNeighbours) in single-step form.
//   source (code 8): every legal move (:51-56)
//   straight d:  d itself; the diagonal d+s for each side s whose cell is blocked (:75-92)
//   diagonal (dx,dy): (dx,0), (0,dy), (dx,dy); (-dx,dy) if (-dx,0) is blocked; (dx,-dy) if (0,-dy) is blocked (:59-73)
// each only if the move is legal (bit set in m).  "Blocked side cell" == the straight move onto it is illegal
// (jps1.blocked is True outside the array, so borders force neighbours exactly like the reference).
__host__ __device__ inline unsigned fx_canon_succ(unsigned code, unsigned m)
{
    if (code >= 8u) return m;
    const int dx = fx_dx((int)code), dy = fx_dy((int)code);
    auto bit = [](int a, int b) { return 1u << fx_dir_of(a, b); };
    unsigned s;
    if (code < 4u) {
        s = m & (1u << code);
        if (dx != 0) {
            if (!(m & bit(0, 1))) s |= m & bit(dx, 1);
            if (!(m & bit(0, -1))) s |= m & bit(dx, -1);
        } else {
            if (!(m & bit(1, 0))) s |= m & bit(1, dy);
            if (!(m & bit(-1, 0))) s |= m & bit(-1, dy);
        }
    } else {
        s = m & (bit(dx, 0) | bit(0, dy) | (1u << code));
        if (!(m & bit(-dx, 0))) s |= m & bit(-dx, dy);
        if (!(m & bit(0, -dy))) s |= m & bit(dx, -dy);
    }
    return s;
}

// Cell index inside the search scratch (cost fields, move masks).  FX_TILED: 8x8-cell tiles, tile-row major
// (a tile is 64 cells = 256 B of costs = 2 lines; a tile row of 8 cells is one 32-byte sector), so that the eight
// neighbours of a cell and the cells of a wavefront arc share sectors and lines in BOTH grid directions.  The
// row-major form is used by the whole-grid field kernel, whose output is the caller's [W][H] array.
#ifndef FX_TILED
#define FX_TILED 1
#endif
__host__ __device__ __forceinline__ int fx_tiles_y(int H) { return (H + 7) >> 3; }
__host__ __device__ __forceinline__ size_t fx_scratch_cells(int W, int H)
{
#if FX_TILED
    return (size_t)((W + 7) >> 3) * (size_t)fx_tiles_y(H) * 64;
#else
    return ((size_t)W * H + 63) / 64 * 64;
#endif
}
__device__ __forceinline__ int fx_cidx(int x, int y, int H, int TY)
{
#if FX_TILED
    (void)H;
    return ((((x >> 3) * TY) + (y >> 3)) << 6) | ((x & 7) << 3) | (y & 7);
#else
    (void)TY;
    return x * H + y;
#endif
}

int fx_grow_bytes(fx_context *ctx, void **p, size_t *cap, size_t want_bytes);
int fx_grow_pinned(fx_context *ctx, size_t want);
#include <functional>
// Launch-bound sizes (one camera frame through the 11 kernels of the cloud filter, a 1024^2 distance transform): the
// launch sequence of an entry point is a function of its arguments and of the scratch pointers it uses -- `key`.  The
// first call with a key runs `enqueue(stream)` as plain launches (and sizes the scratch), the second captures the same
// sequence into a CUDA graph on a private stream, later calls replay it: one launch instead of 8-16, no gaps between the
// kernels.  Device-side decisions (flags, counts) stay inside the kernels, so the replay is exact.  (api.cu)
int fx_graph_run(fx_context *ctx, int slot, const void *key, size_t keylen, cudaStream_t st, const std::function<int(cudaStream_t)> &enqueue);
// pageable host bytes -> start of the pinned staging buffer (non-temporal stores on the host worker pool) -> device (api.cu)
int fx_staged_copy_in(fx_context *ctx, uint8_t *d_dst, const uint8_t *h_src, size_t bytes, cudaStream_t st);
int fx_search_reserve(fx_context *ctx, int which, int W, int H, int max_path, cudaStream_t st);
void fx_search_release(fx_context *ctx, int which);
// band.cu: LPT query order + per-query upper bounds from the band pass (one warp per query)
int fx_band_bounds(fx_context *ctx, const uint8_t *grid, int W, int H, const int32_t *starts_xy, const int32_t *goals_xy, int Q,
                   int metric, cudaStream_t st);
// small.cu: whole query in shared memory; returns 1 when the map is too large for it
int fx_search_small(fx_context *ctx, const uint8_t *grid, int W, int H, const int32_t *starts_xy, const int32_t *goals_xy, int Q,
                    int metric, int32_t *cost_i, double *cost_f, int32_t *path_xy, int32_t *path_len, int max_path, cudaStream_t st);
int fx_build_moves(fx_context *ctx, const uint8_t *grid, int W, int H, bool tiled, cudaStream_t st);
int fx_build_moves_rows(fx_context *ctx, const uint8_t *grid, int W, int H, bool tiled, int x0, int x1, cudaStream_t st);
