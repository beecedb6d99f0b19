// project.cu -- point cloud -> height-sliced 2D occupancy grid (sm_100a).
//
// Fuses the camera->earth rigid transform and z filter of scripts/plc_point2_st.py:244-256 with the
// 3D->2D projection the reference delegates to octomap_server / ccmapping (launch/map_st.launch:3,
// launch/map_ccst.launch:2).  One pass over the cloud: coalesced 16-byte loads (float4 directly, or
// three float4 per four packed-xyz points), 3x4 affine in registers with explicit round-to-nearest
// mul/add/sub/div (no FMA contraction, so the numpy float32 oracle is reproduced bit for bit), then a
// store of 1 into the grid.  The store is idempotent, so the result does not depend on scatter order.
//
// Two scatter forms (same arithmetic, same result):
//   * byte form  -- grids that sit in L2 anyway (<= FX_PROJ_BITS_MIN_CELLS cells) or clear_first == 0:
//                   plain byte stores of 1 straight into the caller's grid.
//   * bit form   -- large grids: a random byte store dirties a 32-byte sector, so on a grid bigger than L2
//                   every point costs a sector fill + a sector write-back in HBM (measured r01: 3.97 GB of
//                   DRAM traffic for 1.34 GB of algorithmic bytes).  Instead the points are RED.OR-ed into a
//                   bit-packed copy of the grid (W*H/8 bytes: 32 MiB at 16384^2, L2-resident, the reductions
//                   resolve in the L2 slices) and one streaming pass expands bits to bytes: the cloud is read
//                   once and the grid written once.
// HBM-bound: N*16 (or N*12) bytes read + the W*H grid written.
#include <stdlib.h>

#include "common.cuh"

struct ProjParams {
    float a[12];
    float zmin, zmax, ox, oy, reso;
    int W, H;
};

struct EmitByte {
    uint8_t *__restrict__ grid;
    __device__ __forceinline__ void operator()(size_t cell) const { grid[cell] = 1; }
};
struct EmitBit {
    unsigned *__restrict__ bits;
    __device__ __forceinline__ void operator()(size_t cell) const { atomicOr(bits + (cell >> 5), 1u << (cell & 31)); }  // RED.OR
};

template <typename Emit>
__device__ __forceinline__ void project_one(const ProjParams &p, float x, float y, float z, const Emit &emit)
{
    // e_k = ((a_k0*x + a_k1*y) + a_k2*z) + a_k3
    float ez = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.a[8], x), __fmul_rn(p.a[9], y)), __fmul_rn(p.a[10], z)), p.a[11]);
    if (!(ez > p.zmin && ez <= p.zmax)) return;
    float ex = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.a[0], x), __fmul_rn(p.a[1], y)), __fmul_rn(p.a[2], z)), p.a[3]);
    float ey = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.a[4], x), __fmul_rn(p.a[5], y)), __fmul_rn(p.a[6], z)), p.a[7]);
    float fx = floorf(__fdiv_rn(__fsub_rn(ex, p.ox), p.reso));
    float fy = floorf(__fdiv_rn(__fsub_rn(ey, p.oy), p.reso));
    if (!(fx >= 0.f && fx < (float)p.W && fy >= 0.f && fy < (float)p.H)) return;  // also rejects NaN
    emit((size_t)(int)fx * p.H + (int)fy);
}

// stride 4: one float4 per point, four independent 16-byte loads in flight per thread
template <typename Emit>
__global__ void __launch_bounds__(256) k_project_f4(const float4 *__restrict__ pts, long long n, ProjParams p, Emit emit, const int *__restrict__ run_flag)
{
    if (run_flag && *run_flag == 0) return;  // fallback of the partition form: only when it gave up
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        float4 v0 = __ldcs(pts + i), v1 = __ldcs(pts + i + stride), v2 = __ldcs(pts + i + 2 * stride),
               v3 = __ldcs(pts + i + 3 * stride);
        project_one(p, v0.x, v0.y, v0.z, emit);
        project_one(p, v1.x, v1.y, v1.z, emit);
        project_one(p, v2.x, v2.y, v2.z, emit);
        project_one(p, v3.x, v3.y, v3.z, emit);
    }
    for (; i < n; i += stride) {
        float4 v = __ldcs(pts + i);
        project_one(p, v.x, v.y, v.z, emit);
    }
}

// stride 3 (PointCloud2 packed xyz, plc_point2_st.py:112-138): four points = three float4
template <typename Emit>
__global__ void __launch_bounds__(256) k_project_f3(const float *__restrict__ pts, long long n, ProjParams p, Emit emit, const int *__restrict__ run_flag)
{
    if (run_flag && *run_flag == 0) return;
    const long long ngroups = n / 4;
    const float4 *__restrict__ p4 = reinterpret_cast<const float4 *>(pts);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; g + stride < ngroups; g += 2 * stride) {
        const long long h = g + stride;
        float4 a = __ldcs(p4 + 3 * g), b = __ldcs(p4 + 3 * g + 1), c = __ldcs(p4 + 3 * g + 2);
        float4 d = __ldcs(p4 + 3 * h), e = __ldcs(p4 + 3 * h + 1), f = __ldcs(p4 + 3 * h + 2);
        project_one(p, a.x, a.y, a.z, emit);
        project_one(p, a.w, b.x, b.y, emit);
        project_one(p, b.z, b.w, c.x, emit);
        project_one(p, c.y, c.z, c.w, emit);
        project_one(p, d.x, d.y, d.z, emit);
        project_one(p, d.w, e.x, e.y, emit);
        project_one(p, e.z, e.w, f.x, emit);
        project_one(p, f.y, f.z, f.w, emit);
    }
    if (g < ngroups) {
        float4 a = __ldcs(p4 + 3 * g), b = __ldcs(p4 + 3 * g + 1), c = __ldcs(p4 + 3 * g + 2);
        project_one(p, a.x, a.y, a.z, emit);
        project_one(p, a.w, b.x, b.y, emit);
        project_one(p, b.z, b.w, c.x, emit);
        project_one(p, c.y, c.z, c.w, emit);
    }
    // tail (< 4 points)
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - ngroups * 4)) {
        long long i = ngroups * 4 + threadIdx.x;
        project_one(p, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], emit);
    }
}

// unaligned / generic stride fallback (scalar loads)
template <typename Emit>
__global__ void __launch_bounds__(256) k_project_generic(const float *__restrict__ pts, long long n, int sf, ProjParams p, Emit emit, const int *__restrict__ run_flag)
{
    if (run_flag && *run_flag == 0) return;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        project_one(p, pts[(size_t)i * sf], pts[(size_t)i * sf + 1], pts[(size_t)i * sf + 2], emit);
}

// bit form, second half: expand the bit-packed grid to the caller's byte grid (16 cells per thread-iteration:
// one 16-bit read, one 16-byte streaming store).  cells16 = ceil(W*H / 16); the tail is written bytewise.
__global__ void __launch_bounds__(256) k_bits_to_bytes(const unsigned short *__restrict__ bits, uint8_t *__restrict__ grid,
                                                       size_t cells, const int *__restrict__ run_flag)
{
    if (run_flag && *run_flag == 0) return;
    const size_t full = cells / 16;
    const size_t T = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < full; i += T) {
        const unsigned b = bits[i];
        uint4 v;
        v.x = ((b & 0xFu) * 0x00204081u) & 0x01010101u;
        v.y = (((b >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
        v.z = (((b >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
        v.w = (((b >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
        __stcs(reinterpret_cast<uint4 *>(grid) + i, v);
    }
    if (blockIdx.x == 0 && threadIdx.x < (unsigned)(cells - full * 16)) {
        const size_t c = full * 16 + threadIdx.x;
        grid[c] = (uint8_t)((bits[c >> 4] >> (c & 15)) & 1u);
    }
}


// ---- partition form (large grids, large clouds) ----------------------------------------------------------------------
// The RED form above costs one L2 transaction per point that passes the height test; on a uniformly random cloud no two
// points of a warp share a word, the L2 request rate (not HBM) is the ceiling (r01: 46 % of the HBM rate at 64 Mi points ->
// 16384^2).  Here the cell indices are partitioned by grid REGION (2^19 cells = 512 KiB of the byte grid, 64 KiB of bits)
// so that a second kernel can rebuild each region's occupancy in shared memory and write the grid once:
//   pass A  k_project_part: a CTA takes chunks of 8192 points, computes the cell indices, counting-sorts them by region
//           in shared memory (histogram -> scan -> placement), reserves room in every region's buffer with ONE global
//           atomic per region and chunk, and copies the sorted run out with coalesced stores (runs of one region abut the
//           runs other chunks appended, so partially written sectors complete in L2);
//   pass B  k_project_fill: one CTA per region ORs the region's cell indices into a 64 KiB shared-memory bitmap and expands
//           it into the caller's byte grid with 16-byte streaming stores.
// Traffic: cloud read once, 4 B per kept point written + read once, grid written once.  A region buffer that overflows
// (a cloud concentrated on a few regions -- exactly the case the RED form is good at, because same-word reductions merge)
// raises a device flag and the RED form redoes the grid; no host synchronisation either way.
#define PJ_REGION_BITS 19
#define PJ_REGION_CELLS (1u << PJ_REGION_BITS)
#define PJ_THREADS 512
#define PJ_PPT 16
#define PJ_CHUNK (PJ_THREADS * PJ_PPT)
#define PJ_MAX_REGIONS 2048
#define PJ_INVALID 0xFFFFFFFFu

struct PartArgs {
    ProjParams p;
    long long n;
    int R;             // regions
    unsigned cap;      // entries per region buffer
    unsigned *count;   // [R] entries appended per region
    unsigned *buf;     // [R][cap] cell indices
    int *flag;         // set to 1 when a region buffer overflowed
};

__device__ __forceinline__ unsigned project_cell(const ProjParams &p, float x, float y, float z)
{
    float ez = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.a[8], x), __fmul_rn(p.a[9], y)), __fmul_rn(p.a[10], z)), p.a[11]);
    if (!(ez > p.zmin && ez <= p.zmax)) return PJ_INVALID;
    float ex = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.a[0], x), __fmul_rn(p.a[1], y)), __fmul_rn(p.a[2], z)), p.a[3]);
    float ey = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.a[4], x), __fmul_rn(p.a[5], y)), __fmul_rn(p.a[6], z)), p.a[7]);
    float fx = floorf(__fdiv_rn(__fsub_rn(ex, p.ox), p.reso));
    float fy = floorf(__fdiv_rn(__fsub_rn(ey, p.oy), p.reso));
    if (!(fx >= 0.f && fx < (float)p.W && fy >= 0.f && fy < (float)p.H)) return PJ_INVALID;
    return (unsigned)((int)fx * p.H + (int)fy);  // < 2^30: this form is only used for grids up to 2^30 cells
}

// SF = floats per point: 4 (float4) or 3 (packed xyz: four points = three float4)
template <int SF>
__global__ void __launch_bounds__(PJ_THREADS, 3) k_project_part(const float *__restrict__ pts, PartArgs a)
{
    extern __shared__ unsigned pj_sm[];
    unsigned *stage = pj_sm;                 // [PJ_CHUNK] cell index per point of the chunk (PJ_INVALID = dropped)
    unsigned *sorted = stage + PJ_CHUNK;     // [PJ_CHUNK] the kept indices grouped by region
    unsigned *cur = sorted + PJ_CHUNK;       // [R] histogram, then placement cursor
    unsigned *offs = cur + a.R;              // [R + 1] exclusive scan of the histogram
    unsigned *base = offs + a.R + 1;         // [R] position of this chunk's run in the region buffer
    __shared__ unsigned s_warp[PJ_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = a.R;
    const int per = (R + PJ_THREADS - 1) / PJ_THREADS;  // regions per thread in the scan (<= 4)
    const long long nchunks = (a.n + PJ_CHUNK - 1) / PJ_CHUNK;
    for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        for (int r = tid; r < R; r += PJ_THREADS) cur[r] = 0;
        __syncthreads();
        const long long p0 = chunk * PJ_CHUNK;
        // phase 1: cell indices + histogram.  All loads of a thread are issued before the first index is computed.
        if (SF == 4) {
            const float4 *__restrict__ p4 = reinterpret_cast<const float4 *>(pts) + p0;
            // four independent 16-byte loads in flight per thread (x 512 threads x 3 CTAs per SM = 96 KB per SM)
#pragma unroll 1
            for (int j0 = 0; j0 < PJ_PPT; j0 += 4) {
                float4 v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const long long i = p0 + (j0 + j) * PJ_THREADS + tid;
                    v[j] = i < a.n ? __ldcs(p4 + (j0 + j) * PJ_THREADS + tid) : make_float4(0.f, 0.f, __int_as_float(0x7FC00000), 0.f);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const unsigned c = project_cell(a.p, v[j].x, v[j].y, v[j].z);  // a NaN z fails the height test
                    stage[(j0 + j) * PJ_THREADS + tid] = c;
                    if (c != PJ_INVALID) atomicAdd(&cur[c >> PJ_REGION_BITS], 1u);
                }
            }
        } else {
            // groups of four points = three float4; the chunk start is a multiple of four points, so groups stay aligned
            const float4 *__restrict__ p4 = reinterpret_cast<const float4 *>(pts + 3 * p0);
            const long long ngroups = (a.n - p0) / 4;  // whole groups available from p0 on
#pragma unroll
            for (int j = 0; j < PJ_PPT / 4; j++) {
                const int gidx = j * PJ_THREADS + tid;
                unsigned c[4] = {PJ_INVALID, PJ_INVALID, PJ_INVALID, PJ_INVALID};
                if (gidx < ngroups) {
                    const float4 u = __ldcs(p4 + 3 * gidx), v = __ldcs(p4 + 3 * gidx + 1), w = __ldcs(p4 + 3 * gidx + 2);
                    c[0] = project_cell(a.p, u.x, u.y, u.z); c[1] = project_cell(a.p, u.w, v.x, v.y);
                    c[2] = project_cell(a.p, v.z, v.w, w.x); c[3] = project_cell(a.p, w.y, w.z, w.w);
                } else {
                    for (int k = 0; k < 4; k++) {  // the last, partial group of the cloud
                        const long long i = p0 + 4ll * gidx + k;
                        if (i < a.n) c[k] = project_cell(a.p, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    stage[4 * gidx + k] = c[k];
                    if (c[k] != PJ_INVALID) atomicAdd(&cur[c[k] >> PJ_REGION_BITS], 1u);
                }
            }
        }
        __syncthreads();
        // phase 2: exclusive scan of the histogram (thread t owns regions [t*per, t*per + per)), one global atomic per
        // non-empty region reserves the run's place in the region buffer
        {
            unsigned h[4] = {0u, 0u, 0u, 0u}, sum = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int r = tid * per + k;
                if (k < per && r < R) { h[k] = cur[r]; sum += h[k]; }
            }
            unsigned incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            unsigned wbase = 0;
            for (int w = 0; w < warp; w++) wbase += s_warp[w];
            unsigned run = wbase + incl - sum;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int r = tid * per + k;
                if (k < per && r < R) {
                    offs[r] = run;
                    cur[r] = run;
                    if (h[k]) base[r] = atomicAdd(a.count + r, h[k]);
                    run += h[k];
                }
            }
            if (tid == PJ_THREADS - 1) offs[R] = run;
        }
        __syncthreads();
        // phase 3: placement
#pragma unroll 4
        for (int j = 0; j < PJ_PPT; j++) {
            const unsigned c = stage[j * PJ_THREADS + tid];
            if (c != PJ_INVALID) sorted[atomicAdd(&cur[c >> PJ_REGION_BITS], 1u)] = c;
        }
        __syncthreads();
        // phase 4: copy the runs out (consecutive threads -> consecutive entries of a run)
        const unsigned total = offs[R];
        for (unsigned i = tid; i < total; i += PJ_THREADS) {
            const unsigned c = sorted[i], r = c >> PJ_REGION_BITS;
            const unsigned g = base[r] + (i - offs[r]);
            if (g < a.cap) a.buf[(size_t)r * a.cap + g] = c;
            else *a.flag = 1;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(PJ_THREADS) k_project_fill(PartArgs a, uint8_t *__restrict__ grid, size_t cells)
{
    extern __shared__ unsigned pj_bm[];  // PJ_REGION_CELLS / 32 words
    if (*a.flag) return;  // a region overflowed: the RED form redoes the grid
    const int r = blockIdx.x, tid = threadIdx.x;
    for (int w = tid; w < (int)(PJ_REGION_CELLS / 32); w += PJ_THREADS) pj_bm[w] = 0u;
    __syncthreads();
    const unsigned cnt = a.count[r];  // <= cap (else the flag is set)
    const unsigned *__restrict__ src = a.buf + (size_t)r * a.cap;
    const uint4 *__restrict__ src4 = reinterpret_cast<const uint4 *>(src);  // cap is a multiple of 4: every region buffer is 16-byte aligned
    const unsigned n4 = cnt / 4;
    for (unsigned i = tid; i < n4; i += PJ_THREADS) {
        const uint4 c = __ldcs(src4 + i);
        atomicOr(&pj_bm[(c.x & (PJ_REGION_CELLS - 1)) >> 5], 1u << (c.x & 31));
        atomicOr(&pj_bm[(c.y & (PJ_REGION_CELLS - 1)) >> 5], 1u << (c.y & 31));
        atomicOr(&pj_bm[(c.z & (PJ_REGION_CELLS - 1)) >> 5], 1u << (c.z & 31));
        atomicOr(&pj_bm[(c.w & (PJ_REGION_CELLS - 1)) >> 5], 1u << (c.w & 31));
    }
    for (unsigned i = n4 * 4 + tid; i < cnt; i += PJ_THREADS) {
        const unsigned c = src[i];
        atomicOr(&pj_bm[(c & (PJ_REGION_CELLS - 1)) >> 5], 1u << (c & 31));
    }
    __syncthreads();
    // expand: 16 cells per thread-iteration (the grid pointer is 16-byte aligned and regions start on multiples of 2^19 cells)
    const size_t c0 = (size_t)r << PJ_REGION_BITS;
    const size_t c1 = c0 + PJ_REGION_CELLS < cells ? c0 + PJ_REGION_CELLS : cells;
    const unsigned short *bm16 = reinterpret_cast<const unsigned short *>(pj_bm);
    const unsigned full = (unsigned)((c1 - c0) / 16);
    uint4 *__restrict__ dst = reinterpret_cast<uint4 *>(grid + c0);
    for (unsigned i = tid; i < full; i += PJ_THREADS) {
        const unsigned b = bm16[i];
        uint4 v;
        v.x = ((b & 0xFu) * 0x00204081u) & 0x01010101u;
        v.y = (((b >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
        v.z = (((b >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
        v.w = (((b >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
        __stcs(dst + i, v);
    }
    for (size_t c = c0 + (size_t)full * 16 + tid; c < c1; c += PJ_THREADS) grid[c] = (uint8_t)((pj_bm[(c - c0) >> 5] >> ((c - c0) & 31)) & 1u);
}

// clears the bit grid of the RED form only when the partition form gave up
__global__ void __launch_bounds__(256) k_clear_if(unsigned *__restrict__ w, size_t n, const int *__restrict__ run_flag)
{
    if (*run_flag == 0) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) w[i] = 0u;
}

#define FX_PROJ_PART_MIN_POINTS (1 << 22) /* below this the three extra launches and the fixed scratch do not pay */

#define FX_PROJ_BITS_MIN_CELLS (32u << 20) /* byte grids up to 32 MiB stay L2-resident: plain byte stores win */

template <typename Emit>
static void launch_project(fx_context *ctx, const float *pts, int64_t n, int stride_floats, const ProjParams &p, Emit emit,
                           cudaStream_t st, const int *run_flag = nullptr)
{
    const bool al16 = ((uintptr_t)pts & 15u) == 0;
    const int maxb = ctx->sm_count * 8;  // 8 resident CTAs of 256 threads per SM: one full wave
    if (stride_floats == 4 && al16) {
        long long want = (n + 1023) / 1024;
        int blocks = (int)(want < maxb ? (want > 0 ? want : 1) : maxb);
        k_project_f4<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(pts), n, p, emit, run_flag);
    } else if (stride_floats == 3 && al16) {
        long long want = (n / 4 + 511) / 512;
        int blocks = (int)(want < maxb ? (want > 0 ? want : 1) : maxb);
        k_project_f3<<<blocks, 256, 0, st>>>(pts, n, p, emit, run_flag);
    } else {
        long long want = (n + 255) / 256;
        int blocks = (int)(want < maxb ? want : maxb);
        k_project_generic<<<blocks, 256, 0, st>>>(pts, n, stride_floats, p, emit, run_flag);
    }
}

extern "C" int fx_project(fx_context *ctx, const float *pts, int64_t n, int stride_floats, const float *h_affine3x4,
                          float zmin, float zmax, float ox, float oy, float reso, int W, int H, uint8_t *grid,
                          int clear_first, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (n < 0 || (n > 0 && !pts) || !h_affine3x4 || !grid || W <= 0 || H <= 0 || stride_floats < 3 || !(reso > 0.f))
        return fx_set_err(ctx, FX_ERR_ARG, "fx_project: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t cells = (size_t)W * H;
    ProjParams p;
    memcpy(p.a, h_affine3x4, sizeof(p.a));
    p.zmin = zmin; p.zmax = zmax; p.ox = ox; p.oy = oy; p.reso = reso; p.W = W; p.H = H;
    const bool bit_form = clear_first && n > 0 && cells > FX_PROJ_BITS_MIN_CELLS && ((uintptr_t)grid & 15u) == 0;
    if (bit_form) {
        const size_t words = (cells + 31) / 32 + 4;
        if (ctx->proj_bits_cap < words) {
            if (ctx->proj_bits) cudaFree(ctx->proj_bits);
            ctx->proj_bits = nullptr; ctx->proj_bits_cap = 0;
            FX_CUDA(ctx, cudaMalloc(&ctx->proj_bits, words * sizeof(unsigned)));
            ctx->proj_bits_cap = words;
        }
        const bool al16 = ((uintptr_t)pts & 15u) == 0;
        const int R = (int)((cells + PJ_REGION_CELLS - 1) >> PJ_REGION_BITS);
        const int *run_flag = nullptr;
        const char *env = getenv("FUXI_B200_PROJ_PART");  // tuning experiments only: 0 = always the RED form
        if (n >= FX_PROJ_PART_MIN_POINTS && n < (1ll << 31) && al16 && (stride_floats == 3 || stride_floats == 4) && R <= PJ_MAX_REGIONS &&
            cells <= (1ull << 30) && !(env && env[0] == '0')) {
            // region buffers: twice the mean load (+ slack), a multiple of 4 entries
            const unsigned cap = (unsigned)(((2 * (size_t)n / R + 16384) + 3) & ~(size_t)3);
            const size_t need = (size_t)R * cap * 4 + (size_t)(R + 16) * 4;
            int rc = fx_grow_bytes(ctx, (void **)&ctx->proj_part, &ctx->proj_part_bytes, need);
            if (rc) return rc;
            PartArgs a;
            a.p = p; a.n = n; a.R = R; a.cap = cap;
            a.buf = ctx->proj_part;
            a.count = ctx->proj_part + (size_t)R * cap;
            a.flag = reinterpret_cast<int *>(a.count + R);
            FX_CUDA(ctx, cudaMemsetAsync(a.count, 0, (size_t)(R + 16) * 4, st));
            const size_t smA = (size_t)(2 * PJ_CHUNK + 3 * R + 1) * 4, smB = PJ_REGION_CELLS / 8;
            if (!ctx->proj_attr_set) {
                FX_CUDA(ctx, cudaFuncSetAttribute(k_project_part<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((2 * PJ_CHUNK + 3 * PJ_MAX_REGIONS + 1) * 4)));
                FX_CUDA(ctx, cudaFuncSetAttribute(k_project_part<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((2 * PJ_CHUNK + 3 * PJ_MAX_REGIONS + 1) * 4)));
                FX_CUDA(ctx, cudaFuncSetAttribute(k_project_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smB));
                ctx->proj_attr_set = 1;
            }
            const long long nchunks = (n + PJ_CHUNK - 1) / PJ_CHUNK;
            const int blocksA = (int)(nchunks < (long long)ctx->sm_count * 3 ? nchunks : (long long)ctx->sm_count * 3);
            if (stride_floats == 4) k_project_part<4><<<blocksA, PJ_THREADS, smA, st>>>(pts, a);
            else k_project_part<3><<<blocksA, PJ_THREADS, smA, st>>>(pts, a);
            FX_LAUNCH_CHECK(ctx);
            k_project_fill<<<R, PJ_THREADS, smB, st>>>(a, grid, cells);
            FX_LAUNCH_CHECK(ctx);
            run_flag = a.flag;  // the launches below return at once unless a region buffer overflowed
            k_clear_if<<<ctx->sm_count * 4, 256, 0, st>>>(ctx->proj_bits, words, run_flag);
            FX_LAUNCH_CHECK(ctx);
        } else {
            FX_CUDA(ctx, cudaMemsetAsync(ctx->proj_bits, 0, words * sizeof(unsigned), st));
        }
        launch_project(ctx, pts, n, stride_floats, p, EmitBit{ctx->proj_bits}, st, run_flag);
        FX_LAUNCH_CHECK(ctx);
        size_t want = (cells / 16 + 255) / 256;
        int blocks = (int)(want < (size_t)ctx->sm_count * 16 ? (want ? want : 1) : (size_t)ctx->sm_count * 16);
        k_bits_to_bytes<<<blocks, 256, 0, st>>>(reinterpret_cast<const unsigned short *>(ctx->proj_bits), grid, cells, run_flag);
        FX_LAUNCH_CHECK(ctx);
        return FX_OK;
    }
    if (clear_first) FX_CUDA(ctx, cudaMemsetAsync(grid, 0, cells, st));
    if (n == 0) return FX_OK;
    launch_project(ctx, pts, n, stride_floats, p, EmitByte{grid}, st);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}
