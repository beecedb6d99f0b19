// project.cu -- point cloud -> height-sliced 2D occupancy grid (sm_100a).
//
// Fuses the camera->earth rigid transform and z filter of scripts/plc_point2_st.py:244-256 with the
// 3D->2D projection the reference delegates to octomap_server / ccmapping (launch/map_st.launch:3,
// launch/map_ccst.launch:2).  One pass over the cloud: coalesced 16-byte loads (float4 directly, or
// three float4 per four packed-xyz points), 3x4 affine in registers with explicit round-to-nearest
// mul/add/sub/div (no FMA contraction, so the numpy float32 oracle is reproduced bit for bit), then a
// byte store of 1 into the grid.  The store is idempotent, so the result does not depend on scatter
// order and needs no atomics.  HBM-bound: N*16 (or N*12) bytes read + the W*H grid.
#include "common.cuh"

struct ProjParams {
    float a[12];
    float zmin, zmax, ox, oy, reso;
    int W, H;
};

__device__ __forceinline__ void project_one(const ProjParams &p, float x, float y, float z, uint8_t *__restrict__ grid)
{
    // e_k = ((a_k0*x + a_k1*y) + a_k2*z) + a_k3
    float ez = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.a[8], x), __fmul_rn(p.a[9], y)), __fmul_rn(p.a[10], z)), p.a[11]);
    if (!(ez > p.zmin && ez <= p.zmax)) return;
    float ex = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.a[0], x), __fmul_rn(p.a[1], y)), __fmul_rn(p.a[2], z)), p.a[3]);
    float ey = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.a[4], x), __fmul_rn(p.a[5], y)), __fmul_rn(p.a[6], z)), p.a[7]);
    float fx = floorf(__fdiv_rn(__fsub_rn(ex, p.ox), p.reso));
    float fy = floorf(__fdiv_rn(__fsub_rn(ey, p.oy), p.reso));
    if (!(fx >= 0.f && fx < (float)p.W && fy >= 0.f && fy < (float)p.H)) return;  // also rejects NaN
    grid[(size_t)(int)fx * p.H + (int)fy] = 1;
}

// stride 4: one float4 per point
__global__ void __launch_bounds__(256) k_project_f4(const float4 *__restrict__ pts, long long n, ProjParams p,
                                                    uint8_t *__restrict__ grid)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // two independent loads in flight per thread
    for (; i + stride < n; i += 2 * stride) {
        float4 v0 = __ldcs(pts + i), v1 = __ldcs(pts + i + stride);
        project_one(p, v0.x, v0.y, v0.z, grid);
        project_one(p, v1.x, v1.y, v1.z, grid);
    }
    if (i < n) {
        float4 v = __ldcs(pts + i);
        project_one(p, v.x, v.y, v.z, grid);
    }
}

// stride 3 (PointCloud2 packed xyz, plc_point2_st.py:112-138): four points = three float4
__global__ void __launch_bounds__(256) k_project_f3(const float *__restrict__ pts, long long n, ProjParams p,
                                                    uint8_t *__restrict__ grid)
{
    const long long ngroups = n / 4;
    const float4 *__restrict__ p4 = reinterpret_cast<const float4 *>(pts);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
        float4 a = __ldcs(p4 + 3 * g), b = __ldcs(p4 + 3 * g + 1), c = __ldcs(p4 + 3 * g + 2);
        project_one(p, a.x, a.y, a.z, grid);
        project_one(p, a.w, b.x, b.y, grid);
        project_one(p, b.z, b.w, c.x, grid);
        project_one(p, c.y, c.z, c.w, grid);
    }
    // tail (< 4 points)
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - ngroups * 4)) {
        long long i = ngroups * 4 + threadIdx.x;
        project_one(p, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], grid);
    }
}

// unaligned / generic stride fallback (scalar loads)
__global__ void __launch_bounds__(256) k_project_generic(const float *__restrict__ pts, long long n, int sf, ProjParams p,
                                                         uint8_t *__restrict__ grid)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        project_one(p, pts[(size_t)i * sf], pts[(size_t)i * sf + 1], pts[(size_t)i * sf + 2], grid);
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
    if (clear_first) FX_CUDA(ctx, cudaMemsetAsync(grid, 0, (size_t)W * H, st));
    if (n == 0) return FX_OK;
    ProjParams p;
    memcpy(p.a, h_affine3x4, sizeof(p.a));
    p.zmin = zmin; p.zmax = zmax; p.ox = ox; p.oy = oy; p.reso = reso; p.W = W; p.H = H;
    const bool al16 = ((uintptr_t)pts & 15u) == 0;
    const int maxb = ctx->sm_count * 8;  // 8 resident CTAs of 256 threads per SM: one full wave
    if (stride_floats == 4 && al16) {
        long long want = (n + 511) / 512;
        int blocks = (int)(want < maxb ? (want > 0 ? want : 1) : maxb);
        k_project_f4<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(pts), n, p, grid);
    } else if (stride_floats == 3 && al16) {
        long long want = (n / 4 + 255) / 256;
        int blocks = (int)(want < maxb ? (want > 0 ? want : 1) : maxb);
        k_project_f3<<<blocks, 256, 0, st>>>(pts, n, p, grid);
    } else {
        long long want = (n + 255) / 256;
        int blocks = (int)(want < maxb ? want : maxb);
        k_project_generic<<<blocks, 256, 0, st>>>(pts, n, stride_floats, p, grid);
    }
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}
