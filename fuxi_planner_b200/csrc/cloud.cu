// cloud.cu -- upstream cloud conditioning on the device:
//   fx_cloud_filter      PassThrough(z) -> VoxelGrid -> RadiusOutlierRemoval, the chain of src/chen_filter_rgb.cpp:52-71
//   fx_distance_filter   convert_plc.distance_filter, scripts/plc_point2_st.py:139-148 (a17)
//
// VoxelGrid without a sort: the voxel index space (PCL's idx = i + j*div_x + k*div_x*div_y) is a bitmap; a prefix
// count over the bitmap ranks the occupied voxels in PCL's output order (ascending idx), points accumulate into
// rank-indexed integer sums (order-independent, so the result is deterministic), and the same bitmap + ranks are the
// neighbour structure of the radius filter: a centroid lies inside its voxel, so everything within `radius` of it lies
// in a fixed window of voxels whose occupied cells are consecutive ranks row by row.
#include <math.h>

#include "common.cuh"

struct CloudState {
    unsigned bb[6];  // ordered-uint encodings of min x,y,z / max x,y,z over the points that pass the z filter
    int min_b[3];
    int div[3];
    unsigned total_bits;  // voxel index space actually used (0 when nothing passes or on overflow)
    unsigned n_pass, n_vox, n_keep;
    long long status;  // 0 ok; > 0: voxel index space needed (exceeds capacity); -1: coordinates outside int range
    unsigned ticket[2];  // "last CTA of the launch" counters of the two prefix-count kernels (zeroed by k_cl_reset)
};

struct CloudArgs {
    const float *pts;
    long long n;
    int stride, rgb_off;
    float lo, hi;
    float inv[3];
};

__device__ __forceinline__ unsigned f2ord(float f)
{
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }

// PassThrough<PointT>::applyFilterIndices: a point survives iff x, y, z are finite and !(z < lo || z > hi)
__device__ __forceinline__ bool cl_load(const CloudArgs &a, long long i, float &x, float &y, float &z)
{
    const float *p = a.pts + i * a.stride;
    x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    return isfinite(x) && isfinite(y) && isfinite(z) && !(z < a.lo || z > a.hi);
}

__global__ void k_cl_reset(CloudState *s)
{
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; k++) s->bb[k] = 0xFFFFFFFFu, s->bb[3 + k] = 0u, s->min_b[k] = 0, s->div[k] = 0;
        s->total_bits = s->n_pass = s->n_vox = s->n_keep = 0;
        s->status = 0;
        s->ticket[0] = s->ticket[1] = 0;
    }
}

__global__ void __launch_bounds__(256) k_cl_bbox(CloudArgs a, CloudState *s)
{
    unsigned mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0, 0, 0}, cnt = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        float v[3];
        if (!cl_load(a, i, v[0], v[1], v[2])) continue;
        cnt++;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            unsigned o = f2ord(v[k]);
            mn[k] = min(mn[k], o), mx[k] = max(mx[k], o);
        }
    }
    // warp reduce, then one set of atomics per CTA: the seven result words are hot addresses, and same-address atomics
    // serialise in L2 (one set per warp made this kernel 45 us on a 307 k-point frame)
    __shared__ unsigned s_mn[3][8], s_mx[3][8], s_c[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const unsigned lo = __reduce_min_sync(0xFFFFFFFFu, mn[k]), hi = __reduce_max_sync(0xFFFFFFFFu, mx[k]);
        if (lane == 0) s_mn[k][warp] = lo, s_mx[k][warp] = hi;
    }
    if (lane == 0) s_c[warp] = cnt;
    __syncthreads();
    if (warp == 0) {
        const unsigned c = lane < 8 ? s_c[lane] : 0u;
        const unsigned tot = __reduce_add_sync(0xFFFFFFFFu, c);
        if (tot == 0) return;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const unsigned lo = __reduce_min_sync(0xFFFFFFFFu, lane < 8 ? s_mn[k][lane] : 0xFFFFFFFFu);
            const unsigned hi = __reduce_max_sync(0xFFFFFFFFu, lane < 8 ? s_mx[k][lane] : 0u);
            if (lane == 0) atomicMin(&s->bb[k], lo), atomicMax(&s->bb[3 + k], hi);
        }
        if (lane == 0) atomicAdd(&s->n_pass, tot);
    }
}

// VoxelGrid<PointT>::applyFilter: min_b = floor(min_p * inverse_leaf), div_b = max_b - min_b + 1
// Evaluated by thread 0 of EVERY CTA of k_cl_zero_bits (same inputs, same result; CTA 0 also stores it): the separate
// one-thread launch that used to do this was 2.6 us of a 90 us frame.  Returns the voxel index space in use.
__device__ unsigned cl_setup(CloudState *s, float ix, float iy, float iz, unsigned long long cap_bits, bool store)
{
    if (s->n_pass == 0) return 0u;
    const float inv[3] = {ix, iy, iz};
    unsigned long long total = 1;
    int mb[3] = {0, 0, 0}, dv[3] = {0, 0, 0};
    long long status = 0;
    for (int k = 0; k < 3; k++) {
        float lo = floorf(__fmul_rn(ord2f(s->bb[k]), inv[k])), hi = floorf(__fmul_rn(ord2f(s->bb[3 + k]), inv[k]));
        if (!(fabsf(lo) < 1.0e9f) || !(fabsf(hi) < 1.0e9f)) { status = -1; break; }
        mb[k] = (int)lo;
        dv[k] = (int)hi - (int)lo + 1;
        total *= (unsigned long long)dv[k];
        if (total > (1ull << 40)) break;
    }
    if (status == 0 && total > cap_bits) status = (long long)total;
    const unsigned used = status == 0 ? (unsigned)total : 0u;
    if (store) {
        if (status != -1)
            for (int k = 0; k < 3; k++) s->min_b[k] = mb[k], s->div[k] = dv[k];
        s->status = status;
        s->total_bits = used;
    }
    return used;
}

__device__ __forceinline__ unsigned cl_voxel(const CloudState *s, const CloudArgs &a, float x, float y, float z)
{
    // voxel_grid.hpp: ijk = static_cast<int>(floor(p * inverse_leaf_size) - static_cast<float>(min_b))
    int i = (int)(floorf(__fmul_rn(x, a.inv[0])) - (float)s->min_b[0]);
    int j = (int)(floorf(__fmul_rn(y, a.inv[1])) - (float)s->min_b[1]);
    int k = (int)(floorf(__fmul_rn(z, a.inv[2])) - (float)s->min_b[2]);
    return (unsigned)i + (unsigned)s->div[0] * ((unsigned)j + (unsigned)s->div[1] * (unsigned)k);
}

// zero the words a scan will read: whole 256-bit groups
__global__ void k_cl_zero_bits(unsigned *bits, CloudState *s, float ix, float iy, float iz, unsigned long long cap_bits)
{
    __shared__ unsigned s_bits;
    if (threadIdx.x == 0) s_bits = cl_setup(s, ix, iy, iz, cap_bits, blockIdx.x == 0);
    __syncthreads();
    const size_t words = ((size_t)s_bits + 255) / 256 * 8;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x) bits[w] = 0;
}

#define CL_SMEM_BITS (1u << 18) /* voxel index spaces up to this many bits are marked in shared memory first (32 KB) */
__global__ void __launch_bounds__(256) k_cl_mark(CloudArgs a, const CloudState *s, unsigned *bits)
{
    __shared__ unsigned s_bits[CL_SMEM_BITS / 32];
    const unsigned total = s->total_bits;
    if (total == 0) return;
    // A depth frame falls into a few hundred bitmap words; RED.ORs to the same 128-byte line serialise in L2 (52 us for
    // one frame).  Small index spaces are therefore marked in a per-CTA shared-memory copy whose non-zero words are
    // OR-ed into the global bitmap once.
    const bool local = total <= CL_SMEM_BITS;
    const unsigned words = (total + 31) / 32;
    if (local) {
        for (unsigned w = threadIdx.x; w < words; w += blockDim.x) s_bits[w] = 0;
        __syncthreads();
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        float x, y, z;
        if (!cl_load(a, i, x, y, z)) continue;
        const unsigned v = cl_voxel(s, a, x, y, z), m = 1u << (v & 31);
        if (local) {
            if (!(s_bits[v >> 5] & m)) atomicOr(&s_bits[v >> 5], m);
        } else if (!(bits[v >> 5] & m)) atomicOr(&bits[v >> 5], m);
    }
    if (local) {
        __syncthreads();
        for (unsigned w = threadIdx.x; w < words; w += blockDim.x) {
            const unsigned m = s_bits[w];
            if (m && (bits[w] & m) != m) atomicOr(&bits[w], m);
        }
    }
}

// ---- rank of set bits: exclusive prefix count per 256-bit group, three kernels (group counts inside 1024-group
// chunks, the chunk totals, then the final per-group base + enumeration of the set bits) -----------------------------
__device__ __forceinline__ unsigned popc8(const unsigned *bits, size_t g, unsigned w[8])
{
    const uint4 a = *reinterpret_cast<const uint4 *>(bits + g * 8), b = *reinterpret_cast<const uint4 *>(bits + g * 8 + 4);
    w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w, w[4] = b.x, w[5] = b.y, w[6] = b.z, w[7] = b.w;
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) c += __popc(w[k]);
    return c;
}

__device__ __forceinline__ unsigned block_excl_scan_1024(unsigned v, unsigned *s_w, unsigned &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned t = s_w[lane], u = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned q = __shfl_up_sync(0xFFFFFFFFu, u, o);
            if (lane >= o) u += q;
        }
        s_w[lane] = u - t;
        if (lane == 31) s_w[32] = u;
    }
    __syncthreads();
    const unsigned r = inc - v + s_w[wid];
    total = s_w[32];
    __syncthreads();
    return r;
}

// per 256-bit group the set bits before it inside its 1024-group chunk, per chunk its total; the CTA that finishes last
// turns the chunk totals into exclusive prefixes and publishes the grand total (was a second, one-CTA launch)
__global__ void __launch_bounds__(1024) k_scan_groups(const unsigned *bits, const unsigned *nbits, unsigned *gpref, unsigned *chunk,
                                                      unsigned *out_total, unsigned *ticket)
{
    __shared__ unsigned s_w[33];
    __shared__ bool s_last;
    const size_t groups = ((size_t)*nbits + 255) / 256, chunks = (groups + 1023) / 1024;
    for (size_t c = blockIdx.x; c < chunks; c += gridDim.x) {
        const size_t g = c * 1024 + threadIdx.x;
        unsigned w[8], cnt = g < groups ? popc8(bits, g, w) : 0u, total;
        const unsigned ex = block_excl_scan_1024(cnt, s_w, total);
        if (g < groups) gpref[g] = ex;
        if (threadIdx.x == 0) chunk[c] = total;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const size_t per = (chunks + 1023) / 1024;  // consecutive chunks per thread
    const size_t b = threadIdx.x * per, e = min(b + per, chunks);
    unsigned sum = 0;
    for (size_t c = b; c < e; c++) sum += __ldcg(chunk + c);
    unsigned total, run = block_excl_scan_1024(sum, s_w, total);
    for (size_t c = b; c < e; c++) {
        unsigned t = __ldcg(chunk + c);
        chunk[c] = run;
        run += t;
    }
    if (threadIdx.x == 0) { *out_total = total; *ticket = 0; }
}

// MODE 0: vidx[rank] = bit index.  MODE 1: out[rank] = src[bit index] (rank < cap).  One thread per bitmap word (eight
// consecutive lanes share a 256-bit group and take their in-group prefix from each other by shuffles); the first lane of
// a group also finalises gpref[g].  (One thread per group walked up to 256 bits one after the other: 75 us per frame.)
template <int MODE>
__global__ void __launch_bounds__(256) k_scan_emit(const unsigned *bits, const unsigned *nbits, unsigned *gpref, const unsigned *chunk,
                                                   unsigned *vidx, const float4 *src, float4 *out, long long cap,
                                                   unsigned long long *zero_acc, const unsigned *n_vox)
{
    if (zero_acc) {  // the voxel accumulators of the next kernel (was a launch of its own)
        const size_t na = (size_t)*n_vox * 5;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < na; i += (size_t)gridDim.x * blockDim.x) zero_acc[i] = 0;
    }
    const size_t groups = ((size_t)*nbits + 255) / 256, nwords = groups * 8;
    const size_t wpad = (nwords + 31) & ~(size_t)31;  // whole warps: the shuffles below need every lane
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < wpad; w += (size_t)gridDim.x * blockDim.x) {
        const bool live = w < nwords;
        const size_t g = w >> 3;
        unsigned m = live ? bits[w] : 0u;
        const unsigned c = __popc(m);
        // exclusive prefix of c inside the 8-lane segment
        unsigned inc = c;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o, 8);
            if ((threadIdx.x & 7) >= o) inc += t;
        }
        if (!live) continue;
        const unsigned base = gpref[g] + chunk[g >> 10];
        unsigned r = base + inc - c;
        while (m) {
            const unsigned idx = (unsigned)(w * 32) + (__ffs(m) - 1);
            m &= m - 1;
            if (MODE == 0) vidx[r] = idx;
            else if ((long long)r < cap) out[r] = src[idx];
            r++;
        }
        __syncwarp(__activemask());
        if ((w & 7) == 0) gpref[g] = base;  // after every lane of the group has read the unfinalised value
    }
}

// set bits strictly below bit index idx
__device__ __forceinline__ unsigned cl_rank(const unsigned *bits, const unsigned *gpref, unsigned idx)
{
    const unsigned g = idx >> 8, wi = (idx >> 5) & 7;
    unsigned w[8];
    const uint4 a = *reinterpret_cast<const uint4 *>(bits + (size_t)g * 8), b = *reinterpret_cast<const uint4 *>(bits + (size_t)g * 8 + 4);
    w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w, w[4] = b.x, w[5] = b.y, w[6] = b.z, w[7] = b.w;
    unsigned r = gpref[g];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (k < (int)wi) r += __popc(w[k]);
        else if (k == (int)wi) r += __popc(w[k] & ((1u << (idx & 31)) - 1u));
    }
    return r;
}

// accumulators per voxel rank: 5 x u64 = {sum x, sum y, sum z (2^-24 m fixed point), r << 32 | g, b << 32 | count}
#define CL_FIX 16777216.0
__global__ void __launch_bounds__(256) k_cl_accum(CloudArgs a, const CloudState *s, const unsigned *bits, const unsigned *gpref,
                                                  unsigned long long *acc)
{
    if (s->total_bits == 0) return;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        float x, y, z;
        if (!cl_load(a, i, x, y, z)) continue;
        const unsigned r = cl_rank(bits, gpref, cl_voxel(s, a, x, y, z));
        unsigned long long *q = acc + (size_t)r * 5;
        atomicAdd(q + 0, (unsigned long long)__double2ll_rn((double)x * CL_FIX));
        atomicAdd(q + 1, (unsigned long long)__double2ll_rn((double)y * CL_FIX));
        atomicAdd(q + 2, (unsigned long long)__double2ll_rn((double)z * CL_FIX));
        unsigned c = a.rgb_off >= 0 ? __float_as_uint(__ldg(a.pts + i * a.stride + a.rgb_off)) : 0u;
        if (a.rgb_off >= 0) atomicAdd(q + 3, ((unsigned long long)((c >> 16) & 255u) << 32) | ((c >> 8) & 255u));
        atomicAdd(q + 4, ((unsigned long long)(c & 255u) << 32) | 1ull);
    }
}

__global__ void k_cl_centroid(const unsigned long long *acc, const unsigned *n_vox, float4 *vox)
{
    const size_t n = *n_vox;
    for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < n; r += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long *q = acc + r * 5;
        const unsigned cnt = (unsigned)q[4];
        const double dn = (double)cnt;
        float4 o;
        o.x = (float)(((double)(long long)q[0] / dn) * (1.0 / CL_FIX));
        o.y = (float)(((double)(long long)q[1] / dn) * (1.0 / CL_FIX));
        o.z = (float)(((double)(long long)q[2] / dn) * (1.0 / CL_FIX));
        const unsigned cr = (unsigned)(q[3] >> 32) / cnt, cg = (unsigned)q[3] / cnt, cb = (unsigned)(q[4] >> 32) / cnt;
        o.w = __uint_as_float((cr << 16) | (cg << 8) | cb);
        vox[r] = o;
    }
}

// RadiusOutlierRemoval<PointT>::applyFilterIndices: k = radiusSearch(p, r) (the point itself included, squared
// distance < r^2 in float: FLANN L2_Simple + RadiusResultSet); kept iff k > min_neighbors
// Eight lanes per centroid: the (2wz+1)(2wy+1) bitmap rows of its window are dealt round-robin to the lanes, the
// partial counts meet in a 3-step shuffle.  A CTA (8 warps x 4 centroids) takes 32 consecutive centroids and writes
// their keep bits as one word.
__global__ void __launch_bounds__(256) k_cl_ror(const float4 *vox, const unsigned *vidx, const unsigned *bits, const unsigned *gpref,
                                                const CloudState *s, int wx, int wy, int wz, float r2, int min_nb, unsigned *keep)
{
    __shared__ unsigned s_nib[8];
    const unsigned n = s->n_vox, nr = (n + 255u) & ~255u;
    const int dx = s->div[0], dy = s->div[1], dz = s->div[2];
    const unsigned lane = threadIdx.x & 31, sub = lane & 7, seg = lane >> 3, warp = threadIdx.x >> 5;
    const int ny = 2 * wy + 1, nrows = (2 * wz + 1) * ny;
    for (unsigned r0 = blockIdx.x * 32; r0 < nr; r0 += gridDim.x * 32) {
        const unsigned r = r0 + 4 * warp + seg;
        int cnt = 0;
        if (r < n) {
            const float4 c = vox[r];
            const unsigned v = vidx[r];
            const int i = (int)(v % (unsigned)dx), j = (int)((v / (unsigned)dx) % (unsigned)dy), k = (int)(v / ((unsigned)dx * (unsigned)dy));
            const int i0 = max(i - wx, 0), i1 = min(i + wx, dx - 1), len = i1 - i0 + 1;
            const unsigned lmask = len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u);
            for (int t = (int)sub; t < nrows; t += 8) {
                const int kk = k - wz + t / ny, jj = j - wy + t % ny;
                if (kk < 0 || kk >= dz || jj < 0 || jj >= dy) continue;
                const unsigned b0 = ((unsigned)kk * (unsigned)dy + (unsigned)jj) * (unsigned)dx + (unsigned)i0;
                const unsigned sh = b0 & 31u;
                const unsigned lo = bits[b0 >> 5], hi = (sh + (unsigned)len > 32u) ? bits[(b0 >> 5) + 1] : 0u;
                unsigned win = __funnelshift_r(lo, hi, sh) & lmask;
                if (!win) continue;
                unsigned q = cl_rank(bits, gpref, b0);
                for (; win; win &= win - 1, q++) {
                    const float4 p = vox[q];
                    const float ex = __fsub_rn(c.x, p.x), ey = __fsub_rn(c.y, p.y), ez = __fsub_rn(c.z, p.z);
                    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
                    cnt += d2 < r2;
                }
            }
        }
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 1);
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 2);
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 4);
        const unsigned m = __ballot_sync(0xFFFFFFFFu, sub == 0 && cnt > min_nb);  // bits 0, 8, 16, 24
        if (lane == 0) s_nib[warp] = (m & 1u) | ((m >> 7) & 2u) | ((m >> 14) & 4u) | ((m >> 21) & 8u);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned word = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) word |= s_nib[w] << (4 * w);
            keep[r0 >> 5] = word;
        }
        __syncthreads();
    }
}

// order-preserving compaction of the kept centroids: one thread per centroid, its output slot is the (not yet
// finalised) group prefix of the keep bitmap + the set bits below it inside the group
__global__ void __launch_bounds__(256) k_cl_compact(const unsigned *keep, const unsigned *n_vox, const unsigned *gpref2, const unsigned *chunk2,
                                                    const float4 *vox, float4 *out, long long cap, const CloudState *s, long long *counts)
{
    // every count is final before this kernel starts (was a one-thread launch of its own)
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = s->n_pass, counts[1] = s->n_vox, counts[2] = s->n_keep, counts[3] = s->status;
    const unsigned n = *n_vox;
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        if (!((keep[r >> 5] >> (r & 31)) & 1u)) continue;
        const unsigned g = r >> 8;
        unsigned d = gpref2[g] + chunk2[g >> 10];
        for (unsigned w = g * 8; w < (r >> 5); w++) d += __popc(keep[w]);
        d += __popc(keep[r >> 5] & ((1u << (r & 31)) - 1u));
        if ((long long)d < cap) out[d] = vox[r];
    }
}

static int cloud_reserve(fx_context *ctx, long long n, unsigned long long cap_bits)
{
    int rc;
    const size_t groups = (size_t)(cap_bits / 256) + 2, chunks = groups / 1024 + 2;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->cl_bits, &ctx->cl_bits_bytes, (size_t)(cap_bits / 8) + 128))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->cl_gpref, &ctx->cl_gpref_bytes, groups * 4))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->cl_chunk, &ctx->cl_chunk_bytes, chunks * 4))) return rc;
    const size_t nn = (size_t)(n > 0 ? n : 1), g2 = nn / 256 + 2, c2 = g2 / 1024 + 2;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->cl_acc, &ctx->cl_acc_bytes, nn * 40))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->cl_vox, &ctx->cl_vox_bytes, nn * 16))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->cl_vidx, &ctx->cl_vidx_bytes, nn * 4))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->cl_keep, &ctx->cl_keep_bytes, g2 * 32 + 128))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->cl_gpref2, &ctx->cl_gpref2_bytes, g2 * 4))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->cl_chunk2, &ctx->cl_chunk2_bytes, c2 * 4))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->cl_state, &ctx->cl_state_bytes, sizeof(CloudState)))) return rc;
    return FX_OK;
}

extern "C" int fx_cloud_reserve(fx_context *ctx, int64_t max_voxel_space)
{
    if (!ctx || max_voxel_space < 256 || max_voxel_space > (1ll << 31)) return fx_set_err(ctx, FX_ERR_ARG, "fx_cloud_reserve: voxel space must be in [256, 2^31]");
    ctx->cl_cap_bits = (unsigned long long)max_voxel_space;
    return FX_OK;
}

static inline int grid_for(fx_context *ctx, long long n, int threads, int per_sm)
{
    long long b = (n + threads - 1) / threads;
    long long cap = (long long)ctx->sm_count * per_sm;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int fx_cloud_filter(fx_context *ctx, const float *pts, int64_t n, const fx_cloud_params *p, float *out, int64_t cap,
                               int64_t *d_counts, void *stream)
{
    if (!ctx || !p || n < 0 || (n > 0 && !pts) || !d_counts || cap < 0 || (cap > 0 && !out)) return fx_set_err(ctx, FX_ERR_ARG, "fx_cloud_filter: bad argument");
    if (p->stride_floats < 3 || p->rgb_offset >= p->stride_floats || (p->rgb_offset >= 0 && p->rgb_offset < 3))
        return fx_set_err(ctx, FX_ERR_ARG, "fx_cloud_filter: stride_floats >= 3 and rgb_offset in [3, stride) or -1");
    if (!(p->leaf_x > 0.f) || !(p->leaf_y > 0.f) || !(p->leaf_z > 0.f) || !(p->radius > 0.0) || p->min_neighbors < 0)
        return fx_set_err(ctx, FX_ERR_ARG, "fx_cloud_filter: leaf sizes and radius must be positive");
    if (n >= (1ll << 32)) return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_cloud_filter: at most 2^32 - 1 points");
    // search window in voxels: a centroid lies inside its voxel
    const int wx = (int)floor(p->radius / (double)p->leaf_x) + 1, wy = (int)floor(p->radius / (double)p->leaf_y) + 1,
              wz = (int)floor(p->radius / (double)p->leaf_z) + 1;
    if (wx > 15) return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_cloud_filter: radius / leaf_x > 14 (window wider than one 32-bit word)");
    cudaStream_t st = (cudaStream_t)stream;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->cl_cap_bits) ctx->cl_cap_bits = 1ull << 28;
    int rc = cloud_reserve(ctx, n, ctx->cl_cap_bits);
    if (rc) return rc;
    CloudState *s = (CloudState *)ctx->cl_state;
    CloudArgs a;
    a.pts = pts, a.n = n, a.stride = p->stride_floats, a.rgb_off = p->rgb_offset, a.lo = p->pass_lo, a.hi = p->pass_hi;
    a.inv[0] = 1.0f / p->leaf_x, a.inv[1] = 1.0f / p->leaf_y, a.inv[2] = 1.0f / p->leaf_z;  // Array4f::Ones() / leaf_size
    const float r2 = (float)(p->radius * p->radius);  // pcl::KdTreeFLANN::radiusSearch: static_cast<float>(radius * radius)
    const int gp = grid_for(ctx, n, 256, 8), gw = ctx->sm_count * 8;
    auto enqueue = [&](cudaStream_t st) -> int {
        k_cl_reset<<<1, 32, 0, st>>>(s);
        FX_LAUNCH_CHECK(ctx);
        k_cl_bbox<<<grid_for(ctx, n, 1024, 4), 256, 0, st>>>(a, s);
        FX_LAUNCH_CHECK(ctx);
        k_cl_zero_bits<<<gw, 256, 0, st>>>(ctx->cl_bits, s, a.inv[0], a.inv[1], a.inv[2], ctx->cl_cap_bits);
        FX_LAUNCH_CHECK(ctx);
        k_cl_mark<<<gp, 256, 0, st>>>(a, s, ctx->cl_bits);
        FX_LAUNCH_CHECK(ctx);
        k_scan_groups<<<ctx->sm_count * 2, 1024, 0, st>>>(ctx->cl_bits, &s->total_bits, ctx->cl_gpref, ctx->cl_chunk, &s->n_vox, &s->ticket[0]);
        FX_LAUNCH_CHECK(ctx);
        k_scan_emit<0><<<gw, 256, 0, st>>>(ctx->cl_bits, &s->total_bits, ctx->cl_gpref, ctx->cl_chunk, ctx->cl_vidx, nullptr, nullptr, 0, ctx->cl_acc, &s->n_vox);
        FX_LAUNCH_CHECK(ctx);
        k_cl_accum<<<gp, 256, 0, st>>>(a, s, ctx->cl_bits, ctx->cl_gpref, ctx->cl_acc);
        FX_LAUNCH_CHECK(ctx);
        k_cl_centroid<<<gw, 256, 0, st>>>(ctx->cl_acc, &s->n_vox, ctx->cl_vox);
        FX_LAUNCH_CHECK(ctx);
        k_cl_ror<<<gw, 256, 0, st>>>(ctx->cl_vox, ctx->cl_vidx, ctx->cl_bits, ctx->cl_gpref, s, wx, wy, wz, r2, p->min_neighbors, ctx->cl_keep);
        FX_LAUNCH_CHECK(ctx);
        k_scan_groups<<<ctx->sm_count * 2, 1024, 0, st>>>(ctx->cl_keep, &s->n_vox, ctx->cl_gpref2, ctx->cl_chunk2, &s->n_keep, &s->ticket[1]);
        FX_LAUNCH_CHECK(ctx);
        k_cl_compact<<<gw, 256, 0, st>>>(ctx->cl_keep, &s->n_vox, ctx->cl_gpref2, ctx->cl_chunk2, ctx->cl_vox, (float4 *)out, cap, s, (long long *)d_counts);
        FX_LAUNCH_CHECK(ctx);
        return FX_OK;
    };
    // everything the 11 launches depend on: arguments, derived constants, scratch pointers
    struct { const float *pts; int64_t n; fx_cloud_params p; float *out; int64_t cap; int64_t *counts; unsigned long long cap_bits;
             void *b[11]; } key;
    memset(&key, 0, sizeof(key));
    key.pts = pts; key.n = n; key.p = *p; key.out = out; key.cap = cap; key.counts = d_counts; key.cap_bits = ctx->cl_cap_bits;
    void *bufs[11] = {ctx->cl_bits, ctx->cl_gpref, ctx->cl_chunk, ctx->cl_vidx, ctx->cl_keep, ctx->cl_gpref2, ctx->cl_chunk2, ctx->cl_acc, ctx->cl_vox, ctx->cl_state, nullptr};
    memcpy(key.b, bufs, sizeof(bufs));
    return fx_graph_run(ctx, FX_GRAPH_CLOUD, &key, sizeof(key), st, enqueue);
}

extern "C" int fx_cloud_filter_host(fx_context *ctx, const float *h_pts, int64_t n, const fx_cloud_params *p, float *h_out, int64_t cap,
                                    int64_t *h_counts)
{
    if (!ctx || !p || n < 0 || (n > 0 && !h_pts) || !h_counts || cap < 0 || (cap > 0 && !h_out)) return fx_set_err(ctx, FX_ERR_ARG, "fx_cloud_filter_host: bad argument");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->own_stream;
    const size_t in_bytes = (size_t)n * p->stride_floats * 4, out_cap = (size_t)(cap < n ? cap : n);
    int rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->d_pts, &ctx->d_pts_cap, in_bytes + 16))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->cl_out, &ctx->cl_out_bytes, out_cap * 16 + 64))) return rc;
    // through the pinned staging buffer (non-temporal host copy, api.cu): a copy straight from the caller's pageable
    // array runs at a third of the PCIe rate
    if (in_bytes) {
        if ((rc = fx_grow_pinned(ctx, in_bytes))) return rc;
        if ((rc = fx_staged_copy_in(ctx, (uint8_t *)ctx->d_pts, (const uint8_t *)h_pts, in_bytes, st))) return rc;
    }
    long long *d_counts = (long long *)((char *)ctx->cl_out + out_cap * 16);
    d_counts = (long long *)(((uintptr_t)d_counts + 15) & ~(uintptr_t)15);
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = fx_cloud_filter(ctx, ctx->d_pts, n, p, (float *)ctx->cl_out, (int64_t)out_cap, (int64_t *)d_counts, st))) return rc;
        FX_CUDA(ctx, cudaMemcpyAsync(h_counts, d_counts, 32, cudaMemcpyDeviceToHost, st));
        FX_CUDA(ctx, cudaStreamSynchronize(st));
        if (h_counts[3] <= 0) break;
        // the cloud's bounding box needs a larger voxel index space than reserved: grow once (PCL's own limit is 2^31)
        if (attempt || h_counts[3] > (1ll << 31))
            return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_cloud_filter: voxel index space %lld exceeds 2^31 (leaf size too small for the cloud's extent)", (long long)h_counts[3]);
        ctx->cl_cap_bits = (unsigned long long)h_counts[3];
    }
    if (h_counts[3] < 0) return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_cloud_filter: coordinates / leaf size outside the int range");
    const size_t kept = (size_t)h_counts[2] < out_cap ? (size_t)h_counts[2] : out_cap;
    if (kept) {
        FX_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->cl_out, kept * 16, cudaMemcpyDeviceToHost, st));
        FX_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return FX_OK;
}

// ---- distance_filter (a17) --------------------------------------------------------------------------------------------
// d = |p| in float64 exactly as np.linalg.norm(axis=1) computes it for three columns (sqrt((x*x + y*y) + z*z), one
// rounding per operation); keep d < dis; order = np.lexsort of the columns (x, y, z, d): by d, then z, then y, then x,
// stable.  Bitonic sort of (d, index) records padded with +inf to a power of two; ties on d look the point up.
struct DfRec {
    double d;
    unsigned idx, pad;
};
#define DF_INVALID 0xFFFFFFFFu
#define DF_LOCAL 2048

__device__ __forceinline__ bool df_less(const DfRec &a, const DfRec &b, const double *pts)
{
    if (a.d != b.d) return a.d < b.d;
    if (a.idx == DF_INVALID || b.idx == DF_INVALID) return a.idx < b.idx;
    const double *p = pts + (size_t)a.idx * 3, *q = pts + (size_t)b.idx * 3;
    if (p[2] != q[2]) return p[2] < q[2];
    if (p[1] != q[1]) return p[1] < q[1];
    if (p[0] != q[0]) return p[0] < q[0];
    return a.idx < b.idx;
}

__global__ void k_df_keys(const double *pts, long long n, long long np2, double dis, DfRec *rec, int *count)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < np2; i += (long long)gridDim.x * blockDim.x) {
        DfRec r;
        r.d = INFINITY, r.idx = DF_INVALID, r.pad = 0;
        bool keep = false;
        if (i < n) {
            const double x = pts[i * 3], y = pts[i * 3 + 1], z = pts[i * 3 + 2];
            const double d = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
            keep = d < dis;
            if (keep) r.d = d, r.idx = (unsigned)i;
        }
        rec[i] = r;
        const unsigned m = __ballot_sync(__activemask(), keep);
        if (keep && (threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) atomicAdd(count, __popc(m));
    }
}

__device__ __forceinline__ void df_cswap(DfRec &a, DfRec &b, bool asc, const double *pts)
{
    if (df_less(b, a, pts) == asc) {
        DfRec t = a;
        a = b;
        b = t;
    }
}

// full sort of each DF_LOCAL-record block (FULL) or the in-block tail j = DF_LOCAL/2 .. 1 of stage k
template <bool FULL>
__global__ void __launch_bounds__(DF_LOCAL / 2) k_df_local(DfRec *rec, const double *pts, unsigned k_stage)
{
    __shared__ DfRec s[DF_LOCAL];
    const size_t base = (size_t)blockIdx.x * DF_LOCAL;
    s[threadIdx.x] = rec[base + threadIdx.x];
    s[threadIdx.x + DF_LOCAL / 2] = rec[base + threadIdx.x + DF_LOCAL / 2];
    __syncthreads();
    for (unsigned k = FULL ? 2u : k_stage; k <= (FULL ? (unsigned)DF_LOCAL : k_stage); k <<= 1) {
        for (unsigned j = min(k >> 1, (unsigned)DF_LOCAL / 2); j > 0; j >>= 1) {
            const unsigned t = threadIdx.x, i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            const bool asc = (((base + i) & (size_t)k) == 0);
            df_cswap(s[i], s[i | j], asc, pts);
            __syncthreads();
        }
        if (!FULL) break;
    }
    rec[base + threadIdx.x] = s[threadIdx.x];
    rec[base + threadIdx.x + DF_LOCAL / 2] = s[threadIdx.x + DF_LOCAL / 2];
}

__global__ void k_df_global(DfRec *rec, const double *pts, long long np2, unsigned long long k, unsigned long long j)
{
    for (unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; t < (unsigned long long)np2 / 2;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        DfRec a = rec[i], b = rec[i | j];
        if (df_less(b, a, pts) == ((i & k) == 0)) rec[i] = b, rec[i | j] = a;
    }
}

__global__ void k_df_emit(const DfRec *rec, const double *pts, const int *count, double *out)
{
    const long long m = *count;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < m; r += (long long)gridDim.x * blockDim.x) {
        const double *p = pts + (size_t)rec[r].idx * 3;
        out[r * 3] = p[0], out[r * 3 + 1] = p[1], out[r * 3 + 2] = p[2];
    }
}

extern "C" int fx_distance_filter(fx_context *ctx, const double *pts, int64_t n, double dis, double *out, int32_t *d_count, void *stream)
{
    if (!ctx || n < 0 || (n > 0 && (!pts || !out)) || !d_count) return fx_set_err(ctx, FX_ERR_ARG, "fx_distance_filter: bad argument");
    if (n >= (1ll << 31)) return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_distance_filter: at most 2^31 - 1 points");
    cudaStream_t st = (cudaStream_t)stream;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    long long np2 = DF_LOCAL;
    while (np2 < n) np2 <<= 1;
    int rc = fx_grow_bytes(ctx, (void **)&ctx->df_rec, &ctx->df_rec_bytes, (size_t)np2 * sizeof(DfRec));
    if (rc) return rc;
    DfRec *rec = (DfRec *)ctx->df_rec;
    auto enqueue = [&](cudaStream_t st) -> int {
        FX_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(int), st));
        k_df_keys<<<grid_for(ctx, np2, 256, 8), 256, 0, st>>>(pts, n, np2, dis, rec, d_count);
        FX_LAUNCH_CHECK(ctx);
        const int nblk = (int)(np2 / DF_LOCAL);
        k_df_local<true><<<nblk, DF_LOCAL / 2, 0, st>>>(rec, pts, 0);
        FX_LAUNCH_CHECK(ctx);
        for (unsigned long long k = 2ull * DF_LOCAL; k <= (unsigned long long)np2; k <<= 1) {
            for (unsigned long long j = k >> 1; j >= DF_LOCAL; j >>= 1) {
                k_df_global<<<grid_for(ctx, np2 / 2, 256, 8), 256, 0, st>>>(rec, pts, np2, k, j);
                FX_LAUNCH_CHECK(ctx);
            }
            k_df_local<false><<<nblk, DF_LOCAL / 2, 0, st>>>(rec, pts, (unsigned)k);
            FX_LAUNCH_CHECK(ctx);
        }
        k_df_emit<<<grid_for(ctx, n > 0 ? n : 1, 256, 8), 256, 0, st>>>(rec, pts, d_count, out);
        FX_LAUNCH_CHECK(ctx);
        return FX_OK;
    };
    struct { const void *pts, *out, *cnt, *rec; int64_t n; double dis; } key = {pts, out, d_count, rec, n, dis};
    return fx_graph_run(ctx, FX_GRAPH_DFILTER, &key, sizeof(key), st, enqueue);
}

extern "C" int fx_distance_filter_host(fx_context *ctx, const double *h_pts, int64_t n, double dis, double *h_out, int64_t *h_count)
{
    if (!ctx || n < 0 || (n > 0 && (!h_pts || !h_out)) || !h_count) return fx_set_err(ctx, FX_ERR_ARG, "fx_distance_filter_host: bad argument");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->own_stream;
    const size_t bytes = (size_t)n * 24;
    int rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->d_pts, &ctx->d_pts_cap, bytes + 16))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->cl_out, &ctx->cl_out_bytes, bytes + 64))) return rc;
    int *d_count = (int *)((char *)ctx->cl_out + ((bytes + 15) & ~(size_t)15));
    // through the pinned staging buffer (non-temporal host copy, api.cu): a copy straight from the caller's pageable
    // array runs at a third of the PCIe rate
    if (bytes) {
        if ((rc = fx_grow_pinned(ctx, bytes))) return rc;
        if ((rc = fx_staged_copy_in(ctx, (uint8_t *)ctx->d_pts, (const uint8_t *)h_pts, bytes, st))) return rc;
    }
    if ((rc = fx_distance_filter(ctx, (const double *)ctx->d_pts, n, dis, (double *)ctx->cl_out, d_count, st))) return rc;
    int cnt = 0;
    FX_CUDA(ctx, cudaMemcpyAsync(&cnt, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    FX_CUDA(ctx, cudaStreamSynchronize(st));
    *h_count = cnt;
    if (cnt) {
        FX_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->cl_out, (size_t)cnt * 24, cudaMemcpyDeviceToHost, st));
        FX_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return FX_OK;
}

// ---- the cloud node's product: transformed + filtered + sorted cloud (scripts/plc_point2_st.py:243-256, 336-339; 351-362) --------
// camera branch (h_R9 != NULL):  b = (z_c + 0.12, -x_c, -y_c);  e_k = ((R_k0*b0 + R_k1*b1) + R_k2*b2) + t_k  (float64, one
//   rounding per operation; numpy hands the same product to BLAS, whose summation order is not specified: parity with the
//   reference's own lines is to a few ulps, with the restatement in oracle/hostref.py bit for bit);  keep e_z > zmin (:255-256)
// octomap-centres branch (h_R9 == NULL):  e = p  (:351-359)
// then q = e - c;  [box > 0: keep |q_k| < box on every axis (:361)];  distance_filter(q, dis) (:139-148);  out = q + c (:339, :362)
struct TfArgs {
    double R[9], t[3], c[3];
    double zmin, box;
    int camera, stride, is_f64;
};

__global__ void k_tf_points(const void *__restrict__ pts, long long n, TfArgs a, double *__restrict__ q)
{
    const double nan = __longlong_as_double(0x7FF8000000000000ll);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double x, y, z;
        if (a.is_f64) {
            const double *p = reinterpret_cast<const double *>(pts) + i * a.stride;
            x = p[0], y = p[1], z = p[2];
        } else {
            const float *p = reinterpret_cast<const float *>(pts) + i * a.stride;
            x = (double)p[0], y = (double)p[1], z = (double)p[2];
        }
        double e0 = x, e1 = y, e2 = z;
        bool keep = true;
        if (a.camera) {
            const double b0 = __dadd_rn(z, 0.12), b1 = -x, b2 = -y;
            e0 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.R[0], b0), __dmul_rn(a.R[1], b1)), __dmul_rn(a.R[2], b2)), a.t[0]);
            e1 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.R[3], b0), __dmul_rn(a.R[4], b1)), __dmul_rn(a.R[5], b2)), a.t[1]);
            e2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.R[6], b0), __dmul_rn(a.R[7], b1)), __dmul_rn(a.R[8], b2)), a.t[2]);
            keep = e2 > a.zmin;
        }
        const double q0 = __dsub_rn(e0, a.c[0]), q1 = __dsub_rn(e1, a.c[1]), q2 = __dsub_rn(e2, a.c[2]);
        if (a.box > 0.0) keep = keep && fabs(q0) < a.box && fabs(q1) < a.box && fabs(q2) < a.box;
        // a rejected point becomes NaN: its norm is NaN, `d < dis` is false, distance_filter drops it
        q[i * 3] = keep ? q0 : nan; q[i * 3 + 1] = keep ? q1 : nan; q[i * 3 + 2] = keep ? q2 : nan;
    }
}

__global__ void k_tf_add(double *__restrict__ out, const int *__restrict__ count, double c0, double c1, double c2)
{
    const long long n = *count;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        out[i * 3] = __dadd_rn(out[i * 3], c0);
        out[i * 3 + 1] = __dadd_rn(out[i * 3 + 1], c1);
        out[i * 3 + 2] = __dadd_rn(out[i * 3 + 2], c2);
    }
}

extern "C" int fx_transform_filter(fx_context *ctx, const void *pts, int64_t n, int stride, int is_f64, const double *h_R9,
                                   const double *h_t3, const double *h_c3, double zmin, double box, double dis, double *out,
                                   int32_t *d_count, void *stream)
{
    if (!ctx || n < 0 || (n > 0 && (!pts || !out)) || !d_count || stride < 3 || !h_c3 || (h_R9 && !h_t3))
        return fx_set_err(ctx, FX_ERR_ARG, "fx_transform_filter: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = fx_grow_bytes(ctx, (void **)&ctx->tf_q, &ctx->tf_q_bytes, (size_t)(n > 0 ? n : 1) * 24);
    if (rc) return rc;
    TfArgs a;
    memset(&a, 0, sizeof(a));
    a.camera = h_R9 != nullptr; a.stride = stride; a.is_f64 = is_f64 != 0; a.zmin = zmin; a.box = box;
    if (h_R9) { memcpy(a.R, h_R9, sizeof(a.R)); memcpy(a.t, h_t3, sizeof(a.t)); }
    memcpy(a.c, h_c3, sizeof(a.c));
    if (n > 0) {
        k_tf_points<<<grid_for(ctx, n, 256, 8), 256, 0, st>>>(pts, n, a, (double *)ctx->tf_q);
        FX_LAUNCH_CHECK(ctx);
    }
    rc = fx_distance_filter(ctx, (const double *)ctx->tf_q, n, dis, out, d_count, stream);
    if (rc) return rc;
    k_tf_add<<<grid_for(ctx, n > 0 ? n : 1, 256, 8), 256, 0, st>>>(out, d_count, a.c[0], a.c[1], a.c[2]);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}

extern "C" int fx_transform_filter_host(fx_context *ctx, const void *h_pts, int64_t n, int stride, int is_f64, const double *h_R9,
                                        const double *h_t3, const double *h_c3, double zmin, double box, double dis, double *h_out,
                                        int64_t *h_count)
{
    if (!ctx || n < 0 || (n > 0 && (!h_pts || !h_out)) || !h_count) return fx_set_err(ctx, FX_ERR_ARG, "fx_transform_filter_host: bad argument");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->own_stream;
    const size_t in_bytes = (size_t)n * stride * (is_f64 ? 8 : 4), out_bytes = (size_t)n * 24;
    int rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->d_pts, &ctx->d_pts_cap, in_bytes + 16))) return rc;
    if ((rc = fx_grow_bytes(ctx, (void **)&ctx->cl_out, &ctx->cl_out_bytes, out_bytes + 64))) return rc;
    int *d_count = (int *)((char *)ctx->cl_out + ((out_bytes + 15) & ~(size_t)15));
    // through the pinned staging buffer (non-temporal host copy, api.cu): a copy straight from the caller's pageable
    // array runs at a third of the PCIe rate
    if (in_bytes) {
        if ((rc = fx_grow_pinned(ctx, in_bytes))) return rc;
        if ((rc = fx_staged_copy_in(ctx, (uint8_t *)ctx->d_pts, (const uint8_t *)h_pts, in_bytes, st))) return rc;
    }
    if ((rc = fx_transform_filter(ctx, ctx->d_pts, n, stride, is_f64, h_R9, h_t3, h_c3, zmin, box, dis, (double *)ctx->cl_out, d_count, st))) return rc;
    int cnt = 0;
    FX_CUDA(ctx, cudaMemcpyAsync(&cnt, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    FX_CUDA(ctx, cudaStreamSynchronize(st));
    *h_count = cnt;
    if (cnt) {
        FX_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->cl_out, (size_t)cnt * 24, cudaMemcpyDeviceToHost, st));
        FX_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return FX_OK;
}
