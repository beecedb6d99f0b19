// inflate.cu -- obstacle inflation (binary dilation) on the uint8 [W][H] grid (sm_100a).
//
// Replaces scripts/global_planner_st.py:256-262 (9-point stencil {-r,0,+r}^2, step == r) and
// scripts/global_planner_ccst.py:442-448 (dense (2r+1)^2 square, step == 1).  Both stencils are
// product sets S x S, so the dilation is separable: OR over y offsets, then OR over x offsets.
//
// Fast path (H % 16 == 0, 16-byte aligned pointers, r <= 16): one CTA owns a tile of TX rows x 512
// cells.  Each warp streams whole 512-byte row segments with one 16-byte load per lane, squeezes the
// 16 cells to 16 bits ("> 0" test), does the y pass on a 48-bit window built from the neighbouring
// lanes' bits (shuffles; the two halo chunks come from one extra load on lanes 0 and 31), and parks
// the row's bits in shared memory.  After one barrier the x pass ORs 2r/step+1 rows of bits, expands
// back to bytes and writes 16 bytes per lane.  HBM traffic is ~ (1 + 2r/TX) B read + 1 B written per cell.
#include "common.cuh"

#define INF_TY 512     /* cells per tile row (32 lanes x 16 B) */
#define INF_MAXR 16
#define INF_MAXTX 64
#define INF_UNROLL 4

__device__ __forceinline__ unsigned bytes_to_bits16(uint4 v)
{
    // byte > 0 -> bit; 4 bytes -> 4 bits via a carry-free multiply
    unsigned r = 0;
    unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        unsigned nz = __vcmpne4(w[i], 0u) & 0x01010101u;
        r |= (((nz * 0x01020408u) >> 24) & 0xFu) << (4 * i);
    }
    return r;
}
__device__ __forceinline__ uint4 bits16_to_bytes(unsigned b)
{
    uint4 v;
    v.x = ((b & 0xFu) * 0x00204081u) & 0x01010101u;
    v.y = (((b >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
    v.z = (((b >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
    v.w = (((b >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
    return v;
}

__global__ void __launch_bounds__(256, 5) k_inflate_tiled(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                                                       int W, int H, int r, int step, int TX)
{
    __shared__ unsigned short rowbits[INF_MAXTX + 2 * INF_MAXR][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int x0 = blockIdx.y * TX, y0 = blockIdx.x * INF_TY;
    const int rows = TX + 2 * r;
    const int yl = y0 + 16 * lane;
    // pass 1: y dilation, one row per warp and step; INF_UNROLL rows are loaded before any is used so that every
    // lane keeps INF_UNROLL independent 16-byte loads in flight (HBM latency x bandwidth needs ~40 KB per SM)
    for (int rr0 = warp; rr0 < rows; rr0 += nwarps * INF_UNROLL) {
        uint4 v[INF_UNROLL], e[INF_UNROLL];
#pragma unroll
        for (int u = 0; u < INF_UNROLL; u++) {
            const int rr = rr0 + u * nwarps, x = x0 - r + rr;
            v[u] = make_uint4(0, 0, 0, 0); e[u] = make_uint4(0, 0, 0, 0);
            if (rr < rows && x >= 0 && x < W) {
                const uint8_t *row = in + (size_t)x * H;
                if (yl < H) v[u] = __ldcs(reinterpret_cast<const uint4 *>(row + yl));
                const int ye = lane == 0 ? y0 - 16 : y0 + INF_TY;
                if ((lane == 0 || lane == 31) && ye >= 0 && ye < H) e[u] = __ldg(reinterpret_cast<const uint4 *>(row + ye));
            }
        }
#pragma unroll
        for (int u = 0; u < INF_UNROLL; u++) {
            const int rr = rr0 + u * nwarps;
            const unsigned b = bytes_to_bits16(v[u]), extra = bytes_to_bits16(e[u]);
            unsigned prev = __shfl_up_sync(0xFFFFFFFFu, b, 1), next = __shfl_down_sync(0xFFFFFFFFu, b, 1);
            if (lane == 0) prev = extra;
            if (lane == 31) next = extra;
            const unsigned long long win = (unsigned long long)prev | ((unsigned long long)b << 16) | ((unsigned long long)next << 32);
            unsigned acc = 0;
            for (int s = -r; s <= r; s += step) acc |= (unsigned)(win >> (16 + s));
            if (rr < rows) rowbits[rr][lane] = (unsigned short)(acc & 0xFFFFu);
        }
    }
    __syncthreads();
    // pass 2: x dilation + expand + store
    for (int ox = warp; ox < TX; ox += nwarps) {
        const int x = x0 + ox;
        if (x >= W) break;
        unsigned acc = 0;
        for (int s = -r; s <= r; s += step) acc |= rowbits[ox + r + s][lane];
        if (yl < H) __stcs(reinterpret_cast<uint4 *>(out + (size_t)x * H + yl), bits16_to_bytes(acc));
    }
}

// ------------------------------------------------------------------------------------------------
// Register rolling-window form (the default for r <= 4, i.e. both reference configurations): no shared memory
// and no block barrier.  A warp owns a strip of 512 columns x INF_RPT rows and walks down the rows: INF_U rows of
// independent 16-byte loads are issued together (8 x 512 B in flight per warp), each row is squeezed to 16 bits per
// lane and y-dilated with the neighbouring lanes' bits, the last 2r+1 rows of bits live in a register ring, their
// OR is the output row, expanded and written with one 512-byte streaming store per warp.
// (A cp.async.bulk + mbarrier staged version of the tile kernel was measured too: 0.26 ms at 16384^2 against
//  0.16 ms for the plain tile kernel -- the bit squeeze is issue-bound at the occupancy its staging buffers allow.)
// ------------------------------------------------------------------------------------------------
#define INF_RPT 64
#define INF_U 8

template <int R, int STEP>
__global__ void __launch_bounds__(256, 3) k_inflate_roll(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int W, int H,
                                                         int strips_y, int ntasks, int rpt)
{
    const int lane = threadIdx.x & 31;
    const int task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (task >= ntasks) return;
    const int y0 = (task % strips_y) * INF_TY, x0 = (task / strips_y) * rpt;
    const int yl = y0 + 16 * lane;
    const bool in_y = yl < H;
    // halo: the 4 cells next to the strip (R <= 4), one 4-byte load on lane 0 (left) / lane 31 (right)
    const int ye = lane == 0 ? y0 - 4 : y0 + INF_TY;
    const bool has_e = (lane == 0 || lane == 31) && ye >= 0 && ye < H;
    const int xo_end = min(x0 + rpt, W);  // output rows [x0, xo_end)
    const int xr_end = xo_end + R;            // input rows fed: [x0 - R, xr_end)
    unsigned ring[2 * R + 1];
#pragma unroll
    for (int i = 0; i < 2 * R + 1; i++) ring[i] = 0;

    for (int xr = x0 - R; xr < xr_end; xr += INF_U) {
        uint4 v[INF_U];
        unsigned e[INF_U];
#pragma unroll
        for (int u = 0; u < INF_U; u++) {
            const int x = xr + u;
            v[u] = make_uint4(0, 0, 0, 0);
            e[u] = 0;
            if (x >= 0 && x < W && x < xr_end) {
                const uint8_t *row = in + (size_t)x * H;
                if (in_y) v[u] = __ldcs(reinterpret_cast<const uint4 *>(row + yl));
                if (has_e) e[u] = __ldg(reinterpret_cast<const unsigned *>(row + ye));
            }
        }
#pragma unroll
        for (int u = 0; u < INF_U; u++) {
            const int x = xr + u;
            if (x >= xr_end) break;  // warp-uniform
            const unsigned b = bytes_to_bits16(v[u]);
            const unsigned eb = ((__vcmpne4(e[u], 0u) & 0x01010101u) * 0x01020408u >> 24) & 0xFu;  // 4 halo cells -> 4 bits
            unsigned prev = __shfl_up_sync(0xFFFFFFFFu, b, 1), next = __shfl_down_sync(0xFFFFFFFFu, b, 1);
            if (lane == 0) prev = eb << 12;   // cells y0-4 .. y0-1 are bits 12..15 of the chunk before
            if (lane == 31) next = eb;        // cells y0+512 .. y0+515 are bits 0..3 of the chunk after
            const unsigned long long win = (unsigned long long)prev | ((unsigned long long)b << 16) | ((unsigned long long)next << 32);
            unsigned acc = 0;
#pragma unroll
            for (int q = -R; q <= R; q += STEP) acc |= (unsigned)(win >> (16 + q));
#pragma unroll
            for (int i = 0; i < 2 * R; i++) ring[i] = ring[i + 1];
            ring[2 * R] = acc & 0xFFFFu;
            const int xo = x - R;
            if (xo >= x0 && xo < xo_end) {
                unsigned o = 0;
#pragma unroll
                for (int i = 0; i <= 2 * R; i += STEP) o |= ring[i];
                if (in_y) __stcs(reinterpret_cast<uint4 *>(out + (size_t)xo * H + yl), bits16_to_bytes(o));
            }
        }
    }
}

template <int R, int STEP>
static void launch_roll(const uint8_t *in, uint8_t *out, int W, int H, int sm_count, cudaStream_t st)
{
    const int strips_y = (H + INF_TY - 1) / INF_TY;
    // rows per warp: 64 on grids that fill the machine anyway (halo re-reads 2R/64); shorter strips on small grids,
    // where a warp walking 64 rows one after the other is the whole run time (1024^2: 32 warps, 20 us -> 256 warps)
    int rpt = INF_RPT;
    while (rpt > 8 && (long long)strips_y * ((W + rpt - 1) / rpt) < 8LL * sm_count) rpt >>= 1;
    const long long ntasks = (long long)strips_y * ((W + rpt - 1) / rpt);
    k_inflate_roll<R, STEP><<<(unsigned)((ntasks + 7) / 8), 256, 0, st>>>(in, out, W, H, strips_y, (int)ntasks, rpt);
}

// any shape / alignment / radius: one thread per output cell, reads the stencil directly
__global__ void __launch_bounds__(256) k_inflate_generic(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                                                         int W, int H, int r, int step)
{
    const size_t total = (size_t)W * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i / H), y = (int)(i - (size_t)x * H);
        unsigned hit = 0;
        for (int a = -r; a <= r && !hit; a += step) {
            const int xx = x + a;
            if (xx < 0 || xx >= W) continue;
            for (int b = -r; b <= r; b += step) {
                const int yy = y + b;
                if (yy < 0 || yy >= H) continue;
                if (__ldg(in + (size_t)xx * H + yy)) { hit = 1; break; }
            }
        }
        out[i] = (uint8_t)hit;
    }
}

extern "C" int fx_inflate(fx_context *ctx, const uint8_t *in, uint8_t *out, int W, int H, int radius, int step, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!in || !out || in == out || W <= 0 || H <= 0 || radius < 0 || step < 0)
        return fx_set_err(ctx, FX_ERR_ARG, "fx_inflate: bad argument");
    if (step == 0 || radius == 0) step = 1;
    if (radius % step != 0) return fx_set_err(ctx, FX_ERR_ARG, "fx_inflate: step must divide radius");
    cudaStream_t st = (cudaStream_t)stream;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool fast = (H % 16 == 0) && (((uintptr_t)in | (uintptr_t)out) & 15u) == 0 && radius <= INF_MAXR;
    // the rolling-window kernels exist for the two reference stencils only (step == 1 dense, step == radius 9-point);
    // any other divisor of the radius (e.g. radius 4, step 2) goes to the tile / generic kernels, which honour `step`
    if (fast && radius >= 1 && radius <= 4 && (step == 1 || step == radius)) {
        const bool dense = step == 1;
        switch (radius * 2 + (dense ? 1 : 0)) {
        case 2: case 3: launch_roll<1, 1>(in, out, W, H, ctx->sm_count, st); break;   // r = 1: the 9-point stencil IS the dense 3x3
        case 5: launch_roll<2, 1>(in, out, W, H, ctx->sm_count, st); break;
        case 4: launch_roll<2, 2>(in, out, W, H, ctx->sm_count, st); break;
        case 7: launch_roll<3, 1>(in, out, W, H, ctx->sm_count, st); break;
        case 6: launch_roll<3, 3>(in, out, W, H, ctx->sm_count, st); break;
        case 9: launch_roll<4, 1>(in, out, W, H, ctx->sm_count, st); break;
        default: launch_roll<4, 4>(in, out, W, H, ctx->sm_count, st); break;
        }
    } else if (fast) {
        // larger radii: shared-memory tile kernel; small grids get shorter tiles so that the grid still covers the SMs
        long long tiles64 = (long long)((W + 63) / 64) * ((H + INF_TY - 1) / INF_TY);
        int TX = tiles64 >= 2LL * ctx->sm_count ? 64 : 16;
        dim3 g((H + INF_TY - 1) / INF_TY, (W + TX - 1) / TX);
        k_inflate_tiled<<<g, 256, 0, st>>>(in, out, W, H, radius, step, TX);
    } else {
        size_t total = (size_t)W * H;
        int blocks = (int)((total + 255) / 256);
        if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
        k_inflate_generic<<<blocks, 256, 0, st>>>(in, out, W, H, radius, step);
    }
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}
