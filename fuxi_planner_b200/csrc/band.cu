// band.cu -- per-query upper bounds for the batched search (sm_100a), one WARP per query.
//
// The exact pass of search.cu prunes with the ellipse g + octile(c, goal) <= U, so it needs an upper bound U on
// the cost of each query.  A wavefront restricted to a +-15-cell band around the start-goal line finds one: any
// path inside the band is a real path.  Its frontier is ~30 cells wide -- one warp's worth -- and a level costs one
// chain of dependent L2 round trips whatever the width, so giving it a whole CTA (as the first version did) left
// three quarters of the warps waiting at the block barrier for ~half of all levels (ncu r01: 45 % barrier stalls).
// Here every warp runs its own query: no block barrier, bucket queues in shared memory, and a cost field in *band
// coordinates* (t = position along the major axis, v = offset from the line, 32 cells = one 128-byte line per t)
// so that a slot needs (max(W,H) + 2*margin) * 128 bytes instead of a full W x H field and is reset with
// streaming stores.
//
// Also here: the longest-processing-time-first query order (k_order_queries).  Query work spans four orders of
// magnitude (it grows with the area of the start-goal parallelogram); handing out the big ones first keeps the
// tail of the batch short.
#include "common.cuh"

#define BAND_B 16                 /* |offset from the line| <= BAND_B - 1 */
#define BAND_M 24                 /* rows allowed before the start / after the goal along the major axis */
#define BAND_Q 128                /* queue entries (8 bytes) per bucket and warp (shared memory) */
#define BAND_WARPS 8
#ifndef BAND_MINB
#define BAND_MINB 6                 /* resident CTAs per SM the register budget is compiled for (40 registers; 5: 44, 8: 32 + spills, measured slower) */
#endif

struct BandParams {
    const uint8_t *grid, *moves;
    int W, H, TY;
    const int32_t *starts, *goals;
    int Q;
    const uint32_t *order;
    uint32_t *ubound;
    uint32_t *bfields;
    size_t bcap;
    unsigned long long *counter;
};

// ---- LPT order: 64 log-scale buckets of estimated work, largest first ---------------------------------------
__device__ __forceinline__ unsigned work_bucket(int sx, int sy, int gx, int gy)
{
    const unsigned dx = (unsigned)abs(gx - sx), dy = (unsigned)abs(gy - sy);
    const unsigned mn = min(dx, dy), mx = max(dx, dy);
    unsigned long long est = (unsigned long long)mn * (mx - mn) + 64ull * mx + 1ull;  // parallelogram area + band
    if (est > 0x7FFFFFFFull) est = 0x7FFFFFFFull;
    const unsigned e = (unsigned)est, lg = 31u - (unsigned)__clz(e);
    return 2u * lg + ((lg > 0) ? ((e >> (lg - 1)) & 1u) : 0u);  // 0..61
}

__global__ void __launch_bounds__(1024) k_order_queries(const int32_t *__restrict__ starts, const int32_t *__restrict__ goals, int Q,
                                                        uint32_t *__restrict__ order)
{
    __shared__ unsigned hist[64], offs[64];
    if (threadIdx.x < 64) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int q = threadIdx.x; q < Q; q += blockDim.x)
        atomicAdd(&hist[work_bucket(starts[2 * q], starts[2 * q + 1], goals[2 * q], goals[2 * q + 1])], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned acc = 0;
        for (int b = 63; b >= 0; b--) { offs[b] = acc; acc += hist[b]; }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
        const unsigned b = work_bucket(starts[2 * q], starts[2 * q + 1], goals[2 * q], goals[2 * q + 1]);
        order[atomicAdd(&offs[b], 1u)] = (uint32_t)q;
    }
}

// ---- band pass ------------------------------------------------------------------------------------------------
// Same relaxation scheme as search.cu (canonical successors from the shared LUT, packed cost|direction words,
// fire-and-forget RED.MIN, entries validated when popped), run by ONE WARP per query on a field in band coordinates.
__device__ __forceinline__ void band_red_min(uint32_t *p, uint32_t v)
{
    asm volatile("red.relaxed.gpu.global.min.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned band_atoms_add(unsigned *p, unsigned v)
{
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
    return old;
}

template <int METRIC>
__global__ void __launch_bounds__(BAND_WARPS * 32, BAND_MINB) k_band_bound(const BandParams P)
{
    constexpr uint32_t WS = Wt<METRIC>::WS, WD = Wt<METRIC>::WD;
    __shared__ uint2 s_queue[BAND_WARPS][3][BAND_Q];
    __shared__ unsigned s_cnt[BAND_WARPS][4];
    __shared__ uint8_t s_lut[9 * 256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int H = P.H, W = P.W;
    uint32_t *__restrict__ field = P.bfields + (size_t)(blockIdx.x * BAND_WARPS + warp) * P.bcap;
    const uint8_t *__restrict__ moves = P.moves;
    uint2(*queue)[BAND_Q] = s_queue[warp];
    unsigned *cnt = s_cnt[warp];
    for (int i = threadIdx.x; i < 9 * 256; i += blockDim.x) s_lut[i] = (uint8_t)fx_canon_succ((unsigned)(i >> 8), (unsigned)(i & 255));
    __syncthreads();

    for (;;) {
        unsigned long long it = 0;
        if (lane == 0) it = atomicAdd(P.counter, 1ull);
        it = __shfl_sync(0xFFFFFFFFu, it, 0);
        if (it >= (unsigned long long)P.Q) break;
        const int q = P.order ? (int)P.order[it] : (int)it;
        const int sx = P.starts[2 * q], sy = P.starts[2 * q + 1], gx = P.goals[2 * q], gy = P.goals[2 * q + 1];
        uint32_t result = FX_INF;
        const bool ok = sx >= 0 && sx < W && sy >= 0 && sy < H && gx >= 0 && gx < W && gy >= 0 && gy < H && !(sx == gx && sy == gy);
        if (!ok || P.grid[(size_t)gx * H + gy] == 1 || moves[fx_cidx(sx, sy, H, P.TY)] == 0) {  // search.cu answers these itself
            if (lane == 0) P.ubound[q] = FX_INF;
            continue;
        }
        // band coordinates: a = major axis, b = minor axis
        const int ddx = gx - sx, ddy = gy - sy;
        const bool xmajor = abs(ddx) >= abs(ddy);
        const int L = xmajor ? abs(ddx) : abs(ddy);
        const int sgn = (xmajor ? ddx : ddy) >= 0 ? 1 : -1;
        const int as = xmajor ? sx : sy, bs = xmajor ? sy : sx;
        // minor offset per major step in 2^-15 fixed point: |slope| <= 2^15 and |tt| < 2^15 + 48, so tt * slope fits 32 bits
        // (a third of this kernel's instructions were the 64-bit form of this product: r02b profile)
        const int slope = (int)(((long long)(xmajor ? ddy : ddx) * 32768) / L);
        const int tmax = L + 2 * BAND_M;
        auto off = [&](int tt) { return (tt * slope + 16384) >> 15; };
        // field slot of cell (x, y), or -1 outside the band
        auto slot_of = [&](int x, int y) {
            const int tt = ((xmajor ? x : y) - as) * sgn;
            const int dev = (xmajor ? y : x) - bs - off(tt);
            const bool in = tt + BAND_M >= 0 && tt + BAND_M <= tmax && dev >= -(BAND_B - 1) && dev <= BAND_B - 1;
            return in ? ((tt + BAND_M) << 5) + (dev + BAND_B - 1) : -1;
        };
        const uint32_t h0 = octile(abs(ddx), abs(ddy), WS, WD - WS);
        const uint64_t U64 = (uint64_t)h0 + h0 / 16 + 64 * WS;
        const uint32_t U = U64 > 0x0FFFE000ull ? 0x0FFFE000u : (uint32_t)U64;  // also keeps every child cost inside 28 bits

        if (lane == 0) {
            queue[0][0] = make_uint2(((uint32_t)sx << 16) | (uint32_t)sy, fx_pack(0u, FX_CODE_START));
            __stcg(field + ((size_t)BAND_M << 5) + (BAND_B - 1), fx_pack(0u, FX_CODE_START));  // start: t = BAND_M, deviation 0
            cnt[0] = 1; cnt[1] = 0; cnt[2] = 0;
        }
        __syncwarp();
        int bk = 0;  // buffer of bucket k; k+1 -> (bk+1)%3, k+2 -> (bk+2)%3
        bool overflow = false;
        for (unsigned k = 0;; k++) {
            const int b1 = bk == 2 ? 0 : bk + 1, b2 = b1 == 2 ? 0 : b1 + 1;
            const unsigned n = cnt[bk], n1 = cnt[b1];
            if (n > BAND_Q) { overflow = true; break; }
            if (n == 0 && n1 == 0) break;
            __syncwarp();
            if (lane == 0) cnt[b2] = 0;  // bucket k+2 starts empty (its buffer held bucket k-1)
            __syncwarp();
            const uint32_t kbase = k * WS;
            uint32_t found = FX_INF;
            for (unsigned i0 = 0; i0 < n; i0 += 32) {
                const unsigned i = i0 + lane;
                bool act = i < n;
                const uint2 e = act ? queue[bk][i] : make_uint2(0u, 0u);
                const int x = (int)(e.x >> 16), y = (int)(e.x & 0xFFFFu);
                uint32_t v = FX_INF;
                unsigned m = 0;
                if (act) {
                    v = __ldcg(field + slot_of(x, y));  // entries are only created for in-band cells
                    m = (unsigned)__ldg(moves + fx_cidx(x, y, H, P.TY));
                }
                const uint32_t g = e.y >> 4;
                act = act && v == e.y;  // this entry's relaxation won and nothing improved the cell since
                if (act && x == gx && y == gy) found = g;
                if (act) act = ((uint64_t)g + octile(abs(x - gx), abs(y - gy), WS, WD - WS)) <= (uint64_t)U;
                unsigned succ = act ? (unsigned)s_lut[((e.y & 15u) << 8) | m] : 0u;
                const bool diag2 = (g - kbase) + WD >= 2u * WS;
                while (succ) {
                    const int d = __ffs(succ) - 1;
                    succ &= succ - 1;
                    const int nx = x + fx_dx(d), ny = y + fx_dy(d);
                    const int ns = slot_of(nx, ny);
                    if (ns < 0) continue;
                    const uint32_t nv = fx_pack(g + (d < 4 ? WS : WD), (unsigned)d);
                    band_red_min(field + ns, nv);
                    const int nb = (d < 4 || !diag2) ? b1 : b2;
                    const unsigned pos = band_atoms_add(&cnt[nb], 1u);
                    if (pos < BAND_Q) queue[nb][pos] = make_uint2(((uint32_t)nx << 16) | (uint32_t)ny, nv);
                }
            }
            found = __reduce_min_sync(0xFFFFFFFFu, found);
            if (found != FX_INF) { result = found; break; }  // the goal was popped: its cost inside the band is final
            __syncwarp();
            bk = b1;
        }
        if (lane == 0) P.ubound[q] = overflow ? FX_INF : result;
        // reset the rows this query could touch (one 128-byte line per t)
        {
            __syncwarp();
            uint4 *f4 = reinterpret_cast<uint4 *>(field);
            const uint4 inf4 = make_uint4(FX_INF, FX_INF, FX_INF, FX_INF);
            const size_t n16 = ((size_t)tmax + 1) * 8;
            for (size_t i = lane; i < n16; i += 32) __stcg(f4 + i, inf4);
        }
        __syncwarp();
    }
}

int fx_band_bounds(fx_context *ctx, const uint8_t *grid, int W, int H, const int32_t *starts_xy, const int32_t *goals_xy, int Q,
                   int metric, cudaStream_t st)
{
    if (ctx->q_cap < (size_t)Q) {
        if (ctx->q_order) cudaFree(ctx->q_order);
        if (ctx->q_ubound) cudaFree(ctx->q_ubound);
        ctx->q_order = ctx->q_ubound = nullptr; ctx->q_cap = 0;
        FX_CUDA(ctx, cudaMalloc(&ctx->q_order, (size_t)Q * sizeof(uint32_t)));
        FX_CUDA(ctx, cudaMalloc(&ctx->q_ubound, (size_t)Q * sizeof(uint32_t)));
        ctx->q_cap = (size_t)Q;
    }
    const size_t bcap = ((size_t)(W > H ? W : H) + 2 * BAND_M + 1) * 32;
    const int bslots = ctx->sm_count * BAND_MINB * BAND_WARPS;
    if (ctx->bcap < bcap || ctx->bslots < bslots) {
        if (ctx->bfields) cudaFree(ctx->bfields);
        ctx->bfields = nullptr; ctx->bcap = 0; ctx->bslots = 0;
        FX_CUDA(ctx, cudaMalloc(&ctx->bfields, (size_t)bslots * bcap * sizeof(uint32_t)));
        FX_CUDA(ctx, cudaMemsetAsync(ctx->bfields, 0xFF, (size_t)bslots * bcap * sizeof(uint32_t), st));
        ctx->bcap = bcap; ctx->bslots = bslots;
    }
    k_order_queries<<<1, 1024, 0, st>>>(starts_xy, goals_xy, Q, ctx->q_order);
    FX_LAUNCH_CHECK(ctx);
    BandParams P;
    P.grid = grid; P.moves = ctx->moves; P.W = W; P.H = H; P.TY = fx_tiles_y(H); P.starts = starts_xy; P.goals = goals_xy; P.Q = Q;
    P.order = ctx->q_order; P.ubound = ctx->q_ubound; P.bfields = ctx->bfields; P.bcap = ctx->bcap; P.counter = ctx->counters + 6;
    int blocks = (Q + BAND_WARPS - 1) / BAND_WARPS;
    if (blocks > ctx->sm_count * BAND_MINB) blocks = ctx->sm_count * BAND_MINB;
    if (metric == 1) k_band_bound<1><<<blocks, BAND_WARPS * 32, 0, st>>>(P);
    else k_band_bound<2><<<blocks, BAND_WARPS * 32, 0, st>>>(P);
    FX_LAUNCH_CHECK(ctx);
    FX_CUDA(ctx, cudaEventRecord(ctx->ev_band[1], st));
    return FX_OK;
}
