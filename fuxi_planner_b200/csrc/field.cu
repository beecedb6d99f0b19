// field.cu -- whole-GPU cost-from-source field (sm_100a): the full-field form of the wavefront in
// search.cu, used for tests, for many-goals-one-source planning and for the row-tiled multi-GPU mode.
//
// Same Dial / delta-stepping wavefront (bucket width == straight weight, four rotating bucket queues,
// warp-ballot compaction) but spread over one persistent cooperative grid: every level is split over
// all CTAs and closed by one grid-wide barrier.  Seeds (cost, cell) are injected when their bucket
// comes up, which is what the row-tiled mode needs: after a halo exchange the improved ghost-row cells
// are the seeds and the slab is relaxed to its fixpoint in bucket order (fx_field_relax).
#include <cooperative_groups.h>

#include "common.cuh"
namespace cg = cooperative_groups;

template <int METRIC> struct FWt;
template <> struct FWt<1> { static constexpr uint32_t WS = 10, WD = 14; };
template <> struct FWt<2> { static constexpr uint32_t WS = FX_EUCLID_WS, WD = FX_EUCLID_WD; };

// device state block (ctx->fstate), 32 unsigned ints
enum { ST_TAIL0 = 0, ST_NSEEDS = 5, ST_MINB = 6, ST_MAXB = 7, ST_FLAGS = 8 /* two parity slots: 8, 9 */, ST_LEVELS = 10,
       ST_SETTLED = 11, ST_WORDS = 32 };
#define FFLAG_OVERFLOW 1u
#define FFLAG_SEEDCAP 2u
#define FFLAG_HISTCAP 4u

struct FieldParams {
    const uint8_t *moves;
    int W, H;
    uint32_t *field;
    uint32_t *fq;
    unsigned fqcap;
    unsigned *st;
    const uint2 *seeds;  // grouped by bucket, (cost, packed xy)
    const unsigned *hist;
    int relax;           // 1: prior finite values may exist -> push on every improvement
    int32_t *changed;
};

template <int METRIC>
__global__ void __launch_bounds__(512) k_field_dial(const FieldParams P)
{
    constexpr uint32_t WS = FWt<METRIC>::WS, WD = FWt<METRIC>::WD;
    cg::grid_group grid = cg::this_grid();
    const int H = P.H;
    const unsigned lane = threadIdx.x & 31;
    const unsigned T = gridDim.x * blockDim.x;
    const unsigned gt = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned *st = P.st;
    const unsigned nseeds = __ldcg(st + ST_NSEEDS);
    const unsigned minb = __ldcg(st + ST_MINB), maxb = __ldcg(st + ST_MAXB);
    const unsigned fqcap = P.fqcap;
    const unsigned *__restrict__ hist = P.hist;  // hist[b - minb] = end offset of bucket b in seeds[]
    unsigned k = minb;
    unsigned levels = 0, my_settled = 0;
    bool any_change = false;
    if (nseeds == 0) return;  // uniform: nothing to do
    for (;;) {
        // everything read here was last written before the preceding grid barrier
        const unsigned n = __ldcg(st + ST_TAIL0 + (k & 3)), n1 = __ldcg(st + ST_TAIL0 + ((k + 1) & 3));
        const unsigned flags = __ldcg(st + ST_FLAGS + (k & 1));  // level k-1 wrote slot k&1, level k writes (k+1)&1
        unsigned s0 = 0, s1 = 0;
        if (k >= minb && k <= maxb) { s0 = k > minb ? __ldcg(hist + (k - minb - 1)) : 0u; s1 = __ldcg(hist + (k - minb)); }
        if (flags & FFLAG_OVERFLOW) break;
        if (n > fqcap || n1 > fqcap) { if (gt == 0) { st[ST_FLAGS] |= FFLAG_OVERFLOW; st[ST_FLAGS + 1] |= FFLAG_OVERFLOW; } break; }
        if (n == 0 && n1 == 0 && s1 == s0) {
            // queues drained: jump to the next bucket that holds seeds, or stop at the fixpoint
            const unsigned nxt = k < minb ? 0u : (k > maxb ? nseeds : s1);
            if (nxt >= nseeds) break;
            const unsigned kb = __ldcg(&P.seeds[nxt].x) / WS;
            if (gt == 0) st[ST_TAIL0 + ((k + 3) & 3)] = 0;  // the only slot that may still hold a stale count
            grid.sync();
            k = kb;
            continue;
        }
        if (gt == 0) st[ST_TAIL0 + ((k + 3) & 3)] = 0;
        const uint32_t *__restrict__ qk = P.fq + (size_t)(k & 3) * fqcap;
        uint32_t *__restrict__ q1 = P.fq + (size_t)((k + 1) & 3) * fqcap;
        uint32_t *__restrict__ q2 = P.fq + (size_t)((k + 2) & 3) * fqcap;
        const unsigned total_q = n * 8u;
        const unsigned total = total_q + (s1 - s0) * 8u;
        for (unsigned i0 = gt - lane; i0 < total; i0 += T) {
            const unsigned i = i0 + lane;
            bool act = i < total;
            const int d = (int)(i & 7u);
            uint32_t xy = 0;
            if (act) xy = i < total_q ? __ldcg(qk + (i >> 3)) : __ldcg(&P.seeds[s0 + ((i - total_q) >> 3)].y);
            // all lanes of a warp stay in the loop for the ballots: no early continue
            const int x = (int)(xy >> 16), y = (int)(xy & 0xFFFFu);
            const size_t idx = (size_t)x * H + y;
            const uint32_t g = act ? __ldcg(P.field + idx) : FX_INF;
            act = act && g != FX_INF && (g / WS) == k;  // stale: the cell has moved to an earlier bucket
            if (act && d == 0) my_settled++;
            const unsigned m = act ? (unsigned)__ldg(P.moves + idx) : 0u;
            act = act && ((m >> d) & 1u);
            const int ddx = fx_dx(d), ddy = fx_dy(d);
            const int nx = x + ddx, ny = y + ddy;
            const uint32_t ng = g + (d < 4 ? WS : WD);
            if (act && ng > 0x7FFFFFFFu) { act = false; atomicOr(st + ST_FLAGS + ((k + 1) & 1), FFLAG_OVERFLOW); }
            const size_t nidx = (size_t)((long long)idx + (long long)ddx * H + ddy);
            uint32_t old = 0;
            if (act) act = ng < __ldcg(P.field + nidx);
            if (act) { old = atomicMin(P.field + nidx, ng); act = ng < old; }
            if (act) any_change = true;
            const unsigned nb = ng / WS;
            const bool push = act && (P.relax || old == FX_INF || old / WS != nb);
            const bool p1 = push && nb == k + 1, p2 = push && nb != k + 1;
            const unsigned m1 = __ballot_sync(0xFFFFFFFFu, p1), m2 = __ballot_sync(0xFFFFFFFFu, p2);
            if (m1) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(st + ST_TAIL0 + ((k + 1) & 3), __popc(m1));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (p1) {
                    unsigned pos = base + __popc(m1 & ((1u << lane) - 1u));
                    if (pos < fqcap) __stcg(q1 + pos, ((uint32_t)nx << 16) | (uint32_t)ny);
                }
            }
            if (m2) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(st + ST_TAIL0 + ((k + 2) & 3), __popc(m2));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (p2) {
                    unsigned pos = base + __popc(m2 & ((1u << lane) - 1u));
                    if (pos < fqcap) __stcg(q2 + pos, ((uint32_t)nx << 16) | (uint32_t)ny);
                }
            }
        }
        grid.sync();
        k++; levels++;
    }
    if (any_change && P.changed) *P.changed = 1;
    if (my_settled) atomicAdd(st + ST_SETTLED, my_settled);
    if (gt == 0) st[ST_LEVELS] = levels;
}

// ---- seed detection for relax: every finite cell that can still improve a neighbour ---------------
template <int METRIC>
__global__ void __launch_bounds__(256) k_find_active(const uint8_t *__restrict__ moves, int W, int H,
                                                     const uint32_t *__restrict__ field, uint2 *__restrict__ seeds,
                                                     unsigned seed_cap, unsigned *st)
{
    constexpr uint32_t WS = FWt<METRIC>::WS, WD = FWt<METRIC>::WD;
    const size_t total = (size_t)W * H;
    const unsigned lane = threadIdx.x & 31;
    const size_t T = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < total; i0 += T) {
        const size_t i = i0 + lane;
        bool active = false;
        uint32_t g = FX_INF;
        int x = 0, y = 0;
        if (i < total) {
            g = field[i];
            if (g != FX_INF) {
                x = (int)(i / H); y = (int)(i - (size_t)x * H);
                const unsigned m = moves[i];
#pragma unroll
                for (int d = 0; d < 8; d++) {
                    if (!((m >> d) & 1u)) continue;
                    const size_t n = (size_t)((long long)i + (long long)fx_dx(d) * H + fx_dy(d));
                    if (g + (d < 4 ? WS : WD) < field[n]) active = true;
                }
            }
        }
        const unsigned mask = __ballot_sync(0xFFFFFFFFu, active);
        if (mask) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(st + ST_NSEEDS, __popc(mask));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (active) {
                unsigned pos = base + __popc(mask & ((1u << lane) - 1u));
                if (pos < seed_cap) seeds[pos] = make_uint2(g, ((uint32_t)x << 16) | (uint32_t)y);
                else atomicOr(st + ST_FLAGS, FFLAG_SEEDCAP);
                atomicMin(st + ST_MINB, g / WS);
                atomicMax(st + ST_MAXB, g / WS);
            }
        }
    }
}

// counting sort of the seeds by bucket (order inside a bucket is irrelevant)
template <int METRIC>
__global__ void k_seed_hist(const uint2 *__restrict__ seeds, unsigned seed_cap, unsigned *st, unsigned *hist, unsigned hist_cap)
{
    constexpr uint32_t WS = FWt<METRIC>::WS;
    unsigned n = min(st[ST_NSEEDS], seed_cap);
    const unsigned minb = st[ST_MINB];
    if (n && st[ST_MAXB] - minb + 1 > hist_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(st + ST_FLAGS, FFLAG_HISTCAP); return; }
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(hist + (seeds[i].x / WS - minb), 1u);
}
__global__ void __launch_bounds__(1024) k_seed_scan(unsigned *st, unsigned *hist, unsigned hist_cap)
{
    // single CTA exclusive scan over [0, range)
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry;
    if (st[ST_NSEEDS] == 0 || (st[ST_FLAGS] & FFLAG_HISTCAP)) return;
    const unsigned range = st[ST_MAXB] - st[ST_MINB] + 1;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < range; base += 1024) {
        unsigned i = base + threadIdx.x;
        unsigned v = i < range ? hist[i] : 0u, s = v;
        for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xFFFFFFFFu, s, o); if ((threadIdx.x & 31) >= o) s += t; }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned ws = warp_sums[threadIdx.x], t2 = ws;
            for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xFFFFFFFFu, t2, o); if (threadIdx.x >= o) t2 += t; }
            warp_sums[threadIdx.x] = t2 - ws;  // exclusive
        }
        __syncthreads();
        unsigned excl = carry + warp_sums[threadIdx.x >> 5] + s - v;
        if (i < range) hist[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}
template <int METRIC>
__global__ void k_seed_scatter(const uint2 *__restrict__ seeds, uint2 *__restrict__ sorted, unsigned seed_cap, unsigned *st, unsigned *hist)
{
    constexpr uint32_t WS = FWt<METRIC>::WS;
    if (st[ST_FLAGS] & FFLAG_HISTCAP) return;
    unsigned n = min(st[ST_NSEEDS], seed_cap);
    const unsigned minb = st[ST_MINB];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint2 s = seeds[i];
        unsigned pos = atomicAdd(hist + (s.x / WS - minb), 1u);
        sorted[pos] = s;
    }
}
__global__ void k_seed_finish(unsigned *st, unsigned seed_cap)
{
    if (st[ST_NSEEDS] > seed_cap) st[ST_NSEEDS] = seed_cap;
    if (st[ST_FLAGS] & (FFLAG_HISTCAP | FFLAG_SEEDCAP)) { st[ST_NSEEDS] = 0; st[ST_FLAGS] |= FFLAG_OVERFLOW; }
    st[ST_FLAGS + 1] = st[ST_FLAGS];
}

__global__ void k_field_init(uint32_t *field, size_t cells, size_t src, uint2 *seeds, uint32_t xy, unsigned *st, unsigned *hist)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t T = (size_t)gridDim.x * blockDim.x;
    for (; i < cells; i += T) field[i] = i == src ? 0u : FX_INF;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        seeds[0] = make_uint2(0u, xy); st[ST_NSEEDS] = 1; st[ST_MINB] = 0; st[ST_MAXB] = 0; hist[0] = 1;
    }
}

static int field_reserve(fx_context *ctx, int W, int H, bool relax)
{
    size_t cells = (size_t)W * H;
    size_t want = 64 * (size_t)(W + H) + (1u << 20);
    if (want > cells + 1024) want = cells + 1024;
    if (ctx->fqcap < want) {
        if (ctx->fq) cudaFree(ctx->fq);
        ctx->fq = nullptr; ctx->fqcap = 0;
        FX_CUDA(ctx, cudaMalloc(&ctx->fq, want * 4 * sizeof(uint32_t)));
        ctx->fqcap = want;
    }
    size_t seed_want = relax ? (cells < (8u << 20) ? cells : (8u << 20)) : 16;
    if (ctx->seed_cap < seed_want) {
        if (ctx->seeds) cudaFree(ctx->seeds);
        if (ctx->seeds_sorted) cudaFree(ctx->seeds_sorted);
        ctx->seeds = ctx->seeds_sorted = nullptr; ctx->seed_cap = 0;
        FX_CUDA(ctx, cudaMalloc(&ctx->seeds, seed_want * sizeof(uint2)));
        FX_CUDA(ctx, cudaMalloc(&ctx->seeds_sorted, seed_want * sizeof(uint2)));
        ctx->seed_cap = seed_want;
    }
    size_t hist_want = relax ? (4u << 20) : 16;
    if (ctx->seed_hist_cap < hist_want) {
        if (ctx->seed_hist) cudaFree(ctx->seed_hist);
        ctx->seed_hist = nullptr; ctx->seed_hist_cap = 0;
        FX_CUDA(ctx, cudaMalloc(&ctx->seed_hist, hist_want * sizeof(unsigned)));
        ctx->seed_hist_cap = hist_want;
    }
    return FX_OK;
}

template <int METRIC>
static int launch_dial(fx_context *ctx, FieldParams &P, cudaStream_t st)
{
    int per_sm = 0;
    FX_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_field_dial<METRIC>, 512, 0));
    if (per_sm < 1) return fx_set_err(ctx, FX_ERR_CUDA, "k_field_dial cannot be resident");
    int blocks = ctx->sm_count;  // one fat CTA per SM keeps the grid barrier cheap
    void *args[] = {(void *)&P};
    FX_CUDA(ctx, cudaLaunchCooperativeKernel((void *)k_field_dial<METRIC>, dim3(blocks), dim3(512), args, 0, st));
    ctx->launches++;
    return FX_OK;
}

static int field_common(fx_context *ctx, const uint8_t *grid, int W, int H, int metric, int32_t *field,
                        int32_t *d_changed, int sx, int sy, bool relax, cudaStream_t st)
{
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = field_reserve(ctx, W, H, relax);
    if (rc) return rc;
    rc = fx_build_moves(ctx, grid, W, H, false, st);
    if (rc) return rc;
    const size_t cells = (size_t)W * H;
    FX_CUDA(ctx, cudaMemsetAsync(ctx->fstate, 0, ST_WORDS * sizeof(unsigned), st));
    FX_CUDA(ctx, cudaMemsetAsync(ctx->fstate + ST_MINB, 0xFF, sizeof(unsigned), st));
    FieldParams P;
    P.moves = ctx->moves; P.W = W; P.H = H; P.field = reinterpret_cast<uint32_t *>(field);
    P.fq = ctx->fq; P.fqcap = (unsigned)ctx->fqcap; P.st = ctx->fstate; P.relax = relax ? 1 : 0; P.changed = d_changed;
    if (!relax) {
        int blocks = (int)((cells + 255) / 256);
        if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
        k_field_init<<<blocks, 256, 0, st>>>(P.field, cells, (size_t)sx * H + sy, ctx->seeds_sorted,
                                             ((uint32_t)sx << 16) | (uint32_t)sy, ctx->fstate, ctx->seed_hist);
        FX_LAUNCH_CHECK(ctx);
    } else {
        int blocks = ctx->sm_count * 8;
        FX_CUDA(ctx, cudaMemsetAsync(ctx->seed_hist, 0, ctx->seed_hist_cap * sizeof(unsigned), st));
        if (metric == 1) k_find_active<1><<<blocks, 256, 0, st>>>(ctx->moves, W, H, P.field, ctx->seeds, (unsigned)ctx->seed_cap, ctx->fstate);
        else k_find_active<2><<<blocks, 256, 0, st>>>(ctx->moves, W, H, P.field, ctx->seeds, (unsigned)ctx->seed_cap, ctx->fstate);
        FX_LAUNCH_CHECK(ctx);
        if (metric == 1) k_seed_hist<1><<<ctx->sm_count, 256, 0, st>>>(ctx->seeds, (unsigned)ctx->seed_cap, ctx->fstate, ctx->seed_hist, (unsigned)ctx->seed_hist_cap);
        else k_seed_hist<2><<<ctx->sm_count, 256, 0, st>>>(ctx->seeds, (unsigned)ctx->seed_cap, ctx->fstate, ctx->seed_hist, (unsigned)ctx->seed_hist_cap);
        FX_LAUNCH_CHECK(ctx);
        k_seed_scan<<<1, 1024, 0, st>>>(ctx->fstate, ctx->seed_hist, (unsigned)ctx->seed_hist_cap);
        FX_LAUNCH_CHECK(ctx);
        if (metric == 1) k_seed_scatter<1><<<ctx->sm_count, 256, 0, st>>>(ctx->seeds, ctx->seeds_sorted, (unsigned)ctx->seed_cap, ctx->fstate, ctx->seed_hist);
        else k_seed_scatter<2><<<ctx->sm_count, 256, 0, st>>>(ctx->seeds, ctx->seeds_sorted, (unsigned)ctx->seed_cap, ctx->fstate, ctx->seed_hist);
        FX_LAUNCH_CHECK(ctx);
        k_seed_finish<<<1, 1, 0, st>>>(ctx->fstate, (unsigned)ctx->seed_cap);
        FX_LAUNCH_CHECK(ctx);
    }
    P.seeds = ctx->seeds_sorted; P.hist = ctx->seed_hist;
    return metric == 1 ? launch_dial<1>(ctx, P, st) : launch_dial<2>(ctx, P, st);
}

extern "C" int fx_field(fx_context *ctx, const uint8_t *grid, int W, int H, int sx, int sy, int metric,
                        int32_t *field, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!grid || !field || W <= 0 || H <= 0 || (metric != 1 && metric != 2) || sx < 0 || sx >= W || sy < 0 || sy >= H)
        return fx_set_err(ctx, FX_ERR_ARG, "fx_field: bad argument");
    if (W > 32767 || H > 32767) return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_field: W,H must be <= 32767");
    return field_common(ctx, grid, W, H, metric, field, nullptr, sx, sy, false, (cudaStream_t)stream);
}

extern "C" int fx_field_relax(fx_context *ctx, const uint8_t *grid, int Wloc, int H, int metric, int32_t *field,
                              int32_t *d_changed, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!grid || !field || Wloc <= 0 || H <= 0 || (metric != 1 && metric != 2))
        return fx_set_err(ctx, FX_ERR_ARG, "fx_field_relax: bad argument");
    if (Wloc > 32767 || H > 32767) return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_field_relax: W,H must be <= 32767");
    return field_common(ctx, grid, Wloc, H, metric, field, d_changed, 0, 0, true, (cudaStream_t)stream);
}

// halo merge for the row-tiled mode: dst[i] = min(dst[i], src[i]) with -1 (unknown) as +inf; *d_changed |= improved
__global__ void __launch_bounds__(256) k_halo_merge(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, size_t n,
                                                    int32_t *changed)
{
    bool any = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t a = dst[i], b = src[i];
        if (b < a) { dst[i] = b; any = true; }
    }
    if (__any_sync(0xFFFFFFFFu, any) && (threadIdx.x & 31) == 0 && changed) *changed = 1;
}

extern "C" int fx_halo_merge(fx_context *ctx, int32_t *dst, const int32_t *src, int64_t n, int32_t *d_changed, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (n < 0 || (n > 0 && (!dst || !src))) return fx_set_err(ctx, FX_ERR_ARG, "fx_halo_merge: bad argument");
    if (n == 0) return FX_OK;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    long long blocks = (n + 255) / 256;
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_halo_merge<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint32_t *>(dst),
                                                                 reinterpret_cast<const uint32_t *>(src), (size_t)n, d_changed);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}

extern "C" int fx_field_status(fx_context *ctx, int64_t *h_levels, int64_t *h_settled)
{
    if (!ctx) return FX_ERR_ARG;
    unsigned st[ST_WORDS];
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    FX_CUDA(ctx, cudaDeviceSynchronize());
    FX_CUDA(ctx, cudaMemcpy(st, ctx->fstate, sizeof(st), cudaMemcpyDeviceToHost));
    if (h_levels) *h_levels = st[ST_LEVELS];
    if (h_settled) *h_settled = st[ST_SETTLED];
    if ((st[ST_FLAGS] | st[ST_FLAGS + 1]) & FFLAG_OVERFLOW)
        return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "field wavefront overflowed (queue, seed list or 31-bit cost range): flags=%u", st[ST_FLAGS] | st[ST_FLAGS + 1]);
    return FX_OK;
}
