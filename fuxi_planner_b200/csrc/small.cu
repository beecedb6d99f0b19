// small.cu -- the search for maps that fit in one SM's shared memory (sm_100a).
//
// Every map the reference ships (maps/*.png: 54 x 12 ... 147 x 113 cells) and every planning grid its 80 Hz loops
// build from them is below 20 000 cells.  At that size a replan is pure latency: the batched kernel of search.cu
// spends a chain of dependent L2 round trips per wavefront level (~1.5 us x ~150 levels x 3 passes, plus the band
// kernel).  Here ONE CTA holds everything in shared memory -- the legal-move mask (built in place from the grid, the
// k_build_moves launch disappears), the packed cost field and three rotating bucket queues -- so a level costs a
// shared-memory round trip and one block barrier, and the whole query is one launch.
//
// Same graph (edges = `not blocked(c, d)`, scripts/jps1.py:14-31), same Dial wavefront with bucket width == the
// straight weight, same canonical successors (scripts/jps1.py:49-93 via fx_canon_succ) and the same parent rule
// (minimum of cost << 5 | direction << 1 | unsettled) as search.cu; no band / ellipse passes -- a map this small is
// settled outright, stopping at the level that pops the goal.  A cell enters a bucket's queue at most once (only
// when its tentative cost moves into that bucket), so a queue can never hold more than `cells` entries: no overflow
// path exists.
#include "common.cuh"

#ifndef SM_THREADS
#define SM_THREADS 256
#endif
#define SM_INF 0xFFFFFFFFu

struct SmallParams {
    const uint8_t *grid;
    int W, H;
    const int32_t *starts, *goals;
    int Q;
    int32_t *cost_i;
    double *cost_f;
    int32_t *path_xy;
    int32_t *path_len;
    int max_path;
    unsigned long long *counters;
};

// shared-memory layout for `cells` cells: field u32[cells] | queues u16[3][cells] | moves u8[cells] | lut u8[9*256] |
// occupancy bits u32[cells/32] (the grid itself is read exactly once: it may live in pinned host memory)
__host__ __device__ inline size_t small_smem_bytes(size_t cells)
{
    const size_t c4 = (cells + 3) & ~(size_t)3;
    return c4 * 4 + 3 * c4 * 2 + c4 + 9 * 256 + ((cells + 31) / 32) * 4;
}

template <int METRIC>
__global__ void __launch_bounds__(SM_THREADS) k_search_small(const SmallParams P)
{
    constexpr uint32_t WS = Wt<METRIC>::WS, WD = Wt<METRIC>::WD;
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ unsigned s_cnt[4];
    __shared__ unsigned s_goal[2];  // cost of the goal when popped, one slot per level parity (written in level k, read after its barrier)
    __shared__ int s_npts;
    __shared__ unsigned s_ab[2];
    const int W = P.W, H = P.H, tid = threadIdx.x, lane = tid & 31;
    const int cells = W * H, c4 = (cells + 3) & ~3;
    uint32_t *field = reinterpret_cast<uint32_t *>(s_raw);
    uint16_t *queue = reinterpret_cast<uint16_t *>(s_raw + (size_t)c4 * 4);
    uint8_t *moves = s_raw + (size_t)c4 * 4 + (size_t)c4 * 6;
    uint8_t *lut = moves + c4;
    unsigned *occb = reinterpret_cast<unsigned *>(lut + 9 * 256);
    for (int i = tid; i < 9 * 256; i += SM_THREADS) lut[i] = (uint8_t)fx_canon_succ((unsigned)(i >> 8), (unsigned)(i & 255));
    // legal-move masks (k_build_moves of search.cu, scripts/jps1.py:14-31), from a shared-memory copy of the grid parked in
    // the queue area
    {
        uint8_t *g = reinterpret_cast<uint8_t *>(queue);
        for (int i = tid; i < (cells + 31) / 32; i += SM_THREADS) occb[i] = 0;
        // the grid may live in mapped pinned host memory (fx_plan_host): 16-byte loads, all in flight at once, so that
        // the copy costs one PCIe round trip instead of one per loop iteration
        if ((reinterpret_cast<uintptr_t>(P.grid) & 15u) == 0) {
            const uint4 *g4 = reinterpret_cast<const uint4 *>(P.grid);
            const int n16 = cells >> 4;
            for (int i0 = tid; i0 < n16; i0 += 4 * SM_THREADS) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = i0 + u * SM_THREADS < n16 ? __ldg(g4 + i0 + u * SM_THREADS) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (i0 + u * SM_THREADS < n16) reinterpret_cast<uint4 *>(g)[i0 + u * SM_THREADS] = v[u];
            }
            for (int i = (n16 << 4) + tid; i < cells; i += SM_THREADS) g[i] = P.grid[i];
        } else {
            for (int i = tid; i < cells; i += SM_THREADS) g[i] = P.grid[i];
        }
        __syncthreads();
        for (int i = tid; i < cells; i += SM_THREADS)
            if (g[i] == 1) atomicOr(&occb[i >> 5], 1u << (i & 31));
        for (int i = tid; i < cells; i += SM_THREADS) {
            const int x = i / H, y = i - x * H;
            unsigned nb = 0;
#pragma unroll
            for (int a = -1; a <= 1; a++)
#pragma unroll
                for (int b = -1; b <= 1; b++) {
                    const int xx = x + a, yy = y + b;
                    bool blk = xx < 0 || xx >= W || yy < 0 || yy >= H;
                    if (!blk) blk = g[xx * H + yy] == 1;
                    nb |= (blk ? 1u : 0u) << ((a + 1) * 3 + (b + 1));
                }
            auto B = [&](int a, int b) { return (nb >> ((a + 1) * 3 + (b + 1))) & 1u; };
            unsigned m = 0;
#pragma unroll
            for (int d = 0; d < 8; d++) {
                const int a = fx_dx(d), b = fx_dy(d);
                bool ok = !B(a, b);
                if (d >= 4) ok = ok && !(B(a, 0) && B(0, b));
                m |= (ok ? 1u : 0u) << d;
            }
            moves[i] = (uint8_t)m;
        }
        __syncthreads();
    }
    unsigned long long tot_settled = 0, tot_levels = 0, nq = 0;

    for (int q = blockIdx.x; q < P.Q; q += gridDim.x) {
        const int sx = P.starts[2 * q], sy = P.starts[2 * q + 1], gx = P.goals[2 * q], gy = P.goals[2 * q + 1];
        int32_t out_cost = FX_COST_UNREACHABLE;
        bool trivial = true;
        const bool s_in = sx >= 0 && sx < W && sy >= 0 && sy < H, g_in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        if (!s_in) out_cost = FX_COST_START_OOB;
        else if (!g_in) out_cost = FX_COST_UNREACHABLE;
        else if (sx == gx && sy == gy) out_cost = 0;                              // jps1.py:199-208
        else if ((occb[(gx * H + gy) >> 5] >> ((gx * H + gy) & 31)) & 1u) out_cost = FX_COST_UNREACHABLE;  // jump() tests the cell first
        else if (moves[sx * H + sy] == 0) out_cost = FX_COST_UNREACHABLE;          // the start cannot move
        else trivial = false;
        if (trivial) {
            if (tid == 0) {
                P.cost_i[q] = out_cost;
                if (P.cost_f) P.cost_f[q] = out_cost == 0 ? 0.0 : -1.0;
                if (P.path_len) P.path_len[q] = out_cost == 0 ? 1 : out_cost;
                if (out_cost == 0 && P.path_xy && P.max_path > 0) {
                    P.path_xy[(size_t)q * P.max_path * 2] = sx;
                    P.path_xy[(size_t)q * P.max_path * 2 + 1] = sy;
                }
            }
            continue;
        }
        nq++;
        const int sidx = sx * H + sy, gidx = gx * H + gy;
        __syncthreads();  // the previous query's path extraction is done with the field
        for (int i = tid; i < cells; i += SM_THREADS) field[i] = SM_INF;
        if (tid == 0) {
            s_cnt[0] = 1; s_cnt[1] = 0; s_cnt[2] = 0; s_cnt[3] = 0;
            s_goal[0] = SM_INF; s_goal[1] = SM_INF;
            queue[0] = (uint16_t)sidx;
        }
        __syncthreads();
        if (tid == 0) field[sidx] = (0u << 5) | (FX_CODE_START << 1) | 1u;
        __syncthreads();
        unsigned k = 0, my_settled = 0;
        uint32_t best = SM_INF;
        for (;;) {
            const unsigned n = s_cnt[k & 3], n1 = s_cnt[(k + 1) & 3];
            if (n == 0 && n1 == 0) break;  // nothing was pushed by level k-1 into k or k+1: the component is exhausted
            if (tid == 0) s_cnt[(k + 3) & 3] = 0;  // counter slot of bucket k-1 (done) becomes bucket k+3's
            const uint16_t *qk = queue + (size_t)(k % 3) * c4;
            uint16_t *q1 = queue + (size_t)((k + 1) % 3) * c4, *q2 = queue + (size_t)((k + 2) % 3) * c4;
            const uint32_t lo = k * WS, hi = lo + WS, hi2 = hi + WS;
            for (unsigned i = (unsigned)tid; i < n; i += SM_THREADS) {
                const int c = qk[i];
                const uint32_t v = field[c];
                const uint32_t g = v >> 5;
                if (!(v & 1u) || g < lo || g >= hi) continue;  // settled already, or improved into an earlier bucket
                field[c] = v & ~1u;  // one entry per cell and bucket: nobody else claims it
                my_settled++;
                if (c == gidx) s_goal[k & 1] = g;
                unsigned succ = lut[(((v >> 1) & 15u) << 8) | moves[c]];
                while (succ) {
                    const int d = __ffs(succ) - 1;
                    succ &= succ - 1;
                    const int nc = c + fx_dx(d) * H + fx_dy(d);
                    const uint32_t ng = g + (d < 4 ? WS : WD);
                    const uint32_t nv = (ng << 5) | ((unsigned)d << 1) | 1u;
                    const uint32_t old = atomicMin(&field[nc], nv);
                    if (nv < old) {
                        // unsettled cells hold costs of buckets k+1 or k+2 (or nothing): queue the cell only when its
                        // cost enters a bucket it was not in yet
                        const bool to2 = ng >= hi2;
                        const bool was2 = (old >> 5) >= hi2;
                        if (old == SM_INF || to2 != was2) {
                            const unsigned pos = atomicAdd(&s_cnt[(k + (to2 ? 2 : 1)) & 3], 1u);
                            (to2 ? q2 : q1)[pos] = (uint16_t)nc;
                        }
                    }
                }
            }
            __syncthreads();
            k++;
            // the goal was popped in the level that just ended (every cell of a popped bucket is final).  Level k writes the
            // other slot, so a warp that is already inside it does not race with this read.
            const unsigned gc = s_goal[(k - 1) & 1];
            if (gc != SM_INF) { best = gc; break; }
        }
        tot_settled += my_settled;
        tot_levels += k;
        __syncthreads();
        if (best == SM_INF) {
            if (tid == 0) {
                P.cost_i[q] = FX_COST_UNREACHABLE;
                if (P.cost_f) P.cost_f[q] = -1.0;
                if (P.path_len) P.path_len[q] = FX_COST_UNREACHABLE;
            }
            continue;
        }
        // path: warp 0 walks goal -> start along the arrival directions, 32 cells of a straight run per step, turning
        // points go straight to the output (reversed in place afterwards)
        int32_t *out = P.path_xy ? P.path_xy + (size_t)q * P.max_path * 2 : nullptr;
        const int cap = P.path_xy ? P.max_path : 0;
        if (tid < 32) {
            int vx = gx, vy = gy, npts = 1, prev_d = -1;
            unsigned a = 0, b = 0;
            uint32_t v = field[gidx];
            if (lane == 0 && cap > 0) { out[0] = vx; out[1] = vy; }
            bool bad = false;
            while (!(vx == sx && vy == sy)) {
                const int d = (int)((v >> 1) & 15u);
                if (v == SM_INF || d > 7) { bad = true; break; }
                if (prev_d >= 0 && d != prev_d) {
                    if (lane == 0 && npts < cap) { out[2 * npts] = vx; out[2 * npts + 1] = vy; }
                    npts++;
                }
                prev_d = d;
                const int ddx = fx_dx(d), ddy = fx_dy(d);
                const int ux = vx - (lane + 1) * ddx, uy = vy - (lane + 1) * ddy;
                uint32_t uv = SM_INF;
                if (ux >= 0 && ux < W && uy >= 0 && uy < H) uv = field[ux * H + uy];
                const unsigned cont = __ballot_sync(0xFFFFFFFFu, uv != SM_INF && (int)((uv >> 1) & 15u) == d);
                int lead = __ffs(~cont) - 1;
                if (cont == 0xFFFFFFFFu) lead = 31;
                const int run = lead + 1;
                v = __shfl_sync(0xFFFFFFFFu, uv, lead);
                vx -= run * ddx; vy -= run * ddy;
                if (d < 4) a += run; else b += run;
            }
            if (lane == 0 && npts < cap) { out[2 * npts] = sx; out[2 * npts + 1] = sy; }
            npts++;
            if (lane == 0) { s_npts = bad ? -1 : npts; s_ab[0] = a; s_ab[1] = b; }
        }
        __syncthreads();
        const int npts = s_npts;
        if (tid == 0) {
            P.cost_i[q] = npts < 0 ? FX_COST_OVERFLOW : (int32_t)best;
            if (P.cost_f) P.cost_f[q] = METRIC == 1 ? (double)best : __dadd_rn((double)s_ab[0], __dmul_rn((double)s_ab[1], 1.4142135623730951));
            if (P.path_len) P.path_len[q] = npts < 0 ? FX_COST_OVERFLOW : npts;
        }
        if (out && npts > 0 && npts <= cap) {
            // stored goal..start -> start..goal (when npts > cap the caller retries with a larger max_path)
            for (int i = tid; i < npts / 2; i += SM_THREADS) {
                const int j = npts - 1 - i;
                const int ax = out[2 * i], ay = out[2 * i + 1];
                out[2 * i] = out[2 * j]; out[2 * i + 1] = out[2 * j + 1];
                out[2 * j] = ax; out[2 * j + 1] = ay;
            }
        }
    }
    tot_settled = __reduce_add_sync(0xFFFFFFFFu, (unsigned)tot_settled);
    if (lane == 0 && tot_settled) atomicAdd(P.counters + 1, tot_settled);
    if (tid == 0) {
        atomicAdd(P.counters + 2, tot_levels);
        atomicAdd(P.counters + 3, nq);
    }
}

// cells <= FX_SMALL_CELLS: the shared-memory form.  Returns FX_OK after enqueueing, or 1 when the map does not qualify.
int fx_search_small(fx_context *ctx, const uint8_t *grid, int W, int H, const int32_t *starts_xy, const int32_t *goals_xy, int Q,
                    int metric, int32_t *cost_i, double *cost_f, int32_t *path_xy, int32_t *path_len, int max_path, cudaStream_t st)
{
    const size_t cells = (size_t)W * H;
    if (cells > FX_SMALL_CELLS) return 1;
    const size_t smem = small_smem_bytes(cells);
    if (!ctx->small_attr_set) {
        FX_CUDA(ctx, cudaFuncSetAttribute(k_search_small<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem_bytes(FX_SMALL_CELLS)));
        FX_CUDA(ctx, cudaFuncSetAttribute(k_search_small<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem_bytes(FX_SMALL_CELLS)));
        ctx->small_attr_set = 1;
    }
    FX_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, 16 * sizeof(unsigned long long), st));
    SmallParams P;
    P.grid = grid; P.W = W; P.H = H; P.starts = starts_xy; P.goals = goals_xy; P.Q = Q;
    P.cost_i = cost_i; P.cost_f = cost_f; P.path_xy = path_xy; P.path_len = path_len; P.max_path = path_xy ? max_path : 0;
    P.counters = ctx->counters;
    const int per_sm = (int)((size_t)200 * 1024 / (smem + 1024));
    const int cap = ctx->sm_count * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
    const int blocks = Q < cap ? Q : cap;
    FX_CUDA(ctx, cudaEventRecord(ctx->ev_search[0], st));
    if (metric == 1) k_search_small<1><<<blocks, SM_THREADS, smem, st>>>(P);
    else k_search_small<2><<<blocks, SM_THREADS, smem, st>>>(P);
    FX_LAUNCH_CHECK(ctx);
    FX_CUDA(ctx, cudaEventRecord(ctx->ev_search[1], st));
    ctx->ev_search_valid = 1;
    return FX_OK;
}
