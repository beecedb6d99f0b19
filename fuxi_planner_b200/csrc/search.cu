// search.cu -- batched shortest-path queries on the 8-connected occupancy grid (sm_100a).
//
// Replaces scripts/jps1.py:183-230 (method) and everything below it for Q queries per launch.
// Graph: edges are exactly `not blocked(c, d)` of scripts/jps1.py:14-31, precomputed once per
// grid into an 8-bit legal-move mask per cell (k_build_moves).
//
// Algorithm (one persistent CTA per concurrent query, queries fetched from an atomic counter):
//   level-synchronous wavefront = Dial / delta-stepping with bucket width == the straight weight.
//   Every edge weighs >= one bucket, so when bucket k is popped all of its cells are final (no
//   intra-bucket re-relaxation), a straight move lands in bucket k+1 and a diagonal in k+1 or k+2:
//   four rotating bucket queues suffice and one __syncthreads() per level orders everything.
//   Frontier compaction: warp-ballot + popc prefix, one shared-memory atomicAdd per warp and bucket.
//   Pruning: a cell is relaxed only if g + octile(cell, goal) <= U (A*-style ellipse); a pass that
//   reaches the goal with cost <= U is exact, because every path of cost <= U lies inside the ellipse.
//   U comes from a first pass restricted to a narrow band around the start-goal line (any path it
//   finds is a valid upper bound); if that cost already equals the octile lower bound it is final.
//   The exact pass is BIDIRECTIONAL: one wavefront from the start and one from the goal advance in the
//   same levels (same bucket index, same queues, one bit of the queue entry says which side it belongs
//   to), each into its own cost field, each pruned by its own ellipse; a cell popped by one side looks
//   at the other side's field and proposes mu = g_s + g_t; once 2*k*WS > mu + WD the proposal is the
//   optimum (see run_pass).  Both halves of the ellipse are settled once, but the chain of dependent
//   levels is half as long and a level carries twice the cells.
//   The cost field lives in a per-slot scratch array in HBM (L2-resident working set); only the
//   128-byte lines a query touched are reset afterwards (dirty flags).
//   Path: warp-parallel descent from the goal along cost-consistent predecessors, 32 cells of a
//   straight run per round trip, emitting turning points.
#include <stdlib.h>

#include <cooperative_groups.h>

#include "common.cuh"
#if !FX_TILED
#error "search.cu relies on the 8x8-tiled scratch layout (step tables)"
#endif

#define FLAG_OVERFLOW 1u
#define FLAG_UNREACH 2u /* bidirectional pass: one side exhausted its component unpruned without meeting the other */
#define FX_LAT_SQ 2048 /* latency form: entries per bucket half kept in shared memory (4 x 2 x 2048 x 8 B = 128 KiB) */
#define FX_POCKET_BUDGET 4096u /* queue pops of the bounded flood from the goal */

// -DFX_PHASE_CLOCKS: tuning build that accumulates, for warp 0 of every CTA, the cycles spent in each dependent step of
// a level (counters[8..15]; read with fx_search_phase_clocks).  Not compiled into the default library.
#ifdef FX_PHASE_CLOCKS
__device__ __forceinline__ long long fx_clk_after(uint32_t dep)
{
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(dep) : "memory");
    return t;
}
#define PH_DECL long long ph_t = 0, ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PH_START(dep) ph_t = fx_clk_after(dep);
#define PH_MARK(i, dep) { const long long t__ = fx_clk_after(dep); ph_acc[i] += t__ - ph_t; ph_t = t__; }
#define PH_COUNT(i, n) ph_acc[i] += (n);
#define PH_FLUSH if (threadIdx.x == 0) { for (int i__ = 0; i__ < 8; i__++) atomicAdd(P.counters + 8 + i__, (unsigned long long)ph_acc[i__]); }
#else
#define PH_DECL
#define PH_START(dep)
#define PH_MARK(i, dep)
#define PH_COUNT(i, n)
#define PH_FLUSH
#endif

__device__ __forceinline__ void fx_red_min(uint32_t *p, uint32_t v)
{
    asm volatile("red.relaxed.gpu.global.min.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// shared-memory atomic add through PTX: keeps ptxas from rewriting it into a warp-aggregated shuffle sequence
__device__ __forceinline__ unsigned fx_atoms_add(unsigned *p, unsigned v)
{
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
    return old;
}

// First prune bound of the latency forms.  How far the optimum lies above the octile lower bound h0 depends, on a given
// map, mostly on the MIX of the query's moves: with t = min(diagonal, straight) / max(diagonal, straight) steps of the
// octile path, queries near t = 0 (nearly pure straight or pure diagonal) have no free choice of lanes and pay for every
// obstacle (5.4 % over h0 on the 20 %-filled headline grid), balanced ones hardly anything (0.05 %).  A fixed guess is
// too small for the first group (a second pass: 39 % of the headline queries) or needlessly wide for the second.  So
// the context keeps, per t-bin, a decaying maximum of the ratios (U* - h0) / h0 it has seen (16 x u32 in units of 2^-20,
// 0 on a fresh context) and guesses from it.  Only a guess: a pass is accepted iff its result is <= the bound it pruned
// with (run_pass), so the table changes how many passes a query takes, never its answer.
#define FX_CALIB_FLOOR 1024u /* 2^-10: what a fresh table guesses (+ 4 WD) */
__device__ __forceinline__ int fx_calib_bin(int adx, int ady)
{
    const int mn = min(adx, ady), st = max(adx, ady) - mn;  // diagonal / straight steps of the octile path
    const int lo = min(mn, st), hi = max(max(mn, st), 1);
    return min(15, (int)(16.f * sqrtf((float)lo / (float)hi)));
}
__device__ __forceinline__ uint64_t fx_first_bound(const uint32_t *calib, int bin, uint32_t h0, uint32_t wd)
{
    if (!calib) return (uint64_t)h0 + h0 / 128 + 4ull * wd;  // FUXI_B200_CALIB=0: the fixed guess of the first version
    uint32_t r = __ldcg(calib + bin);
    if (bin < 15) r = max(r, __ldcg(calib + bin + 1) >> 1);  // a sparsely visited bin borrows from its neighbour
    if (r == 0u) r = 7168u;                                  // nothing learnt yet: the fixed guess (1/128 with the margin below)
    r = min(r, 1u << 19);                                              // never guess beyond 1.5 h0: a maze-like map is served by the widening
    r = r + (r >> 3) + FX_CALIB_FLOOR;                                 // 1/8 margin over the recent maximum
    return (uint64_t)h0 + (((uint64_t)h0 * r) >> 20) + 4ull * wd;
}
__device__ __forceinline__ void fx_calib_update(uint32_t *calib, int bin, uint32_t h0, uint32_t best, uint32_t ws)
{
    if (!calib || h0 < 64u * ws || best < h0) return;  // short queries say little about the ratio
    const uint64_t r64 = (((uint64_t)(best - h0)) << 20) / h0;
    const uint32_t r = r64 > 0x3FFFFFFFull ? 0x3FFFFFFFu : (uint32_t)r64;
    const uint32_t old = __ldcg(calib + bin);
    __stcg(calib + bin, max(r, old - (old >> 5)));  // racy among concurrent queries on purpose: any of the values will do
}

struct SearchParams {
    const uint8_t *grid;
    const uint8_t *moves;
    int W, H, TY;
    const int32_t *starts, *goals;
    int Q;
    int32_t *cost_i;
    double *cost_f;
    int32_t *path_xy;
    int32_t *path_len;
    int max_path;
    uint32_t *fields;
    uint8_t *dirty;
    uint2 *queues;      // [slots][4][qcap] (packed xy, packed cost|direction)
    int32_t *tmp_path;
    size_t cells, dirty_n;
    int qcap, path_cap;
    unsigned long long *counters;
    int band0;
    const uint32_t *order;   // LPT query order (band.cu) or NULL
    const uint32_t *ubound;  // per-query upper bound from the band pass or NULL
    uint32_t *calib;         // latency forms: first-bound table (fx_first_bound)
};

// ------------------------------------------------------------------------------------------------
// legal-move mask: bit d of moves[c] == not blocked(c, dir d)   (scripts/jps1.py:14-31)
// ------------------------------------------------------------------------------------------------
template <bool TILED>
__global__ void __launch_bounds__(256) k_build_moves(const uint8_t *__restrict__ grid, int W, int H,
                                                     uint8_t *__restrict__ moves, int x0, int x1)
{
    const int TY = fx_tiles_y(H);
    size_t total = (size_t)(x1 - x0) * H;  // rows [x0, x1) (their masks read rows x0 - 1 .. x1 of the grid)
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (size_t)gridDim.x * blockDim.x) {
        const size_t i = j + (size_t)x0 * H;
        int x = (int)(i / H), y = (int)(i - (size_t)x * H);
        // 3x3 neighbourhood: 1 = obstacle (== 1) or outside the array
        unsigned nb = 0;  // bit (dx+1)*3 + (dy+1)
#pragma unroll
        for (int a = -1; a <= 1; a++)
#pragma unroll
            for (int b = -1; b <= 1; b++) {
                int xx = x + a, yy = y + b;
                bool blk = xx < 0 || xx >= W || yy < 0 || yy >= H;
                if (!blk) blk = __ldg(grid + (size_t)xx * H + yy) == 1;
                nb |= (blk ? 1u : 0u) << ((a + 1) * 3 + (b + 1));
            }
        auto B = [&](int a, int b) { return (nb >> ((a + 1) * 3 + (b + 1))) & 1u; };
        unsigned m = 0;
#pragma unroll
        for (int d = 0; d < 8; d++) {
            int a = fx_dx(d), b = fx_dy(d);
            bool ok = !B(a, b);
            if (d >= 4) ok = ok && !(B(a, 0) && B(0, b));
            m |= (ok ? 1u : 0u) << d;
        }
        moves[TILED ? (size_t)fx_cidx(x, y, H, TY) : i] = (uint8_t)m;
    }
}

// rows [x0, x1) of the mask (the host upload builds it part by part behind the H2D copies: fx_plan_host)
int fx_build_moves_rows(fx_context *ctx, const uint8_t *grid, int W, int H, bool tiled, int x0, int x1, cudaStream_t st)
{
    size_t total = (size_t)W * H;
    const size_t need = fx_scratch_cells(W, H) > total ? fx_scratch_cells(W, H) : total;
    if (ctx->moves_cap < need) {
        if (ctx->moves) cudaFree(ctx->moves);
        ctx->moves = nullptr; ctx->moves_cap = 0;
        FX_CUDA(ctx, cudaMalloc(&ctx->moves, need));
        ctx->moves_cap = need;
    }
    if (x1 <= x0) return FX_OK;
    const size_t part = (size_t)(x1 - x0) * H;
    int blocks = (int)((part + 255) / 256);
    int maxb = ctx->sm_count * 16;
    if (blocks > maxb) blocks = maxb;
    if (tiled && FX_TILED) k_build_moves<true><<<blocks, 256, 0, st>>>(grid, W, H, ctx->moves, x0, x1);
    else k_build_moves<false><<<blocks, 256, 0, st>>>(grid, W, H, ctx->moves, x0, x1);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}

int fx_build_moves(fx_context *ctx, const uint8_t *grid, int W, int H, bool tiled, cudaStream_t st)
{
    return fx_build_moves_rows(ctx, grid, W, H, tiled, 0, W, st);
}

// ------------------------------------------------------------------------------------------------
// per-CTA shared state
// ------------------------------------------------------------------------------------------------
struct __align__(16) CtaState {
    unsigned tailS[4], tailD[4];  // entries per bucket: cells that arrived by a straight / by a diagonal move
    unsigned long long mu[3];  // bidirectional pass: min over cells seen by both sides of (g_s + g_t) << 32 | packed xy; slot = level % 3
    unsigned goal[2];  // cost of the goal cell once it has been popped (FX_INF = not yet); slot = parity of the level that popped it
    unsigned meet;     // bidirectional pass: packed xy of the cell the two sides met at
    unsigned ovf_level;  // level + 1 in which a cost left the 28-bit range (0 = never)
    unsigned U;      // prune bound on g + h
    unsigned flags;
    int xlo, xhi;     // x-rows this query has touched since the last reset (bounds the dirty-flag scan)
    unsigned pruned;  // this pass rejected a legal move by the ellipse or the band (so a drained queue proves nothing)
    unsigned prl[3][2];  // bidirectional pass: side s pruned a cell in the level whose index % 3 is the slot (rotates like mu)
    unsigned alive[2][4];  // bidirectional pass: side s has pushed an entry into bucket slot b
    int q;
    unsigned long long settled, levels;
};

// One search pass.  Returns (to every thread) the goal cost or FX_INF.  bandL < 0 disables the band.
// budget > 0 stops the pass (returning FX_INF with *budget_hit = true) once more than `budget` queue entries were popped.
//
// Successor generation is CANONICAL (the pruning of scripts/jps1.py:49-93 `nodeNeighbours`, applied at every cell
// instead of only at jump points): a cell popped with arrival direction `code` relaxes only its natural and forced
// neighbours, s_lut[code][moves] (fx_canon_succ) -- ~1.4 relaxations per settled cell instead of 8.  The cost field
// packs (cost << 4 | arrival direction) so that one atomic min keeps cost and parent together; ties on cost resolve
// to the lower direction code, which the pruning tolerates (every optimal parent works: the JPS argument; checked
// against plain Dijkstra with random tie-breaking on whole fields).
//
// No memory round trip sits between popping a cell and pushing its children: a relaxation is a fire-and-forget
// RED.MIN on the packed word plus an unconditional queue entry (child xy, packed value) in the bucket of the new
// cost.  Whether the relaxation won is decided when the entry is POPPED: it is expanded iff the field still holds
// exactly the entry's value (the writer of the final minimum is unique), so losers and superseded entries drop out
// there.  A level is therefore: queue load -> field + move-mask load -> ALU -> stores, then one block barrier.
//
// BIDIR: a second wavefront starts at the goal and advances in the same levels.  Bit 31 of a queue entry's packed
// xy says which side it belongs to; side 1 uses the second cost field of the slot (field + P.cells) and prunes with
// the ellipse around the START.  Among free cells the move graph is symmetric (a diagonal tests the same two
// orthogonal cells both ways; the goal is free, and the search never enters an obstacle), so the wavefront from the
// goal settles exact cost-TO-goal values.  Every valid pop in the late levels (k >= h0/(2 WS) - 2: two labels of one
// cell sum to at least h0) also reads the other side's word of the same cell and proposes mu = g + g_other, always
// the cost of a real path.  Termination: at the top of level k both sides have settled every cell of cost < k*WS.  Walk
// the optimal path P from the start and let c be its last cell with g_s(c) < k*WS: g_s(c) >= k*WS - WD, so
// g_t(c) = U* - g_s(c) < k*WS as soon as 2*k*WS > U* + WD, i.e. c was popped by both sides with exact values and
// proposed mu = U*.  The loop tests 2*k*WS > mu + WD with the smallest proposal so far; mu >= U* makes the test
// imply the condition above, so whatever it returns is the optimum, and it fires at the first level the condition
// holds.  A start on an obstacle can leave it but cannot be entered: the goal side never labels it, and c != start
// because the test needs k >= 2.  The cell that proposed the minimum is returned in S.meet.
//
// SQ > 0 (latency form, one CTA per SM): the first SQ entries of each half of each bucket live in shared memory
// (sq[4][2][SQ], 64 * SQ bytes), the rest spills to the global queue.  A level of a 4096^2 query holds a few hundred
// entries per half, so the pop of a level no longer waits for an L2 round trip on the queue before it can issue the
// field load: one dependent L2 access per level instead of two.
template <int METRIC, bool BIDIR, int SQ>
__device__ uint32_t run_pass(const SearchParams &P, CtaState &S, const uint8_t *__restrict__ s_lut, const int *__restrict__ s_step,
                             uint32_t *__restrict__ field, uint8_t *__restrict__ dirty, uint2 *__restrict__ queue, uint2 *__restrict__ sq,
                             int sx, int sy, int gx, int gy, uint32_t U0, float bandL, unsigned budget, bool *budget_hit)
{
    constexpr uint32_t WS = Wt<METRIC>::WS, WD = Wt<METRIC>::WD;
    const int H = P.H;
    const int tid = threadIdx.x, lane = tid & 31, nthreads = blockDim.x;
    const unsigned qcap = (unsigned)P.qcap, qhalf = qcap >> 1;
    const int TY = P.TY;
    const unsigned side_cells = (unsigned)P.cells;  // word offset of the goal side's field (2 * cells <= 2^31: W, H <= 32767)
    const unsigned sidx = (unsigned)fx_cidx(sx, sy, H, TY), gidx = (unsigned)fx_cidx(gx, gy, H, TY);
    const float qdx = (float)(gx - sx), qdy = (float)(gy - sy);
    const uint8_t *__restrict__ moves = P.moves;
    // cell index (move masks) and field index (cost word: + side_cells for the goal side) of a queue entry's packed xy
    auto cell_of = [&](uint32_t exy) { return (unsigned)fx_cidx((int)((exy >> 16) & 0x7FFFu), (int)(exy & 0xFFFFu), H, TY); };
    auto foff_of = [&](uint32_t exy) { return BIDIR ? (exy >> 31) * side_cells : 0u; };
    // entry j of class c (0: straight arrivals, 1: diagonal arrivals) of bucket slot b
    auto q_load = [&](unsigned b, unsigned c, unsigned j) -> uint2 {
        if (SQ > 0 && j < (unsigned)SQ) return sq[(b * 2u + c) * (unsigned)SQ + j];
        return __ldcg(queue + (size_t)b * qcap + (c ? qhalf : 0u) + j);
    };
    auto q_store = [&](unsigned b, unsigned c, unsigned j, uint2 v) {
        if (SQ > 0 && j < (unsigned)SQ) sq[(b * 2u + c) * (unsigned)SQ + j] = v;
        else __stcg(queue + (size_t)b * qcap + (c ? qhalf : 0u) + j, v);
    };

    if (tid == 0) {
        S.tailS[0] = BIDIR ? 2 : 1; S.tailS[1] = 0; S.tailS[2] = 0; S.tailS[3] = 0;
        S.tailD[0] = 0; S.tailD[1] = 0; S.tailD[2] = 0; S.tailD[3] = 0;
        S.goal[0] = FX_INF; S.goal[1] = FX_INF; S.U = U0; S.pruned = 0; S.ovf_level = 0;
        S.mu[0] = ~0ull; S.mu[1] = ~0ull; S.mu[2] = ~0ull; S.meet = 0;
        for (int a = 0; a < 6; a++) S.prl[a >> 1][a & 1] = 0;
        for (int a = 0; a < 8; a++) S.alive[a >> 2][a & 3] = 0;
        S.alive[0][0] = 1; S.alive[1][0] = 1;
        S.xlo = min(S.xlo, sx - 1); S.xhi = max(S.xhi, sx + 1);
        q_store(0u, 0u, 0u, make_uint2(((uint32_t)sx << 16) | (uint32_t)sy, fx_pack(0u, FX_CODE_START)));
        __stcg(field + sidx, fx_pack(0u, FX_CODE_START));
        dirty[sidx >> FX_DIRTY_SHIFT] = 1;
        if (BIDIR) {
            S.xlo = min(S.xlo, gx - 1); S.xhi = max(S.xhi, gx + 1);
            q_store(0u, 0u, 1u, make_uint2(0x80000000u | ((uint32_t)gx << 16) | (uint32_t)gy, fx_pack(0u, FX_CODE_START)));
            __stcg(field + (side_cells + gidx), fx_pack(0u, FX_CODE_START));
            dirty[(side_cells + gidx) >> FX_DIRTY_SHIFT] = 1;
        }
    }
    __syncthreads();

    unsigned my_settled = 0;
    int my_xlo = 0x7FFFFFFF, my_xhi = -1;
    unsigned k = 0, popped = 0;
    unsigned k3 = 0;  // k % 3
    bool my_pruned = false;
    uint32_t result = FX_INF;
    unsigned long long mu = ~0ull;
    unsigned prl0 = 0, prl1 = 0;
    // the first level in which a cell can carry labels of both sides
    const uint32_t h0 = octile(abs(sx - gx), abs(sy - gy), WS, WD - WS);
    const unsigned k_meet = h0 / (2u * WS) > 2u ? h0 / (2u * WS) - 2u : 0u;
    const bool start_free = BIDIR && P.grid[(size_t)sx * H + sy] != 1;
    *budget_hit = false;
    PH_DECL
    for (;;) {
        // A bucket's array holds the straight-arrival entries in its first half and the diagonal-arrival entries in
        // its second half: a straight arrival has 1 natural (+ <= 2 forced) successor, a diagonal one 3 (+ <= 2), so
        // warps that pop one class run the relaxation loop about the same number of times in every lane.
        const unsigned nS = S.tailS[k & 3], nD = S.tailD[k & 3], n = nS + nD;
        const unsigned n1 = S.tailS[(k + 1) & 3] + S.tailD[(k + 1) & 3];
        PH_START(n)
        // Shared state read here must look the same to a warp that is still at the top of level k and to one that is
        // already inside it: S.goal is double-buffered by level parity (level k writes slot k & 1, this reads the slot of
        // level k-1, complete since the last barrier); S.mu rotates over three slots the same way (level k writes slot
        // k % 3, this reads the slot of level k-1 and clears the slot of level k+1, last read at the top of level k-1);
        // S.ovf_level only counts once its level is over; n is complete since the last barrier; n1 is still growing but
        // only matters when n == 0, i.e. when nobody pushes in this level.
        if (BIDIR) {
            const unsigned kp = k3 == 0 ? 2 : k3 - 1, kn = k3 == 2 ? 0 : k3 + 1;
            const unsigned long long m_prev = S.mu[kp];
            mu = m_prev < mu ? m_prev : mu;
            prl0 |= S.prl[kp][0]; prl1 |= S.prl[kp][1];  // what level k-1 pruned (complete); every thread keeps the running OR
            const uint32_t mc = (uint32_t)(mu >> 32);
            if (mc != FX_INF && 2ull * k * WS > (unsigned long long)mc + WD) { result = mc; break; }  // see the proof above
            // A side with nothing in buckets k and k+1 is finished (bucket k+2 is only filled by this level's own pops).  If
            // it never pruned, it has labelled its whole component: no proposal by now means the other end is not in it
            // (the start side pops the goal -- labelled by the goal side from level 0 -- before it can run dry; the goal side
            // pops the start likewise if the start is a free cell.  A start on an obstacle can be left but never entered: there
            // the goal side instead pops a free cell n the start can step into, after level 0 -- n carries the start side's
            // label from level 0 on -- unless n is the goal itself, which the start side pops in level 1: so from level 2 on
            // the goal side's exhaustion counts for obstacle starts as well).
            // The alive words are stable here: nobody pushes for a side that has no entry in bucket k.
            if (mc == FX_INF) {
                const bool dead0 = !S.alive[0][k & 3] && !S.alive[0][(k + 1) & 3], dead1 = !S.alive[1][k & 3] && !S.alive[1][(k + 1) & 3];
                if ((dead0 && !prl0) || (dead1 && !prl1 && (start_free || k >= 2u))) { if (tid == 0) S.flags |= FLAG_UNREACH; break; }
            }
            if (tid == 0) { S.mu[kn] = ~0ull; S.prl[kn][0] = 0; S.prl[kn][1] = 0; S.alive[0][(k + 3) & 3] = 0; S.alive[1][(k + 3) & 3] = 0; }
        } else {
            const unsigned goalc = S.goal[(k + 1) & 1];
            if (goalc != FX_INF) { result = goalc; break; }  // the goal was popped in the previous level: final
        }
        const unsigned ovl = S.ovf_level;
        if (ovl != 0 && ovl <= k) { if (tid == 0) S.flags |= FLAG_OVERFLOW; break; }
        if ((n == 0 && n1 == 0) || (S.flags & FLAG_OVERFLOW)) {
            // drained: every cell inside the pruning region is settled (by both sides), so the smallest proposal is exact
            if (BIDIR) result = (uint32_t)(mu >> 32);
            break;
        }
        popped += n;
        if (budget && popped > budget) { *budget_hit = true; break; }
        if (nS > qhalf || nD > qhalf) { if (tid == 0) S.flags |= FLAG_OVERFLOW; break; }  // entries beyond a half were dropped at push time
        if (tid == 0) { S.tailS[(k + 3) & 3] = 0; S.tailD[(k + 3) & 3] = 0; }  // bucket k-1 is done; levels k+1.. will refill this slot
        const uint32_t U = S.U;
        const unsigned bk = k & 3u, b1 = (k + 1u) & 3u, b2 = (k + 2u) & 3u;
        const uint32_t kbase = k * WS;
        const bool meet_level = BIDIR && k >= k_meet;
        auto pop = [&](unsigned i) { return i < nS ? q_load(bk, 0u, i) : q_load(bk, 1u, i - nS); };
        // Two-deep software pipeline over the rounds of a level: while round r is processed, the cost word and move mask
        // of round r+1 and the queue entry of round r+2 are already in flight, so only the first round of a level waits
        // for two dependent L2 round trips.  Loading a cell's word before this level's own reductions are issued is safe:
        // the cells popped in level k hold costs of bucket k and every reduction of level k carries a cost of bucket
        // k+1 or later, so their words do not change during the level.
        uint2 e_cur = (unsigned)tid < n ? pop((unsigned)tid) : make_uint2(0u, 0u);
        uint2 e_nxt = (unsigned)tid + (unsigned)nthreads < n ? pop((unsigned)tid + (unsigned)nthreads) : make_uint2(0u, 0u);
        unsigned idx_cur = cell_of(e_cur.x), fo_cur = foff_of(e_cur.x);
        uint32_t v_cur = FX_INF;
        unsigned m_cur = 0;
        if ((unsigned)tid < n) { v_cur = __ldcg(field + (idx_cur + fo_cur)); m_cur = (unsigned)__ldg(moves + idx_cur); }
        for (unsigned i0 = (unsigned)(tid - lane); i0 < n; i0 += (unsigned)nthreads) {
            const unsigned i = i0 + lane;
            bool act = i < n;
            const uint2 e = e_cur;
            const unsigned idx = idx_cur, fo = fo_cur;  // idx < 2^30 (W, H <= 32767)
            const uint32_t v = v_cur;
            const unsigned m = m_cur;
            e_cur = e_nxt;
            idx_cur = cell_of(e_cur.x); fo_cur = foff_of(e_cur.x);
            v_cur = FX_INF;
            m_cur = 0;
            if (i + nthreads < n) { v_cur = __ldcg(field + (idx_cur + fo_cur)); m_cur = (unsigned)__ldg(moves + idx_cur); }
            if (i + 2u * nthreads < n) e_nxt = pop(i + 2u * nthreads);
            const int x = (int)((e.x >> 16) & 0x7FFFu), y = (int)(e.x & 0xFFFFu);
            const bool side = BIDIR && (e.x >> 31) != 0u;
            PH_MARK(0, e.x)  // queue entry arrived
            PH_MARK(1, v + m)  // cost + move mask arrived
            const uint32_t g = e.y >> 4;
            act = act && v == e.y;  // this entry's relaxation won and nothing improved the cell since
            uint32_t other = FX_INF;
            if (meet_level && act) other = __ldcg(field + (idx + (side ? 0u : side_cells)));
            if (act) {
                // every relaxed cell has an entry that carries its final word (the winner's): marking the line when THAT
                // entry is popped -- or swept below if it never is -- covers every touched line; stale entries skip the store
                dirty[(idx + fo) >> FX_DIRTY_SHIFT] = 1;
                if (!BIDIR && idx == gidx) S.goal[k & 1] = g;  // unique winner: plain store
                // prune at POP time: a cell outside the ellipse g + h <= U (or outside the band) keeps its cost but is
                // not expanded.  Every cell of a path of cost <= U satisfies g*(c) + h(c) <= U (h is consistent), and so
                // do the cells of the alternative paths the canonical pruning relies on: exactness holds.
                const int tx = side ? sx : gx, ty = side ? sy : gy;
                const uint32_t h = octile(abs(x - tx), abs(y - ty), WS, WD - WS);
                bool keep = ((uint64_t)g + h) <= (uint64_t)U;
                if (bandL >= 0.f) {
                    const float lat = (float)(x - sx) * qdy - (float)(y - sy) * qdx;
                    keep = keep && fabsf(lat) <= bandL;
                }
                if (!keep) { my_pruned = true; act = false; if (BIDIR) S.prl[k3][side ? 1 : 0] = 1; }
            }
            if (BIDIR && other != FX_INF)  // both sides have labelled this cell: a real start-goal path through it
                atomicMin(&S.mu[k3], ((unsigned long long)(g + (other >> 4)) << 32) | (unsigned long long)(e.x & 0x7FFFFFFFu));
            unsigned succ = 0;
            if (act) {
                my_settled++; my_xlo = min(my_xlo, x); my_xhi = max(my_xhi, x);
                succ = s_lut[((e.y & 15u) << 8) | m];
                if (g + WD > FX_COST_MAX28) { S.ovf_level = k + 1; succ = 0; }  // 28-bit cost range (benign race: same value)
            }
            PH_MARK(2, succ)  // successor mask ready
            unsigned posS = 0, posD = 0;
            // a straight child lands in bucket k+1; all diagonal children of this cell land in the same bucket, k+1 or k+2
            const bool diag2 = (g - kbase) + WD >= 2u * WS;
            if (succ) {
                // queue space: one shared-memory atomic per lane and class (the unit serialises same-address lanes in
                // ~1-2 cycles each).  Measured against a per-warp reservation (packed prefix scan over five shuffles, one
                // three-address atomic by lanes 0..2, two shuffles to hand the bases back): the scan costs more issue slots
                // than the serialised atomics cost LSU wavefronts, k_search_batch 92.4 -> 96.0 ms (r02j), so this stays.
                const unsigned ns = (unsigned)__popc(succ & 0x0Fu), nd = (unsigned)__popc(succ & 0xF0u);
                if (ns) posS = fx_atoms_add(&S.tailS[(k + 1) & 3], ns);
                if (nd) posD = fx_atoms_add(&S.tailD[(k + (diag2 ? 2 : 1)) & 3], nd);
                if (posS + ns > qhalf || posD + nd > qhalf) succ = 0;  // no room: the tail counts flag the overflow at the next level
                if (BIDIR) {
                    if (ns) S.alive[side ? 1 : 0][(k + 1) & 3] = 1;
                    if (nd) S.alive[side ? 1 : 0][(k + (diag2 ? 2 : 1)) & 3] = 1;
                }
            }
            PH_MARK(3, posS + posD)  // queue space reserved
            // the loop below runs as often as the busiest lane of the warp has successors, so it is kept short: the
            // child's packed xy, its tiled index and its packed value are one table lookup + one add each
            const unsigned bD = diag2 ? b2 : b1;
            // global slots (pointer bumps in the loop below: the throughput form's inner loop is exactly the r01 one) and,
            // with SQ, the shared-memory heads of the same bucket halves
            uint2 *__restrict__ gS = queue + (size_t)b1 * qcap + posS;
            uint2 *__restrict__ gD = queue + (size_t)bD * qcap + qhalf + posD;
            uint2 *__restrict__ sS = SQ > 0 ? sq + (b1 * 2u) * (unsigned)SQ : nullptr;
            uint2 *__restrict__ sD = SQ > 0 ? sq + (bD * 2u + 1u) * (unsigned)SQ : nullptr;
            const uint32_t nvS = fx_pack(g + WS, 0u), nvD = fx_pack(g + WD, 0u);
            const int *__restrict__ stp = s_step + 2 * (idx & 63);  // tile-local position (x & 7) << 3 | (y & 7)
            uint32_t *__restrict__ fbase = field + (idx + fo);
            while (succ) {
                const int d = __ffs(succ) - 1;
                succ &= succ - 1;
                const int2 st = *reinterpret_cast<const int2 *>(stp + 2 * 64 * d);  // (packed xy step, tiled index step)
                const uint32_t nv = (d < 4 ? nvS : nvD) | (unsigned)d;
                fx_red_min(fbase + st.y, nv);
                const uint2 child = make_uint2(e.x + (uint32_t)st.x, nv);  // keeps the side bit
                if (d < 4) {
                    if (SQ > 0 && posS < (unsigned)SQ) sS[posS] = child; else __stcg(gS, child);
                    gS++; posS++;
                } else {
                    if (SQ > 0 && posD < (unsigned)SQ) sD[posD] = child; else __stcg(gD, child);
                    gD++; posD++;
                }
            }
            PH_MARK(4, posS)  // children relaxed and appended
            PH_COUNT(6, 1)    // rounds (warp 0)
        }
        __syncthreads();
        PH_MARK(5, k)  // barrier (includes waiting for the other warps' rounds)
        PH_COUNT(7, 1)  // levels
        k++;
        k3 = k3 == 2 ? 0 : k3 + 1;
    }
    PH_FLUSH
    // entries that were never popped (buckets k and k+1; k+2 is still empty at the top of a level): mark their lines too
    for (unsigned b = k; b <= k + 1; b++) {
        const unsigned mS = min(S.tailS[b & 3], qhalf), mD = min(S.tailD[b & 3], qhalf);
        for (unsigned i = (unsigned)tid; i < mS + mD; i += (unsigned)nthreads) {
            const uint32_t xy = (i < mS ? q_load(b & 3u, 0u, i) : q_load(b & 3u, 1u, i - mS)).x;
            dirty[(cell_of(xy) + foff_of(xy)) >> FX_DIRTY_SHIFT] = 1;
        }
    }
    // every thread leaves the loop at the same k with the same decision (all read the same shared state
    // after the same barrier); one more barrier so that nobody is still reading S when it is re-initialised
    if (my_xhi >= 0) { atomicMin(&S.xlo, my_xlo - 1); atomicMax(&S.xhi, my_xhi + 1); }
    if (my_pruned) S.pruned = 1;  // S.pruned was zeroed before the first barrier of this pass; nobody reads it until the next one
    if (BIDIR && tid == 0) S.meet = (unsigned)(mu & 0x7FFFFFFFull);
    __syncthreads();
    {
        // one 64-bit shared-memory atomic per warp (it is a CAS loop in SASS: 128 lanes on one address would spin)
        const unsigned ws = __reduce_add_sync(0xFFFFFFFFu, my_settled);
        if (lane == 0 && ws) atomicAdd(&S.settled, (unsigned long long)ws);
    }
    if (tid == 0) S.levels += k;
    return result;
}

// reset every 128-byte field line this query touched: only the dirty flags of the x-rows [xlo, xhi] are scanned
// (`both`: in the goal side's field as well -- after a bidirectional pass)
__device__ void reset_slot(CtaState &S, uint32_t *__restrict__ field, uint8_t *__restrict__ dirty, size_t dirty_n, size_t cells, int H, bool both)
{
    __syncthreads();
    const int xlo = max(S.xlo, 0), xhi = S.xhi;
    // after an overflow some relaxed cells have no queue entry (and so no dirty flag): reset every line of the x-range
    const bool force = (S.flags & FLAG_OVERFLOW) != 0;
    __syncthreads();
    if (threadIdx.x == 0) { S.xlo = 0x7FFFFFFF; S.xhi = -1; }
    if (xhi < xlo) { __syncthreads(); return; }
    const size_t c_lo = (size_t)(xlo >> 3) * fx_tiles_y(H) * 64;
    size_t c_hi = (size_t)((xhi >> 3) + 1) * fx_tiles_y(H) * 64;
    if (c_hi > cells) c_hi = cells;
    const size_t n16 = dirty_n / 16;  // dirty_n is padded to a multiple of 16 and covers both fields (2 * cells words)
    uint4 *d4 = reinterpret_cast<uint4 *>(dirty);
    const uint4 inf4 = make_uint4(FX_INF, FX_INF, FX_INF, FX_INF);
    for (int f = 0; f < (both ? 2 : 1); f++) {
        const size_t base = (size_t)f * cells;  // cells is a multiple of 512: the second field's flags start on a uint4
        size_t i0 = ((base + c_lo) >> FX_DIRTY_SHIFT) / 16, i1 = (((base + c_hi) >> FX_DIRTY_SHIFT) + 15) / 16;
        if (i1 > n16) i1 = n16;
        for (size_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
            uint4 v = __ldcg(d4 + i);
            if (force) v = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
            if ((v.x | v.y | v.z | v.w) == 0u) continue;
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int a = 0; a < 4; a++) {
                if (!w[a]) continue;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if (!((w[a] >> (8 * b)) & 0xFFu)) continue;
                    const size_t chunk = i * 16 + a * 4 + b;
                    const size_t c0 = chunk << FX_DIRTY_SHIFT;
                    if (c0 + 32 <= 2 * cells) {
                        uint4 *f4 = reinterpret_cast<uint4 *>(field + c0);
#pragma unroll
                        for (int t = 0; t < 8; t++) __stcg(f4 + t, inf4);
                    }
                }
            }
            __stcg(d4 + i, make_uint4(0, 0, 0, 0));
        }
    }
    __syncthreads();
}

// Warp 0 walks from (vx0, vy0) back to the source (sx, sy) of `field` along the arrival directions stored in the packed
// words and records turning points into tmp ((vx0, vy0) first, the source last).  Returns the number of points (may
// exceed cap: only cap are stored), -1 if the field is inconsistent (cannot happen after a successful pass);
// straight/diagonal step counts are ADDED to *na, *nb; *first_dir = arrival direction of (vx0, vy0) (8 for the source).
template <int METRIC>
__device__ int extract_path(const SearchParams &P, const uint32_t *__restrict__ field, int sx, int sy, int vx0, int vy0,
                            int32_t *__restrict__ tmp, int cap, unsigned *na, unsigned *nb, int *first_dir)
{
    const int H = P.H, W = P.W, TY = P.TY, lane = threadIdx.x & 31;
    int vx = vx0, vy = vy0;
    uint32_t v = __ldcg(field + fx_cidx(vx, vy, H, TY));
    int npts = 1, prev_d = -1;
    unsigned a = 0, b = 0;
    *first_dir = (int)(v & 15u);
    if (lane == 0 && cap > 0) { tmp[0] = vx; tmp[1] = vy; }
    const unsigned long long max_steps = (unsigned long long)W * H;
    unsigned long long steps = 0;
    while (!(vx == sx && vy == sy)) {
        const int d = (int)(v & 15u);
        if (v == FX_INF || d > 7 || steps > max_steps) return -1;
        if (prev_d >= 0 && d != prev_d) {  // v is a turning point
            if (lane == 0 && npts < cap) { tmp[2 * npts] = vx; tmp[2 * npts + 1] = vy; }
            npts++;
        }
        prev_d = d;
        // the parent of v is u_1 = v - dir(d); u_j keeps the run going while it arrived in direction d as well.
        // Lane j reads u_(j+1): 32 cells of a straight run per round trip.
        const int ddx = fx_dx(d), ddy = fx_dy(d);
        const int ux = vx - (lane + 1) * ddx, uy = vy - (lane + 1) * ddy;
        uint32_t uv = FX_INF;
        if (ux >= 0 && ux < W && uy >= 0 && uy < H) uv = __ldcg(field + fx_cidx(ux, uy, H, TY));
        const unsigned cont = __ballot_sync(0xFFFFFFFFu, uv != FX_INF && (int)(uv & 15u) == d);
        int lead = __ffs(~cont) - 1;  // u_1 .. u_lead arrived in direction d
        if (cont == 0xFFFFFFFFu) lead = 31;  // move to u_32 and look again
        const int run = lead + 1;
        v = __shfl_sync(0xFFFFFFFFu, uv, lead);
        vx -= run * ddx; vy -= run * ddy;
        steps += (unsigned)run;
        if (d < 4) a += run; else b += run;
    }
    if (!(vx0 == sx && vy0 == sy)) {
        if (lane == 0 && npts < cap) { tmp[2 * npts] = sx; tmp[2 * npts + 1] = sy; }
        npts++;
    }
    *na += a; *nb += b;
    return npts;
}

// THREADS / MINB: the throughput form (128 threads, 8 CTAs per SM hide each other's L2 round trips) and the latency
// form for batches smaller than the machine (FX_SEARCH_WIDE threads: a whole level of a single query -- a few hundred
// frontier cells -- is one round of loads instead of three).
// LAT (latency form): no band kernel, no pocket pre-pass.  The exact pass is bidirectional and its prune bound is a
// GUESS that is widened until it holds: a pass pruned with U0 is accepted iff it returns mu <= U0 -- then U0 was a real
// upper bound on the optimum, the pruning was sound and the termination proof of run_pass applies.  mu > U0 is still
// the cost of a real path, so the next pass with U0 = mu is final; no proposal at all widens the guess fourfold (about
// half of the queries of the headline workload are within 0.8 % of the octile lower bound, nearly all within 7 %).
// A pass in which one side exhausts its component without pruning and without meeting the other proves "unreachable"
// at the cost of the SMALLER component.
template <int METRIC, int THREADS, int MINB, bool LAT>
__global__ void __launch_bounds__(THREADS, MINB) k_search_batch(const SearchParams P)
{
    constexpr uint32_t WS = Wt<METRIC>::WS, WD = Wt<METRIC>::WD;
    __shared__ CtaState S;
    __shared__ uint8_t s_lut[9 * 256];
    __shared__ __align__(8) int s_step[8 * 64 * 2];  // [direction][x&7][y&7] -> (step of the packed xy, step of the tiled index)
    __shared__ int s_npts, s_seg[3];
    __shared__ unsigned s_ab[2];
    extern __shared__ __align__(16) unsigned char fx_dyn_smem[];
    uint2 *sq = reinterpret_cast<uint2 *>(fx_dyn_smem);  // latency form: heads of the bucket queues (run_pass, SQ)
    (void)sq;
    const int tid = threadIdx.x;
    const int slot = blockIdx.x;
    uint32_t *field = P.fields + (size_t)slot * P.cells * (LAT ? 2 : 1);  // LAT: [start side | goal side]
    uint8_t *dirty = P.dirty + (size_t)slot * P.dirty_n;
    uint2 *queue = P.queues + (size_t)slot * 4 * P.qcap;
    int32_t *tmp = P.tmp_path + (size_t)slot * P.path_cap * 4;  // two runs of path_cap points
    const int W = P.W, H = P.H;
    if (tid == 0) { S.settled = 0; S.levels = 0; S.flags = 0; S.xlo = 0x7FFFFFFF; S.xhi = -1; }
    for (int i = tid; i < 9 * 256; i += blockDim.x) s_lut[i] = (uint8_t)fx_canon_succ((unsigned)(i >> 8), (unsigned)(i & 255));
    for (int i = tid; i < 8 * 64; i += blockDim.x) {
        const int d = i >> 6, xi = (i >> 3) & 7, yi = i & 7, dx = fx_dx(d), dy = fx_dy(d);
        // stepping over a tile edge changes the tile part of the index by TY (x) or 1 (y) tiles and wraps the in-tile part
        const int tx = (xi + dx) >> 3, ty = (yi + dy) >> 3;  // -1, 0, +1 (arithmetic shift)
        s_step[2 * i] = dx * 65536 + dy;  // (x << 16 | y) + this == (x + dx) << 16 | (y + dy) for in-range children
        s_step[2 * i + 1] = (tx * P.TY + ty) * 64 + ((((xi + dx) & 7) - xi) << 3) + (((yi + dy) & 7) - yi);
    }
    unsigned long long passes = 0, band_only = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) { S.q = (int)atomicAdd(P.counters + 0, 1ull); S.flags = 0; }
        __syncthreads();
        if (S.q >= P.Q) break;
        const int q = P.order ? (int)P.order[S.q] : S.q;
        const int sx = P.starts[2 * q], sy = P.starts[2 * q + 1], gx = P.goals[2 * q], gy = P.goals[2 * q + 1];
        int32_t out_cost = FX_COST_UNREACHABLE;
        bool trivial = true;
        const bool s_in = sx >= 0 && sx < W && sy >= 0 && sy < H, g_in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        if (!s_in) out_cost = FX_COST_START_OOB;
        else if (!g_in) out_cost = FX_COST_UNREACHABLE;
        else if (sx == gx && sy == gy) out_cost = 0;                                       // jps1.py:199-208
        else if (P.grid[(size_t)gx * H + gy] == 1) out_cost = FX_COST_UNREACHABLE;         // jump() tests the cell first
        else if (P.moves[fx_cidx(sx, sy, H, P.TY)] == 0) out_cost = FX_COST_UNREACHABLE;        // start cannot move
        else {
            // can anything step INTO the goal?  (cheap rejection of sealed-off goals)
            bool any = false;
            for (int d = 0; d < 8; d++) {
                int ux = gx - fx_dx(d), uy = gy - fx_dy(d);
                if (ux >= 0 && ux < W && uy >= 0 && uy < H && ((P.moves[fx_cidx(ux, uy, H, P.TY)] >> d) & 1)) any = true;
            }
            if (any) trivial = false;
        }
        if (trivial) {
            if (tid == 0) {
                P.cost_i[q] = out_cost;
                if (P.cost_f) P.cost_f[q] = out_cost == 0 ? 0.0 : -1.0;
                if (P.path_len) P.path_len[q] = out_cost == 0 ? 1 : out_cost;
                if (out_cost == 0 && P.path_xy && P.max_path > 0) {
                    P.path_xy[(size_t)q * P.max_path * 2] = sx; P.path_xy[(size_t)q * P.max_path * 2 + 1] = sy;
                }
            }
            continue;
        }

        const uint32_t h0 = octile(abs(sx - gx), abs(sy - gy), WS, WD - WS);
        const float L = (float)max(abs(gx - sx), abs(gy - sy));
        uint32_t best = FX_INF;
        bool exact = false, overflow = false;
        bool hit = false, unreachable = false;
        bool bidir = false;
        if constexpr (LAT) {
            const int cbin = fx_calib_bin(abs(sx - gx), abs(sy - gy));
            uint64_t U_try = fx_first_bound(P.calib, cbin, h0, WD);
            for (int attempt = 0; attempt < 6; attempt++) {
                const bool last = attempt == 5 || U_try >= 0x7FFFFFFFull;
                const uint32_t U0 = last ? 0x7FFFFFFFu : (uint32_t)U_try;
                const uint32_t r = run_pass<METRIC, true, FX_LAT_SQ>(P, S, s_lut, s_step, field, dirty, queue, sq, sx, sy, gx, gy, U0, -1.f, 0u, &hit);
                passes++;
                bidir = true;
                overflow = (S.flags & FLAG_OVERFLOW) != 0;
                if (overflow) break;
                if (r != FX_INF && r <= U0) { best = r; if (tid == 0) fx_calib_update(P.calib, cbin, h0, r, WS); break; }  // the bound held: exact
                if ((S.flags & FLAG_UNREACH) || last || (r == FX_INF && !S.pruned)) { unreachable = true; break; }
                U_try = r != FX_INF ? (uint64_t)r : (uint64_t)h0 + (U_try - h0) * 4;  // a real path's cost / a wider guess
                __syncthreads();
                reset_slot(S, field, dirty, P.dirty_n, P.cells, H, true);
            }
        } else {
            // pocket check: a bounded flood FROM THE GOAL.  Among free cells the move graph is symmetric (a diagonal
            // tests the same two orthogonal cells both ways), so if the flood drains below the budget the goal sits in a
            // small sealed component: unless it met the start (or, for a start on an obstacle, a cell the start can
            // step into) the query is unreachable and the forward search need not flood the start's whole component.
            // If the flood reaches the start its cost is the exact answer and pass A is skipped.
            {
                uint32_t back = run_pass<METRIC, false, 0>(P, S, s_lut, s_step, field, dirty, queue, nullptr, gx, gy, sx, sy, 0x7FFFFFFFu, -1.f, FX_POCKET_BUDGET, &hit);
                passes++;
                overflow = (S.flags & FLAG_OVERFLOW) != 0;
                const bool start_free = P.grid[(size_t)sx * H + sy] != 1;
                if (!overflow && !hit) {
                    if (back != FX_INF && start_free) { best = back; }        // symmetric cost; pass B with U = best proves it
                    else if (back == FX_INF) {
                        bool touch = false;
                        if (!start_free) {
                            const unsigned ms = P.moves[fx_cidx(sx, sy, H, P.TY)];
                            for (int d = 0; d < 8; d++)
                                if (((ms >> d) & 1u) && __ldcg(field + fx_cidx(sx + fx_dx(d), sy + fx_dy(d), H, P.TY)) != FX_INF) touch = true;
                        }
                        unreachable = !touch;
                    }
                }
                __syncthreads();
                reset_slot(S, field, dirty, P.dirty_n, P.cells, H, false);
            }
            // upper bound from the warp-per-query band pass (band.cu).  If it equals the octile lower bound it is the
            // answer and only the path is still needed: one pass inside the same band with U = h0 recovers it (the band
            // of band.cu is a subset of this one, so the pass finds a path of that cost).  Otherwise it seeds pass B.
            const uint32_t hint = P.ubound ? P.ubound[q] : FX_INF;
            if (!overflow && !unreachable && best == FX_INF && hint != FX_INF) {
                if (hint == h0) {
                    best = run_pass<METRIC, false, 0>(P, S, s_lut, s_step, field, dirty, queue, nullptr, sx, sy, gx, gy, h0, (float)P.band0 * L, 0u, &hit);
                    passes++;
                    overflow = (S.flags & FLAG_OVERFLOW) != 0;
                    if (best != FX_INF) { exact = true; band_only++; }
                    else if (!overflow) reset_slot(S, field, dirty, P.dirty_n, P.cells, H, false);
                } else {
                    best = hint;
                }
            }
            // pass A (only without a usable hint): narrow band, generous bound; escalate if the band is sealed
            float band = (float)P.band0;
            uint32_t slack = h0 / 16 + 64 * WS;
            for (int attempt = 0; attempt < 3 && !overflow && !unreachable && best == FX_INF; attempt++) {
                const bool last = attempt == 2;
                uint64_t U64 = (uint64_t)h0 + slack;
                uint32_t U0 = (last || U64 > 0x7FFFFFFFull) ? 0x7FFFFFFFu : (uint32_t)U64;
                float bandL = last ? -1.f : band * L;
                best = run_pass<METRIC, false, 0>(P, S, s_lut, s_step, field, dirty, queue, nullptr, sx, sy, gx, gy, U0, bandL, 0u, &hit);
                passes++;
                overflow = (S.flags & FLAG_OVERFLOW) != 0;
                if (overflow) break;
                if (best != FX_INF) { exact = last || best == h0; if (attempt == 0 && exact) band_only++; break; }
                if (last || !S.pruned) break;  // nothing was pruned and the queue drained: the start's component is exhausted
                __syncthreads();
                reset_slot(S, field, dirty, P.dirty_n, P.cells, H, false);
                band *= 8.f; slack = slack * 4;
            }
            // pass B: no band, U = the upper bound -> exact
            if (!overflow && !unreachable && best != FX_INF && !exact) {
                __syncthreads();
                reset_slot(S, field, dirty, P.dirty_n, P.cells, H, false);
                best = run_pass<METRIC, false, 0>(P, S, s_lut, s_step, field, dirty, queue, nullptr, sx, sy, gx, gy, best, -1.f, 0u, &hit);
                passes++;
                overflow = (S.flags & FLAG_OVERFLOW) != 0;
            }
        }
        if (overflow || (best != FX_INF && best > 0x7FFFFFFFu)) {
            if (tid == 0) {
                P.cost_i[q] = FX_COST_OVERFLOW;
                if (P.cost_f) P.cost_f[q] = -3.0;
                if (P.path_len) P.path_len[q] = FX_COST_OVERFLOW;
            }
        } else if (best == FX_INF) {
            if (tid == 0) {
                P.cost_i[q] = FX_COST_UNREACHABLE;
                if (P.cost_f) P.cost_f[q] = -1.0;
                if (P.path_len) P.path_len[q] = FX_COST_UNREACHABLE;
            }
        } else {
            // tmp holds two runs of turning points: [0, cap) the walk from the meeting cell (or the goal) back to the
            // start, [cap, 2 cap) the walk from the meeting cell back to the goal through the goal side's field
            const int cap = P.path_cap;
            if (tid < 32) {
                unsigned a = 0, b = 0;
                int n1, n2 = 1, d1 = 8, d2 = 8;
                if (bidir) {
                    const int mx = (int)(S.meet >> 16), my = (int)(S.meet & 0xFFFFu);
                    n1 = extract_path<METRIC>(P, field, sx, sy, mx, my, tmp, cap, &a, &b, &d1);
                    n2 = extract_path<METRIC>(P, field + P.cells, gx, gy, mx, my, tmp + 2 * (size_t)cap, cap, &a, &b, &d2);
                } else {
                    n1 = extract_path<METRIC>(P, field, sx, sy, gx, gy, tmp, cap, &a, &b, &d1);
                }
                if (tid == 0) {
                    // the meeting cell is a turning point unless the travel direction into it (d1) equals the travel
                    // direction out of it (the opposite of the goal side's arrival direction d2)
                    const int opp2 = d2 < 4 ? (d2 ^ 1) : (d2 < 8 ? 11 - d2 : 8);
                    s_seg[0] = n1; s_seg[1] = n2;
                    s_seg[2] = (bidir && n1 > 1 && n2 > 1 && d1 == opp2) ? 1 : 0;
                    s_npts = (n1 < 0 || n2 < 0) ? -1 : n1 + n2 - 1 - s_seg[2];
                    s_ab[0] = a; s_ab[1] = b;
                }
            }
            __syncthreads();
            const int npts = s_npts;
            if (tid == 0) {
                P.cost_i[q] = npts < 0 ? FX_COST_OVERFLOW : (int32_t)best;
                if (P.cost_f)
                    P.cost_f[q] = METRIC == 1 ? (double)best : __dadd_rn((double)s_ab[0], __dmul_rn((double)s_ab[1], 1.4142135623730951));
                if (P.path_len) P.path_len[q] = npts < 0 ? FX_COST_OVERFLOW : npts;
            }
            if (P.path_xy && npts > 0 && npts <= P.max_path && s_seg[0] <= cap && s_seg[1] <= cap) {
                // output start..goal: the first run reversed (minus the meeting cell when it is collinear), then the second
                // run without its first point.  A path that does not fit (path_len > max_path) is not written at all.
                const int n1 = s_seg[0], drop = s_seg[2], nfirst = n1 - drop;
                int32_t *out = P.path_xy + (size_t)q * P.max_path * 2;
                for (int i = tid; i < npts; i += blockDim.x) {
                    const int32_t *src = i < nfirst ? tmp + 2 * (size_t)(n1 - 1 - i) : tmp + 2 * (size_t)cap + 2 * (size_t)(i - nfirst + 1);
                    out[2 * i] = __ldcg(src); out[2 * i + 1] = __ldcg(src + 1);
                }
            }
        }
        __syncthreads();
        reset_slot(S, field, dirty, P.dirty_n, P.cells, H, bidir);
    }
    if (tid == 0) {
        atomicAdd(P.counters + 1, S.settled);
        atomicAdd(P.counters + 2, S.levels);
        atomicAdd(P.counters + 3, passes);
        atomicAdd(P.counters + 4, band_only);
    }
}

// ------------------------------------------------------------------------------------------------
// cluster form of the latency kernel: one query per THREAD-BLOCK CLUSTER (FX_CL CTAs on FX_CL SMs)
// ------------------------------------------------------------------------------------------------
// One SM issues ~19 instructions per settled cell whatever the form (r02f: 18.7 M warp instructions, 5.4 ms, for one
// 983 k-cell query on one SM at 42 % issue utilisation), so a single query is bound by what ONE SM can issue.  Here the
// entries of a level are dealt over the FX_CL CTAs of a cluster.  A first version kept the wavefront state in rank 0's
// shared memory and let the other CTAs read it and reserve queue space through distributed shared memory: ~7 dependent
// remote round trips per level, no faster than one CTA (r02e: 3.3 us per level either way; r02f: 92 % of the issue slots
// idle, barrier + long-scoreboard stalls).  This version has NO remote read and NO remote atomic on the critical path:
//   * every CTA appends the children it creates to its OWN queue segment (local shared-memory tails, one atomic per lane
//     exactly like the one-CTA kernel);
//   * at the end of a level every CTA broadcasts, with fire-and-forget stores into the shared memory of all CTAs of the
//     cluster, the few words the others need: its segment's entry counts for the next two buckets, its best meeting
//     proposal, its alive / pruned / overflow bits; one cluster barrier (barrier.cluster arrive.release / wait.acquire)
//     publishes them, double-buffered by level parity;
//   * at the top of a level every thread derives the same decisions from its CTA's local copy, and the level's entries
//     -- the concatenation of the 2 * FX_CL (owner, class) runs -- are dealt out by global index: thread t of CTA r takes
//     index r * blockDim + t (+ rounds), found in the runs by a 4-step binary search over a per-warp prefix table.
// Same algorithm and results as run_pass<METRIC, true, *> (bidirectional; proof there).
#define FX_CL 8
#define FX_CL_THREADS 256

struct ClState {
    // local to this CTA
    unsigned tailS[4], tailD[4];    // entries this CTA appended to its segment of bucket slot b (straight / diagonal arrivals)
    unsigned alive[2][4];           // this CTA pushed an entry of side s into bucket slot b
    unsigned prl[2];                // this CTA pruned a cell of side s
    unsigned ovf;                   // this CTA saw a cost leave the 28-bit range / ran out of segment room
    unsigned long long mu[2];       // this CTA's best proposal of the level (slot = level parity)
    int xlo, xhi;                   // x-rows this CTA touched since the last reset
    unsigned long long settled, levels;
    // written by every CTA of the cluster (index = source rank), buffer = parity of the level that wrote it
    uint4 r_cnt[2][FX_CL];          // {nS, nD of the next bucket, nS, nD of the one after}
    unsigned long long r_mu[2][FX_CL];
    unsigned r_flags[2][FX_CL];     // bits 0..7 alive[s][b] (s * 4 + b), 8..9 prl, 10 overflow
    int r_x[2][FX_CL];              // pass end: xlo / xhi of every CTA
    int r_q;                        // the query index rank 0 fetched
    unsigned flags;                 // FLAG_* of the pass (every CTA derives the same value)
    unsigned pruned, meet;
};

template <int METRIC>
__device__ uint32_t run_pass_cluster(const SearchParams &P, ClState &L, ClState *const *peers, const uint8_t *__restrict__ s_lut,
                                     const int *__restrict__ s_step, uint32_t *__restrict__ field, uint8_t *__restrict__ dirty,
                                     uint2 *__restrict__ queue, int sx, int sy, int gx, int gy, uint32_t U)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr uint32_t WS = Wt<METRIC>::WS, WD = Wt<METRIC>::WD;
    const int H = P.H, TY = P.TY;
    const unsigned rank = cluster.block_rank();
    const unsigned tid = threadIdx.x, lane = tid & 31u, NTC = blockDim.x, NT = FX_CL * blockDim.x;
    const unsigned segcap = ((unsigned)P.qcap / (2u * FX_CL)) & ~1u;   // entries per (owner, class) run of a bucket
    const unsigned side_cells = (unsigned)P.cells;
    const unsigned sidx = (unsigned)fx_cidx(sx, sy, H, TY), gidx = (unsigned)fx_cidx(gx, gy, H, TY);
    const uint8_t *__restrict__ moves = P.moves;
    auto cell_of = [&](uint32_t exy) { return (unsigned)fx_cidx((int)((exy >> 16) & 0x7FFFu), (int)(exy & 0xFFFFu), H, TY); };
    auto foff_of = [&](uint32_t exy) { return (exy >> 31) * side_cells; };
    // run (owner o, class c) of bucket slot b starts here
    auto run_base = [&](unsigned b, unsigned o, unsigned c) { return queue + ((size_t)(b * FX_CL + o) * 2u + c) * segcap; };

    if (tid == 0) {
        for (int b = 0; b < 4; b++) { L.tailS[b] = 0; L.tailD[b] = 0; L.alive[0][b] = 0; L.alive[1][b] = 0; }
        L.prl[0] = 0; L.prl[1] = 0; L.ovf = 0; L.mu[0] = ~0ull; L.mu[1] = ~0ull;
        L.flags &= ~(FLAG_UNREACH);
        L.pruned = 0; L.meet = 0;
        // "level -1" (parity 1): the two seeds sit in rank 0's straight run of bucket 0, both sides alive in slot 0
        for (int r = 0; r < FX_CL; r++) {
            L.r_cnt[1][r] = r == 0 ? make_uint4(2u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
            L.r_mu[1][r] = ~0ull;
            L.r_flags[1][r] = r == 0 ? ((1u << 0) | (1u << 4)) : 0u;
        }
        if (rank == 0) {
            L.tailS[0] = 2; L.alive[0][0] = 1; L.alive[1][0] = 1;
            L.xlo = min(L.xlo, min(sx, gx) - 1); L.xhi = max(L.xhi, max(sx, gx) + 1);
            uint2 *q0 = run_base(0u, 0u, 0u);
            __stcg(q0, make_uint2(((uint32_t)sx << 16) | (uint32_t)sy, fx_pack(0u, FX_CODE_START)));
            __stcg(q0 + 1, make_uint2(0x80000000u | ((uint32_t)gx << 16) | (uint32_t)gy, fx_pack(0u, FX_CODE_START)));
            __stcg(field + sidx, fx_pack(0u, FX_CODE_START));
            __stcg(field + (side_cells + gidx), fx_pack(0u, FX_CODE_START));
            dirty[sidx >> FX_DIRTY_SHIFT] = 1;
            dirty[(side_cells + gidx) >> FX_DIRTY_SHIFT] = 1;
        }
    }
    cluster.sync();

    unsigned my_settled = 0;
    int my_xlo = 0x7FFFFFFF, my_xhi = -1;
    unsigned k = 0;
    uint32_t result = FX_INF;
    unsigned long long mu = ~0ull;
    unsigned prl_all = 0;
    const uint32_t h0 = octile(abs(sx - gx), abs(sy - gy), WS, WD - WS);
    const unsigned k_meet = h0 / (2u * WS) > 2u ? h0 / (2u * WS) - 2u : 0u;
    const bool start_free = P.grid[(size_t)sx * H + sy] != 1;
    unsigned exit_flags = 0;
    for (;;) {
        const unsigned pb = (k + 1u) & 1u;  // buffer written by level k-1
        // ---- the replicated words of level k-1: every thread of every CTA computes the same values from its local copy
        unsigned n = 0, n1 = 0, fl = 0;
        // per-warp prefix table of the 2 * FX_CL runs of bucket k (lane rho = owner * 2 + class holds the exclusive prefix)
        unsigned mylen = 0;
        if (lane < 2u * FX_CL) {
            const uint4 c = L.r_cnt[pb][lane >> 1];
            mylen = (lane & 1u) ? c.y : c.x;
        }
        unsigned incl = mylen;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (unsigned)o) incl += t; }
        const unsigned pre = incl - mylen;  // exclusive prefix (lanes >= 2 * FX_CL: the total)
        n = __shfl_sync(0xFFFFFFFFu, incl, 31);
#pragma unroll
        for (int r = 0; r < FX_CL; r++) {
            const uint4 c = L.r_cnt[pb][r];
            n1 += c.z + c.w;
            fl |= L.r_flags[pb][r];
            const unsigned long long m = L.r_mu[pb][r];
            mu = m < mu ? m : mu;
        }
        prl_all |= (fl >> 8) & 3u;
        const uint32_t mc = (uint32_t)(mu >> 32);
        if (mc != FX_INF && 2ull * k * WS > (unsigned long long)mc + WD) { result = mc; break; }
        if (fl & (1u << 10)) { exit_flags = FLAG_OVERFLOW; break; }
        if (mc == FX_INF) {
            // (see run_pass: a side with nothing in buckets k and k+1 that never pruned has exhausted its component)
            const unsigned a0 = (k & 3u), a1 = ((k + 1u) & 3u);
            const bool dead0 = !((fl >> a0) & 1u) && !((fl >> a1) & 1u), dead1 = !((fl >> (4u + a0)) & 1u) && !((fl >> (4u + a1)) & 1u);
            if ((dead0 && !(prl_all & 1u)) || (dead1 && !(prl_all & 2u) && (start_free || k >= 2u))) { exit_flags = FLAG_UNREACH; break; }
        }
        if (n == 0 && n1 == 0) { result = mc; break; }  // drained: the smallest proposal is exact (see run_pass)
        if (tid == 0) {
            // this CTA's slot of bucket k-1 will be bucket k+3: its local words start empty
            L.tailS[(k + 3) & 3] = 0; L.tailD[(k + 3) & 3] = 0; L.alive[0][(k + 3) & 3] = 0; L.alive[1][(k + 3) & 3] = 0;
        }
        const unsigned bk = k & 3u, b1 = (k + 1u) & 3u, b2 = (k + 2u) & 3u;
        const uint32_t kbase = k * WS;
        const bool meet_level = k >= k_meet;
        for (unsigned i0 = rank * NTC + (tid - lane); i0 < n; i0 += NT) {
            const unsigned i = i0 + lane;
            bool act = i < n;
            // which run holds global index i: 4-step binary search over the prefix table held by lanes 0 .. 2 FX_CL - 1
            unsigned rho = 0;
#pragma unroll
            for (unsigned step = FX_CL; step >= 1u; step >>= 1) {
                const unsigned p = __shfl_sync(0xFFFFFFFFu, pre, rho + step);
                if (i >= p) rho += step;
            }
            const unsigned rstart = __shfl_sync(0xFFFFFFFFu, pre, rho);
            const uint2 e = act ? __ldcg(run_base(bk, rho >> 1, rho & 1u) + (i - rstart)) : make_uint2(0u, 0u);
            const unsigned idx = cell_of(e.x), fo = foff_of(e.x);
            uint32_t v = FX_INF;
            unsigned m = 0;
            if (act) { v = __ldcg(field + (idx + fo)); m = (unsigned)__ldg(moves + idx); }
            const int x = (int)((e.x >> 16) & 0x7FFFu), y = (int)(e.x & 0xFFFFu);
            const bool side = (e.x >> 31) != 0u;
            const uint32_t g = e.y >> 4;
            act = act && v == e.y;
            uint32_t other = FX_INF;
            if (meet_level && act) other = __ldcg(field + (idx + (side ? 0u : side_cells)));
            if (act) {
                dirty[(idx + fo) >> FX_DIRTY_SHIFT] = 1;
                const int tx = side ? sx : gx, ty = side ? sy : gy;
                const uint32_t h = octile(abs(x - tx), abs(y - ty), WS, WD - WS);
                if (((uint64_t)g + h) > (uint64_t)U) { act = false; L.prl[side ? 1 : 0] = 1; }
            }
            if (other != FX_INF)
                atomicMin(&L.mu[k & 1u], ((unsigned long long)(g + (other >> 4)) << 32) | (unsigned long long)(e.x & 0x7FFFFFFFu));
            unsigned succ = 0;
            if (act) {
                my_settled++; my_xlo = min(my_xlo, x); my_xhi = max(my_xhi, x);
                succ = s_lut[((e.y & 15u) << 8) | m];
                if (g + WD > FX_COST_MAX28) { L.ovf = 1; succ = 0; }
            }
            const bool diag2 = (g - kbase) + WD >= 2u * WS;
            unsigned posS = 0, posD = 0;
            if (succ) {
                const unsigned ns = (unsigned)__popc(succ & 0x0Fu), nd = (unsigned)__popc(succ & 0xF0u);
                if (ns) posS = fx_atoms_add(&L.tailS[b1], ns);
                if (nd) posD = fx_atoms_add(&L.tailD[diag2 ? b2 : b1], nd);
                if (posS + ns > segcap || posD + nd > segcap) { L.ovf = 1; succ = 0; }
                else {
                    if (ns) L.alive[side ? 1 : 0][b1] = 1;
                    if (nd) L.alive[side ? 1 : 0][diag2 ? b2 : b1] = 1;
                }
            }
            uint2 *__restrict__ gS = run_base(b1, rank, 0u) + posS;
            uint2 *__restrict__ gD = run_base(diag2 ? b2 : b1, rank, 1u) + posD;
            const uint32_t nvS = fx_pack(g + WS, 0u), nvD = fx_pack(g + WD, 0u);
            const int *__restrict__ stp = s_step + 2 * (idx & 63);
            uint32_t *__restrict__ fbase = field + (idx + fo);
            while (succ) {
                const int d = __ffs(succ) - 1;
                succ &= succ - 1;
                const int2 st = *reinterpret_cast<const int2 *>(stp + 2 * 64 * d);
                const uint32_t nv = (d < 4 ? nvS : nvD) | (unsigned)d;
                fx_red_min(fbase + st.y, nv);
                const uint2 child = make_uint2(e.x + (uint32_t)st.x, nv);
                if (d < 4) __stcg(gS++, child);
                else __stcg(gD++, child);
            }
        }
        __syncthreads();  // this CTA's tails, flags and proposal of level k are complete
        // ---- broadcast: lane r of warp 0 writes this CTA's words into CTA r's copy (fire and forget)
        if (tid < FX_CL) {
            ClState *dst = peers[tid];
            const unsigned wb = k & 1u;
            dst->r_cnt[wb][rank] = make_uint4(L.tailS[b1], L.tailD[b1], L.tailS[b2], L.tailD[b2]);
            dst->r_mu[wb][rank] = L.mu[k & 1u];
            unsigned f = (L.prl[0] ? 1u << 8 : 0u) | (L.prl[1] ? 1u << 9 : 0u) | (L.ovf ? 1u << 10 : 0u);
#pragma unroll
            for (int b = 0; b < 4; b++) f |= (L.alive[0][b] ? 1u << b : 0u) | (L.alive[1][b] ? 1u << (4 + b) : 0u);
            dst->r_flags[wb][rank] = f;
        }
        if (tid == FX_CL) L.mu[(k + 1u) & 1u] = ~0ull;  // the other slot is free again: level k-1's proposal went out one level ago
        cluster.sync();
        k++;
    }
    // entries that were never popped (buckets k and k+1 of this CTA's own segment): mark their lines too
    for (unsigned b = k; b <= k + 1; b++) {
        const unsigned mS = min(L.tailS[b & 3], segcap), mD = min(L.tailD[b & 3], segcap);
        const uint2 *__restrict__ rS = run_base(b & 3u, rank, 0u), *__restrict__ rD = run_base(b & 3u, rank, 1u);
        for (unsigned i = tid; i < mS + mD; i += NTC) {
            const uint32_t xy = __ldcg(i < mS ? rS + i : rD + (i - mS)).x;
            dirty[(cell_of(xy) + foff_of(xy)) >> FX_DIRTY_SHIFT] = 1;
        }
    }
    {
        const int wlo = __reduce_min_sync(0xFFFFFFFFu, my_xlo), whi = __reduce_max_sync(0xFFFFFFFFu, my_xhi);
        const unsigned ws = __reduce_add_sync(0xFFFFFFFFu, my_settled);
        if (lane == 0) {
            if (whi >= 0) { atomicMin(&L.xlo, wlo - 1); atomicMax(&L.xhi, whi + 1); }
            if (ws) atomicAdd(&L.settled, (unsigned long long)ws);
        }
    }
    if (tid == 0) {
        L.flags |= exit_flags;
        L.pruned = prl_all;
        L.meet = (unsigned)(mu & 0x7FFFFFFFull);
        if (rank == 0) L.levels += k;
    }
    cluster.sync();
    return result;
}

// reset of the lines the query touched, by all threads of the cluster (see reset_slot); the x-range is the union of the
// CTAs' ranges, exchanged through each other's shared memory
__device__ void reset_slot_cluster(ClState &L, ClState *const *peers, uint32_t *__restrict__ field, uint8_t *__restrict__ dirty,
                                   size_t dirty_n, size_t cells, int H)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const unsigned ctid = rank * blockDim.x + threadIdx.x, NT = FX_CL * blockDim.x;
    __syncthreads();
    if (threadIdx.x < FX_CL) { peers[threadIdx.x]->r_x[0][rank] = L.xlo; peers[threadIdx.x]->r_x[1][rank] = L.xhi; }
    cluster.sync();
    int xlo = 0x7FFFFFFF, xhi = -1;
#pragma unroll
    for (int r = 0; r < FX_CL; r++) { xlo = min(xlo, L.r_x[0][r]); xhi = max(xhi, L.r_x[1][r]); }
    xlo = max(xlo, 0);
    const bool force = (L.flags & FLAG_OVERFLOW) != 0;
    __syncthreads();
    if (threadIdx.x == 0) { L.xlo = 0x7FFFFFFF; L.xhi = -1; }
    if (xhi >= xlo) {
        const size_t c_lo = (size_t)(xlo >> 3) * fx_tiles_y(H) * 64;
        size_t c_hi = (size_t)((xhi >> 3) + 1) * fx_tiles_y(H) * 64;
        if (c_hi > cells) c_hi = cells;
        const size_t n16 = dirty_n / 16;
        uint4 *d4 = reinterpret_cast<uint4 *>(dirty);
        const uint4 inf4 = make_uint4(FX_INF, FX_INF, FX_INF, FX_INF);
        for (int f = 0; f < 2; f++) {
            const size_t base = (size_t)f * cells;
            size_t i0 = ((base + c_lo) >> FX_DIRTY_SHIFT) / 16, i1 = (((base + c_hi) >> FX_DIRTY_SHIFT) + 15) / 16;
            if (i1 > n16) i1 = n16;
            for (size_t i = i0 + ctid; i < i1; i += NT) {
                uint4 v = __ldcg(d4 + i);
                if (force) v = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
                if ((v.x | v.y | v.z | v.w) == 0u) continue;
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    if (!w[a]) continue;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        if (!((w[a] >> (8 * b)) & 0xFFu)) continue;
                        const size_t c0 = (i * 16 + a * 4 + b) << FX_DIRTY_SHIFT;
                        if (c0 + 32 <= 2 * cells) {
                            uint4 *f4 = reinterpret_cast<uint4 *>(field + c0);
#pragma unroll
                            for (int t = 0; t < 8; t++) __stcg(f4 + t, inf4);
                        }
                    }
                }
                __stcg(d4 + i, make_uint4(0, 0, 0, 0));
            }
        }
    }
    cluster.sync();
}

template <int METRIC>
__global__ void __launch_bounds__(FX_CL_THREADS, 1) k_search_cluster(const SearchParams P)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr uint32_t WS = Wt<METRIC>::WS, WD = Wt<METRIC>::WD;
    __shared__ ClState L;
    __shared__ ClState *peers[FX_CL];
    __shared__ uint8_t s_lut[9 * 256];
    __shared__ __align__(8) int s_step[8 * 64 * 2];
    __shared__ int s_out[6];  // rank 0: npts, n1, n2, drop, a, b of the extracted path
    const unsigned rank = cluster.block_rank();
    const int tid = threadIdx.x;
    const unsigned ctid = rank * blockDim.x + tid, NT = FX_CL * blockDim.x;
    const int slot = blockIdx.x / FX_CL;
    uint32_t *field = P.fields + (size_t)slot * P.cells * 2;
    uint8_t *dirty = P.dirty + (size_t)slot * P.dirty_n;
    uint2 *queue = P.queues + (size_t)slot * 4 * P.qcap;
    int32_t *tmp = P.tmp_path + (size_t)slot * P.path_cap * 4;
    const int W = P.W, H = P.H;
    if (tid < FX_CL) peers[tid] = cluster.map_shared_rank(&L, tid);
    if (tid == 0) { L.settled = 0; L.levels = 0; L.flags = 0; L.xlo = 0x7FFFFFFF; L.xhi = -1; }
    for (int i = tid; i < 9 * 256; i += blockDim.x) s_lut[i] = (uint8_t)fx_canon_succ((unsigned)(i >> 8), (unsigned)(i & 255));
    for (int i = tid; i < 8 * 64; i += blockDim.x) {
        const int d = i >> 6, xi = (i >> 3) & 7, yi = i & 7, dx = fx_dx(d), dy = fx_dy(d);
        const int tx = (xi + dx) >> 3, ty = (yi + dy) >> 3;
        s_step[2 * i] = dx * 65536 + dy;
        s_step[2 * i + 1] = (tx * P.TY + ty) * 64 + ((((xi + dx) & 7) - xi) << 3) + (((yi + dy) & 7) - yi);
    }
    __syncthreads();
    int *r_out = cluster.map_shared_rank(s_out, 0);
    unsigned long long passes = 0;
    for (;;) {
        cluster.sync();
        if (rank == 0 && tid < FX_CL) {
            int q = 0;
            if (tid == 0) q = (int)atomicAdd(P.counters + 0, 1ull);
            q = __shfl_sync((1u << FX_CL) - 1u, q, 0);
            peers[tid]->r_q = q;
        }
        if (tid == 0) L.flags = 0;
        cluster.sync();
        const int q = L.r_q;
        if (q >= P.Q) break;
        const int sx = P.starts[2 * q], sy = P.starts[2 * q + 1], gx = P.goals[2 * q], gy = P.goals[2 * q + 1];
        int32_t out_cost = FX_COST_UNREACHABLE;
        bool trivial = true;
        const bool s_in = sx >= 0 && sx < W && sy >= 0 && sy < H, g_in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        if (!s_in) out_cost = FX_COST_START_OOB;
        else if (!g_in) out_cost = FX_COST_UNREACHABLE;
        else if (sx == gx && sy == gy) out_cost = 0;
        else if (P.grid[(size_t)gx * H + gy] == 1) out_cost = FX_COST_UNREACHABLE;
        else if (P.moves[fx_cidx(sx, sy, H, P.TY)] == 0) out_cost = FX_COST_UNREACHABLE;
        else {
            bool any = false;
            for (int d = 0; d < 8; d++) {
                const int ux = gx - fx_dx(d), uy = gy - fx_dy(d);
                if (ux >= 0 && ux < W && uy >= 0 && uy < H && ((P.moves[fx_cidx(ux, uy, H, P.TY)] >> d) & 1)) any = true;
            }
            if (any) trivial = false;
        }
        if (trivial) {
            if (ctid == 0) {
                P.cost_i[q] = out_cost;
                if (P.cost_f) P.cost_f[q] = out_cost == 0 ? 0.0 : -1.0;
                if (P.path_len) P.path_len[q] = out_cost == 0 ? 1 : out_cost;
                if (out_cost == 0 && P.path_xy && P.max_path > 0) {
                    P.path_xy[(size_t)q * P.max_path * 2] = sx; P.path_xy[(size_t)q * P.max_path * 2 + 1] = sy;
                }
            }
            continue;
        }
        const uint32_t h0 = octile(abs(sx - gx), abs(sy - gy), WS, WD - WS);
        uint32_t best = FX_INF;
        bool overflow = false;
        const int cbin = fx_calib_bin(abs(sx - gx), abs(sy - gy));
        uint64_t U_try = fx_first_bound(P.calib, cbin, h0, WD);
        for (int attempt = 0; attempt < 6; attempt++) {
            const bool last = attempt == 5 || U_try >= 0x7FFFFFFFull;
            const uint32_t U0 = last ? 0x7FFFFFFFu : (uint32_t)U_try;
            const uint32_t r = run_pass_cluster<METRIC>(P, L, peers, s_lut, s_step, field, dirty, queue, sx, sy, gx, gy, U0);
            passes++;
            const unsigned fl = L.flags, pruned = L.pruned;  // identical in every CTA (derived from the replicated words)
            overflow = (fl & FLAG_OVERFLOW) != 0;
            if (overflow) break;
            if (r != FX_INF && r <= U0) { best = r; if (ctid == 0) fx_calib_update(P.calib, cbin, h0, r, WS); break; }
            if ((fl & FLAG_UNREACH) || last || (r == FX_INF && !pruned)) break;
            U_try = r != FX_INF ? (uint64_t)r : (uint64_t)h0 + (U_try - h0) * 4;
            reset_slot_cluster(L, peers, field, dirty, P.dirty_n, P.cells, H);
        }
        if (overflow || (best != FX_INF && best > 0x7FFFFFFFu)) {
            if (ctid == 0) {
                P.cost_i[q] = FX_COST_OVERFLOW;
                if (P.cost_f) P.cost_f[q] = -3.0;
                if (P.path_len) P.path_len[q] = FX_COST_OVERFLOW;
            }
        } else if (best == FX_INF) {
            if (ctid == 0) {
                P.cost_i[q] = FX_COST_UNREACHABLE;
                if (P.cost_f) P.cost_f[q] = -1.0;
                if (P.path_len) P.path_len[q] = FX_COST_UNREACHABLE;
            }
        } else {
            const int cap = P.path_cap;
            if (rank == 0 && tid < 32) {
                unsigned a = 0, b = 0;
                int d1 = 8, d2 = 8;
                const int mx = (int)(L.meet >> 16), my = (int)(L.meet & 0xFFFFu);
                const int n1 = extract_path<METRIC>(P, field, sx, sy, mx, my, tmp, cap, &a, &b, &d1);
                const int n2 = extract_path<METRIC>(P, field + P.cells, gx, gy, mx, my, tmp + 2 * (size_t)cap, cap, &a, &b, &d2);
                if (tid == 0) {
                    const int opp2 = d2 < 4 ? (d2 ^ 1) : (d2 < 8 ? 11 - d2 : 8);
                    const int drop = (n1 > 1 && n2 > 1 && d1 == opp2) ? 1 : 0;
                    s_out[0] = (n1 < 0 || n2 < 0) ? -1 : n1 + n2 - 1 - drop;
                    s_out[1] = n1; s_out[2] = n2; s_out[3] = drop; s_out[4] = (int)a; s_out[5] = (int)b;
                }
            }
            cluster.sync();
            const int npts = r_out[0], n1 = r_out[1], n2 = r_out[2], drop = r_out[3];
            if (ctid == 0) {
                P.cost_i[q] = npts < 0 ? FX_COST_OVERFLOW : (int32_t)best;
                if (P.cost_f)
                    P.cost_f[q] = METRIC == 1 ? (double)best : __dadd_rn((double)(unsigned)r_out[4], __dmul_rn((double)(unsigned)r_out[5], 1.4142135623730951));
                if (P.path_len) P.path_len[q] = npts < 0 ? FX_COST_OVERFLOW : npts;
            }
            if (P.path_xy && npts > 0 && npts <= P.max_path && n1 <= cap && n2 <= cap) {
                const int nfirst = n1 - drop;
                int32_t *out = P.path_xy + (size_t)q * P.max_path * 2;
                for (int i = (int)ctid; i < npts; i += (int)NT) {
                    const int32_t *src = i < nfirst ? tmp + 2 * (size_t)(n1 - 1 - i) : tmp + 2 * (size_t)cap + 2 * (size_t)(i - nfirst + 1);
                    out[2 * i] = __ldcg(src); out[2 * i + 1] = __ldcg(src + 1);
                }
            }
        }
        reset_slot_cluster(L, peers, field, dirty, P.dirty_n, P.cells, H);
    }
    cluster.sync();  // nobody leaves while its shared memory may still be written or read by a peer
    if (tid == 0) {
        atomicAdd(P.counters + 1, L.settled);
        if (rank == 0) { atomicAdd(P.counters + 2, L.levels); atomicAdd(P.counters + 3, passes); }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void fx_search_release(fx_context *ctx, int which)
{
    fx_context::SearchScratch &S = ctx->scr[which];
    if (S.fields) cudaFree(S.fields);
    if (S.dirty) cudaFree(S.dirty);
    if (S.queues) cudaFree(S.queues);
    if (S.tmp_path) cudaFree(S.tmp_path);
    memset(&S, 0, sizeof(S));
}

int fx_search_reserve(fx_context *ctx, int which, int W, int H, int max_path, cudaStream_t st)
{
    fx_context::SearchScratch &S = ctx->scr[which];
    size_t cells = fx_scratch_cells(W, H);
    int path_cap = max_path > 0 ? max_path : 1;
    if (S.fields && S.sW == W && S.sH == H && S.path_cap >= path_cap) return FX_OK;
    if (S.fields && S.sW == W && S.sH == H) {
        // only the path staging is too small (a path with more turning points than any before): grow that alone -- the
        // cost fields of the latency forms are 20 GB at 4096^2, freeing and refilling them cost a 25 ms outlier per retry
        FX_CUDA(ctx, cudaStreamSynchronize(st));
        if (S.tmp_path) cudaFree(S.tmp_path);
        S.tmp_path = nullptr;
        FX_CUDA(ctx, cudaMalloc(&S.tmp_path, (size_t)S.slots * path_cap * 16));
        S.path_cap = path_cap;
        return FX_OK;
    }
    fx_search_release(ctx, which);

    // cells padded to a multiple of 512: field rows stay 16-byte aligned for the uint4 reset and a second field's dirty
    // flags start on a uint4
    const int nfields = which == 1 ? 2 : 1;
    const size_t cells_al = (cells + 511) / 512 * 512;
    size_t dirty_n = (nfields * cells_al) >> FX_DIRTY_SHIFT;  // a multiple of 16
    // entries per bucket queue.  The latency forms get eight times the room: there are few slots, and the cluster form cuts
    // a bucket half into one segment per CTA (1 / 16 of qcap), which a wide first bound on a small grid can fill from a
    // single CTA (r02j fuzz: a 121 x 42 grid, effectively unpruned, overflowed 145-entry segments)
    int qcap = which == 1 ? 64 * (W + H) + 8192 : 8 * (W + H) + 1024;
    size_t per_slot = nfields * cells_al * 4 + dirty_n + (size_t)qcap * 32 + (size_t)path_cap * 16;
    size_t free_b = 0, total_b = 0;
    FX_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    int slots = which == 1 ? ctx->sm_count : (ctx->cfg_slots > 0 ? ctx->cfg_slots : ctx->sm_count * FX_SEARCH_MINB);
    size_t budget = free_b / 2;  // leave half of what is free to the caller
    if ((size_t)slots * per_slot > budget) slots = (int)(budget / per_slot);
    if (slots < 1) return fx_set_err(ctx, FX_ERR_NOMEM, "search scratch for a %dx%d grid does not fit (%zu B per slot, %zu free)", W, H, per_slot, free_b);
    FX_CUDA(ctx, cudaMalloc(&S.fields, (size_t)slots * nfields * cells_al * 4));
    FX_CUDA(ctx, cudaMalloc(&S.dirty, (size_t)slots * dirty_n));
    FX_CUDA(ctx, cudaMalloc(&S.queues, (size_t)slots * 4 * qcap * sizeof(uint2)));
    FX_CUDA(ctx, cudaMalloc(&S.tmp_path, (size_t)slots * path_cap * 16));
    // on the LAUNCH stream: the synchronous-API memset runs on the legacy default stream, which a cudaStreamNonBlocking
    // stream (the context's own stream of the *_host entry points) does not wait for
    FX_CUDA(ctx, cudaMemsetAsync(S.fields, 0xFF, (size_t)slots * nfields * cells_al * 4, st));
    FX_CUDA(ctx, cudaMemsetAsync(S.dirty, 0, (size_t)slots * dirty_n, st));
    S.sW = W; S.sH = H; S.slots = slots; S.qcap = qcap; S.path_cap = path_cap; S.nfields = nfields;
    S.cells = cells_al; S.dirty_n = dirty_n;
    return FX_OK;
}

extern "C" int fx_search_batch(fx_context *ctx, const uint8_t *grid, int W, int H, const int32_t *starts_xy,
                               const int32_t *goals_xy, int Q, int metric, int32_t *cost_i, double *cost_f,
                               int32_t *path_xy, int32_t *path_len, int max_path, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!grid || W <= 0 || H <= 0 || Q < 0 || (Q > 0 && (!starts_xy || !goals_xy || !cost_i)) || max_path < 0 ||
        (metric != 1 && metric != 2))
        return fx_set_err(ctx, FX_ERR_ARG, "fx_search_batch: bad argument");
    if (W > 32767 || H > 32767) return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_search_batch: W,H must be <= 32767");
    if (Q == 0) return FX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    {
        // maps that fit one SM's shared memory (every map the reference ships): one launch, no global scratch.
        // FUXI_B200_SMALL=0 routes them through the batched kernel instead (tests cover both forms on the same maps).
        const char *e = getenv("FUXI_B200_SMALL");
        if (!(e && e[0] == '0')) {
            const int r = fx_search_small(ctx, grid, W, H, starts_xy, goals_xy, Q, metric, cost_i, cost_f, path_xy, path_len, max_path, st);
            if (r <= 0) return r;
        }
    }
    // batches of at most this many queries run the latency form (one wide CTA per query, bidirectional, no band kernel)
    const int wide = ctx->cfg_wide_below >= 0 ? ctx->cfg_wide_below : ctx->sm_count;
    const int which = Q <= wide ? 1 : 0;
    int rc = fx_search_reserve(ctx, which, W, H, path_xy ? max_path : 1, st);
    if (rc) return rc;
    if (ctx->moves_prebuilt_for == grid) ctx->moves_prebuilt_for = nullptr;  // fx_plan_host built the mask behind its upload
    else {
        rc = fx_build_moves(ctx, grid, W, H, true, st);
        if (rc) return rc;
    }
    FX_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, 16 * sizeof(unsigned long long), st));
    const fx_context::SearchScratch &X = ctx->scr[which];
    SearchParams P;
    P.grid = grid; P.moves = ctx->moves; P.W = W; P.H = H; P.TY = fx_tiles_y(H);
    P.starts = starts_xy; P.goals = goals_xy; P.Q = Q;
    P.cost_i = cost_i; P.cost_f = cost_f; P.path_xy = path_xy; P.path_len = path_len;
    P.max_path = path_xy ? max_path : 0;
    P.fields = X.fields; P.dirty = X.dirty; P.queues = reinterpret_cast<uint2 *>(X.queues); P.tmp_path = X.tmp_path;
    P.cells = X.cells; P.dirty_n = X.dirty_n; P.qcap = X.qcap; P.path_cap = X.path_cap;
    P.counters = ctx->counters;
    {
        const char *e = getenv("FUXI_B200_CALIB");  // tuning experiments only
        P.calib = (e && e[0] == '0') ? nullptr : reinterpret_cast<uint32_t *>(ctx->counters + 16);
        // The table describes the maps this context plans on.  It survives a change of the grid's shape (a planner's grid
        // grows and shifts from replan to replan while its obstacle statistics stay: the decaying maximum follows them,
        // and a stale value only makes a first pass wider or a second one necessary); another metric starts it afresh.
        if (P.calib && ctx->calib_metric != metric) {
            FX_CUDA(ctx, cudaMemsetAsync(P.calib, 0, 16 * sizeof(uint32_t), st));
            ctx->calib_metric = metric;
        }
        ctx->calib_W = W; ctx->calib_H = H;
    }
    // half-width (cells along the minor axis) of the band-limited passes of the search kernel; one more than the band kernel's
    // 15 + rounding of its fixed-point centre line, so that a path the band kernel found lies inside it
    P.band0 = ctx->cfg_band0 > 0 ? ctx->cfg_band0 : 17;
    P.order = nullptr; P.ubound = nullptr;
    FX_CUDA(ctx, cudaEventRecord(ctx->ev_band[0], st));
    if (which == 0) {
        rc = fx_band_bounds(ctx, grid, W, H, starts_xy, goals_xy, Q, metric, st);
        if (rc) return rc;
        P.order = ctx->q_order; P.ubound = ctx->q_ubound;
    } else {
        FX_CUDA(ctx, cudaEventRecord(ctx->ev_band[1], st));
    }
    int blocks = X.slots < Q ? X.slots : Q;
    FX_CUDA(ctx, cudaEventRecord(ctx->ev_search[0], st));
    const bool use_cluster = which == 1 && ctx->cfg_cluster && Q <= ctx->sm_count / FX_CL;
    if (use_cluster) {
        // at most one query per cluster of FX_CL SMs: the entries of a level are dealt over FX_CL SMs
        cudaLaunchConfig_t cfg = {};
        const int nclusters = Q < X.slots ? Q : X.slots;
        cfg.gridDim = dim3((unsigned)(nclusters * FX_CL), 1, 1);
        cfg.blockDim = dim3(FX_CL_THREADS, 1, 1);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = FX_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (metric == 1) FX_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_search_cluster<1>, P));
        else FX_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_search_cluster<2>, P));
    } else if (which == 1) {
        const size_t sm = (size_t)4 * 2 * FX_LAT_SQ * sizeof(uint2);
        if (!ctx->lat_attr_set) {
            FX_CUDA(ctx, cudaFuncSetAttribute(k_search_batch<1, FX_SEARCH_WIDE, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            FX_CUDA(ctx, cudaFuncSetAttribute(k_search_batch<2, FX_SEARCH_WIDE, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            ctx->lat_attr_set = 1;
        }
        if (metric == 1) k_search_batch<1, FX_SEARCH_WIDE, 1, true><<<blocks, FX_SEARCH_WIDE, sm, st>>>(P);
        else k_search_batch<2, FX_SEARCH_WIDE, 1, true><<<blocks, FX_SEARCH_WIDE, sm, st>>>(P);
    } else {
        if (metric == 1) k_search_batch<1, FX_SEARCH_THREADS, FX_SEARCH_MINB, false><<<blocks, FX_SEARCH_THREADS, 0, st>>>(P);
        else k_search_batch<2, FX_SEARCH_THREADS, FX_SEARCH_MINB, false><<<blocks, FX_SEARCH_THREADS, 0, st>>>(P);
    }
    FX_LAUNCH_CHECK(ctx);
    FX_CUDA(ctx, cudaEventRecord(ctx->ev_search[1], st));
    ctx->ev_search_valid = 1;
    return FX_OK;
}

extern "C" int fx_canon_successors(int code, int moves)
{
    if (code < 0 || code > 8 || moves < 0 || moves > 255) return FX_ERR_ARG;
    return (int)fx_canon_succ((unsigned)code, (unsigned)moves);
}

/* tuning builds only (-DFX_PHASE_CLOCKS): cycles per dependent step, summed over warp 0 of every CTA */
extern "C" int fx_search_phase_clocks(fx_context *ctx, int64_t *h_8)
{
    if (!ctx || !h_8) return FX_ERR_ARG;
    unsigned long long c[16];
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    FX_CUDA(ctx, cudaDeviceSynchronize());
    FX_CUDA(ctx, cudaMemcpy(c, ctx->counters, sizeof(c), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 8; i++) h_8[i] = (int64_t)c[8 + i];
    return FX_OK;
}

/* duration of the last k_search_batch launch alone (CUDA events recorded on the launching stream around it) */
extern "C" int fx_search_kernel_ms(fx_context *ctx, float *h_ms)
{
    if (!ctx || !h_ms) return FX_ERR_ARG;
    if (!ctx->ev_search_valid) return fx_set_err(ctx, FX_ERR_ARG, "fx_search_kernel_ms: no search has run on this context");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    FX_CUDA(ctx, cudaEventSynchronize(ctx->ev_search[1]));
    FX_CUDA(ctx, cudaEventElapsedTime(h_ms, ctx->ev_search[0], ctx->ev_search[1]));
    return FX_OK;
}

/* {k_band_bound, k_search_batch} durations of the last batch that ran the batched kernel */
extern "C" int fx_search_timings(fx_context *ctx, float *h_ms2)
{
    if (!ctx || !h_ms2) return FX_ERR_ARG;
    if (!ctx->ev_search_valid) return fx_set_err(ctx, FX_ERR_ARG, "fx_search_timings: no search has run on this context");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    FX_CUDA(ctx, cudaEventSynchronize(ctx->ev_search[1]));
    FX_CUDA(ctx, cudaEventElapsedTime(h_ms2 + 0, ctx->ev_band[0], ctx->ev_band[1]));
    FX_CUDA(ctx, cudaEventElapsedTime(h_ms2 + 1, ctx->ev_search[0], ctx->ev_search[1]));
    return FX_OK;
}

extern "C" int fx_search_stats(fx_context *ctx, int64_t *h_stats4)
{
    if (!ctx || !h_stats4) return FX_ERR_ARG;
    unsigned long long c[8];
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    FX_CUDA(ctx, cudaDeviceSynchronize());
    FX_CUDA(ctx, cudaMemcpy(c, ctx->counters, sizeof(c), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; i++) h_stats4[i] = (int64_t)c[i + 1];
    return FX_OK;
}
