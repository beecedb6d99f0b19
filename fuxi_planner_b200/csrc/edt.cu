// edt.cu -- exact squared Euclidean distance transform of the occupancy grid (sm_100a).
//
// Not in the reference (its inflation is the square dilation in inflate.cu); added by the north star
// as the exact-clearance counterpart.  Oracle: scipy.ndimage.distance_transform_edt.
// Separable: pass 1 along y (the contiguous axis) gives g(x,y) = distance to the nearest occupied
// cell in the same row (uint16, 0xFFFF = none); pass 2 along x takes min over x' of (x-x')^2 + g(x',y)^2.
//   pass 1: one warp per row; the row is squeezed to a bit mask in shared memory (H/8 bytes), nearest
//           set bits come from clz/ffs on the mask words plus a word-level prefix.
//   pass 2a (windowed, fully parallel): only rows with |x-x'| < g(x,y) can win, so each cell scans a
//           window that closes as soon as d^2 >= best.  Cells that have not closed within EDT_WIN rows
//           raise a flag ...
//   pass 2b (Meijster lower envelope, one thread per column, coalesced across columns) ... which makes
//           this kernel redo the whole grid exactly.  It returns immediately when the flag is clear.
#include "common.cuh"

#define EDT_WIN 48
#define EDT_NONE 0xFFFFu

// ---- pass 1 -------------------------------------------------------------------------------------
// one warp per row; smem per warp: (H+31)/32 mask words + 2 * that for the prefix arrays
__global__ void __launch_bounds__(256) k_edt_rows(const uint8_t *__restrict__ occ, uint16_t *__restrict__ g, int W, int H, const int *__restrict__ run_flag)
{
    if (run_flag && *run_flag == 0) return;  // fallback of the bit-parallel path: only when it gave up
    extern __shared__ unsigned sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nwords = (H + 31) / 32;
    unsigned *mask = sm + (size_t)warp * 3 * nwords;
    int *leftp = reinterpret_cast<int *>(mask + nwords);   // nearest occupied y strictly before word w (or -1)
    int *rightp = leftp + nwords;                          // nearest occupied y strictly after word w (or -1)
    const bool vec = (H % 16 == 0) && (((uintptr_t)occ & 15u) == 0);
    for (int x = blockIdx.x * nwarps + warp; x < W; x += gridDim.x * nwarps) {
        const uint8_t *row = occ + (size_t)x * H;
        // build the bit mask
        if (vec) {
            const int nchunks = H / 16;  // 16 cells per lane-load
            for (int c0 = 0; c0 < nchunks; c0 += 32) {
                const int c = c0 + lane;
                unsigned b = 0;
                if (c < nchunks) {
                    uint4 v = __ldcs(reinterpret_cast<const uint4 *>(row) + c);
                    unsigned w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        unsigned nz = __vcmpne4(w4[i], 0u) & 0x01010101u;
                        b |= (((nz * 0x01020408u) >> 24) & 0xFu) << (4 * i);
                    }
                }
                // two 16-bit chunks -> one 32-bit word
                unsigned hi = __shfl_down_sync(0xFFFFFFFFu, b, 1);
                if (!(lane & 1) && c < nchunks) mask[c >> 1] = b | (hi << 16);
            }
        } else {
            for (int w = lane; w < nwords; w += 32) {
                unsigned m = 0;
                for (int j = 0; j < 32; j++) {
                    int y = w * 32 + j;
                    if (y < H && row[y]) m |= 1u << j;
                }
                mask[w] = m;
            }
        }
        __syncwarp();
        // word-level prefix: last occupied y before word w / first occupied y after word w
        int carry = -1;
        for (int w0 = 0; w0 < nwords; w0 += 32) {
            const int w = w0 + lane;
            unsigned m = w < nwords ? mask[w] : 0u;
            int mine = m ? (w * 32 + 31 - __clz(m)) : -1;  // highest set bit position
            int incl = mine;
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl = max(incl, t); }
            int excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
            if (lane == 0) excl = -1;
            excl = max(excl, carry);
            if (w < nwords) leftp[w] = excl;
            carry = max(carry, __shfl_sync(0xFFFFFFFFu, incl, 31));
        }
        carry = 0x7FFFFFFF;
        for (int w0 = ((nwords - 1) / 32) * 32; w0 >= 0; w0 -= 32) {
            const int w = w0 + lane;
            unsigned m = w < nwords ? mask[w] : 0u;
            int mine = m ? (w * 32 + __ffs(m) - 1) : 0x7FFFFFFF;  // lowest set bit position
            int incl = mine;
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_down_sync(0xFFFFFFFFu, incl, o); if (lane + o < 32) incl = min(incl, t); }
            int excl = __shfl_down_sync(0xFFFFFFFFu, incl, 1);
            if (lane == 31) excl = 0x7FFFFFFF;
            excl = min(excl, carry);
            if (w < nwords) rightp[w] = excl == 0x7FFFFFFF ? -1 : excl;
            carry = min(carry, __shfl_sync(0xFFFFFFFFu, incl, 0));
        }
        __syncwarp();
        // distances: each lane owns 16 consecutive cells (32-byte store)
        uint16_t *grow = g + (size_t)x * H;
        for (int c0 = 0; c0 * 16 < H; c0 += 32) {
            const int c = c0 + lane, ybase = c * 16;
            if (ybase >= H) continue;
            const int w = ybase >> 5, sh = ybase & 31;
            const unsigned m = mask[w];
            const int lp = leftp[w], rp = rightp[w];
            unsigned short d16[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int bit = sh + j, y = ybase + j;
                const unsigned below = m & (0xFFFFFFFFu >> (31 - bit));   // bits <= bit
                const unsigned above = m >> bit;                          // bits >= bit (shifted)
                int dl = below ? bit - (31 - __clz(below)) : (lp >= 0 ? y - lp : 0x7FFFFFFF);
                int dr = above ? __ffs(above) - 1 : (rp >= 0 ? rp - y : 0x7FFFFFFF);
                int dd = min(dl, dr);
                d16[j] = (unsigned short)(dd > 0xFFFE ? EDT_NONE : dd);
            }
            if (ybase + 16 <= H && vec) {
                uint4 lo, hi;
                lo.x = d16[0] | (d16[1] << 16); lo.y = d16[2] | (d16[3] << 16); lo.z = d16[4] | (d16[5] << 16); lo.w = d16[6] | (d16[7] << 16);
                hi.x = d16[8] | (d16[9] << 16); hi.y = d16[10] | (d16[11] << 16); hi.z = d16[12] | (d16[13] << 16); hi.w = d16[14] | (d16[15] << 16);
                uint4 *dst = reinterpret_cast<uint4 *>(grow + ybase);
                dst[0] = lo; dst[1] = hi;
            } else {
                for (int j = 0; j < 16 && ybase + j < H; j++) grow[ybase + j] = d16[j];
            }
        }
        __syncwarp();
    }
}

// ---- pass 2a: windowed ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_edt_cols_window(const uint16_t *__restrict__ g, int32_t *__restrict__ out,
                                                         int W, int H, int *__restrict__ flag)
{
    const size_t total = (size_t)W * H;
    bool open = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i / H);
        const unsigned g0 = g[i];
        long long best = g0 == EDT_NONE ? (long long)1 << 40 : (long long)g0 * g0;
        int d = 1;
        for (; d <= EDT_WIN; d++) {
            if ((long long)d * d >= best) break;
            if (x - d < 0 && x + d >= W) break;  // nothing left on either side
            if (x - d >= 0) {
                unsigned gv = __ldg(g + i - (size_t)d * H);
                if (gv != EDT_NONE) best = min(best, (long long)d * d + (long long)gv * gv);
            }
            if (x + d < W) {
                unsigned gv = __ldg(g + i + (size_t)d * H);
                if (gv != EDT_NONE) best = min(best, (long long)d * d + (long long)gv * gv);
            }
        }
        if (d > EDT_WIN && (long long)d * d < best && !(x - d < 0 && x + d >= W)) open = true;
        out[i] = best > 0x7FFFFFFFLL ? 0x7FFFFFFF : (int32_t)best;
    }
    if (open) *flag = 1;
}

// ---- pass 2a, tiled: the same window scan out of shared memory ------------------------------------
// One CTA owns EDT_TX x EDT_TYC output cells and stages the (EDT_TX + 2*EDT_WIN) x EDT_TYC block of g it can
// need (coalesced 16-byte loads, every g value is read from HBM/L2 once per tile instead of once per probe).
// All arithmetic fits 32 unsigned bits: g <= 65534 -> g*g + EDT_WIN^2 < 2^32.
#define EDT_TX 64
#define EDT_TYC 128
__global__ void __launch_bounds__(256) k_edt_cols_tile(const uint16_t *__restrict__ g, int32_t *__restrict__ out, int W, int H,
                                                       int *__restrict__ flag, const int *__restrict__ run_flag)
{
    if (run_flag && *run_flag == 0) return;
    __shared__ __align__(16) uint16_t tile[EDT_TX + 2 * EDT_WIN][EDT_TYC];
    constexpr int ROWS = EDT_TX + 2 * EDT_WIN;
    const int ntx = (H + EDT_TYC - 1) / EDT_TYC, nty = (W + EDT_TX - 1) / EDT_TX;
    bool open = false;
    // the grid is capped (as the conditional fallback of the bit-parallel path this kernel usually returns at once: tens
    // of thousands of CTAs doing that cost 20 us at 16384^2), a CTA walks over tiles
    for (int t = blockIdx.x; t < ntx * nty; t += gridDim.x) {
        const int x0 = (t / ntx) * EDT_TX, y0 = (t % ntx) * EDT_TYC;
        // stage: 16 threads x 16 bytes per row, 16 rows per step
        {
            const int cchunk = (threadIdx.x & 15) * 8, r0 = threadIdx.x >> 4;
            for (int rr = r0; rr < ROWS; rr += 16) {
                const int x = x0 - EDT_WIN + rr;
                uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                if (x >= 0 && x < W && y0 + cchunk < H) v = __ldcs(reinterpret_cast<const uint4 *>(g + (size_t)x * H + y0 + cchunk));
                *reinterpret_cast<uint4 *>(&tile[rr][cchunk]) = v;
            }
        }
        __syncthreads();
        const int c = threadIdx.x & (EDT_TYC - 1);
        if (y0 + c < H) {
            for (int r = threadIdx.x >> 7; r < EDT_TX; r += 2) {
                const int x = x0 + r;
                if (x >= W) break;
                const int tr = r + EDT_WIN;
                const unsigned g0 = tile[tr][c];
                unsigned best = g0 == EDT_NONE ? 0xFFFFFFFFu : g0 * g0;
                int d = 1;
                for (; d <= EDT_WIN; d++) {
                    const unsigned dd = (unsigned)(d * d);
                    if (dd >= best) break;
                    if (x - d < 0 && x + d >= W) break;  // nothing left on either side
                    const unsigned ga = tile[tr - d][c], gb = tile[tr + d][c];  // rows outside the grid were staged as NONE
                    if (ga != EDT_NONE) best = min(best, dd + ga * ga);
                    if (gb != EDT_NONE) best = min(best, dd + gb * gb);
                }
                if (d > EDT_WIN && (unsigned)(d * d) < best && !(x - d < 0 && x + d >= W)) open = true;
                __stcs(out + (size_t)x * H + y0 + c, best > 0x7FFFFFFFu ? 0x7FFFFFFF : (int32_t)best);
            }
        }
        __syncthreads();
    }
    if (open) *flag = 1;
}

// ---- pass 2b: Meijster et al. lower envelope, one thread per column ------------------------------
__device__ __forceinline__ long long edt_f(int x, int i, long long gi2) { return (long long)(x - i) * (x - i) + gi2; }
__global__ void __launch_bounds__(128) k_edt_cols_exact(const uint16_t *__restrict__ g, int32_t *__restrict__ out, int W, int H,
                                                        uint16_t *__restrict__ S, uint16_t *__restrict__ T,
                                                        const int *__restrict__ flag)
{
    if (*flag == 0) return;
    const long long BIG = (long long)1 << 40;
    for (int y = blockIdx.x * blockDim.x + threadIdx.x; y < H; y += gridDim.x * blockDim.x) {
        auto G2 = [&](int i) -> long long { unsigned v = g[(size_t)i * H + y]; return v == EDT_NONE ? BIG : (long long)v * v; };
        int q = 0;
        int s_top = 0, t_top = 0;
        long long gs_top = G2(0);
        S[y] = 0; T[y] = 0;
        for (int u = 1; u < W; u++) {
            const long long gu = G2(u);
            while (q >= 0 && edt_f(t_top, s_top, gs_top) > edt_f(t_top, u, gu)) {
                q--;
                if (q >= 0) { s_top = S[(size_t)q * H + y]; t_top = T[(size_t)q * H + y]; gs_top = G2(s_top); }
            }
            if (q < 0) {
                q = 0; s_top = u; t_top = 0; gs_top = gu;
                S[y] = (uint16_t)u; T[y] = 0;
            } else {
                // Sep(i,u) = (u^2 - i^2 + g(u)^2 - g(i)^2) div (2(u-i)), floor division (numerator may be negative)
                long long num = (long long)u * u - (long long)s_top * s_top + gu - gs_top;
                long long den = 2LL * (u - s_top);
                long long sep = num >= 0 ? num / den : -((-num + den - 1) / den);
                long long w = 1 + sep;
                if (w < W) {
                    if (w < 0) w = 0;
                    q++; s_top = u; t_top = (int)w; gs_top = gu;
                    S[(size_t)q * H + y] = (uint16_t)u; T[(size_t)q * H + y] = (uint16_t)w;
                }
            }
        }
        for (int u = W - 1; u >= 0; u--) {
            long long v = edt_f(u, s_top, gs_top);
            out[(size_t)u * H + y] = v > 0x7FFFFFFFLL ? 0x7FFFFFFF : (int32_t)v;
            if (u == t_top && q > 0) {
                q--;
                s_top = S[(size_t)q * H + y]; t_top = T[(size_t)q * H + y]; gs_top = G2(s_top);
            }
        }
    }
}

// ---- dense-map fast path: bit-parallel exact EDT ----------------------------------------------------
// dist^2(x,y) = min over row offsets d of d^2 + g(x+-d, y)^2 with g the integer distance to the nearest occupied cell
// in that row, so dist^2 <= D  <=>  some (d, r) with d^2 + r^2 <= D has B_r(x+-d, y) set, where B_r is the row's
// occupancy mask dilated by r along y.  Walking the distinct values D = d^2 + r^2 in increasing order and OR-ing the
// masks of the pairs with d^2 + r^2 == D settles 32 cells per operation: the level at which a cell's bit first turns
// on IS its exact squared distance.  Up to EDT_R (cells farther than that from every obstacle are rare on the maps
// this path is for: they go to a fix-up list; if that list overflows the grid is redone by the windowed path above).
//   k_edt_pack:   occupancy bytes -> bit words along y, stored column-major (268 MB -> 33 MB at 16384^2)
//   k_edt_strips: one WARP per strip of 32 columns (one bit word) x a segment of rows, sliding down the rows 64 at a time
//                 (two adjacent rows per lane); the masks B_0..B_R of the 64 + 2 R rows around them live in the warp's
//                 own shared memory.  Traffic: bits in, int32 out -- the algorithmic 5 B/cell.
//   k_edt_fix:    brute-force search on the bit grid for the listed cells.
// History of the bound (16384^2, 2 % fill; profiles/r02i_edt_*): rounds 1-2 ran one CTA per 64 x 256-cell tile (stage raw
// words, barrier, build the masks incl. 2 R halo rows, barrier, walk; a thread = one row x one word, 128 bytes of output
// each): 0.52 ms, integer pipe saturated (math-pipe throttle the top stall).  Halving the instruction count (items
// below) moved the bound to the two block barriers (7.0 stalls per issue), removing the barriers (warp strips) to the
// LSU data pipe -- 86 % busy, three quarters of it the output stores: a store in which every lane writes 16 bytes of its
// own row is 32 wavefronts, not 4 -- and coalescing those to the raw-word loads (every lane its own row again).  Now:
//   * no block barrier, no per-tile prologue: a warp is a self-contained pipeline (__syncwarp only); the raw words of
//     the next 64 rows are loaded before the current 64 rows are walked; halo rows are built once per segment, not
//     once per tile; going to the next step the last 2 R rows of masks move to the front of the buffer (one lane per
//     row) and 64 new rows are built behind them, so the walk keeps its immediate shared-memory offsets;
//   * a lane owns two ADJACENT rows: row x + 1 at row offset d reads the word row x reads at offset d + 1 -- the same
//     address in the unrolled code, loaded once (246 instead of 466 loads for both rows over the whole walk); rows at
//     even and odd buffer positions are stored in separate halves so that the lanes' stride-2 rows hit 32 different banks;
//   * coalesced global traffic on both sides: the bit grid is column-major (a lane's two rows are one 8-byte load, a
//     warp's 64 rows one 256-byte run), the output is staged through shared memory as bytes and written so that eight
//     lanes cover one row's 128 bytes (edt_flush_rows);
//   * masks: a 64-bit window (16 bits of the left word | the word | 16 bits of the right word) is dilated by one cell
//     per r (6 logic ops) and the centre word extracted with one funnel shift -- exact while r <= 16;
//   * walk: `cum |= OR of the level's pairs` is the whole per-level work; the squared distance is recorded bit-sliced
//     (plane k = cells whose D has bit k set), but per RUN of consecutive levels whose D has bit k set instead of per
//     level: plane k |= cum(after the run) & ~cum(before it); `all 64 cells settled` is tested every third level;
//   * planes -> bytes: the 8 planes are 4 x (8 x 8) bit matrices side by side (one per byte); three rounds of delta swaps
//     between plane registers transpose all four at once (60 ops), two 4 x 4 byte transposes (16 permutes) put four
//     consecutive cells into one word, one byte permute per cell widens them after the staging.  (Round 1: 8 x 8 x 5
//     spread / multiply / mask operations + 3 per cell to substitute INT32_MAX in unsettled cells = 450 ops; unsettled
//     cells are overwritten by the fix-up anyway, so whatever the planes hold for them is stored.)
#ifndef EDT_R
#define EDT_R 14 /* cells farther than this from every obstacle go to the fix-up kernel: 0.18 % of a 2 %-filled grid at R = 10 (146 us at 16384^2), 0.014 % at 12 (40 us), 0.0004 % at 14 */
#endif
#define EDT_STEP 64                       /* rows a warp settles per step: two adjacent rows per lane */
#define EDT_NB (EDT_STEP + 2 * EDT_R)     /* rows of masks a warp keeps: the step's rows + R on either side */
#define EDT_HALF (EDT_NB / 2)
#ifndef EDT_WS_WARPS
#define EDT_WS_WARPS 8  /* warps per CTA (no block barrier: the CTA is only a container) */
#endif
#define EDT_STAGE_WORDS (33 * 8)
#define EDT_MAX_PAIRS 256
#define EDT_MAX_LEVELS 128
#define EDT_WS_SMEM (EDT_WS_WARPS * ((EDT_R + 1) * EDT_NB + 2 * EDT_STAGE_WORDS) * 4)
static_assert(EDT_R <= 16, "the 64-bit dilation window is exact for the centre word while r <= 16");
static_assert(EDT_R % 2 == 0 && 2 * EDT_R <= 32, "lane l owns buffer rows R + 2 l (even) and R + 2 l + 1; the 2 R kept rows are copied by one lane each");
// distinct D = d^2 + r^2 (0 <= d, r <= EDT_R) in increasing order, limited to D <= EDT_R^2 (beyond that a pair with a
// larger d or r, which the buffer does not hold, could win); built at compile time so that the level walk is
// straight-line code with immediate shared-memory offsets
struct EdtLevels {
    int nlevels, npairs;
    int D[EDT_MAX_LEVELS];
    int start[EDT_MAX_LEVELS + 1];
    int pd[EDT_MAX_PAIRS], pr[EDT_MAX_PAIRS];
};
__host__ __device__ constexpr EdtLevels edt_make_levels()
{
    EdtLevels h{};
    int np = 0, nl = 0;
    for (int D = 0; D <= EDT_R * EDT_R; D++) {
        bool any = false;
        for (int d = 0; d <= EDT_R; d++)
            for (int r = 0; r <= EDT_R; r++)
                if (d * d + r * r == D) {
                    if (!any) { h.D[nl] = D; h.start[nl] = np; any = true; }
                    h.pd[np] = d; h.pr[np] = r; np++;
                }
        if (any) nl++;
    }
    h.start[nl] = np;
    h.nlevels = nl; h.npairs = np;
    return h;
}
static_assert(edt_make_levels().nlevels <= EDT_MAX_LEVELS && edt_make_levels().npairs <= EDT_MAX_PAIRS, "EDT level table sizes");
static_assert(EDT_R * EDT_R < 256, "eight bit planes");

// The buffer of one r: rows at even positions first, then the rows at odd positions (lane l's rows are positions
// R + 2 l and R + 2 l + 1, so a warp reads 32 consecutive words whatever the row offset: no bank conflicts).
// Word offset of mask r of the row `e` positions below the lane's first row, relative to T + lane:
__host__ __device__ constexpr int edt_off(int r, int e) { return r * EDT_NB + ((EDT_R + e) & 1) * EDT_HALF + ((EDT_R + e) >> 1); }

// OR of the masks of pairs [P, PEND) for the lane's row ROW (0 / 1).  Row 1 at offset d reads the word row 0 reads at
// offset d + 1 -- the same immediate address, which the compiler loads once.
template <int P, int PEND, int ROW>
__device__ __forceinline__ unsigned edt_pairs(const unsigned *__restrict__ base)
{
    if constexpr (P >= PEND) return 0u;
    else {
        constexpr EdtLevels tab = edt_make_levels();
        constexpr int d = tab.pd[P], r = tab.pr[P];
        if constexpr (d == 0) return base[edt_off(r, ROW)] | edt_pairs<P + 1, PEND, ROW>(base);
        else return base[edt_off(r, ROW - d)] | base[edt_off(r, ROW + d)] | edt_pairs<P + 1, PEND, ROW>(base);
    }
}
// Level LV: cum = cells within squared distance D[LV] of an obstacle.  A cell's squared distance is the D of the level
// that turned its bit on; plane k collects the cells whose D has bit k set, one update per run of consecutive levels
// with that bit set (sv[k] = cum before the run; in the unrolled code it is just the name of an older register).
template <int LV>
__device__ __forceinline__ void edt_walk(const unsigned *__restrict__ base, unsigned &cumA, unsigned &cumB, unsigned (&plA)[8], unsigned (&plB)[8],
                                         unsigned (&svA)[8], unsigned (&svB)[8])
{
    constexpr EdtLevels tab = edt_make_levels();
    if constexpr (LV < tab.nlevels) {
        constexpr int D = tab.D[LV];
        constexpr int Dprev = LV > 0 ? tab.D[LV - 1] : 0;
        constexpr int Dnext = LV + 1 < tab.nlevels ? tab.D[LV + 1] : 0;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (((D >> k) & 1) && !((Dprev >> k) & 1)) { svA[k] = cumA; svB[k] = cumB; }
        cumA |= edt_pairs<tab.start[LV], tab.start[LV + 1], 0>(base);
        cumB |= edt_pairs<tab.start[LV], tab.start[LV + 1], 1>(base);
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (((D >> k) & 1) && !((Dnext >> k) & 1)) { plA[k] |= cumA & ~svA[k]; plB[k] |= cumB & ~svB[k]; }
        if constexpr (LV % 3 == 2 && LV + 1 < tab.nlevels) {
            if ((cumA & cumB) == 0xFFFFFFFFu) {  // every cell of both words is settled: close the runs that are still open and stop
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (((D >> k) & 1) && ((Dnext >> k) & 1)) { plA[k] |= cumA & ~svA[k]; plB[k] |= cumB & ~svB[k]; }
                return;
            }
        }
        edt_walk<LV + 1>(base, cumA, cumB, plA, plB, svA, svB);
    }
}

// occupancy bytes -> bit words, stored COLUMN-major: word w of row x at bitsT[w * Wp + x] (Wp = W rounded up to even, the
// pad row is zero).  A strip (one word column) is then contiguous along x: the two rows a lane of k_edt_strips owns
// are one 8-byte load and a warp's 64 rows one 256-byte run (row-major, every lane fetched its own 32-byte sector three
// times per step: as many data-pipe wavefronts as the whole walk).  One CTA per 32 rows x 32 words, transposed through
// shared memory so that both the byte reads and the word writes are coalesced.
__global__ void __launch_bounds__(256) k_edt_pack(const uint8_t *__restrict__ occ, unsigned *__restrict__ bitsT, int W, int HW, int Wp)
{
    __shared__ unsigned tile[32][33];
    const int x0 = blockIdx.y * 32, w0 = blockIdx.x * 32;
    const int wi = threadIdx.x & 31, r0 = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int x = x0 + r0 + 8 * j, w = w0 + wi;
        unsigned m = 0;
        if (x < W && w < HW) {
            const uint4 *src = reinterpret_cast<const uint4 *>(occ + ((size_t)x * HW + w) * 32);
            const uint4 a = __ldcs(src), b = __ldcs(src + 1);
            const unsigned v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const unsigned nz = __vcmpne4(v[i], 0u) & 0x01010101u;
                m |= (((nz * 0x01020408u) >> 24) & 0xFu) << (4 * i);
            }
        }
        tile[r0 + 8 * j][wi] = m;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int w = w0 + r0 + 8 * j, x = x0 + wi;
        if (w < HW && x < Wp) bitsT[(size_t)w * Wp + x] = tile[wi][r0 + 8 * j];
    }
}

// a <-> b: the bits of `a` selected by (m << s) change places with the bits of `b` selected by m
#define EDT_DSWAP(a, b, s, m) { const unsigned t__ = (((a) >> (s)) ^ (b)) & (m); (b) ^= t__; (a) ^= t__ << (s); }

// 8 bit planes of one word -> the 32 squared distances as bytes, parked in the warp's staging buffer (row-major, 8 words
// per row, every group of four rows shifted by one word: conflict-free for the row-wise writes here and for the
// column-wise reads of edt_flush_rows)
__device__ __forceinline__ void edt_stage_row(unsigned (&q)[8], unsigned *__restrict__ st /* stage + 33 (lane >> 2) + 8 (lane & 3) */)
{
    // transpose the four 8 x 8 bit matrices (rows: planes, columns: the bits of one byte): afterwards byte j of q[i] is the
    // squared distance of cell 8 j + i
#pragma unroll
    for (int k = 0; k < 4; k++) EDT_DSWAP(q[k], q[k + 4], 4, 0x0F0F0F0Fu)
    EDT_DSWAP(q[0], q[2], 2, 0x33333333u) EDT_DSWAP(q[1], q[3], 2, 0x33333333u)
    EDT_DSWAP(q[4], q[6], 2, 0x33333333u) EDT_DSWAP(q[5], q[7], 2, 0x33333333u)
#pragma unroll
    for (int k = 0; k < 8; k += 2) EDT_DSWAP(q[k], q[k + 1], 1, 0x55555555u)
    // 4 x 4 byte transposes: word c = cells 4 c .. 4 c + 3 = byte (c >> 1) of q[4 (c & 1) + 0..3]
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const unsigned t0 = __byte_perm(q[4 * h], q[4 * h + 1], 0x5140u), t1 = __byte_perm(q[4 * h], q[4 * h + 1], 0x7362u);
        const unsigned t2 = __byte_perm(q[4 * h + 2], q[4 * h + 3], 0x5140u), t3 = __byte_perm(q[4 * h + 2], q[4 * h + 3], 0x7362u);
        st[h] = __byte_perm(t0, t2, 0x5410u);
        st[2 + h] = __byte_perm(t0, t2, 0x7632u);
        st[4 + h] = __byte_perm(t1, t3, 0x5410u);
        st[6 + h] = __byte_perm(t1, t3, 0x7632u);
    }
}
// The staged rows go out coalesced: in store i the eight lanes 8 k .. 8 k + 7 write the 128 bytes of staged row 4 i + k
// (grid row x0 + 2 (4 i + k): a lane's two rows are adjacent, so a staging buffer holds every other row).  A store in
// which every lane writes its own row costs 32 data-pipe wavefronts instead of 4 -- that, not DRAM, bounded the
// transform (ncu: LSU data pipe 86 % busy, 3/4 of it these stores).
__device__ __forceinline__ void edt_flush_rows(const unsigned *__restrict__ stage, int32_t *__restrict__ out, int x0, int xe, int HW, int w, int lane)
{
    const int k = lane >> 3, c = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const unsigned v = stage[lane + 33 * i];
        const int x = x0 + 2 * (4 * i + k);
        if (x < xe) {
            int4 o;
            o.x = (int)__byte_perm(v, 0u, 0x4440u); o.y = (int)__byte_perm(v, 0u, 0x4441u);
            o.z = (int)__byte_perm(v, 0u, 0x4442u); o.w = (int)__byte_perm(v, 0u, 0x4443u);
            __stcs(reinterpret_cast<int4 *>(out + ((size_t)x * HW + w) * 32) + c, o);
        }
    }
}

__global__ void __launch_bounds__(EDT_WS_WARPS * 32) k_edt_strips(const unsigned *__restrict__ bitsT, int32_t *__restrict__ out, int W, int HW, int Wp, int seg_rows,
                                                                  unsigned *__restrict__ fix_list, unsigned *__restrict__ fix_count, unsigned fix_cap)
{
    extern __shared__ __align__(16) unsigned edt_buf[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.x * EDT_WS_WARPS + warp;  // the strip: bit word w of every row = columns 32 w .. 32 w + 31
    if (w >= HW) return;                             // (whole warps leave; there is no block barrier below)
    unsigned *__restrict__ T = edt_buf + warp * ((EDT_R + 1) * EDT_NB + 2 * EDT_STAGE_WORDS);  // T[r][even rows | odd rows]
    unsigned *__restrict__ stage = T + (EDT_R + 1) * EDT_NB;  // two staging buffers: the lanes' first rows, the lanes' second rows
    unsigned *__restrict__ st_mine = stage + 33 * (lane >> 2) + 8 * (lane & 3);
    const int xs = blockIdx.y * seg_rows, xe = min(xs + seg_rows, W);
    const bool hasL = w > 0, hasR = w + 1 < HW;
    unsigned raw[2][3];
    auto load_raw = [&](int x) {  // the three words around the strip of rows x (even) and x + 1 (zero outside the grid)
#pragma unroll
        for (int i = 0; i < 3; i++) { raw[0][i] = 0; raw[1][i] = 0; }
        if (x >= 0 && x < W) {
            const uint2 *__restrict__ p = reinterpret_cast<const uint2 *>(bitsT + (size_t)w * Wp + x);
            const uint2 c = __ldg(p);
            raw[0][1] = c.x; raw[1][1] = c.y;
            if (hasL) { const uint2 l = __ldg(p - (Wp >> 1)); raw[0][0] = l.x; raw[1][0] = l.y; }
            if (hasR) { const uint2 r = __ldg(p + (Wp >> 1)); raw[0][2] = r.x; raw[1][2] = r.y; }
        }
    };
    auto build = [&](unsigned *__restrict__ t, const unsigned (&v)[3]) {  // B_r for r = 0..EDT_R of one row, t = its slot in the r = 0 buffer
        unsigned lo = __funnelshift_r(v[0], v[1], 16), hi = __funnelshift_r(v[1], v[2], 16);
        t[0] = v[1];
#pragma unroll
        for (int r = 1; r <= EDT_R; r++) {
            const unsigned nlo = lo | (lo << 1) | __funnelshift_r(lo, hi, 1);
            const unsigned nhi = hi | __funnelshift_l(lo, hi, 1) | (hi >> 1);
            lo = nlo; hi = nhi;
            t[r * EDT_NB] = __funnelshift_r(lo, hi, 16);
        }
    };
    // Buffer position p holds row a - R + p of the current step (rows a .. a + 63 are settled, lane l: positions R + 2 l
    // and R + 2 l + 1).  Going to the next step the last 2 R rows move to the front (one lane per row), 64 new rows are
    // built behind them.  Prologue: the 2 R rows in front of the segment, built where the first step's move expects them.
    if (lane < EDT_R) {
        load_raw(xs - EDT_R + 2 * lane);
        build(T + EDT_STEP / 2 + lane, raw[0]); build(T + EDT_HALF + EDT_STEP / 2 + lane, raw[1]);
    }
    load_raw(xs + EDT_R + 2 * lane);
    unsigned *__restrict__ mv = T + (lane & 1) * EDT_HALF + (lane >> 1);  // lane j < 2 R moves position 64 + j to position j
    for (int a = xs; a < xe; a += EDT_STEP) {
        __syncwarp();
        if (lane < 2 * EDT_R) {
#pragma unroll
            for (int r = 0; r <= EDT_R; r++) mv[r * EDT_NB] = mv[r * EDT_NB + EDT_STEP / 2];
        }
        __syncwarp();
        build(T + EDT_R + lane, raw[0]); build(T + EDT_HALF + EDT_R + lane, raw[1]);  // positions 2 R + 2 l, 2 R + 2 l + 1 = rows a + R + 2 l (+ 1)
        if (a + EDT_STEP < xe) load_raw(a + EDT_STEP + EDT_R + 2 * lane);
        __syncwarp();
        const int x = a + 2 * lane;
        unsigned umA = 0, umB = 0;
        if (x < xe) {
            unsigned cumA = 0, cumB = 0;
            unsigned qA[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, qB[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            unsigned svA[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, svB[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            edt_walk<0>(T + lane, cumA, cumB, qA, qB, svA, svB);
            umA = ~cumA;  // farther than EDT_R from every obstacle: fix-up list
            edt_stage_row(qA, st_mine);
            if (x + 1 < xe) umB = ~cumB;
            edt_stage_row(qB, st_mine + EDT_STAGE_WORDS);
        }
        __syncwarp();
        edt_flush_rows(stage, out, a, xe, HW, w, lane);
        edt_flush_rows(stage + EDT_STAGE_WORDS, out, a + 1, xe, HW, w, lane);
        // The unresolved cells go to the global list, one reservation per warp.  A list longer than fix_cap makes
        // k_edt_fix raise the flag and the windowed path redoes the grid.
        const unsigned mine = (unsigned)(__popc(umA) + __popc(umB));
        const unsigned total = __reduce_add_sync(0xFFFFFFFFu, mine);
        if (total) {
            unsigned incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
            unsigned rbase = 0;
            if (lane == 0) rbase = atomicAdd(fix_count, total);
            rbase = __shfl_sync(0xFFFFFFFFu, rbase, 0);
            size_t pos = (size_t)rbase + (incl - mine);
            while (umA) {
                const int b = __ffs(umA) - 1;
                umA &= umA - 1;
                if (pos < fix_cap) { fix_list[2 * pos] = (unsigned)x; fix_list[2 * pos + 1] = (unsigned)(w * 32 + b); }
                pos++;
            }
            while (umB) {
                const int b = __ffs(umB) - 1;
                umB &= umB - 1;
                if (pos < fix_cap) { fix_list[2 * pos] = (unsigned)(x + 1); fix_list[2 * pos + 1] = (unsigned)(w * 32 + b); }
                pos++;
            }
        }
    }
}

// nearest set bit to column y in one row of the bit grid, looking no farther than `lim` cells; returns the distance or -1
// (rowbits[w * stride] is word w of the row: the bit grid is stored column-major, k_edt_pack)
__device__ int edt_row_nearest(const unsigned *__restrict__ rowbits, size_t stride, int HW, int y, int lim)
{
    const int w = y >> 5, b = y & 31;
    int best = -1;
    {
        // the word of the cell and its two neighbours in one go (three independent loads: everything within 32 cells on
        // either side, which is where the answer nearly always is -- a listed cell is just over EDT_R from an obstacle)
        const unsigned m = rowbits[w * stride];
        const unsigned ml = w > 0 ? rowbits[(w - 1) * stride] : 0u, mr = w + 1 < HW ? rowbits[(w + 1) * stride] : 0u;
        const unsigned below = m & (0xFFFFFFFFu >> (31 - b)), above = m >> b;
        if (below) best = b - (31 - __clz(below));
        if (above) { const int d = __ffs(above) - 1; if (best < 0 || d < best) best = d; }
        if (ml) { const int d = y - ((w - 1) * 32 + 31 - __clz(ml)); if (best < 0 || d < best) best = d; }
        if (mr) { const int d = (w + 1) * 32 + __ffs(mr) - 1 - y; if (best < 0 || d < best) best = d; }
    }
    for (int k = 2; (k - 1) * 32 < lim && (best < 0 || (k - 1) * 32 < best); k++) {
        if (w - k >= 0) { const unsigned m = rowbits[(w - k) * stride]; if (m) { const int d = y - ((w - k) * 32 + 31 - __clz(m)); if (best < 0 || d < best) best = d; } }
        if (w + k < HW) { const unsigned m = rowbits[(w + k) * stride]; if (m) { const int d = (w + k) * 32 + __ffs(m) - 1 - y; if (best < 0 || d < best) best = d; } }
    }
    return (best >= 0 && best <= lim) ? best : -1;
}

// exact squared distance of cell (x, y) by brute force on the bit grid, the 32 lanes of a warp sharing the row offsets:
// lane l looks at the rows x +- d for d = l, l + 32, ...; the warp's best so far bounds every lane's search after each round
__device__ __forceinline__ long long edt_cell_warp(const unsigned *__restrict__ bits, int W, int HW, int Wp, int x, int y, unsigned lane)
{
    const int H = HW * 32;
    long long best = (long long)1 << 40;
    for (int d0 = 0; (long long)d0 * d0 < best && (x - d0 >= 0 || x + d0 < W); d0 += 32) {
        const int d = d0 + (int)lane;
        long long mine = (long long)1 << 40;
        if ((long long)d * d < best) {
            const long long room = best - (long long)d * d;
            int lim = H;
            if (room < (long long)H * H) { lim = (int)sqrtf((float)room) + 1; if (lim > H) lim = H; }
            if (x - d >= 0) { const int g = edt_row_nearest(bits + (x - d), (size_t)Wp, HW, y, lim); if (g >= 0) mine = min(mine, (long long)d * d + (long long)g * g); }
            if (d > 0 && x + d < W) { const int g = edt_row_nearest(bits + (x + d), (size_t)Wp, HW, y, lim); if (g >= 0) mine = min(mine, (long long)d * d + (long long)g * g); }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mine = min(mine, __shfl_xor_sync(0xFFFFFFFFu, mine, o));
        best = min(best, mine);
    }
    return best;
}

__global__ void __launch_bounds__(128) k_edt_fix(const unsigned *__restrict__ bits, int32_t *__restrict__ out, int W, int HW, int Wp,
                                                 const unsigned *__restrict__ fix_list, const unsigned *__restrict__ fix_count, unsigned fix_cap,
                                                 int *__restrict__ flag)
{
    const unsigned n = *fix_count;
    if (n > fix_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) *flag = 1; return; }  // too many: the windowed path redoes the grid
    const int H = HW * 32;
    const unsigned lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    if (n > 8u * nwarps) {
        // long list (large sparse grids): one thread per cell, the list itself is the parallelism
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const int x = (int)fix_list[2 * (size_t)i], y = (int)fix_list[2 * (size_t)i + 1];
            long long best = (long long)1 << 40;
            for (int d = 0; (long long)d * d < best && (x - d >= 0 || x + d < W); d++) {
                const long long room = best - (long long)d * d;
                int lim = H;
                if (room < (long long)H * H) { lim = (int)sqrtf((float)room) + 1; if (lim > H) lim = H; }
                if (x - d >= 0) { const int g = edt_row_nearest(bits + (x - d), (size_t)Wp, HW, y, lim); if (g >= 0) best = min(best, (long long)d * d + (long long)g * g); }
                if (d > 0 && x + d < W) { const int g = edt_row_nearest(bits + (x + d), (size_t)Wp, HW, y, lim); if (g >= 0) best = min(best, (long long)d * d + (long long)g * g); }
            }
            out[(size_t)x * H + y] = best > 0x7FFFFFFFLL ? 0x7FFFFFFF : (int32_t)best;
        }
        return;
    }
    // short list (small grids): one warp per cell (a listed cell is > EDT_R from everything, so the brute force walks tens
    // of rows: one thread doing that alone was the longest kernel of the whole transform at 1024^2)
    for (unsigned i = warp; i < n; i += nwarps) {
        const int x = (int)fix_list[2 * (size_t)i], y = (int)fix_list[2 * (size_t)i + 1];
        const long long best = edt_cell_warp(bits, W, HW, Wp, x, y, lane);
        if (lane == 0) out[(size_t)x * H + y] = best > 0x7FFFFFFFLL ? 0x7FFFFFFF : (int32_t)best;
    }
}

static int edt_upload_levels(fx_context *ctx)
{
    static bool done[64] = {false};
    if (ctx->device >= 0 && ctx->device < 64 && done[ctx->device]) return FX_OK;
    FX_CUDA(ctx, cudaFuncSetAttribute(k_edt_strips, cudaFuncAttributeMaxDynamicSharedMemorySize, EDT_WS_SMEM));
    if (ctx->device >= 0 && ctx->device < 64) done[ctx->device] = true;
    return FX_OK;
}

static int edt_reserve(fx_context *ctx, size_t cells)
{
    if (ctx->edt_cap >= cells) return FX_OK;
    if (ctx->edt_g) cudaFree(ctx->edt_g);
    if (ctx->edt_s) cudaFree(ctx->edt_s);
    if (ctx->edt_t) cudaFree(ctx->edt_t);
    ctx->edt_g = ctx->edt_s = ctx->edt_t = nullptr; ctx->edt_cap = 0;
    FX_CUDA(ctx, cudaMalloc(&ctx->edt_g, cells * 2 + 64));
    FX_CUDA(ctx, cudaMalloc(&ctx->edt_s, cells * 2));
    FX_CUDA(ctx, cudaMalloc(&ctx->edt_t, cells * 2));
    ctx->edt_cap = cells;
    return FX_OK;
}
static int edt_rows_launch(fx_context *ctx, const uint8_t *occ, uint16_t *g, int W, int H, cudaStream_t st);
static int edt_cols_grid(fx_context *ctx, int W, int H)
{
    const long long tiles = (long long)((H + EDT_TYC - 1) / EDT_TYC) * ((W + EDT_TX - 1) / EDT_TX);
    const long long cap = (long long)ctx->sm_count * 16;
    return (int)(tiles < cap ? tiles : cap);
}
static int edt_cols_launch(fx_context *ctx, const uint16_t *g, int32_t *dist2, int W, int H, cudaStream_t st);

extern "C" int fx_edt(fx_context *ctx, const uint8_t *occ, int32_t *dist2, int W, int H, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!occ || !dist2 || W <= 0 || H <= 0) return fx_set_err(ctx, FX_ERR_ARG, "fx_edt: bad argument");
    if (W > 65534 || H > 65534) return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_edt: W,H must be <= 65534");
    // the output is int32 and INT32_MAX is the "no obstacle" sentinel: the largest possible squared distance must stay below it
    if ((long long)(W - 1) * (W - 1) + (long long)(H - 1) * (H - 1) >= 0x7FFFFFFFll)
        return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_edt: (W-1)^2 + (H-1)^2 must be < 2^31 - 1 (int32 squared distances)");
    cudaStream_t st = (cudaStream_t)stream;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t cells = (size_t)W * H;
    int rc0 = edt_reserve(ctx, cells);
    if (rc0) return rc0;
    if (H % 32 == 0 && ((uintptr_t)occ & 15u) == 0 && ((uintptr_t)dist2 & 15u) == 0) {
        // dense-map fast path; falls through to the windowed path (conditionally, on the device flag) if too many
        // cells are farther than EDT_R from every obstacle
        int rc = edt_upload_levels(ctx);
        if (rc) return rc;
        const int HW = H / 32;
        const int Wp = (W + 1) & ~1;  // column stride of the bit grid (even: a lane's two rows are one aligned 8-byte load)
        unsigned *bitsb = reinterpret_cast<unsigned *>(ctx->edt_s);        // scratch reuse: cells*2 bytes >= cells/8
        unsigned *fix_list = reinterpret_cast<unsigned *>(ctx->edt_t);     // cells*2 bytes -> cells/4 entries of 8 bytes
        unsigned *fix_count = reinterpret_cast<unsigned *>(ctx->edt_flag) + 1;
        size_t cap = cells / 4; if (cap > (1u << 22)) cap = 1u << 22;
        // segments of rows per strip (multiples of the 64-row step; a segment starts with 2 R extra rows of masks): about
        // four CTAs per resident slot on large grids (strips differ in work: the walk stops when a word is settled), one
        // wave of shorter segments when that would leave a segment less than four steps
        const int gx = (HW + EDT_WS_WARPS - 1) / EDT_WS_WARPS;
        int resident = 0;
        FX_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_edt_strips, EDT_WS_WARPS * 32, EDT_WS_SMEM));
        if (resident < 1) resident = 1;
        const int slots = ctx->sm_count * resident;
        int segs = slots * 4 / gx;
        if (segs < 1 || W / segs < 4 * EDT_STEP) segs = slots / gx;
        if (segs < 1) segs = 1;
        const int seg_rows = ((W + segs - 1) / segs + EDT_STEP - 1) / EDT_STEP * EDT_STEP;
        segs = (W + seg_rows - 1) / seg_rows;
        auto enqueue = [&](cudaStream_t st) -> int {
            FX_CUDA(ctx, cudaMemsetAsync(ctx->edt_flag, 0, 4 * sizeof(int), st));  // [0] fallback flag, [1] fix_count, [2] window-left-open flag
            k_edt_pack<<<dim3((HW + 31) / 32, (Wp + 31) / 32), 256, 0, st>>>(occ, bitsb, W, HW, Wp);
            FX_LAUNCH_CHECK(ctx);
            k_edt_strips<<<dim3(gx, segs), EDT_WS_WARPS * 32, EDT_WS_SMEM, st>>>(bitsb, dist2, W, HW, Wp, seg_rows, fix_list, fix_count, (unsigned)cap);
            FX_LAUNCH_CHECK(ctx);
            k_edt_fix<<<ctx->sm_count * 4, 128, 0, st>>>(bitsb, dist2, W, HW, Wp, fix_list, fix_count, (unsigned)cap, ctx->edt_flag);
            FX_LAUNCH_CHECK(ctx);
            // the windowed path below runs only if the flag was raised
            const int nwords = (H + 31) / 32;
            int warps = 8;
            size_t smem = (size_t)warps * 3 * nwords * 4;
            while (smem > 48 * 1024 && warps > 1) { warps >>= 1; smem = (size_t)warps * 3 * nwords * 4; }
            int blocks = (W + warps - 1) / warps;
            if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
            k_edt_rows<<<blocks, warps * 32, smem, st>>>(occ, ctx->edt_g, W, H, ctx->edt_flag);
            FX_LAUNCH_CHECK(ctx);
            k_edt_cols_tile<<<edt_cols_grid(ctx, W, H), 256, 0, st>>>(ctx->edt_g, dist2, W, H, ctx->edt_flag + 2, ctx->edt_flag);
            FX_LAUNCH_CHECK(ctx);
            k_edt_cols_exact<<<(H + 127) / 128, 128, 0, st>>>(ctx->edt_g, dist2, W, H, ctx->edt_s, ctx->edt_t, ctx->edt_flag + 2);
            FX_LAUNCH_CHECK(ctx);
            return FX_OK;
        };
        struct { const void *occ, *out, *g, *s, *t, *flag; int W, H; } key = {occ, dist2, ctx->edt_g, ctx->edt_s, ctx->edt_t, ctx->edt_flag, W, H};
        return fx_graph_run(ctx, FX_GRAPH_EDT, &key, sizeof(key), st, enqueue);
    }
    FX_CUDA(ctx, cudaMemsetAsync(ctx->edt_flag, 0, 4 * sizeof(int), st));
    int rc = edt_rows_launch(ctx, occ, ctx->edt_g, W, H, st);
    if (rc) return rc;
    return edt_cols_launch(ctx, ctx->edt_g, dist2, W, H, st);
}

// ---- the two separable passes as entry points: the row-tiled multi-GPU mode runs pass 1 on its x-slab, transposes
// the row distances between ranks, and runs pass 2 on whole columns (tiled.edt_tiled) ------------------------------
static int edt_rows_launch(fx_context *ctx, const uint8_t *occ, uint16_t *g, int W, int H, cudaStream_t st)
{
    const int nwords = (H + 31) / 32;
    // warps per CTA limited by shared memory (3 arrays of nwords per warp)
    int warps = 8;
    size_t smem = (size_t)warps * 3 * nwords * 4;
    while (smem > 48 * 1024 && warps > 1) { warps >>= 1; smem = (size_t)warps * 3 * nwords * 4; }
    int blocks = (W + warps - 1) / warps;
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_edt_rows<<<blocks, warps * 32, smem, st>>>(occ, g, W, H, nullptr);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}

static int edt_cols_launch(fx_context *ctx, const uint16_t *g, int32_t *dist2, int W, int H, cudaStream_t st)
{
    const size_t cells = (size_t)W * H;
    int b2 = (int)((cells + 255) / 256);
    if (b2 > ctx->sm_count * 16) b2 = ctx->sm_count * 16;
    if (H % 8 == 0 && ((uintptr_t)g & 15u) == 0) {
        k_edt_cols_tile<<<edt_cols_grid(ctx, W, H), 256, 0, st>>>(g, dist2, W, H, ctx->edt_flag, nullptr);
    } else {
        k_edt_cols_window<<<b2, 256, 0, st>>>(g, dist2, W, H, ctx->edt_flag);
    }
    FX_LAUNCH_CHECK(ctx);
    k_edt_cols_exact<<<(H + 127) / 128, 128, 0, st>>>(g, dist2, W, H, ctx->edt_s, ctx->edt_t, ctx->edt_flag);
    FX_LAUNCH_CHECK(ctx);
    return FX_OK;
}

extern "C" int fx_edt_rows(fx_context *ctx, const uint8_t *occ, uint16_t *g, int W, int H, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!occ || !g || W <= 0 || H <= 0) return fx_set_err(ctx, FX_ERR_ARG, "fx_edt_rows: bad argument");
    if (H > 65534) return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_edt_rows: H must be <= 65534");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    return edt_rows_launch(ctx, occ, g, W, H, (cudaStream_t)stream);
}

extern "C" int fx_edt_cols(fx_context *ctx, const uint16_t *g, int32_t *dist2, int W, int H, void *stream)
{
    if (!ctx) return FX_ERR_ARG;
    if (!g || !dist2 || W <= 0 || H <= 0) return fx_set_err(ctx, FX_ERR_ARG, "fx_edt_cols: bad argument");
    if (W > 65534) return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_edt_cols: W must be <= 65534");
    // int32 output with INT32_MAX as the "no obstacle" sentinel: the caller guarantees (W-1)^2 + (row extent - 1)^2 < 2^31 - 1
    // (fx_edt checks it for the whole grid; tiled.edt_tiled for the stitched one); here only W itself can be checked
    if ((long long)(W - 1) * (W - 1) >= 0x7FFFFFFFll) return fx_set_err(ctx, FX_ERR_UNSUPPORTED, "fx_edt_cols: (W-1)^2 must be < 2^31 - 1");
    cudaStream_t st = (cudaStream_t)stream;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = edt_reserve(ctx, (size_t)W * H);
    if (rc) return rc;
    FX_CUDA(ctx, cudaMemsetAsync(ctx->edt_flag, 0, 4 * sizeof(int), st));
    return edt_cols_launch(ctx, g, dist2, W, H, st);
}
