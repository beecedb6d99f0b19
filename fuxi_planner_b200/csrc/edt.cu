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
    const int x0 = blockIdx.y * EDT_TX, y0 = blockIdx.x * EDT_TYC;
    constexpr int ROWS = EDT_TX + 2 * EDT_WIN;
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
    bool open = false;
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
//   k_edt_pack: occupancy bytes -> bit words along y (268 MB -> 33 MB at 16384^2)
//   k_edt_bits: one CTA per 64 x 256-cell tile; shared memory holds B_0..B_R for the tile's rows + R halo rows
//               (built from three-word windows with funnel shifts), the per-cell level index, and the output is
//               written with 16-byte streaming stores.  Traffic: bits in, int32 out -- the algorithmic 5 B/cell.
//   k_edt_fix:  brute-force search on the bit grid for the listed cells.
#define EDT_R 12 /* 10 left 0.18 % of the cells of a 2 %-filled grid to the fix-up kernel (22 % of the time); 12 leaves 0.011 % */
#define EDT_BT_ROWS 64
#define EDT_BT_WORDS 8
#define EDT_MAX_PAIRS 256
#define EDT_MAX_LEVELS 128
#define EDT_BITS_SMEM (4 * (EDT_R + 1) * (EDT_BT_ROWS + 2 * EDT_R) * EDT_BT_WORDS + 4 * (EDT_BT_ROWS + 2 * EDT_R) * (EDT_BT_WORDS + 2))
// distinct D = d^2 + r^2 (0 <= d, r <= EDT_R) in increasing order, limited to D <= EDT_R^2 (beyond that a pair with a
// larger d or r, which the tile does not hold, could win); built at compile time so that the level walk is straight-line
// code with immediate shared-memory offsets
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

// OR of the masks of pairs [P, PEND): base points at T[0][tile row][word]
template <int P, int PEND>
__device__ __forceinline__ unsigned edt_pairs(const unsigned *__restrict__ base)
{
    if constexpr (P >= PEND) return 0u;
    else {
        constexpr EdtLevels tab = edt_make_levels();
        constexpr int ROWS = EDT_BT_ROWS + 2 * EDT_R;
        constexpr int d = tab.pd[P], r = tab.pr[P];
        constexpr int offA = (r * ROWS - d) * EDT_BT_WORDS, offB = (r * ROWS + d) * EDT_BT_WORDS;
        if constexpr (d == 0) return base[offA] | edt_pairs<P + 1, PEND>(base);
        else return base[offA] | base[offB] | edt_pairs<P + 1, PEND>(base);
    }
}
// level LV settles the cells in `nw`; their squared distance D (a compile-time constant) is recorded bit-sliced:
// plane k collects the cells whose D has bit k set -- no per-cell work, no divergence
template <int LV>
__device__ __forceinline__ void edt_levels(const unsigned *__restrict__ base, unsigned &done, unsigned (&pl)[8])
{
    constexpr EdtLevels tab = edt_make_levels();
    if constexpr (LV < tab.nlevels) {
        const unsigned nw = edt_pairs<tab.start[LV], tab.start[LV + 1]>(base) & ~done;
        done |= nw;
        constexpr int D = tab.D[LV];
        static_assert(D < 256, "eight bit planes");
#pragma unroll
        for (int k = 0; k < 8; k++)
            if ((D >> k) & 1) pl[k] |= nw;
        if (done != 0xFFFFFFFFu) edt_levels<LV + 1>(base, done, pl);
    }
}

__global__ void __launch_bounds__(256) k_edt_pack(const uint8_t *__restrict__ occ, unsigned *__restrict__ bits, size_t nwords)
{
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (size_t)gridDim.x * blockDim.x) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4 *>(occ) + 2 * w), b = __ldcs(reinterpret_cast<const uint4 *>(occ) + 2 * w + 1);
        const unsigned v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        unsigned m = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const unsigned nz = __vcmpne4(v[i], 0u) & 0x01010101u;
            m |= (((nz * 0x01020408u) >> 24) & 0xFu) << (4 * i);
        }
        bits[w] = m;
    }
}

__global__ void __launch_bounds__(EDT_BT_ROWS * EDT_BT_WORDS) k_edt_bits(const unsigned *__restrict__ bits, int32_t *__restrict__ out, int W, int HW,
                                                                         unsigned *__restrict__ fix_list, unsigned *__restrict__ fix_count, unsigned fix_cap,
                                                                         int *__restrict__ flag)
{
    (void)flag;
    constexpr int ROWS = EDT_BT_ROWS + 2 * EDT_R, TW = EDT_BT_WORDS;
    extern __shared__ __align__(16) unsigned char edt_smem[];
    unsigned(*T)[ROWS][TW] = reinterpret_cast<unsigned(*)[ROWS][TW]>(edt_smem);                                   // [EDT_R + 1]
    unsigned(*raw)[TW + 2] = reinterpret_cast<unsigned(*)[TW + 2]>(edt_smem + sizeof(unsigned) * (EDT_R + 1) * ROWS * TW);
    const int x0 = blockIdx.y * EDT_BT_ROWS, w0 = blockIdx.x * TW;
    const int tid = threadIdx.x;
    // raw occupancy words of the tile + halo (zero outside the grid)
    for (int i = tid; i < ROWS * (TW + 2); i += blockDim.x) {
        const int rr = i / (TW + 2), ww = i - rr * (TW + 2);
        const int x = x0 - EDT_R + rr, w = w0 - 1 + ww;
        raw[rr][ww] = (x >= 0 && x < W && w >= 0 && w < HW) ? __ldg(bits + (size_t)x * HW + w) : 0u;
    }
    __syncthreads();
    // B_r for r = 0..EDT_R from a three-word window (exact for the centre word while r <= 32)
    for (int i = tid; i < ROWS * TW; i += blockDim.x) {
        const int rr = i / TW, ww = i - rr * TW;
        unsigned L = raw[rr][ww], C = raw[rr][ww + 1], R = raw[rr][ww + 2];
        T[0][rr][ww] = C;
#pragma unroll
        for (int r = 1; r <= EDT_R; r++) {
            const unsigned nl = L | (L << 1) | __funnelshift_r(L, C, 1);
            const unsigned nc = C | __funnelshift_l(L, C, 1) | __funnelshift_r(C, R, 1);
            const unsigned nr = R | __funnelshift_l(C, R, 1) | (R >> 1);
            L = nl; C = nc; R = nr;
            T[r][rr][ww] = C;
        }
    }
    __syncthreads();
    // level walk + write-out: thread = (row, word) = 32 cells = 128 contiguous output bytes
    {
        const int row = tid / TW, ww = tid - row * TW, tr = row + EDT_R;
        const int x = x0 + row, w = w0 + ww;
        unsigned um = 0;
        if (x < W && w < HW) {
            unsigned done = 0;
            unsigned pl[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            edt_levels<0>(&T[0][tr][ww], done, pl);
            int4 *dst = reinterpret_cast<int4 *>(out + ((size_t)x * HW + w) * 32);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                unsigned acc = 0;  // four cells, one byte each
#pragma unroll
                for (int k = 0; k < 8; k++) acc |= ((((pl[k] >> (4 * j)) & 0xFu) * 0x00204081u) & 0x01010101u) << k;
                const unsigned dn = done >> (4 * j);
                int4 v;
                v.x = (dn & 1u) ? (int)(acc & 0xFFu) : 0x7FFFFFFF;
                v.y = (dn & 2u) ? (int)((acc >> 8) & 0xFFu) : 0x7FFFFFFF;
                v.z = (dn & 4u) ? (int)((acc >> 16) & 0xFFu) : 0x7FFFFFFF;
                v.w = (dn & 8u) ? (int)(acc >> 24) : 0x7FFFFFFF;
                __stcs(dst + j, v);
            }
            um = ~done;  // farther than EDT_R from every obstacle: fix-up list
        }
        // The unresolved cells go straight to the global list, one reservation per warp: no block barrier after the walk (its
        // length differs from word to word -- the barrier that used to collect a per-tile list was the top stall of this
        // kernel, r01: 8.2 per issue), warps retire as they finish.  A list longer than fix_cap makes k_edt_fix raise the
        // flag and the windowed path redoes the grid.
        const unsigned lane = (unsigned)tid & 31u;
        const unsigned mine = (unsigned)__popc(um);
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (unsigned)o) incl += t; }
        const unsigned total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (total) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(fix_count, total);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            size_t pos = (size_t)base + (incl - mine);
            while (um) {
                const int b = __ffs(um) - 1;
                um &= um - 1;
                if (pos < fix_cap) { fix_list[2 * pos] = (unsigned)x; fix_list[2 * pos + 1] = (unsigned)(w * 32 + b); }
                pos++;
            }
        }
    }
}

// nearest set bit to column y in one row of the bit grid, looking no farther than `lim` cells; returns the distance or -1
__device__ int edt_row_nearest(const unsigned *__restrict__ rowbits, int HW, int y, int lim)
{
    const int w = y >> 5, b = y & 31;
    int best = -1;
    {
        const unsigned m = rowbits[w];
        const unsigned below = m & (0xFFFFFFFFu >> (31 - b)), above = m >> b;
        if (below) best = b - (31 - __clz(below));
        if (above) { const int d = __ffs(above) - 1; if (best < 0 || d < best) best = d; }
    }
    for (int k = 1; (k - 1) * 32 < lim && (best < 0 || (k - 1) * 32 < best); k++) {
        if (w - k >= 0) { const unsigned m = rowbits[w - k]; if (m) { const int d = y - ((w - k) * 32 + 31 - __clz(m)); if (best < 0 || d < best) best = d; } }
        if (w + k < HW) { const unsigned m = rowbits[w + k]; if (m) { const int d = (w + k) * 32 + __ffs(m) - 1 - y; if (best < 0 || d < best) best = d; } }
    }
    return (best >= 0 && best <= lim) ? best : -1;
}

__global__ void __launch_bounds__(128) k_edt_fix(const unsigned *__restrict__ bits, int32_t *__restrict__ out, int W, int HW,
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
                if (x - d >= 0) { const int g = edt_row_nearest(bits + (size_t)(x - d) * HW, HW, y, lim); if (g >= 0) best = min(best, (long long)d * d + (long long)g * g); }
                if (d > 0 && x + d < W) { const int g = edt_row_nearest(bits + (size_t)(x + d) * HW, HW, y, lim); if (g >= 0) best = min(best, (long long)d * d + (long long)g * g); }
            }
            out[(size_t)x * H + y] = best > 0x7FFFFFFFLL ? 0x7FFFFFFF : (int32_t)best;
        }
        return;
    }
    // short list (small grids): one warp per cell, lane l looks at the rows x +- d for d = l, l + 32, ...; the warp's
    // best so far bounds every lane's search after each round (a listed cell is > EDT_R from everything, so the brute
    // force walks tens of rows: one thread doing that alone was the longest kernel of the whole transform at 1024^2)
    for (unsigned i = warp; i < n; i += nwarps) {
        const int x = (int)fix_list[2 * (size_t)i], y = (int)fix_list[2 * (size_t)i + 1];
        long long best = (long long)1 << 40;
        for (int d0 = 0; (long long)d0 * d0 < best && (x - d0 >= 0 || x + d0 < W); d0 += 32) {
            const int d = d0 + (int)lane;
            long long mine = (long long)1 << 40;
            if ((long long)d * d < best) {
                const long long room = best - (long long)d * d;
                int lim = H;
                if (room < (long long)H * H) { lim = (int)sqrtf((float)room) + 1; if (lim > H) lim = H; }
                if (x - d >= 0) { const int g = edt_row_nearest(bits + (size_t)(x - d) * HW, HW, y, lim); if (g >= 0) mine = min(mine, (long long)d * d + (long long)g * g); }
                if (d > 0 && x + d < W) { const int g = edt_row_nearest(bits + (size_t)(x + d) * HW, HW, y, lim); if (g >= 0) mine = min(mine, (long long)d * d + (long long)g * g); }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) mine = min(mine, __shfl_xor_sync(0xFFFFFFFFu, mine, o));
            best = min(best, mine);
        }
        if (lane == 0) out[(size_t)x * H + y] = best > 0x7FFFFFFFLL ? 0x7FFFFFFF : (int32_t)best;
    }
}

static int edt_upload_levels(fx_context *ctx)
{
    static bool done[64] = {false};
    if (ctx->device >= 0 && ctx->device < 64 && done[ctx->device]) return FX_OK;
    FX_CUDA(ctx, cudaFuncSetAttribute(k_edt_bits, cudaFuncAttributeMaxDynamicSharedMemorySize, EDT_BITS_SMEM));
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
    FX_CUDA(ctx, cudaMemsetAsync(ctx->edt_flag, 0, 4 * sizeof(int), st));
    if (H % 32 == 0 && ((uintptr_t)occ & 15u) == 0 && ((uintptr_t)dist2 & 15u) == 0) {
        // dense-map fast path; falls through to the windowed path (conditionally, on the device flag) if too many
        // cells are farther than EDT_R from every obstacle
        int rc = edt_upload_levels(ctx);
        if (rc) return rc;
        const int HW = H / 32;
        const size_t nw = (size_t)W * HW;
        unsigned *bitsb = reinterpret_cast<unsigned *>(ctx->edt_s);        // scratch reuse: cells*2 bytes >= cells/8
        unsigned *fix_list = reinterpret_cast<unsigned *>(ctx->edt_t);     // cells*2 bytes -> cells/4 entries of 8 bytes
        unsigned *fix_count = reinterpret_cast<unsigned *>(ctx->edt_flag) + 1;
        size_t cap = cells / 4; if (cap > (1u << 22)) cap = 1u << 22;
        FX_CUDA(ctx, cudaMemsetAsync(fix_count, 0, sizeof(unsigned), st));
        int pb = (int)((nw + 255) / 256); if (pb > ctx->sm_count * 16) pb = ctx->sm_count * 16;
        k_edt_pack<<<pb, 256, 0, st>>>(occ, bitsb, nw);
        FX_LAUNCH_CHECK(ctx);
        dim3 gt((HW + EDT_BT_WORDS - 1) / EDT_BT_WORDS, (W + EDT_BT_ROWS - 1) / EDT_BT_ROWS);
        k_edt_bits<<<gt, EDT_BT_ROWS * EDT_BT_WORDS, EDT_BITS_SMEM, st>>>(bitsb, dist2, W, HW, fix_list, fix_count, (unsigned)cap, ctx->edt_flag);
        FX_LAUNCH_CHECK(ctx);
        k_edt_fix<<<ctx->sm_count * 4, 128, 0, st>>>(bitsb, dist2, W, HW, fix_list, fix_count, (unsigned)cap, ctx->edt_flag);
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
        dim3 g2((H + EDT_TYC - 1) / EDT_TYC, (W + EDT_TX - 1) / EDT_TX);
        k_edt_cols_tile<<<g2, 256, 0, st>>>(ctx->edt_g, dist2, W, H, ctx->edt_flag + 2, ctx->edt_flag);
        FX_LAUNCH_CHECK(ctx);
        k_edt_cols_exact<<<(H + 127) / 128, 128, 0, st>>>(ctx->edt_g, dist2, W, H, ctx->edt_s, ctx->edt_t, ctx->edt_flag + 2);
        FX_LAUNCH_CHECK(ctx);
        return FX_OK;
    }
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
        dim3 gt((H + EDT_TYC - 1) / EDT_TYC, (W + EDT_TX - 1) / EDT_TX);
        k_edt_cols_tile<<<gt, 256, 0, st>>>(g, dist2, W, H, ctx->edt_flag, nullptr);
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
