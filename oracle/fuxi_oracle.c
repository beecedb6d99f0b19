/*
 * fuxi_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's global-planning hot path, used only as
 * the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under fuxi_planner_b200/ may link or call this.
 *
 * What is restated (reference file:line, relative to the fuxi-planner repo):
 *   - move legality            scripts/jps1.py:14-31   (blocked)
 *   - squeezed-diagonal test   scripts/jps1.py:34-38   (dblock)
 *   - JPS neighbour pruning    scripts/jps1.py:49-93   (nodeNeighbours)
 *   - jump                     scripts/jps1.py:95-164  (jump)
 *   - A* over jump points      scripts/jps1.py:183-230 (method) + :232-246 (lenght) + :3-12 (heuristic)
 *   - square / 9-point dilation scripts/global_planner_st.py:256-262, global_planner_ccst.py:442-448
 *   - PCL conditioning chain   src/chen_filter_rgb.cpp:52-71 (PassThrough, VoxelGrid, RadiusOutlierRemoval; PCL is
 *                              un-vendored: published algorithm restated, PARITY UNPINNED)
 * plus a second, independent oracle: exact Dijkstra (Dial buckets) on the 8-connected
 * graph whose edges are "not blocked(c, d)" -- the graph-equivalence property SURVEY.md
 * section 0 verified against the reference.
 *
 * Parity pin: tests/golden/jps1_golden.json holds costs AND jump-point paths produced by
 * the unmodified reference (tests/golden/make_golden.py, run in the build container);
 * tests/test_oracle.py checks this file against every one of them.
 *
 * Grid convention: occ[x*H + y], W = x extent, H = y extent (y fastest), value 1 = the
 * reference's "matrix[x][y] == 1".  Anything else is free (the reference tests == 1 only).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    const uint8_t *occ;
    int W, H;
} grid_t;

/* scripts/jps1.py:14-31.  (dX,dY) == (0,0) falls to the "straight, dX == 0" arm and tests the cell itself. */
static inline int blocked(const grid_t *g, int cX, int cY, int dX, int dY)
{
    int tx = cX + dX, ty = cY + dY;
    if (tx < 0 || tx >= g->W) return 1;
    if (ty < 0 || ty >= g->H) return 1;
    const uint8_t *m = g->occ;
    int H = g->H;
    if (dX != 0 && dY != 0) {
        if (m[tx * H + cY] == 1 && m[cX * H + ty] == 1) return 1;
        if (m[tx * H + ty] == 1) return 1;
        return 0;
    }
    if (dX != 0) return m[tx * H + cY] == 1;
    return m[cX * H + ty] == 1;
}

/* scripts/jps1.py:34-38 (no bounds test in the reference; the callers guarantee the cells exist) */
static inline int dblock(const grid_t *g, int cX, int cY, int dX, int dY)
{
    return g->occ[(cX - dX) * g->H + cY] == 1 && g->occ[cX * g->H + (cY - dY)] == 1;
}

/* straight part of scripts/jps1.py:95-164; returns 1 and the jump point, or 0 for None */
static int jump_straight(const grid_t *g, int cX, int cY, int dX, int dY, int gx, int gy, int *rx, int *ry)
{
    int nX = cX + dX, nY = cY + dY;
    if (blocked(g, nX, nY, 0, 0)) return 0;
    if (nX == gx && nY == gy) { *rx = nX; *ry = nY; return 1; }
    int oX = nX, oY = nY;
    if (dX != 0) {
        for (;;) {
            if ((!blocked(g, oX, nY, dX, 1) && blocked(g, oX, nY, 0, 1)) ||
                (!blocked(g, oX, nY, dX, -1) && blocked(g, oX, nY, 0, -1))) {
                *rx = oX; *ry = nY; return 1;
            }
            oX += dX;
            if (blocked(g, oX, nY, 0, 0)) return 0;
            if (oX == gx && nY == gy) { *rx = oX; *ry = nY; return 1; }
        }
    } else {
        for (;;) {
            if ((!blocked(g, nX, oY, 1, dY) && blocked(g, nX, oY, 1, 0)) ||
                (!blocked(g, nX, oY, -1, dY) && blocked(g, nX, oY, -1, 0))) {
                *rx = nX; *ry = oY; return 1;
            }
            oY += dY;
            if (blocked(g, nX, oY, 0, 0)) return 0;
            if (nX == gx && oY == gy) { *rx = nX; *ry = oY; return 1; }
        }
    }
}

/* scripts/jps1.py:95-164 */
static int jump(const grid_t *g, int cX, int cY, int dX, int dY, int gx, int gy, int *rx, int *ry)
{
    if (dX == 0 || dY == 0) return jump_straight(g, cX, cY, dX, dY, gx, gy, rx, ry);
    int nX = cX + dX, nY = cY + dY;
    if (blocked(g, nX, nY, 0, 0)) return 0;
    if (nX == gx && nY == gy) { *rx = nX; *ry = nY; return 1; }
    int oX = nX, oY = nY, tx, ty;
    for (;;) {
        if ((!blocked(g, oX, oY, -dX, dY) && blocked(g, oX, oY, -dX, 0)) ||
            (!blocked(g, oX, oY, dX, -dY) && blocked(g, oX, oY, 0, -dY))) {
            *rx = oX; *ry = oY; return 1;
        }
        if (jump_straight(g, oX, oY, dX, 0, gx, gy, &tx, &ty) ||
            jump_straight(g, oX, oY, 0, dY, gx, gy, &tx, &ty)) {
            *rx = oX; *ry = oY; return 1;
        }
        oX += dX; oY += dY;
        if (blocked(g, oX, oY, 0, 0)) return 0;
        if (dblock(g, oX, oY, dX, dY)) return 0;
        if (oX == gx && oY == gy) { *rx = oX; *ry = oY; return 1; }
    }
}

static inline int sgn(int v) { return (v > 0) - (v < 0); }

/* scripts/jps1.py:49-93; writes up to 8 neighbours, returns the count */
static int node_neighbours(const grid_t *g, int cX, int cY, int has_parent, int pX, int pY, int *nx, int *ny)
{
    int n = 0;
    if (!has_parent) {
        static const int di[8] = {-1, 0, 1, 0, -1, -1, 1, 1};
        static const int dj[8] = {0, -1, 0, 1, -1, 1, -1, 1};
        for (int k = 0; k < 8; k++)
            if (!blocked(g, cX, cY, di[k], dj[k])) { nx[n] = cX + di[k]; ny[n] = cY + dj[k]; n++; }
        return n;
    }
    int dX = sgn(cX - pX), dY = sgn(cY - pY);
    if (dX != 0 && dY != 0) {
        int by = blocked(g, cX, cY, 0, dY), bx = blocked(g, cX, cY, dX, 0);
        if (!by) { nx[n] = cX; ny[n] = cY + dY; n++; }
        if (!bx) { nx[n] = cX + dX; ny[n] = cY; n++; }
        if ((!by || !bx) && !blocked(g, cX, cY, dX, dY)) { nx[n] = cX + dX; ny[n] = cY + dY; n++; }
        if (blocked(g, cX, cY, -dX, 0) && !by) { nx[n] = cX - dX; ny[n] = cY + dY; n++; }
        if (blocked(g, cX, cY, 0, -dY) && !bx) { nx[n] = cX + dX; ny[n] = cY - dY; n++; }
    } else if (dX == 0) {
        /* jps1.py:77 guards with blocked(c, dX=0, 0), i.e. a test of the cell itself */
        if (!blocked(g, cX, cY, 0, 0)) {
            if (!blocked(g, cX, cY, 0, dY)) { nx[n] = cX; ny[n] = cY + dY; n++; }
            if (blocked(g, cX, cY, 1, 0)) { nx[n] = cX + 1; ny[n] = cY + dY; n++; }
            if (blocked(g, cX, cY, -1, 0)) { nx[n] = cX - 1; ny[n] = cY + dY; n++; }
        }
    } else {
        if (!blocked(g, cX, cY, dX, 0)) {
            nx[n] = cX + dX; ny[n] = cY; n++;
            if (blocked(g, cX, cY, 0, 1)) { nx[n] = cX + dX; ny[n] = cY + 1; n++; }
            if (blocked(g, cX, cY, 0, -1)) { nx[n] = cX + dX; ny[n] = cY - 1; n++; }
        }
    }
    return n;
}

/* scripts/jps1.py:3-12 */
static inline double heuristic(int ax, int ay, int bx, int by, int hchoice)
{
    if (hchoice == 1) {
        double xd = fabs((double)(bx - ax)), yd = fabs((double)(by - ay));
        return xd > yd ? 14 * yd + 10 * (xd - yd) : 14 * xd + 10 * (yd - xd);
    }
    double dx = bx - ax, dy = by - ay;
    return sqrt(dx * dx + dy * dy);
}

/* scripts/jps1.py:232-246 */
static inline double seg_length(int cx, int cy, int jx, int jy, int hchoice)
{
    double lX = fabs((double)(cx - jx)), lY = fabs((double)(cy - jy));
    if (hchoice == 1) {
        int dX = cx != jx, dY = cy != jy;
        if (dX && dY) return lX * 14;
        return (dX * lX + dY * lY) * 10;
    }
    double dx = cx - jx, dy = cy - jy;
    return sqrt(dx * dx + dy * dy);
}

/* min-heap of (f, x, y) tuples, lexicographic like Python's tuple order used by heapq (jps1.py:192,228) */
typedef struct { double f; int x, y; } hent_t;
typedef struct { hent_t *a; size_t n, cap; } heap_t;

static inline int hless(const hent_t *p, const hent_t *q)
{
    if (p->f != q->f) return p->f < q->f;
    if (p->x != q->x) return p->x < q->x;
    return p->y < q->y;
}
static int heap_push(heap_t *h, hent_t e)
{
    if (h->n == h->cap) {
        size_t nc = h->cap ? h->cap * 2 : 1024;
        hent_t *na = (hent_t *)realloc(h->a, nc * sizeof(hent_t));
        if (!na) return -1;
        h->a = na; h->cap = nc;
    }
    size_t i = h->n++;
    while (i > 0) {
        size_t p = (i - 1) / 2;
        if (!hless(&e, &h->a[p])) break;
        h->a[i] = h->a[p]; i = p;
    }
    h->a[i] = e;
    return 0;
}
static hent_t heap_pop(heap_t *h)
{
    hent_t top = h->a[0], last = h->a[--h->n];
    size_t i = 0;
    for (;;) {
        size_t c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && hless(&h->a[c + 1], &h->a[c])) c++;
        if (!hless(&h->a[c], &last)) break;
        h->a[i] = h->a[c]; i = c;
    }
    if (h->n) h->a[i] = last;
    return top;
}

/* per-thread scratch so batches do not pay O(W*H) per query */
typedef struct {
    int cells;
    double *g;        /* gscore; valid iff seen[] */
    int32_t *parent;  /* came_from as cell index, -1 = none */
    uint8_t *seen;    /* bit0: has gscore, bit1: in close_set */
    int32_t *inheap;  /* number of live heap entries naming this cell (jps1.py:224 membership test) */
    int32_t *touched; size_t ntouched, captouched;
    heap_t heap;
} jps_ws_t;

static jps_ws_t *ws_new(int cells)
{
    jps_ws_t *w = (jps_ws_t *)calloc(1, sizeof(*w));
    if (!w) return NULL;
    w->cells = cells;
    w->g = (double *)malloc(sizeof(double) * (size_t)cells);
    w->parent = (int32_t *)malloc(sizeof(int32_t) * (size_t)cells);
    w->seen = (uint8_t *)calloc((size_t)cells, 1);
    w->inheap = (int32_t *)calloc((size_t)cells, sizeof(int32_t));
    w->captouched = 4096;
    w->touched = (int32_t *)malloc(sizeof(int32_t) * w->captouched);
    return w;
}
static void ws_free(jps_ws_t *w)
{
    if (!w) return;
    free(w->g); free(w->parent); free(w->seen); free(w->inheap); free(w->touched); free(w->heap.a); free(w);
}
static inline void ws_touch(jps_ws_t *w, int c)
{
    if (w->ntouched == w->captouched) {
        w->captouched *= 2;
        w->touched = (int32_t *)realloc(w->touched, sizeof(int32_t) * w->captouched);
    }
    w->touched[w->ntouched++] = c;
}
static void ws_reset(jps_ws_t *w)
{
    for (size_t i = 0; i < w->ntouched; i++) { int c = w->touched[i]; w->seen[c] = 0; w->inheap[c] = 0; }
    w->ntouched = 0; w->heap.n = 0;
}

/*
 * scripts/jps1.py:183-230.  Returns 1 = path found (cost = gscore[goal], the value the reference
 * prints at :207), 0 = exhausted (reference returns (0, t)), <0 = argument/alloc error.
 * path_xy receives up to max_path (x,y) jump points start..goal; *path_len gets the full count.
 * *expansions (optional) counts heap pops.  Out-of-range start raises IndexError in the reference;
 * here it is -2.
 */
static int jps_run(jps_ws_t *w, const grid_t *g, int sx, int sy, int gx, int gy, int hchoice,
                   double *cost, int32_t *path_xy, int max_path, int32_t *path_len, int64_t *expansions)
{
    int H = g->H;
    if (sx < 0 || sx >= g->W || sy < 0 || sy >= g->H) return -2;
    int s = sx * H + sy;
    ws_reset(w);
    w->g[s] = 0.0; w->parent[s] = -1; w->seen[s] = 1; ws_touch(w, s);
    hent_t e0 = {heuristic(sx, sy, gx, gy, hchoice), sx, sy};
    if (heap_push(&w->heap, e0)) return -3;
    w->inheap[s] = 1;
    int64_t pops = 0;
    while (w->heap.n) {
        hent_t cur = heap_pop(&w->heap);
        int cX = cur.x, cY = cur.y, c = cX * H + cY;
        w->inheap[c]--;
        pops++;
        if (cX == gx && cY == gy) {
            int n = 0;
            for (int v = c; v >= 0; v = w->parent[v]) n++;
            if (path_len) *path_len = n;
            if (path_xy) {
                int i = n - 1;
                for (int v = c; v >= 0; v = w->parent[v], i--)
                    if (i < max_path) { path_xy[2 * i] = v / H; path_xy[2 * i + 1] = v % H; }
            }
            if (cost) *cost = w->g[c];
            if (expansions) *expansions = pops;
            return 1;
        }
        w->seen[c] |= 2;
        int nx[8], ny[8];
        int has_parent = w->parent[c] >= 0;
        int pX = has_parent ? w->parent[c] / H : 0, pY = has_parent ? w->parent[c] % H : 0;
        int nn = node_neighbours(g, cX, cY, has_parent, pX, pY, nx, ny);
        for (int k = 0; k < nn; k++) {
            int jx, jy;
            if (!jump(g, cX, cY, nx[k] - cX, ny[k] - cY, gx, gy, &jx, &jy)) continue;
            int j = jx * H + jy;
            if (w->seen[j] & 2) continue;
            double tentative = w->g[c] + seg_length(cX, cY, jx, jy, hchoice);
            double known = (w->seen[j] & 1) ? w->g[j] : 0.0; /* gscore.get(jp, 0) */
            if (tentative < known || w->inheap[j] == 0) {
                if (!w->seen[j]) ws_touch(w, j);
                w->parent[j] = c; w->g[j] = tentative; w->seen[j] |= 1;
                hent_t e = {tentative + heuristic(jx, jy, gx, gy, hchoice), jx, jy};
                if (heap_push(&w->heap, e)) return -3;
                w->inheap[j]++;
            }
        }
    }
    if (expansions) *expansions = pops;
    return 0;
}

int fxo_jps(const uint8_t *occ, int W, int H, int sx, int sy, int gx, int gy, int hchoice,
            double *cost, int32_t *path_xy, int max_path, int32_t *path_len, int64_t *expansions)
{
    if (!occ || W <= 0 || H <= 0 || (hchoice != 1 && hchoice != 2)) return -1;
    grid_t g = {occ, W, H};
    jps_ws_t *w = ws_new(W * H);
    if (!w) return -3;
    int r = jps_run(w, &g, sx, sy, gx, gy, hchoice, cost, path_xy, max_path, path_len, expansions);
    ws_free(w);
    return r;
}

/* Q independent queries over `threads` host threads (0 = all).  status[q] as fxo_jps; cost[q] valid iff 1. */
int fxo_jps_batch(const uint8_t *occ, int W, int H, const int32_t *starts_xy, const int32_t *goals_xy, int Q,
                  int hchoice, int threads, double *cost, int32_t *status, int64_t *expansions)
{
    if (!occ || W <= 0 || H <= 0 || (hchoice != 1 && hchoice != 2)) return -1;
    grid_t g = {occ, W, H};
    int used = 1;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
    used = threads;
#pragma omp parallel num_threads(threads)
#endif
    {
        jps_ws_t *w = ws_new(W * H);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int q = 0; q < Q; q++) {
            double c = 0; int64_t ex = 0;
            int r = w ? jps_run(w, &g, starts_xy[2 * q], starts_xy[2 * q + 1], goals_xy[2 * q], goals_xy[2 * q + 1],
                                hchoice, &c, NULL, 0, NULL, &ex) : -3;
            status[q] = r; cost[q] = c;
            if (expansions) expansions[q] = ex;
        }
        ws_free(w);
    }
    return used;
}

/* ------------------------------------------------------------------------------------------
 * Second oracle: exact single-source shortest paths on the "not blocked" graph (forward moves
 * out of the source; the source cell itself is never tested, jps1.py:14-31), integer weights
 * ws (straight) / wd (diagonal), Dial buckets via a binary heap of (cost, cell).
 * field[x*H+y] = cost from (sx,sy), or -1 if unreachable.  If (gx,gy) >= 0 stops once the goal
 * is settled (cells not yet settled then hold -1 or a tentative value -- only the goal is exact).
 * ------------------------------------------------------------------------------------------ */
typedef struct { int64_t d; int32_t c; } dent_t;
typedef struct { dent_t *a; size_t n, cap; } dheap_t;
static int dpush(dheap_t *h, dent_t e)
{
    if (h->n == h->cap) {
        size_t nc = h->cap ? h->cap * 2 : 4096;
        dent_t *na = (dent_t *)realloc(h->a, nc * sizeof(dent_t));
        if (!na) return -1;
        h->a = na; h->cap = nc;
    }
    size_t i = h->n++;
    while (i > 0) { size_t p = (i - 1) / 2; if (h->a[p].d <= e.d) break; h->a[i] = h->a[p]; i = p; }
    h->a[i] = e;
    return 0;
}
static dent_t dpop(dheap_t *h)
{
    dent_t top = h->a[0], last = h->a[--h->n];
    size_t i = 0;
    for (;;) {
        size_t c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && h->a[c + 1].d < h->a[c].d) c++;
        if (h->a[c].d >= last.d) break;
        h->a[i] = h->a[c]; i = c;
    }
    if (h->n) h->a[i] = last;
    return top;
}

int fxo_sssp(const uint8_t *occ, int W, int H, int sx, int sy, int gx, int gy, int64_t ws, int64_t wd,
             int64_t *field, int64_t *settled)
{
    if (!occ || !field || W <= 0 || H <= 0) return -1;
    if (sx < 0 || sx >= W || sy < 0 || sy >= H) return -2;
    grid_t g = {occ, W, H};
    size_t cells = (size_t)W * H;
    for (size_t i = 0; i < cells; i++) field[i] = -1;
    uint8_t *done = (uint8_t *)calloc(cells, 1);
    dheap_t h = {0};
    if (!done) return -3;
    static const int di[8] = {-1, 1, 0, 0, -1, -1, 1, 1};
    static const int dj[8] = {0, 0, -1, 1, -1, 1, -1, 1};
    int s = sx * H + sy, goal = (gx >= 0 && gy >= 0 && gx < W && gy < H) ? gx * H + gy : -1;
    field[s] = 0;
    dent_t e0 = {0, s};
    dpush(&h, e0);
    int64_t ns = 0;
    int found = 0;
    while (h.n) {
        dent_t e = dpop(&h);
        if (done[e.c] || e.d != field[e.c]) continue;
        done[e.c] = 1; ns++;
        if (e.c == goal) { found = 1; break; }
        int cx = e.c / H, cy = e.c % H;
        for (int k = 0; k < 8; k++) {
            if (blocked(&g, cx, cy, di[k], dj[k])) continue;
            int v = (cx + di[k]) * H + cy + dj[k];
            int64_t nd = e.d + (k < 4 ? ws : wd);
            if (field[v] < 0 || nd < field[v]) { field[v] = nd; dent_t ne = {nd, v}; if (dpush(&h, ne)) { free(done); free(h.a); return -3; } }
        }
    }
    if (settled) *settled = ns;
    free(done); free(h.a);
    return goal >= 0 ? found : 1;
}

/* goal-directed batch of the same (threads as above); cost[q] = -1 if unreachable */
int fxo_sssp_batch(const uint8_t *occ, int W, int H, const int32_t *starts_xy, const int32_t *goals_xy, int Q,
                   int64_t ws, int64_t wd, int threads, int64_t *cost)
{
    if (!occ || W <= 0 || H <= 0) return -1;
    int used = 1;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
    used = threads;
#pragma omp parallel num_threads(threads)
#endif
    {
        int64_t *field = (int64_t *)malloc(sizeof(int64_t) * (size_t)W * H);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int q = 0; q < Q; q++) {
            int gx = goals_xy[2 * q], gy = goals_xy[2 * q + 1];
            int r = field ? fxo_sssp(occ, W, H, starts_xy[2 * q], starts_xy[2 * q + 1], gx, gy, ws, wd, field, NULL) : -3;
            cost[q] = (r == 1 && gx >= 0 && gx < W && gy >= 0 && gy < H) ? field[gx * H + gy] : -1;
        }
        free(field);
    }
    return used;
}

/* scripts/global_planner_st.py:256-262 (step = radius -> 9-point stencil {-r,0,+r}^2) and
 * scripts/global_planner_ccst.py:442-448 (step = 1 -> dense (2r+1)^2).  Sources are cells > 0;
 * every stencil target becomes 1, everything else 0 ... except that untouched cells keep their
 * input value in the reference; since all sources are themselves targets (offset 0,0) the only
 * surviving values are 0 and 1.  Targets outside the array are dropped (the reference pads so
 * that none exist, :230-250). */
int fxo_inflate(const uint8_t *in, uint8_t *out, int W, int H, int radius, int step)
{
    if (!in || !out || W <= 0 || H <= 0 || radius < 0) return -1;
    if (step <= 0) step = 1;
    memset(out, 0, (size_t)W * H);
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            if (in[(size_t)x * H + y] == 0) continue;
            if (radius == 0) { out[(size_t)x * H + y] = 1; continue; }
            for (int i = -radius; i <= radius; i += step)
                for (int j = -radius; j <= radius; j += step) {
                    int tx = x + i, ty = y + j;
                    if (tx < 0 || tx >= W || ty < 0 || ty >= H) continue;
                    out[(size_t)tx * H + ty] = 1;
                }
        }
    return 0;
}

/* exact squared Euclidean distance to the nearest cell > 0 (not in the reference: north-star addition;
 * checked against scipy in tests).  Felzenszwalb-Huttenlocher two-pass lower envelope in int64.
 * dist2 = INT32_MAX where the grid holds no occupied cell. */
int fxo_edt(const uint8_t *occ, int32_t *dist2, int W, int H)
{
    if (!occ || !dist2 || W <= 0 || H <= 0) return -1;
    const int64_t INF = (int64_t)1 << 40;
    int64_t *g = (int64_t *)malloc(sizeof(int64_t) * (size_t)W * H);
    int n = W > H ? W : H;
    int *v = (int *)malloc(sizeof(int) * (size_t)n);
    double *z = (double *)malloc(sizeof(double) * ((size_t)n + 1));
    int64_t *f = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    if (!g || !v || !z || !f) { free(g); free(v); free(z); free(f); return -3; }
    for (int x = 0; x < W; x++) { /* pass 1 along y */
        int64_t *row = g + (size_t)x * H;
        int last = -1;
        for (int y = 0; y < H; y++) {
            if (occ[(size_t)x * H + y]) last = y;
            row[y] = last < 0 ? INF : (int64_t)(y - last) * (y - last);
        }
        last = -1;
        for (int y = H - 1; y >= 0; y--) {
            if (occ[(size_t)x * H + y]) last = y;
            if (last >= 0) { int64_t d = (int64_t)(last - y) * (last - y); if (d < row[y]) row[y] = d; }
        }
    }
    for (int y = 0; y < H; y++) { /* pass 2 along x: brute-force-exact lower envelope in integers */
        for (int x = 0; x < W; x++) f[x] = g[(size_t)x * H + y];
        int k = -1;
        for (int q = 0; q < W; q++) {
            if (f[q] >= INF) continue;
            while (k >= 0) {
                /* intersection of parabolas v[k] and q: s = ((f[q]+q^2)-(f[v]+v^2)) / (2q-2v) */
                double s = ((double)(f[q] + (int64_t)q * q) - (double)(f[v[k]] + (int64_t)v[k] * v[k])) / (2.0 * q - 2.0 * v[k]);
                if (s <= z[k]) k--; else { z[k + 1] = s; break; }
            }
            if (k < 0) z[0] = -1e30;
            k++; v[k] = q; z[k + 1] = 1e30;
            if (k == 0) z[0] = -1e30;
        }
        if (k < 0) { for (int x = 0; x < W; x++) dist2[(size_t)x * H + y] = INT32_MAX; continue; }
        int j = 0;
        for (int x = 0; x < W; x++) {
            while (z[j + 1] < x) j++;
            /* guard against double rounding at envelope ties: test the neighbours too */
            int64_t best = (int64_t)(x - v[j]) * (x - v[j]) + f[v[j]];
            if (j > 0) { int64_t b = (int64_t)(x - v[j - 1]) * (x - v[j - 1]) + f[v[j - 1]]; if (b < best) best = b; }
            if (j < k) { int64_t b = (int64_t)(x - v[j + 1]) * (x - v[j + 1]) + f[v[j + 1]]; if (b < best) best = b; }
            dist2[(size_t)x * H + y] = best > INT32_MAX ? INT32_MAX : (int32_t)best;
        }
    }
    free(g); free(v); free(z); free(f);
    return 0;
}

/* ---- upstream cloud conditioning: PassThrough -> VoxelGrid -> RadiusOutlierRemoval ------------------------------
 * src/chen_filter_rgb.cpp:52-71 calls PCL (un-vendored, unpinned: CMakeLists.txt:10-18) -> PARITY UNPINNED.  This
 * restates PCL's published algorithms (pcl/filters/impl/passthrough.hpp, voxel_grid.hpp, radius_outlier_removal.hpp,
 * FLANN L2_Simple + RadiusResultSet) with the one deliberate change include/fuxi_b200.h documents: voxel centroids
 * accumulate in 2^-24 fixed point (order independent) because PCL's own float sums depend on an unstable sort.
 * out: float [cap][4] = x, y, z, packed rgb word; counts = {after PassThrough, voxels, kept, status}. */
typedef struct { uint32_t idx; int32_t pt; } vox_ent_t;
static int vox_cmp(const void *a, const void *b)
{
    const vox_ent_t *p = (const vox_ent_t *)a, *q = (const vox_ent_t *)b;
    if (p->idx != q->idx) return p->idx < q->idx ? -1 : 1;
    return (p->pt > q->pt) - (p->pt < q->pt);
}
int fxo_cloud_filter(const float *pts, int64_t n, int stride, int rgb_off, float lo, float hi, float leaf_x, float leaf_y,
                     float leaf_z, double radius, int min_nb, float *out, int64_t cap, int64_t *counts)
{
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    vox_ent_t *e = (vox_ent_t *)malloc(sizeof(vox_ent_t) * (size_t)(n > 0 ? n : 1));
    if (!e) return -3;
    const float inv[3] = {1.0f / leaf_x, 1.0f / leaf_y, 1.0f / leaf_z};
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        const float *p = pts + i * stride;
        if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
        if (p[2] < lo || p[2] > hi) continue; /* passthrough.hpp: removed iff value < min || value > max */
        e[m++].pt = (int32_t)i;
        for (int k = 0; k < 3; k++) {
            if (p[k] < mn[k]) mn[k] = p[k];
            if (p[k] > mx[k]) mx[k] = p[k];
        }
    }
    counts[0] = m;
    if (m == 0) { free(e); return 0; }
    int min_b[3], div[3];
    uint64_t total = 1;
    for (int k = 0; k < 3; k++) {
        float a = floorf(mn[k] * inv[k]), b = floorf(mx[k] * inv[k]);
        if (!(fabsf(a) < 1.0e9f) || !(fabsf(b) < 1.0e9f)) { counts[3] = -1; free(e); return 0; }
        min_b[k] = (int)a;
        div[k] = (int)b - (int)a + 1;
        total *= (uint64_t)div[k];
        if (total > (1ull << 40)) break;
    }
    if (total > (1ull << 31)) { counts[3] = (int64_t)total; free(e); return 0; }
    for (int64_t t = 0; t < m; t++) {
        const float *p = pts + (int64_t)e[t].pt * stride;
        int i = (int)(floorf(p[0] * inv[0]) - (float)min_b[0]);
        int j = (int)(floorf(p[1] * inv[1]) - (float)min_b[1]);
        int k = (int)(floorf(p[2] * inv[2]) - (float)min_b[2]);
        e[t].idx = (uint32_t)i + (uint32_t)div[0] * ((uint32_t)j + (uint32_t)div[1] * (uint32_t)k);
    }
    qsort(e, (size_t)m, sizeof(vox_ent_t), vox_cmp);
    float *vox = (float *)malloc(sizeof(float) * 4 * (size_t)m);
    if (!vox) { free(e); return -3; }
    int64_t nv = 0;
    for (int64_t t = 0; t < m;) {
        int64_t u = t, sx = 0, sy = 0, sz = 0;
        uint64_t sr = 0, sg = 0, sb = 0;
        for (; u < m && e[u].idx == e[t].idx; u++) {
            const float *p = pts + (int64_t)e[u].pt * stride;
            sx += llrint((double)p[0] * 16777216.0);
            sy += llrint((double)p[1] * 16777216.0);
            sz += llrint((double)p[2] * 16777216.0);
            if (rgb_off >= 0) {
                uint32_t c;
                memcpy(&c, p + rgb_off, 4);
                sr += (c >> 16) & 255u, sg += (c >> 8) & 255u, sb += c & 255u;
            }
        }
        const double dn = (double)(u - t);
        float *o = vox + 4 * nv++;
        o[0] = (float)(((double)sx / dn) * (1.0 / 16777216.0));
        o[1] = (float)(((double)sy / dn) * (1.0 / 16777216.0));
        o[2] = (float)(((double)sz / dn) * (1.0 / 16777216.0));
        uint32_t c = ((uint32_t)(sr / (uint64_t)(u - t)) << 16) | ((uint32_t)(sg / (uint64_t)(u - t)) << 8) | (uint32_t)(sb / (uint64_t)(u - t));
        memcpy(o + 3, &c, 4);
        t = u;
    }
    counts[1] = nv;
    /* radius_outlier_removal.hpp: k = radiusSearch(point, radius) includes the point; outlier iff k <= min_pts */
    const float r2 = (float)(radius * radius);
    uint8_t *keep = (uint8_t *)calloc((size_t)nv, 1);
    if (!keep) { free(e); free(vox); return -3; }
#pragma omp parallel for schedule(static)
    for (int64_t a = 0; a < nv; a++) {
        int k = 0;
        const float *c = vox + 4 * a;
        for (int64_t b = 0; b < nv; b++) {
            const float *q = vox + 4 * b;
            const float ex = c[0] - q[0], ey = c[1] - q[1], ez = c[2] - q[2];
            const float d2 = (ex * ex + ey * ey) + ez * ez;
            k += d2 < r2;
        }
        keep[a] = k > min_nb;
    }
    int64_t kept = 0;
    for (int64_t a = 0; a < nv; a++)
        if (keep[a]) {
            if (kept < cap) memcpy(out + 4 * kept, vox + 4 * a, 16);
            kept++;
        }
    counts[2] = kept;
    free(e); free(vox); free(keep);
    return 0;
}

/* scripts/jps1.py:95-164 on its own (tests of the jump-point path form): returns 1 and the jump point, or 0 for None */
int fxo_jump(const uint8_t *occ, int W, int H, int cX, int cY, int dX, int dY, int gx, int gy, int32_t *rxy)
{
    grid_t g = {occ, W, H};
    int rx = 0, ry = 0;
    if (!jump(&g, cX, cY, dX, dY, gx, gy, &rx, &ry)) return 0;
    rxy[0] = rx; rxy[1] = ry;
    return 1;
}

int fxo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
