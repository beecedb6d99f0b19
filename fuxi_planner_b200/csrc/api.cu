// api.cu -- context lifetime, error reporting and the host-buffer entry points of libfuxi_b200.so.
#include <stdarg.h>
#include <stdlib.h>

#include <thread>
#include <vector>

#include "common.cuh"

static char g_create_err[512] = "";

int fx_set_err(fx_context *ctx, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx ? ctx->err : g_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" int fx_version(void) { return 1; }

extern "C" const char *fx_last_error(fx_context *ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" int64_t fx_launch_count(fx_context *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int64_t fx_last_d2h_bytes(fx_context *ctx) { return ctx ? ctx->last_d2h_bytes : 0; }

extern "C" int fx_create(int device, fx_context **out)
{
    if (!out) return FX_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fx_set_err(nullptr, FX_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fx_set_err(nullptr, FX_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    fx_context *ctx = (fx_context *)calloc(1, sizeof(fx_context));
    if (!ctx) return FX_ERR_NOMEM;
    ctx->device = device;
    cudaDeviceProp prop;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        fx_set_err(nullptr, FX_ERR_CUDA, "cudaSetDevice/GetDeviceProperties: %s", cudaGetErrorString(e));
        free(ctx);
        return FX_ERR_CUDA;
    }
    if (prop.major < 10) {
        fx_set_err(nullptr, FX_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        free(ctx);
        return FX_ERR_UNSUPPORTED;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cfg_wide_below = -1;
    ctx->cfg_cluster = 1;
    if (const char *e = getenv("FUXI_B200_CLUSTER")) ctx->cfg_cluster = e[0] != '0';  // tests run the latency form both ways
    if (const char *e = getenv("FUXI_B200_WIDE_BELOW")) ctx->cfg_wide_below = atoi(e);  // tuning experiments only
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    e = cudaMalloc(&ctx->counters, 16 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->counters, 0, 16 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->fstate, 32 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(ctx->fstate, 0, 32 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->edt_flag, 4 * sizeof(int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_search[0]);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_search[1]);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_band[0]);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_band[1]);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();  // the memsets above ran on the legacy stream; own_stream does not wait for it
    if (e != cudaSuccess) {
        fx_set_err(nullptr, FX_ERR_CUDA, "context allocation: %s", cudaGetErrorString(e));
        fx_destroy(ctx);
        return FX_ERR_CUDA;
    }
    *out = ctx;
    return FX_OK;
}

extern "C" int fx_destroy(fx_context *ctx)
{
    if (!ctx) return FX_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    fx_search_release(ctx, 0);
    fx_search_release(ctx, 1);
    void *dev[] = {ctx->moves, ctx->counters, ctx->fq, ctx->fstate,
                   ctx->seeds, ctx->seeds_sorted, ctx->seed_hist, ctx->edt_g, ctx->edt_s, ctx->edt_t, ctx->edt_flag,
                   ctx->d_grid, ctx->d_grid2, ctx->d_q, ctx->d_out_i, ctx->d_out_f, ctx->d_path, ctx->d_pts, ctx->proj_bits, ctx->q_order, ctx->q_ubound, ctx->bfields,
                   ctx->d_msg, ctx->d_rp, ctx->d_cpath, ctx->d_coff, ctx->cl_bits, ctx->cl_gpref, ctx->cl_chunk, ctx->cl_vidx, ctx->cl_keep, ctx->cl_gpref2,
                   ctx->cl_chunk2, ctx->cl_acc, ctx->cl_vox, ctx->cl_state, ctx->cl_out, ctx->df_rec, ctx->tf_q};
    for (size_t i = 0; i < sizeof(dev) / sizeof(dev[0]); i++)
        if (dev[i]) cudaFree(dev[i]);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->ev_search[0]) cudaEventDestroy(ctx->ev_search[0]);
    if (ctx->ev_search[1]) cudaEventDestroy(ctx->ev_search[1]);
    if (ctx->ev_band[0]) cudaEventDestroy(ctx->ev_band[0]);
    if (ctx->ev_band[1]) cudaEventDestroy(ctx->ev_band[1]);
    free(ctx);
    return FX_OK;
}

extern "C" int fx_set_search_tuning(fx_context *ctx, int slots, int band0)
{
    if (!ctx || slots < 0 || band0 < 0) return FX_ERR_ARG;
    if (slots != ctx->cfg_slots) {  // force re-allocation of the per-slot scratch
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        fx_search_release(ctx, 0);
        fx_search_release(ctx, 1);
    }
    ctx->cfg_slots = slots;
    ctx->cfg_band0 = band0;
    return FX_OK;
}

extern "C" int fx_set_search_form(fx_context *ctx, int form)
{
    if (!ctx || form < 0 || form > 2) return FX_ERR_ARG;
    ctx->cfg_wide_below = form == 0 ? -1 : (form == 1 ? 0 : 0x7FFFFFFF);
    return FX_OK;
}

int fx_grow_bytes(fx_context *ctx, void **p, size_t *cap, size_t want_bytes)
{
    if (*cap >= want_bytes && *p) return FX_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    FX_CUDA(ctx, cudaMalloc(p, want_bytes));
    *cap = want_bytes;
    return FX_OK;
}
template <typename T>
static int grow(fx_context *ctx, T **p, size_t *cap, size_t want_bytes)
{
    return fx_grow_bytes(ctx, (void **)p, cap, want_bytes);
}
int fx_grow_pinned(fx_context *ctx, size_t want)
{
    if (ctx->h_pin_cap >= want && ctx->h_pin) return FX_OK;
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    ctx->h_pin = nullptr; ctx->h_pin_cap = 0;
    FX_CUDA(ctx, cudaMallocHost(&ctx->h_pin, want));
    ctx->h_pin_cap = want;
    return FX_OK;
}
static int grow_pinned(fx_context *ctx, size_t want) { return fx_grow_pinned(ctx, want); }

// Host-buffer planning call: H2D (grid + queries) -> search -> D2H (costs, path lengths, paths).
// Replaces the whole of `jps1.method(...)` for Q queries (scripts/jps1.py:183-230) from the caller's view.
// The reference hands jps1.method a float64 matrix (np.zeros, scripts/global_planner_st.py:248) and tests `== 1`
// (scripts/jps1.py:20-29).  Converting 16 Mi cells with numpy costs 15 ms on the host (r01f), more than the search:
// the f64 entry point does `== 1.0 -> uint8` straight into the pinned staging buffer with a few host threads,
// chunk by chunk, each chunk's H2D copy overlapping the conversion of the next.
static void convert_f64_chunk(const double *src, uint8_t *dst, size_t n, int nthreads)
{
    auto work = [=](size_t a, size_t b) {
        for (size_t i = a; i < b; i++) dst[i] = src[i] == 1.0 ? (uint8_t)1 : (uint8_t)0;
    };
    if (nthreads <= 1 || n < (1u << 18)) { work(0, n); return; }
    std::vector<std::thread> th;
    const size_t per = (n + nthreads - 1) / nthreads;
    for (int t = 1; t < nthreads; t++) {
        const size_t a = (size_t)t * per, b = a + per < n ? a + per : n;
        if (a < b) th.emplace_back(work, a, b);
    }
    work(0, per < n ? per : n);
    for (auto &t : th) t.join();
}

// pageable <-> pinned staging copies of tens of MB (the grid in, the path buffers out): a few host threads instead of
// one memcpy (67 MB of path buffers per 8192-query batch took 8 ms single-threaded, 6 % of the end-to-end step)
static void par_memcpy(void *dst, const void *src, size_t bytes)
{
    unsigned hc = std::thread::hardware_concurrency();
    const int nthreads = (int)(hc == 0 ? 1 : (hc > 8 ? 8 : hc));
    if (nthreads <= 1 || bytes < ((size_t)4 << 20)) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t per = ((bytes + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
    for (int t = 1; t < nthreads; t++) {
        const size_t a = (size_t)t * per;
        if (a >= bytes) break;
        const size_t nb = a + per < bytes ? per : bytes - a;
        th.emplace_back([=]() { memcpy((char *)dst + a, (const char *)src + a, nb); });
    }
    memcpy(dst, src, per < bytes ? per : bytes);
    for (auto &t : th) t.join();
}

// csr != NULL: paths are returned packed (offsets[Q+1] + xy[total][2]) instead of in padded rows
struct CsrOut { int64_t *h_offsets; int32_t *h_xy; int64_t cap; int64_t *h_total; };
static int plan_host_impl(fx_context *ctx, const uint8_t *h_grid, const double *h_matrix, int W, int H, const int32_t *h_starts_xy,
                          const int32_t *h_goals_xy, int Q, int metric, int32_t *h_cost_i, double *h_cost_f,
                          int32_t *h_path_xy, int32_t *h_path_len, int max_path, const CsrOut *csr);

extern "C" int fx_plan_host(fx_context *ctx, const uint8_t *h_grid, int W, int H, const int32_t *h_starts_xy,
                            const int32_t *h_goals_xy, int Q, int metric, int32_t *h_cost_i, double *h_cost_f,
                            int32_t *h_path_xy, int32_t *h_path_len, int max_path)
{
    if (!ctx) return FX_ERR_ARG;
    if (!h_grid) return fx_set_err(ctx, FX_ERR_ARG, "fx_plan_host: bad argument");
    return plan_host_impl(ctx, h_grid, nullptr, W, H, h_starts_xy, h_goals_xy, Q, metric, h_cost_i, h_cost_f, h_path_xy, h_path_len, max_path, nullptr);
}

extern "C" int fx_plan_host_f64(fx_context *ctx, const double *h_matrix, int W, int H, const int32_t *h_starts_xy,
                                const int32_t *h_goals_xy, int Q, int metric, int32_t *h_cost_i, double *h_cost_f,
                                int32_t *h_path_xy, int32_t *h_path_len, int max_path)
{
    if (!ctx) return FX_ERR_ARG;
    if (!h_matrix) return fx_set_err(ctx, FX_ERR_ARG, "fx_plan_host_f64: bad argument");
    return plan_host_impl(ctx, nullptr, h_matrix, W, H, h_starts_xy, h_goals_xy, Q, metric, h_cost_i, h_cost_f, h_path_xy, h_path_len, max_path, nullptr);
}

extern "C" int fx_plan_host_csr(fx_context *ctx, const uint8_t *h_grid, int W, int H, const int32_t *h_starts_xy,
                                const int32_t *h_goals_xy, int Q, int metric, int32_t *h_cost_i, double *h_cost_f,
                                int32_t *h_path_len, int max_path, int64_t *h_offsets, int32_t *h_xy, int64_t cap, int64_t *h_total)
{
    if (!ctx) return FX_ERR_ARG;
    if (!h_grid || !h_offsets || cap < 0 || (cap > 0 && !h_xy) || max_path <= 0) return fx_set_err(ctx, FX_ERR_ARG, "fx_plan_host_csr: bad argument");
    CsrOut csr = {h_offsets, h_xy, cap, h_total};
    return plan_host_impl(ctx, h_grid, nullptr, W, H, h_starts_xy, h_goals_xy, Q, metric, h_cost_i, h_cost_f, nullptr, h_path_len, max_path, &csr);
}

static int plan_host_impl(fx_context *ctx, const uint8_t *h_grid, const double *h_matrix, int W, int H, const int32_t *h_starts_xy,
                          const int32_t *h_goals_xy, int Q, int metric, int32_t *h_cost_i, double *h_cost_f,
                          int32_t *h_path_xy, int32_t *h_path_len, int max_path, const CsrOut *csr)
{
    if (csr && Q == 0) { csr->h_offsets[0] = 0; if (csr->h_total) *csr->h_total = 0; }
    if (W <= 0 || H <= 0 || Q < 0 || (Q > 0 && (!h_starts_xy || !h_goals_xy || !h_cost_i)) || max_path < 0)
        return fx_set_err(ctx, FX_ERR_ARG, "fx_plan_host: bad argument");
    if (Q == 0) return FX_OK;
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->own_stream;
    const size_t cells = (size_t)W * H;
    const bool want_path = (h_path_xy || csr) && max_path > 0;
    int rc;
    if ((rc = grow(ctx, &ctx->d_grid, &ctx->d_grid_cap, cells))) return rc;
    if ((rc = grow(ctx, &ctx->d_q, &ctx->d_q_cap, (size_t)Q * 4 * sizeof(int32_t)))) return rc;
    // d_out_i and d_out_f share one capacity counter (both sized by Q)
    if (ctx->d_out_cap < (size_t)Q) {
        if (ctx->d_out_i) cudaFree(ctx->d_out_i);
        if (ctx->d_out_f) cudaFree(ctx->d_out_f);
        ctx->d_out_i = nullptr; ctx->d_out_f = nullptr; ctx->d_out_cap = 0;
        FX_CUDA(ctx, cudaMalloc(&ctx->d_out_i, (size_t)Q * 2 * sizeof(int32_t)));
        FX_CUDA(ctx, cudaMalloc(&ctx->d_out_f, (size_t)Q * sizeof(double)));
        ctx->d_out_cap = (size_t)Q;
    }
    const size_t path_bytes = want_path ? (size_t)Q * max_path * 2 * sizeof(int32_t) : 0;
    if (want_path && (rc = grow(ctx, &ctx->d_path, &ctx->d_path_cap, path_bytes))) return rc;
    // compact form of the paths on the device: offsets[Q+1] + packed points (paths.cu)
    if (want_path && (rc = grow(ctx, &ctx->d_cpath, &ctx->d_cpath_cap, path_bytes))) return rc;
    if (want_path && (rc = grow(ctx, &ctx->d_coff, &ctx->d_coff_cap, ((size_t)Q + 1) * sizeof(int64_t)))) return rc;
    // pinned staging: [grid | starts | goals | cost_i | path_len | cost_f | offsets | paths]
    const size_t qb = (size_t)Q * 2 * sizeof(int32_t);
    size_t off_grid = 0, off_s = (cells + 15) / 16 * 16, off_g = off_s + qb, off_ci = off_g + qb,
           off_pl = off_ci + (size_t)Q * 4, off_cf = (off_pl + (size_t)Q * 4 + 7) / 8 * 8, off_o = off_cf + (size_t)Q * 8,
           off_p = off_o + ((size_t)Q + 1) * 8;
    if ((rc = grow_pinned(ctx, off_p + path_bytes))) return rc;
    char *pin = (char *)ctx->h_pin;
    memcpy(pin + off_s, h_starts_xy, qb);
    memcpy(pin + off_g, h_goals_xy, qb);
    {
        // Maps of the reference's own size, a handful of queries: no copies at all.  The pinned staging buffer is mapped
        // into the device's address space (unified addressing), so the shared-memory kernel reads the grid and the
        // queries from it once and writes its few result bytes straight back: one launch, one synchronisation.
        const char *e = getenv("FUXI_B200_SMALL");
        if (cells <= FX_SMALL_CELLS && Q <= 64 && !(e && e[0] == '0')) {
            if (h_matrix) convert_f64_chunk(h_matrix, (uint8_t *)pin + off_grid, cells, 1);
            else memcpy(pin + off_grid, h_grid, cells);
            rc = fx_search_small(ctx, (const uint8_t *)pin + off_grid, W, H, (const int32_t *)(pin + off_s), (const int32_t *)(pin + off_g), Q,
                                 metric, (int32_t *)(pin + off_ci), (double *)(pin + off_cf), want_path ? (int32_t *)(pin + off_p) : nullptr,
                                 (int32_t *)(pin + off_pl), want_path ? max_path : 0, st);
            if (rc) return rc < 0 ? rc : fx_set_err(ctx, FX_ERR_ARG, "fx_plan_host: small-map path refused the map");
            FX_CUDA(ctx, cudaStreamSynchronize(st));
            memcpy(h_cost_i, pin + off_ci, (size_t)Q * 4);
            if (h_path_len) memcpy(h_path_len, pin + off_pl, (size_t)Q * 4);
            if (h_cost_f) memcpy(h_cost_f, pin + off_cf, (size_t)Q * 8);
            if (want_path) {
                // only the points each query produced (the buffers are [Q][max_path][2])
                const int32_t *pl = (const int32_t *)(pin + off_pl);
                int64_t run = 0;
                for (int q = 0; q < Q; q++) {
                    const int np = (pl[q] <= 0 || pl[q] > max_path) ? 0 : pl[q];
                    if (csr) {
                        csr->h_offsets[q] = run;
                        for (int i = 0; i < np; i++)
                            if (run + i < csr->cap) memcpy(csr->h_xy + (run + i) * 2, pin + off_p + ((size_t)q * max_path + i) * 8, 8);
                        run += np;
                    } else if (np) {
                        memcpy(h_path_xy + (size_t)q * max_path * 2, pin + off_p + (size_t)q * max_path * 8, (size_t)np * 8);
                    }
                }
                if (csr) { csr->h_offsets[Q] = run; if (csr->h_total) *csr->h_total = run; }
            }
            return FX_OK;
        }
    }
    if (h_matrix) {
        unsigned hc = std::thread::hardware_concurrency();
        const int nthreads = (int)(hc == 0 ? 1 : (hc > 16 ? 16 : hc));
        const size_t chunk = cells > (1u << 22) ? (cells + 3) / 4 : cells;
        for (size_t a = 0; a < cells; a += chunk) {
            const size_t nb = a + chunk < cells ? chunk : cells - a;
            convert_f64_chunk(h_matrix + a, (uint8_t *)pin + off_grid + a, nb, nthreads);
            FX_CUDA(ctx, cudaMemcpyAsync(ctx->d_grid + a, pin + off_grid + a, nb, cudaMemcpyHostToDevice, st));
        }
    } else {
        par_memcpy(pin + off_grid, h_grid, cells);
        FX_CUDA(ctx, cudaMemcpyAsync(ctx->d_grid, pin + off_grid, cells, cudaMemcpyHostToDevice, st));
    }
    FX_CUDA(ctx, cudaMemcpyAsync(ctx->d_q, pin + off_s, 2 * qb, cudaMemcpyHostToDevice, st));
    int32_t *d_s = ctx->d_q, *d_g = ctx->d_q + (size_t)Q * 2;
    int32_t *d_ci = ctx->d_out_i, *d_pl = ctx->d_out_i + Q;
    rc = fx_search_batch(ctx, ctx->d_grid, W, H, d_s, d_g, Q, metric, d_ci, ctx->d_out_f, want_path ? ctx->d_path : nullptr,
                         d_pl, want_path ? max_path : 0, (void *)st);
    if (rc) return rc;
    if (want_path) {
        rc = fx_paths_compact(ctx, ctx->d_path, d_pl, Q, max_path, ctx->d_coff, ctx->d_cpath, (int64_t)Q * max_path, (void *)st);
        if (rc) return rc;
        FX_CUDA(ctx, cudaMemcpyAsync(pin + off_o, ctx->d_coff, ((size_t)Q + 1) * 8, cudaMemcpyDeviceToHost, st));
    }
    FX_CUDA(ctx, cudaMemcpyAsync(pin + off_ci, d_ci, (size_t)Q * 8, cudaMemcpyDeviceToHost, st));  // cost_i + path_len
    FX_CUDA(ctx, cudaMemcpyAsync(pin + off_cf, ctx->d_out_f, (size_t)Q * 8, cudaMemcpyDeviceToHost, st));
    FX_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(h_cost_i, pin + off_ci, (size_t)Q * 4);
    if (h_path_len) memcpy(h_path_len, pin + off_pl, (size_t)Q * 4);
    if (h_cost_f) memcpy(h_cost_f, pin + off_cf, (size_t)Q * 8);
    ctx->last_d2h_bytes = (int64_t)Q * 16;
    if (want_path) {
        // only the points the batch produced cross the bus: `total` is known after the first (small) copy
        const int64_t *offs = (const int64_t *)(pin + off_o);
        const int64_t total = offs[Q];
        if (total > 0) {
            FX_CUDA(ctx, cudaMemcpyAsync(pin + off_p, ctx->d_cpath, (size_t)total * 8, cudaMemcpyDeviceToHost, st));
            FX_CUDA(ctx, cudaStreamSynchronize(st));
        }
        ctx->last_d2h_bytes += ((int64_t)Q + 1) * 8 + total * 8;
        if (csr) {
            memcpy(csr->h_offsets, offs, ((size_t)Q + 1) * 8);
            if (csr->h_total) *csr->h_total = total;
            const int64_t ncopy = total < csr->cap ? total : csr->cap;
            if (ncopy > 0) par_memcpy(csr->h_xy, pin + off_p, (size_t)ncopy * 8);
        } else {
            // scatter into the caller's padded rows: a few host threads when the batch is large (the rows of a fresh
            // numpy array are untouched pages: the page faults, not the copies, dominate a single thread)
            auto rows = [=](int a, int b) {
                for (int q = a; q < b; q++) {
                    const int64_t n = offs[q + 1] - offs[q];
                    if (n > 0) memcpy(h_path_xy + (size_t)q * max_path * 2, pin + off_p + (size_t)offs[q] * 8, (size_t)n * 8);
                }
            };
            unsigned hc = std::thread::hardware_concurrency();
            const int nthreads = total < (1 << 18) ? 1 : (int)(hc == 0 ? 1 : (hc > 8 ? 8 : hc));
            if (nthreads <= 1) rows(0, Q);
            else {
                std::vector<std::thread> th;
                const int per = (Q + nthreads - 1) / nthreads;
                for (int t = 1; t < nthreads; t++)
                    if (t * per < Q) th.emplace_back(rows, t * per, (t + 1) * per < Q ? (t + 1) * per : Q);
                rows(0, per < Q ? per : Q);
                for (auto &t : th) t.join();
            }
        }
    }
    return FX_OK;
}

// Host-buffer map call: cloud -> occupancy grid -> inflated grid.
extern "C" int fx_map_host(fx_context *ctx, const float *h_pts, int64_t n, int stride_floats, const float *h_affine3x4,
                           float zmin, float zmax, float ox, float oy, float reso, int W, int H, int radius, int step,
                           uint8_t *h_grid_out)
{
    if (!ctx) return FX_ERR_ARG;
    if (n < 0 || (n > 0 && !h_pts) || !h_grid_out || W <= 0 || H <= 0 || (stride_floats != 3 && stride_floats != 4))
        return fx_set_err(ctx, FX_ERR_ARG, "fx_map_host: bad argument");
    FX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->own_stream;
    const size_t cells = (size_t)W * H, pbytes = (size_t)n * stride_floats * sizeof(float);
    int rc;
    if ((rc = grow(ctx, &ctx->d_grid, &ctx->d_grid_cap, cells))) return rc;
    if ((rc = grow(ctx, &ctx->d_grid2, &ctx->d_grid2_cap, cells))) return rc;
    if ((rc = grow(ctx, &ctx->d_pts, &ctx->d_pts_cap, pbytes + 16))) return rc;
    FX_CUDA(ctx, cudaMemcpyAsync(ctx->d_pts, h_pts, pbytes, cudaMemcpyHostToDevice, st));
    rc = fx_project(ctx, ctx->d_pts, n, stride_floats, h_affine3x4, zmin, zmax, ox, oy, reso, W, H, ctx->d_grid, 1, (void *)st);
    if (rc) return rc;
    const uint8_t *result = ctx->d_grid;
    if (radius > 0) {
        rc = fx_inflate(ctx, ctx->d_grid, ctx->d_grid2, W, H, radius, step, (void *)st);
        if (rc) return rc;
        result = ctx->d_grid2;
    }
    FX_CUDA(ctx, cudaMemcpyAsync(h_grid_out, result, cells, cudaMemcpyDeviceToHost, st));
    FX_CUDA(ctx, cudaStreamSynchronize(st));
    return FX_OK;
}
