// api.cu -- context lifetime, error reporting and the host-buffer entry points of libfuxi_b200.so.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
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
    ctx->cfg_graphs = 1;
    if (const char *e = getenv("FUXI_B200_GRAPHS")) ctx->cfg_graphs = e[0] != '0';
    // [0..15] per-launch counters (zeroed by every launch), [16..23] the 16 x u32 first-bound table of the latency forms
    // of the search (persists across launches, search.cu: fx_first_bound)
    e = cudaMalloc(&ctx->counters, 24 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->counters, 0, 24 * sizeof(unsigned long long));
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
    for (auto &g : ctx->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->ev_search[0]) cudaEventDestroy(ctx->ev_search[0]);
    if (ctx->ev_search[1]) cudaEventDestroy(ctx->ev_search[1]);
    if (ctx->ev_band[0]) cudaEventDestroy(ctx->ev_band[0]);
    if (ctx->ev_band[1]) cudaEventDestroy(ctx->ev_band[1]);
    free(ctx);
    return FX_OK;
}

int fx_graph_run(fx_context *ctx, int slot, const void *key, size_t keylen, cudaStream_t st, const std::function<int(cudaStream_t)> &enqueue)
{
    fx_context::GraphSlot &g = ctx->graphs[slot];
    if (!ctx->cfg_graphs || keylen > sizeof(g.key)) return enqueue(st);
    const bool same = g.keylen == keylen && memcmp(g.key, key, keylen) == 0;
    if (same && g.exec) {
        ctx->launches++;
        FX_CUDA(ctx, cudaGraphLaunch(g.exec, st));
        return FX_OK;
    }
    if (!same) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        memcpy(g.key, key, keylen);
        g.keylen = keylen;
        g.sightings = 1;
        return enqueue(st);
    }
    if (++g.sightings < 2) return enqueue(st);
    // the capture must not be disturbed by (and must not disturb) work the caller has in flight on `st`
    if (!ctx->cap_stream) FX_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
    if (cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); g.keylen = 0; return enqueue(st); }
    const int64_t launches0 = ctx->launches;
    const int rc = enqueue(ctx->cap_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(ctx->cap_stream, &graph);
    ctx->launches = launches0;  // nothing ran yet
    if (rc != FX_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        g.keylen = 0;  // do not try again with this key until it has been seen twice more
        return enqueue(st);
    }
    const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { cudaGetLastError(); g.exec = nullptr; g.keylen = 0; return enqueue(st); }
    ctx->launches++;
    FX_CUDA(ctx, cudaGraphLaunch(g.exec, st));
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
// Host worker pool (process-wide, created on first use): pageable <-> pinned staging copies of tens of MB and the
// float64 -> uint8 conversion run on a few host threads.  Round 1 spawned std::threads per call: 15 spawns per chunk
// x 4 chunks cost more than a millisecond of a 7 ms single-query call.
namespace {
struct HostPool {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void(int)> fn;
    int ntasks = 0;
    std::atomic<int> next{0}, pending{0};
    unsigned long long gen = 0;
    bool stop = false;
    explicit HostPool(int n)
    {
        for (int i = 0; i < n; i++)
            workers.emplace_back([this]() {
                unsigned long long seen = 0;
                for (;;) {
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&]() { return stop || gen != seen; });
                        if (stop) return;
                        seen = gen;
                    }
                    drain();
                }
            });
    }
    ~HostPool()
    {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto &t : workers) t.join();
    }
    void drain()
    {
        for (;;) {
            const int i = next.fetch_add(1, std::memory_order_acq_rel);
            if (i >= ntasks) return;
            fn(i);
            pending.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    // run f(0..n-1) on the pool and the calling thread; `between`, if given, is called by the calling thread after each
    // of its own tasks (used to issue H2D copies of the parts that are complete while the rest is still being filled)
    void run(int n, std::function<void(int)> f, const std::function<void()> &between = nullptr)
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            fn = std::move(f); ntasks = n; next.store(0); pending.store(n); gen++;
        }
        cv.notify_all();
        for (;;) {
            const int i = next.fetch_add(1, std::memory_order_acq_rel);
            if (i >= ntasks) break;
            fn(i);
            pending.fetch_sub(1, std::memory_order_acq_rel);
            if (between) between();
        }
        while (pending.load(std::memory_order_acquire) > 0) {
            if (between) between();
            std::this_thread::yield();
        }
    }
};
HostPool &host_pool()
{
    static std::mutex one;       // one job at a time (contexts on several threads share the pool)
    static HostPool *pool = nullptr;
    std::lock_guard<std::mutex> lk(one);
    if (!pool) {
        unsigned hc = std::thread::hardware_concurrency();
        int n = (int)(hc == 0 ? 1 : (hc > 16 ? 16 : hc)) - 1;
        if (const char *e = getenv("FUXI_B200_HOST_THREADS")) n = atoi(e) - 1;  // tuning experiments only
        pool = new HostPool(n < 0 ? 0 : n);  // never destroyed: worker threads must not be joined from a static destructor
    }
    return *pool;
}
std::mutex g_pool_job;
}  // namespace

// pageable <-> pinned staging copies of tens of MB (the grid in, the path buffers out): a few host threads instead of
// one memcpy (67 MB of path buffers per 8192-query batch took 8 ms single-threaded, 6 % of the end-to-end step)
static void par_memcpy(void *dst, const void *src, size_t bytes)
{
    if (bytes < ((size_t)4 << 20)) { memcpy(dst, src, bytes); return; }
    const size_t per = (size_t)1 << 20;
    const int n = (int)((bytes + per - 1) / per);
    std::lock_guard<std::mutex> job(g_pool_job);
    host_pool().run(n, [=](int i) {
        const size_t a = (size_t)i * per, nb = a + per < bytes ? per : bytes - a;
        memcpy((char *)dst + a, (const char *)src + a, nb);
    });
}

// Filling the pinned staging buffer with NON-TEMPORAL stores: written with ordinary stores, the 16 MB of a 4096^2 grid
// sit dirty in the caches of the worker threads' cores, and the DMA engine's reads have to snoop them out one line at a
// time -- the H2D copy then runs at 7 GB/s instead of 29 (r02j trace: 2.3 ms instead of 0.6 for the same bytes).
static inline void nt_copy_u8(uint8_t *dst, const uint8_t *src, size_t n)
{
#if defined(__x86_64__)
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 15u)) { dst[i] = src[i]; i++; }
    for (; i + 16 <= n; i += 16) _mm_stream_si128((__m128i *)(dst + i), _mm_loadu_si128((const __m128i *)(src + i)));
    for (; i < n; i++) dst[i] = src[i];
    _mm_sfence();
#else
    memcpy(dst, src, n);
#endif
}
static inline void nt_eq1_f64(uint8_t *dst, const double *src, size_t n)  // dst[i] = src[i] == 1.0 (scripts/jps1.py:20-29)
{
#if defined(__x86_64__)
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 15u)) { dst[i] = src[i] == 1.0 ? (uint8_t)1 : (uint8_t)0; i++; }
    const __m128d one = _mm_set1_pd(1.0);
    for (; i + 16 <= n; i += 16) {
        __m128i w[4];
        for (int k = 0; k < 4; k++) {  // 4 doubles -> 4 x 32-bit masks
            const __m128 lo = _mm_castpd_ps(_mm_cmpeq_pd(_mm_loadu_pd(src + i + 4 * k), one));
            const __m128 hi = _mm_castpd_ps(_mm_cmpeq_pd(_mm_loadu_pd(src + i + 4 * k + 2), one));
            w[k] = _mm_castps_si128(_mm_shuffle_ps(lo, hi, _MM_SHUFFLE(2, 0, 2, 0)));
        }
        const __m128i b = _mm_packs_epi16(_mm_packs_epi32(w[0], w[1]), _mm_packs_epi32(w[2], w[3]));  // 0 / -1 per byte
        _mm_stream_si128((__m128i *)(dst + i), _mm_and_si128(b, _mm_set1_epi8(1)));
    }
    for (; i < n; i++) dst[i] = src[i] == 1.0 ? (uint8_t)1 : (uint8_t)0;
    _mm_sfence();
#else
    for (size_t i = 0; i < n; i++) dst[i] = src[i] == 1.0 ? (uint8_t)1 : (uint8_t)0;
#endif
}

// The grid into the pinned staging buffer and on to the device: `fill(a, b)` produces cells [a, b) of the staging buffer
// (a copy of the caller's uint8 grid, or `== 1.0` of its float64 matrix); the grid is cut into tasks, the tasks into four
// upload parts, and the calling thread issues a part's H2D copy as soon as its tasks are done, so the copies overlap the
// filling of the rest.
static int staged_upload(fx_context *ctx, uint8_t *d_dst, uint8_t *pin_grid, size_t cells, const std::function<void(size_t, size_t)> &fill, cudaStream_t st,
                         const std::function<cudaError_t(size_t)> &after_part = nullptr /* called with the bytes uploaded so far */)
{
    if (cells < ((size_t)1 << 20)) {
        fill(0, cells);
        FX_CUDA(ctx, cudaMemcpyAsync(d_dst, pin_grid, cells, cudaMemcpyHostToDevice, st));
        if (after_part) FX_CUDA(ctx, after_part(cells));
        return FX_OK;
    }
    constexpr int PARTS = 4, TPP = 16;  // tasks per part
    const size_t per = ((cells + PARTS * TPP - 1) / (PARTS * TPP) + 63) & ~(size_t)63;
    const int ntasks = (int)((cells + per - 1) / per);
    std::atomic<int> done[PARTS];
    for (auto &d : done) d.store(0);
    int issued = 0;
    cudaError_t err = cudaSuccess;
    auto part_tasks = [&](int p) { const int lo = p * TPP, hi = lo + TPP < ntasks ? lo + TPP : ntasks; return hi > lo ? hi - lo : 0; };
    auto issue = [&]() {
        while (issued < PARTS && done[issued].load(std::memory_order_acquire) >= part_tasks(issued)) {
            const size_t a = (size_t)issued * TPP * per;
            if (a < cells && err == cudaSuccess) {
                const size_t nb = a + TPP * per < cells ? TPP * per : cells - a;
                err = cudaMemcpyAsync(d_dst + a, pin_grid + a, nb, cudaMemcpyHostToDevice, st);
                if (err == cudaSuccess && after_part) err = after_part(a + nb);
            }
            issued++;
        }
    };
    {
        std::lock_guard<std::mutex> job(g_pool_job);
        host_pool().run(ntasks, [&](int i) {
            const size_t a = (size_t)i * per, b = a + per < cells ? a + per : cells;
            fill(a, b);
            done[i / TPP].fetch_add(1, std::memory_order_acq_rel);
        }, issue);
    }
    issue();
    FX_CUDA(ctx, err);
    return FX_OK;
}

// h_src (pageable) -> start of the pinned staging buffer (which must hold `bytes`) -> d_dst, as above
int fx_staged_copy_in(fx_context *ctx, uint8_t *d_dst, const uint8_t *h_src, size_t bytes, cudaStream_t st)
{
    uint8_t *pp = (uint8_t *)ctx->h_pin;
    return staged_upload(ctx, d_dst, pp, bytes, [=](size_t a, size_t b) { nt_copy_u8(pp + a, h_src + a, b - a); }, st);
}

extern "C" int fx_plan_host_stages(fx_context *ctx, double *h_us6)
{
    if (!ctx || !h_us6) return FX_ERR_ARG;
    if (!ctx->last_stage_valid) return fx_set_err(ctx, FX_ERR_ARG, "fx_plan_host_stages: the last fx_plan_host call was not traced (FUXI_B200_TRACE=1|2) or took the shared-memory path");
    memcpy(h_us6, ctx->last_stage_us, 6 * sizeof(double));
    return FX_OK;
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
    // FUXI_B200_TRACE=1: wall-clock of the host stages of this call on stderr (tuning aid)
    const char *trace_env = getenv("FUXI_B200_TRACE");  // read per call: 1 = print, 2 = record only (fx_plan_host_stages)
    const int trace_level = trace_env ? atoi(trace_env) : 0;
    const bool trace = trace_level > 0;
    ctx->last_stage_valid = 0;
    auto now = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = trace ? now() : 0.0;
    double t_up = 0.0, t_enq = 0.0, t_sync = 0.0;
    cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (trace) { for (auto &e : tev) cudaEventCreate(&e); cudaEventRecord(tev[0], st); }
    memcpy(pin + off_s, h_starts_xy, qb);
    memcpy(pin + off_g, h_goals_xy, qb);
    {
        // Maps of the reference's own size, a handful of queries: no copies at all.  The pinned staging buffer is mapped
        // into the device's address space (unified addressing), so the shared-memory kernel reads the grid and the
        // queries from it once and writes its few result bytes straight back: one launch, one synchronisation.
        const char *e = getenv("FUXI_B200_SMALL");
        if (cells <= FX_SMALL_CELLS && Q <= 64 && !(e && e[0] == '0')) {
            if (h_matrix) { uint8_t *pg = (uint8_t *)pin + off_grid; for (size_t i = 0; i < cells; i++) pg[i] = h_matrix[i] == 1.0 ? (uint8_t)1 : (uint8_t)0; }
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
    {
        uint8_t *pg = (uint8_t *)pin + off_grid;
        static const bool nt = !(getenv("FUXI_B200_NT") && getenv("FUXI_B200_NT")[0] == '0');  // tuning experiments only
        // the legal-move mask of the rows that have arrived is built behind each part of the upload (rows x need rows
        // x - 1 .. x + 1), so that only the last part's share of that kernel is left on the critical path
        int rows_done = 0, mrc = FX_OK;
        auto after_part = [&](size_t bytes_up) -> cudaError_t {
            const int avail = (int)(bytes_up / (size_t)H);                 // complete rows on the device
            const int upto = bytes_up >= cells ? W : (avail > 0 ? avail - 1 : 0);
            if (upto > rows_done && mrc == FX_OK) { mrc = fx_build_moves_rows(ctx, ctx->d_grid, W, H, true, rows_done, upto, st); rows_done = upto; }
            return cudaSuccess;
        };
        if (h_matrix && nt) rc = staged_upload(ctx, ctx->d_grid, pg, cells, [=](size_t a, size_t b) { nt_eq1_f64(pg + a, h_matrix + a, b - a); }, st, after_part);
        else if (h_matrix) rc = staged_upload(ctx, ctx->d_grid, pg, cells, [=](size_t a, size_t b) { for (size_t i = a; i < b; i++) pg[i] = h_matrix[i] == 1.0 ? (uint8_t)1 : (uint8_t)0; }, st, after_part);
        else if (nt) rc = staged_upload(ctx, ctx->d_grid, pg, cells, [=](size_t a, size_t b) { nt_copy_u8(pg + a, h_grid + a, b - a); }, st, after_part);
        else rc = staged_upload(ctx, ctx->d_grid, pg, cells, [=](size_t a, size_t b) { memcpy(pg + a, h_grid + a, b - a); }, st, after_part);
        if (rc) return rc;
        if (mrc) return mrc;
        ctx->moves_prebuilt_for = ctx->d_grid;
    }
    if (trace) { t_up = now(); cudaEventRecord(tev[1], st); }
    FX_CUDA(ctx, cudaMemcpyAsync(ctx->d_q, pin + off_s, 2 * qb, cudaMemcpyHostToDevice, st));
    int32_t *d_s = ctx->d_q, *d_g = ctx->d_q + (size_t)Q * 2;
    int32_t *d_ci = ctx->d_out_i, *d_pl = ctx->d_out_i + Q;
    rc = fx_search_batch(ctx, ctx->d_grid, W, H, d_s, d_g, Q, metric, d_ci, ctx->d_out_f, want_path ? ctx->d_path : nullptr,
                         d_pl, want_path ? max_path : 0, (void *)st);
    ctx->moves_prebuilt_for = nullptr;  // (a batch that took the shared-memory kernel did not consume it)
    if (rc) return rc;
    if (trace) cudaEventRecord(tev[2], st);
    if (want_path) {
        rc = fx_paths_compact(ctx, ctx->d_path, d_pl, Q, max_path, ctx->d_coff, ctx->d_cpath, (int64_t)Q * max_path, (void *)st);
        if (rc) return rc;
        FX_CUDA(ctx, cudaMemcpyAsync(pin + off_o, ctx->d_coff, ((size_t)Q + 1) * 8, cudaMemcpyDeviceToHost, st));
    }
    FX_CUDA(ctx, cudaMemcpyAsync(pin + off_ci, d_ci, (size_t)Q * 8, cudaMemcpyDeviceToHost, st));  // cost_i + path_len
    FX_CUDA(ctx, cudaMemcpyAsync(pin + off_cf, ctx->d_out_f, (size_t)Q * 8, cudaMemcpyDeviceToHost, st));
    if (trace) { t_enq = now(); cudaEventRecord(tev[3], st); }
    FX_CUDA(ctx, cudaStreamSynchronize(st));
    if (trace) {
        t_sync = now();
        float kms = 0.f;
        if (ctx->ev_search_valid) cudaEventElapsedTime(&kms, ctx->ev_search[0], ctx->ev_search[1]);
        float d01 = 0.f, d12 = 0.f, d23 = 0.f;
        cudaEventElapsedTime(&d01, tev[0], tev[1]); cudaEventElapsedTime(&d12, tev[1], tev[2]); cudaEventElapsedTime(&d23, tev[2], tev[3]);
        for (auto &e : tev) cudaEventDestroy(e);
        const double st6[6] = {t_up - t_begin, t_enq - t_up, t_sync - t_enq, 1e3 * d01, 1e3 * d12, 1e3 * d23};
        memcpy(ctx->last_stage_us, st6, sizeof(st6));
        ctx->last_stage_valid = 1;
        if (trace_level == 1) fprintf(stderr, "fx_plan_host trace: host: fill+upload issue %.0f us, enqueue %.0f us, wait %.0f us | device: upload %.0f us, fx_search_batch %.0f us (search kernel %.0f us), paths+D2H %.0f us\n",
                t_up - t_begin, t_enq - t_up, t_sync - t_enq, 1e3 * d01, 1e3 * d12, 1e3 * kms, 1e3 * d23);
    }
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
            if (total < (1 << 18)) rows(0, Q);
            else {
                const int per = 64, n = (Q + per - 1) / per;
                std::lock_guard<std::mutex> job(g_pool_job);
                host_pool().run(n, [=](int i) { rows(i * per, (i + 1) * per < Q ? (i + 1) * per : Q); });
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
    // the cloud goes through the pinned staging buffer (non-temporal host copy on the worker pool, H2D of the finished
    // parts overlapping the rest): a cudaMemcpyAsync from the caller's pageable array ran at a third of the PCIe rate
    if ((rc = grow_pinned(ctx, pbytes > cells ? pbytes : cells))) return rc;
    {
        uint8_t *pp = (uint8_t *)ctx->h_pin;
        const uint8_t *src = (const uint8_t *)h_pts;
        if (pbytes && (rc = staged_upload(ctx, (uint8_t *)ctx->d_pts, pp, pbytes, [=](size_t a, size_t b) { nt_copy_u8(pp + a, src + a, b - a); }, st))) return rc;
    }
    rc = fx_project(ctx, ctx->d_pts, n, stride_floats, h_affine3x4, zmin, zmax, ox, oy, reso, W, H, ctx->d_grid, 1, (void *)st);
    if (rc) return rc;
    const uint8_t *result = ctx->d_grid;
    if (radius > 0) {
        rc = fx_inflate(ctx, ctx->d_grid, ctx->d_grid2, W, H, radius, step, (void *)st);
        if (rc) return rc;
        result = ctx->d_grid2;
    }
    FX_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, result, cells, cudaMemcpyDeviceToHost, st));  // (stream order: after the H2D copies out of the same buffer)
    FX_CUDA(ctx, cudaStreamSynchronize(st));
    par_memcpy(h_grid_out, ctx->h_pin, cells);
    return FX_OK;
}
