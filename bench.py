#!/usr/bin/env python
"""bench.py -- headline benchmark of the FUXI global-planning hot path on B200.

Metric (BASELINE.json): planning queries/s on a 4096^2 random-obstacle grid (20 % fill), batched
start/goal queries sharded query-parallel over N GPUs (cfg4 of SURVEY.md §8d: 65 536 queries at N = 8,
i.e. 8 192 queries per GPU, "weak" scaling), plus p50 single-replan latency of the drop-in
``jps1.method``.  One "step" = one fx_search_batch launch over this rank's batch of queries
(move-mask build + search + path extraction), Euclidean metric (hchoice 2, what the planners use).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path (cfg4, the headline)
    python bench.py --impl reference ...                                # CPU arm (see below)
    torchrun --nproc-per-node N bench.py --gpus N ...                   # N > 1
    python bench.py --config cfg2|cfg3|cfg5 ...                         # BASELINE.json's other configurations, same JSON contract

`value`   queries/s with grid and queries resident in HBM, CUDA-event timed, max over ranks.
`e2e`     the same queries through the host-buffer C-ABI call (fx_plan_host): H2D of the grid and the
          queries and D2H of costs + paths inside the timed region.
`roofline`  of the dominant kernel (k_search_batch); `kernels` = the map-side kernels (projection,
          inflation, EDT) at the BASELINE sizes and at HBM-exercising scaled sizes.
`cpu_baseline` / `--impl reference`: the reference itself is Python (scripts/jps1.py) and
          /root/reference does not exist on the GPU box, so the CPU arm is the C restatement of jps1.py in
          oracle/ (kind "port"), OpenMP over ALL host cores (the thread count is passed explicitly: torchrun exports
          OMP_NUM_THREADS=1), 16 queries per thread per step, rank 0 only.
`per_rank`  ms_per_step / k_search_batch ms / k_band_bound ms / median SM clock of every rank (N > 1: names the straggler).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC_NAME = "planning queries/s, 4096^2 grid (20% fill), batched start/goal queries, Euclidean metric"
L2_FLUSH_BYTES = 512 << 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=4096, help="grid edge in cells")
    ap.add_argument("--queries", type=int, default=8192, help="queries per GPU per step")
    ap.add_argument("--hchoice", type=int, default=2, choices=[1, 2])
    ap.add_argument("--max-path", type=int, default=1024)
    ap.add_argument("--no-extras", action="store_true", help="skip map-kernel rooflines, latency and cpu_baseline")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work for the cpu_baseline sample")
    ap.add_argument("--config", default="cfg4", choices=["cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json configuration (cfg4 = the headline metric; the others print the same JSON contract)")
    return ap.parse_args()


def host_threads():
    """Cores this process may use.  OMP_NUM_THREADS is ignored on purpose (torchrun sets it to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------ workload
def make_workload(n, q_total):
    """cfg4 of SURVEY.md §8d: grid default_rng(4) at 20 % fill, queries default_rng(5) over free cells."""
    m = (np.random.default_rng(4).random((n, n)) < 0.2).astype(np.uint8)
    free = np.argwhere(m == 0)
    rng = np.random.default_rng(5)
    s = free[rng.integers(len(free), size=q_total)].astype(np.int32)
    g = free[rng.integers(len(free), size=q_total)].astype(np.int32)
    return m, s, g


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, n, Q, hchoice):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture of the same
    workload (profiles/ncu_traffic.json, written by scripts/ncu_traffic.py from the .ncu-rep), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            for rec in json.load(fh):
                if rec["kernel"] == kernel and rec["grid"] == n and rec["queries"] == Q and rec["hchoice"] == hchoice:
                    return float(rec["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------ reference arm
def cpu_sample(m, s, g, hchoice, target_s, oracle):
    """Times the C restatement of jps1.py over all host threads on the first S queries; S is doubled until
    the sample costs about target_s seconds of wall clock (bounded: S <= 64 x threads)."""
    threads = host_threads()
    S = min(len(s), 16 * threads)
    while True:
        t0 = time.perf_counter()
        cost, status, used = oracle.jps_batch(m, s[:S], g[:S], hchoice, threads=threads)
        dt = time.perf_counter() - t0
        if dt >= target_s / 3 or S >= len(s) or S >= 64 * threads:
            return S, dt, used, cost, status
        S = min(len(s), S * 2)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    oracle.build()
    if args.config != "cfg4":
        return run_reference_other(args, oracle)
    n, Q = args.grid, args.queries
    m, s, g = make_workload(n, Q * args.gpus)
    threads = host_threads()
    # one step = a bounded sample of the workload: the first 16 queries per host thread (~0.17 s per query per core at
    # 4096^2, dynamic schedule), all cores of the box whatever OMP_NUM_THREADS says
    S = min(len(s), 16 * threads)
    for _ in range(min(args.warmup, 1)):
        oracle.jps_batch(m, s[:S], g[:S], args.hchoice, threads=threads)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        _, _, used = oracle.jps_batch(m, s[:S], g[:S], args.hchoice, threads=threads)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = S * args.steps / total
    cfg = workload_config(args, args.queries)
    cfg["cpu_sample_queries_per_step"] = S
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": int(used), "kind": "port",
                         "sample": "first %d of the workload's queries per step (16 per thread; 1 warm-up step), C restatement of "
                                   "scripts/jps1.py (oracle/fuxi_oracle.c), OpenMP over %d host threads" % (S, used)},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, q_per_step):
    return {"workload": "cfg4 shard: %d start/goal queries per GPU per step on a %dx%d random-obstacle grid "
                        "(20%% fill, default_rng(4)/(5)), hchoice %d" % (q_per_step, args.grid, args.grid, args.hchoice),
            "grid": [args.grid, args.grid], "queries_per_gpu": q_per_step, "global_queries": q_per_step * args.gpus,
            "parallelism": "query-parallel x%d (grid replicated, no data-path collective)" % args.gpus,
            "l2": "flushed between timed steps (%d MiB write)" % (L2_FLUSH_BYTES >> 20)}


# ------------------------------------------------------------------------------------------ map kernels
def time_kernel(torch, fn, flush, reps=5, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


def map_kernel_rooflines(torch, fx, dev, flush, peak):
    """Projection / inflation / EDT at the BASELINE sizes (cfg2: 1 M points, 1024^2) and at scaled sizes where
    HBM is exercised (64 M points, 16384^2).  Algorithmic bytes: projection N*16 + W*H (float4
    points) or N*12 + W*H (packed xyz); inflation 2*W*H; EDT 5*W*H."""
    out = {}
    rng = np.random.default_rng(1)

    def entry(name, alg_bytes, fn):
        mean_ms, min_ms = time_kernel(torch, fn, flush)
        gbs = alg_bytes / (mean_ms * 1e-3) / 1e9
        out[name] = {"ms": mean_ms, "ms_min": min_ms, "alg_bytes": int(alg_bytes), "achieved_gbs": gbs, "frac": gbs / peak}

    # 64 Mi points into cfg4's 4096^2 grid: the grid (16 MiB) stays in L2, so this isolates the streaming side of the
    # projection kernel (byte-store form); the 16384^2 case below adds the L2-resident bit-packed scatter (RED.OR).
    N, n = 64 << 20, 4096
    half = n * 0.1
    pts = torch.empty((N, 4), dtype=torch.float32, device=dev)
    pts[:, 0:2].uniform_(-half, half)
    pts[:, 2].uniform_(-0.5, 3.0)
    pts[:, 3] = 0
    grid = torch.empty((n, n), dtype=torch.uint8, device=dev)
    entry("project_f4_64Mpts_4096", N * 16 + n * n, lambda: fx.project(pts, None, 0.3, float("inf"), (-half, -half), 0.2, out=grid))
    del pts, grid
    torch.cuda.empty_cache()
    for label, N, n in (("cfg2", 1 << 20, 1024), ("scaled", 64 << 20, 16384)):
        half = n * 0.1
        pts = torch.empty((N, 4), dtype=torch.float32, device=dev)
        pts[:, 0:2].uniform_(-half, half)
        pts[:, 2].uniform_(-0.5, 3.0)
        pts[:, 3] = 0
        grid = torch.empty((n, n), dtype=torch.uint8, device=dev)
        entry("project_f4_%s" % label, N * 16 + n * n,
              lambda: fx.project(pts, None, 0.3, float("inf"), (-half, -half), 0.2, out=grid))
        p3 = pts[:, :3].contiguous()
        entry("project_xyz_%s" % label, N * 12 + n * n,
              lambda: fx.project(p3, None, 0.3, float("inf"), (-half, -half), 0.2, out=grid))
        del pts, p3
        occ = (torch.rand((n, n), device=dev) < 0.02).to(torch.uint8)
        o2 = torch.empty_like(occ)
        entry("inflate_r2_%s" % label, 2 * n * n, lambda: fx.inflate(occ, 2, "ccst", out=o2))
        entry("inflate_r1_%s" % label, 2 * n * n, lambda: fx.inflate(occ, 1, "st", out=o2))
        d2 = torch.empty((n, n), dtype=torch.int32, device=dev)
        entry("edt_%s" % label, 5 * n * n, lambda: fx.edt(occ, out=d2))
        del occ, o2, d2, grid
        torch.cuda.empty_cache()
    # upstream cloud conditioning: one depth frame (640 x 480 PointXYZRGB, point_step 32) and a scaled
    # cloud; algorithmic bytes = the points read once + the kept points written (the kernels read the cloud three times:
    # bounding box, voxel marking, accumulation -- L2-resident at frame size)
    for label, N in (("frame_640x480", 640 * 480), ("16Mpts", 16 << 20)):
        c = torch.zeros((N, 8), dtype=torch.float32, device=dev)
        c[:, 0].uniform_(-3.0, 3.0)
        c[:, 1].uniform_(-2.0, 2.0)
        c[:, 2].uniform_(-0.5, 5.0)
        c[: N // 2, 2] = 3.0 + 0.03 * torch.randn(N // 2, device=dev)      # a wall: something survives the radius filter
        # the node filters every frame into the same buffers (the launch sequence is then replayed as one CUDA graph)
        state = {"out": torch.empty((N, 4), dtype=torch.float32, device=dev), "counts": torch.empty(4, dtype=torch.int64, device=dev)}

        def run_cf():
            fx.cloud.cloud_filter(c, rgb_offset=4, out=state["out"], counts=state["counts"])
        run_cf()
        kept = int(state["counts"][2].item())
        entry("cloud_filter_%s" % label, N * 32 + kept * 16, run_cf)
        out["cloud_filter_%s" % label]["counts"] = [int(v) for v in state["counts"].tolist()]
        del c
    d = torch.empty((1 << 20, 3), dtype=torch.float64, device=dev).uniform_(-6.0, 6.0)
    mkept = int(fx.cloud.distance_filter(d, 4.0)[1].item())
    entry("distance_filter_1Mpts", (1 << 20) * 24 + mkept * 24, lambda: fx.cloud.distance_filter(d, 4.0))
    del rng, d
    return out


def other_configs(torch, fx, dev, flush, hchoice):
    """BASELINE.json's other single-GPU configurations as reported numbers (not the bench line): cfg3 = 4096 queries on
    a 1024^2 grid at 20 % fill (grid default_rng(2), queries default_rng(3)), both metrics, device-resident;
    plus the host-side baselines SURVEY §8d asks for next to the map kernels (numpy restatements of the reference's
    inline inflation blocks and of distance_filter, timed on this box's host)."""
    import oracle
    out = {}
    n, Q = 1024, 4096
    m = (np.random.default_rng(2).random((n, n)) < 0.2).astype(np.uint8)
    free = np.argwhere(m == 0)
    rng = np.random.default_rng(3)
    s = free[rng.integers(len(free), size=Q)].astype(np.int32)
    g = free[rng.integers(len(free), size=Q)].astype(np.int32)
    d_m, d_s, d_g = torch.from_numpy(m).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(g).to(dev)
    for metric in (1, 2):
        ms, _ = time_kernel(torch, lambda: fx.plan_batch(d_m, d_s, d_g, metric=metric, max_path=1024), flush)
        settled = fx.search_stats()[0]
        res = fx.plan_batch(d_m, d_s, d_g, metric=metric, max_path=1024)
        S = 256
        want, status, _ = oracle.jps_batch(m, s[:S], g[:S], metric)
        ci, cf = res.cost_i[:S].cpu().numpy(), res.cost_f[:S].cpu().numpy()
        ok = status == 1
        bad = (ci[ok] != want[ok]) if metric == 1 else (np.abs(cf[ok] - want[ok]) > 1e-5 * np.maximum(want[ok], 1e-12))
        out["cfg3_1024_4096q_metric%d" % metric] = {"ms": ms, "queries_per_s": Q / (ms * 1e-3), "settled_cells": int(settled),
                                                    "nodes_per_s": settled / (ms * 1e-3),
                                                    "parity_mismatches_first_256": int(bad.sum()) + int(((ci >= 0) != ok).sum())}
    try:
        out["host_numpy_baselines_ms"] = host_baselines()
    except Exception as exc:          # a reported baseline must never take the bench line down
        out["host_numpy_baselines_ms"] = {"error": repr(exc)}
    return out


def host_baselines():
    """The host-side baselines SURVEY §8d lists next to the map kernels, timed on this box's host (one core, no GPU):
    numpy restatements of the reference's inline inflation blocks (a10, a11) and of distance_filter (a17), and the
    cloud node's DBSCAN step (a19, scikit-learn as in the reference, plc_point2_st.py:299-300: eps 0.4, min_samples 6)
    on two synthetic frames of 5 000 points.  Reported as baselines only."""
    import oracle
    occ = (np.random.default_rng(1).random((1024, 1024)) < 0.02).astype(np.float64)
    occ[:4] = 0; occ[-4:] = 0; occ[:, :4] = 0; occ[:, -4:] = 0      # the reference pads before it inflates (st:230-250)
    t0 = time.perf_counter(); oracle.hostref.inflate_ccst(occ.copy(), 2); t1 = time.perf_counter()
    oracle.hostref.inflate_st(occ.copy(), 1); t2 = time.perf_counter()
    pts = np.random.default_rng(1).uniform(-6, 6, (1 << 20, 3))
    t3 = time.perf_counter(); oracle.hostref.distance_filter(pts, 4.0); t4 = time.perf_counter()
    res = {"inflate_ccst_r2_1024": 1e3 * (t1 - t0), "inflate_st_r1_1024": 1e3 * (t2 - t1),
           "distance_filter_1Mpts": 1e3 * (t4 - t3), "cores": 1}
    try:
        from sklearn.cluster import DBSCAN
        rng = np.random.default_rng(2)
        frames = [np.concatenate([c + 0.15 * rng.standard_normal((500, 3)) for c in rng.uniform(-4, 4, (10, 3))]) for _ in range(2)]
        t5 = time.perf_counter()
        labels = [DBSCAN(eps=0.4, min_samples=6).fit(f).labels_ for f in frames]
        res["dbscan_2x5000pts"] = 1e3 * (time.perf_counter() - t5)
        res["dbscan_clusters"] = [int(l.max()) + 1 for l in labels]
    except ImportError:
        res["dbscan_2x5000pts"] = None
    return res


def latency_probe(fx, m4096, s, g, hchoice):
    """p50 of one drop-in replan: jps1.method(matrix, start, goal, 2) wall clock, host buffers in and out."""
    import contextlib
    import io
    res = {}
    z = np.load(os.path.join(ROOT, "tests", "golden", "maps.npz"))
    m1 = z["-16.40-4.80_out.png"].astype(np.float64)
    free = np.argwhere(m1 == 0)
    rng = np.random.default_rng(0)
    pairs = [(tuple(free[rng.integers(len(free))]), tuple(free[rng.integers(len(free))])) for _ in range(1050)]   # 1000 reps after 50 warm-ups
    sink = io.StringIO()

    def run(mat, pairs, warm):
        ts = []
        with contextlib.redirect_stdout(sink):
            for i, (a, b) in enumerate(pairs):
                t0 = time.perf_counter()
                fx.jps1.method(mat, a, b, hchoice)
                dt = time.perf_counter() - t0
                if i >= warm:
                    ts.append(dt)
        ts = np.array(ts) * 1e3
        return {"p50_ms": float(np.percentile(ts, 50)), "p90_ms": float(np.percentile(ts, 90)),
                "p99_ms": float(np.percentile(ts, 99)), "n": len(ts)}

    res["cfg1_map_148x52"] = run(m1, pairs, 50)
    # the same 1000 queries through the C restatement of jps1.py on ONE host core of this box (the reference's own
    # Python is ~2 orders of magnitude slower: BASELINE.md §2 measured p50 10.5 ms on another machine)
    try:
        import oracle
        occ1 = (m1 == 1).astype(np.uint8)
        ts = []
        for i, (a, b) in enumerate(pairs):
            t0 = time.perf_counter()
            oracle.capi.jps(occ1, a, b, hchoice, max_path=4096)
            if i >= 50:
                ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        res["cfg1_map_148x52"]["cpu_port_p50_ms"] = float(np.percentile(ts, 50))
        res["cfg1_map_148x52"]["cpu_port_p99_ms"] = float(np.percentile(ts, 99))
        res["cfg1_map_148x52"]["cpu_port"] = "oracle/fuxi_oracle.c fxo_jps, 1 core, same box, same 1000 queries"
    except Exception as exc:
        res["cfg1_map_148x52"]["cpu_port_error"] = repr(exc)
    m4 = m4096.astype(np.float64)
    pairs4 = [(tuple(int(v) for v in s[i]), tuple(int(v) for v in g[i])) for i in range(450)]
    res["grid_%dx%d" % m4096.shape] = run(m4, pairs4, 50)
    # the same through the uint8 host entry point (no float64 -> uint8 conversion of 16 Mi cells on the host)
    tsu = []
    for i, (a, b) in enumerate(pairs4):
        t0 = time.perf_counter()
        fx.plan_host(m4096, np.array([a], dtype=np.int32), np.array([b], dtype=np.int32), metric=hchoice, max_path=2048)
        if i >= 50:
            tsu.append(time.perf_counter() - t0)
    tsu = np.array(tsu) * 1e3
    res["grid_%dx%d_uint8_host" % m4096.shape] = {"p50_ms": float(np.percentile(tsu, 50)), "p90_ms": float(np.percentile(tsu, 90)),
                                                  "p99_ms": float(np.percentile(tsu, 99)), "n": len(tsu)}
    # where such a call spends its time: medians over 60 traced calls, host wall clock and CUDA events
    try:
        names = ("host_fill_and_upload_issue", "host_enqueue", "host_wait", "device_grid_upload", "device_search_incl_move_mask", "device_paths_and_d2h")
        for label, mat in (("uint8", m4096), ("float64", m4)):
            os.environ["FUXI_B200_TRACE"] = "2"
            rows = []
            for (a, b) in pairs4[20:80]:
                fx.plan_host(mat, np.array([a], dtype=np.int32), np.array([b], dtype=np.int32), metric=hchoice, max_path=2048)
                rows.append(fx.plan_host_stages())
            os.environ.pop("FUXI_B200_TRACE", None)
            med = np.median(np.array(rows), axis=0)
            res["grid_%dx%d_%s_host_stages_us_median" % (m4096.shape + (label,))] = {k: float(v) for k, v in zip(names, med)}
    except Exception as exc:
        os.environ.pop("FUXI_B200_TRACE", None)
        res["host_stages_error"] = repr(exc)
    # the whole planner iteration in one call (decode + pad/shift + inflate + goal relocation + search + shortcutting +
    # world coordinates, fx_replan_host): OccupancyGrid message of the cfg1 map in, world path out
    from fuxi_planner_b200 import planner
    msg = planner.array_to_occupancy_grid(z["-16.40-4.80_out.png"])
    Wm, Hm = m1.shape
    ts = []
    for i, (a, b) in enumerate(pairs[:220]):
        st_xy = (-16.4 + 0.2 * (a[0] + 0.5), -4.8 + 0.2 * (a[1] + 0.5))
        go_xy = (-16.4 + 0.2 * (b[0] + 0.5), -4.8 + 0.2 * (b[1] + 0.5))
        t0 = time.perf_counter()
        planner.replan_fused(msg, Wm, Hm, (-16.4, -4.8), 0.2, st_xy, go_xy, ifa=1, variant="st", hchoice=hchoice)
        dt = time.perf_counter() - t0
        if i >= 20:
            ts.append(dt)
    ts = np.array(ts) * 1e3
    res["replan_fused_cfg1"] = {"p50_ms": float(np.percentile(ts, 50)), "p90_ms": float(np.percentile(ts, 90)),
                                "p99_ms": float(np.percentile(ts, 99)), "n": len(ts)}
    return res


def multi_gpu_checks(torch, dist, fx, dev, world, rank, m):
    """N > 1 only: the sharded modes of SURVEY §8e beside the query-parallel headline, each compared BIT FOR BIT with the
    single-GPU kernel on the cfg4 grid (every rank checks its own slab, all-reduced MIN), warm wall-clock times, max over
    ranks: row-tiled inflation (one halo exchange), row-tiled EDT (two transposes), point-sharded projection (one
    all-reduce), row-tiled cost field (iterated halo exchange).  These are capacity modes, not the metric."""
    from fuxi_planner_b200 import tiled
    n = m.shape[0]
    x0, x1 = tiled.slab_bounds(n, world, rank)
    gm = torch.from_numpy(m).to(dev)
    own = gm[x0:x1].contiguous()
    out = {}

    def timed(fn):
        fn()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return r, float(t.item()) * 1e3

    def same(a, b):
        ok = torch.tensor([int(torch.equal(a, b))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return bool(ok.item())

    for radius, variant in ((2, "ccst"), (3, "st")):
        r, ms = timed(lambda: tiled.inflate_tiled(own, radius, variant))
        out["inflate_tiled_r%d_%s" % (radius, variant)] = {"ms": ms, "bit_exact_vs_single_gpu": same(r, fx.inflate(gm, radius, variant)[x0:x1])}
    r, ms = timed(lambda: tiled.edt_tiled(own, n))
    out["edt_tiled"] = {"ms": ms, "bit_exact_vs_single_gpu": same(r, fx.edt(gm)[x0:x1])}
    npts = 1 << 22
    prng = np.random.default_rng(9)
    pts = torch.from_numpy(np.c_[prng.uniform(0, n * 0.2, (npts, 2)), prng.uniform(-0.5, 3.0, npts)].astype(np.float32)).to(dev)
    p0, p1 = tiled.shard_queries(npts, world, rank)
    mine = pts[p0:p1].contiguous()
    r, ms = timed(lambda: tiled.project_sharded(mine, None, 0.3, float("inf"), (0.0, 0.0), 0.2, (n, n)))
    out["project_sharded_4Mpts"] = {"ms": ms, "bit_exact_vs_single_gpu": same(r, fx.project(pts, None, 0.3, float("inf"), (0.0, 0.0), 0.2, (n, n)))}
    free = np.argwhere(m == 0)
    src = tuple(int(v) for v in free[0])
    (fld, rounds), ms = timed(lambda: tiled.field_tiled(own, n, src, 2))
    t0 = time.perf_counter()
    full = fx.field(gm, src, 2)
    torch.cuda.synchronize()
    out["field_tiled"] = {"ms": ms, "exchange_rounds": int(rounds), "bit_exact_vs_single_gpu": same(fld, full[x0:x1]),
                          "single_gpu_ms_rank0": 1e3 * (time.perf_counter() - t0)}
    out["grid"] = [n, n]
    out["world"] = world
    return out


# ------------------------------------------------------------------------------------------ main arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import fuxi_planner_b200 as fx
    ctx = fx.default_context(local)
    if os.environ.get("FUXI_SLOTS"):  # tuning experiments only: concurrent search slots
        ctx.check(ctx.lib.fx_set_search_tuning(ctx.handle, int(os.environ["FUXI_SLOTS"]), 0), "fx_set_search_tuning")

    if args.config != "cfg4":
        rc = run_other_config(args, torch, dist, fx, ctx, dev, world, rank, local)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return rc
    n, Q = args.grid, args.queries
    m, s_all, g_all = make_workload(n, Q * world)
    s, g = s_all[rank * Q:(rank + 1) * Q], g_all[rank * Q:(rank + 1) * Q]
    d_m = torch.from_numpy(m).to(dev)
    d_s, d_g = torch.from_numpy(s).to(dev), torch.from_numpy(g).to(dev)
    flush_buf = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def flush():
        flush_buf.fill_(1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return fx.plan_batch(d_m, d_s, d_g, metric=args.hchoice, max_path=args.max_path)

    for _ in range(args.warmup):
        res = step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()                      # every rank samples its own GPU (per_rank names a slow one)
    launches0 = ctx.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    kernel_ms, band_ms = [], []
    for a, b in ev:
        flush()
        a.record()
        res = step()
        b.record()
        bm, km = fx.search_timings(ctx)                 # waits for this step's k_search_batch; the next step starts after it
        kernel_ms.append(km)
        band_ms.append(bm)
    barrier()
    clocks_dev = sampler.stop()                         # clocks under the device-timed region only
    launches = ctx.launches - launches0
    t_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    settled, levels, passes, band_only = fx.search_stats(ctx)
    tt = torch.tensor([t_ms], dtype=torch.float64, device=dev)
    st = torch.tensor([float(settled)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(st, op=dist.ReduceOp.SUM)
    t_ms_max = float(tt.item())
    value = Q * world * args.steps / (t_ms_max * 1e-3)
    mine = torch.tensor([t_ms / args.steps, float(np.mean(kernel_ms)), float(np.mean(band_ms)), float(clocks_dev.get("sm_mhz") or 0.0),
                         float(settled)], dtype=torch.float64, device=dev)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    per_rank = [{"rank": r, "ms_per_step": float(v[0]), "search_kernel_ms": float(v[1]), "band_kernel_ms": float(v[2]),
                 "sm_mhz": float(v[3]), "settled_cells": int(v[4])} for r, v in enumerate(t.cpu() for t in allr)]
    answered = int((res.cost_i >= 0).sum().item())

    # ---- e2e: host buffers through the batched host call (fx_plan_host_csr: H2D grid + queries, D2H costs + the paths in
    # compact form), wall clock, max over ranks; the padded-row form (fx_plan_host) is timed beside it
    e2e_steps = args.steps
    cap = 0
    for _ in range(2):                                     # warm (allocates staging), and learns the batch's point count
        ci, cf, pl, offs, xy = fx.plan_host_csr(m, s, g, metric=args.hchoice, max_path=args.max_path, cap=cap or None, ctx=ctx)
        cap = int(offs[-1]) + 1024
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ci, cf, pl, offs, xy = fx.plan_host_csr(m, s, g, metric=args.hchoice, max_path=args.max_path, cap=cap, ctx=ctx)
    torch.cuda.synchronize()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    clocks = clocks_dev if rank == 0 else None
    e2e_value = Q * world * e2e_steps / float(e2e_t.item())
    assert np.array_equal(ci, res.cost_i.cpu().numpy()), "host-buffer path and device path disagree"
    h2d = m.nbytes + s.nbytes + g.nbytes
    d2h = int(ctx.lib.fx_last_d2h_bytes(ctx.handle))    # costs + lengths + offsets + the points of the paths found (compact form)
    # every path the host call returned starts at its start and ends at its goal (equal-cost paths may differ between two
    # runs: ties are decided by which relaxation lands first), and has the device run's length class
    pl_dev = res.path_len.cpu().numpy()
    assert np.array_equal(pl >= 0, pl_dev >= 0)
    for q in range(0, Q, max(1, Q // 64)):
        k = int(offs[q + 1] - offs[q])
        if k > 0:
            assert tuple(xy[offs[q]]) == tuple(s[q]) and tuple(xy[offs[q + 1] - 1]) == tuple(g[q]), "host path is not start..goal"
    fx.plan_host(m, s, g, metric=args.hchoice, max_path=args.max_path, ctx=ctx)
    t0 = time.perf_counter()
    for _ in range(max(1, e2e_steps // 2)):
        fx.plan_host(m, s, g, metric=args.hchoice, max_path=args.max_path, ctx=ctx)
    e2e_padded = Q * max(1, e2e_steps // 2) / (time.perf_counter() - t0)

    peak, peak_src = measured_peak()
    settled_all = float(st.item())
    ms_step = t_ms_max / args.steps
    alg_bytes = settled / 1.0 * 5.0            # this rank's last step: settled cells x (1 B occupancy + 4 B cost)
    k_ms = float(np.mean(kernel_ms))           # k_search_batch alone: CUDA events on its own stream, inside the C library
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC_NAME, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 (cost in 1/2378 cell, packed with the arrival direction)" if args.hchoice == 2 else "u32",
        "data": "synthetic", "config": workload_config(args, Q),
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "api": "fx_plan_host_csr (fuxi_planner_b200.plan_host_csr), host numpy buffers in, costs + offsets[Q+1] + xy[total][2] out",
                "padded_rows_form_rank0": e2e_padded},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "k_search_batch", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic("k_search_batch", n, Q, args.hchoice),
                     "peak_source": peak_src, "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_step,
                     "algorithmic_bytes": "settled cells x 5 B (SURVEY §8d early-exit form)",
                     "full_field_form_gbs": Q * n * n * 5.0 / (k_ms * 1e-3) / 1e9},
        "search": {"settled_cells_per_step_rank0": int(settled), "nodes_per_s": settled_all / (ms_step * 1e-3),
                   "levels": int(levels), "passes": int(passes), "band_only": int(band_only),
                   "answered_rank0": answered, "queries_rank0": Q},
        "clocks": clocks,
        "per_rank": per_rank,
    }
    line["roofline"]["band_kernel_ms"] = float(np.mean(band_ms))
    if rank == 0 and world == 1 and not args.no_extras:
        # reported extras: a failure in one of them is recorded, it must not take the bench line down
        for key, fn in (("kernels", lambda: map_kernel_rooflines(torch, fx, dev, flush, peak)),
                        ("latency", lambda: latency_probe(fx, m, s, g, args.hchoice)),
                        ("other_configs", lambda: other_configs(torch, fx, dev, flush, args.hchoice))):
            try:
                line[key] = fn()
            except Exception as exc:
                line[key] = {"error": repr(exc)}
        import oracle
        oracle.build()
        S, dt, used, cost, status = cpu_sample(m, s, g, args.hchoice, args.cpu_seconds, oracle)
        line["cpu_baseline"] = {"value": S / dt, "unit": "queries/s", "cores": int(used), "kind": "port",
                                "sample": "first %d queries of the step's batch (16 per thread), C restatement of scripts/jps1.py "
                                          "(oracle/fuxi_oracle.c) over %d OpenMP threads, %.1f s" % (S, used, dt)}
        # the sample doubles as a parity spot-check of the timed batch
        got = cf[:S]
        ok = status == 1
        bad = np.abs(got[ok] - cost[ok]) > 1e-5 * np.maximum(cost[ok], 1e-12) if args.hchoice == 2 else ci[:S][ok] != cost[ok]
        line["cpu_baseline"]["parity_mismatches"] = int(bad.sum()) + int(((ci[:S] >= 0) != ok).sum())
    if world > 1 and not args.no_extras:
        try:
            line["multi_gpu_checks"] = multi_gpu_checks(torch, dist, fx, dev, world, rank, m)
        except Exception as exc:          # a reported extra never takes the bench line down
            line["multi_gpu_checks"] = {"error": repr(exc)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------ other BASELINE configurations
CFG_METRIC = {
    "cfg2": ("cloud -> inflated occupancy grid, 1 Mi points -> 1024^2 cells, dense r=2 inflation: points/s", "points/s", True),
    "cfg3": ("planning queries/s, 1024^2 grid (20% fill), 4096 batched start/goal queries", "queries/s", True),
    "cfg5": ("single-query latency, 16384^2 grid (20% fill), first free cell -> last free cell", "ms", False),
}


def cfg2_cloud():
    """default_rng(1), x,y ~ U(-102.4, 102.4), z ~ U(-0.5, 3.0), float32; origin (-102.4, -102.4), reso 0.2."""
    rng = np.random.default_rng(1)
    N = 1 << 20
    pts = np.empty((N, 4), dtype=np.float32)
    pts[:, 0:2] = rng.uniform(-102.4, 102.4, (N, 2))
    pts[:, 2] = rng.uniform(-0.5, 3.0, N)
    pts[:, 3] = 0
    return pts


def cfg3_workload():
    m = (np.random.default_rng(2).random((1024, 1024)) < 0.2).astype(np.uint8)
    free = np.argwhere(m == 0)
    rng = np.random.default_rng(3)
    s = free[rng.integers(len(free), size=4096)].astype(np.int32)
    g = free[rng.integers(len(free), size=4096)].astype(np.int32)
    return m, s, g


def cfg5_workload(n=16384):
    m = (np.random.default_rng(6).random((n, n)) < 0.2).astype(np.uint8)
    ff = np.flatnonzero(m.reshape(-1) == 0)
    s = np.array(np.unravel_index(ff[0], m.shape), dtype=np.int32)
    g = np.array(np.unravel_index(ff[-1], m.shape), dtype=np.int32)
    return m, s, g


def base_line(args, world, value, ms_step, cfgd, dtype):
    name, unit, hib = CFG_METRIC[args.config]
    return {"metric": name, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": hib, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "config": cfgd}


def run_other_config(args, torch, dist, fx, ctx, dev, world, rank, local):
    peak, peak_src = measured_peak()
    flush_buf = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def flush():
        flush_buf.fill_(1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        l0 = ctx.launches
        ts = []
        for _ in range(args.steps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        barrier()
        clocks = sampler.stop()
        t = torch.tensor([float(sum(ts))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / args.steps, clocks, ctx.launches - l0

    if args.config == "cfg3":
        m, s_all, g_all = cfg3_workload()
        Q = len(s_all) // world
        s, g = s_all[rank * Q:(rank + 1) * Q], g_all[rank * Q:(rank + 1) * Q]
        d_m, d_s, d_g = torch.from_numpy(m).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(g).to(dev)
        ms, clocks, launches = timed(lambda: fx.plan_batch(d_m, d_s, d_g, metric=args.hchoice, max_path=args.max_path))
        band_ms, k_ms = fx.search_timings(ctx)
        settled = fx.search_stats(ctx)[0]
        fx.plan_host(m, s, g, metric=args.hchoice, max_path=args.max_path, ctx=ctx)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ci, cf, pxy, pl = fx.plan_host(m, s, g, metric=args.hchoice, max_path=args.max_path, ctx=ctx)
        e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        line = base_line(args, world, Q * world / (ms * 1e-3), ms,
                         {"workload": "cfg3: 4096 start/goal queries on a 1024x1024 random-obstacle grid (20%% fill, default_rng(2)/(3)), "
                                      "hchoice %d" % args.hchoice, "l2": "flushed between timed steps"},
                         "u32 (packed cost | arrival direction)")
        line["e2e"] = {"value": Q * world * args.steps / float(e2e_t.item()), "unit": "queries/s",
                       "h2d_bytes_per_step": int(m.nbytes + s.nbytes + g.nbytes), "d2h_bytes_per_step": int(ctx.lib.fx_last_d2h_bytes(ctx.handle)),
                       "api": "fx_plan_host"}
        line["gpu_launches"] = int(launches)
        ach = settled * 5.0 / (k_ms * 1e-3) / 1e9
        line["roofline"] = {"kernel": "k_search_batch", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                            "traffic": None, "peak_source": peak_src, "kernel_ms": k_ms, "band_kernel_ms": band_ms,
                            "algorithmic_bytes": "settled cells x 5 B"}
        line["search"] = {"settled_cells": int(settled), "nodes_per_s": settled / (ms * 1e-3)}
        line["clocks"] = clocks
        if rank == 0:
            import oracle
            oracle.build()
            S, dt, used, cost, status = cpu_sample(m, s, g, args.hchoice, args.cpu_seconds, oracle)
            ok = status == 1
            bad = np.abs(cf[:S][ok] - cost[ok]) > 1e-5 * np.maximum(cost[ok], 1e-12) if args.hchoice == 2 else ci[:S][ok] != cost[ok]
            line["cpu_baseline"] = {"value": S / dt, "unit": "queries/s", "cores": int(used), "kind": "port",
                                    "sample": "first %d queries, C restatement of scripts/jps1.py over %d threads, %.1f s" % (S, used, dt),
                                    "parity_mismatches": int(bad.sum()) + int(((ci[:S] >= 0) != ok).sum())}
            print(json.dumps(line), flush=True)
        return 0

    if args.config == "cfg2":
        import oracle
        pts = cfg2_cloud()
        N, W, H = len(pts), 1024, 1024
        d_p = torch.from_numpy(pts).to(dev)
        grid = torch.empty((W, H), dtype=torch.uint8, device=dev)
        infl = torch.empty_like(grid)

        def step():
            fx.project(d_p, None, 0.3, float("inf"), (-102.4, -102.4), 0.2, out=grid)
            fx.inflate(grid, 2, "ccst", out=infl)
        ms, clocks, launches = timed(step)
        want = oracle.hostref.project(pts[:, :3], np.eye(3, 4), 0.3, np.inf, -102.4, -102.4, 0.2, W, H)
        parity = int(np.array_equal(grid.cpu().numpy(), want)) + int(np.array_equal(infl.cpu().numpy(), oracle.inflate(want, 2, 1)))
        out = fx.map_host(pts, None, 0.3, float("inf"), (-102.4, -102.4), 0.2, (W, H), radius=2, variant="ccst", ctx=ctx)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = fx.map_host(pts, None, 0.3, float("inf"), (-102.4, -102.4), 0.2, (W, H), radius=2, variant="ccst", ctx=ctx)
        e2e_s = (time.perf_counter() - t0) / args.steps
        parity += int(np.array_equal(out, oracle.inflate(want, 2, 1)))
        alg = N * 16 + W * H + 2 * W * H
        line = base_line(args, world, N / (ms * 1e-3), ms,
                         {"workload": "cfg2: 1 Mi float4 points (default_rng(1)) -> 1024x1024 grid (reso 0.2, z > 0.3) + dense r=2 inflation",
                          "l2": "flushed between timed steps"}, "f32 transform / u8 grid")
        line["e2e"] = {"value": N / e2e_s, "unit": "points/s", "h2d_bytes_per_step": int(pts.nbytes), "d2h_bytes_per_step": W * H,
                       "api": "fx_map_host (host cloud in, inflated grid out)", "ms": e2e_s * 1e3}
        line["gpu_launches"] = int(launches)
        ach = alg / (ms * 1e-3) / 1e9
        line["roofline"] = {"kernel": "k_project_f4 + k_inflate_roll", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                            "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                            "algorithmic_bytes": "N*16 + W*H (projection) + 2*W*H (inflation); 20 MB per step: launch-latency-bound at this size, "
                                                 "see kernels.*_scaled of the cfg4 line for the HBM-exercising sizes"}
        line["clocks"] = clocks
        line["parity_checks_passed_of_3"] = parity
        # host baseline: the reference's numpy transform + height filter (a15/a16), the projection restatement, the a11 block
        t0 = time.perf_counter()
        p64 = pts[:, :3].astype(np.float64)
        e = oracle.hostref.transform_cloud(p64, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0))
        e = oracle.hostref.height_filter(e, 0.3)
        t1 = time.perf_counter()
        gr = oracle.hostref.project(pts[:, :3], np.eye(3, 4), 0.3, np.inf, -102.4, -102.4, 0.2, W, H)
        t2 = time.perf_counter()
        pad = np.zeros((W + 8, H + 8)); pad[4:-4, 4:-4] = gr
        oracle.hostref.inflate_ccst(pad, 2)
        t3 = time.perf_counter()
        line["cpu_baseline"] = {"value": N / (t3 - t0), "unit": "points/s", "cores": 1, "kind": "port",
                                "sample": "whole step once: numpy a15/a16 transform + height filter %.1f ms, projection restatement %.1f ms, "
                                          "reference inflation block a11 %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2))}
        if rank == 0:
            print(json.dumps(line), flush=True)
        return 0

    # ---- cfg5: one query on a 16384^2 grid.  N = 1: goal-directed batched search (bidirectional, ellipse-pruned) and the
    # whole cost field; N > 1: the row-tiled field over N slabs (halo exchange), with rank 0's one-GPU numbers beside it.
    from fuxi_planner_b200 import tiled
    n = 16384
    m, s, g = cfg5_workload(n)
    one = {}
    if rank == 0:
        d_m = torch.from_numpy(m).to(dev)
        d_s, d_g = torch.from_numpy(s[None]).to(dev), torch.from_numpy(g[None]).to(dev)
        res = fx.plan_batch(d_m, d_s, d_g, metric=args.hchoice, max_path=8192)
        torch.cuda.synchronize()
        ts = []
        for _ in range(max(args.steps, 1)):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); res = fx.plan_batch(d_m, d_s, d_g, metric=args.hchoice, max_path=8192); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        st = fx.search_stats(ctx)
        one["goal_directed_ms"] = float(np.mean(ts))
        one["goal_directed_settled_cells"] = int(st[0])
        one["goal_directed_levels"] = int(st[1])
        one["cost_i"] = int(res.cost_i[0])
        one["launches"] = 0
        fld = torch.empty((n, n), dtype=torch.int32, device=dev)
        fx.field(d_m, tuple(int(v) for v in s), metric=args.hchoice, out=fld)
        torch.cuda.synchronize()
        ts = []
        for _ in range(max(min(args.steps, 3), 1)):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fx.field(d_m, tuple(int(v) for v in s), metric=args.hchoice, out=fld, check=False); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        one["full_field_ms"] = float(np.mean(ts))
        one["full_field_goal_cost"] = int(fld[int(g[0]), int(g[1])])
        one["full_field_cells_reached"] = int((fld >= 0).sum())
        del fld
        if world > 1:
            del d_m
        torch.cuda.empty_cache()
    cfgd = {"workload": "cfg5: one query on a 16384x16384 random-obstacle grid (20%% fill, default_rng(6)), first free cell -> last free "
                        "cell, hchoice %d" % args.hchoice, "l2": "flushed between timed steps"}
    if world == 1:
        ms = one["goal_directed_ms"]
        line = base_line(args, world, ms, ms, cfgd, "u32 (packed cost | arrival direction)")
        line["scaling"] = "replicas only (a single query does not shard profitably; see DESIGN.md §5)"
        launches = 3
    else:
        x0, x1 = tiled.slab_bounds(n, world, rank)
        own = torch.from_numpy(m[x0:x1]).to(dev)
        fld, rounds = tiled.field_tiled(own, n, tuple(int(v) for v in s), metric=args.hchoice)      # warm-up (allocations, NCCL)
        barrier()
        ts = []
        for _ in range(max(min(args.steps, 3), 1)):
            t0 = time.perf_counter()
            fld, rounds = tiled.field_tiled(own, n, tuple(int(v) for v in s), metric=args.hchoice)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        t = torch.tensor([float(np.mean(ts))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        goal_cost = torch.tensor([int(fld[int(g[0]) - x0, int(g[1])]) if x0 <= int(g[0]) < x1 else -1], dtype=torch.int64, device=dev)
        dist.all_reduce(goal_cost, op=dist.ReduceOp.MAX)
        ms = float(t.item()) * 1e3
        line = base_line(args, world, ms, ms, cfgd, "i32 field")
        line["row_tiled"] = {"field_ms": ms, "exchange_rounds": int(rounds), "goal_cost": int(goal_cost.item()),
                             "slab_rows": n // world, "timing": "wall clock around the whole exchange loop, warm, max over ranks"}
        launches = 0
    line["one_gpu"] = one
    line["e2e"] = {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "note": "device-resident grid (256 MiB); the host-buffer form adds one 256 MiB H2D copy (~5 ms)"}
    line["gpu_launches"] = launches
    settled = one.get("goal_directed_settled_cells", 0)
    ach = settled * 5.0 / (one.get("goal_directed_ms", 1.0) * 1e-3) / 1e9 if settled else 0.0
    line["roofline"] = {"kernel": "k_search_batch (one CTA)", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes": "settled cells x 5 B"}
    line["clocks"] = None
    if rank == 0:
        # bounded CPU sample: the same grid, a query 1/8 of the way along the diagonal (the whole query costs > 1 min on one core)
        import oracle
        oracle.build()
        ff = np.argwhere(m[2040:2056, 2040:2056] == 0)[0] + 2040
        gs = np.array([ff], dtype=np.int32)
        t0 = time.perf_counter()
        cost, status, used = oracle.jps_batch(m, s[None], gs, args.hchoice, threads=1)
        dt = time.perf_counter() - t0
        r2 = fx.plan_batch(torch.from_numpy(m).to(dev), torch.from_numpy(s[None]).to(dev), torch.from_numpy(gs).to(dev), metric=args.hchoice, max_path=8192)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r2 = fx.plan_batch(torch.from_numpy(m).to(dev), torch.from_numpy(s[None]).to(dev), torch.from_numpy(gs).to(dev), metric=args.hchoice, max_path=8192); b.record()
        torch.cuda.synchronize()
        got = float(r2.cost_f[0]) if args.hchoice == 2 else float(r2.cost_i[0])
        line["cpu_baseline"] = {"value": dt * 1e3, "unit": "ms", "cores": 1, "kind": "port",
                                "sample": "same grid, first free cell -> (%d, %d) (1/8 of the diagonal): C restatement of jps1.py %.0f ms; "
                                          "this library on the same sub-query (incl. H2D of the grid) %.1f ms; costs agree: %s"
                                          % (int(gs[0][0]), int(gs[0][1]), dt * 1e3, a.elapsed_time(b), abs(got - float(cost[0])) <= 1e-5 * float(cost[0]))}
        print(json.dumps(line), flush=True)
    return 0


def run_reference_other(args, oracle):
    """CPU arm of the other configurations: the C restatement (or the numpy restatements for cfg2) on a bounded sample."""
    name, unit, hib = CFG_METRIC[args.config]
    threads = host_threads()
    if args.config == "cfg3":
        m, s, g = cfg3_workload()
        S = min(len(s), 16 * threads)
        fn = lambda: oracle.jps_batch(m, s[:S], g[:S], args.hchoice, threads=threads)
        units, sample = S, "first %d of the 4096 queries per step, C restatement of jps1.py over %d threads" % (S, threads)
    elif args.config == "cfg2":
        pts = cfg2_cloud()
        def fn():
            e = oracle.hostref.height_filter(oracle.hostref.transform_cloud(pts[:, :3].astype(np.float64), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0)), 0.3)
            gr = oracle.hostref.project(pts[:, :3], np.eye(3, 4), 0.3, np.inf, -102.4, -102.4, 0.2, 1024, 1024)
            pad = np.zeros((1032, 1032)); pad[4:-4, 4:-4] = gr
            oracle.hostref.inflate_ccst(pad, 2)
        units, sample, threads = len(pts), "the whole step: numpy a15/a16 + projection restatement + reference inflation block a11, 1 core", 1
    else:
        m, s, g = cfg5_workload()
        ff = np.argwhere(m[2040:2056, 2040:2056] == 0)[0] + 2040
        fn = lambda: oracle.jps_batch(m, s[None], np.array([ff], dtype=np.int32), args.hchoice, threads=1)
        units, sample, threads = 1, "same grid, first free cell -> (%d, %d) (1/8 of the diagonal), C restatement of jps1.py, 1 core" % (ff[0], ff[1]), 1
    for _ in range(min(args.warmup, 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    value = dt * 1e3 if unit == "ms" else units / dt
    line = {"impl": "reference", "metric": name, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": hib, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.config},
            "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(run_reference(a) if a.impl == "reference" else run_b200(a))
