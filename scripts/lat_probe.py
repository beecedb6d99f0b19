"""Single-query latency on the cfg4 grid (4096^2): device-only (CUDA events around fx_search_batch, grid resident) and through
the host-buffer calls.  Tuning aid for the latency form (FUXI_B200_SO selects a build variant)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fuxi_planner_b200 as fx

n = 4096
m = (np.random.default_rng(4).random((n, n)) < 0.2).astype(np.uint8)
free = np.argwhere(m == 0)
rng = np.random.default_rng(5)
s = free[rng.integers(len(free), size=8192)].astype(np.int32)
g = free[rng.integers(len(free), size=8192)].astype(np.int32)
dev = torch.device("cuda:0")
dm = torch.from_numpy(m).to(dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 220
ts, passes, levels = [], [], []
for i in range(N):
    ds, dg = torch.from_numpy(s[i:i + 1]).to(dev), torch.from_numpy(g[i:i + 1]).to(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); res = fx.plan_batch(dm, ds, dg, metric=2, max_path=2048); b.record()
    torch.cuda.synchronize()
    if i >= 20:
        ts.append(a.elapsed_time(b))
        st = fx.search_stats()
        passes.append(st[2]); levels.append(st[1])
ts = np.array(ts)
print("device-only (moves + search, 1 query): p50 %.3f p90 %.3f p99 %.3f max %.3f ms; passes mean %.2f; levels mean %.0f; us/level %.2f"
      % (np.percentile(ts, 50), np.percentile(ts, 90), np.percentile(ts, 99), ts.max(), np.mean(passes), np.mean(levels), 1e3 * ts.sum() / np.sum(levels)))
for name, mat in (("uint8 host", m), ("float64 host", m.astype(np.float64))):
    tt = []
    for i in range(N):
        t0 = time.perf_counter()
        fx.plan_host(mat, s[i:i + 1], g[i:i + 1], metric=2, max_path=2048)
        if i >= 20:
            tt.append(1e3 * (time.perf_counter() - t0))
    tt = np.array(tt)
    print("%s: p50 %.3f p90 %.3f p99 %.3f ms" % (name, np.percentile(tt, 50), np.percentile(tt, 90), np.percentile(tt, 99)))
# a batch of 148 (one per SM)
ds, dg = torch.from_numpy(s[:148]).to(dev), torch.from_numpy(g[:148]).to(dev)
fx.plan_batch(dm, ds, dg, metric=2, max_path=2048)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); fx.plan_batch(dm, ds, dg, metric=2, max_path=2048); b.record(); torch.cuda.synchronize()
print("148 queries, one launch: %.3f ms" % a.elapsed_time(b))
