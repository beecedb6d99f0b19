"""Experiment: how long do the largest queries of the bench batch take on their own? (tail analysis)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import fuxi_planner_b200 as fx

dev = torch.device("cuda:0")
m, s, g = bench.make_workload(4096, int(os.environ.get("Q", "8192")))
dx = np.abs(s[:, 0] - g[:, 0]).astype(np.int64); dy = np.abs(s[:, 1] - g[:, 1]).astype(np.int64)
est = np.minimum(dx, dy) * np.abs(dx - dy) + 64 * np.maximum(dx, dy)
order = np.argsort(-est)
d_m = torch.from_numpy(m).to(dev)

def run(idx, label):
    ds, dg = torch.from_numpy(s[idx]).to(dev), torch.from_numpy(g[idx]).to(dev)
    for _ in range(2):
        fx.plan_batch(d_m, ds, dg, metric=2, max_path=1024)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fx.plan_batch(d_m, ds, dg, metric=2, max_path=1024); b.record(); torch.cuda.synchronize()
    st = fx.search_stats()
    ms = a.elapsed_time(b)
    print("%-28s Q=%6d  %8.2f ms  settled %12d  levels %9d  -> %.1f Mnodes/s/query-slot-equivalent, %.2f Gnodes/s" %
          (label, len(idx), ms, st[0], st[1], st[0] / ms / 1e3 / max(1, min(len(idx), 888)), st[0] / ms / 1e6))

run(order[:1], "largest 1")
run(order[1:2], "2nd largest")
run(order[:8], "largest 8")
run(order[:148], "largest 148")
run(order[:888], "largest 888")
run(order[:2048], "largest 2048")
run(order[888:], "all but largest 888")
run(order[2048:], "all but largest 2048")
run(order, "all (LPT is internal anyway)")
run(order[-2048:], "smallest 2048")
