"""CUDA-event timings of the map-side kernels at the scaled sizes (64 Mi points -> 16384^2), L2 flushed between runs.
Tuning aid for A/B runs (e.g. FUXI_B200_PROJ_PART=0); bench.py's `kernels` block is the reported number."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fuxi_planner_b200 as fx

dev = torch.device("cuda:0")
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def t(fn, reps=7):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush_buf.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


N, n = 64 << 20, 16384
half = n * 0.1
pts = torch.empty((N, 4), dtype=torch.float32, device=dev)
pts[:, 0:2].uniform_(-half, half); pts[:, 2].uniform_(-0.5, 3.0); pts[:, 3] = 0
grid = torch.empty((n, n), dtype=torch.uint8, device=dev)
peak = 6547.2
m, mn = t(lambda: fx.project(pts, None, 0.3, float("inf"), (-half, -half), 0.2, out=grid))
print("project_f4 64Mi -> 16384^2: %.3f ms (min %.3f)  frac %.3f" % (m, mn, (N * 16 + n * n) / (m * 1e-3) / 1e9 / peak))
p3 = pts[:, :3].contiguous()
m, mn = t(lambda: fx.project(p3, None, 0.3, float("inf"), (-half, -half), 0.2, out=grid))
print("project_xyz 64Mi -> 16384^2: %.3f ms (min %.3f)  frac %.3f" % (m, mn, (N * 12 + n * n) / (m * 1e-3) / 1e9 / peak))
# a spatially coherent cloud (a scan: neighbouring points fall into neighbouring cells)
idx = torch.arange(N, device=dev, dtype=torch.float32)
pts[:, 0] = ((idx % 16384) * 0.2 - half + 0.05)
pts[:, 1] = ((idx // 16384 * 4 % 16384) * 0.2 - half + 0.05)
m, mn = t(lambda: fx.project(pts, None, 0.3, float("inf"), (-half, -half), 0.2, out=grid))
print("project_f4 coherent 64Mi -> 16384^2: %.3f ms (min %.3f)  frac %.3f" % (m, mn, (N * 16 + n * n) / (m * 1e-3) / 1e9 / peak))
del pts, p3, idx
for fill in (0.02, 0.2):
    occ = (torch.rand((n, n), device=dev) < fill).to(torch.uint8)
    d2 = torch.empty((n, n), dtype=torch.int32, device=dev)
    m, mn = t(lambda: fx.edt(occ, out=d2))
    print("edt 16384^2 fill %.2f: %.3f ms (min %.3f)  frac %.3f" % (fill, m, mn, 5 * n * n / (m * 1e-3) / 1e9 / peak))
    del occ, d2
