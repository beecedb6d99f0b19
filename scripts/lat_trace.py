"""Host-stage trace of single-query calls on the 4096^2 grid (FUXI_B200_TRACE=1) + raw pinned H2D time of the grid."""
import os, sys, time
os.environ["FUXI_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fuxi_planner_b200 as fx
n = 4096
m = (np.random.default_rng(4).random((n, n)) < 0.2).astype(np.uint8)
free = np.argwhere(m == 0)
rng = np.random.default_rng(5)
s = free[rng.integers(len(free), size=64)].astype(np.int32)
g = free[rng.integers(len(free), size=64)].astype(np.int32)
hp = torch.from_numpy(m).pin_memory(); d = torch.empty((n, n), dtype=torch.uint8, device="cuda:0")
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(hp, non_blocking=True); torch.cuda.synchronize()
    print("pinned H2D 16.7 MB: %.0f us" % (1e6 * (time.perf_counter() - t0)), file=sys.stderr)
for name, mat in (("uint8", m), ("float64", m.astype(np.float64))):
    print("==", name, file=sys.stderr)
    for i in range(30):
        t0 = time.perf_counter()
        fx.plan_host(mat, s[i:i + 1], g[i:i + 1], metric=2, max_path=2048)
        print("   python wall %.0f us" % (1e6 * (time.perf_counter() - t0)), file=sys.stderr)
