"""cfg3 (4096 queries, 1024^2, 20 % fill): one plan_batch per metric -- ncu launch-list target / stats."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fuxi_planner_b200 as fx
n, Q = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = (np.random.default_rng(2).random((n, n)) < 0.2).astype(np.uint8)
free = np.argwhere(m == 0)
rng = np.random.default_rng(3)
s = free[rng.integers(len(free), size=Q)].astype(np.int32)
g = free[rng.integers(len(free), size=Q)].astype(np.int32)
d_m, d_s, d_g = torch.from_numpy(m).cuda(), torch.from_numpy(s).cuda(), torch.from_numpy(g).cuda()
for _ in range(3):
    res = fx.plan_batch(d_m, d_s, d_g, metric=2, max_path=1024)
torch.cuda.synchronize()
print("stats (settled, levels, passes, band_only):", fx.search_stats(), "kernel ms", fx.search_kernel_ms())
