"""One single-query launch per form on the cfg4 grid (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fuxi_planner_b200 as fx
n = 4096
m = (np.random.default_rng(4).random((n, n)) < 0.2).astype(np.uint8)
free = np.argwhere(m == 0)
rng = np.random.default_rng(5)
s = free[rng.integers(len(free), size=8192)].astype(np.int32)
g = free[rng.integers(len(free), size=8192)].astype(np.int32)
dev = torch.device("cuda:0")
dm = torch.from_numpy(m).to(dev)
q = int(os.environ.get("QI", "3"))
ds, dg = torch.from_numpy(s[q:q + 1]).to(dev), torch.from_numpy(g[q:q + 1]).to(dev)
for _ in range(3):
    res = fx.plan_batch(dm, ds, dg, metric=2, max_path=2048)
torch.cuda.synchronize()
print(int(res.cost_i[0]), fx.search_stats())
