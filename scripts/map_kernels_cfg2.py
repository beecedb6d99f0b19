"""Launches each map-side kernel at the BASELINE cfg2 size (1 Mi points, 1024^2) -- ncu launch-list target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fuxi_planner_b200 as fx
dev = torch.device("cuda:0")
N, n = 1 << 20, 1024
half = n * 0.1
pts = torch.empty((N, 4), dtype=torch.float32, device=dev)
pts[:, 0:2].uniform_(-half, half); pts[:, 2].uniform_(-0.5, 3.0); pts[:, 3] = 0
p3 = pts[:, :3].contiguous()
grid = torch.empty((n, n), dtype=torch.uint8, device=dev)
occ = (torch.rand((n, n), device=dev) < 0.02).to(torch.uint8)
o2 = torch.empty_like(occ)
d2 = torch.empty((n, n), dtype=torch.int32, device=dev)
for _ in range(3):
    fx.project(pts, None, 0.3, float("inf"), (-half, -half), 0.2, out=grid)
    fx.project(p3, None, 0.3, float("inf"), (-half, -half), 0.2, out=grid)
    fx.inflate(occ, 2, "ccst", out=o2)
    fx.inflate(occ, 1, "st", out=o2)
    fx.edt(occ, out=d2)
torch.cuda.synchronize()
print("ok")
