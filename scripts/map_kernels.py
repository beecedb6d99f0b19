"""Launches each map-side kernel once at the scaled sizes (64 M points, 16384^2) -- the ncu target for profiles/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fuxi_planner_b200 as fx

dev = torch.device("cuda:0")
N, n = 64 << 20, 16384
half = n * 0.1
pts = torch.empty((N, 4), dtype=torch.float32, device=dev)
pts[:, 0:2].uniform_(-half, half); pts[:, 2].uniform_(-0.5, 3.0); pts[:, 3] = 0
grid = torch.empty((n, n), dtype=torch.uint8, device=dev)
for _ in range(2):
    fx.project(pts, None, 0.3, float("inf"), (-half, -half), 0.2, out=grid)
p3 = pts[:, :3].contiguous()
for _ in range(2):
    fx.project(p3, None, 0.3, float("inf"), (-half, -half), 0.2, out=grid)
del pts, p3
occ = (torch.rand((n, n), device=dev) < 0.02).to(torch.uint8)
o2 = torch.empty_like(occ)
for _ in range(2):
    fx.inflate(occ, 2, "ccst", out=o2)
for _ in range(2):
    fx.inflate(occ, 1, "st", out=o2)
d2 = torch.empty((n, n), dtype=torch.int32, device=dev)
for _ in range(2):
    fx.edt(occ, out=d2)
torch.cuda.synchronize()
print("ok")
