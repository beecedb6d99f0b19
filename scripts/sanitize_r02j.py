"""Small workload through the kernels written in the last sessions of round 2 (compute-sanitizer target of
scripts/sanitize_r02j.sh): the warp-strip EDT (incl. ragged shapes, the fix-up list and the windowed fallback), graph
replays of the EDT / cloud filter / distance filter, the staged host uploads, the mapped fx_replan_host path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fuxi_planner_b200 as fx
import oracle
from fuxi_planner_b200 import planner

dev = torch.device("cuda:0")
rng = np.random.default_rng(7)
for (W, H, fill) in ((70, 64, 0.05), (130, 96, 0.02), (257, 160, 0.3), (96, 224, 0.001), (64, 32, 0.0)):
    occ = (rng.random((W, H)) < fill).astype(np.uint8)
    t = torch.from_numpy(occ).to(dev); d2 = torch.empty((W, H), dtype=torch.int32, device=dev)
    for _ in range(3):                       # third call replays the graph
        fx.edt(t, out=d2)
        assert np.array_equal(d2.cpu().numpy(), oracle.edt(occ)), (W, H, fill)
print("edt ok")
pts = torch.empty((4000, 3), dtype=torch.float32, device=dev)
out = torch.empty((4000, 4), dtype=torch.float32, device=dev); cnt = torch.empty(4, dtype=torch.int64, device=dev)
for it in range(3):
    host = np.c_[rng.uniform(-2, 2, (4000, 2)), 2.0 + 0.05 * rng.standard_normal(4000)].astype(np.float32)
    pts.copy_(torch.from_numpy(host))
    fx.cloud.cloud_filter(pts, out=out, counts=cnt)
    want, wc = oracle.cloud_filter(host)
    assert cnt.cpu().numpy().tolist() == wc.tolist()
print("cloud ok")
occ = (rng.random((1100, 1000)) < 0.2).astype(np.uint8)      # > 2^20 cells: staged upload
free = np.argwhere(occ == 0)
s = free[rng.integers(len(free), size=3)].astype(np.int32); g = free[rng.integers(len(free), size=3)].astype(np.int32)
want = oracle.sssp_batch(occ, s, g, 1)
for grid in (occ, occ.astype(np.float64)):
    r = fx.plan_host(grid, s, g, metric=1, max_path=512)
    assert np.array_equal(np.asarray(r[0], dtype=np.int64), want)
print("staged plan_host ok")
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "maps.npz"))
m1 = z["-16.40-4.80_out.png"]
msg = planner.array_to_occupancy_grid(m1)
fr = np.argwhere(m1 == 0)
for variant in ("st", "ccst"):
    for i in range(3):
        a, b = fr[rng.integers(len(fr))], fr[rng.integers(len(fr))]
        planner.replan_fused(msg, m1.shape[0], m1.shape[1], (-16.4, -4.8), 0.2, (-16.4 + 0.2 * (a[0] + 0.5), -4.8 + 0.2 * (a[1] + 0.5)),
                             (-16.4 + 0.2 * (b[0] + 0.5), -4.8 + 0.2 * (b[1] + 0.5)), ifa=1, variant=variant, hchoice=2)
print("replan ok")
