"""Small workload through every kernel added in round 2 (the compute-sanitizer target of scripts/sanitize_r02.sh): the
three forms of the batched search on one grid (cluster: 8 queries, one CTA per query: 40, throughput: 400), compact and
jump-point path forms, the node-cloud kernel, EDT with the per-warp fix list; everything checked against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FUXI_B200_SMALL"] = "0"
import numpy as np
import torch
import fuxi_planner_b200 as fx
import oracle

dev = torch.device("cuda:0")
rng = np.random.default_rng(3)
m = (rng.random((300, 340)) < 0.25).astype(np.uint8)
m[100:140, 200] = 1; m[100:140, 240] = 1; m[100, 200:241] = 1; m[139, 200:241] = 1; m[110:130, 210:230] = 0   # a sealed room
free = np.argwhere(m == 0)
gm = torch.from_numpy(m).to(dev)
for Q in (8, 40, 400):
    s = free[rng.integers(len(free), size=Q)].astype(np.int32)
    g = free[rng.integers(len(free), size=Q)].astype(np.int32)
    g[1] = (120, 220)                      # goal inside the room: unreachable
    s[2] = np.argwhere(m == 1)[5]          # start on an obstacle
    for metric in (1, 2):
        want = oracle.sssp_batch(m, s, g, metric)
        res = fx.plan_batch(gm, torch.from_numpy(s).to(dev), torch.from_numpy(g).to(dev), metric=metric, max_path=1024)
        got = res.cost_i.cpu().numpy().astype(np.int64)
        assert np.array_equal(got, want), (Q, metric, np.flatnonzero(got != want)[:5])
        off, xy = fx.paths_compact(res.path_xy, res.path_len)
        assert int(off[-1]) == int(res.path_len.clamp(min=0).sum())
    print("search ok", Q, int((want >= 0).sum()), "reachable")
ctx = fx.Context(0)
ctx.set_search_form("throughput")
s = free[rng.integers(len(free), size=12)].astype(np.int32); g = free[rng.integers(len(free), size=12)].astype(np.int32)
res = fx.plan_batch(gm, torch.from_numpy(s).to(dev), torch.from_numpy(g).to(dev), metric=1, max_path=1024, ctx=ctx)
jxy, jl = fx.paths_jump_points(gm, res.path_xy, res.path_len, ctx=ctx)
jxy, jl = jxy.cpu().numpy(), jl.cpu().numpy()
for q in range(12):
    if jl[q] > 1:
        jp = [tuple(int(v) for v in p) for p in jxy[q, :jl[q]]]
        for p, r in zip(jp[:-1], jp[1:]):
            d = (int(np.sign(r[0] - p[0])), int(np.sign(r[1] - p[1])))
            assert oracle.jump(m, p, d, tuple(int(v) for v in g[q])) == r
ctx.close()
print("jump points ok")
cam = np.c_[rng.uniform(-4, 4, 3000), rng.uniform(-3, 3, 3000), rng.uniform(0.2, 7.0, 3000)].astype(np.float32)
got = fx.cloud.node_cloud_host(cam, (0.1, -0.05, 1.2), (1.0, 2.0, 1.5), 0.02, (0.1, 0.2, -0.1), (1.0, 0.0, 0.2))
want = oracle.hostref.node_cloud(cam, (0.1, -0.05, 1.2), (1.0, 2.0, 1.5), 0.02, (0.1, 0.2, -0.1), (1.0, 0.0, 0.2))
assert np.array_equal(got, want)
print("node cloud ok", len(got))
occ = (rng.random((192, 256)) < 0.004).astype(np.uint8)
assert np.array_equal(fx.edt(torch.from_numpy(occ).to(dev)).cpu().numpy(), oracle.edt(occ))
print("edt ok")
