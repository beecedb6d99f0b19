"""Randomised differential test of the map-side kernels against the oracle (test infrastructure, GPU box):
cloud filter, distance filter, projection, inflation, EDT.  Usage: python scripts/fuzz_map.py [seconds] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fuxi_planner_b200 as fx
import oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 99)
dev = torch.device("cuda:0")
t_end = time.time() + budget
n_cases = {"cloud": 0, "dist": 0, "project": 0, "inflate": 0, "edt": 0}
def fail(what, **kw):
    print("MISMATCH", what, kw); sys.exit(1)
while time.time() < t_end:
    # cloud filter
    n = int(rng.choice([0, 1, 50, 3000, 40000]))
    ext = rng.choice([1.0, 4.0, 30.0])
    stride = int(rng.choice([3, 4, 8]))
    p = np.zeros((n, stride), dtype=np.float32)
    p[:, :3] = rng.uniform(-ext, ext, (n, 3))
    if n > 10:
        k = n // 2
        p[:k, 2] = 2.0 + 0.05 * rng.standard_normal(k)          # a sheet
        p[rng.integers(n, size=3), rng.integers(3, size=3)] = [np.nan, np.inf, -np.inf]
    rgb = 4 if stride == 8 and rng.random() < 0.7 else -1
    if rgb >= 0:
        p[:, 4] = rng.integers(0, 1 << 24, size=n).astype(np.uint32).view(np.float32)
    leaf = tuple(float(v) for v in rng.choice([0.05, 0.1, 0.17, 0.2, 0.5], size=3))
    radius = float(rng.choice([0.12, 0.35, 0.6]))
    if radius / leaf[0] > 14:
        radius = leaf[0] * 3
    kw = dict(rgb_offset=rgb, pass_lim=(float(rng.uniform(-1, 1)), float(rng.uniform(2, 6))), leaf=leaf, radius=radius,
              min_neighbors=int(rng.integers(0, 20)))
    want, wc = oracle.cloud_filter(p, **kw)
    got, gc = fx.cloud.cloud_filter_host(p, **kw)
    if gc.tolist() != wc.tolist() or not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
        fail("cloud_filter", n=n, stride=stride, kw=kw, counts=(gc, wc))
    n_cases["cloud"] += 1
    # distance filter
    n = int(rng.choice([0, 7, 2048, 5000, 70000]))
    q = rng.uniform(-6, 6, (n, 3))
    if n > 100:
        q[: n // 3] = np.round(q[: n // 3])                        # exact ties
    dis = float(rng.choice([0.5, 4.0, 100.0]))
    w = oracle.hostref.distance_filter(q, dis)
    gt = fx.cloud.distance_filter_host(q, dis)
    if gt.shape != w.shape or not np.array_equal(gt.view(np.uint64), w.view(np.uint64)):
        fail("distance_filter", n=n, dis=dis)
    n_cases["dist"] += 1
    # projection + inflation + EDT on one grid
    W, H = int(rng.integers(1, 900)), int(rng.choice([16, 48, 130, 512, 1000, 1024]))
    npts = int(rng.choice([0, 5, 10000, 300000]))
    stride = int(rng.choice([3, 4]))
    pts = np.zeros((npts, stride), dtype=np.float32)
    pts[:, 0] = rng.uniform(-5, W * 0.2 + 5, npts); pts[:, 1] = rng.uniform(-5, H * 0.2 + 5, npts); pts[:, 2] = rng.uniform(-0.5, 3, npts)
    A = oracle.hostref.cloud_affine(tuple(rng.uniform(-0.2, 0.2, 3)), tuple(rng.uniform(-1, 1, 3))) if rng.random() < 0.5 else None
    zmax = float(rng.choice([np.inf, 2.0]))
    wantg = oracle.hostref.project(pts[:, :3], np.eye(3, 4) if A is None else A, 0.3, zmax, 0.0, 0.0, 0.2, W, H)
    gotg = fx.project(torch.from_numpy(pts).to(dev), A, 0.3, zmax, (0.0, 0.0), 0.2, (W, H))
    if not np.array_equal(gotg.cpu().numpy(), wantg):
        fail("project", W=W, H=H, npts=npts, stride=stride)
    n_cases["project"] += 1
    occ = (rng.random((W, H)) < rng.choice([0.0, 0.001, 0.02, 0.3])).astype(np.uint8) * rng.choice(np.array([1, 100, 7], dtype=np.uint8))
    r = int(rng.choice([0, 1, 2, 3, 4, 6, 16]))
    variant = str(rng.choice(["st", "ccst"]))
    step = 1 if variant == "ccst" else max(r, 1)
    if not np.array_equal(fx.inflate(torch.from_numpy(occ).to(dev), r, variant).cpu().numpy(), oracle.inflate(occ, r, step)):
        fail("inflate", W=W, H=H, r=r, variant=variant)
    n_cases["inflate"] += 1
    if not np.array_equal(fx.edt(torch.from_numpy(occ).to(dev)).cpu().numpy(), oracle.edt(occ)):
        fail("edt", W=W, H=H)
    n_cases["edt"] += 1
print("fuzz ok:", n_cases)
