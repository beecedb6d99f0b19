"""Randomised differential test of the search against the Dijkstra oracle (test infrastructure: runs on the GPU box).
Random shapes, fills, structured obstacles (walls with gaps, rooms), both metrics, all three kernel forms
(shared-memory, batched throughput, batched wide/latency).  Usage: python scripts/fuzz_search.py [seconds] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import fuxi_planner_b200 as fx
import oracle
from util import validate_path

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1234)
dev = torch.device("cuda:0")
t_end = time.time() + budget
cases = queries = 0
while time.time() < t_end:
    kind = rng.integers(4)
    if kind == 0:      # small map -> shared-memory kernel (or batched when forced)
        W, H = int(rng.integers(1, 200)), int(rng.integers(1, 100))
    elif kind == 1:
        W, H = int(rng.integers(100, 700)), int(rng.integers(100, 700))
    elif kind == 2:
        W, H = int(rng.integers(2, 40)), int(rng.integers(500, 3000))
    else:
        W, H = int(rng.integers(300, 1500)), int(rng.integers(300, 1500))
    fill = float(rng.choice([0.0, 0.05, 0.2, 0.3, 0.42, 0.55]))
    m = (rng.random((W, H)) < fill).astype(np.uint8)
    for _ in range(int(rng.integers(0, 6))):            # walls with a gap, rooms
        if rng.random() < 0.5 and W > 4:
            x = int(rng.integers(W)); m[x, :] = 1; m[x, int(rng.integers(H))] = 0
        elif H > 4:
            y = int(rng.integers(H)); m[:, y] = 1; m[int(rng.integers(W)), y] = 0
    if rng.random() < 0.3:
        m[m == 0] = rng.choice(np.array([0, 0, 0, 100, 2], dtype=np.uint8), size=int((m == 0).sum()))   # non-1 values are free
    Q = int(rng.choice([1, 3, 40, 200, 1500]))
    s = np.c_[rng.integers(-1, W + 1, size=Q), rng.integers(-1, H + 1, size=Q)].astype(np.int32)
    g = np.c_[rng.integers(-1, W + 1, size=Q), rng.integers(-1, H + 1, size=Q)].astype(np.int32)
    small_off = rng.random() < 0.5
    if small_off:
        os.environ["FUXI_B200_SMALL"] = "0"
    else:
        os.environ.pop("FUXI_B200_SMALL", None)
    for metric in (1, 2):
        inb = (s[:, 0] >= 0) & (s[:, 0] < W) & (s[:, 1] >= 0) & (s[:, 1] < H)
        want = np.full(Q, -2, dtype=np.int64)
        if inb.any():
            want[inb] = oracle.sssp_batch(m, s[inb], g[inb], metric)
        res = fx.plan_batch(torch.from_numpy(m).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(g).to(dev), metric=metric, max_path=4096)
        got = res.cost_i.cpu().numpy().astype(np.int64)
        if not np.array_equal(got, want):
            bad = np.flatnonzero(got != want)[:5]
            print("MISMATCH shape", (W, H), "fill", fill, "Q", Q, "metric", metric, "small_off", small_off, "queries", bad, got[bad], want[bad], s[bad], g[bad])
            np.savez("gpurun_out/fuzz_fail.npz", m=m, s=s, g=g)
            sys.exit(1)
        pl = res.path_len.cpu().numpy()
        for i in np.flatnonzero(want > 0)[:10]:
            if pl[i] <= 4096:
                validate_path((m == 1).astype(np.uint8), res.path(int(i)), tuple(s[i]), tuple(g[i]))
        queries += Q
    cases += 1
print("fuzz ok: %d cases, %d queries x 2 metrics, no mismatch" % (cases, queries))
