"""CUDA-event timings of fx_edt at 16384^2 for several fills (L2 flushed between runs) + exactness against the windowed
path on a sub-grid.  Tuning aid; bench.py's `kernels.edt_scaled` is the reported number."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fuxi_planner_b200 as fx
import oracle

dev = torch.device("cuda:0")
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
peak = 6547.2


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


# exactness on shapes that exercise tile edges, the inline fix, the fix list and the windowed fallback
rng = np.random.default_rng(11)
for (W, H, fill) in ((192, 256, 0.004), (1000, 1024, 0.02), (1031, 1056, 0.2), (640, 2048, 0.0005), (2048, 2048, 0.00002), (517, 96, 0.01), (128, 128, 0.0)):
    occ = (rng.random((W, H)) < fill).astype(np.uint8)
    got = fx.edt(torch.from_numpy(occ).to(dev)).cpu().numpy()
    want = oracle.edt(occ)
    assert np.array_equal(got, want), (W, H, fill, int((got != want).sum()))
    print("exact", W, H, fill)
n = 16384
for fill in (0.02, 0.2, 0.005):
    occ = (torch.rand((n, n), device=dev) < fill).to(torch.uint8)
    d2 = torch.empty((n, n), dtype=torch.int32, device=dev)
    m, mn = t(lambda: fx.edt(occ, out=d2))
    print("edt 16384^2 fill %.3f: %.3f ms (min %.3f)  frac %.3f" % (fill, m, mn, 5 * n * n / (m * 1e-3) / 1e9 / peak))
    del occ, d2
