#!/bin/bash
# graph cache on / off: tests of the wrapped entry points + small-size timings
OUT=gpurun_out/${1:-graphs}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "cloud or edt or distance or filter or node" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
for g in 0 1; do
echo "== FUXI_B200_GRAPHS=$g"
FUXI_B200_GRAPHS=$g timeout 600 python - <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import fuxi_planner_b200 as fx
dev = torch.device("cuda:0")
def t(fn, reps=20):
    for _ in range(5): fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
N = 640 * 480
c = torch.zeros((N, 8), dtype=torch.float32, device=dev)
c[:, 0].uniform_(-3.0, 3.0); c[:, 1].uniform_(-2.0, 2.0); c[:, 2].uniform_(-0.5, 5.0)
c[: N // 2, 2] = 3.0 + 0.03 * torch.randn(N // 2, device=dev)
print("cloud_filter frame: %.4f ms" % t(lambda: fx.cloud.cloud_filter(c, rgb_offset=4)))
occ = (torch.rand((1024, 1024), device=dev) < 0.02).to(torch.uint8); d2 = torch.empty((1024, 1024), dtype=torch.int32, device=dev)
print("edt 1024^2: %.4f ms" % t(lambda: fx.edt(occ, out=d2)))
d = torch.empty((5000, 3), dtype=torch.float64, device=dev).uniform_(-6.0, 6.0)
print("distance_filter 5000 pts: %.4f ms" % t(lambda: fx.cloud.distance_filter(d, 4.0)))
d = torch.empty((1 << 20, 3), dtype=torch.float64, device=dev).uniform_(-6.0, 6.0)
print("distance_filter 1Mi pts: %.4f ms" % t(lambda: fx.cloud.distance_filter(d, 4.0)))
PY
done
