"""p50 of the fused replan (fx_replan_host) and of the drop-in method() on the cfg1 map; tuning aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fuxi_planner_b200 as fx
from fuxi_planner_b200 import planner
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "maps.npz"))
m1 = z["-16.40-4.80_out.png"].astype(np.float64)
free = np.argwhere(m1 == 0)
rng = np.random.default_rng(0)
pairs = [(tuple(free[rng.integers(len(free))]), tuple(free[rng.integers(len(free))])) for _ in range(300)]
msg = planner.array_to_occupancy_grid(z["-16.40-4.80_out.png"])
Wm, Hm = m1.shape
for variant in ("st", "ccst"):
    ts = []
    for i, (a, b) in enumerate(pairs[:220]):
        st_xy = (-16.4 + 0.2 * (a[0] + 0.5), -4.8 + 0.2 * (a[1] + 0.5))
        go_xy = (-16.4 + 0.2 * (b[0] + 0.5), -4.8 + 0.2 * (b[1] + 0.5))
        t0 = time.perf_counter()
        planner.replan_fused(msg, Wm, Hm, (-16.4, -4.8), 0.2, st_xy, go_xy, ifa=1, variant=variant, hchoice=2)
        dt = time.perf_counter() - t0
        if i >= 20:
            ts.append(dt)
    ts = np.array(ts) * 1e3
    print("replan_fused %s: p50 %.4f p90 %.4f p99 %.4f ms" % (variant, np.percentile(ts, 50), np.percentile(ts, 90), np.percentile(ts, 99)))
import io, contextlib
sink = io.StringIO(); ts = []
with contextlib.redirect_stdout(sink):
    for i, (a, b) in enumerate(pairs):
        t0 = time.perf_counter(); fx.jps1.method(m1, a, b, 2); dt = time.perf_counter() - t0
        if i >= 20: ts.append(dt)
ts = np.array(ts) * 1e3
print("method() cfg1: p50 %.4f p90 %.4f p99 %.4f ms" % (np.percentile(ts, 50), np.percentile(ts, 90), np.percentile(ts, 99)))
