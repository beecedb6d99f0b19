"""p50 of the drop-in jps1.method on the 4096^2 headline grid (float64 matrix), 200 queries; FUXI_B200_TRACE=1 for the host stages."""
import sys, os, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fuxi_planner_b200 as fx
from bench import make_workload
m, s, g = make_workload(4096, 8192)
m4 = m.astype(np.float64)
ts = []
sink = io.StringIO()
with contextlib.redirect_stdout(sink):
    for i in range(220):
        t0 = time.perf_counter(); fx.jps1.method(m4, tuple(int(v) for v in s[i]), tuple(int(v) for v in g[i]), 2); dt = time.perf_counter() - t0
        if i >= 20: ts.append(dt * 1e3)
ts = np.array(ts)
print("method() f64 4096^2: p50 %.2f p90 %.2f p99 %.2f max %.2f ms" % (np.percentile(ts, 50), np.percentile(ts, 90), np.percentile(ts, 99), ts.max()))
tt = []
for i in range(220):
    t0 = time.perf_counter(); fx.plan_host(m4, s[i:i + 1], g[i:i + 1], metric=2, max_path=2048); dt = time.perf_counter() - t0
    if i >= 20: tt.append(dt * 1e3)
print("plan_host f64: p50 %.2f" % np.percentile(tt, 50))
