"""Where does one 4096^2 drop-in replan spend its time?  (tuning experiment, run on the GPU box)
FUXI_B200_WIDE_BELOW=0 disables the wide-CTA form for comparison."""
import contextlib, io, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fuxi_planner_b200 as fx
from bench import make_workload

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m, s, g = make_workload(n, 64)
mf = m.astype(np.float64)
def med(f, k=9):
    ts = []
    for _ in range(k):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))
print("WIDE_BELOW=%s" % os.environ.get("FUXI_B200_WIDE_BELOW", "default"))
print("numpy (==1).astype(u8): %.2f ms" % med(lambda: (mf == 1).astype(np.uint8)))
sink = io.StringIO()
fx.plan_host(mf, [s[0]], [g[0]], metric=2, max_path=1024)
gm = torch.from_numpy(m).cuda()
tot = []
for q in range(12):
    with contextlib.redirect_stdout(sink):
        tm = med(lambda: fx.jps1.method(mf, tuple(s[q]), tuple(g[q]), 2), 7)
    th = med(lambda: fx.plan_host(m, [s[q]], [g[q]], metric=2, max_path=1024), 7)
    ss = torch.from_numpy(s[q:q+1]).cuda(); gg = torch.from_numpy(g[q:q+1]).cuda()
    def dev():
        fx.plan_batch(gm, ss, gg, metric=2, max_path=1024); torch.cuda.synchronize()
    td = med(dev, 7)
    tot.append(tm)
    print("q%d dist=%d method(f64) %.2f ms, plan_host(u8) %.2f ms, device-resident %.2f ms" % (q, int(np.abs(s[q]-g[q]).max()), tm, th, td))
print("median method(): %.2f ms" % float(np.median(tot)))
