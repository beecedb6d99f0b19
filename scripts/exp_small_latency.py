import sys, time, io, contextlib, os
import numpy as np
sys.path.insert(0, ".")
import fuxi_planner_b200 as fx
z = np.load("tests/golden/maps.npz")
for name in ("-16.40-4.80_out.png", "-16.20-11.40_out.png"):
    m1 = z[name].astype(np.float64)
    free = np.argwhere(m1 == 0); rng = np.random.default_rng(0)
    pairs = [(tuple(free[rng.integers(len(free))]), tuple(free[rng.integers(len(free))])) for _ in range(400)]
    ts = []; ks = []
    with contextlib.redirect_stdout(io.StringIO()):
        for i, (a, b) in enumerate(pairs):
            t0 = time.perf_counter(); fx.jps1.method(m1, a, b, 2); dt = time.perf_counter() - t0
            if i >= 50: ts.append(dt); ks.append(fx.search_kernel_ms())
    ts = np.array(ts) * 1e3
    print(os.environ.get("FUXI_B200_SO", "default")[-12:], name, m1.shape, "p50 %.4f p90 %.4f kernel p50 %.4f" % (np.percentile(ts, 50), np.percentile(ts, 90), np.median(ks)))
