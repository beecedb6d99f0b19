"""Tuning experiment: per-phase cycle breakdown of a level (needs a -DFX_PHASE_CLOCKS build selected by FUXI_B200_SO)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import fuxi_planner_b200 as fx

dev = torch.device("cuda:0")
Q = int(os.environ.get("Q", "8192"))
m, s, g = bench.make_workload(4096, Q)
ctx = fx.default_context(0)
if os.environ.get("FUXI_SLOTS"):
    ctx.lib.fx_set_search_tuning(ctx.handle, int(os.environ["FUXI_SLOTS"]), 0)
d_m, ds, dg = torch.from_numpy(m).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(g).to(dev)
for _ in range(2):
    fx.plan_batch(d_m, ds, dg, metric=2, max_path=1024)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); fx.plan_batch(d_m, ds, dg, metric=2, max_path=1024); b.record(); torch.cuda.synchronize()
st = fx.search_stats()
ph = (C.c_int64 * 8)()
ctx.lib.fx_search_phase_clocks.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
ctx.lib.fx_search_phase_clocks(ctx.handle, ph)
ph = list(ph)
names = ["queue load", "field+moves load", "ALU+scan", "reserve (smem atomic+shfl)", "relax+append", "barrier"]
rounds, levels = max(ph[6], 1), max(ph[7], 1)
print("SO=%s slots=%s Q=%d: %.1f ms, settled %d, levels %d (clocked levels %d, rounds %d = %.2f/level)" %
      (os.path.basename(os.environ.get("FUXI_B200_SO", "default")), os.environ.get("FUXI_SLOTS", "auto"), Q, a.elapsed_time(b), st[0], st[1], levels, rounds, rounds / levels))
tot = sum(ph[:6])
for n, v in zip(names, ph[:6]):
    print("   %-18s %8.0f cycles/level  %5.1f%%" % (n, v / levels, 100.0 * v / tot))
print("   total %.0f cycles/level" % (tot / levels))
