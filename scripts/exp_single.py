"""One device-resident single-query plan_batch per query (for an ncu launch list: which kernel is the latency?)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fuxi_planner_b200 as fx
from bench import make_workload
m, s, g = make_workload(4096, 64)
gm = torch.from_numpy(m).cuda()
for rep in range(2):
    for q in (5, 0, 6):
        ss = torch.from_numpy(s[q:q+1]).cuda(); gg = torch.from_numpy(g[q:q+1]).cuda()
        fx.plan_batch(gm, ss, gg, metric=2, max_path=1024); torch.cuda.synchronize()
