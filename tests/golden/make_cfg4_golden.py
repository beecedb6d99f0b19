#!/usr/bin/env python
"""Golden costs for the BASELINE headline configuration (cfg4: 4096^2 grid default_rng(4), queries default_rng(5))
from the UNMODIFIED reference scripts/jps1.py -- build container only (needs /root/reference).

    python tests/golden/make_cfg4_golden.py [n_queries]

The reference needs minutes per query at this size (pure Python, O(|open|) list rebuild per successor,
jps1.py:224), so only a handful of the batch's queries are pinned this way: the three shortest of the first 256
plus the ones listed in PICK (a medium and a long one).  Every other query of the batch is checked against the C
restatement (oracle/fuxi_oracle.c), which is itself pinned on jps1_golden.json.
Writes tests/golden/cfg4_golden.json: [{"index", "start", "goal", "h", "cost", "secs"}].
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402


def workload(n=4096, q=256):
    m = (np.random.default_rng(4).random((n, n)) < 0.2).astype(np.uint8)
    free = np.argwhere(m == 0)
    rng = np.random.default_rng(5)
    s = free[rng.integers(len(free), size=8192)].astype(np.int32)[:q]
    g = free[rng.integers(len(free), size=8192)].astype(np.int32)[:q]
    return m, s, g


def main():
    assert refload.available(), "reference tree not found"
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    m, s, g = workload()
    d = np.maximum(np.abs(s[:, 0] - g[:, 0]), np.abs(s[:, 1] - g[:, 1]))
    order = np.argsort(d, kind="stable")
    # three shortest, then one near the median and one at the 90th percentile of the first 256 (by Chebyshev distance)
    pick = [int(v) for v in order[:3]] + [int(order[len(order) // 2]), int(order[int(len(order) * 0.9)])]
    pick = pick[:nq]
    mf = m.astype(np.float64)
    out = []
    path_out = os.path.join(HERE, "cfg4_golden.json")
    for q in pick:
        for h in (2, 1):
            t0 = time.time()
            path, cost, _ = refload.method(mf, tuple(int(v) for v in s[q]), tuple(int(v) for v in g[q]), h)
            rec = {"index": q, "start": [int(s[q][0]), int(s[q][1])], "goal": [int(g[q][0]), int(g[q][1])], "h": h,
                   "cost": None if (path == 0 and not isinstance(path, list)) else repr(float(cost)),
                   "secs": round(time.time() - t0, 1)}
            out.append(rec)
            print(rec, flush=True)
            with open(path_out, "w") as fh:
                json.dump(out, fh, separators=(",", ":"))
    print("done")


if __name__ == "__main__":
    main()
