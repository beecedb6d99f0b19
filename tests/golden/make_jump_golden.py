#!/usr/bin/env python
"""Golden vectors for jps1.jump (scripts/jps1.py:95-164) from the UNMODIFIED reference: random (cell, direction, goal)
probes on five of the repo maps and on random grids.  Pins oracle.jump, which the GPU tests of the jump-point path form
use as the checker.  Build container only:   python tests/golden/make_jump_golden.py   -> jump_golden.json"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402

DIRS = [(-1, 0), (1, 0), (0, -1), (0, 1), (-1, -1), (-1, 1), (1, -1), (1, 1)]


def main():
    assert refload.available()
    mod = refload.jps1_module()
    z = np.load(os.path.join(HERE, "maps.npz"))
    rng = np.random.default_rng(5)
    out = []
    grids = [(k, z[k]) for k in sorted(z.files)[::7]]
    for i in range(6):
        W, H = int(rng.integers(8, 60)), int(rng.integers(8, 60))
        grids.append(("rand%d" % i, (rng.random((W, H)) < rng.choice([0.05, 0.2, 0.35])).astype(np.uint8)))
    for name, m in grids:
        mf = m.astype(np.float64)
        W, H = m.shape
        probes = []
        for _ in range(400):
            c = (int(rng.integers(W)), int(rng.integers(H)))
            d = DIRS[int(rng.integers(8))]
            g = (int(rng.integers(W)), int(rng.integers(H)))
            if rng.random() < 0.3:      # goal on the ray: the goal test fires
                k = int(rng.integers(1, 12))
                g = (c[0] + k * d[0], c[1] + k * d[1])
            try:
                r = mod.jump(c[0], c[1], d[0], d[1], mf, g)
            except IndexError:          # dblock has no bounds test (jps1.py:34-38)
                continue
            probes.append([c[0], c[1], d[0], d[1], g[0], g[1]] + ([int(r[0]), int(r[1])] if r is not None else [-1, -1]))
        rec = {"name": name, "probes": probes}
        if name.startswith("rand"):
            rec["W"], rec["H"], rec["grid"] = W, H, np.packbits(m).tolist()
        out.append(rec)
        print(name, len(probes), sum(p[6] >= 0 for p in probes))
    with open(os.path.join(HERE, "jump_golden.json"), "w") as fh:
        json.dump(out, fh, separators=(",", ":"))


if __name__ == "__main__":
    main()
