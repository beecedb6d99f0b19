#!/usr/bin/env python
"""Golden vectors for the line-of-sight shortcutting of global_planner_ccst.py:515-521, generated with the
UNMODIFIED reference's global_planner.map_line_col (ccst:258-283).

Run in the build container only (needs /root/reference):   python tests/golden/make_shortcut_golden.py
The greedy loop itself is inline in the reference's __main__ block (not importable); it is reproduced here
statement for statement around the reference's own function:

    ii=1
    while ii<len(path2_c)-1:
        if planner.map_line_col(path2_c[ii+1],path2_c[ii-1],mapu[min(..x..):max(..x..),min(..y..):max(..y..)]):
            path2_c  = np.delete(path2_c, ii ,axis =0)
        else:
            ii+=1

Inputs: the jump-point paths of jps1_golden.json on the reference maps (hchoice 2) and random polylines on random
grids.  Writes shortcut_golden.json next to this file.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import refload  # noqa: E402
import make_golden  # noqa: E402  (stub_ros)


def shortcut(gp, path, mapu):
    path2_c = np.array(path)
    ii = 1
    with contextlib.redirect_stdout(io.StringIO()):
        while ii < len(path2_c) - 1:
            a, b = path2_c[ii - 1], path2_c[ii + 1]
            if gp.map_line_col(path2_c[ii + 1], path2_c[ii - 1],
                               mapu[min(a[0], b[0]):max(a[0], b[0]), min(a[1], b[1]):max(a[1], b[1])]):
                path2_c = np.delete(path2_c, ii, axis=0)
            else:
                ii += 1
    return path2_c.tolist()


def main():
    assert refload.available(), "reference tree not found"
    make_golden.stub_ros()
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(refload.REF_ROOT, "scripts"))
    import global_planner_ccst as ref_gp
    gp = ref_gp.global_planner.__new__(ref_gp.global_planner)
    maps = np.load(os.path.join(HERE, "maps.npz"))
    gold = json.load(open(os.path.join(HERE, "jps1_golden.json")))
    out = {"maps": {}, "random": []}
    for name, recs in gold["maps"].items():
        m = maps[name].astype(np.float64)
        rows = []
        for r in recs:
            if r["h"] != 2 or r.get("path") is None or len(r["path"]) < 3:
                continue
            rows.append({"path": r["path"], "out": shortcut(gp, r["path"], m)})
        out["maps"][name] = rows
    rng = np.random.default_rng(21)
    for i in range(60):
        W, H = int(rng.integers(8, 70)), int(rng.integers(8, 70))
        m = (rng.random((W, H)) < rng.uniform(0.02, 0.3)).astype(np.float64)
        n = int(rng.integers(2, 14))
        path = np.c_[rng.integers(0, W, n), rng.integers(0, H, n)].tolist()
        if i % 5 == 0:      # axis-aligned and repeated points: the degenerate crops
            path[1] = [path[0][0], path[1][1]]
            if n > 3:
                path[3] = list(path[2])
        out["random"].append({"W": W, "H": H, "grid": np.packbits(m.astype(np.uint8)).tolist(), "path": path,
                              "out": shortcut(gp, path, m)})
    with open(os.path.join(HERE, "shortcut_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("maps:", sum(len(v) for v in out["maps"].values()), "random:", len(out["random"]))


if __name__ == "__main__":
    main()
