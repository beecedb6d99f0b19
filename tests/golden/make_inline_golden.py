#!/usr/bin/env python
"""Golden vectors for the planner loops' INLINE blocks, produced by executing the
UNMODIFIED reference source lines themselves.

The index/pad/shift + inflation + goal-relocation code sits in the `if __name__ == '__main__':` bodies of
scripts/global_planner_st.py (lines 226-275) and scripts/global_planner_ccst.py (lines 411-464), so it cannot be
called; this script reads those exact line ranges from the reference files, drops comment-only lines, dedents them and
`exec`s them in a namespace holding the variables the loop has at that point (mapu, map_o, map_c, map_r, map_reso,
global_goal, px, py, ifa, np).  Inputs and the resulting variables are stored; nothing is restated here.

Run in the build container only (needs /root/reference):   python tests/golden/make_inline_golden.py
Writes inline_golden.json next to this file (arrays: zlib + base64 of the uint8 bytes).
"""
import base64
import json
import os
import sys
import textwrap
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402

SLICES = {"st": ("global_planner_st.py", 226, 275), "ccst": ("global_planner_ccst.py", 411, 464)}


def load_block(variant):
    fname, lo, hi = SLICES[variant]
    with open(os.path.join(refload.REF_ROOT, "scripts", fname), encoding="utf-8", errors="replace") as fh:
        lines = fh.read().split("\n")[lo - 1:hi]
    first = lines[0].strip()
    assert first.startswith("map_goal=((global_goal[0:2]-map_o)/map_reso).astype(int)"), first
    assert lines[-1].strip() == "end_occu = 0", lines[-1]
    body = [ln for ln in lines if ln.strip() and not ln.strip().startswith("#")]
    return compile(textwrap.dedent("\n".join(body)) + "\n", "%s:%d-%d" % (fname, lo, hi), "exec")


def b64(a, dt):
    return base64.b64encode(zlib.compress(np.ascontiguousarray(a, dtype=dt).tobytes(), 9)).decode()


def main():
    assert refload.available(), "reference tree not found"
    rng = np.random.default_rng(77)
    out = {"slices": {k: list(v) for k, v in SLICES.items()}, "cases": []}
    code = {v: load_block(v) for v in SLICES}
    for i in range(160):
        variant = "st" if i % 2 == 0 else "ccst"
        ifa = int([1, 2, 3, 1][(i // 2) % 4])
        W, H = int(rng.integers(3, 70)), int(rng.integers(3, 70))
        fill = float([0.0, 0.02, 0.1, 0.3, 0.6][i % 5])
        # decoded OccupancyGrid values (map_callback: 100 -> 1, -1 -> 0, 1..99 survive and count as occupied in `> 0`)
        vals = rng.integers(1, 100, (W, H))
        vals[rng.random((W, H)) < 0.7] = 1
        X = ((rng.random((W, H)) < fill) * vals).astype(np.int64)
        if i % 9 == 4:                      # dense map: rows without a free cell -> relocation falls back to the column
            X[:, :] = np.where(rng.random((W, H)) < 0.5, 1, X)
        if i % 40 == 13:                     # everything occupied: only the zero padding the block adds is free (so the row scan always succeeds)
            X[:, :] = 1
        reso = float(rng.choice([0.1, 0.2, 0.25]))
        map_o = rng.uniform(-20, 20, 2).round(2)
        ext = np.array([W, H]) * reso
        lo = -0.35 if i % 3 == 0 else 0.02       # start / goal sometimes below the map origin (negative indices)
        hi = 1.3 if i % 4 == 1 else 0.98         # ... or beyond the far edge
        start = map_o + rng.uniform(lo, hi, 2) * ext
        goal = map_o + rng.uniform(lo, hi, 2) * ext
        if i % 5 == 2 and X.any():               # goal exactly on an occupied cell
            occ = np.argwhere(X > 0)
            c = occ[rng.integers(len(occ))]
            goal = map_o + (c + 0.5) * reso
        ns = {"np": np, "mapu": X.copy(), "map_o": [float(map_o[0]), float(map_o[1])], "map_c": W, "map_r": H,
              "map_reso": reso, "global_goal": np.array([goal[0], goal[1], 1.0]), "px": float(start[0]), "py": float(start[1]),
              "ifa": ifa}
        rec = {"variant": variant, "ifa": ifa, "W": W, "H": H, "X": b64(X, np.uint8), "map_o": [repr(float(v)) for v in map_o],
               "reso": repr(reso), "start": [repr(float(v)) for v in start], "goal": [repr(float(v)) for v in goal]}
        try:
            exec(code[variant], ns)
            g = np.asarray(ns["mapu"])
            rec["out"] = {"shape": list(g.shape), "grid": b64(g, np.uint8), "map_start": [int(v) for v in ns["map_start"]],
                          "map_goal": [int(v) for v in ns["map_goal"]], "map_goal0": [int(v) for v in ns["map_goal0"]],
                          "map_d": [int(v) for v in ns["map_d"]], "map_o": [repr(float(v)) for v in ns["map_o"]],
                          "map_c": int(ns["map_c"]), "map_r": int(ns["map_r"]), "end_occu": int(ns["end_occu"])}
            assert set(np.unique(g)).issubset(set(range(0, 100)))
        except Exception as exc:                 # the reference raises (e.g. goal index beyond the padded grid)
            rec["out"] = None
            rec["raises"] = type(exc).__name__
        out["cases"].append(rec)
    with open(os.path.join(HERE, "inline_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    ok = [r for r in out["cases"] if r["out"] is not None]
    print("cases:", len(out["cases"]), "ok:", len(ok), "raises:", len(out["cases"]) - len(ok),
          "relocated:", sum(r["out"]["map_goal"] != [a + b - (1 if r["variant"] == "st" else 0) for a, b in zip(r["out"]["map_goal0"], r["out"]["map_d"])] for r in ok),
          "end_occu:", sum(r["out"]["end_occu"] for r in ok))


if __name__ == "__main__":
    main()
