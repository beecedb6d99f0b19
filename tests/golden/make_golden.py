#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Writes, next to this file:
  maps.npz            the 35 reference maps/*.png decoded per SURVEY Appendix B (uint8 [x][y], 1 = occupied)
  jps1_golden.json    jps1.method results (cost printed at jps1.py:207 + jump-point path) for
                        * every map: first-free -> last-free + 20 random free pairs (default_rng(1)), hchoice 1 and 2
                        * cfg1 map: 300 random free pairs (default_rng(0)), hchoice 1 and 2 (costs only)
                        * edge-case toy grids
                        * random small grids (<= 32x32, 5-50 % fill)
                        * a few queries on 256^2 / 512^2 / 1024^2 20 %-fill grids (cfg3 seeds)
  hostfn_golden.npz   outputs of reference functions importable with stubbed ROS modules:
                        utils.body_to_earth_frame, convert_plc.distance_filter, global_planner.map_line_col
Nothing here is read from /root/reference at test time; only the files written are.
"""
import json
import os
import sys
import types

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refload, hostref  # noqa: E402


def run(m, s, g, h, keep_path=True):
    path, cost, _ = refload.method(m, tuple(int(v) for v in s), tuple(int(v) for v in g), h)
    rec = {"start": [int(s[0]), int(s[1])], "goal": [int(g[0]), int(g[1])], "h": h,
           "cost": None if path == 0 and not isinstance(path, list) else repr(float(cost))}
    if keep_path and rec["cost"] is not None:
        rec["path"] = [[int(p[0]), int(p[1])] for p in path]
    return rec


def stub_ros():
    names = ["rospy", "tf", "roslib", "sensor_msgs", "sensor_msgs.msg", "sensor_msgs.point_cloud2",
             "geometry_msgs", "geometry_msgs.msg", "visualization_msgs", "visualization_msgs.msg",
             "nav_msgs", "nav_msgs.msg", "message_filters", "std_msgs", "std_msgs.msg", "sklearn", "sklearn.cluster"]

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return type(k, (), {"__init__": lambda self, *a, **kw: None})

    for n in names:
        if n not in sys.modules:
            sys.modules[n] = _Any(n)


def main():
    assert refload.available(), "reference tree not found"
    out = {"maps": {}, "cfg1": {}, "edge": [], "random_small": [], "large": []}
    grids = {}
    for f in refload.map_files():
        name = os.path.basename(f)
        m = hostref.png_to_grid(np.array(Image.open(f).convert("L")))
        grids[name] = m
        mf = m.astype(np.float64)
        free = np.argwhere(m == 0)
        rng = np.random.default_rng(1)
        recs = []
        pairs = [(free[0], free[-1])]
        for _ in range(20):
            s = free[rng.integers(len(free))]
            g = free[rng.integers(len(free))]
            pairs.append((s, g))
        for s, g in pairs:
            for h in (1, 2):
                recs.append(run(mf, s, g, h))
        out["maps"][name] = recs
        print(name, m.shape, int(m.sum()), recs[0]["cost"], recs[1]["cost"], flush=True)
    np.savez_compressed(os.path.join(HERE, "maps.npz"), **grids)

    # cfg1: 300 random pairs on maps/-16.40-4.80_out.png
    m = grids["-16.40-4.80_out.png"]
    mf = m.astype(np.float64)
    free = np.argwhere(m == 0)
    rng = np.random.default_rng(0)
    recs = []
    for _ in range(300):
        s = free[rng.integers(len(free))]
        g = free[rng.integers(len(free))]
        for h in (1, 2):
            recs.append(run(mf, s, g, h, keep_path=False))
    out["cfg1"] = recs

    # edge cases
    def edge(name, m, s, g):
        for h in (1, 2):
            r = run(np.array(m, dtype=np.float64), s, g, h)
            r["name"] = name
            r["grid"] = np.array(m).tolist()
            out["edge"].append(r)

    z = np.zeros((6, 6))
    edge("start_eq_goal", z, (2, 2), (2, 2))
    m = z.copy(); m[5, 5] = 1
    edge("goal_on_obstacle", m, (0, 0), (5, 5))
    m = z.copy(); m[3, 3] = 1
    edge("start_on_obstacle", m, (3, 3), (5, 5))
    m = z.copy(); m[3, :] = 1
    edge("wall_between", m, (0, 0), (5, 5))
    m = np.zeros((4, 4)); m[1, 0] = 1; m[0, 1] = 1
    edge("diagonal_squeeze", m, (0, 0), (3, 3))
    m = np.zeros((4, 4)); m[1, 0] = 1
    edge("single_corner_cut", m, (0, 0), (1, 1))
    m = np.zeros((5, 5)); m[2, :] = 100
    edge("value_100_is_free", m, (0, 2), (4, 2))
    m = np.zeros((5, 5)); m[2, 1:] = 1
    edge("gap_at_edge", m, (0, 4), (4, 4))
    m = np.zeros((1, 7))
    edge("one_row", m, (0, 0), (0, 6))
    m = np.zeros((7, 1)); m[3, 0] = 1
    edge("one_col_blocked", m, (0, 0), (6, 0))

    # random small grids
    rng = np.random.default_rng(7)
    for i in range(120):
        W, H = int(rng.integers(3, 33)), int(rng.integers(3, 33))
        fill = float(rng.uniform(0.05, 0.5))
        m = (rng.random((W, H)) < fill).astype(np.uint8)
        free = np.argwhere(m == 0)
        if len(free) < 2:
            continue
        qs = []
        for _ in range(4):
            s = free[rng.integers(len(free))]
            g = free[rng.integers(len(free))]
            for h in (1, 2):
                qs.append(run(m.astype(np.float64), s, g, h, keep_path=(i < 30)))
        out["random_small"].append({"W": W, "H": H, "grid": np.packbits(m).tolist(), "queries": qs})

    # a few large queries (cfg3 seeds: grid default_rng(2), queries default_rng(3))
    for n, nq in ((256, 6), (512, 4), (1024, 3)):
        if n == 1024:
            m = (np.random.default_rng(2).random((1024, 1024)) < 0.2).astype(np.uint8)
            qrng = np.random.default_rng(3)
        else:
            m = (np.random.default_rng(100 + n).random((n, n)) < 0.2).astype(np.uint8)
            qrng = np.random.default_rng(200 + n)
        free = np.argwhere(m == 0)
        qs = []
        for _ in range(nq):
            s = free[qrng.integers(len(free))]
            g = free[qrng.integers(len(free))]
            for h in (1, 2):
                qs.append(run(m.astype(np.float64), s, g, h, keep_path=False))
                print("large", n, qs[-1], flush=True)
        out["large"].append({"n": n, "grid_seed": 2 if n == 1024 else 100 + n,
                             "query_seed": 3 if n == 1024 else 200 + n, "queries": qs})

    with open(os.path.join(HERE, "jps1_golden.json"), "w") as fh:
        json.dump(out, fh, separators=(",", ":"))

    # ---- reference host functions (ROS stubbed) -------------------------------------------------
    stub_ros()
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(refload.REF_ROOT, "scripts"))
    import utils as ref_utils  # noqa: E402
    import plc_point2_st as ref_plc  # noqa: E402
    import global_planner_ccst as ref_gp  # noqa: E402
    rng = np.random.default_rng(11)
    rpy = rng.uniform(-1.0, 1.0, (16, 3))
    Rs = np.stack([ref_utils.body_to_earth_frame(*a) for a in rpy])
    conv = ref_plc.convert_plc.__new__(ref_plc.convert_plc)
    pts = rng.uniform(-5, 5, (500, 3))
    pts[10] = pts[11]  # exact tie on every key
    df = conv.distance_filter(pts.copy(), 4)
    gp = ref_gp.global_planner.__new__(ref_gp.global_planner)
    g = (rng.random((40, 30)) < 0.15).astype(np.float64)
    segs = rng.integers(0, 30, (200, 4))
    segs[:, 0] = rng.integers(0, 40, 200)
    segs[:, 2] = rng.integers(0, 40, 200)
    import contextlib, io
    los = []
    for a in segs:
        with contextlib.redirect_stdout(io.StringIO()):
            try:
                los.append(int(bool(gp.map_line_col(np.array(a[0:2], dtype=float), np.array(a[2:4], dtype=float), g))))
            except Exception:
                los.append(-1)
    np.savez_compressed(os.path.join(HERE, "hostfn_golden.npz"), rpy=rpy, R=Rs, df_in=pts, df_out=df,
                        los_grid=g, los_segs=segs, los_out=np.array(los))
    print("done")


if __name__ == "__main__":
    main()
