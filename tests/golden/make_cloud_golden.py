#!/usr/bin/env python
"""Golden vectors for the cloud node's published product (SURVEY rows a15 -> a16 -> a17 as one pipeline), produced by
executing the UNMODIFIED reference source lines of scripts/plc_point2_st.py: 244-251 (axis swap, time compensation,
rotation with utils.body_to_earth_frame, translation), 255-256 (height filter), 336-339 (range filter around the vehicle
through convert_plc.distance_filter) and the octomap-centres branch 360-362.  The lines sit in the `__main__` loop, so
they are read from the file, dedented and exec'd in a namespace holding the loop's variables; `convert` is an instance
of the reference's own convert_plc class (ROS modules stubbed), so distance_filter is the reference's method.

Run in the build container only (needs /root/reference):   python tests/golden/make_cloud_golden.py
Writes cloud_golden.npz next to this file.
"""
import os
import sys
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import refload  # noqa: E402
import make_golden  # noqa: E402  (stub_ros)


def block(lines, lo, hi):
    body = [ln for ln in lines[lo - 1:hi] if ln.strip() and not ln.strip().startswith("#")]
    return compile(textwrap.dedent("\n".join(body)) + "\n", "plc_point2_st.py:%d-%d" % (lo, hi), "exec")


def main():
    assert refload.available()
    make_golden.stub_ros()
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(refload.REF_ROOT, "scripts"))
    import plc_point2_st as ref
    with open(os.path.join(refload.REF_ROOT, "scripts", "plc_point2_st.py"), encoding="utf-8", errors="replace") as fh:
        lines = fh.read().split("\n")
    assert lines[243].strip().startswith("plc_c[:,0]=plc[:,2]+0.12") and lines[250].strip().startswith("plc_c=np.matmul(b2e,plc_c.T).T")
    assert lines[255].strip() == "plc1=plc_c[plc_c[:,2]>0.3]" and lines[336].strip() == "plc1=plc1-local_pos"
    assert lines[360].strip() == "octo_plc=octo_plc[(abs(octo_plc)<4).all(axis=1)]"
    transform = block(lines, 244, 251)        # the body of the `if` at :243
    hfilter = block(lines, 256, 256)
    rfilter = block(lines, 337, 339)          # the body of `if len(plc1)>0:` up to `plc1=plc1+local_pos`
    octo = block(lines, 360, 362)
    rng = np.random.default_rng(21)
    out = {}
    ncase = 12
    for i in range(ncase):
        n = int(rng.integers(50, 4000))
        cam = np.c_[rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(0.2, 7.0, n)].astype(np.float32)   # PointCloud2 data
        plc = np.array([tuple(float(v) for v in row) for row in cam])       # what list(read_points(...)) -> np.array gives
        conv = ref.convert_plc.__new__(ref.convert_plc)
        conv.plc_time, conv.pos_time = 10.0 + float(rng.uniform(0, 0.05)), 10.0
        conv.ang_vel = rng.uniform(-0.5, 0.5, 3)
        conv.line_vel = rng.uniform(-2, 2, 3)
        rpy = rng.uniform(-0.4, 0.4, 3); rpy[2] = rng.uniform(-3.1, 3.1)
        pos = rng.uniform(-10, 10, 3); pos[2] = rng.uniform(0.5, 3.0)
        local_pos = pos + (rng.uniform(-0.05, 0.05, 3) if i % 3 == 0 else 0.0)
        ns = {"np": np, "plc": plc, "plc_c": np.zeros([n, 3]), "length": n, "convert": conv, "body_to_earth_frame": ref.body_to_earth_frame,
              "r": float(rpy[0]), "p": float(rpy[1]), "y": float(rpy[2]), "px": float(pos[0]), "py": float(pos[1]), "pz": float(pos[2]),
              "local_pos": local_pos.copy()}
        exec(transform, ns)
        exec(hfilter, ns)
        exec(rfilter, ns)
        out["cam_%d" % i] = cam
        out["par_%d" % i] = np.concatenate([rpy, pos, local_pos, conv.ang_vel, conv.line_vel, [conv.plc_time - conv.pos_time]])
        out["out_%d" % i] = np.asarray(ns["plc1"], dtype=np.float64)
        # octomap centres around the vehicle
        cen = (np.round(rng.uniform(-8, 8, (int(rng.integers(20, 3000)), 3)) / 0.2) * 0.2 + local_pos.round(1)).astype(np.float32)
        octo_plc = np.array([tuple(float(v) for v in row) for row in cen])
        ns2 = {"np": np, "octo_plc": octo_plc, "local_pos": local_pos.copy(), "convert": conv}
        exec(octo, ns2)
        out["cen_%d" % i] = cen
        out["octo_%d" % i] = np.asarray(ns2["octo_plc"], dtype=np.float64)
        print(i, n, len(out["out_%d" % i]), len(cen), len(out["octo_%d" % i]), flush=True)
    out["ncase"] = np.array([ncase])
    np.savez_compressed(os.path.join(HERE, "cloud_golden.npz"), **out)


if __name__ == "__main__":
    main()
