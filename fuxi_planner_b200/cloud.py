"""Cloud-side host glue of the hot path: everything around fx_project that the reference's cloud node does on
the host (scripts/plc_point2_st.py, scripts/utils.py).  Small host math only -- the per-point work runs in
fx_project (transform + height filter + scatter in one kernel).

  cloud_affine      camera->earth rigid transform of plc_point2_st.py:244-251 (+ utils.py:21-28) as ONE 3x4 matrix
  pointcloud2_xyz   PointCloud2 <-> float32 [N,3] (layout of plc_point2_st.py:112-138: x,y,z FLOAT32 at 0/4/8, step 12)
  cloud_to_grid     host cloud -> inflated grid through fx_map_host
"""
import math

import numpy as np

from . import api

CAMERA_LEVER_ARM = 0.12   # plc_point2_st.py:244  x_b = z_c + 0.12


def rotation_zyx(roll, pitch, yaw):
    """Body->earth rotation Rz(yaw) @ Ry(pitch) @ Rx(roll) (the matrix utils.py:21-28 spells out element by element)."""
    cr, sr = math.cos(roll), math.sin(roll)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cy, sy = math.cos(yaw), math.sin(yaw)
    rz = np.array([[cy, -sy, 0.0], [sy, cy, 0.0], [0.0, 0.0, 1.0]])
    ry = np.array([[cp, 0.0, sp], [0.0, 1.0, 0.0], [-sp, 0.0, cp]])
    rx = np.array([[1.0, 0.0, 0.0], [0.0, cr, -sr], [0.0, sr, cr]])
    return rz @ ry @ rx


def cloud_affine(rpy, pos, dt=0.0, ang_vel=(0.0, 0.0, 0.0), line_vel=(0.0, 0.0, 0.0)):
    """3x4 matrix A with  earth = A[:, :3] @ (x_c, y_c, z_c) + A[:, 3]  for camera-frame points.

    Folds, in the reference's order (plc_point2_st.py:244-251): the camera->body axis permutation
    (x_b, y_b, z_b) = (z_c + 0.12, -x_c, -y_c), the attitude extrapolated by dt * ang_vel, the body->earth
    rotation, and the position extrapolated by dt * line_vel."""
    att = np.asarray(rpy, dtype=np.float64) + float(dt) * np.asarray(ang_vel, dtype=np.float64)
    R = rotation_zyx(*att)
    # columns of (R @ P) for the permutation P: x_c feeds -y_b, y_c feeds -z_b, z_c feeds +x_b
    A = np.empty((3, 4), dtype=np.float64)
    A[:, 0] = -R[:, 1]
    A[:, 1] = -R[:, 2]
    A[:, 2] = R[:, 0]
    A[:, 3] = R[:, 0] * CAMERA_LEVER_ARM + np.asarray(pos, dtype=np.float64) + float(dt) * np.asarray(line_vel, dtype=np.float64)
    return A


def xyz_to_pointcloud2_fields(points):
    """float [N,3] -> the PointCloud2 payload the reference publishes (plc_point2_st.py:112-138)."""
    p = np.ascontiguousarray(points, dtype="<f4").reshape(-1, 3)
    return {"height": 1, "width": int(p.shape[0]), "point_step": 12, "row_step": 12 * int(p.shape[0]),
            "is_bigendian": False, "is_dense": int(np.isfinite(p).all()), "data": p.tobytes()}


def pointcloud2_to_xyz(data, point_step=12, offsets=(0, 4, 8)):
    """PointCloud2 byte payload -> float32 [N,3] without the per-point Python tuples of
    ``list(read_points(...))`` (plc_point2_st.py:44).  A 12-byte xyz payload is returned as a zero-copy view
    (the packed layout fx_project reads directly with stride 3)."""
    buf = np.frombuffer(data, dtype=np.uint8)
    n = buf.size // point_step
    if point_step == 12 and tuple(offsets) == (0, 4, 8):
        return buf[: n * 12].view("<f4").reshape(n, 3)
    rec = buf[: n * point_step].reshape(n, point_step)
    return np.stack([rec[:, o:o + 4].copy().view("<f4").reshape(n) for o in offsets], axis=1)


def cloud_to_grid(points, rpy=None, pos=None, origin=(0.0, 0.0), reso=0.2, shape=(1024, 1024), zmin=0.3,
                  zmax=math.inf, radius=0, variant="ccst", dt=0.0, ang_vel=(0.0, 0.0, 0.0), line_vel=(0.0, 0.0, 0.0),
                  ctx=None, device=0):
    """Host cloud (camera frame if rpy/pos are given, else already in the earth frame) -> inflated uint8 grid [W][H]."""
    A = None if rpy is None else cloud_affine(rpy, pos if pos is not None else (0.0, 0.0, 0.0), dt, ang_vel, line_vel)
    return api.map_host(points, A, zmin, zmax, origin, reso, shape, radius, variant, ctx=ctx, device=device)
