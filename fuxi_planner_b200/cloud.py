"""Cloud-side host glue of the hot path: everything around fx_project that the reference's cloud node does on
the host (scripts/plc_point2_st.py, scripts/utils.py).  Small host math only -- the per-point work runs in
fx_project (transform + height filter + scatter in one kernel).

  cloud_affine      camera->earth rigid transform of plc_point2_st.py:244-251 (+ utils.py:21-28) as ONE 3x4 matrix
  pointcloud2_xyz   PointCloud2 <-> float32 [N,3] (layout of plc_point2_st.py:112-138: x,y,z FLOAT32 at 0/4/8, step 12)
  cloud_to_grid     host cloud -> inflated grid through fx_map_host
  cloud_filter(_host)     PassThrough -> VoxelGrid -> RadiusOutlierRemoval of src/chen_filter_rgb.cpp:52-71 (fx_cloud_filter)
  distance_filter(_host)  convert_plc.distance_filter, plc_point2_st.py:139-148 (fx_distance_filter)
"""
import ctypes as C

import math

import numpy as np

import torch

from . import api
from ._lib import CloudParams, FuxiError, default_context

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


def _cloud_params(stride, rgb_offset, pass_lim, leaf, radius, min_neighbors):
    return CloudParams(int(stride), int(rgb_offset), float(pass_lim[0]), float(pass_lim[1]), float(leaf[0]), float(leaf[1]),
                       float(leaf[2]), int(min_neighbors), float(radius))


def cloud_filter(points, rgb_offset=-1, pass_lim=(0.0, 4.0), leaf=(0.17, 0.17, 0.2), radius=0.35, min_neighbors=13, ctx=None,
                 out=None, counts=None):
    """PCL chain of src/chen_filter_rgb.cpp:52-71 on a float32 CUDA tensor [n, stride] (x, y, z first; rgb_offset = column
    of PCL's packed rgb word or -1).  Returns (out float32 [n, 4] = x, y, z, rgb word; counts int64 CUDA tensor [4] =
    {after PassThrough, voxels, kept, status}); rows [0, counts[2]) of out are valid.  Enqueues only (no host sync).
    `out` / `counts`: reuse these tensors (a node that filters every frame into the same buffers replays one CUDA graph)."""
    if not (isinstance(points, torch.Tensor) and points.is_cuda and points.dtype == torch.float32 and points.dim() == 2
            and points.shape[1] >= 3):
        raise FuxiError("points must be a float32 CUDA tensor [n, stride >= 3]")
    points = points.contiguous()
    ctx = api._ctx(ctx, points)
    n, stride = points.shape
    if out is None:
        out = torch.empty((max(n, 1), 4), dtype=torch.float32, device=points.device)
    if counts is None:
        counts = torch.empty(4, dtype=torch.int64, device=points.device)
    if not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.dim() == 2 and out.shape[1] == 4
            and counts.is_cuda and counts.dtype == torch.int64 and counts.numel() >= 4 and counts.is_contiguous()):
        raise FuxiError("out must be a contiguous float32 CUDA tensor [cap, 4], counts an int64 CUDA tensor [4]")
    prm = _cloud_params(stride, rgb_offset, pass_lim, leaf, radius, min_neighbors)
    ctx.check(ctx.lib.fx_cloud_filter(ctx.handle, api._ptr(points), n, C.byref(prm), api._ptr(out), out.shape[0],
                                      api._ptr(counts), api._stream()), "fx_cloud_filter")
    return out, counts


def cloud_filter_host(points, rgb_offset=-1, pass_lim=(0.0, 4.0), leaf=(0.17, 0.17, 0.2), radius=0.35, min_neighbors=13,
                      ctx=None, device=0):
    """Host numpy float32 [n, stride] -> (float32 [kept, 4], counts int64[4]) through fx_cloud_filter_host."""
    p = np.ascontiguousarray(points, dtype=np.float32)
    if p.ndim != 2 or p.shape[1] < 3:
        raise FuxiError("points must be [n, stride >= 3]")
    ctx = ctx or default_context(device)
    n, stride = p.shape
    out = np.empty((max(n, 1), 4), dtype=np.float32)
    counts = np.zeros(4, dtype=np.int64)
    prm = _cloud_params(stride, rgb_offset, pass_lim, leaf, radius, min_neighbors)
    ctx.check(ctx.lib.fx_cloud_filter_host(ctx.handle, p.ctypes.data_as(C.c_void_p), n, C.byref(prm),
                                           out.ctypes.data_as(C.c_void_p), out.shape[0],
                                           counts.ctypes.data_as(C.POINTER(C.c_int64))), "fx_cloud_filter_host")
    return out[:counts[2]].copy(), counts


def distance_filter(points, dis, ctx=None, out=None, count=None):
    """convert_plc.distance_filter (plc_point2_st.py:139-148) on a float64 CUDA tensor [n, 3]: returns (out float64
    [n, 3], count int32 CUDA tensor [1]); rows [0, count) are the points with |p| < dis ordered by (|p|, z, y, x)."""
    if not (isinstance(points, torch.Tensor) and points.is_cuda and points.dtype == torch.float64 and points.dim() == 2
            and points.shape[1] == 3):
        raise FuxiError("points must be a float64 CUDA tensor [n, 3]")
    points = points.contiguous()
    ctx = api._ctx(ctx, points)
    if out is None:
        out = torch.empty_like(points)
    if count is None:
        count = torch.empty(1, dtype=torch.int32, device=points.device)
    if not (out.is_cuda and out.dtype == torch.float64 and out.is_contiguous() and out.shape == points.shape
            and count.is_cuda and count.dtype == torch.int32 and count.numel() >= 1):
        raise FuxiError("out must be a contiguous float64 CUDA tensor [n, 3], count an int32 CUDA tensor [1]")
    ctx.check(ctx.lib.fx_distance_filter(ctx.handle, api._ptr(points), points.shape[0], float(dis), api._ptr(out),
                                         api._ptr(count), api._stream()), "fx_distance_filter")
    return out, count


def distance_filter_host(plc, dis, ctx=None, device=0):
    """Drop-in for ``convert_plc.distance_filter(plc, dis)``: host array in, float64 [m, 3] out."""
    p = np.ascontiguousarray(np.asarray(plc, dtype=np.float64).reshape(-1, 3))
    ctx = ctx or default_context(device)
    out = np.empty_like(p)
    cnt = C.c_int64(0)
    ctx.check(ctx.lib.fx_distance_filter_host(ctx.handle, p.ctypes.data_as(C.c_void_p), p.shape[0], float(dis),
                                              out.ctypes.data_as(C.c_void_p), C.byref(cnt)), "fx_distance_filter_host")
    return out[:cnt.value].copy()


def rotation_elementwise(roll, pitch, yaw):
    """The body->earth matrix with the products grouped the way utils.body_to_earth_frame evaluates them
    (scripts/utils.py:21-28: each entry left to right, e.g. (cos y * sin p) * sin r - sin y * cos r), so that the float64
    entries -- and with them every transformed coordinate -- carry the same bits as the reference's.  rotation_zyx above
    multiplies three matrices instead and differs in the last place."""
    sr, cr = math.sin(roll), math.cos(roll)
    sp, cp = math.sin(pitch), math.cos(pitch)
    sy, cy = math.sin(yaw), math.cos(yaw)
    m = np.empty((3, 3), dtype=np.float64)
    m[0] = (cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr)
    m[1] = (sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr)
    m[2] = (-sp, cp * sr, cp * cr)
    return m


def node_cloud_host(points, rpy, pos, dt=0.0, ang_vel=(0.0, 0.0, 0.0), line_vel=(0.0, 0.0, 0.0), local_pos=None, zmin=0.3, dis=4.0,
                    ctx=None, device=0):
    """The cloud `plc_point2_st.py` publishes on /points_global_all (lines 243-256 + 336-339): camera-frame points (host
    array [n, >= 3], float32 PointCloud2 data or float64) -> transformed to the earth frame with the time-compensated
    attitude / position, height filter, range filter around `local_pos` (default: pos) sorted by distance.  float64 [m, 3].
    One call of fx_transform_filter_host; the [n_dyn, 0, 0] row the node appends (:341) is the caller's."""
    p = np.asarray(points)
    is64 = p.dtype == np.float64
    p = np.ascontiguousarray(p, dtype=np.float64 if is64 else np.float32)
    if p.ndim != 2 or p.shape[1] < 3:
        raise FuxiError("points must be [n, >= 3]")
    r, pp, y = np.array(rpy, dtype=np.float64) + float(dt) * np.array(ang_vel, dtype=np.float64)        # :248
    R = np.ascontiguousarray(rotation_elementwise(r, pp, y))
    t = np.ascontiguousarray(float(dt) * np.array(line_vel, dtype=np.float64) + np.array(pos, dtype=np.float64))   # :250
    c = np.ascontiguousarray(np.array(pos if local_pos is None else local_pos, dtype=np.float64))
    return _tf_host(p, is64, R, t, c, zmin, 0.0, dis, ctx, device)


def octomap_local_host(centres, local_pos, box=4.0, dis=4.0, ctx=None, device=0):
    """/octomap_point_cloud_centers_local (plc_point2_st.py:351-362): voxel centres within `box` of the vehicle on every
    axis and within `dis` of it, sorted by distance.  float64 [m, 3]."""
    p = np.asarray(centres)
    is64 = p.dtype == np.float64
    p = np.ascontiguousarray(p, dtype=np.float64 if is64 else np.float32)
    c = np.ascontiguousarray(np.array(local_pos, dtype=np.float64))
    return _tf_host(p, is64, None, None, c, -math.inf, box, dis, ctx, device)


def _tf_host(p, is64, R, t, c, zmin, box, dis, ctx, device):
    ctx = ctx or default_context(device)
    n, stride = p.shape
    out = np.empty((max(n, 1), 3), dtype=np.float64)
    cnt = C.c_int64(0)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None
    ctx.check(ctx.lib.fx_transform_filter_host(ctx.handle, p.ctypes.data_as(C.c_void_p), n, stride, int(is64), dp(R), dp(t), dp(c),
                                               float(zmin), float(box), float(dis), out.ctypes.data_as(C.c_void_p), C.byref(cnt)),
              "fx_transform_filter_host")
    return out[:cnt.value].copy()
