"""Python surface over the C ABI.  torch is used only for device buffers and streams.

Grid convention everywhere: uint8 ``[W][H]`` C-contiguous, ``grid[x, y]`` == the reference's
``matrix[x][y]`` (scripts/global_planner_st.py:16-18); for the search 1 is an obstacle, anything
else is free (scripts/jps1.py:20-29).
"""
import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import FuxiError, default_context


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _ctx(ctx, tensor=None):
    if ctx is not None:
        return ctx
    dev = tensor.device.index if tensor is not None and tensor.is_cuda else torch.cuda.current_device()
    return default_context(dev if dev is not None else 0)


def _u8_grid(grid):
    if not (isinstance(grid, torch.Tensor) and grid.is_cuda):
        raise FuxiError("grid must be a CUDA tensor (use plan_host / map_host for host buffers)")
    if grid.dim() != 2:
        raise FuxiError("grid must be 2-D [W][H]")
    if grid.dtype != torch.uint8:
        raise FuxiError("grid must be uint8")
    return grid.contiguous()


def project(points, affine=None, zmin=0.3, zmax=math.inf, origin=(0.0, 0.0), reso=0.2, shape=None, out=None,
            clear=True, ctx=None):
    """Point cloud -> occupancy grid (fx_project).  ``points``: float32 CUDA tensor [N,3] or [N,4];
    ``affine``: 3x4 (camera->earth incl. axis swap, see cloud.cloud_affine) or None for identity."""
    if not (isinstance(points, torch.Tensor) and points.is_cuda and points.dtype == torch.float32 and points.dim() == 2
            and points.shape[1] in (3, 4)):
        raise FuxiError("points must be a float32 CUDA tensor of shape [N,3] or [N,4]")
    points = points.contiguous()
    ctx = _ctx(ctx, points)
    if out is None:
        if shape is None:
            raise FuxiError("shape=(W,H) or out= is required")
        out = torch.empty((int(shape[0]), int(shape[1])), dtype=torch.uint8, device=points.device)
    W, H = out.shape
    A = np.eye(3, 4, dtype=np.float32) if affine is None else np.ascontiguousarray(np.asarray(affine, dtype=np.float32).reshape(3, 4))
    rc = ctx.lib.fx_project(ctx.handle, _ptr(points), points.shape[0], points.shape[1],
                            A.ctypes.data_as(C.POINTER(C.c_float)), float(zmin), float(zmax), float(origin[0]),
                            float(origin[1]), float(reso), W, H, _ptr(out), 1 if clear else 0, _stream())
    ctx.check(rc, "fx_project")
    return out


def inflate(grid, radius, variant="ccst", out=None, ctx=None):
    """Obstacle inflation (fx_inflate).  variant "ccst": dense (2r+1)^2 square (global_planner_ccst.py:442-448);
    "st": 9-point stencil {-r,0,+r}^2 (global_planner_st.py:256-262; identical to dense for r == 1)."""
    grid = _u8_grid(grid)
    ctx = _ctx(ctx, grid)
    if out is None:
        out = torch.empty_like(grid)
    step = 1 if variant == "ccst" else max(int(radius), 1)
    W, H = grid.shape
    ctx.check(ctx.lib.fx_inflate(ctx.handle, _ptr(grid), _ptr(out), W, H, int(radius), step, _stream()), "fx_inflate")
    return out


def edt(grid, out=None, ctx=None):
    """Exact squared Euclidean distance (cells^2) to the nearest cell > 0 (fx_edt)."""
    grid = _u8_grid(grid)
    ctx = _ctx(ctx, grid)
    if out is None:
        out = torch.empty(grid.shape, dtype=torch.int32, device=grid.device)
    W, H = grid.shape
    ctx.check(ctx.lib.fx_edt(ctx.handle, _ptr(grid), _ptr(out), W, H, _stream()), "fx_edt")
    return out


@dataclass
class PlanResult:
    cost_i: torch.Tensor    # int32 [Q]   metric 1: 10/14 cost; metric 2: units of 1/FX_EUCLID_WS cell; <0: FX_COST_*
    cost_f: torch.Tensor    # float64 [Q] metric 2: straight + diagonal*sqrt(2); metric 1: float(cost_i)
    path_xy: torch.Tensor   # int32 [Q, max_path, 2] turning points (start .. goal) or None
    path_len: torch.Tensor  # int32 [Q]   number of turning points (<0: FX_COST_*)

    def path(self, q):
        """Query q as the reference returns it: list of (x, y) tuples, or the int 0 when there is no path."""
        n = int(self.path_len[q])
        if n <= 0:
            return 0
        if n > self.path_xy.shape[1]:
            raise FuxiError("path of query %d has %d turning points > max_path=%d" % (q, n, self.path_xy.shape[1]))
        return [tuple(int(v) for v in p) for p in self.path_xy[q, :n].tolist()]


def plan_batch(grid, starts, goals, metric=2, max_path=512, ctx=None):
    """Q shortest-path queries in one launch (fx_search_batch).  starts/goals: int32 CUDA tensors [Q,2]."""
    grid = _u8_grid(grid)
    ctx = _ctx(ctx, grid)
    starts = starts.to(device=grid.device, dtype=torch.int32).contiguous()
    goals = goals.to(device=grid.device, dtype=torch.int32).contiguous()
    if starts.shape != goals.shape or starts.dim() != 2 or starts.shape[1] != 2:
        raise FuxiError("starts/goals must both be [Q,2]")
    Q = starts.shape[0]
    W, H = grid.shape
    cost_i = torch.empty(Q, dtype=torch.int32, device=grid.device)
    cost_f = torch.empty(Q, dtype=torch.float64, device=grid.device)
    path_len = torch.empty(Q, dtype=torch.int32, device=grid.device)
    path_xy = torch.empty((Q, max_path, 2), dtype=torch.int32, device=grid.device) if max_path > 0 else None
    rc = ctx.lib.fx_search_batch(ctx.handle, _ptr(grid), W, H, _ptr(starts), _ptr(goals), Q, int(metric), _ptr(cost_i),
                                 _ptr(cost_f), _ptr(path_xy), _ptr(path_len), int(max_path), _stream())
    ctx.check(rc, "fx_search_batch")
    return PlanResult(cost_i, cost_f, path_xy, path_len)


def search_stats(ctx=None):
    """(settled cells, wavefront levels, passes, band-only answers) of the last batch (synchronises)."""
    ctx = _ctx(ctx)
    a = (C.c_int64 * 4)()
    ctx.check(ctx.lib.fx_search_stats(ctx.handle, a), "fx_search_stats")
    return tuple(int(v) for v in a)


def search_kernel_ms(ctx=None):
    """Duration (ms) of the last k_search_batch launch alone, from CUDA events on its stream (fx_search_kernel_ms)."""
    ctx = _ctx(ctx)
    ms = C.c_float(0.0)
    ctx.check(ctx.lib.fx_search_kernel_ms(ctx.handle, C.byref(ms)), "fx_search_kernel_ms")
    return float(ms.value)


def search_timings(ctx=None):
    """(k_band_bound ms, k_search_batch ms) of the last batch (fx_search_timings)."""
    ctx = _ctx(ctx)
    ms = (C.c_float * 2)()
    ctx.check(ctx.lib.fx_search_timings(ctx.handle, ms), "fx_search_timings")
    return float(ms[0]), float(ms[1])


def plan_host_stages(ctx=None):
    """Stage times (us) of the last traced plan_host call (FUXI_B200_TRACE=1|2 in the environment before the library is
    first used): host fill + upload issue, host enqueue, host wait, device upload, device search, device paths + D2H."""
    ctx = _ctx(ctx)
    us = (C.c_double * 6)()
    ctx.check(ctx.lib.fx_plan_host_stages(ctx.handle, us), "fx_plan_host_stages")
    return tuple(float(v) for v in us)


def field(grid, source, metric=1, out=None, ctx=None, check=True):
    """Cost-from-source field, int32 [W][H], -1 = unreachable (fx_field)."""
    grid = _u8_grid(grid)
    ctx = _ctx(ctx, grid)
    if out is None:
        out = torch.empty(grid.shape, dtype=torch.int32, device=grid.device)
    W, H = grid.shape
    ctx.check(ctx.lib.fx_field(ctx.handle, _ptr(grid), W, H, int(source[0]), int(source[1]), int(metric), _ptr(out),
                               _stream()), "fx_field")
    if check:
        ctx.check(ctx.lib.fx_field_status(ctx.handle, None, None), "fx_field")
    return out


def field_relax(grid, fld, metric=1, changed=None, ctx=None, check=True):
    """Relax an int32 field holding seeds to its fixpoint in place (fx_field_relax); returns the `changed` tensor."""
    grid = _u8_grid(grid)
    ctx = _ctx(ctx, grid)
    if not (fld.is_cuda and fld.dtype == torch.int32 and fld.is_contiguous() and fld.shape == grid.shape):
        raise FuxiError("field must be a contiguous int32 CUDA tensor of the grid's shape")
    if changed is None:
        changed = torch.zeros(1, dtype=torch.int32, device=grid.device)
    W, H = grid.shape
    ctx.check(ctx.lib.fx_field_relax(ctx.handle, _ptr(grid), W, H, int(metric), _ptr(fld), _ptr(changed), _stream()),
              "fx_field_relax")
    if check:
        ctx.check(ctx.lib.fx_field_status(ctx.handle, None, None), "fx_field_relax")
    return changed


def field_status(ctx=None):
    ctx = _ctx(ctx)
    lv, st = C.c_int64(0), C.c_int64(0)
    ctx.check(ctx.lib.fx_field_status(ctx.handle, C.byref(lv), C.byref(st)), "fx_field_status")
    return lv.value, st.value


# ---------------------------------------------------------------------------------- grid assembly / path post-processing
def grid_decode(msg, width, height, out=None, window=None, paste_at=(0, 0), ctx=None):
    """OccupancyGrid.data (int8 CUDA tensor, index y*width + x) -> uint8 [x][y] with 100 -> 1, -1 -> 0
    (map_callback, global_planner_st.py:15-20), optionally only ``window = (x0, y0, w, h)`` of it, pasted at
    ``paste_at`` of ``out`` (which must then be pre-zeroed by the caller where the window does not reach)."""
    if not (isinstance(msg, torch.Tensor) and msg.is_cuda and msg.dtype == torch.int8 and msg.numel() == width * height):
        raise FuxiError("msg must be an int8 CUDA tensor of width*height elements")
    msg = msg.contiguous()
    ctx = _ctx(ctx, msg)
    x0, y0, w, h = window if window is not None else (0, 0, int(width), int(height))
    if out is None:
        out = torch.zeros((w + paste_at[0], h + paste_at[1]), dtype=torch.uint8, device=msg.device)
    ctx.check(ctx.lib.fx_grid_decode(ctx.handle, _ptr(msg), int(width), int(height), int(x0), int(y0), int(w), int(h), _ptr(out),
                                     out.shape[0], out.shape[1], int(paste_at[0]), int(paste_at[1]), _stream()), "fx_grid_decode")
    return out


def grid_encode(grid, ctx=None):
    """uint8 [x][y] -> OccupancyGrid.data int8 (publish_map, global_planner_st.py:102-115): 1 -> 100, y-major."""
    grid = _u8_grid(grid)
    ctx = _ctx(ctx, grid)
    W, H = grid.shape
    out = torch.empty(W * H, dtype=torch.int8, device=grid.device)
    ctx.check(ctx.lib.fx_grid_encode(ctx.handle, _ptr(grid), W, H, _ptr(out), _stream()), "fx_grid_encode")
    return out


def grid_to_image(grid, ctx=None):
    """uint8 [x][y] -> 'L' image [H rows][W columns] as the planners save it (global_planner_st.py:368-372)."""
    grid = _u8_grid(grid)
    ctx = _ctx(ctx, grid)
    W, H = grid.shape
    img = torch.empty((H, W), dtype=torch.uint8, device=grid.device)
    ctx.check(ctx.lib.fx_grid_to_image(ctx.handle, _ptr(grid), W, H, _ptr(img), _stream()), "fx_grid_to_image")
    return img


def image_to_grid(img, threshold=200, ctx=None):
    """'L' image [H][W] uint8 CUDA tensor -> uint8 [x][y], pixel > threshold = free (pre-map loader, st:176-182)."""
    img = _u8_grid(img)
    ctx = _ctx(ctx, img)
    H, W = img.shape
    grid = torch.empty((W, H), dtype=torch.uint8, device=img.device)
    ctx.check(ctx.lib.fx_image_to_grid(ctx.handle, _ptr(img), W, H, int(threshold), _ptr(grid), _stream()), "fx_image_to_grid")
    return grid


def grid_paste(src, dst, window=None, paste_at=(0, 0), ctx=None):
    """dst[px+i, py+j] = src[x0+i, y0+j] (overwrite), the slice assignments of the pre-map merge / pad step."""
    src, ctx = _u8_grid(src), _ctx(ctx, src)
    if not (dst.is_cuda and dst.dtype == torch.uint8 and dst.dim() == 2 and dst.is_contiguous()):
        raise FuxiError("dst must be a contiguous uint8 CUDA tensor [W][H]")
    x0, y0, w, h = window if window is not None else (0, 0, src.shape[0], src.shape[1])
    ctx.check(ctx.lib.fx_grid_paste(ctx.handle, _ptr(src), src.shape[0], src.shape[1], int(x0), int(y0), int(w), int(h), _ptr(dst),
                                    dst.shape[0], dst.shape[1], int(paste_at[0]), int(paste_at[1]), _stream()), "fx_grid_paste")
    return dst


def grid_bbox(a, width=None, height=None, ctx=None):
    """int32 CUDA tensor {min x, max x, min y, max y} of the non-zero cells: of a uint8 [x][y] array, or (width/height
    given) of a raw int8 OccupancyGrid message (non-zero after the 100/-1 mapping)."""
    ctx = _ctx(ctx, a)
    out = torch.empty(4, dtype=torch.int32, device=a.device)
    if width is None:
        a = _u8_grid(a)
        ctx.check(ctx.lib.fx_grid_bbox(ctx.handle, _ptr(a), a.shape[0], a.shape[1], 0, _ptr(out), _stream()), "fx_grid_bbox")
    else:
        a = a.contiguous()
        ctx.check(ctx.lib.fx_grid_bbox(ctx.handle, _ptr(a), int(width), int(height), 1, _ptr(out), _stream()), "fx_grid_bbox")
    return out


def relocate_goal(grid, goal, ifa=1, variant="st", ctx=None):
    """Goal on an obstacle -> nearest free cell of its row, else column (global_planner_st.py:268-275).
    Returns an int32 CUDA tensor {gx, gy, moved, end_occu}."""
    grid = _u8_grid(grid)
    ctx = _ctx(ctx, grid)
    g = torch.tensor([int(goal[0]), int(goal[1]), 0, 0], dtype=torch.int32, device=grid.device)
    ctx.check(ctx.lib.fx_relocate_goal(ctx.handle, _ptr(grid), grid.shape[0], grid.shape[1], _ptr(g), int(ifa),
                                       1 if variant == "ccst" else 0, _stream()), "fx_relocate_goal")
    return g


def path_post(grid, path_xy, path_len, shortcut=True, drop=None, world=None, ctx=None):
    """Path post-processing of the ccmapping planner on Q paths at once (fx_path_post): optional near-vehicle drop
    ``drop = (px, py, pz, radius)`` (ccst:507-513), line-of-sight shortcutting (ccst:258-283, 515-521), world
    coordinates ``world = (reso, origin_x, origin_y, off_x, off_y)``.  Returns (out_xy, out_len, out_world | None)."""
    grid = _u8_grid(grid)
    ctx = _ctx(ctx, grid)
    path_xy = path_xy.to(torch.int32).contiguous()
    path_len = path_len.to(torch.int32).contiguous()
    Q, max_path = path_xy.shape[0], path_xy.shape[1]
    out_xy = torch.empty_like(path_xy)
    out_len = torch.empty_like(path_len)
    out_world = torch.empty((Q, max_path, 3), dtype=torch.float64, device=grid.device) if world is not None else None
    d4 = (C.c_double * 4)(*[float(v) for v in drop]) if drop is not None else None
    w5 = (C.c_double * 5)(*[float(v) for v in world]) if world is not None else None
    if drop is not None and world is None:
        raise FuxiError("drop= needs world= (the drop radius is in world units)")
    ctx.check(ctx.lib.fx_path_post(ctx.handle, _ptr(grid), grid.shape[0], grid.shape[1], _ptr(path_xy), _ptr(path_len), Q, max_path,
                                   1 if shortcut else 0, d4, w5, _ptr(out_xy), _ptr(out_len), _ptr(out_world), _stream()),
              "fx_path_post")
    return out_xy, out_len, out_world


# ---------------------------------------------------------------------------------- host-buffer calls
def replan_host(map_data, width, height, origin, reso, start_xy, goal_xy, ifa=1, variant="st", hchoice=2, layout="msg",
                crop=False, shortcut=False, drop=None, max_path=1024, want_grid=False, ctx=None, device=0):
    """One global replan through fx_replan_host: numpy in, numpy out, one H2D + one D2H + one sync.
    ``map_data``: OccupancyGrid.data (int8, layout "msg") or a uint8 [x][y] array (layout "array", width = W, height = H).
    Returns (ReplanOut struct, path cells int32 [k,2], path world float64 [k,3], grid uint8 [W][H] | None)."""
    ctx = ctx or default_context(device)
    if layout == "msg":
        m = np.ascontiguousarray(map_data, dtype=np.int8).reshape(-1)
    else:
        m = np.ascontiguousarray(map_data, dtype=np.uint8).reshape(-1)
    if m.size != int(width) * int(height):
        raise FuxiError("map has %d cells, expected %d x %d" % (m.size, width, height))
    rin = _lib.ReplanIn(1 if variant == "ccst" else 0, 0 if layout == "msg" else 1, 1 if crop else 0, int(ifa), int(hchoice),
                        1 if shortcut else 0, float(origin[0]), float(origin[1]), float(reso), float(start_xy[0]),
                        float(start_xy[1]), float(goal_xy[0]), float(goal_xy[1]), *(tuple(float(v) for v in drop) if drop else (0.0,) * 4))
    rout = _lib.ReplanOut()
    pxy = np.empty((max_path, 2), dtype=np.int32)
    pw = np.empty((max_path, 3), dtype=np.float64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    ctx.check(ctx.lib.fx_replan_host(ctx.handle, vp(m), int(width), int(height), C.byref(rin), C.byref(rout), vp(pxy), vp(pw),
                                     int(max_path)), "fx_replan_host")
    n = max(int(rout.path_len), 0)
    grid = None
    if want_grid and rout.skipped != 2:
        grid = np.empty((rout.W, rout.H), dtype=np.uint8)
        ctx.check(ctx.lib.fx_replan_grid_host(ctx.handle, vp(grid), grid.size, None, None), "fx_replan_grid_host")
    return rout, pxy[:n].copy(), pw[:n].copy(), grid



def paths_compact(path_xy, path_len, ctx=None):
    """Padded path rows -> (offsets int64 [Q+1], xy int32 [total, 2]) on the device (fx_paths_compact)."""
    ctx = _ctx(ctx, path_xy)
    Q, max_path = int(path_xy.shape[0]), int(path_xy.shape[1])
    offsets = torch.empty(Q + 1, dtype=torch.int64, device=path_xy.device)
    n = path_len.clamp(min=0)
    cap = int(torch.where(n <= max_path, n, torch.zeros_like(n)).sum().item())
    out = torch.empty((max(cap, 1), 2), dtype=torch.int32, device=path_xy.device)
    ctx.check(ctx.lib.fx_paths_compact(ctx.handle, _ptr(path_xy), _ptr(path_len), Q, max_path, _ptr(offsets), _ptr(out), cap,
                                       _stream()), "fx_paths_compact")
    return offsets, out[:cap]


def paths_jump_points(grid, path_xy, path_len, max_out=None, ctx=None):
    """Jump-point form of the padded path rows (fx_paths_jump_points): returns (out_xy int32 [Q, max_out, 2], out_len int32 [Q])."""
    grid = _u8_grid(grid)
    ctx = _ctx(ctx, grid)
    Q, max_path = int(path_xy.shape[0]), int(path_xy.shape[1])
    W, H = grid.shape
    max_out = int(max_out) if max_out is not None else min(4 * max_path, W + H + 2)
    out = torch.empty((Q, max_out, 2), dtype=torch.int32, device=grid.device)
    out_len = torch.empty(Q, dtype=torch.int32, device=grid.device)
    ctx.check(ctx.lib.fx_paths_jump_points(ctx.handle, _ptr(grid), W, H, _ptr(path_xy.contiguous()), _ptr(path_len.contiguous()), Q, max_path,
                                           _ptr(out), _ptr(out_len), max_out, _stream()), "fx_paths_jump_points")
    return out, out_len


def jump_points_host(grid, path, ctx=None, device=0):
    """One path (sequence of (x, y) turning points, start first) -> the list of (x, y) jump points the reference would
    return for the same cell path (fx_jump_points_host).  grid: host array, obstacle iff == 1."""
    ctx = ctx or default_context(device)
    g = np.asarray(grid)
    g = np.ascontiguousarray(g) if g.dtype == np.uint8 else np.ascontiguousarray((g == 1).astype(np.uint8))
    p = np.ascontiguousarray(np.asarray(path, dtype=np.int32).reshape(-1, 2))
    W, H = g.shape
    cap = max(16, 4 * len(p))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    while True:
        out = np.empty((cap, 2), dtype=np.int32)
        n = C.c_int32(0)
        ctx.check(ctx.lib.fx_jump_points_host(ctx.handle, vp(g), W, H, vp(p), len(p), vp(out), cap, C.byref(n)), "fx_jump_points_host")
        if n.value <= cap:
            return [(int(a), int(b)) for a, b in out[:n.value]]
        cap = n.value


def plan_host_csr(grid, starts, goals, metric=2, max_path=512, cap=None, ctx=None, device=0):
    """plan_host with the paths in compact form (fx_plan_host_csr): returns (cost_i, cost_f, path_len, offsets int64 [Q+1],
    xy int32 [total, 2]).  cap = room for the points (default 64 per query; retried once with the exact total if it is short)."""
    ctx = ctx or default_context(device)
    grid = np.asarray(grid)
    g = np.ascontiguousarray(grid) if grid.dtype == np.uint8 else np.ascontiguousarray((grid == 1).astype(np.uint8))
    s = np.ascontiguousarray(starts, dtype=np.int32).reshape(-1, 2)
    t = np.ascontiguousarray(goals, dtype=np.int32).reshape(-1, 2)
    Q = len(s)
    W, H = g.shape
    cost_i = np.empty(Q, dtype=np.int32)
    cost_f = np.empty(Q, dtype=np.float64)
    path_len = np.empty(Q, dtype=np.int32)
    offsets = np.zeros(Q + 1, dtype=np.int64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    cap = int(cap) if cap is not None else 64 * max(Q, 1)
    while True:
        xy = np.empty((max(cap, 1), 2), dtype=np.int32)
        total = C.c_int64(0)
        rc = ctx.lib.fx_plan_host_csr(ctx.handle, vp(g), W, H, vp(s), vp(t), Q, int(metric), vp(cost_i), vp(cost_f), vp(path_len),
                                      int(max_path), vp(offsets), vp(xy), cap, C.byref(total))
        ctx.check(rc, "fx_plan_host_csr")
        if total.value <= cap:
            return cost_i, cost_f, path_len, offsets, xy[:total.value]
        cap = int(total.value)


def plan_host(grid, starts, goals, metric=2, max_path=512, ctx=None, device=0):
    """numpy in / numpy out through fx_plan_host (H2D + search + D2H inside the call).  A float64 grid is taken as the
    reference's matrix (obstacle iff == 1.0, fx_plan_host_f64); uint8 is passed as is; any other dtype becomes `grid == 1`.
    Returns (cost_i int32[Q], cost_f float64[Q], path_xy int32[Q,max_path,2] or None, path_len int32[Q])."""
    ctx = ctx or default_context(device)
    grid = np.asarray(grid)
    f64 = grid.dtype == np.float64          # the reference's own matrix type: `== 1` is evaluated inside the library
    if f64 or grid.dtype == np.uint8:
        g = np.ascontiguousarray(grid)
    else:
        # the reference's test is `matrix[x][y] == 1` (jps1.py:20-29): 257 or 1.5 are FREE cells; a plain astype(uint8)
        # would wrap / truncate them to 1
        g = np.ascontiguousarray((grid == 1).astype(np.uint8))
    s = np.ascontiguousarray(starts, dtype=np.int32).reshape(-1, 2)
    t = np.ascontiguousarray(goals, dtype=np.int32).reshape(-1, 2)
    Q = len(s)
    W, H = g.shape
    cost_i = np.empty(Q, dtype=np.int32)
    cost_f = np.empty(Q, dtype=np.float64)
    path_len = np.empty(Q, dtype=np.int32)
    path_xy = np.empty((Q, max_path, 2), dtype=np.int32) if max_path > 0 else None
    vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    fn = ctx.lib.fx_plan_host_f64 if f64 else ctx.lib.fx_plan_host
    rc = fn(ctx.handle, vp(g), W, H, vp(s), vp(t), Q, int(metric), vp(cost_i), vp(cost_f), vp(path_xy), vp(path_len), int(max_path))
    ctx.check(rc, "fx_plan_host")
    return cost_i, cost_f, path_xy, path_len


def map_host(points, affine=None, zmin=0.3, zmax=math.inf, origin=(0.0, 0.0), reso=0.2, shape=(1024, 1024), radius=0,
             variant="ccst", ctx=None, device=0):
    """numpy cloud [N,3|4] float32 -> inflated uint8 grid [W][H] through fx_map_host."""
    ctx = ctx or default_context(device)
    p = np.ascontiguousarray(points, dtype=np.float32)
    A = np.eye(3, 4, dtype=np.float32) if affine is None else np.ascontiguousarray(np.asarray(affine, dtype=np.float32).reshape(3, 4))
    W, H = int(shape[0]), int(shape[1])
    out = np.empty((W, H), dtype=np.uint8)
    step = 1 if variant == "ccst" else max(int(radius), 1)
    rc = ctx.lib.fx_map_host(ctx.handle, p.ctypes.data_as(C.c_void_p), p.shape[0], p.shape[1],
                             A.ctypes.data_as(C.POINTER(C.c_float)), float(zmin), float(zmax), float(origin[0]),
                             float(origin[1]), float(reso), W, H, int(radius), step, out.ctypes.data_as(C.c_void_p))
    ctx.check(rc, "fx_map_host")
    return out
