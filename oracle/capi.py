"""ctypes binding of oracle/fuxi_oracle.c (CPU ORACLE -- test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "fuxi_oracle.c")
_SO = os.path.join(_HERE, "libfuxi_oracle.so")
_lib = None

# integer Euclidean weights shared with the CUDA path (DESIGN.md "metric 2"): 2378 : 3363 ~ 1 : sqrt(2) (rel. error 4.4e-8)
FX_WS = 2378
FX_WD = 3363


def lib_path():
    return _SO


def build(force=False):
    """Compile the C oracle with gcc (needs nothing but libc/libm/libgomp)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp", "-std=c11", "-o", _SO, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        build(force=True)
    lib = C.CDLL(_SO)
    u8p, i32p, i64p, f64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                             C.POINTER(C.c_double))
    lib.fxo_jps.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                            f64p, i32p, C.c_int, i32p, i64p]
    lib.fxo_jps_batch.argtypes = [u8p, C.c_int, C.c_int, i32p, i32p, C.c_int, C.c_int, C.c_int, f64p, i32p, i64p]
    lib.fxo_sssp.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                             i64p, i64p]
    lib.fxo_sssp_batch.argtypes = [u8p, C.c_int, C.c_int, i32p, i32p, C.c_int, C.c_int64, C.c_int64, C.c_int, i64p]
    lib.fxo_inflate.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.fxo_edt.argtypes = [u8p, i32p, C.c_int, C.c_int]
    f32p = C.POINTER(C.c_float)
    lib.fxo_cloud_filter.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_double, C.c_int, f32p, C.c_int64, i64p]
    lib.fxo_cloud_filter.restype = C.c_int
    lib.fxo_jump.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    lib.fxo_jump.restype = C.c_int
    for f in ("fxo_jps", "fxo_jps_batch", "fxo_sssp", "fxo_sssp_batch", "fxo_inflate", "fxo_edt", "fxo_num_threads"):
        getattr(lib, f).restype = C.c_int
    _lib = lib
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _occ(matrix):
    """The reference's obstacle test is ``matrix[x][y] == 1`` (jps1.py:20-29): anything else is free."""
    m = np.asarray(matrix)
    return np.ascontiguousarray((m == 1).astype(np.uint8))


def num_threads():
    return int(_load().fxo_num_threads())


def jps(matrix, start, goal, hchoice, max_path=1 << 16):
    """Restated jps1.method.  Returns (path or 0, cost or None, expansions)."""
    occ = _occ(matrix)
    W, H = occ.shape
    cost = C.c_double(0.0)
    plen = C.c_int32(0)
    exp = C.c_int64(0)
    path = np.zeros((max_path, 2), dtype=np.int32)
    r = _load().fxo_jps(_p(occ, C.c_uint8), W, H, int(start[0]), int(start[1]), int(goal[0]), int(goal[1]),
                        int(hchoice), C.byref(cost), _p(path, C.c_int32), max_path, C.byref(plen), C.byref(exp))
    if r == -2:
        raise IndexError("start outside the grid (the reference raises IndexError too)")
    if r < 0:
        raise RuntimeError("fxo_jps failed: %d" % r)
    if r == 0:
        return 0, None, exp.value
    n = min(plen.value, max_path)
    return [tuple(int(v) for v in p) for p in path[:n]], cost.value, exp.value


def jump(matrix, cell, direction, goal):
    """Restated jps1.jump(cX, cY, dX, dY, matrix, goal): the jump point as a tuple, or None."""
    occ = _occ(matrix)
    W, H = occ.shape
    out = np.zeros(2, dtype=np.int32)
    r = _load().fxo_jump(_p(occ, C.c_uint8), W, H, int(cell[0]), int(cell[1]), int(direction[0]), int(direction[1]),
                         int(goal[0]), int(goal[1]), _p(out, C.c_int32))
    return (int(out[0]), int(out[1])) if r else None


def jps_batch(matrix, starts, goals, hchoice, threads=0):
    """Q queries over host threads.  Returns (cost float64[Q] (nan = no path), status int32[Q], threads used)."""
    occ = _occ(matrix)
    W, H = occ.shape
    s = np.ascontiguousarray(starts, dtype=np.int32).reshape(-1, 2)
    g = np.ascontiguousarray(goals, dtype=np.int32).reshape(-1, 2)
    Q = len(s)
    cost = np.zeros(Q, dtype=np.float64)
    status = np.zeros(Q, dtype=np.int32)
    exp = np.zeros(Q, dtype=np.int64)
    used = _load().fxo_jps_batch(_p(occ, C.c_uint8), W, H, _p(s, C.c_int32), _p(g, C.c_int32), Q, int(hchoice),
                                 int(threads), _p(cost, C.c_double), _p(status, C.c_int32), _p(exp, C.c_int64))
    if used < 0:
        raise RuntimeError("fxo_jps_batch failed: %d" % used)
    cost[status != 1] = np.nan
    return cost, status, used


def _weights(metric):
    if metric == 1:
        return 10, 14
    if metric == 2:
        return FX_WS, FX_WD
    raise ValueError("metric must be 1 or 2")


def sssp_field(matrix, source, metric=1, weights=None):
    """Exact cost-from-source field (int64, -1 unreachable) on the 'not blocked' graph."""
    occ = _occ(matrix)
    W, H = occ.shape
    ws, wd = weights if weights is not None else _weights(metric)
    field = np.empty((W, H), dtype=np.int64)
    r = _load().fxo_sssp(_p(occ, C.c_uint8), W, H, int(source[0]), int(source[1]), -1, -1, ws, wd,
                         _p(field, C.c_int64), None)
    if r < 0:
        raise RuntimeError("fxo_sssp failed: %d" % r)
    return field


def sssp_cost(matrix, start, goal, metric=1, weights=None):
    occ = _occ(matrix)
    W, H = occ.shape
    ws, wd = weights if weights is not None else _weights(metric)
    field = np.empty((W, H), dtype=np.int64)
    st = C.c_int64(0)
    r = _load().fxo_sssp(_p(occ, C.c_uint8), W, H, int(start[0]), int(start[1]), int(goal[0]), int(goal[1]),
                         ws, wd, _p(field, C.c_int64), C.byref(st))
    if r < 0:
        raise RuntimeError("fxo_sssp failed: %d" % r)
    return (int(field[int(goal[0]), int(goal[1])]) if r == 1 else -1), st.value


def sssp_batch(matrix, starts, goals, metric=1, threads=0, weights=None):
    occ = _occ(matrix)
    W, H = occ.shape
    ws, wd = weights if weights is not None else _weights(metric)
    s = np.ascontiguousarray(starts, dtype=np.int32).reshape(-1, 2)
    g = np.ascontiguousarray(goals, dtype=np.int32).reshape(-1, 2)
    cost = np.zeros(len(s), dtype=np.int64)
    r = _load().fxo_sssp_batch(_p(occ, C.c_uint8), W, H, _p(s, C.c_int32), _p(g, C.c_int32), len(s), ws, wd,
                               int(threads), _p(cost, C.c_int64))
    if r < 0:
        raise RuntimeError("fxo_sssp_batch failed: %d" % r)
    return cost


def inflate(grid, radius, step=1):
    g = np.ascontiguousarray(np.asarray(grid) > 0, dtype=np.uint8)
    out = np.empty_like(g)
    r = _load().fxo_inflate(_p(g, C.c_uint8), _p(out, C.c_uint8), g.shape[0], g.shape[1], int(radius), int(step))
    if r < 0:
        raise RuntimeError("fxo_inflate failed: %d" % r)
    return out


def edt(grid):
    g = np.ascontiguousarray(np.asarray(grid) > 0, dtype=np.uint8)
    out = np.empty(g.shape, dtype=np.int32)
    r = _load().fxo_edt(_p(g, C.c_uint8), _p(out, C.c_int32), g.shape[0], g.shape[1])
    if r < 0:
        raise RuntimeError("fxo_edt failed: %d" % r)
    return out


def cloud_filter(points, rgb_offset=-1, pass_lim=(0.0, 4.0), leaf=(0.17, 0.17, 0.2), radius=0.35, min_neighbors=13):
    """PassThrough(z) -> VoxelGrid -> RadiusOutlierRemoval (src/chen_filter_rgb.cpp:52-71; PCL restated, parity unpinned).
    points: float32 [n, stride]; returns (float32 [kept, 4] = x, y, z, rgb word, counts int64[4])."""
    p = np.ascontiguousarray(points, dtype=np.float32)
    n, stride = p.shape
    out = np.zeros((max(n, 1), 4), dtype=np.float32)
    counts = np.zeros(4, dtype=np.int64)
    r = _load().fxo_cloud_filter(_p(p, C.c_float), n, stride, int(rgb_offset), float(pass_lim[0]), float(pass_lim[1]),
                                 float(leaf[0]), float(leaf[1]), float(leaf[2]), float(radius), int(min_neighbors),
                                 _p(out, C.c_float), n, _p(counts, C.c_int64))
    if r < 0:
        raise RuntimeError("fxo_cloud_filter failed: %d" % r)
    return out[:counts[2]].copy(), counts
