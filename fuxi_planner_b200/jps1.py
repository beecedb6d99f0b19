"""Drop-in for the reference's ``scripts/jps1.py``: same name, same signature, same return and print
conventions for the one function the planners call --

    path1 = jps1.method(mapu, tuple(map_start), tuple(map_goal), 2)      # global_planner_st.py:285
                                                                          # global_planner_ccst.py:477

-- but the search runs on the B200 through ``fx_plan_host`` (include/fuxi_b200.h).  Put this module's
directory first on ``sys.path`` (or replace the file) and ``import jps1`` in the planners picks it up.

Contract kept from scripts/jps1.py:183-230:
  * ``matrix`` is indexed ``[x][y]``; a cell is an obstacle iff ``matrix[x][y] == 1`` (:20-29).
  * success  -> ``(path, secs)``: list of 2-tuples, ``path[0]`` is the ``start`` object, ``path[-1]`` the goal,
    consecutive points joined by a straight 8-direction run; ``secs = round(wall, 6)``; the cost is printed
    on stdout like the reference's ``print(gscore[goal])`` (:207).
  * no path  -> ``(0, secs)`` with the int literal 0 (callers test ``path1[0] is 0``, global_planner_st.py:287).
  * start outside the array -> IndexError (numpy raises it in the reference).
Interior points: by default the turning points of an optimal path (all the callers need: they index path[1], path[2]
and convert to an ndarray).  ``jps1.POINTS = "jump"`` returns the jump points instead -- every cell of the path at which
the reference's ``jump`` (:95-164) would have stopped, i.e. the list the reference itself returns for that cell path
(when several optimal paths exist the reference's heap order picks one of them, the wavefront another: same cost).
"""
import time

import numpy as np

from . import api
from ._lib import FX_COST_OVERFLOW, FX_COST_START_OOB, FuxiError

_MAX_PATH = 4096    # turning points kept per call before a retry (a path across a 4096^2 grid at 20 % fill has 300-1100)
POINTS = "turning"      # or "jump": the reference's jump-point list (fx_jump_points_host)


class _Fast:
    """Per-process buffers of the single-query call: on the reference's own maps a replan is ~0.1 ms, of which building
    seven numpy arrays and their ctypes pointers per call was 20 us.  Query and result buffers are allocated once and
    their addresses cached; the matrix pointer is the only thing taken per call."""

    def __init__(self):
        ctx = api.default_context(0)
        self.ctx = ctx
        self.fn = ctx.lib.fx_plan_host_f64
        self.q = np.zeros(4, dtype=np.int32)                      # start x, y, goal x, y
        self.cost_i = np.zeros(1, dtype=np.int32)
        self.cost_f = np.zeros(1, dtype=np.float64)
        self.path_len = np.zeros(1, dtype=np.int32)
        self.path_xy = np.zeros((1, _MAX_PATH, 2), dtype=np.int32)
        self.p_s = self.q.ctypes.data
        self.p_g = self.q.ctypes.data + 8
        self.p_ci, self.p_cf = self.cost_i.ctypes.data, self.cost_f.ctypes.data
        self.p_len, self.p_path = self.path_len.ctypes.data, self.path_xy.ctypes.data


_fast = None
_INT = (int, np.integer)


def _plan_one(occ, start, goal, hchoice):
    """(cost_i, cost_f, n, path points int32 [n][2]) of one query; float64 C-contiguous matrices with integer endpoints take the cached
    buffers, everything else the general wrapper."""
    global _fast
    if (occ.dtype == np.float64 and occ.ndim == 2 and occ.flags.c_contiguous and isinstance(start[0], _INT)
            and isinstance(start[1], _INT) and isinstance(goal[0], _INT) and isinstance(goal[1], _INT)
            and max(abs(int(start[0])), abs(int(start[1])), abs(int(goal[0])), abs(int(goal[1]))) < 2 ** 31):
        f = _fast
        if f is None or f.ctx.handle is None:
            f = _fast = _Fast()
        q = f.q
        q[0], q[1], q[2], q[3] = start[0], start[1], goal[0], goal[1]
        rc = f.fn(f.ctx.handle, occ.ctypes.data, occ.shape[0], occ.shape[1], f.p_s, f.p_g, 1, hchoice, f.p_ci, f.p_cf,
                  f.p_path, f.p_len, _MAX_PATH)
        f.ctx.check(rc, "fx_plan_host_f64")
        n = int(f.path_len[0])
        if n <= _MAX_PATH:
            return int(f.cost_i[0]), float(f.cost_f[0]), n, f.path_xy[0, :max(n, 0)]
        first = n                      # longer than the cached buffer: one more search with room for exactly that many
    else:
        first = _MAX_PATH
    if occ.dtype != np.float64:        # float64 (what the planners pass) is compared `== 1` inside the library
        occ = (occ == 1).astype(np.uint8)
    max_path = first
    while True:
        cost_i, cost_f, path_xy, path_len = api.plan_host(occ, [start], [goal], metric=hchoice, max_path=max_path)
        n = int(path_len[0])
        if n <= max_path:
            break
        max_path = n
    return int(cost_i[0]), float(cost_f[0]), n, path_xy[0, :max(n, 0)]


def method(matrix, start, goal, hchoice):
    starttime = time.time()
    if hchoice not in (1, 2):
        raise ValueError("hchoice must be 1 or 2")
    occ = np.asarray(matrix)
    if POINTS == "jump":
        # the jump-point list needs a forward-canonical path: forward searches only (fx_set_search_form)
        ctx = api.default_context(0)
        ctx.set_search_form("throughput")
        try:
            cost_i, cost_f, n, rows = _plan_one(occ, start, goal, hchoice)
        finally:
            ctx.set_search_form("auto")
    else:
        cost_i, cost_f, n, rows = _plan_one(occ, start, goal, hchoice)
    if cost_i == FX_COST_START_OOB:
        raise IndexError("index %r is out of bounds for the %dx%d map" % (tuple(start), occ.shape[0], occ.shape[1]))
    if cost_i == FX_COST_OVERFLOW:
        raise FuxiError("search overflowed its 31-bit cost range or frontier queue on this map")
    endtime = time.time()
    if cost_i < 0:
        return (0, round(endtime - starttime, 6))
    if POINTS == "jump" and n > 1:
        rows = np.asarray(api.jump_points_host(occ, rows.tolist()), dtype=np.int64).reshape(-1, 2)
    data = list(zip(rows[:, 0].tolist(), rows[:, 1].tolist()))   # one pass in C: a 1000-point path cost 0.23 ms as a comprehension
    data[0] = start
    if n == 1:
        print(0)
    else:
        print(float(cost_i) if hchoice == 1 else cost_f)
    return (data, round(endtime - starttime, 6))
