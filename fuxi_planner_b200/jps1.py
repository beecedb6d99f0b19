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
Difference (documented in DESIGN.md): interior points are turning points of an optimal path, not the
reference's jump points; the cost is identical.
"""
import time

import numpy as np

from . import api
from ._lib import FX_COST_OVERFLOW, FX_COST_START_OOB, FuxiError

_MAX_PATH = 1024


def method(matrix, start, goal, hchoice):
    starttime = time.time()
    if hchoice not in (1, 2):
        raise ValueError("hchoice must be 1 or 2")
    occ = np.asarray(matrix)
    if occ.dtype != np.float64:        # float64 (what the planners pass) is compared `== 1` inside the library
        occ = (occ == 1).astype(np.uint8)
    max_path = _MAX_PATH
    while True:
        cost_i, cost_f, path_xy, path_len = api.plan_host(occ, [start], [goal], metric=hchoice, max_path=max_path)
        n = int(path_len[0])
        if n <= max_path:
            break
        max_path = n
    if cost_i[0] == FX_COST_START_OOB:
        raise IndexError("index %r is out of bounds for the %dx%d map" % (tuple(start), occ.shape[0], occ.shape[1]))
    if cost_i[0] == FX_COST_OVERFLOW:
        raise FuxiError("search overflowed its 31-bit cost range or frontier queue on this map")
    endtime = time.time()
    if cost_i[0] < 0:
        return (0, round(endtime - starttime, 6))
    data = [tuple(int(v) for v in p) for p in path_xy[0, :n]]
    data[0] = start
    if n == 1:
        print(0)
    else:
        print(float(cost_i[0]) if hchoice == 1 else float(cost_f[0]))
    return (data, round(endtime - starttime, 6))
