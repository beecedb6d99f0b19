"""Planner-side host glue around the kernels: what the two planner main loops do between receiving an
OccupancyGrid and publishing a path (scripts/global_planner_st.py:15-25,226-298; scripts/global_planner_ccst.py:
17-31,411-495), with the inflation and the search on the B200.  Host integer bookkeeping stays on the host
(it is a few dozen scalar operations per replan); the grid work runs in fx_inflate / fx_search_batch.

`variant` is "st" (octomap planner: 9-point stencil, index offset -1, path offset [1,1]) or "ccst"
(ccmapping planner: dense square, no index offset, path offset [1,0]).
"""
from dataclasses import dataclass

import math

import numpy as np
import torch

from . import api
from ._lib import FuxiError


def occupancy_grid_to_array(data, width, height):
    """nav_msgs/OccupancyGrid.data (row-major int8, index y*width + x) -> int array [x][y] with the reference's
    value mapping (global_planner_st.py:15-20): 100 -> 1, -1 -> 0, everything else unchanged (1..99 stay and count
    as occupied for the inflation, whose source test is ``> 0``)."""
    a = np.asarray(data).reshape(int(height), int(width)).T.copy()
    a[a == 100] = 1
    a[a == -1] = 0
    return a


def array_to_occupancy_grid(mapu):
    """Inverse for publishing (global_planner_st.py:102-115): 1 -> 100, flattened y-major, int8."""
    a = np.array(mapu)
    a[a == 1] = 100
    return a.T.reshape(-1).astype(np.int8)


@dataclass
class Assembly:
    shape: tuple        # (W, H) of the padded planning grid
    paste_at: tuple     # where the incoming map's [0,0] lands in the padded grid (the reference's map_d)
    origin: tuple       # world coordinates of padded cell [0,0] (the reference's shifted map_o)
    start: tuple        # start cell in the padded grid
    goal: tuple         # goal cell in the padded grid


def plan_assembly(map_shape, map_origin, reso, start_xy, goal_xy, ifa, variant="st"):
    """Index / pad / shift arithmetic of global_planner_st.py:226-250,266-267 (ccst:411-436,452-453).

    World -> cell uses truncation toward zero like the reference's ``astype(int)``; the grid grows by 2*ifa below
    (more if start or goal has a negative index) and 4*ifa in total size beyond the furthest of map, start, goal."""
    W0, H0 = int(map_shape[0]), int(map_shape[1])
    o = np.asarray(map_origin, dtype=np.float64)[:2]
    cell = lambda p: np.trunc((np.asarray(p, dtype=np.float64)[:2] - o) / reso).astype(np.int64)
    s, g = cell(start_xy), cell(goal_xy)
    shift = np.array([-2 * ifa, -2 * ifa], dtype=np.int64)
    for ax in (0, 1):
        if s[ax] < 0 or g[ax] < 0:
            shift[ax] += min(s[ax], g[ax])
    d = np.abs(shift)
    W = max(W0, int(g[0]), int(s[0])) + int(d[0]) + 4 * ifa
    H = max(H0, int(g[1]), int(s[1])) + int(d[1]) + 4 * ifa
    off = -1 if variant == "st" else 0
    return Assembly((W, H), (int(d[0]), int(d[1])), tuple((shift * reso + o).tolist()),
                    tuple(int(v) for v in (s + d + off)), tuple(int(v) for v in (g + d + off)))


def relocate_goal(grid, goal):
    """Goal on an obstacle -> nearest free cell (== 0) along the goal's x-row, else along its y-column
    (global_planner_st.py:268-275).  Returns (goal, moved).  `grid` is a host array [x][y]."""
    gx, gy = int(goal[0]), int(goal[1])
    if grid[gx, gy] != 1:
        return (gx, gy), False
    free = np.flatnonzero(grid[gx, :] == 0)
    if free.size:
        return (gx, int(free[np.argmin(np.abs(free - gy))])), True
    free = np.flatnonzero(grid[:, gy] == 0)
    if free.size:
        return (int(free[np.argmin(np.abs(free - gx))]), gy), True
    raise FuxiError("goal (%d, %d): no free cell in its row or column" % (gx, gy))


def path_cells_to_world(path, reso, origin, variant="st"):
    """global_planner_st.py:292-298 / global_planner_ccst.py:487-495: cells (+[1,1] st / +[1,0] ccst) * reso + origin, z = 0."""
    p = np.asarray(path, dtype=np.float64).reshape(-1, 2) + (np.array([1.0, 1.0]) if variant == "st" else np.array([1.0, 0.0]))
    xy = p * reso + np.asarray(origin, dtype=np.float64)[:2]
    return np.concatenate([xy, np.zeros((len(xy), 1))], axis=1)


def inflate_host(mapu, ifa, variant="st", device=0, ctx=None):
    """Drop-in for the inline inflation blocks (global_planner_st.py:256-262, global_planner_ccst.py:442-448):
    host array in, float64 {0,1} host array out (the dtype the planners hand to jps1.method), inflation on the GPU."""
    src = torch.from_numpy(np.ascontiguousarray(np.asarray(mapu) > 0, dtype=np.uint8)).to("cuda:%d" % device)
    out = api.inflate(src, int(ifa), variant, ctx=ctx)
    return out.cpu().numpy().astype(np.float64)


@dataclass
class Replan:
    path_cells: object      # list of (x, y) in the padded grid, or 0 (the reference's "no path" literal)
    path_world: object      # float64 [k,3] or None
    cost: float             # gscore[goal] as jps1.method prints it, or None
    goal_moved: bool
    assembly: Assembly
    grid: torch.Tensor      # inflated uint8 planning grid on the device


def replan(mapu, map_origin, reso, start_xy, goal_xy, ifa=1, variant="st", hchoice=2, device=0, ctx=None, max_path=1024):
    """One global replan as the planner loops do it (st:226-298 / ccst:411-495): pad/shift, inflate, relocate the
    goal if it sits on an obstacle, search, convert the path to world coordinates.  Grid work on the GPU."""
    mapu = np.asarray(mapu)
    asm = plan_assembly(mapu.shape, map_origin, reso, start_xy, goal_xy, ifa, variant)
    dev = torch.device("cuda", device)
    padded = torch.zeros(asm.shape, dtype=torch.uint8, device=dev)
    px, py = asm.paste_at
    padded[px:px + mapu.shape[0], py:py + mapu.shape[1]] = torch.from_numpy(np.ascontiguousarray(mapu > 0, dtype=np.uint8)).to(dev)
    grid = api.inflate(padded, int(ifa), variant, ctx=ctx)
    W, H = asm.shape
    goal, moved = asm.goal, False
    if 0 <= goal[0] < W and 0 <= goal[1] < H and int(grid[goal[0], goal[1]]) == 1:
        row = grid[goal[0], :].cpu().numpy()
        if (row == 0).any():
            goal, moved = relocate_goal(row[None, :], (0, goal[1]))
            goal = (asm.goal[0], goal[1])
        else:
            col = grid[:, goal[1]].cpu().numpy()
            goal, moved = relocate_goal(col[:, None], (goal[0], 0))
            goal = (goal[0], asm.goal[1])
    s = torch.tensor([asm.start], dtype=torch.int32, device=dev)
    g = torch.tensor([goal], dtype=torch.int32, device=dev)
    res = api.plan_batch(grid, s, g, metric=hchoice, max_path=max_path, ctx=ctx)
    n = int(res.path_len[0])
    if n <= 0:
        return Replan(0, None, None, moved, asm, grid)
    cells = res.path(0)
    cost = float(res.cost_i[0]) if hchoice == 1 else float(res.cost_f[0])
    return Replan(cells, path_cells_to_world(cells, reso, asm.origin, variant), cost, moved, asm, grid)


def replan_fused(map_data, width, height, map_origin, reso, start_xy, goal_xy, ifa=1, variant="st", hchoice=2, layout="msg",
                 crop=None, shortcut=None, drop_radius=None, pos_z=0.0, ctx=None, device=0, max_path=1024, want_grid=False):
    """The same replan in ONE library call (fx_replan_host): the raw OccupancyGrid message goes to the device once,
    crop / pad / decode / inflate / goal relocation / search / path post-processing run there back to back, and the
    result comes back in one copy.  Defaults follow the two planners: "ccst" crops to the occupied bounding box
    (global_planner_ccst.py:36-63), shortcuts the path (ccst:515-521) and drops points within 1.5 m of the vehicle
    (ccst:507-513); "st" does none of these.  Returns a Replan (grid = host uint8 array when want_grid)."""
    cc = variant == "ccst"
    crop = cc if crop is None else crop
    shortcut = cc if shortcut is None else shortcut
    drop_radius = (1.5 if cc else 0.0) if drop_radius is None else drop_radius
    drop = (float(start_xy[0]), float(start_xy[1]), float(pos_z), float(drop_radius)) if drop_radius > 0 else None
    out, cells, world, grid = api.replan_host(map_data, width, height, map_origin, reso, start_xy, goal_xy, ifa=ifa, variant=variant,
                                              hchoice=hchoice, layout=layout, crop=crop, shortcut=shortcut, drop=drop,
                                              max_path=max_path, want_grid=want_grid, ctx=ctx, device=device)
    asm = Assembly((out.W, out.H), (out.paste_x, out.paste_y), (out.origin_x, out.origin_y), (out.start_x, out.start_y),
                   (out.goal_x, out.goal_y))
    if out.skipped or out.path_len <= 0:
        return Replan(0, None, None, out.goal_moved == 1, asm, grid)
    cost = float(out.cost_i) if hchoice == 1 else float(out.cost_f)
    return Replan([tuple(int(v) for v in p) for p in cells.tolist()], world, cost, out.goal_moved == 1, asm, grid)


def waypoint_ccst(path_world, global_goal):
    """global_planner_ccst.py:523-526: blend of the second and third path points, or the goal for short paths."""
    p = np.asarray(path_world)
    return (p[1] * 1.4 + p[2] * 0.6) / 2 if len(p) > 2 else np.asarray(global_goal)


def waypoint_st(path_cells, map_start, reso, origin, global_goal, pos, end_occu, prev_wp=None,
                dis_wp_tre=2.0, ang_wp_tre=math.pi / 4):
    """Next waypoint as the Octomap planner picks it (global_planner_st.py:291-325).

    path_cells: the path `jps1.method` returned (cells of the planning grid); the planner works on ``path2 = path + [1, 1]``.
    Walk the path until the bearing from the vehicle stops closing in on the bearing of the path's end (and the point
    is more than 2 cells away): the point before that is the candidate.  It is kept only if the path has more than
    two points and the candidate is farther than `dis_wp_tre` or the last bearing gap lies in (ang_wp_tre, pi/2);
    otherwise the goal itself is the waypoint; with `end_occu == 1` the vehicle holds position and that position becomes
    the goal.  `prev_wp` is the waypoint of the previous iteration (the reference's `wp` survives between iterations and
    is what `uav2next_wp` measures when no candidate is found).  Returns (wp, global_goal, ang_wp)."""
    path2 = np.asarray(path_cells, dtype=np.int64) + np.array([1, 1])
    start = np.asarray(map_start)
    goal = np.asarray(global_goal, dtype=np.float64)
    end_bearing = math.atan2(*(path2[-1] - start))          # atan2(dx, dy): the reference's argument order
    wp, gap = prev_wp, 0.0
    for k in range(1, len(path2)):
        rel = path2[k] - start
        here = abs(end_bearing - math.atan2(rel[0], rel[1]))
        if here <= gap and np.linalg.norm(rel) > 2:
            wp = path2[k - 1] * reso + np.asarray(origin, dtype=np.float64)
            break
        gap = here
    if wp is None:
        wp = goal
    far = float(np.linalg.norm(np.asarray(wp)[0:2] - np.asarray(pos, dtype=np.float64)[0:2]))
    if end_occu == 1:
        wp = np.asarray(pos, dtype=np.float64)
        goal = wp
    elif not (len(path2) > 2 and (far > dis_wp_tre or (ang_wp_tre < gap < math.pi * 0.5))):
        wp = goal
    return wp, goal, gap
