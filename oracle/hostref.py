"""numpy restatements of the reference's inline (``__main__``-body) hot-path blocks.

CPU ORACLE -- test infrastructure only.  Each function cites the reference lines it follows
(paths relative to the fuxi-planner repo).  These blocks are not callable in the reference
(they live inside ``if __name__ == '__main__':`` loops), hence the restatement.
"""
import math

import numpy as np


# --------------------------------------------------------------------------- a12
def decode_occupancy_grid(data, width, height):
    """scripts/global_planner_st.py:15-20 (= global_planner_ccst.py:17-24).

    nav_msgs/OccupancyGrid row-major int8 (index = y*width + x) -> array[x][y]; 100 -> 1, -1 -> 0;
    other values (1..99) survive and count as occupied for the inflation (``> 0``)."""
    m = np.array(data).reshape(height, width).T
    m = m.copy()
    m[np.where(m == 100)] = 1
    m[np.where(m == -1)] = 0
    return m


def encode_occupancy_grid(mapu):
    """scripts/global_planner_st.py:102-115 (publish_map): 1 -> 100, then data = mapu.T.flatten()."""
    m = np.array(mapu).copy()
    m[np.where(m == 1)] = 100
    return m.T.reshape(-1).astype(np.int8)


# --------------------------------------------------------------------------- a13
def assemble_grid(mapu, map_o, map_reso, start_xy, goal_xy, ifa, variant="st"):
    """scripts/global_planner_st.py:226-250,266-267 / global_planner_ccst.py:411-436,452-453.

    Returns (padded float64 grid, map_start, map_goal, new map_o, map_d).  ``astype(int)`` truncates
    toward zero (not floor), as in the reference."""
    mapu = np.asarray(mapu)
    map_c, map_r = mapu.shape
    map_o = np.array(map_o, dtype=float)
    map_goal = ((np.array(goal_xy[0:2], dtype=float) - map_o) / map_reso).astype(int)
    map_start = ((np.array(start_xy[0:2], dtype=float) - map_o) / map_reso).astype(int)
    map_o2 = np.array([-2 * ifa, -2 * ifa])
    if map_goal[0] < 0 or map_start[0] < 0:
        map_o2[0] = min(map_goal[0], map_start[0]) + map_o2[0]
    if map_goal[1] < 0 or map_start[1] < 0:
        map_o2[1] = min(map_goal[1], map_start[1]) + map_o2[1]
    map_d = abs(map_o2)
    new_o = list(map_o2 * map_reso + map_o)
    map_c = max(map_c, map_goal[0], map_start[0]) + map_d[0]
    map_r = max(map_r, map_goal[1], map_start[1]) + map_d[1]
    mapu0 = np.zeros([map_c + 4 * ifa, map_r + 4 * ifa])
    mapu0[map_d[0]:len(mapu) + map_d[0], map_d[1]:len(mapu[0]) + map_d[1]] = mapu
    if variant == "st":
        map_start = map_start + map_d - 1
        map_goal = map_goal + map_d - 1
    else:
        map_start = map_start + map_d
        map_goal = map_goal + map_d
    return mapu0, map_start, map_goal, new_o, map_d


# --------------------------------------------------------------------------- a10 / a11
def inflate_st(mapu, ifa):
    """scripts/global_planner_st.py:256-262: 9-point stencil {-ifa, 0, +ifa}^2 (dense 3x3 for ifa=1)."""
    mapu = np.array(mapu, dtype=np.float64)
    occ = np.where(mapu > 0)
    for i in range(-ifa, ifa + 1, ifa):
        for j in range(-ifa, ifa + 1, ifa):
            mapu[(occ[0] + i, occ[1] + j)] = 1
    return mapu


def inflate_ccst(mapu, ifa):
    """scripts/global_planner_ccst.py:442-448: dense (2*ifa+1)^2 square dilation."""
    mapu = np.array(mapu, dtype=np.float64)
    occ = np.where(mapu > 0)
    for i in range(-ifa, ifa + 1, 1):
        for j in range(-ifa, ifa + 1, 1):
            mapu[(occ[0] + i, occ[1] + j)] = 1
    return mapu


# --------------------------------------------------------------------------- a14
def relocate_goal(mapu, map_goal):
    """scripts/global_planner_st.py:268-275 (= ccst:454-458).  Returns (goal, end_occu)."""
    map_goal = np.array(map_goal).copy()
    if mapu[map_goal[0], map_goal[1]] == 1:
        try:
            free = np.where(mapu[map_goal[0], :] == 0)
            map_goal[1] = free[0][np.argmin(abs(free - map_goal[1]))]
        except Exception:
            free = np.where(mapu[:, map_goal[1]] == 0)
            map_goal[0] = free[0][np.argmin(abs(free - map_goal[0]))]
        return map_goal, 1
    return map_goal, 0


def path_to_world(path, map_reso, map_o, variant="st"):
    """scripts/global_planner_st.py:292-298 (offset [1,1]) / global_planner_ccst.py:487-495 (offset [1,0])."""
    off = np.array([1, 1]) if variant == "st" else np.array([1, 0])
    path2 = np.array(path) + off
    path3 = path2 * map_reso + map_o
    return np.c_[path3, np.zeros([len(path3), 1])]


def map_line_col(p2, p1, mapu):
    """scripts/global_planner_ccst.py:258-283: True = no occupied cell on the sampled line (one sample per
    integer x strictly between the endpoints, y = rint(slope * x) + int(p1.y)); False = collision."""
    blc = np.where(mapu == 1)
    if len(blc[0]) > 0:
        p1 = np.array(p1)
        p2 = np.array(p2)
        p0 = np.array([min(p1[0], p2[0]), min(p1[1], p2[1])]).astype(float)
        p1 = p1 - p0
        p2 = p2 - p0
        if p2[0] < p1[0]:
            p1, p2 = p2.copy(), p1.copy()
        xs = np.arange(p1[0] + 1, p2[0], 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            ys = np.rint((p2[1] - p1[1]) / (p2[0] - p1[0]) * xs).astype(int) + int(p1[1])
        occ = set(zip(blc[0].tolist(), blc[1].tolist()))
        for x, y in zip(xs.astype(int).tolist(), ys.tolist()):
            if (x, y) in occ:
                return False
    return True


def shortcut_path(path, mapu):
    """scripts/global_planner_ccst.py:515-521: greedy line-of-sight shortcutting of the jump-point list -- drop the
    middle point ii whenever map_line_col sees no occupied cell between its neighbours, looking only at the part of
    the map inside their bounding box (exclusive upper edges, so axis-aligned pairs always "see" each other)."""
    pts = np.array(path)
    ii = 1
    while ii < len(pts) - 1:
        a, b = pts[ii - 1], pts[ii + 1]
        crop = mapu[min(a[0], b[0]):max(a[0], b[0]), min(a[1], b[1]):max(a[1], b[1])]
        if map_line_col(pts[ii + 1], pts[ii - 1], crop):
            pts = np.delete(pts, ii, axis=0)
        else:
            ii += 1
    return pts.tolist()


def near_drop(path_cells, path_world, pos, radius=1.5):
    """scripts/global_planner_ccst.py:507-513: if the path has more than two points, delete every point ii >= 1 whose
    world position (z = 0, ``path3``) is closer than ``radius`` to the vehicle ``pos = (px, py, pz)``.
    Returns (cells, world) after the deletion."""
    path_cells = np.array(path_cells)
    path_world = np.array(path_world)
    del_path = []
    if len(path_world) > 2:
        for ii in range(1, len(path_world)):
            if np.linalg.norm(path_world[ii] - np.array(pos)) < radius:
                del_path.append(ii)
        path_world = np.delete(path_world, del_path, axis=0)
        path_cells = np.delete(path_cells, del_path, axis=0)
    return path_cells, path_world


def remove_zero_rowscols(X, px, py, map_o, map_reso):
    """scripts/global_planner_ccst.py:36-63 (the arithmetic, without the ROS wait loop and the object state).

    Returns None for an all-zero map (the reference returns 0), else (crop, map_c, map_r, new map_o): the crop window is
    ``X[min(min_row, start_x):max_row, min(min_col, start_y):max_col]`` -- exclusive upper ends, so the last occupied
    row and column are dropped -- while map_c / map_r are computed from the indices (they can differ from the crop's
    shape when the start index is negative: numpy wraps a negative slice start)."""
    X = np.asarray(X)
    rows, cols = X.nonzero()
    u_row, u_col = np.unique(rows), np.unique(cols)
    if len(u_row) == 0:
        return None
    map_o = np.array(map_o, dtype=float)
    map_start0 = ((np.array([px, py]) - map_o) / map_reso).astype(int)
    map_mat_o = [min(u_row), min(u_col)]
    r0, c0 = min(map_mat_o[0], map_start0[0]), min(map_mat_o[1], map_start0[1])
    map_c = max(u_row) - r0
    map_r = max(u_col) - c0
    new_o = np.array([r0, c0]) * map_reso + map_o
    x_row = X[r0:max(u_row)]
    return x_row[:, c0:max(u_col)], int(map_c), int(map_r), new_o


def replan_pipeline(mapu, map_o, map_reso, start_xy, goal_xy, ifa, variant="st", crop=False):
    """The planner loop from the decoded map to the search call, restated: [crop ccst:36-63] -> index/pad/shift
    (st:226-250 / ccst:411-436) -> inflation (st:256-262 / ccst:442-448) -> goal relocation and end_occu (st:266-275 /
    ccst:452-464) -> the `start beyond the map` guard (st:280 / ccst:466).
    Returns dict(grid float64 {0,1}-ish, start, goal, moved, end_occu, skipped, origin, map_d) or None (empty map)."""
    mapu = np.asarray(mapu)
    map_c, map_r = mapu.shape
    map_o = np.array(map_o, dtype=float)
    if crop:
        r = remove_zero_rowscols(mapu, start_xy[0], start_xy[1], map_o, map_reso)
        if r is None or not (r[1] * r[2] > 0):
            return None
        mapu, map_c, map_r, map_o = r
    map_goal = ((np.array(goal_xy[0:2], dtype=float) - map_o) / map_reso).astype(int)
    map_start = ((np.array(start_xy[0:2], dtype=float) - map_o) / map_reso).astype(int)
    map_o2 = np.array([-2 * ifa, -2 * ifa])
    if map_goal[0] < 0 or map_start[0] < 0:
        map_o2[0] = min(map_goal[0], map_start[0]) + map_o2[0]
    if map_goal[1] < 0 or map_start[1] < 0:
        map_o2[1] = min(map_goal[1], map_start[1]) + map_o2[1]
    map_d = abs(map_o2)
    new_o = list(map_o2 * map_reso + np.array(map_o))
    map_c = max(map_c, map_goal[0], map_start[0]) + map_d[0]
    map_r = max(map_r, map_goal[1], map_start[1]) + map_d[1]
    mapu0 = np.zeros([map_c + 4 * ifa, map_r + 4 * ifa])
    mapu0[map_d[0]:len(mapu) + map_d[0], map_d[1]:(len(mapu[0]) if len(mapu) else 0) + map_d[1]] = mapu
    mapu = inflate_st(mapu0, ifa) if variant == "st" else inflate_ccst(mapu0, ifa)
    off = -1 if variant == "st" else 0
    map_start = map_start + map_d + off
    map_goal = map_goal + map_d + off
    g0 = map_goal.copy()
    map_goal, moved = relocate_goal(mapu, map_goal)
    if variant == "st":
        end_occu = moved
    else:
        end_occu = int((mapu[map_goal[0] - ifa:map_goal[0] + ifa, map_goal[1] - ifa:map_goal[1] + ifa] == 1).any())
    skipped = bool(map_start[0] > map_c or map_start[1] > map_r)
    return dict(grid=mapu, start=tuple(int(v) for v in map_start), goal=tuple(int(v) for v in map_goal),
                goal0=tuple(int(v) for v in g0), moved=int(moved), end_occu=int(end_occu), skipped=skipped,
                origin=[float(v) for v in new_o], map_d=tuple(int(v) for v in map_d))


# --------------------------------------------------------------------------- a15 / a16
def body_to_earth_frame(ii, jj, kk):
    """scripts/utils.py:21-28: R = Rz(kk) * Ry(jj) * Rx(ii)."""
    ci, cj, ck = math.cos(ii), math.cos(jj), math.cos(kk)
    si, sj, sk = math.sin(ii), math.sin(jj), math.sin(kk)
    return np.array([[ck * cj, ck * sj * si - sk * ci, ck * sj * ci + sk * si],
                     [sk * cj, sk * sj * si + ck * ci, sk * sj * ci - ck * si],
                     [-sj, cj * si, cj * ci]])


def transform_cloud(plc, rpy, pos, dt=0.0, ang_vel=(0.0, 0.0, 0.0), line_vel=(0.0, 0.0, 0.0)):
    """scripts/plc_point2_st.py:244-251: camera -> body axis swap (+0.12 m lever arm), attitude and
    position extrapolated by dt, then earth frame.  float64 like the reference."""
    plc = np.asarray(plc, dtype=np.float64)
    plc_c = np.zeros([len(plc), 3])
    plc_c[:, 0] = plc[:, 2] + 0.12
    plc_c[:, 1] = -plc[:, 0]
    plc_c[:, 2] = -plc[:, 1]
    r, p, y = np.array(rpy, dtype=float) + dt * np.array(ang_vel, dtype=float)
    b2e = body_to_earth_frame(r, p, y)
    local_pos1 = dt * np.array(line_vel, dtype=float) + np.array(pos, dtype=float)
    return np.matmul(b2e, plc_c.T).T + np.tile(local_pos1, (len(plc), 1))


def cloud_affine(rpy, pos, dt=0.0, ang_vel=(0.0, 0.0, 0.0), line_vel=(0.0, 0.0, 0.0)):
    """The same transform folded into one 3x4 [R*S | R*l + t] matrix acting on camera-frame (x,y,z):
    S is the axis swap (x_b=z_c, y_b=-x_c, z_b=-y_c), l = (0.12,0,0).  This is what fx_project takes."""
    r, p, y = np.array(rpy, dtype=float) + dt * np.array(ang_vel, dtype=float)
    R = body_to_earth_frame(r, p, y)
    S = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    t = dt * np.array(line_vel, dtype=float) + np.array(pos, dtype=float)
    A = np.zeros((3, 4))
    A[:, :3] = R @ S
    A[:, 3] = R @ np.array([0.12, 0.0, 0.0]) + t
    return A


def height_filter(plc_c, zmin=0.3):
    """scripts/plc_point2_st.py:255-256."""
    plc_c = np.asarray(plc_c)
    return plc_c[plc_c[:, 2] > zmin]


def node_cloud(plc, rpy, pos, dt=0.0, ang_vel=(0.0, 0.0, 0.0), line_vel=(0.0, 0.0, 0.0), local_pos=None, zmin=0.3, dis=4.0):
    """What the cloud node publishes on /points_global_all, scripts/plc_point2_st.py:243-256 + 336-339: transform, height
    filter, `plc1 - local_pos`, distance_filter(.., 4), `+ local_pos`.  The rotation is applied with an explicit summation
    order, ((R_k0*b0 + R_k1*b1) + R_k2*b2) + t_k (the reference hands the product to BLAS through np.matmul, whose order is
    not specified: tests/golden/make_cloud_golden.py records how far the two are apart -- a few ulps)."""
    plc = np.asarray(plc, dtype=np.float64)
    b0, b1, b2 = plc[:, 2] + 0.12, -plc[:, 0], -plc[:, 1]
    r, p, y = np.array(rpy, dtype=float) + dt * np.array(ang_vel, dtype=float)
    R = body_to_earth_frame(r, p, y)
    t = dt * np.array(line_vel, dtype=float) + np.array(pos, dtype=float)
    e = np.stack([((R[k, 0] * b0 + R[k, 1] * b1) + R[k, 2] * b2) + t[k] for k in range(3)], axis=1)
    e = e[e[:, 2] > zmin]
    c = np.array(pos if local_pos is None else local_pos, dtype=float)
    return distance_filter(e - c, dis) + c


def octomap_local(centres, local_pos, box=4.0, dis=4.0):
    """/octomap_point_cloud_centers_local, scripts/plc_point2_st.py:357-362."""
    c = np.array(local_pos, dtype=float)
    q = np.asarray(centres, dtype=np.float64) - c
    q = q[(abs(q) < box).all(axis=1)]
    return distance_filter(q, dis) + c


# --------------------------------------------------------------------------- a17 / a18
def distance_filter(plc, dis):
    """scripts/plc_point2_st.py:139-148: keep |p| < dis, sort by (d, z, y, x) (np.lexsort: last key primary)."""
    plc = np.asarray(plc, dtype=np.float64)
    d_point = np.linalg.norm(plc, axis=1)
    filtered = np.c_[plc, d_point]
    filtered = filtered[d_point < dis]
    if len(filtered) > 0:
        filtered = filtered[np.lexsort(filtered.T)]
    return filtered[:, 0:3]


def pack_pointcloud2(points):
    """scripts/plc_point2_st.py:112-138: float32 x,y,z at offsets 0/4/8, point_step 12, little endian.
    (The reference's ``.tostring()`` is ``.tobytes()`` on numpy >= 2.)"""
    pts = np.asarray(points, np.float32)
    return {"height": 1, "width": len(pts), "point_step": 12, "row_step": 12 * pts.shape[0],
            "is_bigendian": False, "is_dense": int(np.isfinite(pts).all()),
            "data": pts.astype("<f4").tobytes()}


# --------------------------------------------------------------------------- a20 (NEW semantics, parity unpinned)
def project(points, affine, zmin, zmax, ox, oy, reso, W, H):
    """3D -> 2D projection as DEFINED by this framework (the reference delegates it to octomap_server /
    ccmapping, un-vendored: launch/map_st.launch:3, launch/map_ccst.launch:2).

    All arithmetic in float32, one rounding per operation, fixed order, no FMA:
        e_k = ((a_k0*x + a_k1*y) + a_k2*z) + a_k3
        fx  = floor((e_x - ox) / reso),  fy = floor((e_y - oy) / reso)
        occupied[fx, fy] = 1  iff  zmin < e_z <= zmax  and  0 <= fx < W  and  0 <= fy < H
    """
    f = np.float32
    p = np.asarray(points, dtype=f)
    A = np.asarray(affine, dtype=f).reshape(3, 4)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    e = []
    for k in range(3):
        e.append(((A[k, 0] * x + A[k, 1] * y) + A[k, 2] * z) + A[k, 3])
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        fx = np.floor((e[0] - f(ox)) / f(reso))
        fy = np.floor((e[1] - f(oy)) / f(reso))
        ok = (e[2] > f(zmin)) & (e[2] <= f(zmax)) & (fx >= 0) & (fx < W) & (fy >= 0) & (fy < H)
    grid = np.zeros((W, H), dtype=np.uint8)
    grid[fx[ok].astype(np.int64), fy[ok].astype(np.int64)] = 1
    return grid


# --------------------------------------------------------------------------- Appendix B (PNG fixtures)
def png_to_grid(img_l):
    """Inverse of the save block scripts/global_planner_st.py:365-374: 'L' image array -> m[x][y], 1 = occupied."""
    a = np.asarray(img_l)
    return (a[::-1].T == 0).astype(np.uint8)


def grid_to_png_array(mapu):
    """scripts/global_planner_st.py:368-372: 0 -> 255 (free), non-zero -> 0, then .T[::-1]."""
    mapu = np.asarray(mapu)
    ms = mapu.copy()
    ms[mapu != 0] = 0
    ms[mapu == 0] = 255
    return np.uint8(ms.T[::-1])


def load_premap(img_l, threshold=200):
    """scripts/global_planner_st.py:176-182: img.point(lambda x: 0 if x > 200 else 1); map_pre = img[::-1].T."""
    a = np.asarray(img_l)
    return np.where(a > threshold, 0, 1).astype(np.uint8)[::-1].T


def merge_premap(mapu, map_o, map_t, map_pre, ori_pre, reso):
    """scripts/global_planner_st.py:210-224, line by line (float64 grid like np.zeros gives)."""
    mapu = np.asarray(mapu)
    map_c, map_r = mapu.shape
    l1_pre, l2_pre = len(map_pre), len(map_pre[0])
    t_pre = [ori_pre[0] + reso * l1_pre, ori_pre[1] + reso * l2_pre]
    map_o1 = [min(map_o[0], ori_pre[0]), min(map_o[1], ori_pre[1])]
    map_oi = ((np.array(map_o) - map_o1) / reso).astype(int)
    ori_prei = ((np.array(ori_pre) - map_o1) / reso).astype(int)
    map_c1 = int((max(t_pre[0], map_t[0]) - map_o1[0]) / reso)
    map_r1 = int((max(t_pre[1], map_t[1]) - map_o1[1]) / reso)
    mapu0 = np.zeros([map_c1, map_r1])
    mapu0[ori_prei[0]:ori_prei[0] + l1_pre, ori_prei[1]:ori_prei[1] + l2_pre] = map_pre
    mapu0[map_oi[0]:map_oi[0] + map_c, map_oi[1]:map_oi[1] + map_r] = mapu
    return mapu0, map_o1


def waypoint_st(path1, map_start, map_reso, map_o, global_goal, px, py, pz, end_occu, wp=None, dis_wp_tre=2,
                ang_wp_tre=math.pi / 4):
    """scripts/global_planner_st.py:291-325 line by line (`wp` comes in with the previous iteration's value, st:127)."""
    path2 = np.array(path1) + np.array([1, 1])
    ang_wp = 0
    for k in range(1, len(path2)):
        map_wp = path2[k]
        if abs(math.atan2((path2[-1] - map_start)[0], (path2[-1] - map_start)[1]) - math.atan2((map_wp - map_start)[0], (map_wp - map_start)[1])) <= ang_wp and np.linalg.norm(map_wp - map_start) > 2:
            map_wp = path2[k - 1]
            wp = map_wp * map_reso + map_o
            break
        ang_wp = abs(math.atan2((path2[-1] - map_start)[0], (path2[-1] - map_start)[1]) - math.atan2((map_wp - map_start)[0], (map_wp - map_start)[1]))
    if wp is None:
        wp = global_goal
    uav2next_wp = np.linalg.norm(wp[0:2] - np.array([px, py]))
    if end_occu == 1:
        wp = np.array([px, py, pz])
        global_goal = wp
    elif not (len(path2) > 2 and (uav2next_wp > dis_wp_tre or (ang_wp > ang_wp_tre and ang_wp < math.pi * 0.5))):
        wp = global_goal
    return wp, global_goal, ang_wp
