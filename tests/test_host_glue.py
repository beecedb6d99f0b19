"""Product host glue (fuxi_planner_b200/cloud.py, planner.py) against the oracle's line-by-line restatements of the
reference's inline blocks (oracle/hostref.py, themselves pinned to the reference functions that are importable,
tests/golden/hostfn_golden.npz).  CPU only: these are the small host computations around the kernels."""
import numpy as np
import pytest

from fuxi_planner_b200 import cloud, planner


def test_cloud_affine_matches_reference_transform(oracle):
    rng = np.random.default_rng(0)
    for _ in range(100):
        rpy, pos = rng.uniform(-1.2, 1.2, 3), rng.uniform(-20, 20, 3)
        dt, av, lv = rng.uniform(0, 0.1), rng.uniform(-3, 3, 3), rng.uniform(-2, 2, 3)
        A = cloud.cloud_affine(rpy, pos, dt, av, lv)
        pts = rng.uniform(-6, 6, (50, 3))
        want = oracle.hostref.transform_cloud(pts, rpy, pos, dt, av, lv)        # plc_point2_st.py:244-251
        got = pts @ A[:, :3].T + A[:, 3]
        assert np.allclose(got, want, rtol=0, atol=1e-12)
        assert np.allclose(cloud.rotation_zyx(*rpy), oracle.hostref.body_to_earth_frame(*rpy), rtol=0, atol=1e-15)


def test_rotation_matches_reference_utils_golden(hostfn_golden):
    """rpy -> R pairs produced by the UNMODIFIED scripts/utils.py:21-28 (tests/golden/make_golden.py)."""
    for rpy, R in zip(hostfn_golden["rpy"], hostfn_golden["R"]):
        assert np.allclose(cloud.rotation_zyx(*rpy), R, rtol=0, atol=1e-15)


def test_pointcloud2_roundtrip(oracle):
    rng = np.random.default_rng(1)
    pts = rng.normal(size=(1000, 3))
    msg = cloud.xyz_to_pointcloud2_fields(pts)
    ref = oracle.hostref.pack_pointcloud2(pts)                                   # plc_point2_st.py:112-138
    assert msg == ref
    back = cloud.pointcloud2_to_xyz(msg["data"])
    assert back.dtype == np.float32 and np.array_equal(back, pts.astype(np.float32))
    # a 16-byte point step with xyz at 0/4/8 (e.g. an extra intensity field)
    rec = np.zeros((1000, 4), dtype="<f4"); rec[:, :3] = pts
    assert np.array_equal(cloud.pointcloud2_to_xyz(rec.tobytes(), point_step=16), pts.astype(np.float32))


def test_occupancy_grid_codec(oracle):
    rng = np.random.default_rng(2)
    w, h = 37, 21
    data = rng.choice(np.array([-1, 0, 100, 50, 1], dtype=np.int8), size=w * h)
    a = planner.occupancy_grid_to_array(data, w, h)
    assert np.array_equal(a, oracle.hostref.decode_occupancy_grid(data, w, h))   # global_planner_st.py:15-20
    assert a.shape == (w, h)
    assert np.array_equal(planner.array_to_occupancy_grid(a), oracle.hostref.encode_occupancy_grid(a))


@pytest.mark.parametrize("variant", ["st", "ccst"])
def test_plan_assembly_matches_reference_blocks(oracle, variant):
    rng = np.random.default_rng(3)
    for _ in range(400):
        W0, H0 = (int(v) for v in rng.integers(5, 80, 2))
        m = (rng.random((W0, H0)) < 0.2).astype(np.int64)
        o, reso, ifa = rng.uniform(-20, 20, 2), 0.2, int(rng.integers(1, 4))
        s = o + np.array([rng.uniform(-4, W0 * reso + 4), rng.uniform(-4, H0 * reso + 4)])
        g = o + np.array([rng.uniform(-4, W0 * reso + 4), rng.uniform(-4, H0 * reso + 4)])
        ref_grid, ref_s, ref_g, ref_o, ref_d = oracle.hostref.assemble_grid(m, o, reso, s, g, ifa, variant)
        a = planner.plan_assembly(m.shape, o, reso, s, g, ifa, variant)
        assert a.shape == ref_grid.shape
        assert a.start == tuple(int(v) for v in ref_s) and a.goal == tuple(int(v) for v in ref_g)
        assert a.paste_at == tuple(int(v) for v in ref_d)
        assert np.allclose(a.origin, ref_o, rtol=0, atol=1e-12)


def test_relocate_goal_and_path_to_world(oracle):
    rng = np.random.default_rng(4)
    for _ in range(300):
        m = (rng.random((30, 20)) < 0.5).astype(np.float64)
        if rng.random() < 0.2:
            m[int(rng.integers(30)), :] = 1                      # a full row: forces the column fallback
        g = (int(rng.integers(30)), int(rng.integers(20)))
        if m[g[0], g[1]] == 1 and not (m[g[0], :] == 0).any() and not (m[:, g[1]] == 0).any():
            continue
        want, occ = oracle.hostref.relocate_goal(m, g)           # global_planner_st.py:268-275
        got, moved = planner.relocate_goal(m, g)
        assert got == tuple(int(v) for v in want) and moved == bool(occ)
    path = [(3, 4), (7, 8), (7, 20)]
    for v in ("st", "ccst"):
        assert np.array_equal(planner.path_cells_to_world(path, 0.2, (-3.0, 5.0), v),
                              oracle.hostref.path_to_world(path, 0.2, np.array([-3.0, 5.0]), v))


def test_formats_host_side():
    from fuxi_planner_b200 import formats
    assert formats.map_png_name((-16.4, -4.8)) == "-16.40-4.80_out.png"
    assert formats.map_png_name((4.6, 1.0)) == "4.601.00_out.png"
    msg = formats.path_message([(1.0, 2.0, 0.0), (3.5, 4.5, 0.0)], stamp=12.5)
    assert msg["header"]["frame_id"] == "map" and len(msg["poses"]) == 2
    assert msg["poses"][1]["pose"]["position"] == {"x": 3.5, "y": 4.5, "z": 0.0}
    assert msg["poses"][0]["header"] == {"frame_id": "map", "stamp": 12.5}


def test_waypoint_st_matches_the_restated_block(oracle, golden, maps):
    """planner.waypoint_st against the line-by-line restatement of global_planner_st.py:291-325 on real paths (the
    reference's own jump-point paths from the golden file) and on random walks, with and without a previous waypoint."""
    from fuxi_planner_b200 import planner
    rng = np.random.default_rng(5)
    paths = [r["path"] for r in golden["cfg1"] if r.get("path") and len(r["path"]) >= 2][:120]
    for _ in range(80):
        n = int(rng.integers(2, 12))
        paths.append(np.cumsum(rng.integers(-6, 7, size=(n, 2)), axis=0) + 60)
    kept = 0
    for i, p in enumerate(paths):
        p = np.asarray(p)
        start = p[0] + rng.integers(-1, 2, size=2)
        origin = np.array([-16.4, -4.8]) + rng.uniform(-1, 1, 2)
        pos = np.r_[start * 0.2 + origin + rng.uniform(-0.3, 0.3, 2), 1.2]
        goal = np.r_[p[-1] * 0.2 + origin, 1.0]
        prev = None if i % 3 else np.r_[rng.uniform(-5, 5, 2), 0.0]
        for end_occu in (0, 1):
            want = oracle.hostref.waypoint_st(p, start, 0.2, origin, goal, pos[0], pos[1], pos[2], end_occu, wp=prev)
            got = planner.waypoint_st(p, start, 0.2, origin, goal, pos, end_occu, prev_wp=prev)
            assert np.array_equal(np.asarray(got[0], dtype=np.float64), np.asarray(want[0], dtype=np.float64)), i
            assert np.array_equal(np.asarray(got[1], dtype=np.float64), np.asarray(want[1], dtype=np.float64))
            assert got[2] == want[2]
            kept += int(not np.array_equal(np.asarray(got[0]), goal) and end_occu == 0)
    assert kept > 10          # the interesting branch (an intermediate waypoint) is exercised


def test_inline_blocks_restatement_matches_reference_lines(oracle, inline_golden):
    """a10/a11/a13/a14: oracle/hostref.py's restatement of the planner loops' inline blocks against the outputs of the
    reference's OWN source lines (global_planner_st.py:226-275, global_planner_ccst.py:411-464, exec'd by
    tests/golden/make_inline_golden.py): padded + inflated grid bit-exact, start / goal cells after relocation, map_d,
    shifted origin (float64 bit-exact), end_occu."""
    n = 0
    for c in inline_golden:
        want = c["out"]
        assert want is not None
        got = oracle.hostref.replan_pipeline(c["X"].astype(np.int64), c["map_o"], c["reso"], c["start"], c["goal"], c["ifa"], c["variant"])
        assert got["grid"].shape == tuple(want["shape"])
        assert np.array_equal(got["grid"].astype(np.uint8), want["grid"])
        assert list(got["start"]) == want["map_start"] and list(got["goal"]) == want["map_goal"]
        assert list(got["map_d"]) == want["map_d"] and got["origin"] == want["map_o"]
        assert got["end_occu"] == want["end_occu"]
        # the inflation restatements on their own (a10 / a11): padded grid before inflation -> grid after
        pad, _, _, _, _ = oracle.hostref.assemble_grid(c["X"].astype(np.int64), c["map_o"], c["reso"], c["start"], c["goal"], c["ifa"], c["variant"])
        inf = oracle.hostref.inflate_st(pad, c["ifa"]) if c["variant"] == "st" else oracle.hostref.inflate_ccst(pad, c["ifa"])
        assert np.array_equal(inf.astype(np.uint8), want["grid"])
        assert np.array_equal(oracle.inflate((pad > 0).astype(np.uint8), c["ifa"], c["ifa"] if c["variant"] == "st" else 1), want["grid"])
        n += 1
    assert n >= 150


def test_plan_assembly_matches_reference_lines(inline_golden):
    """planner.plan_assembly (the product's host arithmetic for a13) against the reference's own lines."""
    for c in inline_golden:
        want = c["out"]
        a = planner.plan_assembly((c["W"], c["H"]), c["map_o"], c["reso"], c["start"], c["goal"], c["ifa"], c["variant"])
        assert a.shape == tuple(want["shape"]) and a.paste_at == tuple(want["map_d"])
        assert list(a.origin) == want["map_o"]
        # a.start / a.goal are the cells BEFORE relocation: goal0 + map_d (- 1 for st)
        off = -1 if c["variant"] == "st" else 0
        assert list(a.goal) == [want["map_goal0"][k] + want["map_d"][k] + off for k in (0, 1)]
        assert list(a.start) == want["map_start"]


def test_node_cloud_restatement_vs_reference_lines(oracle, cloud_golden):
    """a15 -> a16 -> a17 as the node runs them (plc_point2_st.py:244-256, 336-339, 360-362): the restatement in
    oracle/hostref.py against the outputs of the reference's own lines.  The reference multiplies through np.matmul (BLAS:
    summation order unspecified), the restatement with an explicit order: same points, same order, coordinates to 4 ulps;
    the octomap-centres branch has no product in it and is bit-exact."""
    worst = 0.0
    for c in cloud_golden:
        got = oracle.hostref.node_cloud(c["cam"], c["rpy"], c["pos"], c["dt"], c["ang_vel"], c["line_vel"], c["local_pos"])
        assert got.shape == c["out"].shape
        err = np.abs(got - c["out"]) / np.maximum(np.abs(c["out"]), 1.0)
        worst = max(worst, float(err.max()) if err.size else 0.0)
        assert np.array_equal(oracle.hostref.octomap_local(c["cen"], c["local_pos"]), c["octo"])
    assert worst <= 4 * np.finfo(np.float64).eps, worst


def test_rotation_elementwise_has_the_reference_bits(hostfn_golden):
    from fuxi_planner_b200 import cloud
    for a, R in zip(hostfn_golden["rpy"], hostfn_golden["R"]):
        assert np.array_equal(cloud.rotation_elementwise(*a), R)
