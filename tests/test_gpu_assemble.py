"""GPU parity tests for the rows either side of the search: OccupancyGrid decode / encode, crop
bounding box, paste, goal relocation, path post-processing (near-vehicle drop, line-of-sight shortcutting, world
coordinates) and the fused fx_replan_host call -- all through the C ABI, against the numpy restatements in
oracle/hostref.py and the golden vectors produced by the unmodified reference (shortcut_golden.json,
crop_golden.json).  Everything here is bit-exact (integer / byte work and one-rounding-per-op float64)."""
import base64
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from util import validate_path

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as ge
    ge.build()
    import fuxi_planner_b200 as fx
    return fx


@pytest.fixture(scope="module")
def dev():
    import torch
    return torch.device("cuda:0")


def _t(a, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _b64(s, dt, shape):
    return np.frombuffer(base64.b64decode(s), dtype=dt).reshape(shape)


@pytest.fixture(scope="module")
def crop_golden():
    return json.load(open(os.path.join(GOLDEN, "crop_golden.json")))


@pytest.fixture(scope="module")
def shortcut_golden():
    return json.load(open(os.path.join(GOLDEN, "shortcut_golden.json")))


# ------------------------------------------------------------------------------------------ decode / encode / paste / bbox
def test_decode_vs_reference_map_callback(fx, dev, crop_golden, oracle):
    for r in crop_golden["decode"]:
        data = _b64(r["data"], np.int8, (-1,))
        got = fx.grid_decode(_t(data, dev), r["width"], r["height"]).cpu().numpy()
        assert np.array_equal(got, _b64(r["map"], np.uint8, r["shape"]))
        back = fx.grid_encode(_t(got, dev)).cpu().numpy()
        assert np.array_equal(back, oracle.hostref.encode_occupancy_grid(got))


@pytest.mark.parametrize("w,h", [(1, 1), (31, 33), (257, 64), (1000, 1237), (4096, 4096)])
def test_decode_window_paste_and_roundtrip(fx, dev, oracle, w, h):
    import torch
    rng = np.random.default_rng(w * 7 + h)
    data = rng.choice(np.array([-1, 0, 0, 100, 100, 1, 50, 99], dtype=np.int8), w * h)
    want = oracle.hostref.decode_occupancy_grid(data, w, h).astype(np.uint8)
    d = _t(data, dev)
    got = fx.grid_decode(d, w, h)
    assert np.array_equal(got.cpu().numpy(), want)
    # encode(decode(x)) == x with -1 -> 0 (and 1 -> 100): the round trip the planners do (map_callback -> publish_map)
    rt = fx.grid_encode(got).cpu().numpy()
    exp = data.copy(); exp[exp == -1] = 0; exp[exp == 1] = 100
    assert np.array_equal(rt, exp)
    # window + paste offset == numpy slice assignment into a zero array
    x0, y0 = w // 3, h // 4
    ww, hh = max(1, w // 2), max(1, h // 2)
    ww, hh = min(ww, w - x0), min(hh, h - y0)
    out = torch.zeros((ww + 5, hh + 9), dtype=torch.uint8, device=dev)
    fx.grid_decode(d, w, h, out=out, window=(x0, y0, ww, hh), paste_at=(3, 2))
    ref = np.zeros((ww + 5, hh + 9), dtype=np.uint8)
    ref[3:3 + ww, 2:2 + hh] = want[x0:x0 + ww, y0:y0 + hh]
    assert np.array_equal(out.cpu().numpy(), ref)
    # array -> array paste, overwriting
    dst = torch.full((w + 4, h + 6), 7, dtype=torch.uint8, device=dev)
    fx.grid_paste(got, dst, paste_at=(2, 5))
    ref2 = np.full((w + 4, h + 6), 7, dtype=np.uint8)
    ref2[2:2 + w, 5:5 + h] = want
    assert np.array_equal(dst.cpu().numpy(), ref2)


def test_bbox_vs_numpy(fx, dev):
    rng = np.random.default_rng(5)
    for i in range(30):
        W, H = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        a = (rng.random((W, H)) < [0.0, 0.001, 0.02, 0.5][i % 4]).astype(np.uint8) * rng.integers(1, 100, (W, H)).astype(np.uint8)
        got = fx.grid_bbox(_t(a, dev)).cpu().numpy().tolist()
        xs, ys = a.nonzero()
        want = [int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())] if len(xs) else [2**31 - 1, -1, 2**31 - 1, -1]
        assert got == want
        # the same cells as a raw message (100 = occupied, -1 = unknown counts as zero)
        msg = np.where(a.T == 1, 100, a.T).astype(np.int8).reshape(-1)
        msg[(msg == 0) & (rng.random(msg.size) < 0.3)] = -1
        assert fx.grid_bbox(_t(msg, dev), width=W, height=H).cpu().numpy().tolist() == want


# ------------------------------------------------------------------------------------------ goal relocation
def test_relocate_goal_vs_restatement(fx, dev, oracle):
    rng = np.random.default_rng(8)
    n_moved = n_col = 0
    for i in range(120):
        W, H = int(rng.integers(3, 80)), int(rng.integers(3, 80))
        m = (rng.random((W, H)) < rng.choice([0.1, 0.5, 0.9])).astype(np.uint8)
        gx, gy = int(rng.integers(W)), int(rng.integers(H))
        if i % 3 == 0:
            m[gx, gy] = 1
        if i % 6 == 0:
            m[gx, :] = 1          # no free cell in the row: the reference's except branch scans the column
            m[int(rng.integers(W)), gy] = 0
        ifa = int(rng.integers(1, 4))
        has_free = (m[gx, :] == 0).any() or (m[:, gy] == 0).any()
        got = fx.relocate_goal(_t(m, dev), (gx, gy), ifa=ifa, variant="st").cpu().numpy().tolist()
        got_cc = fx.relocate_goal(_t(m, dev), (gx, gy), ifa=ifa, variant="ccst").cpu().numpy().tolist()
        if m[gx, gy] == 1 and not has_free:
            assert got[2] == -1      # the reference raises here
            continue
        want, occ = oracle.hostref.relocate_goal(m.astype(np.float64), np.array([gx, gy]))
        assert got[:3] == [int(want[0]), int(want[1]), occ] and got[3] == occ
        assert got_cc[:3] == got[:3]
        wx, wy = int(want[0]), int(want[1])
        assert got_cc[3] == int((m[wx - ifa:wx + ifa, wy - ifa:wy + ifa] == 1).any())     # ccst:460-463, numpy slice semantics
        n_moved += occ
        n_col += int(occ and want[0] != gx)
    assert n_moved > 20 and n_col > 3


# ------------------------------------------------------------------------------------------ path post-processing
def _run_post(fx, dev, grid, paths, **kw):
    import torch
    mp = max(len(p) for p in paths)
    xy = np.zeros((len(paths), mp, 2), dtype=np.int32)
    ln = np.zeros(len(paths), dtype=np.int32)
    for i, p in enumerate(paths):
        xy[i, :len(p)] = p
        ln[i] = len(p)
    oxy, oln, ow = fx.path_post(_t(grid, dev), _t(xy, dev), _t(ln, dev), **kw)
    torch.cuda.synchronize()
    oxy, oln = oxy.cpu().numpy(), oln.cpu().numpy()
    return [oxy[i, :oln[i]].tolist() for i in range(len(paths))], (ow.cpu().numpy() if ow is not None else None), oln


def test_shortcut_vs_reference_golden(fx, dev, maps, shortcut_golden):
    """Every vector of shortcut_golden.json (made with the reference's own map_line_col): batched per map."""
    n = 0
    for name, rows in shortcut_golden["maps"].items():
        got, _, _ = _run_post(fx, dev, maps[name].astype(np.uint8), [r["path"] for r in rows], shortcut=True)
        for r, g in zip(rows, got):
            assert g == r["out"], (name, r["path"])
            n += 1
    for r in shortcut_golden["random"]:
        m = np.unpackbits(np.array(r["grid"], dtype=np.uint8))[:r["W"] * r["H"]].reshape(r["W"], r["H"])
        got, _, _ = _run_post(fx, dev, m, [r["path"]], shortcut=True)
        assert got[0] == r["out"], r["path"]
        n += 1
    assert n > 700


def test_shortcut_vs_restatement_long_lines(fx, dev, oracle):
    """Long segments on a 1024^2 grid (more than one 32-sample round per line), random polylines, both slopes."""
    rng = np.random.default_rng(12)
    m = (rng.random((1024, 1024)) < 0.002).astype(np.uint8)
    paths = []
    for i in range(64):
        k = int(rng.integers(3, 40))
        paths.append(np.c_[rng.integers(0, 1024, k), rng.integers(0, 1024, k)].tolist())
    got, _, _ = _run_post(fx, dev, m, paths, shortcut=True)
    for p, g in zip(paths, got):
        assert g == oracle.hostref.shortcut_path(p, m.astype(np.float64))
        assert g[0] == p[0] and g[-1] == p[-1]


def test_near_drop_and_world_vs_restatement(fx, dev, oracle):
    rng = np.random.default_rng(13)
    m = np.zeros((200, 200), dtype=np.uint8)
    paths, wants_c, wants_w = [], [], []
    reso, origin = 0.2, (-7.3, 4.1)
    pos = (2.5, 14.0, 0.3)
    for i in range(50):
        k = int(rng.integers(1, 30))
        p = np.c_[rng.integers(35, 65, k), rng.integers(35, 65, k)].tolist()
        w = oracle.hostref.path_to_world(p, reso, np.array(origin), "ccst")
        c2, w2 = oracle.hostref.near_drop(p, w, pos, 1.5)
        paths.append(p); wants_c.append(np.asarray(c2).reshape(-1, 2).tolist()); wants_w.append(w2)
    got, world, oln = _run_post(fx, dev, m, paths, shortcut=False, drop=(*pos, 1.5), world=(reso, origin[0], origin[1], 1, 0))
    dropped = 0
    for i, (g, wc, ww) in enumerate(zip(got, wants_c, wants_w)):
        assert g == wc
        assert np.array_equal(world[i, :oln[i]], np.asarray(ww).reshape(-1, 3))     # bit-exact float64
        dropped += len(paths[i]) - len(g)
    assert dropped > 20
    # st offsets, no drop
    got, world, oln = _run_post(fx, dev, m, paths, shortcut=False, world=(reso, origin[0], origin[1], 1, 1))
    for i, p in enumerate(paths):
        assert np.array_equal(world[i, :oln[i]], oracle.hostref.path_to_world(p, reso, np.array(origin), "st"))


def test_path_post_passes_failures_through(fx, dev):
    m = np.zeros((16, 16), dtype=np.uint8)
    import torch
    xy = torch.zeros((3, 8, 2), dtype=torch.int32, device=dev)
    ln = torch.tensor([-1, 0, -2], dtype=torch.int32, device=dev)
    _, oln, _ = fx.path_post(_t(m, dev), xy, ln, shortcut=True)
    assert oln.cpu().numpy().tolist() == [-1, 0, -2]


# ------------------------------------------------------------------------------------------ fused replan
def _check_replan(fx, oracle, mapu, origin, reso, start, goal, ifa, variant, crop=False, hchoice=2):
    """fx_replan_host against hostref.replan_pipeline + the oracle's cost on the pipeline's grid."""
    W0, H0 = mapu.shape
    msg = oracle.hostref.encode_occupancy_grid(mapu)
    want = oracle.hostref.replan_pipeline(mapu, origin, reso, start, goal, ifa, variant, crop=crop)
    out, cells, world, grid = fx.replan_host(msg, W0, H0, origin, reso, start, goal, ifa=ifa, variant=variant, hchoice=hchoice,
                                             crop=crop, shortcut=False, want_grid=True)
    if want is None:
        assert out.skipped == 2
        return None
    assert (out.W, out.H) == want["grid"].shape
    assert np.array_equal(grid, want["grid"].astype(np.uint8))
    assert (out.paste_x, out.paste_y) == want["map_d"]
    assert [out.origin_x, out.origin_y] == want["origin"]         # bit-exact float64
    assert (out.start_x, out.start_y) == want["start"]
    assert (out.goal_x, out.goal_y) == want["goal"] and out.goal_moved == want["moved"] and out.end_occu == want["end_occu"]
    assert bool(out.skipped) == want["skipped"]
    if want["skipped"]:
        return None
    g = want["grid"]
    s, t = want["start"], want["goal"]
    if not (0 <= s[0] < g.shape[0] and 0 <= s[1] < g.shape[1]):
        assert out.path_len == -2
        return None
    ref = oracle.capi.jps(g, s, t, hchoice)
    if ref[0] == 0:
        assert out.path_len == -1
        return None
    if hchoice == 1:
        assert out.cost_i == int(ref[1])
    else:
        assert abs(out.cost_f - ref[1]) <= 1e-5 * max(ref[1], 1.0)       # north-star tolerance
    path = [tuple(p) for p in cells.tolist()]
    validate_path((g == 1).astype(np.uint8), path, s, t)
    assert np.array_equal(world, oracle.hostref.path_to_world(path, reso, np.array(want["origin"]), variant))
    # post-processing on the same raw path: the fused call with shortcut + drop == restatement applied to `path`
    pos = (start[0], start[1], 1.0)
    out2, cells2, world2, _ = fx.replan_host(msg, W0, H0, origin, reso, start, goal, ifa=ifa, variant=variant, hchoice=hchoice,
                                             crop=crop, shortcut=True, drop=(*pos, 1.5))
    c2, w2 = oracle.hostref.near_drop(path, oracle.hostref.path_to_world(path, reso, np.array(want["origin"]), variant), pos, 1.5)
    sc = oracle.hostref.shortcut_path(np.asarray(c2).reshape(-1, 2).tolist(), g) if len(c2) else []
    assert cells2.tolist() == [list(p) for p in sc]
    assert out2.raw_len == len(path)
    if len(sc):
        assert np.array_equal(world2, oracle.hostref.path_to_world(sc, reso, np.array(want["origin"]), variant))
    return out


@pytest.mark.parametrize("variant,ifa", [("st", 1), ("ccst", 2), ("st", 3)])
def test_replan_repo_maps(fx, oracle, maps, variant, ifa):
    rng = np.random.default_rng(17)
    n = 0
    for name in sorted(maps)[::3]:
        m = maps[name].astype(np.int64)
        W0, H0 = m.shape
        reso, origin = 0.2, (-16.4, -4.8)
        for k in range(4):
            # world positions inside, and occasionally outside (negative index / beyond the map) like a vehicle near the edge
            lo, hi = (-0.2, 1.2) if k == 3 else (0.0, 1.0)
            start = (origin[0] + rng.uniform(lo, hi) * W0 * reso, origin[1] + rng.uniform(lo, hi) * H0 * reso)
            goal = (origin[0] + rng.uniform(lo, hi) * W0 * reso, origin[1] + rng.uniform(lo, hi) * H0 * reso)
            r = _check_replan(fx, oracle, m, origin, reso, start, goal, ifa, variant, crop=(variant == "ccst" and k % 2 == 0))
            n += r is not None
    assert n > 10


def test_replan_random_grids_with_values_and_crop(fx, oracle):
    rng = np.random.default_rng(19)
    n = 0
    for i in range(40):
        W0, H0 = int(rng.integers(8, 120)), int(rng.integers(8, 120))
        m = (rng.random((W0, H0)) < rng.choice([0.0, 0.02, 0.1])).astype(np.int64) * rng.choice([1, 1, 1, 40, 99], (W0, H0))
        if i % 5 == 1:
            m[: W0 // 2] = 0     # occupied cells only in one half: the crop moves the origin
        reso = float(rng.choice([0.1, 0.2]))
        origin = tuple(rng.uniform(-5, 5, 2).round(1))
        start = (origin[0] + rng.uniform(0, 1) * W0 * reso, origin[1] + rng.uniform(0, 1) * H0 * reso)
        goal = (origin[0] + rng.uniform(-0.1, 1.3) * W0 * reso, origin[1] + rng.uniform(-0.1, 1.3) * H0 * reso)
        variant = "ccst" if i % 2 else "st"
        r = _check_replan(fx, oracle, m, origin, reso, start, goal, int(rng.integers(1, 3)), variant, crop=(variant == "ccst"),
                          hchoice=1 if i % 4 == 0 else 2)
        n += r is not None
    assert n > 15


def test_replan_array_layout_equals_message_layout(fx, oracle, maps):
    m = maps["-16.40-4.80_out.png"].astype(np.uint8)
    msg = oracle.hostref.encode_occupancy_grid(m)
    a = fx.replan_host(msg, m.shape[0], m.shape[1], (0.0, 0.0), 0.2, (0.5, 0.5), (28.0, 9.0), ifa=1, variant="st", want_grid=True)
    b = fx.replan_host(m, m.shape[0], m.shape[1], (0.0, 0.0), 0.2, (0.5, 0.5), (28.0, 9.0), ifa=1, variant="st", layout="array", want_grid=True)
    assert np.array_equal(a[3], b[3]) and a[1].tolist() == b[1].tolist() and a[0].cost_f == b[0].cost_f and a[0].path_len > 1


# ------------------------------------------------------------------------------------------ reference-executed goldens
def test_replan_vs_reference_inline_lines(fx, dev, oracle, inline_golden):
    """a10/a11/a13/a14 on the device against the outputs of the reference's OWN source lines
    (global_planner_st.py:226-275 / global_planner_ccst.py:411-464, exec'd by tests/golden/make_inline_golden.py):
    fx_replan_host's planning grid (pad/shift + inflation) bit-exact, start / goal cells after relocation, map_d,
    shifted origin bit-exact, end_occu; fx_inflate and fx_relocate_goal alone on the same data."""
    n = n_moved = 0
    for c in inline_golden:
        want = c["out"]
        X = c["X"]
        msg = oracle.hostref.encode_occupancy_grid(X.astype(np.int64))
        out, cells, world, grid = fx.replan_host(msg, c["W"], c["H"], tuple(c["map_o"]), c["reso"], tuple(c["start"]), tuple(c["goal"]),
                                                 ifa=c["ifa"], variant=c["variant"], hchoice=2, shortcut=False, want_grid=True)
        assert (out.W, out.H) == tuple(want["shape"])
        assert np.array_equal(grid, want["grid"])
        assert [out.paste_x, out.paste_y] == want["map_d"]
        assert [out.origin_x, out.origin_y] == want["map_o"]
        assert [out.start_x, out.start_y] == want["map_start"]
        assert [out.goal_x, out.goal_y] == want["map_goal"]
        assert out.end_occu == want["end_occu"]
        # a10 / a11 alone: fx_inflate of the padded grid == the reference's inflated grid
        pad = np.zeros(tuple(want["shape"]), dtype=np.uint8)
        d = want["map_d"]
        pad[d[0]:d[0] + c["W"], d[1]:d[1] + c["H"]] = X
        step = "st" if c["variant"] == "st" else "ccst"
        assert np.array_equal(fx.inflate(_t(pad, dev), c["ifa"], step).cpu().numpy(), want["grid"])
        # a14 alone: fx_relocate_goal from the unrelocated goal cell
        off = -1 if c["variant"] == "st" else 0
        g0 = [want["map_goal0"][k] + d[k] + off for k in (0, 1)]
        rel = fx.relocate_goal(_t(want["grid"], dev), tuple(g0), ifa=c["ifa"], variant=c["variant"]).cpu().numpy().tolist()
        assert rel[:2] == want["map_goal"] and rel[3] == want["end_occu"]
        n_moved += int(rel[2] == 1)
        n += 1
    assert n >= 150 and n_moved > 20
