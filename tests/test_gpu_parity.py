"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed golden
vectors produced by the unmodified reference.  Bit-exact for grids, masks, fields and integer costs;
1e-5 relative for Euclidean path lengths (the north-star tolerance)."""
import numpy as np
import pytest

from conftest import unpack_grid, large_grid
from util import validate_path, random_queries

pytestmark = pytest.mark.gpu
SQRT2 = 1.4142135623730951
RTOL = 1e-5   # north-star tolerance for float path lengths


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


# ------------------------------------------------------------------------------------------ projection
@pytest.mark.parametrize("stride,n", [(3, 100003), (4, 100003), (3, 7), (4, 1), (3, 1 << 20)])
def test_project_bit_exact(fx, dev, oracle, stride, n):
    rng = np.random.default_rng(stride * 1000 + n % 97)
    pts = np.zeros((n, stride), dtype=np.float32)
    pts[:, 0:2] = rng.uniform(-110, 110, (n, 2))
    pts[:, 2] = rng.uniform(-0.5, 3.0, n)
    if n > 100:
        pts[5] = np.nan
        pts[6, 0] = np.inf
        pts[7, :3] = (-102.4, -102.4, 0.3)       # exactly on the low edges / z threshold
        pts[8, :3] = (102.4, 0.0, 1.0)           # exactly on the high edge -> outside
        pts[9:40, 0] = (np.arange(31) * 0.2 - 3.0).astype(np.float32)   # on cell boundaries
    A = oracle.hostref.cloud_affine((0.03, -0.02, 0.7), (1.0, -2.0, 0.4), dt=0.01, ang_vel=(0.1, 0.2, -0.1), line_vel=(1, 0, 0))
    for affine in (None, A):
        want = oracle.hostref.project(pts[:, :3], np.eye(3, 4) if affine is None else affine, 0.3, np.inf, -102.4, -102.4, 0.2, 1024, 1024)
        got = fx.project(_t(pts, dev), affine, 0.3, np.inf, (-102.4, -102.4), 0.2, (1024, 1024))
        assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("stride", [3, 4])
def test_project_large_grid_bit_form(fx, dev, oracle, stride):
    """Grids beyond L2 size take the bit form (RED.OR into an L2-resident bit-packed grid + streaming expand): same
    grid as the oracle bit for bit, with a cell count that is not a multiple of the expand kernel's 16-cell step and a
    cloud concentrated in a small patch (heavy same-word contention)."""
    rng = np.random.default_rng(stride)
    W, H, reso = 6001, 5599, 0.2                      # 33.6 M cells > 32 Mi, odd cell count
    n = (4 << 20) + 37
    pts = np.zeros((n, stride), dtype=np.float32)
    pts[:, 0] = rng.uniform(-5.0, W * reso + 5.0, n)
    pts[:, 1] = rng.uniform(-5.0, H * reso + 5.0, n)
    pts[:, 2] = rng.uniform(-0.5, 3.0, n)
    pts[11] = np.nan
    want = oracle.hostref.project(pts[:, :3], np.eye(3, 4), 0.3, np.inf, 0.0, 0.0, reso, W, H)
    got = fx.project(_t(pts, dev), None, 0.3, np.inf, (0.0, 0.0), reso, (W, H))
    assert np.array_equal(got.cpu().numpy(), want)
    pts[:, 0] = rng.uniform(600.0, 640.0, n)
    pts[:, 1] = rng.uniform(500.0, 540.0, n)
    want = oracle.hostref.project(pts[:, :3], np.eye(3, 4), 0.3, np.inf, 0.0, 0.0, reso, W, H)
    got = fx.project(_t(pts, dev), None, 0.3, np.inf, (0.0, 0.0), reso, (W, H))
    assert np.array_equal(got.cpu().numpy(), want)
    assert 199 * 199 <= int(want.sum()) <= 202 * 202


def test_project_unaligned_and_accumulate(fx, dev, oracle):
    import torch
    rng = np.random.default_rng(3)
    pts = rng.uniform(-10, 10, (1001, 3)).astype(np.float32)
    buf = torch.zeros(1001 * 3 + 1, dtype=torch.float32, device=dev)
    buf[1:] = _t(pts.reshape(-1), dev)
    view = buf[1:].view(1001, 3)            # 4-byte aligned only -> generic kernel
    want = oracle.hostref.project(pts, np.eye(3, 4), -1.0, 5.0, -10, -10, 0.5, 40, 56)
    got = fx.project(view, None, -1.0, 5.0, (-10, -10), 0.5, (40, 56))
    assert np.array_equal(got.cpu().numpy(), want)
    pts2 = rng.uniform(-10, 10, (500, 3)).astype(np.float32)
    fx.project(_t(pts2, dev), None, -1.0, 5.0, (-10, -10), 0.5, out=got, clear=False)
    want2 = want | oracle.hostref.project(pts2, np.eye(3, 4), -1.0, 5.0, -10, -10, 0.5, 40, 56)
    assert np.array_equal(got.cpu().numpy(), want2)


# ------------------------------------------------------------------------------------------ inflation
@pytest.mark.parametrize("shape", [(1024, 1024), (300, 528), (148, 52), (1, 16), (17, 1), (2048, 4096)])
@pytest.mark.parametrize("radius,variant", [(1, "st"), (2, "ccst"), (3, "st"), (5, "ccst"), (16, "ccst"), (20, "ccst"), (0, "ccst")])
def test_inflate_bit_exact(fx, dev, oracle, shape, radius, variant):
    if shape == (2048, 4096) and radius not in (2, 3):
        pytest.skip("large case only for the reference radii")
    rng = np.random.default_rng(radius * 7 + shape[0])
    m = ((rng.random(shape) < 0.03) * rng.integers(1, 101, shape)).astype(np.uint8)   # occupancy values 1..100 count (> 0)
    step = 1 if variant == "ccst" else max(radius, 1)
    want = oracle.inflate(m, radius, step)
    got = fx.inflate(_t(m, dev), radius, variant)
    assert np.array_equal(got.cpu().numpy(), want)


def test_inflate_matches_reference_formulation_on_padded_map(fx, dev, oracle, maps):
    # the exact numpy fancy-index formulation of the planners, on a map padded as global_planner_st.py:230-250 does
    for ifa, fn, variant in ((1, oracle.hostref.inflate_st, "st"), (2, oracle.hostref.inflate_ccst, "ccst"), (3, oracle.hostref.inflate_st, "st")):
        core = maps["-16.20-11.40_out.png"]
        m = np.zeros((core.shape[0] + 6 * ifa, core.shape[1] + 6 * ifa))
        m[2 * ifa:2 * ifa + core.shape[0], 2 * ifa:2 * ifa + core.shape[1]] = core * 100      # OccupancyGrid-style 100
        want = fn(m, ifa)
        got = fx.inflate(_t((m > 0).astype(np.uint8) * 100, dev), ifa, variant)
        assert np.array_equal(got.cpu().numpy().astype(np.float64), want)


# ------------------------------------------------------------------------------------------ EDT
@pytest.mark.parametrize("shape,fill", [((512, 512), 0.2), ((300, 272), 0.001), ((64, 48), 0.0), ((1000, 37), 0.01),
                                        ((1, 100), 0.05), ((2048, 1024), 0.00001),
                                        # H % 32 == 0: the bit-parallel path (dense: all resolved in-tile; 0.5-2 %: fix-up
                                        # list; sparse tiles: falls back to the windowed path on the device flag)
                                        ((1024, 1024), 0.02), ((256, 64), 0.05), ((1100, 2048), 0.005), ((96, 32), 0.3),
                                        ((130, 96), 1.0), ((2048, 4096), 0.02), ((777, 1056), 0.0004),
                                        # warp strips: fewer rows than one 64-row step, one word column, ragged last step
                                        ((1, 32), 0.5), ((2, 64), 0.1), ((63, 32), 0.03), ((65, 64), 0.03), ((129, 288), 0.01)])
def test_edt_exact(fx, dev, oracle, shape, fill):
    rng = np.random.default_rng(int(fill * 1e6) + shape[0])
    m = (rng.random(shape) < fill).astype(np.uint8)
    want = oracle.edt(m)
    got = fx.edt(_t(m, dev)).cpu().numpy()
    assert np.array_equal(got, want)


def test_edt_graph_replay_follows_the_data(fx, dev, oracle):
    """fx_edt replays a captured CUDA graph from its third call with the same buffers on: the device-side decisions (fix-up
    list, fallback to the windowed path on the overflow flag) must still follow the data of each call."""
    import torch
    rng = np.random.default_rng(5)
    shape = (640, 1024)
    occ = torch.empty(shape, dtype=torch.uint8, device=dev)
    d2 = torch.empty(shape, dtype=torch.int32, device=dev)
    for it, fill in enumerate((0.02, 0.02, 0.02, 0.0002, 0.3, 0.0, 0.004)):  # dense / sparse (list overflows) / empty / fix-up list
        m = (rng.random(shape) < fill).astype(np.uint8)
        occ.copy_(torch.from_numpy(m))
        fx.edt(occ, out=d2)
        assert np.array_equal(d2.cpu().numpy(), oracle.edt(m)), (it, fill)


def test_edt_single_obstacle_and_clusters(fx, dev, oracle):
    m = np.zeros((128, 128), dtype=np.uint8)
    m[5, 120] = 1
    assert np.array_equal(fx.edt(_t(m, dev)).cpu().numpy(), oracle.edt(m))
    # dense blobs next to large empty areas: resolved tiles, fix-up cells and (here) sparse tiles in one grid
    rng = np.random.default_rng(3)
    m = np.zeros((512, 1024), dtype=np.uint8)
    m[:200, :300] = rng.random((200, 300)) < 0.1
    m[300:330, 900:] = 1
    assert np.array_equal(fx.edt(_t(m, dev)).cpu().numpy(), oracle.edt(m))
    m[:, :] = 0
    m[::9, ::11] = 1            # every cell within sqrt(4^2 + 5^2) of an obstacle: nothing left for the fix-up list
    assert np.array_equal(fx.edt(_t(m, dev)).cpu().numpy(), oracle.edt(m))


@pytest.mark.parametrize("shape,fill", [((300, 200), 0.01), ((64, 1024), 0.2), ((257, 33), 0.001), ((128, 128), 0.0)])
def test_edt_separable_passes(fx, dev, oracle, shape, fill):
    """fx_edt_rows / fx_edt_cols (the entry points of the row-tiled mode) compose to fx_edt; the column pass also works
    on a column block on its own (what a rank holds after the transpose)."""
    from fuxi_planner_b200 import tiled
    m = (np.random.default_rng(shape[0]).random(shape) < fill).astype(np.uint8)
    want = oracle.edt(m)
    ops = tiled.CudaOps()
    g = ops.edt_rows(_t(m, dev))
    assert np.array_equal(ops.edt_cols(g).cpu().numpy(), want)
    a, b = shape[1] // 3, shape[1] // 3 + max(shape[1] // 2, 1)
    assert np.array_equal(ops.edt_cols(g[:, a:b].contiguous()).cpu().numpy(), want[:, a:b])
    assert np.array_equal(tiled.edt_tiled(_t(m, dev), shape[0]).cpu().numpy(), want)


def test_edt_vs_scipy(fx, dev):
    from scipy.ndimage import distance_transform_edt
    rng = np.random.default_rng(12)
    m = (rng.random((700, 400)) < 0.002).astype(np.uint8)
    m[0, 0] = 1
    ref = np.rint(distance_transform_edt(m == 0) ** 2).astype(np.int64)
    assert np.array_equal(fx.edt(_t(m, dev)).cpu().numpy().astype(np.int64), ref)


# ------------------------------------------------------------------------------------------ search
def _check_batch(fx, dev, oracle, m, recs, max_path=1024):
    """recs: golden records (start, goal, h, cost).  Costs: metric 1 bit-exact, metric 2 within RTOL; paths legal."""
    import torch
    for h in (1, 2):
        rs = [r for r in recs if r["h"] == h]
        if not rs:
            continue
        s = np.array([r["start"] for r in rs], dtype=np.int32)
        g = np.array([r["goal"] for r in rs], dtype=np.int32)
        res = fx.plan_batch(_t(m, dev), _t(s, dev), _t(g, dev), metric=h, max_path=max_path)
        torch.cuda.synchronize()
        ci, cf, pl = res.cost_i.cpu().numpy(), res.cost_f.cpu().numpy(), res.path_len.cpu().numpy()
        pxy = res.path_xy.cpu().numpy()
        for q, r in enumerate(rs):
            if r["cost"] is None:
                assert ci[q] == -1 and pl[q] == -1, (r, ci[q])
                continue
            want = float(r["cost"])
            if h == 1:
                assert ci[q] == int(want), (r, ci[q])
                assert cf[q] == want
            else:
                assert abs(cf[q] - want) <= RTOL * max(want, 1e-12), (r, cf[q])
            n = pl[q]
            assert 1 <= n <= max_path, (r, n)
            path = [tuple(p) for p in pxy[q, :n].tolist()]
            if r["start"] == r["goal"]:
                assert path == [tuple(r["start"])]
                continue
            a, b = validate_path(m, path, r["start"], r["goal"])
            if h == 1:
                assert 10 * a + 14 * b == ci[q]
            else:
                assert a * fx.FX_EUCLID_WS + b * fx.FX_EUCLID_WD == ci[q]
                assert cf[q] == a + b * SQRT2
            # turning points really turn
            for p0, p1, p2 in zip(path[:-2], path[1:-1], path[2:]):
                d1 = (np.sign(p1[0] - p0[0]), np.sign(p1[1] - p0[1]))
                d2 = (np.sign(p2[0] - p1[0]), np.sign(p2[1] - p1[1]))
                assert d1 != d2


def test_search_all_reference_maps(fx, dev, oracle, golden, maps, search_form):
    for name, recs in golden["maps"].items():
        _check_batch(fx, dev, oracle, maps[name], recs)


@pytest.fixture(params=["shared-memory", "batched"])
def search_form(request, monkeypatch):
    """Maps below FX_SMALL_CELLS (every map the reference ships) are searched by the one-CTA shared-memory kernel
    (small.cu); FUXI_B200_SMALL=0 sends the same maps through the batched kernel (search.cu): both are checked against
    the same golden vectors."""
    if request.param == "batched":
        monkeypatch.setenv("FUXI_B200_SMALL", "0")
    return request.param


def test_search_cfg1_300_pairs(fx, dev, oracle, golden, maps, search_form):
    _check_batch(fx, dev, oracle, maps["-16.40-4.80_out.png"], golden["cfg1"])


def test_search_edge_cases(fx, dev, oracle, golden, search_form):
    for rec in golden["edge"]:
        m = (np.array(rec["grid"]) == 1).astype(np.uint8)
        _check_batch(fx, dev, oracle, m, [rec])


def test_search_random_small(fx, dev, oracle, golden, search_form):
    for g in golden["random_small"]:
        _check_batch(fx, dev, oracle, unpack_grid(g), g["queries"])


def test_search_large_golden(fx, dev, oracle, golden):
    for g in golden["large"]:
        _check_batch(fx, dev, oracle, large_grid(g), g["queries"], max_path=4096)


def test_search_start_oob_and_goal_oob(fx, dev, search_form):
    import torch
    m = np.zeros((8, 8), dtype=np.uint8)
    s = np.array([[8, 0], [0, 0], [-1, 3]], dtype=np.int32)
    g = np.array([[1, 1], [9, 9], [2, 2]], dtype=np.int32)
    res = fx.plan_batch(_t(m, dev), _t(s, dev), _t(g, dev), metric=1)
    torch.cuda.synchronize()
    assert res.cost_i.cpu().tolist() == [-2, -1, -2]


@pytest.mark.parametrize("n,fill,Q", [(256, 0.2, 512), (512, 0.3, 256), (1024, 0.2, 256), (300, 0.42, 200)])
def test_search_vs_oracle_random(fx, dev, oracle, n, fill, Q):
    """cfg3-style: random-obstacle grid used as the final grid; GPU batch vs the restated JPS (threads on the host)."""
    import torch
    rng = np.random.default_rng(n + Q)
    m = (rng.random((n, n + 16)) < fill).astype(np.uint8)
    s, g = random_queries(m, Q, rng)
    for h in (1, 2):
        want, status, _ = oracle.jps_batch(m, s, g, h)
        res = fx.plan_batch(_t(m, dev), _t(s, dev), _t(g, dev), metric=h, max_path=8192)
        torch.cuda.synchronize()
        ci, cf, pl = res.cost_i.cpu().numpy(), res.cost_f.cpu().numpy(), res.path_len.cpu().numpy()
        pxy = res.path_xy.cpu().numpy()
        assert ((status == 1) == (ci >= 0)).all()
        ok = status == 1
        if h == 1:
            assert np.array_equal(ci[ok].astype(np.float64), want[ok])
        else:
            assert (np.abs(cf[ok] - want[ok]) <= RTOL * np.maximum(want[ok], 1e-12)).all()
        for q in np.flatnonzero(ok)[:64]:
            path = [tuple(p) for p in pxy[q, :pl[q]].tolist()]
            if tuple(s[q]) != tuple(g[q]):
                validate_path(m, path, s[q], g[q])


def test_search_maze_needs_wide_passes(fx, dev, oracle):
    """Serpentine walls: the band-limited first pass must fail and escalate; the result is still exact."""
    import torch
    n = 96
    m = np.zeros((n, n), dtype=np.uint8)
    for i, x in enumerate(range(4, n - 4, 4)):
        m[x, :] = 1
        if i % 2 == 0:
            m[x, n - 3:n - 1] = 0
        else:
            m[x, 1:3] = 0
    s = np.array([[0, 0], [0, n // 2], [n - 1, n - 1]], dtype=np.int32)
    g = np.array([[n - 1, n - 1], [n - 1, n // 2], [0, 0]], dtype=np.int32)
    for h in (1, 2):
        want, status, _ = oracle.jps_batch(m, s, g, h)
        assert (status == 1).all()
        res = fx.plan_batch(_t(m, dev), _t(s, dev), _t(g, dev), metric=h, max_path=1024)
        torch.cuda.synchronize()
        if h == 1:
            assert res.cost_i.cpu().tolist() == [int(v) for v in want]
        else:
            assert np.allclose(res.cost_f.cpu().numpy(), want, rtol=RTOL, atol=0)


# ------------------------------------------------------------------------------------------ field
def test_field_bit_exact(fx, dev, oracle, maps):
    rng = np.random.default_rng(9)
    cases = [(maps["-16.20-11.40_out.png"], (0, 0)), (maps["-16.00-8.80_out.png"], (144, 71))]
    m = (rng.random((384, 272)) < 0.25).astype(np.uint8)
    free = np.argwhere(m == 0)
    cases.append((m, tuple(free[len(free) // 2])))
    occ_src = np.argwhere(m == 1)[10]
    cases.append((m, tuple(occ_src)))           # source on an obstacle may leave it (jps1.py never tests the source)
    for grid, src in cases:
        for metric in (1, 2):
            want = oracle.sssp_field(grid, src, metric)
            got = fx.field(_t(grid, dev), src, metric).cpu().numpy()
            assert np.array_equal(got.astype(np.int64), want)


def test_field_relax_two_slabs(fx, dev, oracle):
    """Row-tiled mode with two virtual ranks on one GPU: slabs + one ghost row each, halo exchange until
    nothing changes; the stitched field must equal the single-GPU field bit for bit."""
    import torch
    rng = np.random.default_rng(21)
    W, H = 200, 144
    m = (rng.random((W, H)) < 0.3).astype(np.uint8)
    src = tuple(np.argwhere(m == 0)[5])
    want = oracle.sssp_field(m, src, 1)
    cut = 90
    # slab 0 owns rows [0,cut), holds [0,cut+1); slab 1 owns [cut,W), holds [cut-1,W)
    g0, g1 = _t(m[:cut + 1], dev), _t(m[cut - 1:], dev)
    f0 = torch.full((cut + 1, H), -1, dtype=torch.int32, device=dev)
    f1 = torch.full((W - cut + 1, H), -1, dtype=torch.int32, device=dev)
    if src[0] < cut:
        f0[src[0], src[1]] = 0
    else:
        f1[src[0] - cut + 1, src[1]] = 0
    umin = lambda a, b: torch.minimum(a.view(torch.int32).to(torch.int64) & 0xFFFFFFFF, b.to(torch.int64) & 0xFFFFFFFF)
    for it in range(64):
        c0 = fx.field_relax(g0, f0, 1)
        c1 = fx.field_relax(g1, f1, 1)
        torch.cuda.synchronize()
        if it > 0 and int(c0) == 0 and int(c1) == 0:
            break
        # exchange: each side's ghost row <- min(own ghost, neighbour's owned boundary row), both directions
        a = umin(f0[cut], f1[1]); b = umin(f0[cut - 1], f1[0])
        a = torch.where(a == 0xFFFFFFFF, torch.full_like(a, -1), a).to(torch.int32)
        b = torch.where(b == 0xFFFFFFFF, torch.full_like(b, -1), b).to(torch.int32)
        f0[cut] = a; f1[1] = a; f0[cut - 1] = b; f1[0] = b
    else:
        raise AssertionError("tiled relaxation did not converge")
    got = torch.cat([f0[:cut], f1[1:]]).cpu().numpy().astype(np.int64)
    assert np.array_equal(got, want)


# ------------------------------------------------------------------------------------------ drop-in
def test_jps1_dropin_contract(fx, oracle, golden, maps, capsys):
    from fuxi_planner_b200 import jps1
    m = maps["-16.40-4.80_out.png"].astype(np.float64)
    start = (np.int64(0), np.int64(0))
    path, secs = jps1.method(m, start, (147, 51), 1)
    out = capsys.readouterr().out.strip()
    assert out == "1674.0" and isinstance(secs, float)
    assert path[0] is start and path[-1] == (147, 51)
    assert np.array(path).shape[1] == 2                       # callers do np.array(path1[0]) + [1,1]
    validate_path(maps["-16.40-4.80_out.png"], [tuple(int(v) for v in p) for p in path], (0, 0), (147, 51))
    path, _ = jps1.method(m, (0, 0), (147, 51), 2)
    assert abs(float(capsys.readouterr().out.strip()) - 168.12489168102775) <= RTOL * 168.12489168102775
    z = np.zeros((6, 6)); z[3, :] = 1
    r = jps1.method(z, (0, 0), (5, 5), 2)
    assert r[0] is 0                                          # noqa: F632  (the planners' identity test)
    r = jps1.method(np.zeros((6, 6)), (2, 2), (2, 2), 2)
    assert r[0] == [(2, 2)] and capsys.readouterr().out.strip() == "0"
    m100 = np.zeros((5, 5)); m100[2, :] = 100                 # 100 is free: the test is == 1
    r = jps1.method(m100, (0, 2), (4, 2), 1)
    assert r[0] == [(0, 2), (4, 2)] and capsys.readouterr().out.strip() == "40.0"
    with pytest.raises(IndexError):
        jps1.method(np.zeros((6, 6)), (6, 0), (1, 1), 1)


def test_host_pipeline_map_then_plan(fx, oracle):
    rng = np.random.default_rng(5)
    pts = np.c_[rng.uniform(-20, 20, (5000, 2)), rng.uniform(0, 2, 5000)].astype(np.float32)
    g = fx.map_host(pts, None, 0.3, np.inf, (-25.6, -25.6), 0.2, (256, 256), radius=2, variant="ccst")
    want = oracle.inflate(oracle.hostref.project(pts, np.eye(3, 4), 0.3, np.inf, -25.6, -25.6, 0.2, 256, 256), 2, 1)
    assert np.array_equal(g, want)
    s, t = random_queries(g, 32, rng)
    ci, cf, pxy, pl = fx.plan_host(g, s, t, metric=1, max_path=512)
    wantc, status, _ = oracle.jps_batch(g, s, t, 1)
    assert ((status == 1) == (ci >= 0)).all()
    assert np.array_equal(ci[status == 1].astype(np.float64), wantc[status == 1])


# ------------------------------------------------------------------------------------------ multi-GPU building blocks
def test_halo_merge_kernel(fx, dev):
    import torch
    from fuxi_planner_b200 import tiled
    rng = np.random.default_rng(5)
    a = rng.integers(-1, 1000, size=(2, 1000)).astype(np.int32)
    b = rng.integers(-1, 1000, size=(2, 1000)).astype(np.int32)
    want = np.minimum(a.view(np.uint32), b.view(np.uint32)).view(np.int32)
    da, db = _t(a, dev), _t(b, dev)
    ch = torch.zeros(1, dtype=torch.int32, device=dev)
    tiled.CudaOps().merge(da, db, ch)
    assert np.array_equal(da.cpu().numpy(), want) and int(ch) == 1
    ch.zero_()
    tiled.CudaOps().merge(da, db, ch)           # idempotent: nothing improves the second time
    assert np.array_equal(da.cpu().numpy(), want) and int(ch) == 0


def test_field_tiled_single_rank_equals_field(fx, dev, oracle):
    """field_tiled with one rank (no process group) is the plain relax-to-fixpoint path used by every rank."""
    from fuxi_planner_b200 import tiled
    m = (np.random.default_rng(33).random((300, 200)) < 0.25).astype(np.uint8)
    src = tuple(int(v) for v in np.argwhere(m == 0)[3])
    for metric in (1, 2):
        fld, rounds = tiled.field_tiled(_t(m, dev), 300, src, metric)
        assert rounds == 1
        assert np.array_equal(fld.cpu().numpy().astype(np.int64), oracle.sssp_field(m, src, metric))


def test_search_pockets_and_obstacle_start(fx, dev, oracle):
    """Sealed pockets (goal side: bounded flood from the goal; start side: drained queue with nothing pruned) must
    come back unreachable without flooding the whole map, and a start on an obstacle next to / inside a pocket
    keeps the reference's semantics (the source cell is never tested, jps1.py:14-31)."""
    n = 600
    m = np.zeros((n, n), dtype=np.uint8)
    m[100:107, 100] = 1; m[100:107, 106] = 1; m[100, 100:107] = 1; m[106, 100:107] = 1      # sealed 5x5 room A
    m[400:407, 300] = 1; m[400:407, 306] = 1; m[400, 300:307] = 1; m[406, 300:307] = 1      # sealed 5x5 room B
    m[(np.random.default_rng(8).random((n, n)) < 0.1) & (m == 0)] = 1
    m[101:106, 101:106] = 0; m[401:406, 301:306] = 0
    q = [((5, 5), (103, 103)),        # goal in a pocket
         ((103, 103), (5, 5)),        # start in a pocket
         ((102, 102), (104, 104)),    # both inside the same pocket: reachable, met by the flood from the goal
         ((103, 103), (403, 303)),    # different pockets
         ((100, 103), (104, 104)),    # start ON the wall of room A, goal inside: reachable
         ((100, 103), (90, 103)),     # start on the wall, goal outside: reachable too (steps out)
         ((100, 103), (403, 303)),    # wall start, goal in the other pocket: unreachable
         ((7, 9), (590, 580))]        # ordinary long query
    s = np.array([a for a, _ in q], dtype=np.int32)
    g = np.array([b for _, b in q], dtype=np.int32)
    for metric in (1, 2):
        want = oracle.sssp_batch(m, s, g, metric)
        res = fx.plan_batch(_t(m, dev), _t(s, dev), _t(g, dev), metric=metric, max_path=2048)
        got = res.cost_i.cpu().numpy().astype(np.int64)
        assert np.array_equal(got, want), (metric, got, want)
        assert list(want[[0, 1, 3, 6]]) == [-1] * 4 and (want[[2, 4, 5, 7]] > 0).all()
        for i in np.flatnonzero(want > 0):
            validate_path(m, res.path(int(i)), tuple(s[i]), tuple(g[i]))
    settled = fx.search_stats()[0]
    assert settled < 8 * 40000, "pocket queries flooded the map (%d cells settled)" % settled


@pytest.mark.parametrize("shape,fill", [((120, 140), 0.1), ((141, 141), 0.3), ((199, 100), 0.45), ((3, 5000), 0.2), ((1, 40), 0.0)])
def test_search_small_maps_vs_oracle(fx, dev, oracle, shape, fill, search_form):
    """Maps of the reference's own size (< 20 000 cells) against the Dijkstra oracle, through both kernel forms: random
    queries including starts ON obstacles and goals in sealed pockets, both metrics, paths validated move by move."""
    rng = np.random.default_rng(shape[0] * 7 + int(fill * 100))
    m = (rng.random(shape) < fill).astype(np.uint8)
    Q = 200
    s = np.c_[rng.integers(shape[0], size=Q), rng.integers(shape[1], size=Q)].astype(np.int32)      # any cell, obstacles too
    g = np.c_[rng.integers(shape[0], size=Q), rng.integers(shape[1], size=Q)].astype(np.int32)
    for metric in (1, 2):
        want = oracle.sssp_batch(m, s, g, metric)
        res = fx.plan_batch(_t(m, dev), _t(s, dev), _t(g, dev), metric=metric, max_path=4096)
        got = res.cost_i.cpu().numpy().astype(np.int64)
        assert np.array_equal(got, want), (metric, np.flatnonzero(got != want)[:5])
        for i in np.flatnonzero(want > 0)[:60]:
            a, b = validate_path(m, res.path(int(i)), tuple(s[i]), tuple(g[i]))
            assert a * (10 if metric == 1 else fx.FX_EUCLID_WS) + b * (14 if metric == 1 else fx.FX_EUCLID_WD) == want[i]


def test_plan_host_float64_matrix_and_wide_ctas(fx, oracle):
    """fx_plan_host_f64 evaluates the reference's `matrix[x][y] == 1` (jps1.py:20-29) on the float64 matrix itself
    (threaded, chunked into the pinned staging buffer): same answers as the uint8 entry point, values other than
    exactly 1.0 are free.  Small batches run the wide-CTA (latency) form of the search kernel: checked against the
    oracle here, and against the throughput form on the same queries."""
    rng = np.random.default_rng(77)
    W, H = 2304, 2048                       # > 2^22 cells: exercises the chunked conversion
    occ = (rng.random((W, H)) < 0.2).astype(np.uint8)
    mat = occ.astype(np.float64)
    free = mat == 0
    mat[free & (rng.random((W, H)) < 0.05)] = 100.0          # "occupied" in the message encoding, free for the search
    mat[free & (rng.random((W, H)) < 0.01)] = 0.999999
    mat[free & (rng.random((W, H)) < 0.01)] = np.nan
    s, g = random_queries(occ, 6, rng)
    want, status, _ = oracle.jps_batch(occ, s, g, 1)
    a = fx.plan_host(mat, s, g, metric=1, max_path=64)
    b = fx.plan_host(occ, s, g, metric=1, max_path=64)
    assert np.array_equal(a[0], b[0])           # (equal-cost paths may differ between two runs: compare costs, validate paths)
    for q in range(6):
        assert (a[0][q] == int(want[q])) if status[q] == 1 else (a[0][q] == -1)
        for r in (a, b):
            if status[q] == 1 and r[3][q] <= 64:
                na, nb = validate_path(occ, [tuple(p) for p in r[2][q, :r[3][q]]], tuple(s[q]), tuple(g[q]))
                assert 10 * na + 14 * nb == r[0][q]
    # the same six queries inside a batch large enough for the throughput form
    S, G = random_queries(occ, 400, rng)
    S[:6], G[:6] = s, g
    c = fx.plan_host(occ, S, G, metric=1, max_path=0)
    assert np.array_equal(c[0][:6], a[0])


def test_plan_host_odd_sizes_through_the_staged_upload(fx, oracle):
    """Grids above 2^20 cells reach the device through the worker pool's non-temporal fill of the pinned staging buffer,
    cut into 64 tasks / four upload parts (csrc/api.cu staged_upload): odd extents leave ragged last tasks and unaligned
    rows, for the uint8 copy and for the float64 `== 1.0` conversion alike."""
    rng = np.random.default_rng(123)
    for W, H in ((1031, 1057), (1500, 701)):
        occ = (rng.random((W, H)) < 0.2).astype(np.uint8)
        s, g = random_queries(occ, 5, rng)
        want, status, _ = oracle.jps_batch(occ, s, g, 1)
        mat = occ.astype(np.float64)
        mat[(occ == 0) & (rng.random((W, H)) < 0.03)] = 1.0000000000000002   # not == 1.0: free
        for grid in (occ, mat, np.asfortranarray(occ)):
            r = fx.plan_host(grid, s, g, metric=1, max_path=0)
            for q in range(5):
                assert (r[0][q] == int(want[q])) if status[q] == 1 else (r[0][q] == -1), (W, H, grid.dtype, q)


def test_plan_host_stage_trace(fx, oracle, monkeypatch):
    """FUXI_B200_TRACE=2 records where a host-buffer call spends its time (fx_plan_host_stages: the latency break-down
    SURVEY 8d asks for); an untraced call, or one that took the shared-memory kernel, has none."""
    rng = np.random.default_rng(8)
    occ = (rng.random((1100, 1000)) < 0.2).astype(np.uint8)
    s, g = random_queries(occ, 2, rng)
    monkeypatch.setenv("FUXI_B200_TRACE", "2")
    r = fx.plan_host(occ, s, g, metric=1, max_path=0)
    st = fx.plan_host_stages()
    assert len(st) == 6 and all(v >= 0.0 for v in st) and st[3] > 0.0 and st[4] > 0.0
    want, status, _ = oracle.jps_batch(occ, s, g, 1)
    assert all((r[0][q] == int(want[q])) if status[q] == 1 else (r[0][q] == -1) for q in range(2))
    monkeypatch.delenv("FUXI_B200_TRACE")
    fx.plan_host(occ, s, g, metric=1, max_path=0)
    with pytest.raises(fx.FuxiError):
        fx.plan_host_stages()


def test_latency_forms_after_a_maze_of_the_same_shape(fx, dev, oracle, monkeypatch):
    """The latency forms size their first pass from a per-context table of how far above the octile bound earlier optima
    lay (search.cu fx_first_bound).  A serpentine maze teaches it ratios of 10 and more; the next map of the same shape is
    an open one.  The guess is capped (1.5 h0) and the cluster form's per-CTA queue segments must hold the wide levels of
    a barely pruned search on a small grid (r02j fuzz: FX_COST_OVERFLOW on a 121 x 42 grid).  Answers never depend on
    the table."""
    monkeypatch.setenv("FUXI_B200_SMALL", "0")
    rng = np.random.default_rng(2024)
    W, H = 121, 42
    maze = np.zeros((W, H), dtype=np.uint8)
    for x in range(3, W, 6):
        maze[x, :] = 1
        maze[x, (H - 2) if (x // 6) % 2 else 1] = 0
    ms = np.array([[0, 0], [1, 20], [0, 41], [2, 5]], dtype=np.int32)
    mg = np.array([[120, 41], [119, 3], [120, 0], [118, 30]], dtype=np.int32)
    open_map = (rng.random((W, H)) < 0.05).astype(np.uint8)
    s, g = random_queries(open_map, 12, rng)
    s[0], g[0] = (68, 12), (112, 17)
    open_map[68, 12] = 0; open_map[112, 17] = 0
    for metric in (1, 2):
        for rep in range(3):                                   # the table decays slowly: the open map sees the maze's values
            r = fx.plan_batch(_t(maze, dev), _t(ms, dev), _t(mg, dev), metric=metric, max_path=4096)
            assert np.array_equal(r.cost_i.cpu().numpy().astype(np.int64), oracle.sssp_batch(maze, ms, mg, metric))
        for Q in (1, 3, 12):                                   # cluster form (<= 18 queries)
            r = fx.plan_batch(_t(open_map, dev), _t(s[:Q], dev), _t(g[:Q], dev), metric=metric, max_path=4096)
            assert np.array_equal(r.cost_i.cpu().numpy().astype(np.int64), oracle.sssp_batch(open_map, s[:Q], g[:Q], metric)), (metric, Q)


def test_search_extreme_shapes_and_empty_batches(fx, dev, oracle):
    """The largest extent the packed (x << 16 | y) queue entries allow (32767), one-cell-wide corridors, an empty batch,
    max_path = 0, and the refusal above the limit."""
    import torch
    rng = np.random.default_rng(9)
    for shape in ((32767, 4), (4, 32767)):
        m = (rng.random(shape) < 0.15).astype(np.uint8)
        m[:, 0] = 0 if shape[1] == 4 else m[:, 0]
        if shape[0] == 4:
            m[0, :] = 0
        free = np.argwhere(m == 0)
        s = free[rng.integers(len(free), size=24)].astype(np.int32)
        g = free[rng.integers(len(free), size=24)].astype(np.int32)
        s[0], g[0] = free[0], free[-1]                      # end to end: ~32766 straight steps
        want = oracle.sssp_batch(m, s, g, 2)
        res = fx.plan_batch(_t(m, dev), _t(s, dev), _t(g, dev), metric=2, max_path=8192)
        got = res.cost_i.cpu().numpy().astype(np.int64)
        assert np.array_equal(got, want), np.flatnonzero(got != want)[:5]
        assert want[0] > 32000 * fx.FX_EUCLID_WS
        validate_path(m, res.path(0), tuple(s[0]), tuple(g[0]))
    m = np.zeros((64, 64), dtype=np.uint8)
    empty = torch.zeros((0, 2), dtype=torch.int32, device=dev)
    res = fx.plan_batch(_t(m, dev), empty, empty, metric=1)
    assert res.cost_i.numel() == 0
    one = torch.tensor([[1, 1]], dtype=torch.int32, device=dev), torch.tensor([[60, 50]], dtype=torch.int32, device=dev)
    res = fx.plan_batch(_t(m, dev), one[0], one[1], metric=1, max_path=0)      # costs only
    assert res.path_xy is None and int(res.cost_i[0]) == 14 * 49 + 10 * 10
    with pytest.raises(fx.FuxiError):
        fx.plan_batch(torch.zeros((32768, 2), dtype=torch.uint8, device=dev), one[0], one[1], metric=1)


def test_search_slot_reuse_resets_completely(fx, dev, oracle):
    """Two search slots, 400 queries: every slot answers ~200 queries one after the other on the same scratch field, so a
    field line that a query touched and the reset missed would corrupt a later answer.  Costs against the oracle."""
    import torch
    rng = np.random.default_rng(31)
    m = (rng.random((300, 340)) < 0.25).astype(np.uint8)
    s, g = random_queries(m, 400, rng)
    ctx = fx.Context(0)
    ctx.check(ctx.lib.fx_set_search_tuning(ctx.handle, 2, 0), "fx_set_search_tuning")
    try:
        for metric in (1, 2):
            want = oracle.sssp_batch(m, s, g, metric)
            res = fx.plan_batch(_t(m, dev), _t(s, dev), _t(g, dev), metric=metric, max_path=0, ctx=ctx)
            torch.cuda.synchronize()
            assert np.array_equal(res.cost_i.cpu().numpy().astype(np.int64), want), metric
    finally:
        ctx.close()


# ------------------------------------------------------------------------------------------ BASELINE configurations
def _cfg4_workload(n=4096, q=256):
    """cfg4 of SURVEY §8d exactly as bench.py builds it: grid default_rng(4) at 20 % fill, queries default_rng(5)."""
    m = (np.random.default_rng(4).random((n, n)) < 0.2).astype(np.uint8)
    free = np.argwhere(m == 0)
    rng = np.random.default_rng(5)
    s = free[rng.integers(len(free), size=8192)].astype(np.int32)[:q]
    g = free[rng.integers(len(free), size=8192)].astype(np.int32)[:q]
    return m, s, g


def test_search_cfg4_4096(fx, dev, oracle, cfg4_golden):
    """The headline configuration (cfg4: 4096^2, 20 % fill): the first 256 queries of the benchmark's batch, both
    metrics, against the C restatement of jps1.py, every path validated move by move; plus the queries whose costs
    were produced by the UNMODIFIED reference at this size (tests/golden/cfg4_golden.json, minutes per query)."""
    m, s, g = _cfg4_workload()
    gm = _t(m, dev)
    for metric in (1, 2):
        want, status, _ = oracle.jps_batch(m, s, g, metric)
        res = fx.plan_batch(gm, _t(s, dev), _t(g, dev), metric=metric, max_path=2048)
        ci, cf = res.cost_i.cpu().numpy(), res.cost_f.cpu().numpy()
        assert (status == 1).sum() > 250
        for q in range(len(s)):
            if status[q] != 1:
                assert ci[q] == -1, (metric, q, ci[q])
            elif metric == 1:
                assert ci[q] == int(want[q]), (metric, q, ci[q], want[q])
            else:
                assert abs(cf[q] - want[q]) <= RTOL * max(want[q], 1e-12), (metric, q, cf[q], want[q])
        for q in np.flatnonzero(status == 1)[::4]:
            a, b = validate_path(m, res.path(int(q)), tuple(s[q]), tuple(g[q]))
            if metric == 1:
                assert 10 * a + 14 * b == ci[q]
            else:
                assert a * fx.FX_EUCLID_WS + b * fx.FX_EUCLID_WD == ci[q] and cf[q] == a + b * SQRT2
        # the reference's own answers
        recs = [r for r in cfg4_golden if r["h"] == metric]
        assert len(recs) >= 3
        for r in recs:
            q = r["index"]
            assert list(s[q]) == r["start"] and list(g[q]) == r["goal"]
            if metric == 1:
                assert ci[q] == int(float(r["cost"])), (q, ci[q], r["cost"])
            else:
                assert abs(cf[q] - float(r["cost"])) <= RTOL * float(r["cost"]), (q, cf[q], r["cost"])
    # the latency forms on the same grid: one query per thread-block cluster (at most sm_count / 8 queries), one query per
    # CTA (at most sm_count queries; also what FUXI_B200_CLUSTER=0 selects for the smallest batches): same costs, valid paths
    import os
    for nq, env in ((8, None), (8, "0"), (40, None)):
        old = os.environ.get("FUXI_B200_CLUSTER")
        if env is not None:
            os.environ["FUXI_B200_CLUSTER"] = env
        try:
            ctx = fx.Context(0)
        finally:
            if env is not None:
                if old is None:
                    del os.environ["FUXI_B200_CLUSTER"]
                else:
                    os.environ["FUXI_B200_CLUSTER"] = old
        try:
            res1 = fx.plan_batch(gm, _t(s[:nq], dev), _t(g[:nq], dev), metric=2, max_path=2048, ctx=ctx)
            assert np.array_equal(res1.cost_i.cpu().numpy(), ci[:nq]), (nq, env)
            for q in range(0, nq, 3):
                if ci[q] > 0:
                    a, b = validate_path(m, res1.path(q), tuple(s[q]), tuple(g[q]))
                    assert a * fx.FX_EUCLID_WS + b * fx.FX_EUCLID_WD == ci[q]
        finally:
            ctx.close()


def test_cfg5_field_vs_oracle(fx, dev, oracle):
    """cfg5 (16384^2, default_rng(6), 20 % fill; query = first free cell -> last free cell): the whole cost field of one
    GPU against the C oracle's Dijkstra field bit for bit, and the goal's cost against the goal-directed batched search
    (fx_search_batch) on the same query.  The oracle's field costs ~75 s of host time."""
    import torch
    n = 16384
    m = (np.random.default_rng(6).random((n, n)) < 0.2).astype(np.uint8)
    free_first = np.unravel_index(np.flatnonzero(m.reshape(-1) == 0)[0], m.shape)
    free_last = np.unravel_index(np.flatnonzero(m.reshape(-1) == 0)[-1], m.shape)
    s = np.array([free_first], dtype=np.int32)
    g = np.array([free_last], dtype=np.int32)
    gm = _t(m, dev)
    want = oracle.sssp_field(m, tuple(int(v) for v in s[0]), metric=2)          # int64, -1 unreachable
    field = fx.field(gm, tuple(int(v) for v in s[0]), metric=2)
    got = field.cpu().numpy()
    assert got.dtype == np.int32 and got.shape == (n, n)
    assert np.array_equal(got, want.astype(np.int32))
    goal_cost = int(want[g[0][0], g[0][1]])
    assert goal_cost > 16000 * fx.FX_EUCLID_WS
    del field, got
    torch.cuda.empty_cache()
    res = fx.plan_batch(gm, _t(s, dev), _t(g, dev), metric=2, max_path=8192)
    assert int(res.cost_i[0]) == goal_cost
    a, b = validate_path(m, res.path(0), tuple(s[0]), tuple(g[0]))
    assert a * fx.FX_EUCLID_WS + b * fx.FX_EUCLID_WD == goal_cost


def test_paths_compact_and_host_csr(fx, dev, oracle):
    """Compact (CSR) path output: fx_paths_compact of the padded rows and fx_plan_host_csr against the padded forms,
    including unreachable queries, start == goal, and paths that do not fit max_path (they contribute no point)."""
    rng = np.random.default_rng(41)
    m = (rng.random((700, 650)) < 0.3).astype(np.uint8)
    s, g = random_queries(m, 300, rng)
    s[5] = g[5]
    m2 = m.copy()
    m2[200:260, 300] = 1; m2[200:260, 360] = 1; m2[200, 300:361] = 1; m2[259, 300:361] = 1   # a sealed room
    m2[220:240, 320:340] = 0
    g[7] = (230, 330)
    for max_path in (256, 12):
        res = fx.plan_batch(_t(m2, dev), _t(s, dev), _t(g, dev), metric=2, max_path=max_path)
        pl = res.path_len.cpu().numpy()
        pxy = res.path_xy.cpu().numpy()
        off, xy = fx.paths_compact(res.path_xy, res.path_len)
        off, xy = off.cpu().numpy(), xy.cpu().numpy()
        n = np.where((pl > 0) & (pl <= max_path), pl, 0)
        assert np.array_equal(off, np.concatenate([[0], np.cumsum(n)]))
        for q in range(len(s)):
            assert np.array_equal(xy[off[q]:off[q + 1]], pxy[q, :n[q]])
        assert pl[5] == 1 and pl[7] == -1
        if max_path == 12:
            assert (pl > max_path).any()
        ci, cf, hxy, hpl = fx.plan_host(m2, s, g, metric=2, max_path=max_path)
        ci2, cf2, hpl2, hoff, hx = fx.plan_host_csr(m2, s, g, metric=2, max_path=max_path, cap=16)   # forces the retry
        assert np.array_equal(ci, res.cost_i.cpu().numpy()) and np.array_equal(ci, ci2) and np.array_equal(cf, cf2)
        assert np.array_equal(hpl, pl) and np.array_equal(hpl2, pl) and np.array_equal(hoff, off) and np.array_equal(hx, xy)
        for q in range(len(s)):
            assert np.array_equal(hxy[q, :n[q]], pxy[q, :n[q]])
    # a map of the reference's own size goes through the shared-memory kernel: same contract
    ms = (rng.random((120, 90)) < 0.2).astype(np.uint8)
    ss, gs = random_queries(ms, 40, rng)
    a = fx.plan_host(ms, ss, gs, metric=1, max_path=64)
    b = fx.plan_host_csr(ms, ss, gs, metric=1, max_path=64)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[2])
    for q in range(40):
        k = max(int(a[3][q]), 0)
        assert np.array_equal(b[4][b[3][q]:b[3][q + 1]], a[2][q, :k])


def _check_jump_list(oracle, m, jp, tp, s, g, want_cost, ws, wd):
    """jp: jump-point list of the path whose turning points are tp."""
    a, b = validate_path(m, jp, s, g)
    assert a * ws + b * wd == want_cost
    it = iter(jp)
    assert all(any(tuple(p) == tuple(q) for q in it) for p in tp), "turning points must be a subsequence of the jump points"
    for p, q in zip(jp[:-1], jp[1:]):
        d = (int(np.sign(q[0] - p[0])), int(np.sign(q[1] - p[1])))
        assert oracle.jump(m, p, d, g) == tuple(q), ("not what jps1.jump returns", p, d, q)


def test_jump_point_path_form(fx, dev, oracle, maps, golden):
    """Opt-in jump-point output (fx_paths_jump_points / jps1.POINTS = "jump"): every consecutive pair of the returned list
    is exactly what the reference's jump() returns from the first point in that direction (checked with the restated
    jump, itself pinned on the reference in tests/test_oracle.py), the list contains the turning points, costs unchanged.
    Where the reference's own path (jps1_golden.json) runs through the same cells the two lists are identical."""
    import contextlib
    import io
    rng = np.random.default_rng(12)
    same = total = 0
    for name in sorted(maps)[::4]:
        m = maps[name]
        recs = [r for r in golden["maps"][name] if r["h"] == 1 and r["cost"] is not None and "path" in r]
        s = np.array([r["start"] for r in recs], dtype=np.int32)
        g = np.array([r["goal"] for r in recs], dtype=np.int32)
        res = fx.plan_batch(_t(m, dev), _t(s, dev), _t(g, dev), metric=1, max_path=512)
        jxy, jl = fx.paths_jump_points(_t(m, dev), res.path_xy, res.path_len)
        jxy, jl = jxy.cpu().numpy(), jl.cpu().numpy()
        for q, r in enumerate(recs):
            if s[q].tolist() == g[q].tolist():
                continue
            tp = res.path(q)
            jp = [tuple(int(v) for v in p) for p in jxy[q, :jl[q]]]
            _check_jump_list(oracle, m, jp, tp, tuple(s[q]), tuple(g[q]), int(float(r["cost"])), 10, 14)
            total += 1
            same += jp == [tuple(p) for p in r["path"]]
    assert total > 100 and same >= 1, (same, total)
    print("jump-point lists identical to the reference's own:", same, "of", total)
    # a large random grid, both metrics
    m = (rng.random((1024, 1024)) < 0.2).astype(np.uint8)
    s, g = random_queries(m, 64, rng)
    ctx = fx.Context(0)
    ctx.set_search_form("throughput")      # forward searches only: forward-canonical paths (a batch this small would be bidirectional)
    try:
        for metric, ws, wd in ((1, 10, 14), (2, fx.FX_EUCLID_WS, fx.FX_EUCLID_WD)):
            res = fx.plan_batch(_t(m, dev), _t(s, dev), _t(g, dev), metric=metric, max_path=2048, ctx=ctx)
            jxy, jl = fx.paths_jump_points(_t(m, dev), res.path_xy, res.path_len, max_out=4096, ctx=ctx)
            jxy, jl, ci = jxy.cpu().numpy(), jl.cpu().numpy(), res.cost_i.cpu().numpy()
            for q in range(0, 64, 4):
                if ci[q] > 0:
                    _check_jump_list(oracle, m, [tuple(int(v) for v in p) for p in jxy[q, :jl[q]]], res.path(q), tuple(s[q]), tuple(g[q]), int(ci[q]), ws, wd)
    finally:
        ctx.close()
    # the drop-in on a map beyond the shared-memory kernel: jump mode switches to forward searches by itself
    mf = m.astype(np.float64)
    old = fx.jps1.POINTS
    try:
        fx.jps1.POINTS = "jump"
        with contextlib.redirect_stdout(io.StringIO()):
            jp, _ = fx.jps1.method(mf, tuple(int(v) for v in s[1]), tuple(int(v) for v in g[1]), 1)
    finally:
        fx.jps1.POINTS = old
    for p, q in zip(jp[:-1], jp[1:]):
        d = (int(np.sign(q[0] - p[0])), int(np.sign(q[1] - p[1])))
        assert oracle.jump(m, p, d, tuple(int(v) for v in g[1])) == tuple(q)
    # the drop-in
    name = "-16.40-4.80_out.png"
    mf = maps[name].astype(np.float64)
    old = fx.jps1.POINTS
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            fx.jps1.POINTS = "turning"
            tp, _ = fx.jps1.method(mf, (0, 0), (147, 51), 1)
            fx.jps1.POINTS = "jump"
            jp, _ = fx.jps1.method(mf, (0, 0), (147, 51), 1)
    finally:
        fx.jps1.POINTS = old
    assert len(jp) >= len(tp) and jp[0] == (0, 0) and jp[-1] == (147, 51)
    _check_jump_list(oracle, maps[name], jp, tp, (0, 0), (147, 51), 1674, 10, 14)
