"""GPU parity tests for the wire / disk formats either side of the path: the planners' map image dump
and pre-map loader (global_planner_st.py:176-182, 365-374), the pre-map merge (st:210-224) -- against the numpy
restatements in oracle/hostref.py on the reference's own maps (tests/golden/maps.npz).  Byte work: bit-exact."""
import numpy as np
import pytest

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


def test_grid_image_all_maps(fx, oracle, maps):
    for name, m in maps.items():
        m100 = m.astype(np.uint8) * 100            # what the planners hold when they save (publish_map rewrote 1 -> 100)
        img = fx.formats.grid_to_image(m100)
        want = oracle.hostref.grid_to_png_array(m100)
        assert img.shape == want.shape and np.array_equal(img, want), name
        back = fx.formats.image_to_grid(img, threshold=0)
        assert np.array_equal(back, oracle.hostref.png_to_grid(want)) and np.array_equal(back, m), name


@pytest.mark.parametrize("shape", [(1, 1), (31, 33), (64, 64), (257, 130), (1000, 37)])
def test_premap_loader_gray_images(fx, oracle, shape):
    rng = np.random.default_rng(shape[0])
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)      # [rows = y][cols = x]
    img.flat[:3] = (200, 201, 199)                               # the threshold itself: > 200 is free
    got = fx.formats.image_to_grid(img, threshold=200)
    want = oracle.hostref.load_premap(img, 200)
    assert got.shape == want.shape == (shape[1], shape[0]) and np.array_equal(got, want)


def test_png_roundtrip_on_disk(fx, maps, tmp_path):
    m = maps["-16.40-4.80_out.png"]
    path = fx.formats.save_map_png(m, (-16.4, -4.8), str(tmp_path))
    assert path.endswith("-16.40-4.80_out.png")
    assert np.array_equal(fx.formats.load_map_png(path), m)


def test_premap_merge(fx, oracle, maps):
    import torch
    rng = np.random.default_rng(3)
    pre = maps["-16.40-4.80_out.png"]
    for case in range(12):
        det = (rng.random((int(rng.integers(5, 90)), int(rng.integers(5, 70)))) < 0.3).astype(np.uint8)
        reso = 0.2
        ori_pre = [-15.0, -15.0]
        map_o = [float(np.round(rng.uniform(-22, -8), 1)), float(np.round(rng.uniform(-22, -8), 1))]
        map_t = [map_o[0] + reso * det.shape[0], map_o[1] + reso * det.shape[1]]
        try:
            want, want_o = oracle.hostref.merge_premap(det, map_o, map_t, pre, ori_pre, reso)
        except ValueError:
            with pytest.raises(ValueError):
                fx.formats.merge_premap(torch.from_numpy(det).cuda(), map_o, map_t, torch.from_numpy(pre).cuda(), ori_pre, reso)
            continue
        got, got_o = fx.formats.merge_premap(torch.from_numpy(det).cuda(), map_o, map_t, torch.from_numpy(pre).cuda(), ori_pre, reso)
        assert list(got_o) == list(want_o)
        assert np.array_equal(got.cpu().numpy(), want.astype(np.uint8)), case
