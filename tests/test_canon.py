"""Successor rule of the batched search (fx_canon_successors / csrc/common.cuh:fx_canon_succ).

CPU tests: the rule is a pure host function of the C-ABI library, so it is checked here without a GPU
  * against the golden table generated from the reference's own jps1.nodeNeighbours (tests/golden/make_canon_golden.py),
  * live against the reference when /root/reference is present,
  * and as an algorithm: a wavefront that relaxes only those successors, with random tie-breaking between
    equal-cost parents (what the GPU's atomics amount to), reproduces the exact cost field of the oracle's Dijkstra.
"""
import json
import os
import random

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
DIRS = [(-1, 0), (1, 0), (0, -1), (0, 1), (-1, -1), (-1, 1), (1, -1), (1, 1)]


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    import fuxi_planner_b200 as fx
    return fx.load()


def test_successor_rule_matches_reference_golden(lib):
    g = json.load(open(os.path.join(HERE, "golden", "canon_succ_golden.json")))
    assert [tuple(d) for d in g["dirs"]] == DIRS
    checked = 0
    for rec in g["records"]:
        for code, want in enumerate(rec["succ"]):
            if want is None:
                continue
            got = lib.fx_canon_successors(code, rec["moves"])
            assert got == want, (rec["where"], rec["pattern"], code, rec["moves"], got, want)
            checked += 1
    assert checked > 2500
    assert lib.fx_canon_successors(9, 0) < 0 and lib.fx_canon_successors(0, 256) < 0


def test_successor_rule_matches_live_reference(lib):
    from oracle import refload
    if not refload.available():
        pytest.skip("reference tree not present (GPU box)")
    j = refload.jps1_module()
    rng = np.random.default_rng(7)
    for _ in range(300):
        m = (rng.random((7, 7)) < rng.uniform(0.05, 0.6)).astype(np.float64)
        cx, cy = int(rng.integers(7)), int(rng.integers(7))
        moves = sum(1 << d for d, (dx, dy) in enumerate(DIRS) if not j.blocked(cx, cy, dx, dy, m))
        for code, (dx, dy) in enumerate(DIRS):
            px, py = cx - dx, cy - dy
            if not (0 <= px < 7 and 0 <= py < 7) or j.blocked(px, py, dx, dy, m):
                continue
            want = 0
            for (nx, ny) in j.nodeNeighbours(cx, cy, (px, py), m):
                if not j.blocked(cx, cy, nx - cx, ny - cy, m):
                    want |= 1 << DIRS.index((nx - cx, ny - cy))
            assert lib.fx_canon_successors(code, moves) == want


def _canonical_field(lib, occ, src, ws, wd, rnd):
    """Dial wavefront (bucket width ws) that relaxes only fx_canon_successors, arbitrary winner among equal costs."""
    W, H = occ.shape

    def blk(x, y):
        return x < 0 or x >= W or y < 0 or y >= H or occ[x, y] == 1
    moves = np.zeros((W, H), dtype=np.int64)
    for x in range(W):
        for y in range(H):
            for d, (dx, dy) in enumerate(DIRS):
                ok = not blk(x + dx, y + dy)
                if d >= 4:
                    ok = ok and not (blk(x + dx, y) and blk(x, y + dy))
                moves[x, y] |= int(ok) << d
    INF = 1 << 60
    g = np.full((W, H), INF, dtype=np.int64)
    code = np.full((W, H), 8, dtype=np.int64)
    g[src] = 0
    buckets = {0: [src]}
    k = 0
    while buckets:
        if k not in buckets:
            k = min(buckets)
        cells = buckets.pop(k)
        rnd.shuffle(cells)
        for (x, y) in cells:
            if g[x, y] // ws != k:
                continue
            succ = lib.fx_canon_successors(int(code[x, y]), int(moves[x, y]))
            for d, (dx, dy) in enumerate(DIRS):
                if not (succ >> d) & 1:
                    continue
                ng = g[x, y] + (ws if d < 4 else wd)
                c = (x + dx, y + dy)
                if ng < g[c] or (ng == g[c] and rnd.random() < 0.5):
                    new = ng < g[c]
                    g[c], code[c] = ng, d
                    if new:
                        buckets.setdefault(int(ng // ws), []).append(c)
        k += 1
    g[g == INF] = -1
    return g


@pytest.mark.parametrize("seed", range(6))
def test_canonical_wavefront_equals_dijkstra(lib, seed):
    import oracle
    rnd = random.Random(seed)
    rng = np.random.default_rng(100 + seed)
    for trial in range(12):
        W, H = int(rng.integers(2, 22)), int(rng.integers(2, 22))
        occ = (rng.random((W, H)) < rng.uniform(0.0, 0.55)).astype(np.uint8)
        if trial % 3 == 0:        # walls with one gap
            for x in range(2, W, 4):
                gap = int(rng.integers(H))
                occ[x, :] = 1
                occ[x, gap] = 0
        src = (int(rng.integers(W)), int(rng.integers(H)))      # may sit on an obstacle, like the reference allows
        metric = 1 + (trial & 1)
        ws, wd = (10, 14) if metric == 1 else (oracle.capi.FX_WS, oracle.capi.FX_WD)
        want = oracle.sssp_field(occ, src, metric)
        got = _canonical_field(lib, occ, src, ws, wd, rnd)
        assert np.array_equal(got, want), (seed, trial, W, H, src)
