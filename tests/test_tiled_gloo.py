"""Host-side logic of the multi-GPU modes over gloo (world_size 2 and 3, CPU): partitioning, the halo-exchange
protocol and its termination.  The two device operations are replaced by CPU stand-ins (a heap Dijkstra seeded by
every finite cell, numpy min) so that only the protocol in fuxi_planner_b200/tiled.py is under test; the stitched
field must equal the oracle's single-source field bit for bit.  The CUDA ops are covered by tests/test_gpu_parity.py."""
import heapq
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import legal_move


class CpuOps:
    """Stand-ins with the same contract as tiled.CudaOps (test infrastructure)."""

    def relax(self, grid, field, metric):
        ws, wd = (10, 14) if metric == 1 else (2378, 3363)
        occ = grid.numpy()
        f = field.numpy()
        W, H = occ.shape
        heap = [(int(f[x, y]), x, y) for x in range(W) for y in range(H) if f[x, y] >= 0]
        heapq.heapify(heap)
        changed = False
        while heap:
            g, x, y = heapq.heappop(heap)
            if g != f[x, y]:
                continue
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    if (dx or dy) and legal_move(occ, x, y, dx, dy):
                        ng = g + (wd if dx and dy else ws)
                        if f[x + dx, y + dy] < 0 or ng < f[x + dx, y + dy]:
                            f[x + dx, y + dy] = ng
                            changed = True
                            heapq.heappush(heap, (ng, x + dx, y + dy))
        return torch.tensor([1 if changed else 0], dtype=torch.int32)

    def merge(self, dst, src, changed):
        a = dst.numpy().view(np.uint32)
        b = src.numpy().view(np.uint32)
        better = b < a
        if better.any():
            a[better] = b[better]
            changed[0] = 1

    def inflate(self, grid, radius, variant):
        import oracle
        step = 1 if variant == "ccst" else max(radius, 1)
        return torch.from_numpy(oracle.inflate(grid.numpy(), radius, step))


    def edt_rows(self, grid):
        occ = grid.numpy() > 0
        w, H = occ.shape
        g = np.full((w, H), 0xFFFF, dtype=np.uint16)
        ys = np.arange(H)
        for x in range(w):
            idx = np.flatnonzero(occ[x])
            if len(idx):
                g[x] = np.abs(ys[:, None] - idx[None, :]).min(axis=1)
        return torch.from_numpy(g.view(np.int16))

    def edt_cols(self, g):
        gv = g.numpy().view(np.uint16).astype(np.int64)
        W, hb = gv.shape
        g2 = np.where(gv == 0xFFFF, 1 << 40, gv * gv)
        xs = np.arange(W)
        d = ((xs[:, None] - xs[None, :]) ** 2)[:, :, None] + g2[None, :, :]      # [x][x'][y]
        return torch.from_numpy(np.minimum(d.min(axis=1), 0x7FFFFFFF).astype(np.int32))

    def project(self, points, affine, zmin, zmax, origin, reso, shape):
        import oracle
        A = np.eye(3, 4) if affine is None else affine
        return torch.from_numpy(oracle.hostref.project(points.numpy(), A, zmin, zmax, origin[0], origin[1], reso, shape[0], shape[1]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cloud(seed, W, H):
    rng = np.random.default_rng(seed + 1)
    return np.c_[rng.uniform(-1, 0.5 * W + 1, 500), rng.uniform(-1, 0.5 * H + 1, 500), rng.uniform(-0.5, 3, 500)].astype(np.float32)


def _worker(rank, world, port, W, H, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fuxi_planner_b200 import tiled
        m = (np.random.default_rng(seed).random((W, H)) < 0.3).astype(np.uint8)
        src = tuple(int(v) for v in np.argwhere(m == 0)[7])
        x0, x1 = tiled.slab_bounds(W, world, rank)
        own = torch.from_numpy(m[x0:x1].copy())
        fld, rounds = tiled.field_tiled(own, W, src, metric=1, ops=CpuOps())
        inf = tiled.inflate_tiled(own, 2, "ccst", ops=CpuOps())
        inf_st = tiled.inflate_tiled(own, 3, "st", ops=CpuOps())
        edt = tiled.edt_tiled(own, W, ops=CpuOps())
        pts = _cloud(seed, W, H)
        p0, p1 = tiled.shard_queries(len(pts), world, rank)
        proj = tiled.project_sharded(torch.from_numpy(pts[p0:p1].copy()), None, 0.3, float("inf"), (0.0, 0.0), 0.5, (W, H), ops=CpuOps())
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), field=fld.numpy(), inflate=inf.numpy(), inflate_st=inf_st.numpy(),
                 rounds=rounds, x0=x0, x1=x1, edt=edt.numpy(), proj=proj.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,W,H", [(2, 48, 40), (3, 61, 33)])
def test_row_tiled_field_and_inflation_match_single_device(tmp_path, oracle, world, W, H):
    seed = 100 + world
    mp.spawn(_worker, args=(world, _free_port(), W, H, seed, str(tmp_path)), nprocs=world, join=True)
    m = (np.random.default_rng(seed).random((W, H)) < 0.3).astype(np.uint8)
    src = tuple(int(v) for v in np.argwhere(m == 0)[7])
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    assert [int(p["x0"]) for p in parts] == [parts[r - 1]["x1"] if r else 0 for r in range(world)]
    assert int(parts[-1]["x1"]) == W
    field = np.concatenate([p["field"] for p in parts]).astype(np.int64)
    assert np.array_equal(field, oracle.sssp_field(m, src, 1))
    assert np.array_equal(np.concatenate([p["inflate"] for p in parts]), oracle.inflate(m, 2, 1))
    assert np.array_equal(np.concatenate([p["inflate_st"] for p in parts]), oracle.inflate(m, 3, 3))
    assert np.array_equal(np.concatenate([p["edt"] for p in parts]), oracle.edt(m))
    want = oracle.hostref.project(_cloud(seed, W, H), np.eye(3, 4), 0.3, np.inf, 0.0, 0.0, 0.5, W, H)
    assert want.sum() > 100 and all(np.array_equal(p["proj"], want) for p in parts)
    assert all(int(p["rounds"]) == int(parts[0]["rounds"]) for p in parts)     # every rank leaves the loop together
    assert int(parts[0]["rounds"]) >= 2


def test_partitions_cover_everything():
    from fuxi_planner_b200 import tiled
    for n in (1, 2, 3, 8):
        for W in (8, 61, 4096, 65536 + 3):
            b = [tiled.slab_bounds(W, n, r) for r in range(n)]
            assert b[0][0] == 0 and b[-1][1] == W
            assert all(b[i][1] == b[i + 1][0] for i in range(n - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1
            assert [tiled.shard_queries(W, n, r) for r in range(n)] == b
