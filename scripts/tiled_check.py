"""Multi-GPU check + timing of the row-tiled and query-parallel modes (run under torchrun, one rank per GPU):

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node N scripts/tiled_check.py [--size 8192] [--out file.json]

* row-tiled single-source field (BASELINE cfg5 shape: grid default_rng(6), 20 % fill, first free -> last free cell)
  over N x-slabs with NCCL halo exchange; rank 0 also computes the single-GPU field (when the grid fits its scratch)
  and every rank compares its slab bit for bit;
* row-tiled inflation (radius 2 dense, radius 3 sparse) against the single-GPU kernel;
* query-parallel batch: costs gathered from the shards equal the single-GPU batch.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=8192)
    ap.add_argument("--queries", type=int, default=1024)
    ap.add_argument("--out", default="")
    ap.add_argument("--no-verify", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import fuxi_planner_b200 as fx
    from fuxi_planner_b200 import tiled

    n = a.n
    m = (np.random.default_rng(6).random((n, n)) < 0.2).astype(np.uint8)
    free = np.argwhere(m == 0)
    src, dst = tuple(int(v) for v in free[0]), tuple(int(v) for v in free[-1])
    x0, x1 = tiled.slab_bounds(n, world, rank)
    own = torch.from_numpy(m[x0:x1]).to(dev)
    res = {"n": n, "world": world}

    for metric in (1, 2):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        fld, rounds = tiled.field_tiled(own, n, src, metric)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        goal_cost = torch.tensor([int(fld[dst[0] - x0, dst[1]]) if x0 <= dst[0] < x1 else -1], device=dev)
        dist.all_reduce(goal_cost, op=dist.ReduceOp.MAX)
        reached = torch.tensor([int((fld >= 0).sum())], device=dev, dtype=torch.int64)
        dist.all_reduce(reached)
        ok = None
        if not a.no_verify:
            full = fx.field(torch.from_numpy(m).to(dev), src, metric)        # every rank: the single-GPU field
            ok_t = torch.tensor([int(torch.equal(full[x0:x1], fld))], device=dev)
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
            ok = bool(ok_t.item())
            del full
        res["field_metric%d" % metric] = {"seconds": dt, "rounds": rounds, "goal_cost": int(goal_cost.item()),
                                          "cells_reached": int(reached.item()), "nodes_per_s": int(reached.item()) / dt,
                                          "bit_exact_vs_single_gpu": ok}
        del fld
    for radius, variant in ((2, "ccst"), (3, "st")):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        inf = tiled.inflate_tiled(own, radius, variant)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        full = fx.inflate(torch.from_numpy(m).to(dev), radius, variant)
        ok_t = torch.tensor([int(torch.equal(full[x0:x1], inf))], device=dev)
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        res["inflate_r%d_%s" % (radius, variant)] = {"seconds": dt, "bit_exact_vs_single_gpu": bool(ok_t.item())}
        del full, inf
    # row-tiled EDT (two transposes) and point-sharded projection (one all-reduce) against the single-GPU kernels
    sparse = torch.from_numpy((np.random.default_rng(8).random((n, n)) < 0.002).astype(np.uint8))
    for name, grid in (("edt_20pct", torch.from_numpy(m)), ("edt_0.2pct", sparse)):
        own_e = grid[x0:x1].to(dev)
        tiled.edt_tiled(own_e, n)                      # untimed: the first all-to-all sets up the NCCL peer connections
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        d = tiled.edt_tiled(own_e, n)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        full = fx.edt(grid.to(dev))
        ok_t = torch.tensor([int(torch.equal(full[x0:x1], d))], device=dev)
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        res[name] = {"seconds": dt, "bit_exact_vs_single_gpu": bool(ok_t.item())}
        del full, d, own_e
    npts = 1 << 24
    prng = np.random.default_rng(9)
    pts = torch.from_numpy(np.c_[prng.uniform(0, n * 0.2, (npts, 2)), prng.uniform(-0.5, 3.0, npts)].astype(np.float32))
    p0, p1 = tiled.shard_queries(npts, world, rank)
    mine = pts[p0:p1].to(dev)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    gr = tiled.project_sharded(mine, None, 0.3, float("inf"), (0.0, 0.0), 0.2, (n, n))
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    full = fx.project(pts.to(dev), None, 0.3, float("inf"), (0.0, 0.0), 0.2, (n, n))
    ok_t = torch.tensor([int(torch.equal(full, gr))], device=dev)
    dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    res["project_sharded_16Mpts"] = {"seconds": dt, "bit_exact_vs_single_gpu": bool(ok_t.item()), "occupied": int(full.sum())}
    del full, gr, mine
    # query-parallel
    rng = np.random.default_rng(7)
    Q = a.queries
    s = torch.from_numpy(free[rng.integers(len(free), size=Q)].astype(np.int32)).to(dev)
    g = torch.from_numpy(free[rng.integers(len(free), size=Q)].astype(np.int32)).to(dev)
    gm = torch.from_numpy(m).to(dev)
    _, (q0, q1), allc = tiled.plan_batch_sharded(gm, s, g, metric=1, max_path=0, gather_costs=True)
    single = fx.plan_batch(gm, s, g, metric=1, max_path=0).cost_i
    res["query_parallel"] = {"Q": Q, "shard": [q0, q1], "costs_equal_single_gpu": bool(torch.equal(allc, single)),
                             "answered": int((single >= 0).sum())}
    if rank == 0:
        line = json.dumps(res)
        print(line, flush=True)
        if a.out:
            os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
            open(a.out, "w").write(line + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
