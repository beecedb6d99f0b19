"""Multi-GPU modes of the hot path (one process per GPU, torch.distributed for the plumbing).

* query-parallel: queries are independent units -> rank r takes a contiguous slice,
  the grid is replicated, there is NO collective on the data path (``shard_queries`` / ``plan_batch_sharded``).
* projection of one cloud (§8e row 2): points are independent -> every rank bins its slice of the cloud into a full
  grid and the grids are OR-ed with one all-reduce (MAX on uint8); bit-identical to one GPU (``project_sharded``).
* EDT on row tiles (§8e row 4): pass 1 (along y) is local to an x-slab, pass 2 (along x) needs whole columns -> one
  transpose of the uint16 row distances between ranks (grouped send/recv), the column pass on a y-block, and one
  transpose of the int32 result back (``edt_tiled``).
* row-tiled (cfg5): a grid too large for one GPU's scratch is cut into x-slabs (x-rows are contiguous in the
  ``[x][y]`` layout).  Inflation needs ONE halo exchange of ``radius`` rows; the single-source cost field needs an
  iterative exchange: every rank relaxes its slab to the local fixpoint (``fx_field_relax``, a Dial wavefront
  seeded by whatever improved), then neighbours swap their two boundary rows (NCCL send/recv over NVLink:
  2 x H x 4 bytes per direction), merge them with ``fx_halo_merge`` and an all-reduce of one flag decides
  termination.  The slab sub-graph (owned rows + one ghost row each side) is an exact sub-graph of the full
  grid graph, so every intermediate value is the cost of a real path and the fixpoint is the exact field:
  the stitched result is bit-identical to the single-GPU field (tests check this).

The reference has no multi-device mode at all; this replaces nothing in it but is what
BASELINE.json's configs 4 and 5 ask for.

The exchange protocol is backend-agnostic (NCCL on GPUs; the CPU tests drive it over gloo with a stand-in for
the two device ops), the compute is not: the default ops are the CUDA kernels and raise without a GPU.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import api
from ._lib import FuxiError


# ------------------------------------------------------------------------------------------ partitioning
def slab_bounds(W, nranks, rank):
    """Owned x-rows [x0, x1) of `rank`: contiguous, sizes differ by at most one row."""
    base, rem = divmod(int(W), int(nranks))
    x0 = rank * base + min(rank, rem)
    return x0, x0 + base + (1 if rank < rem else 0)


def shard_queries(Q, nranks, rank):
    """Contiguous slice [q0, q1) of the Q queries handled by `rank` (rank r takes [r*Q/G, (r+1)*Q/G))."""
    return slab_bounds(Q, nranks, rank)


def plan_batch_sharded(grid, starts, goals, metric=2, max_path=512, group=None, gather_costs=False, ctx=None):
    """Query-parallel batched planning: this rank answers its slice of (starts, goals) on its replica of the grid.
    Returns (PlanResult of the local slice, (q0, q1)).  With gather_costs=True also all-gathers the int32 costs
    (the only optional collective; Q*4 bytes) and returns them as a third element."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = dist.get_world_size(group) if dist.is_initialized() else 1
    Q = starts.shape[0]
    q0, q1 = shard_queries(Q, n, rank)
    res = api.plan_batch(grid, starts[q0:q1], goals[q0:q1], metric=metric, max_path=max_path, ctx=ctx)
    if not gather_costs:
        return res, (q0, q1)
    if n == 1:
        return res, (q0, q1), res.cost_i
    sizes = [shard_queries(Q, n, r) for r in range(n)]
    pad = max(b - a for a, b in sizes)
    mine = torch.full((pad,), -1, dtype=torch.int32, device=res.cost_i.device)
    mine[: q1 - q0] = res.cost_i
    bufs = [torch.empty_like(mine) for _ in range(n)]
    dist.all_gather(bufs, mine, group=group)
    allc = torch.cat([b[: e - s] for b, (s, e) in zip(bufs, sizes)])
    return res, (q0, q1), allc


# ------------------------------------------------------------------------------------------ device ops
class CudaOps:
    """The two device operations of the row-tiled loop, on the CUDA kernels."""

    def __init__(self, ctx=None):
        self.ctx = ctx

    def relax(self, grid, field, metric):
        changed = torch.zeros(1, dtype=torch.int32, device=field.device)
        api.field_relax(grid, field, metric, changed=changed, ctx=self.ctx, check=True)
        return changed

    def merge(self, dst_rows, src_rows, changed):
        ctx = api._ctx(self.ctx, dst_rows)
        if not (dst_rows.is_contiguous() and src_rows.is_contiguous()):
            raise FuxiError("halo rows must be contiguous")
        rc = ctx.lib.fx_halo_merge(ctx.handle, C.c_void_p(dst_rows.data_ptr()), C.c_void_p(src_rows.data_ptr()),
                                   dst_rows.numel(), C.c_void_p(changed.data_ptr()), api._stream())
        ctx.check(rc, "fx_halo_merge")

    def inflate(self, grid, radius, variant):
        return api.inflate(grid, radius, variant, ctx=self.ctx)

    def edt_rows(self, grid):
        """uint8 [w][H] -> int16 view of uint16 row distances (0xFFFF = none); fx_edt_rows"""
        ctx = api._ctx(self.ctx, grid)
        g = torch.empty(grid.shape, dtype=torch.int16, device=grid.device)
        ctx.check(ctx.lib.fx_edt_rows(ctx.handle, api._ptr(grid), api._ptr(g), grid.shape[0], grid.shape[1], api._stream()), "fx_edt_rows")
        return g

    def edt_cols(self, g):
        """row distances for whole columns [W][hb] -> int32 squared distances [W][hb]; fx_edt_cols"""
        ctx = api._ctx(self.ctx, g)
        out = torch.empty(g.shape, dtype=torch.int32, device=g.device)
        ctx.check(ctx.lib.fx_edt_cols(ctx.handle, api._ptr(g), api._ptr(out), g.shape[0], g.shape[1], api._stream()), "fx_edt_cols")
        return out

    def project(self, points, affine, zmin, zmax, origin, reso, shape):
        return api.project(points, affine, zmin, zmax, origin, reso, shape, ctx=self.ctx)


def _exchange(send_lo, send_hi, recv_lo, recv_hi, rank, n, group):
    """Swap boundary blocks with rank-1 (lo) and rank+1 (hi) in one batch of P2P ops."""
    ops = []
    if rank > 0:
        ops += [dist.P2POp(dist.isend, send_lo, dist.get_global_rank(group, rank - 1) if group else rank - 1, group),
                dist.P2POp(dist.irecv, recv_lo, dist.get_global_rank(group, rank - 1) if group else rank - 1, group)]
    if rank < n - 1:
        ops += [dist.P2POp(dist.isend, send_hi, dist.get_global_rank(group, rank + 1) if group else rank + 1, group),
                dist.P2POp(dist.irecv, recv_hi, dist.get_global_rank(group, rank + 1) if group else rank + 1, group)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


# ------------------------------------------------------------------------------------------ row-tiled inflation
def inflate_tiled(own_rows, radius, variant="ccst", group=None, ops=None):
    """Inflate a grid split into x-slabs: own_rows = this rank's owned rows [x1-x0, H] (uint8).  One halo exchange
    of `radius` rows with each neighbour, then the local kernel on slab + halo, cropped back to the owned rows.
    Equals the single-GPU fx_inflate of the whole grid bit for bit.  Every slab must hold >= radius rows."""
    ops = ops or CudaOps()
    rank, n = dist.get_rank(group), dist.get_world_size(group)
    r = int(radius)
    h = own_rows.shape[0]
    if r == 0 or n == 1:
        return ops.inflate(own_rows.contiguous(), r, variant)
    if h < r:
        raise FuxiError("slab of %d rows is thinner than the inflation radius %d" % (h, r))
    H = own_rows.shape[1]
    recv_lo = torch.zeros((r, H), dtype=own_rows.dtype, device=own_rows.device)
    recv_hi = torch.zeros((r, H), dtype=own_rows.dtype, device=own_rows.device)
    _exchange(own_rows[:r].contiguous(), own_rows[h - r:].contiguous(), recv_lo, recv_hi, rank, n, group)
    parts = ([recv_lo] if rank > 0 else []) + [own_rows] + ([recv_hi] if rank < n - 1 else [])
    out = ops.inflate(torch.cat(parts).contiguous(), r, variant)
    lo = r if rank > 0 else 0
    return out[lo:lo + h]


# ------------------------------------------------------------------------------------------ row-tiled cost field
def field_tiled(own_rows, W, source, metric=1, group=None, ops=None, max_rounds=100000):
    """Single-source cost field of a grid split into x-slabs over the ranks of `group`.

    own_rows: uint8 [x1-x0, H], this rank's owned rows (slab_bounds(W, nranks, rank)).
    Returns (field int32 [x1-x0, H] for the owned rows (-1 unreachable), rounds) -- rounds = number of
    relax+exchange rounds until the global fixpoint.
    """
    ops = ops or CudaOps()
    rank, n = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    x0, x1 = slab_bounds(W, n, rank)
    h, H = own_rows.shape
    if h != x1 - x0:
        raise FuxiError("own_rows has %d rows, slab_bounds says %d" % (h, x1 - x0))
    if n > 1 and h < 2:
        raise FuxiError("row-tiled search needs at least two rows per slab")
    dev = own_rows.device
    has_lo, has_hi = rank > 0, rank < n - 1
    # 1) ghost rows of the GRID (one exchange): neighbour's nearest owned row
    g_lo = torch.zeros((1, H), dtype=own_rows.dtype, device=dev)
    g_hi = torch.zeros((1, H), dtype=own_rows.dtype, device=dev)
    if n > 1:
        _exchange(own_rows[:1].contiguous(), own_rows[h - 1:].contiguous(), g_lo, g_hi, rank, n, group)
    slab = torch.cat(([g_lo] if has_lo else []) + [own_rows] + ([g_hi] if has_hi else [])).contiguous()
    off = 1 if has_lo else 0                     # slab row of owned row 0
    field = torch.full(slab.shape, -1, dtype=torch.int32, device=dev)
    sx, sy = int(source[0]), int(source[1])
    if x0 <= sx < x1:
        field[sx - x0 + off, sy] = 0
    # 2) relax / exchange until nobody changes.  Both rows next to a cut live on both ranks (owned on one side,
    #    ghost on the other) and both ranks improve both, so each side sends its copy of the pair and merges by min.
    rounds = 0
    recv_lo = torch.empty((2, H), dtype=torch.int32, device=dev)
    recv_hi = torch.empty((2, H), dtype=torch.int32, device=dev)
    L = slab.shape[0]
    while True:
        rounds += 1
        if rounds > max_rounds:
            raise FuxiError("row-tiled field did not converge in %d rounds" % max_rounds)
        changed = ops.relax(slab, field, metric)
        if n == 1:
            break
        pair_lo = field[0:2] if has_lo else None           # (ghost, first owned)
        pair_hi = field[L - 2:L] if has_hi else None       # (last owned, ghost)
        _exchange(pair_lo, pair_hi, recv_lo, recv_hi, rank, n, group)
        if has_lo:
            ops.merge(pair_lo, recv_lo, changed)
        if has_hi:
            ops.merge(pair_hi, recv_hi, changed)
        flag = changed.to(torch.int32).reshape(1).clone()
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()) == 0:
            break
    return field[off:off + h], rounds


# ------------------------------------------------------------------------------------------ sharded projection
def project_sharded(points, affine=None, zmin=0.3, zmax=float("inf"), origin=(0.0, 0.0), reso=0.2, shape=None, group=None,
                    ops=None):
    """points: THIS rank's slice of the cloud.  Every rank ends up with the full grid: local fx_project, then one
    all-reduce (MAX) over the uint8 grids -- occupancy is a set union, so the result is bit-identical to projecting
    the whole cloud on one GPU."""
    ops = ops or CudaOps()
    grid = ops.project(points, affine, zmin, zmax, origin, reso, shape)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grid, op=dist.ReduceOp.MAX, group=group)
    return grid


# ------------------------------------------------------------------------------------------ row-tiled EDT
def _peer(group, r):
    return dist.get_global_rank(group, r) if group else r


def _transpose_blocks(send, recv, rank, n, group):
    """send[j] goes to rank j, recv[i] comes from rank i (one batch of P2P ops; the own block is copied locally)."""
    recv[rank].copy_(send[rank])
    ops = []
    for k in range(1, n):
        to, frm = (rank + k) % n, (rank - k) % n
        # byte views: NCCL has no 16-bit integer type, and the payload is opaque to the transport anyway
        ops.append(dist.P2POp(dist.isend, send[to].view(torch.uint8), _peer(group, to), group))
        ops.append(dist.P2POp(dist.irecv, recv[frm].view(torch.uint8), _peer(group, frm), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def edt_tiled(own_rows, W, group=None, ops=None):
    """Exact squared EDT of a grid split into x-slabs: own_rows = uint8 [x1-x0, H] -> int32 [x1-x0, H], bit-identical
    to fx_edt of the whole grid.  Two all-to-all transposes (uint16 in: W*H*2/n bytes per rank, int32 back)."""
    ops = ops or CudaOps()
    rank, n = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    x0, x1 = slab_bounds(W, n, rank)
    h, H = own_rows.shape
    if h != x1 - x0:
        raise FuxiError("own_rows has %d rows, slab_bounds says %d" % (h, x1 - x0))
    if (W - 1) ** 2 + (H - 1) ** 2 >= 2 ** 31 - 1:          # int32 squared distances, INT32_MAX = no obstacle (fx_edt's own check)
        raise FuxiError("edt_tiled: (W-1)^2 + (H-1)^2 must be < 2^31 - 1")
    g = ops.edt_rows(own_rows.contiguous())
    if n == 1:
        return ops.edt_cols(g)
    dev = own_rows.device
    xb = [slab_bounds(W, n, r) for r in range(n)]
    yb = [slab_bounds(H, n, r) for r in range(n)]
    y0, y1 = yb[rank]
    send = [g[:, a:b].contiguous() for a, b in yb]
    recv = [torch.empty((b - a, y1 - y0), dtype=g.dtype, device=dev) for a, b in xb]
    _transpose_blocks(send, recv, rank, n, group)
    d = ops.edt_cols(torch.cat(recv).contiguous())          # [W][y1-y0]
    send = [d[a:b].contiguous() for a, b in xb]
    recv = [torch.empty((h, b - a), dtype=d.dtype, device=dev) for a, b in yb]
    _transpose_blocks(send, recv, rank, n, group)
    return torch.cat(recv, dim=1).contiguous()
