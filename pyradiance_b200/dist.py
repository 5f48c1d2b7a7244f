"""Multi-GPU plumbing: one process per GPU, sensor records sharded contiguously,
scene replicated, no exchange while tracing; the only collective is the final
gather of matrix rows.  This replaces the reference's process
fan-out (rt/rc3.c:329-413,579-654; rt/RcontribSimulManager.cpp:677-689,
820-854: fork + pipes, each child owning a disjoint range of records).

torch.distributed is used for rendezvous / barrier / gather only; all compute
goes through the C ABI.  RNG streams are keyed by the GLOBAL record index
(row_base), so the matrix does not depend on the number of ranks.
"""
from __future__ import annotations

import numpy as np


def shard_range(nrecords: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [r0, r1) of records owned by `rank` (blocks differ by
    at most one record)."""
    base, rem = divmod(nrecords, world)
    r0 = rank * base + min(rank, rem)
    r1 = r0 + base + (1 if rank < rem else 0)
    return r0, r1


def local_rays(rays: np.ndarray, accum: int, rank: int, world: int):
    """Slice of the ray list this rank traces, plus its record range."""
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
    nrec = (rays.shape[0] + accum - 1) // accum
    r0, r1 = shard_range(nrec, rank, world)
    return rays[r0 * accum:min(r1 * accum, rays.shape[0])], r0, r1


def rcontrib_sharded(ctx, rays, accum=1, flags=0, rank=0, world=1, dtype=np.float32):
    """Trace this rank's block of records; returns (rows [r1-r0, ncols, 3], r0, r1)."""
    mine, r0, r1 = local_rays(rays, accum, rank, world)
    if r1 > r0:
        m = ctx.rcontrib(mine, accum=accum, flags=flags, row_base=r0, dtype=dtype)
    else:
        m = np.zeros((0, ctx.num_columns(), 3), dtype=dtype)
    return m, r0, r1


def gather_rows(local_rows, nrecords: int, group=None, dst: int = 0, device=None):
    """Gather the row blocks of all ranks on `dst` (NCCL when `device` is a CUDA
    device: rows travel GPU->GPU over NVLink; gloo otherwise).  Returns the full
    [nrecords, ncols, 3] array on dst, None elsewhere."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    t = torch.from_numpy(np.ascontiguousarray(local_rows))
    if device is not None:
        t = t.to(device)
    shape_tail = tuple(t.shape[1:])
    sizes = [shard_range(nrecords, r, world) for r in range(world)]
    maxn = max(b - a for a, b in sizes)
    pad = torch.zeros((maxn,) + shape_tail, dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    if rank == dst:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.gather(pad, gather_list=bufs, dst=dst, group=group)
        out = torch.cat([bufs[r][:sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)
        return out.cpu().numpy()
    dist.gather(pad, gather_list=None, dst=dst, group=group)
    return None
