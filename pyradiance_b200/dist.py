"""Multi-GPU plumbing: one process per GPU, sensor records sharded contiguously,
scene replicated, no exchange while tracing; the only collective is the final
gather of matrix rows.  This replaces the reference's process
fan-out (rt/rc3.c:329-413,579-654; rt/RcontribSimulManager.cpp:677-689,
820-854: fork + pipes, each child owning a disjoint range of records, ONE
caller receiving the `[nrows][ncols*3]` array).

torch.distributed is used for rendezvous / barrier / handle exchange (and the
NCCL form of the gather); all compute goes through the C ABI.  RNG streams are
keyed by the GLOBAL record index (row_base), so the matrix does not depend on
the number of ranks.

Three ways for the rows to reach the gathering rank, fastest first:

  RowWindow         the whole matrix lives in the gathering GPU's HBM and is
                    mapped into every rank (CUDA IPC): the kernel that finishes a
                    batch of records stores its rows straight into that memory
                    over NVLink while the next batch is being traced.  No gather
                    step remains; a barrier ends the job.
  gather_rows       each rank keeps its block in its own HBM; exact-size grouped
                    NCCL send / recv into slices of one destination tensor (no
                    padding, no host bounce).  Also the gloo path of the CPU tests.
  SharedHostMatrix  every rank copies its rows D2H into its slice of ONE pinned
                    host array in POSIX shared memory (N PCIe links in parallel):
                    what a caller that wants the matrix in host memory should use
                    once the matrix is large (27.7 GB for BASELINE configs[2]).
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


def rcontrib_sharded(ctx, rays, accum=1, flags=0, rank=0, world=1, dtype=np.float32, out=None):
    """Trace this rank's block of records; returns (rows [r1-r0, ncols, 3], r0, r1).
    `out` (optional) = where the rows go, e.g. this rank's slice of a SharedHostMatrix."""
    mine, r0, r1 = local_rays(rays, accum, rank, world)
    if r1 > r0:
        m = ctx.rcontrib(mine, accum=accum, flags=flags, row_base=r0, dtype=dtype, out=out)
    else:
        m = np.zeros((0, ctx.num_columns(), 3), dtype=dtype) if out is None else out
    return m, r0, r1


def gather_rows(local_rows, nrecords: int, group=None, dst: int = 0, device=None, out=None):
    """Row blocks of all ranks -> one [nrecords, ...] array on `dst`; None elsewhere.

    `local_rows` is this rank's block [r1 - r0, ...] of the shard_range() split: a torch
    tensor already in HBM (NCCL: the rows travel GPU -> GPU over NVLink, straight from
    where the engine wrote them) or a numpy array (gloo in the CPU tests; with `device`
    it is uploaded first, which only a caller without device-resident rows needs).
    Exact sizes: every sender posts one send of its block, the gatherer one receive
    per peer directly into that peer's slice of the result -- no padding, no
    concatenation, no host bounce.  `out` (dst only, optional) receives the rows.
    Returns the type it was given (tensor -> tensor on its device, numpy -> numpy)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    was_numpy = not isinstance(local_rows, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(local_rows)) if was_numpy else local_rows.contiguous()
    if device is not None and t.device != torch.device(device):
        t = t.to(device)
    r0, r1 = shard_range(nrecords, rank, world)
    assert t.shape[0] == r1 - r0, f"rank {rank} holds {t.shape[0]} rows, its block is [{r0}, {r1})"
    if rank != dst:
        if r1 > r0:
            for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, t, dst, group)]):
                q.wait()
        return None
    if out is None:
        out = torch.empty((nrecords,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    elif not isinstance(out, torch.Tensor):
        out = torch.from_numpy(out)
    ops = []
    for r in range(world):
        a, b = shard_range(nrecords, r, world)
        if r != dst and b > a:
            ops.append(dist.P2POp(dist.irecv, out[a:b], r, group))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    if r1 > r0 and out[r0:r1].data_ptr() != t.data_ptr():
        out[r0:r1].copy_(t)
    for q in reqs:
        q.wait()
    return out.cpu().numpy() if was_numpy else out


def _bcast_object(obj, src, group):
    import torch.distributed as dist
    box = [obj]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


class RowWindow:
    """The whole [nrecords, ncols, 3] matrix in the gathering rank's HBM, mapped into every
    rank's address space (C ABI rb_ipc_export / rb_ipc_open, CUDA IPC with peer access): each
    rank's engine writes its finished rows there directly, over NVLink, batch by batch."""

    def __init__(self, ctx, nrecords: int, ncols: int, group=None, dst: int = 0, dtype=np.float32):
        import torch.distributed as dist
        self.ctx, self.group, self.dst = ctx, group, dst
        self.nrecords, self.ncols, self.dtype = int(nrecords), int(ncols), np.dtype(dtype)
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.row_bytes = self.ncols * 3 * self.dtype.itemsize
        self.owner = self.rank == dst
        handle = None
        if self.owner:
            self.base = ctx.device_alloc(max(256, self.nrecords * self.row_bytes))
            handle = ctx.ipc_export(self.base)
        handle = _bcast_object(handle, dst, group)
        if not self.owner:
            self.base = ctx.ipc_open(handle)

    def ptr(self, row: int) -> int:
        return self.base + int(row) * self.row_bytes

    def rcontrib(self, rays, nrays=None, accum=1, flags=0, rays_on_device=False):
        """Trace this rank's block of the record list straight into the window.
        rays: the WHOLE [n, 6] numpy ray list (each rank slices its block), or, with
        rays_on_device, a raw device pointer to THIS rank's block of `nrays` rays."""
        from . import _lib
        r0, r1 = shard_range(self.nrecords, self.rank, self.world)
        if r1 <= r0:
            return r0, r1
        f = int(flags) | _lib.RB_FLAG_OUT_ON_DEVICE | (_lib.RB_FLAG_OUT_DOUBLE if self.dtype == np.float64 else 0)
        nflt = (r1 - r0) * self.ncols * 3
        if rays_on_device:
            self.ctx._ck(self.ctx.lib.rb_rcontrib(self.ctx.h, rays, int(nrays), int(accum), f | _lib.RB_FLAG_RAYS_ON_DEVICE,
                                                  r0, self.ptr(r0), nflt))
        else:
            mine, a, b = local_rays(rays, accum, self.rank, self.world)
            assert (a, b) == (r0, r1)
            self.ctx._ck(self.ctx.lib.rb_rcontrib(self.ctx.h, mine.ctypes.data, mine.shape[0], int(accum), f, r0,
                                                  self.ptr(r0), nflt))
        return r0, r1

    def complete(self):
        """Every rank's rows are in the window once all ranks passed this point."""
        import torch.distributed as dist
        self.ctx.sync()
        dist.barrier(self.group)

    def as_tensor(self):
        """dst only: the matrix as a torch tensor over the window's memory (no copy; CUDA array interface)."""
        if not self.owner:
            return None
        import torch

        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = {"shape": (self.nrecords, self.ncols, 3), "typestr": self.dtype.str,
                                      "data": (int(self.base), False), "version": 2, "strides": None}
        return torch.as_tensor(v, device="cuda")

    def download(self, out=None):
        """dst only: the matrix as a host array (one D2H over the gathering GPU's link)."""
        if not self.owner:
            return None
        if out is None:
            out = np.empty((self.nrecords, self.ncols, 3), dtype=self.dtype)
        self.ctx.device_download(out, self.base)
        return out

    def close(self):
        import torch.distributed as dist
        if getattr(self, "base", None) is None:
            return
        if not self.owner:
            self.ctx.ipc_close(self.base)
        dist.barrier(self.group)          # nobody has the window mapped any more
        if self.owner:
            self.ctx.device_free(self.base)
        self.base = None


class SharedHostMatrix:
    """One [nrecords, ncols, 3] host array in POSIX shared memory, page-locked in every rank:
    each rank's engine copies its rows D2H into its own slice (N PCIe links in parallel), the
    gathering rank reads the whole matrix in place.  `ctx` may be None (no pinning; CPU tests)."""

    def __init__(self, ctx, nrecords: int, ncols: int, group=None, dst: int = 0, dtype=np.float32):
        import torch.distributed as dist
        from multiprocessing import shared_memory
        self.ctx, self.group, self.dst = ctx, group, dst
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.owner = self.rank == dst
        shape = (int(nrecords), int(ncols), 3)
        nbytes = max(16, int(np.prod(shape)) * np.dtype(dtype).itemsize)
        name = None
        if self.owner:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name = self.shm.name
        name = _bcast_object(name, dst, group)
        if not self.owner:
            self.shm = shared_memory.SharedMemory(name=name)
            try:        # the creator unlinks; keep this process's resource tracker out of it
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:       # noqa: BLE001
                pass
        self.array = np.ndarray(shape, dtype=dtype, buffer=self.shm.buf)
        self.pinned = False
        if ctx is not None and self.array.nbytes:
            ctx.pin(self.array)
            self.pinned = True

    def rows(self):
        r0, r1 = shard_range(self.array.shape[0], self.rank, self.world)
        return self.array[r0:r1], r0, r1

    def complete(self):
        import torch.distributed as dist
        dist.barrier(self.group)

    def close(self):
        import torch.distributed as dist
        if self.shm is None:
            return
        if self.pinned:
            self.ctx.unpin(self.array)
        self.array = None
        dist.barrier(self.group)
        self.shm.close()
        if self.owner:
            self.shm.unlink()
        self.shm = None


def rcontrib_gathered(ctx, rays, accum=1, flags=0, group=None, dst: int = 0, dtype=np.float32, how="auto"):
    """The distributed form of ctx.rcontrib(): every rank calls it with the SAME ray list; rank
    `dst` gets the [nrecords, ncols, 3] matrix (numpy), the others None -- the semantics of the
    reference's `rcontrib -n N` (one caller, one array).  how = "window" (RowWindow + one D2H),
    "host" (SharedHostMatrix) or "auto" (host above 1 GiB, window below)."""
    import torch.distributed as dist
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
    nrec = (rays.shape[0] + accum - 1) // accum
    ncols = ctx.num_columns()
    if how == "auto":
        how = "host" if nrec * ncols * 3 * np.dtype(dtype).itemsize > (1 << 30) else "window"
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if how == "window":
        w = RowWindow(ctx, nrec, ncols, group, dst, dtype)
        try:
            w.rcontrib(rays, accum=accum, flags=flags)
            w.complete()
            return w.download()
        finally:
            w.close()
    h = SharedHostMatrix(ctx, nrec, ncols, group, dst, dtype)
    try:
        mine, r0, r1 = h.rows()
        rcontrib_sharded(ctx, rays, accum, flags, rank, world, dtype, out=mine if r1 > r0 else None)
        h.complete()
        return h.array.copy() if rank == dst else None
    finally:
        h.close()
