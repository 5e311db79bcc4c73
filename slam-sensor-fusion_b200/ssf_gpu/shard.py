"""Host-side logic of the two multi-GPU partitionings.

* scans sharded, map replicated (offline reprocessing, BASELINE config 4): ``scan_range``;
* map sharded by cell columns with a one-cell halo, scans replicated, one small all-reduce per
  iteration (configs 3 and 5): ``global_grid`` / ``partition_columns`` / ``shard_map`` and the
  ``setup_peer_exchange`` for the in-kernel exchange over CUDA IPC peer memory (default), ``setup_nccl``
  for ``ncclAllReduce`` behind the C ABI, and the ``torch_allreduce_hook`` that plugs
  ``torch.distributed`` into ``ssf_icp_set_allreduce``.

Pure numpy (+ optional torch for the hook); covered on CPU by tests/test_sharding_gloo.py.
"""
from __future__ import annotations

import ctypes

import numpy as np

INT32_MIN, INT32_MAX = -2**31, 2**31 - 1


def scan_range(n_scans: int, rank: int, world: int) -> range:
    """Contiguous block of scans for ``rank`` (blocks differ by at most one scan)."""
    base, extra = divmod(n_scans, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def global_grid(xyz_min, max_correspondence_dist: float):
    """(origin float32[3], cell edge float32) every rank must use: one cell covers the search radius."""
    origin = np.asarray(xyz_min, np.float32)[:3].copy()
    cell = np.float32(np.sqrt(np.float32(max_correspondence_dist)) * np.float32(1.01))
    return origin, cell


def column_of(x, origin_x, cell) -> np.ndarray:
    """Cell column of x coordinates with the device's float32 arithmetic: floor((x - o) * (1 / h))."""
    inv = np.float32(1.0) / np.float32(cell)
    u = (np.asarray(x, np.float32) - np.float32(origin_x)).astype(np.float32) * inv
    return np.floor(u.astype(np.float32)).astype(np.int64)


def partition_columns(cols: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Split the column axis into ``world`` ranges [lo, hi) balanced by point count.  The first
    range starts at -inf and the last ends at +inf so every query has exactly one owner."""
    cmin, cmax = int(cols.min()), int(cols.max())
    hist = np.bincount(cols - cmin, minlength=cmax - cmin + 1)
    cum = np.cumsum(hist)
    total = int(cum[-1])
    bounds = []
    for r in range(1, world):
        c = int(np.searchsorted(cum, total * r / world, side="left")) + 1 + cmin
        if bounds and c <= bounds[-1]:
            c = bounds[-1] + 1
        bounds.append(c)
    edges = [INT32_MIN] + bounds + [INT32_MAX]
    return [(edges[i], edges[i + 1]) for i in range(world)]


def shard_map(xyz: np.ndarray, normals, rank: int, world: int, max_correspondence_dist: float,
              origin=None, ranges=None):
    """Points of ``rank``'s shard: columns [lo - 1, hi + 1) (one-cell halo on both sides).

    Returns dict(points, normals, global_index int32, origin, cell, own=(lo, hi), ranges)."""
    if origin is None:
        origin = xyz[:, :3].min(0)
    origin, cell = global_grid(origin, max_correspondence_dist)
    cols = column_of(xyz[:, 0], origin[0], cell)
    if ranges is None:
        ranges = partition_columns(cols, world)
    lo, hi = ranges[rank]
    sel = np.nonzero((cols >= lo - 1) & (cols < hi + 1))[0]
    return dict(points=np.ascontiguousarray(xyz[sel]),
                normals=None if normals is None else np.ascontiguousarray(normals[sel]),
                global_index=sel.astype(np.int32), origin=origin, cell=cell, own=(lo, hi), ranges=ranges)


class ShardInfo(ctypes.Structure):
    _fields_ = [("origin", ctypes.c_float * 3), ("cell_size", ctypes.c_float), ("own_lo", ctypes.c_int32),
                ("own_hi", ctypes.c_int32)]


ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)


def setup_peer_exchange(icp, rank: int, world: int, max_scans: int, group=None) -> None:
    """In-kernel exchange of the per-scan sums (``ssf_icp_exchange_*``): every rank creates its
    buffer, the 64-byte CUDA IPC handles travel through ``torch.distributed.all_gather_object`` once,
    and from then on the kernels store into each other's memory over NVLink -- no per-iteration hook."""
    import torch.distributed as dist
    mine = icp.exchangeCreate(rank, world, max_scans)
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    icp.exchangeOpen(handles)
    dist.barrier(group=group)  # nobody starts storing before every rank has its buffer mapped


def setup_nccl(icp, rank: int, world: int, group=None) -> None:
    """The per-iteration sum through the library's own ``ncclAllReduce`` (``ssf_icp_nccl_init``): rank 0
    creates the unique id, ``torch.distributed`` only carries its 128 bytes to the other ranks."""
    import torch.distributed as dist
    import ssf_gpu
    box = [ssf_gpu.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    icp.ncclInit(box[0], rank, world)
    dist.barrier(group=group)


class _DevArray:
    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def torch_allreduce_hook(device_index: int, group=None):
    """ctypes callback for ssf_icp_set_allreduce: torch.distributed.all_reduce (NCCL) on the
    library's stream.  Keep the returned object alive as long as the handle uses it."""
    import torch
    import torch.distributed as dist

    def hook(_user, buf, count, stream):
        try:
            t = torch.as_tensor(_DevArray(int(buf), int(count)), device=torch.device("cuda", device_index))
            with torch.cuda.stream(torch.cuda.ExternalStream(int(stream), device=device_index)):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception as e:  # never let an exception cross the C boundary
            import sys
            print(f"[ssf_gpu] all-reduce hook failed: {e}", file=sys.stderr)
            return 1

    return ALLREDUCE_FN(hook)
