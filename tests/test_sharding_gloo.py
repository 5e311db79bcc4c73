"""Host-side logic of the N > 1 paths on CPU: torch.distributed with the gloo backend, world
size 2 (MASTER_ADDR 127.0.0.1).  No GPU: the per-rank partial sums come from the CPU oracle on
each rank's map shard, exchanged with one all_reduce, and must equal the unsharded result --
the same partition / halo / ownership rules the CUDA path uses (ssf_gpu/shard.py, ssf.h)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT
from ssf_gpu import shard


def test_scan_range_covers_every_scan_once():
    for n, w in [(10, 3), (64, 8), (5, 8), (0, 2), (10000, 8)]:
        seen = [i for r in range(w) for i in shard.scan_range(n, r, w)]
        assert seen == list(range(n))
        sizes = [len(shard.scan_range(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_partition_ownership_and_halo(small_world):
    xyz = small_world["map"]
    thr = 0.5
    origin, cell = shard.global_grid(xyz[:, :3].min(0), thr)
    cols = shard.column_of(xyz[:, 0], origin[0], cell)
    for world in (2, 4, 8):
        ranges = shard.partition_columns(cols, world)
        assert ranges[0][0] == shard.INT32_MIN and ranges[-1][1] == shard.INT32_MAX
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))  # contiguous, disjoint
        counts = []
        for r in range(world):
            s = shard.shard_map(xyz, None, r, world, thr)
            lo, hi = s["own"]
            c = shard.column_of(s["points"][:, 0], origin[0], cell)
            assert ((c >= lo - 1) & (c < hi + 1)).all()
            owned = (cols >= lo) & (cols < hi)
            counts.append(int(owned.sum()))
            # halo: every map point within sqrt(thr) in x of an owned column is in the shard
            assert set(np.nonzero((cols >= lo - 1) & (cols < hi + 1))[0]) == set(s["global_index"].tolist())
        assert sum(counts) == xyz.shape[0]
        assert max(counts) < 2.0 * xyz.shape[0] / world + 2000  # balanced by point count


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, PKG)
    from oracle import oracle
    from ssf_gpu import shard as sh
    from ssf_gpu import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xyz, nrm, half = synth.make_map(65536, normals=True)
    T_gt = synth.street_pose(3, half=half)
    scan = synth.make_scan(T_gt, beams=16, azimuths=256, scan_id=3, max_range=60.0)
    T0 = synth.perturb_pose(T_gt, 3).astype(np.float32)
    thr = np.float32(0.5)
    s = sh.shard_map(xyz, nrm, rank, world, float(thr))
    q = (scan[:, :3] @ T0[:3, :3].T + T0[:3, 3]).astype(np.float32)
    qcol = sh.column_of(q[:, 0], s["origin"][0], s["cell"])
    mine = (qcol >= s["own"][0]) & (qcol < s["own"][1])
    idx, d2 = oracle.KdTree(s["points"]).nn(q[mine])
    ok = d2 < thr
    gidx = s["global_index"][idx[ok]]
    # partial sums of this rank: K, sum d2, sum of global indices (a checksum of the correspondences)
    part = torch.tensor([ok.sum(), float(d2[ok].astype(np.float64).sum()), float(gidx.astype(np.float64).sum()),
                         float(mine.sum())], dtype=torch.float64)
    dist.all_reduce(part)
    if rank == 0:
        fi, fd = oracle.KdTree(xyz).nn(q)
        fok = fd < thr
        ref = np.array([fok.sum(), fd[fok].astype(np.float64).sum(), fi[fok].astype(np.float64).sum(), q.shape[0]])
        np.save(os.path.join(out_dir, "result.npy"), np.stack([part.numpy(), ref]))
    dist.barrier()
    dist.destroy_process_group()


def test_map_sharded_sums_equal_unsharded_gloo(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got, ref = np.load(tmp_path / "result.npy")
    assert got[0] == ref[0] and got[3] == ref[3]      # every query owned once, same number of correspondences
    assert got[2] == ref[2]                            # same correspondences (global indices)
    assert abs(got[1] - ref[1]) <= 1e-9 * max(1.0, ref[1])
