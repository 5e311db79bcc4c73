"""Map-sharded registration on N GPUs, checked against the unsharded run on one GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/multigpu/run_sharded.py [--map-points 1000000]

Every rank holds one column range of the map (+ one-cell halo), all ranks see all scans, each
rank searches the queries it owns and one all_reduce (NCCL) per iteration sums the 32-double
rows.  Rank 0 also runs the same scans against the whole map and compares: correspondences
must be identical (global indices), poses equal to float rounding.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))


def main():
    import torch
    import torch.distributed as dist
    import ssf_gpu
    from ssf_gpu import shard, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--map-points", type=int, default=1_000_000)
    ap.add_argument("--mode", default="p2plane", choices=["p2plane", "p2p", "o3d"])
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer", "hook"],
                    help="nccl: ncclAllReduce behind the C ABI; peer: in-kernel exchange over CUDA IPC peer memory; "
                         "hook: the caller's all-reduce callback (torch.distributed)")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    xyz, nrm, half = synth.make_map(args.map_points, normals=True)
    scans, inits = [], []
    for k in range(6):
        T = synth.street_pose(int(half / 0.15) + 35 * k - 100, half=half)
        scans.append(synth.make_scan(T, 32, 512, scan_id=k, max_range=80.0))
        inits.append(synth.perturb_pose(T, k))
    mode = {"p2plane": ssf_gpu.MODE_GN_P2PLANE, "p2p": ssf_gpu.MODE_GN_P2P, "o3d": ssf_gpu.MODE_O3D_P2P}[args.mode]
    ctx = ssf_gpu.Context(local)
    icp = ssf_gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, mode=mode, context=ctx)
    s = shard.shard_map(xyz, nrm, rank, world, 0.5)
    icp.setTargetShard(s)
    if args.exchange == "peer":
        shard.setup_peer_exchange(icp, rank, world, max_scans=16)
    elif args.exchange == "nccl":
        shard.setup_nccl(icp, rank, world)
    else:
        icp.setAllreduce(shard.torch_allreduce_hook(local))
    res = icp.align_batch(scans, inits)
    # per-row correspondences of the first scan from every rank (-2 = not owned)
    icp.setSourcePointCloud(scans[0])
    icp.setInitialTransformation(inits[0])
    r0 = icp.calculateAlignment()
    corr = torch.from_numpy(icp.correspondences().astype(np.int64)).cuda()
    dist.all_reduce(corr, op=dist.ReduceOp.MAX)
    owned = torch.tensor([float((icp.correspondences() != -2).sum())], device="cuda", dtype=torch.float64)
    dist.all_reduce(owned)
    ok = True
    if rank == 0:
        full = ssf_gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, mode=mode, context=ctx)
        full.setTargetPointCloud(xyz, nrm)
        ref = full.align_batch(scans, inits)
        full.setSourcePointCloud(scans[0])
        full.setInitialTransformation(inits[0])
        f0 = full.calculateAlignment()
        fcorr = full.correspondences()
        dmax = max(float(np.abs(a.transformation - b.transformation).max()) for a, b in zip(res, ref))
        same_k = all(a.k_final == b.k_final and a.iterations == b.iterations for a, b in zip(res, ref))
        same_corr = bool(np.array_equal(corr.cpu().numpy(), fcorr.astype(np.int64)))
        ok = dmax < 1e-5 and same_k and same_corr and int(owned.item()) == scans[0].shape[0]
        print(f"[sharded x{world}, {args.mode}] shard sizes ~{s['points'].shape[0]} of {xyz.shape[0]}, "
              f"max |dT| vs unsharded = {dmax:.2e}, same K/iterations: {same_k}, correspondences identical: {same_corr}, "
              f"every query owned once: {int(owned.item()) == scans[0].shape[0]} -> {'OK' if ok else 'FAIL'}", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
