"""Experiment: the search kernel alone (no accumulation) on different query sets."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np
import ssf_gpu
from ssf_gpu import synth

xyz, nrm, half = synth.make_map(5_000_000, normals=False)
ctx = ssf_gpu.Context(0)
icp = ssf_gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, context=ctx)
icp.setTargetPointCloud(xyz)
h = np.float32(np.sqrt(np.float32(0.5)) * np.float32(1.01)); o = xyz[:, :3].min(0)
def cellkey(pw):
    c = np.floor((pw - o) / h).astype(np.int64); return (c[:, 2] * 4096 + c[:, 1]) * 4096 + c[:, 0]
init_q, gt_q = [], []
for d in range(16):
    T = synth.street_pose(40 * d, half=half)
    v = ssf_gpu.voxel_down_sample(synth.make_scan(T, 64, 2048, scan_id=40 * d), 0.2, ctx)
    for r in range(4):
        T0 = synth.perturb_pose(T, 4 * d + r)
        init_q.append((v @ T0[:3, :3].T + T0[:3, 3]).astype(np.float32))
        gt_q.append((v @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
for name, qs in (("initial poses", init_q), ("converged poses", gt_q)):
    q = np.concatenate(qs)
    per_scan_sorted = np.concatenate([a[np.argsort(cellkey(a), kind="stable")] for a in qs])
    all_sorted = q[np.argsort(cellkey(q), kind="stable")]
    rng = np.random.default_rng(0); shuffled = q[rng.permutation(len(q))]
    for tag, arr in (("scan order", q), ("sorted per scan", per_scan_sorted), ("sorted across scans", all_sorted), ("shuffled", shuffled)):
        ms = icp.nearest_bench(arr, 0.5, 10)
        print(f"{name:16s} {tag:20s} n={len(arr)}  {ms*1e3:7.1f} us/pass  {len(arr)/ms/1e6:6.2f} Gq/s")
