"""Experiment: distribution of the true NN distance of the bench queries (who is unmatched, and by how much)."""
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
mn, mx = xyz[:, :3].min(0), xyz[:, :3].max(0)
for which in ("initial", "converged"):
    qs = []
    for d in range(16):
        T = synth.street_pose(40 * d, half=half)
        v = ssf_gpu.voxel_down_sample(synth.make_scan(T, 64, 2048, scan_id=40 * d), 0.2, ctx)
        for r in range(4):
            T0 = synth.perturb_pose(T, 4 * d + r) if which == "initial" else T
            qs.append((v @ T0[:3, :3].T + T0[:3, 3]).astype(np.float32))
    q = np.concatenate(qs)
    idx, d2 = icp.nearest(q, 100.0)
    d = np.sqrt(np.where(idx >= 0, d2, 1e4))
    inside = np.all((q >= mn) & (q <= mx), axis=1)
    edges = [0, 0.05, 0.1, 0.15, 0.2, 0.3, 0.4, 0.5, 0.707, 1.0, 1.5, 2.5, 5, 10, 1e9]
    hist = np.histogram(d, edges)[0] / len(d)
    print(which, "n", len(q), "outside bbox", 1 - inside.mean())
    print("  ", " ".join(f"<{e:g}:{h*100:.1f}%" for e, h in zip(edges[1:], hist)))
    far = d >= 0.707
    print("   unmatched", far.mean(), "of which outside bbox", (far & ~inside).sum() / max(1, far.sum()))
