"""Profile target: the search kernel alone on converged (or initial) bench queries."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np
import ssf_gpu
from ssf_gpu import synth

which = os.environ.get("WHICH", "converged")
xyz, nrm, half = synth.make_map(5_000_000, normals=False)
ctx = ssf_gpu.Context(0)
icp = ssf_gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, context=ctx)
icp.setTargetPointCloud(xyz)
qs = []
for d in range(16):
    T = synth.street_pose(40 * d, half=half)
    v = ssf_gpu.voxel_down_sample(synth.make_scan(T, 64, 2048, scan_id=40 * d), 0.2, ctx)
    for r in range(4):
        T0 = synth.perturb_pose(T, 4 * d + r) if which == "initial" else T
        qs.append((v @ T0[:3, :3].T + T0[:3, 3]).astype(np.float32))
q = np.concatenate(qs)
ms = icp.nearest_bench(q, 0.5, int(os.environ.get("REPS", "3")))
print(which, len(q), ms * 1e3, "us")
