"""Experiment: search kernel alone vs. map cell edge (SSF_CELL_SIZE) and library variant (SSF_LIB)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np
from ssf_gpu import capi
if os.environ.get("SSF_LIB"):
    capi.LIB_PATH = capi.LIB_PATH.replace("libssf_gpu.so", os.environ["SSF_LIB"])
import ssf_gpu
from ssf_gpu import synth

xyz, nrm, half = synth.make_map(5_000_000, normals=False)
ctx = ssf_gpu.Context(0)
init_q, gt_q = [], []
for d in range(16):
    T = synth.street_pose(40 * d, half=half)
    v = ssf_gpu.voxel_down_sample(synth.make_scan(T, 64, 2048, scan_id=40 * d), 0.2, ctx)
    for r in range(4):
        T0 = synth.perturb_pose(T, 4 * d + r)
        init_q.append((v @ T0[:3, :3].T + T0[:3, 3]).astype(np.float32))
        gt_q.append((v @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
qi, qg = np.concatenate(init_q), np.concatenate(gt_q)
for cell in [float(c) for c in os.environ.get("CELLS", "0.25,0.3,0.4,0.5,0.72").split(",")]:
    os.environ["SSF_CELL_SIZE"] = str(cell)
    icp = ssf_gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, context=ctx)
    icp.setTargetPointCloud(xyz)
    a = icp.nearest_bench(qi, 0.5, 10); b = icp.nearest_bench(qg, 0.5, 10)
    print(f"{os.environ.get('SSF_LIB','libssf_gpu.so'):22s} cell {cell:5.2f}  initial {a*1e3:7.1f} us  converged {b*1e3:7.1f} us  ({len(qg)/b/1e6:5.2f} Gq/s)")
    del icp
