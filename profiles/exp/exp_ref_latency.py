"""Single-scan latency of the reference loop (STRICT): 32x1024 scan against the 1M-point map, device ms of one
calculateAlignment (median of 20).  Library chosen by SSF_GPU_LIB.   python profiles/exp/exp_ref_latency.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np
import ssf_gpu
from ssf_gpu import synth
xyz, nrm, half = synth.make_map(1_000_000, normals=True)
T = synth.street_pose(100, half=half)
scan = synth.make_scan(T, 32, 1024, scan_id=100)
T0 = synth.perturb_pose(T, 100)
ctx = ssf_gpu.Context(0)
icp = ssf_gpu.ICPPointToPoint(0.5, 10, 0.05, 1e-5, mode=ssf_gpu.MODE_REFERENCE, reduce=ssf_gpu.REDUCE_STRICT, context=ctx)
icp.setTargetPointCloud(xyz)
icp.setSourcePointCloud(scan)
icp.setInitialTransformation(T0)
ms = []
for k in range(25):
    r = icp.calculateAlignment()
    ms.append(r.device_ms)
print(os.path.basename(os.environ.get("SSF_GPU_LIB", "libssf_gpu.so")), "median ms", float(np.median(ms[5:])), "iterations", r.iterations, "searches", r.n_searches)
