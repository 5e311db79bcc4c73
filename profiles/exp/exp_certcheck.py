import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np
import ssf_gpu as gpu
from ssf_gpu import synth
xyz, nrm, half = synth.make_map(65536, normals=True)
T_gt = synth.street_pose(3, half=half)
scan = synth.make_scan(T_gt, beams=16, azimuths=512, scan_id=3, max_range=60.0)
T0 = synth.perturb_pose(T_gt, 3)
icp = gpu.ICPPointToPoint(0.5, 6, 0.0, 0.0, mode=gpu.MODE_GN_P2P)
icp.setTargetPointCloud(xyz, nrm)
icp.setSourcePointCloud(scan)
icp.setInitialTransformation(T0)
r = icp.calculateAlignment()
print("done", r.iterations)
