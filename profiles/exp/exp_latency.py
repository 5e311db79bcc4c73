"""Single-scan latency of ssf_icp_align per mode (device ms by CUDA events, median of 7)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np, time
import ssf_gpu
from ssf_gpu import synth
xyz, nrm, half = synth.make_map(1_000_000, normals=True)
T = synth.street_pose(100, half=half); sc = synth.make_scan(T, 32, 1024, scan_id=100); T0 = synth.perturb_pose(T, 100)
for name, mode, red, acc, eps in (("REFERENCE strict", ssf_gpu.MODE_REFERENCE, ssf_gpu.REDUCE_STRICT, 0.05, 1e-5),
                                  ("REFERENCE fast", ssf_gpu.MODE_REFERENCE, ssf_gpu.REDUCE_FAST, 0.05, 1e-5),
                                  ("GN p2plane", ssf_gpu.MODE_GN_P2PLANE, 0, 0.0, 0.0), ("GN p2p", ssf_gpu.MODE_GN_P2P, 0, 0.0, 0.0),
                                  ("O3D flow (30)", ssf_gpu.MODE_O3D_P2P, 0, 0.0, 0.0)):
    icp = ssf_gpu.ICPPointToPoint(0.5, 30 if "O3D" in name else 10, acc, eps, mode=mode, reduce=red)
    icp.setTargetPointCloud(xyz, nrm); icp.setSourcePointCloud(sc); icp.setInitialTransformation(T0)
    ms, wall = [], []
    for _ in range(7):
        t0 = time.perf_counter(); r = icp.calculateAlignment(); wall.append((time.perf_counter() - t0) * 1e3); ms.append(r.device_ms)
    print(f"{name:18s} n={sc.shape[0]} it={r.iterations} searches={r.n_searches} device {np.median(ms):7.3f} ms  wall {np.median(wall):7.3f} ms")
