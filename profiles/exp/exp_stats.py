"""Search statistics (debug build libssf_gpu_stats.so): eval4 calls / directory loads / rows per query."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np
from ssf_gpu import capi
capi.LIB_PATH = capi.LIB_PATH.replace("libssf_gpu.so", "libssf_gpu_stats.so")
import ssf_gpu
from ssf_gpu import synth
L = capi.lib()
xyz, nrm, half = synth.make_map(5_000_000, normals=True)
ctx = ssf_gpu.Context(0)
icp = ssf_gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, mode=ssf_gpu.MODE_GN_P2PLANE, context=ctx)
icp.setTargetPointCloud(xyz, nrm)
scans, inits = [], []
for d in range(8):
    T = synth.street_pose(40 * d, half=half)
    v = ssf_gpu.voxel_down_sample(synth.make_scan(T, 64, 2048, scan_id=40 * d), 0.2, ctx)
    scans.append(v); inits.append(synth.perturb_pose(T, d))
st = (ctypes.c_ulonglong * 8)()
for iters in (1, 2, 3, 4, 10):
    icp.setNumIterations(iters)
    b = ssf_gpu.Batch(icp, len(scans), sum(s.shape[0] for s in scans) + 1)
    b.upload(scans); b.set_initial(inits)
    L.ssf_debug_nn_stats(st, 1)
    b.run(); ctx.synchronize()
    L.ssf_debug_nn_stats(st, 1)
    q = st[3]
    print(f"iters={iters}: queries {q}  eval4/query {st[0]/q:.2f}  dir loads/query {st[1]/q:.2f}  rows/query {st[2]/q:.2f} "
          f"runs/query {st[6]/q:.2f}  past-ring-1 {st[4]/q:.4f}  matched {st[5]/q:.3f}  reach-mask exits {st[7]/q:.4f}")
    b.close()
