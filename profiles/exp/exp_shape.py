"""Which shape of the fused search kernel (128-thread blocks, four queries per thread / 512-thread blocks, one
query per thread) is faster at which batch size: scans of 32x1024 rays, point-to-plane GN x10 against the
5M-point map, SSF_SEARCH_WIDE forcing the shape.   python profiles/exp/exp_shape.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np
import ssf_gpu
from ssf_gpu import synth

xyz, nrm, half = synth.make_map(5_000_000, normals=True)
ctx = ssf_gpu.Context(0)
icp = ssf_gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, mode=ssf_gpu.MODE_GN_P2PLANE, context=ctx)
icp.setTargetPointCloud(xyz, nrm)
scans, inits = [], []
for d in range(64):
    T = synth.street_pose(37 * d + 5, half=half)
    scans.append(synth.make_scan(T, 32, 1024, scan_id=900 + d))
    inits.append(synth.perturb_pose(T, 900 + d))
for B in (1, 2, 4, 8, 16, 32, 64):
    row = []
    for wide in ("1", "0"):
        os.environ["SSF_SEARCH_WIDE"] = wide
        b = ssf_gpu.Batch(icp, B, sum(s.shape[0] for s in scans[:B]) + 1)
        b.upload(scans[:B]); b.set_initial(inits[:B])
        for _ in range(3):
            b.run()
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(30):
            b.run()
        ctx.synchronize()
        row.append((time.perf_counter() - t0) / 30 * 1e3)
        b.close()
    print(f"B={B:3d}  tiles {B * 62:5d}  wide {row[0]:.3f} ms  narrow {row[1]:.3f} ms  -> {'wide' if row[0] < row[1] else 'narrow'}", flush=True)
