"""Check + timing: one batch of many small scans (config-4 shape: independent scans against one map)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np
import ssf_gpu
from ssf_gpu import synth
N = int(os.environ.get("N", "768"))
xyz, nrm, half = synth.make_map(1_000_000, normals=True)
scans, inits = [], []
base = []
for d in range(24):
    T = synth.street_pose(3 + 17 * d, half=half)
    base.append((T, synth.make_scan(T, 16, 256 + 16 * d, scan_id=d, max_range=60.0)))
for s in range(N):
    T, sc = base[s % len(base)]
    scans.append(sc); inits.append(synth.perturb_pose(T, 1000 + s))
for mode, name in ((ssf_gpu.MODE_GN_P2PLANE, "gn_p2plane"), (ssf_gpu.MODE_REFERENCE, "reference")):
    icp = ssf_gpu.ICPPointToPoint(0.5, 10, 0.05 if mode == ssf_gpu.MODE_REFERENCE else 0.0,
                                  1e-5 if mode == ssf_gpu.MODE_REFERENCE else 0.0, mode=mode)
    icp.setTargetPointCloud(xyz, nrm)
    t0 = time.time(); res = icp.align_batch(scans, inits); dt = time.time() - t0
    t0 = time.time(); res = icp.align_batch(scans, inits); dt = time.time() - t0
    bad = 0
    for k in range(0, N, 97):
        icp.setSourcePointCloud(scans[k]); icp.setInitialTransformation(inits[k])
        r1 = icp.calculateAlignment()
        if not (np.array_equal(r1.transformation.view(np.uint32), res[k].transformation.view(np.uint32)) and r1.k_final == res[k].k_final):
            bad += 1
    print(f"{name}: {N} scans in one batch: {dt*1e3:.1f} ms ({N/dt:.0f} scans/s incl. upload), device {res[0].device_ms:.2f} ms, "
          f"sampled batch-vs-single mismatches: {bad}")
