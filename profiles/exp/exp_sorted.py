"""Experiment: how much does query ordering / placement matter for the search kernel?"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np
import ssf_gpu
from ssf_gpu import synth

xyz, nrm, half = synth.make_map(5_000_000, normals=True)
ctx = ssf_gpu.Context(0)
icp = ssf_gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, mode=ssf_gpu.MODE_GN_P2PLANE, context=ctx)
icp.setTargetPointCloud(xyz, nrm)
h = np.float32(np.sqrt(np.float32(0.5)) * np.float32(1.01))
o = xyz[:, :3].min(0)

def cellkey(pw):
    c = np.floor((pw - o) / h).astype(np.int64)
    return (c[:, 2] * 4096 + c[:, 1]) * 4096 + c[:, 0]

def run(name, k0, sort):
    scans, inits = [], []
    for d in range(16):
        k = k0 + 40 * d
        T = synth.street_pose(k, half=half)
        sc = synth.make_scan(T, 64, 2048, scan_id=k)
        v = ssf_gpu.voxel_down_sample(sc, 0.2, ctx)
        for r in range(4):
            T0 = synth.perturb_pose(T, 4 * d + r)
            if sort:
                pw = v @ T0[:3, :3].T + T0[:3, 3]
                v2 = v[np.argsort(cellkey(pw), kind="stable")]
            else:
                v2 = v
            scans.append(v2); inits.append(T0)
    tot = sum(s.shape[0] for s in scans)
    b = ssf_gpu.Batch(icp, len(scans), tot + 1)
    b.upload(scans); b.set_initial(inits)
    for _ in range(3): b.run()
    ctx.synchronize(); ctx.time_searches(True)
    t0 = time.perf_counter()
    for _ in range(5): b.run()
    ctx.synchronize(); dt = (time.perf_counter() - t0) / 5
    ms, n = ctx.search_time(); ctx.time_searches(False)
    res = b.results()
    print(f"{name:28s} queries/launch {tot:8d} step {dt*1e3:7.3f} ms  search avg {ms/n*1e3:7.1f} us  "
          f"Gq/s {tot/(ms/n*1e-3)/1e9:6.2f}  matched {np.mean([r.fitness for r in res]):.3f}")
    b.close()

a = max(2.0, half - 15.0)
kc = int(a / 0.15)  # pose index at x = 0 (map centre)
run("edge poses, scan order", 0, False)
run("edge poses, cell-sorted", 0, True)
run("centre poses, scan order", kc - 320, False)
run("centre poses, cell-sorted", kc - 320, True)
