"""Per-launch device time of the search kernel on the bench workload (config 2 by default).
   SPS=256 REPS=3 python profiles/exp/exp_launch_times.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np
from ssf_gpu import capi
if os.environ.get('SSF_LIB'):
    capi.LIB_PATH = os.path.join(ROOT, os.environ['SSF_LIB'])
import bench, ssf_gpu

B = int(os.environ.get("SPS", "256")); reps = int(os.environ.get("REPS", "3"))
w, xyz, nrm, half, scans, inits, gts = bench.make_workload(os.environ.get("WORKLOAD", "c2"), B, 0)
ctx = ssf_gpu.Context(0)
mode = {"gn_p2plane": ssf_gpu.MODE_GN_P2PLANE, "gn_p2p": ssf_gpu.MODE_GN_P2P}[w["mode"]]
icp = ssf_gpu.ICPPointToPoint(bench.THR, w.get("iters", bench.ITERS), 0.0, 0.0, mode=mode, context=ctx)
icp.setSourceVoxelLeaf(w["leaf"])
icp.setTargetPointCloud(xyz, nrm)
batch = ssf_gpu.Batch(icp, B, sum(s.shape[0] for s in scans) + 1)
batch.upload(scans); batch.set_initial(inits)
for _ in range(2):
    batch.run()
ctx.synchronize()
ctx.time_searches(True)
for _ in range(reps):
    batch.run()
t = np.array(ctx.search_times()).reshape(reps, -1) * 1e3
ctx.time_searches(False)
import time
for _ in range(2):
    batch.run()
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    batch.run()
ctx.synchronize()
step_us = (time.perf_counter() - t0) / 10 * 1e6
res = batch.results()
err = [bench.pose_delta(r.transformation, T)[0] for r, T in zip(res, gts)]
print("per-launch us (median over reps):", " ".join(f"{x:.0f}" for x in np.median(t, axis=0)))
print(f"step {step_us:.0f} us (graph replay, wall clock over 10 steps); rest of step {step_us - np.median(t.sum(axis=1)):.0f} us")
print(f"search total {np.median(t.sum(axis=1)):.0f} us/step; median pose error {np.median(err):.4f} m; "
      f"checksum {sum(float(np.abs(np.asarray(r.transformation, np.float64)).sum()) for r in res):.9f} k_final {sum(r.k_final for r in res)}")
