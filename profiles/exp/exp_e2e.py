"""Experiment: where does the end-to-end step go?  upload only / run only / sequential / pipelined (2, 3 batches)."""
import sys, os, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np, torch
import bench, ssf_gpu
from ssf_gpu import capi
B = 64
w, xyz, nrm, half, scans, inits, gts = bench.make_workload("c2", B, 0)
ctx = ssf_gpu.Context(0)
icp = ssf_gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, mode=ssf_gpu.MODE_GN_P2PLANE, context=ctx)
icp.setSourceVoxelLeaf(w["leaf"]); icp.setTargetPointCloud(xyz, nrm)
n_pts = [s.shape[0] for s in scans]; total = int(sum(n_pts))
pinned = torch.empty((total, 4), dtype=torch.float32, pin_memory=True); pinned.numpy()[:] = np.concatenate(scans, axis=0)
T_pinned = torch.empty((B, 16), dtype=torch.float32, pin_memory=True)
T_pinned.numpy()[:] = np.stack([np.asarray(T, np.float32).T.reshape(16) for T in inits])
res = (capi.IcpResult * B)()
bs = [ssf_gpu.Batch(icp, B, total + 1) for _ in range(3)]
def sync(): ctx.synchronize(); torch.cuda.synchronize()
def timed(name, fn, n=8):
    fn(2); sync(); t0 = time.perf_counter(); fn(n); sync(); dt = (time.perf_counter() - t0) / n
    print(f"{name:28s} {dt*1e3:7.3f} ms/step  {B/dt:9.0f} scans/s")
def upload_only(n):
    for i in range(n): bs[i % 2].upload_ptr(pinned.data_ptr(), n_pts, 16, wait=False)
    for b in bs[:2]: b.set_initial_ptr(T_pinned.data_ptr())
def run_only(n):
    for i in range(n): bs[0].run()
def sequential(n):
    for i in range(n):
        b = bs[0]; b.upload_ptr(pinned.data_ptr(), n_pts, 16, wait=False); b.set_initial_ptr(T_pinned.data_ptr()); b.run(); b.results_into(res)
def pipelined(k):
    def f(n):
        pend = []
        for i in range(n):
            b = bs[i % k]; b.upload_ptr(pinned.data_ptr(), n_pts, 16, wait=False); b.set_initial_ptr(T_pinned.data_ptr()); b.run()
            pend.append(b)
            if len(pend) >= k: pend.pop(0).results_into(res)
        for b in pend: b.results_into(res)
    return f
timed("upload only", upload_only)
bs[0].upload_ptr(pinned.data_ptr(), n_pts, 16); bs[0].set_initial_ptr(T_pinned.data_ptr())
timed("run only", run_only)
timed("sequential", sequential)
timed("pipelined x2", pipelined(2))
timed("pipelined x3", pipelined(3))
# ---- does the H2D copy slow down while the alignment kernels run (and vice versa)? ----
d = torch.empty((total, 4), dtype=torch.float32, device="cuda")
cs = torch.cuda.Stream()
def h2d(n):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        e0.record(cs)
        for _ in range(n): d.copy_(pinned, non_blocking=True)
        e1.record(cs)
    return e0, e1
sync(); e0, e1 = h2d(8); sync(); print("H2D alone            ", e0.elapsed_time(e1) / 8, "ms per 130 MB")
sync(); t0 = time.perf_counter(); run_only(8); e0, e1 = h2d(8); sync()
print("H2D during alignment ", e0.elapsed_time(e1) / 8, "ms per 130 MB; both done in", (time.perf_counter() - t0) / 8 * 1e3, "ms per step")
# ---- host time of every call of the 2-deep pipeline ----
sync()
log = []
pend = []
tp = time.perf_counter
for i in range(8):
    b = bs[i % 2]
    t0 = tp(); b.upload_ptr(pinned.data_ptr(), n_pts, 16, wait=False)
    t1 = tp(); b.set_initial_ptr(T_pinned.data_ptr())
    t2 = tp(); b.run()
    t3 = tp()
    pend.append(b)
    if len(pend) >= 2: pend.pop(0).results_into(res)
    t4 = tp()
    log.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
for b in pend: b.results_into(res)
sync()
for l in log: print("upload %.3f  set_initial %.3f  run %.3f  results %.3f ms" % tuple(x * 1e3 for x in l))
