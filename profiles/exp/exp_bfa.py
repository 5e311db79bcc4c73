"""Timing of the pose-grid alignment (row N1) at the node's settings: 7 776 poses, map cropped to 10 m and
subsampled x15 with the floor removed, scan with the floor removed (localization_node.cpp:38-43, 207-219)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np
import ssf_gpu
from ssf_gpu import synth
from oracle import oracle
xyz, _, half = synth.make_map(1_000_000)
T = synth.street_pose(int(half / 0.15), half=half)
scan = synth.make_scan(T, 32, 1024, scan_id=5, max_range=100.0)
scan = oracle.subsample(scan, 2)                                           # localization_node.cpp:292
d = np.linalg.norm(xyz[:, :3] - T[:3, 3], axis=1)
m = oracle.remove_floor(oracle.subsample(xyz[d < 10.0], 15))               # :302, :211-212
sc = scan[np.linalg.norm(scan[:, :3], axis=1) < 10.0]                       # :296
s = oracle.remove_floor((sc[:, :3] @ T[:3, :3].T + T[:3, 3]).astype(np.float32))  # map-frame z for the floor test
s = (np.c_[s[:, :3] - T[:3, 3]] @ T[:3, :3]).astype(np.float32)
prm = oracle.BfaParams.node_defaults()
T_prev = synth.perturb_pose(T, 9, xy=0.4, yaw_deg=8.0).astype(np.float32)
bfa = ssf_gpu.BruteForceAlignment()
bfa.setMeanErrorThreshold(prm.mean_error_threshold); bfa.setXYZStep(prm.x_step, prm.y_step, prm.z_step)
bfa.setXYZRange(prm.x_range, prm.y_range, prm.z_range); bfa.setRotationStep(prm.yaw_step); bfa.setRotationRange(prm.yaw_range)
bfa.setInitialGuess(T_prev); bfa.setSourceCloud(s); bfa.setTargetCloud(m)
bfa.alignClouds()
t0 = time.time(); ok = bfa.alignClouds(); dt = time.time() - t0
n_pose = bfa.last_scores.shape[0]
print(f"GPU: map {m.shape[0]} pts, scan {s.shape[0]} pts, {n_pose} poses = {n_pose * s.shape[0] / 1e6:.1f} M unbounded NN queries: "
      f"{dt * 1e3:.1f} ms ({n_pose * s.shape[0] / dt / 1e9:.2f} G queries/s), success {ok}, best score {bfa.best_score:.4f}")
tree = oracle.KdTree(m)
sub = oracle.BfaParams.node_defaults(); sub.x_range = sub.y_range = 0.2
t0 = time.time(); oracle.bfa_align(tree, s, T_prev, sub, no_early_exit=True, threads=oracle.max_threads()); dt2 = time.time() - t0
n2 = oracle.bfa_poses(T_prev, sub).shape[0]
print(f"CPU oracle ({oracle.max_threads()} threads): {n2} poses in {dt2:.2f} s -> {dt2 / n2 * n_pose:.1f} s for the full grid")
