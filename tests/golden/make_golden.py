"""Generates tests/golden/c1_mini.npz: frozen oracle outputs at fixed seeds.

Run from the repository root:  python tests/golden/make_golden.py
The reference itself cannot be run here (PCL / Eigen / ROS 2 absent), so these vectors freeze
the ORACLE (oracle/ssf_oracle.c), which tests/test_oracle.py pins against brute force,
cv2.flann, scipy and numpy.  GPU tests compare the CUDA path with the same file.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))

from oracle import oracle  # noqa: E402
from ssf_gpu import synth  # noqa: E402


def main():
    xyz, nrm, half = synth.make_map(16384, normals=True)
    T_gt = synth.street_pose(5, half=half)
    scan = synth.make_scan(T_gt, beams=8, azimuths=256, scan_id=5, max_range=40.0)
    T0 = synth.perturb_pose(T_gt, 5)
    tree = oracle.KdTree(xyz)
    T0f = T0.astype(np.float32)
    queries = (scan[:, :3] @ T0f[:3, :3].T + T0f[:3, 3]).astype(np.float32)
    nn_idx, nn_d2 = tree.nn(queries)
    ref, ref_corr, _ = oracle.icp_reference(tree, scan, T0)
    gnp, gnp_corr = oracle.icp_gn(tree, scan, T0, mode="p2plane", normals=nrm, num_iterations=10)
    gn2, _ = oracle.icp_gn(tree, scan, T0, mode="p2p", num_iterations=10)
    o3, o3_fit, _ = oracle.icp_o3d(tree, scan, T0, 0.5, 30)
    vox, _ = oracle.voxel_grid(scan, 0.2)
    out = os.path.join(ROOT, "tests", "golden", "c1_mini.npz")
    np.savez_compressed(
        out, map=xyz, normals=nrm, scan=scan, T0=T0, T_gt=T_gt, queries=queries, nn_idx=nn_idx, nn_d2=nn_d2,
        ref_T=ref.T, ref_error=np.float32(ref.error), ref_iterations=ref.iterations, ref_n_searches=ref.n_searches,
        ref_k_final=ref.k_final, ref_corr=ref_corr,
        gn_p2plane_T=gnp.T, gn_p2plane_iterations=gnp.iterations, gn_p2plane_error=np.float32(gnp.error),
        gn_p2plane_k=gnp.k_final, gn_p2p_T=gn2.T, gn_p2p_iterations=gn2.iterations,
        o3d_T=o3.T, o3d_iterations=o3.iterations, o3d_fitness=np.float32(o3_fit), o3d_rmse=np.float32(o3.error),
        vox_02=vox)
    print("wrote", out, os.path.getsize(out), "bytes;", "scan", scan.shape, "ref it", ref.iterations, "searches",
          ref.n_searches)


if __name__ == "__main__":
    main()
