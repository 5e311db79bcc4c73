"""Debug aid: dump the correspondences of the small world after 1..N iterations (lib chosen by SSF_GPU_LIB),
or compare two dumps.   python exp_walkdiff.py dump OUT.npz | python exp_walkdiff.py cmp A.npz B.npz"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np

if sys.argv[1] == "dump":
    import ssf_gpu as gpu
    from ssf_gpu import synth
    xyz, nrm, half = synth.make_map(65536, normals=True)
    T_gt = synth.street_pose(3, half=half)
    scan = synth.make_scan(T_gt, beams=16, azimuths=512, scan_id=3, max_range=60.0)
    T0 = synth.perturb_pose(T_gt, 3)
    out = {"scan": scan, "map": xyz}
    for mode, mname in ((gpu.MODE_GN_P2P, "p2p"), (gpu.MODE_GN_P2PLANE, "p2plane")):
        for it in (1, 2, 3, 4, 6, 10):
            icp = gpu.ICPPointToPoint(0.5, it, 0.0, 0.0, mode=mode)
            icp.setTargetPointCloud(xyz, nrm)
            icp.setSourcePointCloud(scan)
            icp.setInitialTransformation(T0)
            r = icp.calculateAlignment()
            out[f"{mname}_{it}_corr"] = icp.correspondences().copy()
            out[f"{mname}_{it}_T"] = np.asarray(r.transformation)
    np.savez(sys.argv[2], **out)
else:
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    for k in a.files:
        if k.endswith("_corr"):
            ca, cb = a[k], b[k]
            bad = np.nonzero(ca != cb)[0]
            print(k, "n", len(ca), "mismatches", len(bad), "first", bad[:8], ca[bad[:8]], cb[bad[:8]],
                  "T equal", np.array_equal(a[k[:-4] + "T"], b[k[:-4] + "T"]))
