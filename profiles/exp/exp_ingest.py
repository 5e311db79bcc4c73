"""Row N3 at map scale: M = 50M points in 50 binary PCD tiles -> pinned host -> HBM -> voxel grid 0.1 on the
device (ssf_map_from_pcd_folder).  Prints the device time of the ingest and the rate against 2 x 16 x M bytes.
   python profiles/exp/exp_ingest.py [points, default 50_000_000]"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))
import numpy as np
import ssf_gpu
from ssf_gpu import synth, pcd

M = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
# leaf 0.1 over the 966 m extent of this map overflows PCL's int32 voxel index (9660 x 9660 x 300): pcl::VoxelGrid
# refuses and returns the input -- and so does the device path.  0.3 m is the finest leaf that runs.
LEAF = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
xyz, _, _ = synth.make_map(M)
with tempfile.TemporaryDirectory(dir=os.environ.get("SSF_BENCH_TMP", "/tmp")) as d:
    t0 = time.time()
    for k, t in enumerate(np.array_split(xyz[:, :3], 50)):
        pcd.write_pcd_binary(os.path.join(d, f"cloud_{10 * (k + 1)}.pcd"), t)
    print(f"wrote 50 tiles ({M * 12 / 1e6:.0f} MB) in {time.time() - t0:.1f}s", flush=True)
    for rep in range(2):
        t0 = time.time()
        rm = ssf_gpu.ResidentMap.from_pcd_folder(d, "map", LEAF, save=False)
        wall = time.time() - t0
        gbs = 2 * 16 * M / (rm.ingest_ms * 1e-3) / 1e9
        print(f"ingest (leaf {LEAF}): {M} points -> {len(rm)} voxels; stream {rm.ingest_ms:.1f} ms ({gbs:.1f} GB/s against 2 x 16 x M), of which "
              f"the voxel filter {rm.merge_ms:.1f} ms ({2 * 16 * M / (rm.merge_ms * 1e-3) / 1e9:.0f} GB/s); "
              f"wall {wall:.2f}s incl. file reads ({M * 12 / wall / 1e9:.2f} GB/s of PCD)", flush=True)
        del rm
