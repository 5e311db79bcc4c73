"""Row N3: PCD tiles in, merged + voxel-filtered map out."""
import os

import numpy as np
import pytest

from ssf_gpu import pcd


def test_pcd_binary_roundtrip_and_ascii(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.normal(size=(1000, 3)).astype(np.float32)
    p = str(tmp_path / "cloud_10.pcd")
    pcd.write_pcd_binary(p, a)
    assert os.path.getsize(p) == len(open(p, "rb").read().split(b"DATA binary\n")[0]) + len(b"DATA binary\n") + 12000
    assert np.array_equal(pcd.read_pcd(p), a)
    # a tile with extra fields (e.g. intensity), ascii
    q = tmp_path / "ascii.pcd"
    q.write_text("# .PCD v0.7\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
                 "WIDTH 2\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 2\nDATA ascii\n1 2 3 9\n4 5 6 9\n")
    assert np.array_equal(pcd.read_pcd(str(q)), np.array([[1, 2, 3], [4, 5, 6]], np.float32))
    # binary with an extra field keeps x y z
    dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("i", "<f4")])
    raw = np.zeros(3, dt)
    raw["x"], raw["y"], raw["z"] = [1, 2, 3], [4, 5, 6], [7, 8, 9]
    r = tmp_path / "xyzi.pcd"
    r.write_bytes(b"VERSION 0.7\nFIELDS x y z i\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 3\nHEIGHT 1\n"
                  b"POINTS 3\nDATA binary\n" + raw.tobytes())
    assert np.array_equal(pcd.read_pcd(str(r)), np.array([[1, 4, 7], [2, 5, 8], [3, 6, 9]], np.float32))


@pytest.mark.gpu
def test_merge_scans_and_save_matches_oracle(tmp_path, small_world):
    from oracle import oracle
    m = small_world["map"][:, :3]
    parts = np.array_split(m[:40000], 4)
    for k, part in enumerate(parts):
        pcd.write_pcd_binary(str(tmp_path / f"cloud_{10 * (k + 1)}.pcd"), part)
    out = pcd.get_map_cloud(str(tmp_path), "map", 0.1)
    order = [e.name for e in os.scandir(tmp_path) if e.name.startswith("cloud_")]
    merged = np.concatenate([pcd.read_pcd(str(tmp_path / n)) for n in order])
    ref, refused = oracle.voxel_grid(merged, 0.1)
    assert not refused and np.array_equal(out.view(np.uint32), ref[:, :3].copy().view(np.uint32))
    # second call takes the cached map.pcd branch (no filter): identical cloud back
    assert os.path.exists(tmp_path / "map.pcd")
    assert np.array_equal(pcd.get_map_cloud(str(tmp_path), "map", 0.1), out)


def test_native_reader_handles_float64_and_truncation(tmp_path):
    dt = np.dtype([("i", "<u2"), ("x", "<f8"), ("y", "<f8"), ("z", "<f8")])
    raw = np.zeros(4, dt)
    raw["x"], raw["y"], raw["z"] = [1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]
    p = tmp_path / "d.pcd"
    head = b"VERSION 0.7\nFIELDS i x y z\nSIZE 2 8 8 8\nTYPE U F F F\nCOUNT 1 1 1 1\nWIDTH 4\nHEIGHT 1\nPOINTS 4\nDATA binary\n"
    p.write_bytes(head + raw.tobytes())
    assert np.array_equal(pcd.read_pcd(str(p)), np.array([[1, 5, 9], [2, 6, 10], [3, 7, 11], [4, 8, 12]], np.float32))
    from ssf_gpu import SsfError
    p.write_bytes(head + raw.tobytes()[:-5])
    with pytest.raises(SsfError):
        pcd.read_pcd(str(p))
    (tmp_path / "c.pcd").write_bytes(head.replace(b"DATA binary", b"DATA binary_compressed"))
    with pytest.raises(SsfError):
        pcd.read_pcd(str(tmp_path / "c.pcd"))
    with pytest.raises(SsfError):
        pcd.read_pcd(str(tmp_path / "missing.pcd"))


@pytest.mark.gpu
def test_resident_map_crop_is_the_reference_crop(small_world):
    """ssf_map_crop_radius / ssf_map_crop_to_target: same points, same order as cropPointCloudThroughRadius
    (point_cloud_processing.hpp:31-53), and the target built from the crop in HBM gives the oracle's ICP."""
    import ssf_gpu
    from oracle import oracle
    w = small_world
    rm = ssf_gpu.ResidentMap(w["map"])
    rm.subsample(3)                                             # localization_node.cpp:20
    m3 = oracle.subsample(w["map"], 3)
    assert len(rm) == m3.shape[0] and np.array_equal(rm.download(), m3)
    for centre_T, radius in ((w["T0"], 10.0), (w["T_gt"], 3.0), (np.eye(4), 0.0)):
        got, gi = rm.crop(centre_T, radius, return_indices=True)
        want, wi = oracle.crop_radius(centre_T, radius, m3)
        assert np.array_equal(got, want) and np.array_equal(gi, wi)
    icp = ssf_gpu.ICPPointToPoint(0.5, 10, 0.05, 1e-5)
    n = rm.crop_to_target(icp, w["T0"], 10.0)                   # localization_node.cpp:300-305 in HBM
    target, _ = oracle.crop_radius(w["T0"], 10.0, m3)
    assert n == target.shape[0]
    scan = oracle.crop_radius(np.eye(4), 10.0, oracle.subsample(w["scan"], 2))[0]
    icp.setSourcePointCloud(scan)
    icp.setInitialTransformation(w["T0"])
    g = icp.calculateAlignment()
    o, ocorr, _ = oracle.icp_reference(oracle.KdTree(target), scan, w["T0"])
    assert np.array_equal(g.transformation.view(np.uint32), o.T.view(np.uint32)) and g.iterations == o.iterations
    assert np.array_equal(icp.correspondences(), ocorr)


@pytest.mark.gpu
def test_resident_map_from_pcd_folder_at_scale(tmp_path):
    """Row N3 at map scale: tiles -> pinned host -> HBM -> pcl::VoxelGrid(0.1) on the device, bit-exact against the
    oracle, with the measured ingest rate."""
    import ssf_gpu
    from oracle import oracle
    from ssf_gpu import synth
    xyz, _, _ = synth.make_map(2_000_000)
    rng = np.random.default_rng(8)
    dense = (xyz[:, :3] + rng.normal(0, 0.03, (xyz.shape[0], 3))).astype(np.float32)   # overlapping tiles
    tiles = np.array_split(np.concatenate([xyz[:, :3], dense]), 16)
    for k, t in enumerate(tiles):
        pcd.write_pcd_binary(str(tmp_path / f"cloud_{10 * (k + 1)}.pcd"), t)
    rm = ssf_gpu.ResidentMap.from_pcd_folder(str(tmp_path), "map", 0.1, save=False)
    order = [e.name for e in os.scandir(tmp_path) if e.name.startswith("cloud_")]
    merged = np.concatenate([pcd.read_pcd(str(tmp_path / n)) for n in order])
    ref, refused = oracle.voxel_grid(merged, 0.1)
    assert not refused and len(rm) == ref.shape[0]
    assert np.array_equal(rm.download().view(np.uint32), ref[:, :3].copy().view(np.uint32))
    gbs = 2 * 16 * merged.shape[0] / (rm.ingest_ms * 1e-3) / 1e9
    print(f"ingest: {merged.shape[0]} points -> {len(rm)} voxels in {rm.ingest_ms:.2f} ms device time "
          f"({gbs:.1f} GB/s against 2 x 16 x M bytes)")
    assert rm.ingest_ms > 0


@pytest.mark.gpu
def test_pointcloud2_extraction():
    """pcl::fromROSMsg for x / y / z float32 at arbitrary offsets of a PointCloud2 record (e.g. x y z intensity
    ring time: 32-byte point_step), little- and big-endian."""
    import ssf_gpu
    rng = np.random.default_rng(2)
    n = 10_000
    dt = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f8"],
                   "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})
    rec = np.zeros(n, dt)
    xyz = rng.normal(size=(n, 3)).astype(np.float32)
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    rec["intensity"] = 7.0
    got = ssf_gpu.from_pointcloud2(rec.tobytes(), n, 32, (0, 4, 8))
    assert np.array_equal(got, xyz)
    be = np.zeros(n, np.dtype({"names": ["x", "y", "z"], "formats": [">f4", ">f4", ">f4"], "offsets": [8, 0, 4], "itemsize": 16}))
    be["x"], be["y"], be["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    got = ssf_gpu.from_pointcloud2(be.tobytes(), n, 16, (8, 0, 4), is_bigendian=True)
    assert np.array_equal(got, xyz)


@pytest.mark.gpu
def test_voxel_down_sample_open3d_semantics(small_world):
    """The Python node's pcd.voxel_down_sample(0.1) (localization_node.py:47): origin min_bound - v / 2, double
    centroids -- a different point set than pcl::VoxelGrid's."""
    import ssf_gpu
    from oracle import oracle
    rng = np.random.default_rng(3)
    clouds = [(small_world["map"][:60_000], 0.1), (small_world["scan"], 0.2),
              (rng.uniform(-2, 2, (30_000, 3)).astype(np.float32), 0.05),
              (np.array([[0, 0, 0], [np.nan, 1, 1], [0.04, 0.04, 0.04], [5, 5, 5]], np.float32), 0.1)]
    for c, v in clouds:
        g = ssf_gpu.voxel_down_sample(c, v, semantics="open3d")
        o = oracle.voxel_grid_o3d(c, v)
        assert g.shape == o.shape and np.array_equal(g.view(np.uint32), o.view(np.uint32))
    # independent numpy restatement (group-by on the double keys), compared as point sets
    c, v = clouds[0]
    p = c[:, :3].astype(np.float64)
    key = np.floor((p - (p.min(0) - v / 2)) / v).astype(np.int64)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    sums = np.zeros((cnt.size, 3))
    np.add.at(sums, inv.reshape(-1), p)
    want = (sums / cnt[:, None]).astype(np.float32)
    got = ssf_gpu.voxel_down_sample(c, v, semantics="open3d")
    assert got.shape == want.shape
    assert np.allclose(np.sort(got, axis=0), np.sort(want, axis=0), atol=1e-6)
    assert ssf_gpu.voxel_down_sample(c, v).shape != got.shape or not np.array_equal(ssf_gpu.voxel_down_sample(c, v), got)
