"""Row N2: point_cloud_processing.hpp on the device vs the oracle and vs plain numpy."""
import numpy as np
import pytest

from oracle import oracle


def test_oracle_preprocess_vs_numpy(small_world):
    xyz = small_world["scan"][:, :3]
    assert np.array_equal(oracle.subsample(xyz, 2), xyz[::2])
    assert np.array_equal(oracle.subsample(xyz[:3], 5), xyz[:3])           # shorter than the step: unchanged
    assert np.array_equal(oracle.remove_floor(xyz), xyz[xyz[:, 2] > 0])
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = [1.0, -2.0, 0.5]
    pts, idx = oracle.crop_radius(T, 10.0, xyz)
    d = (T[:3, 3] - xyz).astype(np.float32)
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + d[:, 2] * d[:, 2]
    keep = np.nonzero(d2 < np.float32(100.0))[0]
    order = keep[np.lexsort((keep, d2[keep]))]                             # by distance, ties by index
    assert np.array_equal(idx, order.astype(np.int32)) and np.array_equal(pts, xyz[order])


@pytest.mark.gpu
def test_preprocess_gpu_matches_oracle(small_world):
    import ssf_gpu
    rng = np.random.default_rng(3)
    clouds = [small_world["scan"], small_world["map"][:30000], np.zeros((0, 3), np.float32),
              np.repeat(rng.uniform(-5, 5, (500, 3)).astype(np.float32), 3, axis=0)]  # duplicated points: ties
    T = np.eye(4)
    T[:3, 3] = [0.3, -0.7, 0.2]
    for c in clouds:
        for step in (1, 2, 3, 15, 10**6):
            assert np.array_equal(ssf_gpu.applyUniformSubsample(c, step), oracle.subsample(c, step) if len(c) else c[:, :3])
        assert np.array_equal(ssf_gpu.removeFloor(c), oracle.remove_floor(c) if len(c) else c[:, :3])
        for radius in (0.0, 2.5, 10.0, 1e3):
            g, gi = ssf_gpu.cropPointCloudThroughRadius(T, radius, c, return_indices=True)
            if len(c):
                o, oi = oracle.crop_radius(T, radius, c)
                assert np.array_equal(gi, oi) and np.array_equal(g.view(np.uint32), o.view(np.uint32))
            else:
                assert g.shape[0] == 0


@pytest.mark.gpu
def test_node_preprocessing_chain(c1_world):
    """The chain of localization_node.cpp:292-303: stride-2 scan, 10 m crops of scan and map, then ICP."""
    import ssf_gpu
    w = c1_world
    scan = ssf_gpu.applyUniformSubsample(w["scan"], 2)
    scan = ssf_gpu.cropPointCloudThroughRadius(np.eye(4), 10.0, scan)
    ref_map = ssf_gpu.cropPointCloudThroughRadius(w["T0"], 10.0, w["map"])
    o_scan = oracle.crop_radius(np.eye(4), 10.0, oracle.subsample(w["scan"], 2))[0]
    o_map = oracle.crop_radius(w["T0"], 10.0, w["map"])[0]
    assert np.array_equal(scan, o_scan) and np.array_equal(ref_map, o_map)
    icp = ssf_gpu.ICPPointToPoint(0.5, 10, 0.05, 1e-5)
    icp.setTargetPointCloud(ref_map)
    icp.setSourcePointCloud(scan)
    icp.setInitialTransformation(w["T0"])
    g = icp.calculateAlignment()
    o, ocorr, _ = oracle.icp_reference(oracle.KdTree(o_map), o_scan, w["T0"])
    assert np.array_equal(g.transformation.view(np.uint32), o.T.view(np.uint32)) and g.iterations == o.iterations
    assert np.array_equal(icp.correspondences(), ocorr)
