"""The three C++ drop-in headers (slam-sensor-fusion_b200/cpp/localization/) driven like localization_node.cpp
drives the reference's: applyUniformSubsample / cropPointCloudThroughRadius / removeFloor, BruteForceAlignment,
ICPPointToPoint, plus ssf::ResidentMap (the re-crop as a window change of the map kept in HBM).  They compile
without PCL / Eigen / ROS against the C ABI; on a GPU every step must equal the oracle (and oracle/_ref, the
reference's own sources, when it was shipped)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT

DRIVER_SRC = os.path.join(ROOT, "tests", "cpp", "node_driver.cpp")


def _build(tmp_path):
    exe = str(tmp_path / "node_driver")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "cpp"),
           DRIVER_SRC, "-o", exe, "-L", os.path.join(PKG, "csrc"), "-lssf_gpu", "-Wl,-rpath," + os.path.join(PKG, "csrc")]
    subprocess.check_call(cmd)
    return exe


def test_node_flow_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    assert subprocess.call([exe, "--help"], stderr=subprocess.DEVNULL) == 0


def _checksum(c):
    w = (np.arange(c.shape[0]) % 97 + 1).astype(np.float64)
    return np.float32((w * (c[:, 0].astype(np.float64) + 2.0 * c[:, 1] + 3.0 * c[:, 2])).sum())


@pytest.mark.gpu
def test_node_flow_matches_oracle(tmp_path, small_world):
    from oracle import oracle, ref
    w = small_world
    exe = _build(tmp_path)
    prior = w["T0"].astype(np.float32)
    # the scan reaches the node in the sensor frame; crop centre of the map = the prior's translation
    w["map"].astype(np.float32).tofile(tmp_path / "map.f32")
    w["scan"].astype(np.float32).tofile(tmp_path / "scan.f32")
    np.ascontiguousarray(prior.T).tofile(tmp_path / "prior.f32")
    subprocess.check_call([exe, str(tmp_path / "map.f32"), str(tmp_path / "scan.f32"), str(tmp_path / "prior.f32"),
                           str(tmp_path / "out.f32")])
    out = np.fromfile(tmp_path / "out.f32", np.float32)
    it = iter(range(out.size))

    def take(n=1):
        v = out[[next(it) for _ in range(n)]]
        return v if n > 1 else v[0]

    # ---- the oracle's version of the same flow ------------------------------------------------------
    m3 = oracle.subsample(w["map"], 3)
    assert take() == m3.shape[0] and take() == m3.shape[0]
    scan2 = oracle.subsample(w["scan"], 2)
    cropped, _ = oracle.crop_radius(np.eye(4), 10.0, scan2)
    assert take() == cropped.shape[0]
    assert take() == pytest.approx(_checksum(cropped), rel=1e-5)
    ref_map, _ = oracle.crop_radius(prior, 10.0, m3)
    assert take() == ref_map.shape[0]
    assert take() == pytest.approx(_checksum(ref_map), rel=1e-5)
    scan_c, map_c = oracle.remove_floor(cropped), oracle.remove_floor(oracle.subsample(ref_map, 15))
    assert take() == scan_c.shape[0] and take() == map_c.shape[0]
    prm = oracle.BfaParams(0.1, 0.1, 0.05, 0.5, 0.5, 0.1, float(np.float32(np.pi) / np.float32(18.0)),
                           float(np.float32(np.pi) / np.float32(6.0)), 0.1)
    ok, T_bfa, _, _ = oracle.bfa_align(oracle.KdTree(map_c), scan_c, prior, prm, threads=8)
    assert bool(take()) == ok and bool(take()) == ok
    assert np.array_equal(take(16).reshape(4, 4).T, T_bfa)
    fine, _, _ = oracle.icp_reference(oracle.KdTree(ref_map), cropped, prior, 0.5, 10, 0.05, 1e-5)
    for _ in range(2):   # host path, then the resident-map re-crop
        if _ == 1:
            assert take() == ref_map.shape[0]
        row = take(19)
        assert np.array_equal(row[:16].reshape(4, 4).T.view(np.uint32), fine.T.view(np.uint32))
        assert np.float32(row[16]).view(np.uint32) == np.float32(fine.error).view(np.uint32)
        assert int(row[17]) == fine.iterations and bool(row[18]) == bool(fine.has_converged)
    assert take() == ref_map.shape[0]
    assert take() == pytest.approx(_checksum(ref_map), rel=1e-5)
    # ---- and the reference's own sources, when shipped --------------------------------------------------
    if ref.available():
        r = ref.ICPPointToPoint(0.5, 10, 0.05, 1e-5)
        r.setDebugMode(False)
        r.setTargetPointCloud(ref.crop_radius(prior, 10.0, ref.subsample(w["map"], 3)))
        r.setSourcePointCloud(ref.crop_radius(np.eye(4), 10.0, ref.subsample(w["scan"], 2)))
        r.setInitialTransformation(prior)
        rr = r.calculateAlignment()
        assert np.array_equal(rr.T.view(np.uint32), fine.T.view(np.uint32)) and rr.iterations == fine.iterations
