"""The C++ drop-in class (slam-sensor-fusion_b200/cpp/localization/icp_point_to_point.h):
compiles against the C ABI without PCL/Eigen/ROS, and on a GPU reproduces the oracle's
ICPPointToPoint results when driven like localization_node.cpp drives the reference class."""
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT

DRIVER_SRC = os.path.join(ROOT, "tests", "cpp", "shim_driver.cpp")


def _build(tmp_path):
    exe = str(tmp_path / "shim_driver")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "cpp"),
           DRIVER_SRC, "-o", exe, "-L", os.path.join(PKG, "csrc"), "-lssf_gpu",
           "-Wl,-rpath," + os.path.join(PKG, "csrc")]
    subprocess.check_call(cmd)
    return exe


def test_shim_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    assert subprocess.call([exe, "--help"], stderr=subprocess.DEVNULL) == 0


@pytest.mark.gpu
def test_shim_matches_oracle(tmp_path, small_world):
    from oracle import oracle
    w = small_world
    exe = _build(tmp_path)
    w["map"].astype(np.float32).tofile(tmp_path / "map.f32")
    w["scan"].astype(np.float32).tofile(tmp_path / "scan.f32")
    np.ascontiguousarray(w["T0"].astype(np.float32).T).tofile(tmp_path / "T0.f32")
    subprocess.check_call([exe, str(tmp_path / "map.f32"), str(tmp_path / "scan.f32"), str(tmp_path / "T0.f32"),
                           str(tmp_path / "out.f32")])
    out = np.fromfile(tmp_path / "out.f32", np.float32).reshape(4, 19)
    tree = oracle.KdTree(w["map"])
    fine, _, _ = oracle.icp_reference(tree, w["scan"], w["T0"], 0.5, 10, 0.05, 1e-5)
    coarse, _, _ = oracle.icp_reference(tree, w["scan"], w["T0"], 5.0, 80, 0.4, 1e-2)
    for row, o in ((out[0], fine), (out[1], coarse), (out[2], fine)):
        assert np.array_equal(row[:16].reshape(4, 4).T.view(np.uint32), o.T.view(np.uint32))
        assert np.float32(row[16]).view(np.uint32) == np.float32(o.error).view(np.uint32)
        assert int(row[17]) == o.iterations and bool(row[18]) == bool(o.has_converged)
    # sentinel: initial transform, error 1e6, 0 iterations, not converged (icp_point_to_point.cpp:196-200)
    far = np.eye(4, dtype=np.float32)
    far[0, 3] = 1e4
    assert np.array_equal(out[3][:16].reshape(4, 4).T, far)
    assert out[3][16] == np.float32(1e6) and out[3][17] == 0 and out[3][18] == 0
