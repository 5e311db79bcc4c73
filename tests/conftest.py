"""Shared test setup: import paths, in-tree builds, seeded synthetic fixtures."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "slam-sensor-fusion_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _ensure_built():
    """Build the host-side helpers (and the CUDA library when nvcc is present) if missing."""
    need = [os.path.join(PKG, "synth", "libssf_synth.so"), os.path.join(PKG, "csrc", "libssf_gpu.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.call(["make", "-C", PKG, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ROOT, "oracle", "libssf_oracle.so")):
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], stdout=subprocess.DEVNULL,
                        stderr=subprocess.DEVNULL)


_ensure_built()


def has_gpu() -> bool:
    try:
        from ssf_gpu import capi
        import ctypes
        h = ctypes.c_void_p()
        rc = capi.lib().ssf_ctx_create(0, ctypes.byref(h))
        if rc == 0:
            capi.lib().ssf_ctx_destroy(h)
            return True
    except Exception:
        pass
    return False


@pytest.fixture(scope="session")
def small_world():
    """C1-mini: 65 536-point map with normals, one 16x512 scan, ground truth and perturbed pose."""
    from ssf_gpu import synth
    xyz, nrm, half = synth.make_map(65536, normals=True)
    T_gt = synth.street_pose(3, half=half)
    scan = synth.make_scan(T_gt, beams=16, azimuths=512, scan_id=3, max_range=60.0)
    T0 = synth.perturb_pose(T_gt, 3)
    return dict(map=xyz, normals=nrm, half=half, T_gt=T_gt, scan=scan, T0=T0)


@pytest.fixture(scope="session")
def c1_world():
    """Config 1 at full size: 1M-point map, 32x1024 scan."""
    from ssf_gpu import synth
    xyz, nrm, half = synth.make_map(1_000_000, normals=True)
    T_gt = synth.street_pose(100, half=half)
    scan = synth.make_scan(T_gt, beams=32, azimuths=1024, scan_id=100)
    T0 = synth.perturb_pose(T_gt, 100)
    return dict(map=xyz, normals=nrm, half=half, T_gt=T_gt, scan=scan, T0=T0)


def pose_delta(Ta, Tb):
    """(translation distance [m], rotation angle [rad]) between two 4x4 poses."""
    Ta, Tb = np.asarray(Ta, np.float64), np.asarray(Tb, np.float64)
    dt = float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))
    R = Ta[:3, :3].T @ Tb[:3, :3]
    c = max(-1.0, min(1.0, (np.trace(R) - 1.0) / 2.0))
    s = np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2.0
    return dt, float(np.arctan2(s, c))
