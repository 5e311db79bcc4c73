"""ctypes front end of ``oracle/_ref/libssf_ref.so``: the reference's OWN sources
(``localization/src/icp_point_to_point.cpp``, ``brute_force_alignment.cpp``,
``point_cloud_processing.hpp``) compiled unmodified against the stand-in Eigen/PCL headers of
``oracle/ref_stubs/`` (recipe: ``oracle/Makefile`` target ``ref``; wrappers: ``ref_driver.cpp``).

TEST INFRASTRUCTURE ONLY -- used by ``tests/`` to pin ``ssf_oracle.c`` to the reference's text
and by ``bench.py`` as the ``"reference"`` CPU baseline; never imported by the product.  The
library is built in the container that has ``/root/reference``; on the GPU box only the prebuilt
``.so`` (shipped with the snapshot) is used.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libssf_ref.so")
_REF_SRC = "/root/reference/localization/src/icp_point_to_point.cpp"
_lib = None


class Result(ctypes.Structure):
    _fields_ = [("transformation", ctypes.c_float * 16), ("error", ctypes.c_float), ("iterations", ctypes.c_int32),
                ("has_converged", ctypes.c_int32)]

    @property
    def T(self) -> np.ndarray:
        return np.array(self.transformation, dtype=np.float32).reshape(4, 4).T.copy()


class BfaParams(ctypes.Structure):
    _fields_ = [("x_step", ctypes.c_float), ("y_step", ctypes.c_float), ("z_step", ctypes.c_float),
                ("x_range", ctypes.c_float), ("y_range", ctypes.c_float), ("z_range", ctypes.c_float),
                ("yaw_step", ctypes.c_float), ("yaw_range", ctypes.c_float), ("mean_error_threshold", ctypes.c_float)]


def available() -> bool:
    """True when the library exists or can be built here (the reference tree is present)."""
    return os.path.exists(_LIB_PATH) or os.path.exists(_REF_SRC)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if os.path.exists(_REF_SRC):
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"], stdout=subprocess.DEVNULL)
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError("oracle/_ref/libssf_ref.so is missing and /root/reference is not present to build it")
        L = ctypes.CDLL(_LIB_PATH)
        vp, i64, i32, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
        L.ssf_ref_icp_create.restype = vp
        L.ssf_ref_icp_create.argtypes = [f32, i32, f32, f32]
        L.ssf_ref_icp_destroy.argtypes = [vp]
        L.ssf_ref_icp_set_params.argtypes = [vp, f32, i32, f32, f32]
        L.ssf_ref_icp_set_debug.argtypes = [vp, i32]
        L.ssf_ref_icp_set_target.argtypes = [vp, vp, i64, i32]
        L.ssf_ref_icp_set_source.argtypes = [vp, vp, i64, i32]
        L.ssf_ref_icp_set_initial.argtypes = [vp, vp]
        L.ssf_ref_icp_align.argtypes = [vp, ctypes.POINTER(Result)]
        L.ssf_ref_last_stdout.restype = ctypes.c_char_p
        L.ssf_ref_last_stderr.restype = ctypes.c_char_p
        L.ssf_ref_default_result.argtypes = [ctypes.POINTER(Result)]
        L.ssf_ref_bfa_align.argtypes = [vp, i64, i32, vp, i64, i32, vp, ctypes.POINTER(BfaParams), i32, vp, vp]
        L.ssf_ref_crop_radius.restype = i64
        L.ssf_ref_crop_radius.argtypes = [vp, i64, i32, vp, ctypes.c_double, vp]
        L.ssf_ref_subsample.restype = i64
        L.ssf_ref_subsample.argtypes = [vp, i64, i32, i64, vp]
        L.ssf_ref_remove_floor.restype = i64
        L.ssf_ref_remove_floor.argtypes = [vp, i64, i32, vp]
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] not in (3, 4):
        raise ValueError(f"expected (n, 3|4) array, got {a.shape}")
    return a


def _colmajor(T) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).reshape(16)


class ICPPointToPoint:
    """The reference's class, same method names (icp_point_to_point.h:41-85)."""

    def __init__(self, max_correspondence_dist, num_iterations, acceptable_mean_error, transformation_epsilon):
        self._h = lib().ssf_ref_icp_create(max_correspondence_dist, num_iterations, acceptable_mean_error,
                                           transformation_epsilon)
        self._prm = [max_correspondence_dist, num_iterations, acceptable_mean_error, transformation_epsilon]
        self.stdout = ""
        self.stderr = ""

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ssf_ref_icp_destroy(self._h)
            self._h = None

    def _push(self):
        lib().ssf_ref_icp_set_params(self._h, *self._prm)

    def setMaxCorrespondenceDist(self, v):
        self._prm[0] = v
        self._push()

    def setNumIterations(self, v):
        self._prm[1] = v
        self._push()

    def setAcceptableMeanError(self, v):
        self._prm[2] = v
        self._push()

    def setTransformationEpsilon(self, v):
        self._prm[3] = v
        self._push()

    def setDebugMode(self, on):
        lib().ssf_ref_icp_set_debug(self._h, 1 if on else 0)

    def setTargetPointCloud(self, xyz):
        a = _f32(xyz)
        lib().ssf_ref_icp_set_target(self._h, a.ctypes.data, a.shape[0], a.shape[1])

    def setSourcePointCloud(self, xyz):
        a = _f32(xyz)
        lib().ssf_ref_icp_set_source(self._h, a.ctypes.data, a.shape[0], a.shape[1])

    def setInitialTransformation(self, T):
        Tc = _colmajor(T)
        lib().ssf_ref_icp_set_initial(self._h, Tc.ctypes.data)

    def calculateAlignment(self) -> Result:
        r = Result()
        lib().ssf_ref_icp_align(self._h, ctypes.byref(r))
        self.stdout = lib().ssf_ref_last_stdout().decode()
        self.stderr = lib().ssf_ref_last_stderr().decode()
        return r


def default_result() -> Result:
    r = Result()
    lib().ssf_ref_default_result(ctypes.byref(r))
    return r


def bfa_align(target, src, T_guess, prm: BfaParams, n_calls: int = 1):
    """BruteForceAlignment: n_calls consecutive alignClouds(); returns (success[n_calls], T[n_calls, 4, 4])."""
    t, s = _f32(target), _f32(src)
    ok = np.zeros(n_calls, np.int32)
    Ts = np.zeros((n_calls, 16), np.float32)
    Tc = _colmajor(T_guess)
    lib().ssf_ref_bfa_align(t.ctypes.data, t.shape[0], t.shape[1], s.ctypes.data, s.shape[0], s.shape[1], Tc.ctypes.data,
                            ctypes.byref(prm), n_calls, ok.ctypes.data, Ts.ctypes.data)
    return ok.astype(bool), Ts.reshape(n_calls, 4, 4).transpose(0, 2, 1).copy()


def crop_radius(T, radius, xyz) -> np.ndarray:
    a = _f32(xyz)
    out = np.empty((max(1, a.shape[0]), 4), np.float32)
    Tc = _colmajor(T)
    n = lib().ssf_ref_crop_radius(a.ctypes.data, a.shape[0], a.shape[1], Tc.ctypes.data, float(radius), out.ctypes.data)
    return out[:n, :3].copy()


def subsample(xyz, step) -> np.ndarray:
    a = _f32(xyz)
    out = np.empty((max(1, a.shape[0]), 4), np.float32)
    n = lib().ssf_ref_subsample(a.ctypes.data, a.shape[0], a.shape[1], step, out.ctypes.data)
    return out[:n, :3].copy()


def remove_floor(xyz) -> np.ndarray:
    a = _f32(xyz)
    out = np.empty((max(1, a.shape[0]), 4), np.float32)
    n = lib().ssf_ref_remove_floor(a.ctypes.data, a.shape[0], a.shape[1], out.ctypes.data)
    return out[:n, :3].copy()
