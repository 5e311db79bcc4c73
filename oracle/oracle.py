"""ctypes front end of the CPU oracle (``ssf_oracle.c``).

TEST INFRASTRUCTURE ONLY -- see the header of ``ssf_oracle.c``.  Imported by ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py``; never by the product
package.  The reference has no behavioural tests; ``tests/test_ref_pin.py`` pins this restatement
bit for bit to the reference's own sources compiled unmodified (``oracle/_ref``, see ``ref.py``),
and ``tests/test_oracle.py`` pins the third-party pieces (FLANN k=1, Jacobi SVD, VoxelGrid)
against brute force, cv2.flann, scipy and numpy.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libssf_oracle.so")
_lib = None


class Params(ctypes.Structure):
    _fields_ = [("max_correspondence_dist", ctypes.c_float), ("num_iterations", ctypes.c_int32),
                ("acceptable_mean_error", ctypes.c_float), ("transformation_epsilon", ctypes.c_float)]


class Result(ctypes.Structure):
    _fields_ = [("transformation", ctypes.c_float * 16), ("error", ctypes.c_float), ("iterations", ctypes.c_int32),
                ("has_converged", ctypes.c_int32), ("n_searches", ctypes.c_int32), ("k_final", ctypes.c_int32),
                ("aborted", ctypes.c_int32)]

    @property
    def T(self) -> np.ndarray:
        """4x4 float32, row-major numpy view of the column-major field."""
        return np.array(self.transformation, dtype=np.float32).reshape(4, 4).T.copy()


class _Trace(ctypes.Structure):
    _fields_ = [("cap_searches", ctypes.c_int32), ("n_source", ctypes.c_int32), ("count", ctypes.c_void_p),
                ("queries", ctypes.c_void_p), ("rows", ctypes.c_void_p), ("idx", ctypes.c_void_p),
                ("d2", ctypes.c_void_p), ("iter_err", ctypes.c_void_p), ("iter_searched", ctypes.c_void_p)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "ssf_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i64, i32, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
        L.ssf_oracle_kdtree_build.restype = vp
        L.ssf_oracle_kdtree_build.argtypes = [vp, i64, i32]
        L.ssf_oracle_kdtree_free.argtypes = [vp]
        L.ssf_oracle_kdtree_nn.argtypes = [vp, vp, i64, i32, vp, vp, i32]
        L.ssf_oracle_nn_brute.argtypes = [vp, i64, i32, vp, i64, i32, vp, vp]
        L.ssf_oracle_max_threads.restype = i32
        L.ssf_oracle_svd3.argtypes = [vp, vp, vp, vp]
        L.ssf_oracle_kabsch.argtypes = [vp, vp, i64, vp]
        L.ssf_oracle_icp_reference.restype = i32
        L.ssf_oracle_icp_reference.argtypes = [vp, vp, i64, i32, vp, ctypes.POINTER(Params), ctypes.POINTER(Result),
                                               vp, vp, i32]
        L.ssf_oracle_icp_gn.restype = i32
        L.ssf_oracle_icp_gn.argtypes = [vp, vp, vp, i64, i32, vp, ctypes.POINTER(Params), i32,
                                        ctypes.POINTER(Result), vp, i32]
        L.ssf_oracle_icp_o3d.restype = i32
        L.ssf_oracle_icp_o3d.argtypes = [vp, vp, i64, i32, vp, ctypes.POINTER(Params), ctypes.POINTER(Result), vp, vp,
                                         i32]
        L.ssf_oracle_voxel_grid.restype = i64
        L.ssf_oracle_voxel_grid.argtypes = [vp, i64, i32, f32, vp, vp]
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().ssf_oracle_max_threads())


def _f32(a, cols=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols is not None and (a.ndim != 2 or a.shape[1] not in cols):
        raise ValueError(f"expected (n, {cols}) array, got {a.shape}")
    return a


def _colmajor(T) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).reshape(16)


class KdTree:
    """Exact 1-NN over an (M, 3|4) float32 cloud; ties -> lowest index."""

    def __init__(self, xyz):
        self.xyz = _f32(xyz, (3, 4))
        self.n = self.xyz.shape[0]
        self._h = lib().ssf_oracle_kdtree_build(self.xyz.ctypes.data, self.n, self.xyz.shape[1])

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ssf_oracle_kdtree_free(self._h)
            self._h = None

    def nn(self, q, threads: int = 1):
        q = _f32(q, (3, 4))
        idx = np.empty(q.shape[0], dtype=np.int32)
        d2 = np.empty(q.shape[0], dtype=np.float32)
        lib().ssf_oracle_kdtree_nn(self._h, q.ctypes.data, q.shape[0], q.shape[1], idx.ctypes.data, d2.ctypes.data,
                                   threads)
        return idx, d2


def nn_brute(map_xyz, q):
    m, q = _f32(map_xyz, (3, 4)), _f32(q, (3, 4))
    idx = np.empty(q.shape[0], dtype=np.int32)
    d2 = np.empty(q.shape[0], dtype=np.float32)
    lib().ssf_oracle_nn_brute(m.ctypes.data, m.shape[0], m.shape[1], q.ctypes.data, q.shape[0], q.shape[1],
                              idx.ctypes.data, d2.ctypes.data)
    return idx, d2


def svd3(H):
    """H (3x3, numpy row-major) -> U, S, V with H = U diag(S) V^T."""
    Hc = np.ascontiguousarray(np.asarray(H, dtype=np.float32).T)
    U, S, V = np.empty((3, 3), np.float32), np.empty(3, np.float32), np.empty((3, 3), np.float32)
    lib().ssf_oracle_svd3(Hc.ctypes.data, U.ctypes.data, S.ctypes.data, V.ctypes.data)
    return U.T.copy(), S, V.T.copy()


def kabsch(src, tgt) -> np.ndarray:
    s, t = _f32(src, (3,)), _f32(tgt, (3,))
    T = np.empty(16, np.float32)
    lib().ssf_oracle_kabsch(s.ctypes.data, t.ctypes.data, s.shape[0], T.ctypes.data)
    return T.reshape(4, 4).T.copy()


class Trace:
    def __init__(self, cap_searches: int, n_source: int, num_iterations: int):
        n = max(1, n_source)
        self.count = np.zeros(cap_searches, np.int32)
        self.queries = np.zeros((cap_searches, n, 3), np.float32)
        self.rows = np.zeros((cap_searches, n), np.int32)
        self.idx = np.zeros((cap_searches, n), np.int32)
        self.d2 = np.zeros((cap_searches, n), np.float32)
        self.iter_err = np.zeros(max(1, num_iterations), np.float32)
        self.iter_searched = np.zeros(max(1, num_iterations), np.int32)
        self.c = _Trace(cap_searches, n_source, self.count.ctypes.data, self.queries.ctypes.data,
                        self.rows.ctypes.data, self.idx.ctypes.data, self.d2.ctypes.data, self.iter_err.ctypes.data,
                        self.iter_searched.ctypes.data)


def icp_reference(tree: KdTree, src, T_init, max_correspondence_dist=0.5, num_iterations=10,
                  acceptable_mean_error=0.05, transformation_epsilon=1e-5, trace: bool = False, threads: int = 1):
    """Restatement of ICPPointToPoint::calculateAlignment (icp_point_to_point.cpp:185-254).

    Returns (Result, corr[int32 per source row], Trace|None)."""
    s = _f32(src, (3, 4))
    prm = Params(max_correspondence_dist, num_iterations, acceptable_mean_error, transformation_epsilon)
    res = Result()
    corr = np.empty(s.shape[0], np.int32)
    tr = Trace(num_iterations + 1, s.shape[0], num_iterations) if trace else None
    Tc = _colmajor(T_init)
    rc = lib().ssf_oracle_icp_reference(tree._h, s.ctypes.data, s.shape[0], s.shape[1], Tc.ctypes.data,
                                        ctypes.byref(prm), ctypes.byref(res),
                                        ctypes.addressof(tr.c) if tr else None, corr.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError(f"oracle icp_reference rc={rc}")
    return res, corr, tr


def icp_gn(tree: KdTree, src, T_init, mode: str = "p2p", normals=None, max_correspondence_dist=0.5,
           num_iterations=10, acceptable_mean_error=0.0, transformation_epsilon=0.0, threads: int = 1):
    s = _f32(src, (3, 4))
    nrm = _f32(normals, (4,)) if normals is not None else None
    prm = Params(max_correspondence_dist, num_iterations, acceptable_mean_error, transformation_epsilon)
    res = Result()
    corr = np.empty(s.shape[0], np.int32)
    Tc = _colmajor(T_init)
    rc = lib().ssf_oracle_icp_gn(tree._h, nrm.ctypes.data if nrm is not None else None, s.ctypes.data, s.shape[0],
                                 s.shape[1], Tc.ctypes.data, ctypes.byref(prm), 1 if mode == "p2plane" else 0,
                                 ctypes.byref(res), corr.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError(f"oracle icp_gn rc={rc}")
    return res, corr


def icp_o3d(tree: KdTree, src, T_init, max_correspondence_distance=0.5, max_iteration=30, threads: int = 1):
    """Open3D registration_icp control flow on float32 geometry; the metric radius is squared here."""
    s = _f32(src, (3, 4))
    thr = np.float32(max_correspondence_distance) * np.float32(max_correspondence_distance)
    prm = Params(float(thr), max_iteration, 0.0, 0.0)
    res = Result()
    corr = np.empty(s.shape[0], np.int32)
    fit = ctypes.c_float(0)
    Tc = _colmajor(T_init)
    rc = lib().ssf_oracle_icp_o3d(tree._h, s.ctypes.data, s.shape[0], s.shape[1], Tc.ctypes.data, ctypes.byref(prm),
                                  ctypes.byref(res), ctypes.addressof(fit), corr.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError(f"oracle icp_o3d rc={rc}")
    return res, float(fit.value), corr


def voxel_grid(xyz, leaf: float):
    """pcl::VoxelGrid semantics; returns ((n_out, 4) float32, overflow_refused: bool)."""
    a = _f32(xyz, (3, 4))
    out = np.empty((max(1, a.shape[0]), 4), np.float32)
    st = ctypes.c_int32(0)
    n = lib().ssf_oracle_voxel_grid(a.ctypes.data, a.shape[0], a.shape[1], leaf, out.ctypes.data, ctypes.addressof(st))
    return out[:n].copy(), bool(st.value)


def voxel_grid_o3d(xyz, voxel: float) -> np.ndarray:
    """open3d voxel_down_sample semantics (double arithmetic, origin min_bound - voxel / 2); (n_out, 3) float32."""
    a = _f32(xyz, (3, 4))
    out = np.empty((max(1, a.shape[0]), 4), np.float32)
    L = lib()
    L.ssf_oracle_voxel_grid_o3d.restype = ctypes.c_int64
    L.ssf_oracle_voxel_grid_o3d.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
    n = L.ssf_oracle_voxel_grid_o3d(a.ctypes.data, a.shape[0], a.shape[1], float(voxel), out.ctypes.data)
    return out[:n, :3].copy()


def _preproc_setup():
    L = lib()
    if not getattr(L, "_preproc_ready", False):
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        L.ssf_oracle_subsample.restype = i64
        L.ssf_oracle_subsample.argtypes = [vp, i64, i32, i64, vp]
        L.ssf_oracle_remove_floor.restype = i64
        L.ssf_oracle_remove_floor.argtypes = [vp, i64, i32, vp]
        L.ssf_oracle_crop_radius.restype = i64
        L.ssf_oracle_crop_radius.argtypes = [vp, i64, i32, vp, ctypes.c_double, vp, vp]
        L._preproc_ready = True
    return L


def subsample(xyz, step: int) -> np.ndarray:
    """applyUniformSubsample (point_cloud_processing.hpp:55-74)."""
    a = _f32(xyz, (3, 4))
    out = np.empty((max(1, a.shape[0]), 4), np.float32)
    n = _preproc_setup().ssf_oracle_subsample(a.ctypes.data, a.shape[0], a.shape[1], step, out.ctypes.data)
    return out[:n, :3].copy()


def remove_floor(xyz) -> np.ndarray:
    """removeFloor (point_cloud_processing.hpp:76-92)."""
    a = _f32(xyz, (3, 4))
    out = np.empty((max(1, a.shape[0]), 4), np.float32)
    n = _preproc_setup().ssf_oracle_remove_floor(a.ctypes.data, a.shape[0], a.shape[1], out.ctypes.data)
    return out[:n, :3].copy()


def crop_radius(T, radius: float, xyz):
    """cropPointCloudThroughRadius (point_cloud_processing.hpp:31-53); returns (points, indices)."""
    a = _f32(xyz, (3, 4))
    c = np.ascontiguousarray(np.asarray(T, np.float32)[:3, 3])
    out = np.empty((max(1, a.shape[0]), 4), np.float32)
    idx = np.empty(max(1, a.shape[0]), np.int32)
    n = _preproc_setup().ssf_oracle_crop_radius(a.ctypes.data, a.shape[0], a.shape[1], c.ctypes.data, float(radius),
                                                out.ctypes.data, idx.ctypes.data)
    return out[:n, :3].copy(), idx[:n].copy()


class BfaParams(ctypes.Structure):
    """Defaults = localization_node.cpp:38-43."""
    _fields_ = [("x_step", ctypes.c_float), ("y_step", ctypes.c_float), ("z_step", ctypes.c_float),
                ("x_range", ctypes.c_float), ("y_range", ctypes.c_float), ("z_range", ctypes.c_float),
                ("yaw_step", ctypes.c_float), ("yaw_range", ctypes.c_float), ("mean_error_threshold", ctypes.c_float)]

    @staticmethod
    def node_defaults() -> "BfaParams":
        pi = np.float32(np.pi)
        return BfaParams(0.1, 0.1, 0.05, 1.5, 1.5, 0.1, float(pi / np.float32(18.0)), float(pi / np.float32(6.0)), 0.1)


def bfa_poses(T_prev, prm: BfaParams) -> np.ndarray:
    """Candidate transforms prev * T(x, y, z, yaw) in the reference's loop order, (n, 4, 4) row-major."""
    L = lib()
    L.ssf_oracle_bfa_poses.restype = ctypes.c_int64
    L.ssf_oracle_bfa_poses.argtypes = [ctypes.c_void_p, ctypes.POINTER(BfaParams), ctypes.c_void_p]
    Tc = _colmajor(T_prev)
    n = L.ssf_oracle_bfa_poses(Tc.ctypes.data, ctypes.byref(prm), None)
    out = np.empty((n, 16), np.float32)
    L.ssf_oracle_bfa_poses(Tc.ctypes.data, ctypes.byref(prm), out.ctypes.data)
    return out.reshape(n, 4, 4).transpose(0, 2, 1).copy()


def bfa_align(tree: KdTree, src, T_prev, prm: BfaParams, no_early_exit: bool = False, threads: int = 1):
    """BruteForceAlignment::alignClouds; returns (success, T_best 4x4, best_score, scores)."""
    L = lib()
    L.ssf_oracle_bfa_align.restype = ctypes.c_int
    L.ssf_oracle_bfa_align.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.POINTER(BfaParams), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    s = _f32(src, (3, 4))
    n_pose = bfa_poses(T_prev, prm).shape[0]
    Tc = _colmajor(T_prev)
    T_out = np.empty(16, np.float32)
    score = ctypes.c_float(0)
    ok = ctypes.c_int32(0)
    scores = np.empty(n_pose, np.float32)
    rc = L.ssf_oracle_bfa_align(tree._h, s.ctypes.data, s.shape[0], s.shape[1], Tc.ctypes.data, ctypes.byref(prm),
                                1 if no_early_exit else 0, T_out.ctypes.data, ctypes.addressof(score),
                                ctypes.addressof(ok), scores.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError("oracle bfa_align failed")
    return bool(ok.value), T_out.reshape(4, 4).T.copy(), float(score.value), scores
