"""ssf_gpu -- Python host side of libssf_gpu.so (ctypes + numpy only, no PyTorch).

Mirrors the two registration interfaces of viniciusvidal2/slam-sensor-fusion:

* ``ICPPointToPoint`` -- same constructor, setters and ``calculateAlignment`` as the C++
  class of ``localization/include/localization/icp_point_to_point.h:41-85`` (clouds are
  (N, 3|4) float32 arrays instead of ``pcl::PointCloud``, transforms are 4x4 arrays);
* ``registration_icp`` / ``voxel_down_sample`` -- the Open3D calls made by
  ``localization_python/localization_python/localization_node.py:47,233-237``.

Every call runs on the GPU through the C ABI; nothing here computes a result on the CPU.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import (MODE_GN_P2P, MODE_GN_P2PLANE, MODE_O3D_P2P, MODE_REFERENCE, REDUCE_FAST, REDUCE_STRICT, IcpParams,
                   IcpResult, SsfError)

__all__ = ["Context", "ICPPointToPoint", "ICPResult", "Batch", "ResidentMap", "from_pointcloud2", "nccl_unique_id",
           "registration_icp", "voxel_down_sample",
           "RegistrationResult", "BruteForceAlignment", "applyUniformSubsample", "removeFloor", "cropPointCloudThroughRadius", "ICPConvergenceCriteria", "TransformationEstimationPointToPoint",
           "TransformationEstimationPointToPlane", "SsfError", "default_context",
           "MODE_REFERENCE", "MODE_GN_P2P", "MODE_GN_P2PLANE", "MODE_O3D_P2P", "REDUCE_STRICT", "REDUCE_FAST"]


def _cloud(a) -> np.ndarray:
    a = np.asarray(a)
    if a.ndim != 2 or a.shape[1] not in (3, 4):
        raise ValueError(f"cloud must be (N, 3) or (N, 4), got {a.shape}")
    return np.ascontiguousarray(a, dtype=np.float32)


def _colmajor(T) -> np.ndarray:
    T = np.asarray(T, dtype=np.float32)
    if T.shape != (4, 4):
        raise ValueError("transform must be 4x4")
    return np.ascontiguousarray(T.T).reshape(16)


def _rowmajor(t16) -> np.ndarray:
    return np.array(t16, dtype=np.float32).reshape(4, 4).T.copy()


class Context:
    """One CUDA device (``ssf_ctx``)."""

    def __init__(self, device: int = 0):
        self._h = ctypes.c_void_p()
        capi.check(capi.lib().ssf_ctx_create(device, ctypes.byref(self._h)))
        self.device = device

    def synchronize(self) -> None:
        capi.check(capi.lib().ssf_ctx_synchronize(self._h))

    @property
    def stream(self) -> int:
        return int(capi.lib().ssf_ctx_stream(self._h) or 0)

    def time_searches(self, enable: bool) -> None:
        capi.check(capi.lib().ssf_ctx_time_searches(self._h, 1 if enable else 0))

    def search_time(self):
        """(total ms, launches) of the NN-search kernels since the last call (CUDA events)."""
        ms, n = ctypes.c_double(0), ctypes.c_uint64(0)
        capi.check(capi.lib().ssf_ctx_search_time(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, int(n.value)

    def search_times(self, cap: int = 4096):
        """Per-launch times (ms) of the NN-search kernels recorded so far (list not cleared)."""
        buf, n = (ctypes.c_float * cap)(), ctypes.c_uint64(0)
        capi.check(capi.lib().ssf_ctx_search_times(self._h, buf, cap, ctypes.byref(n)))
        return [float(buf[i]) for i in range(min(cap, int(n.value)))]

    def close(self) -> None:
        if getattr(self, "_h", None):
            capi.lib().ssf_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


@dataclass
class ICPResult:
    """``struct ICPResult`` (icp_point_to_point.h:28-39) plus diagnostics."""
    transformation: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    error: float = 1e6
    iterations: int = 0
    has_converged: bool = False
    n_searches: int = 0
    k_final: int = 0
    aborted: bool = False
    fitness: float = 0.0
    n_source: int = 0
    device_ms: float = 0.0

    @staticmethod
    def from_c(r: IcpResult) -> "ICPResult":
        return ICPResult(_rowmajor(r.transformation), float(r.error), int(r.iterations), bool(r.has_converged),
                         int(r.n_searches), int(r.k_final), bool(r.aborted), float(r.fitness), int(r.n_source),
                         float(r.device_ms))


class ICPPointToPoint:
    """Drop-in for the reference class of the same name (icp_point_to_point.h:41-85).

    ``mode`` / ``reduce`` select the solver (see ssf.h); the defaults reproduce the reference.
    """

    def __init__(self, max_correspondence_dist: float, num_iterations: int, acceptable_mean_error: float,
                 transformation_epsilon: float, *, mode: int = MODE_REFERENCE, reduce: int = REDUCE_STRICT,
                 context: Context | None = None):
        self._ctx = context or default_context()
        self._p = IcpParams(max_correspondence_dist, num_iterations, acceptable_mean_error, transformation_epsilon,
                            mode, reduce, 0, 0.0)
        self._h = ctypes.c_void_p()
        capi.check(capi.lib().ssf_icp_create(self._ctx._h, ctypes.byref(self._p), ctypes.byref(self._h)))
        self._n_source = 0

    # -- setters (icp_point_to_point.cpp:14-55) ------------------------------------------------
    def _push(self) -> None:
        capi.check(capi.lib().ssf_icp_set_params(self._h, ctypes.byref(self._p)))

    def setMaxCorrespondenceDist(self, v: float) -> None:
        self._p.max_correspondence_dist = v
        self._push()

    def setNumIterations(self, v: int) -> None:
        self._p.num_iterations = v
        self._push()

    def setTransformationEpsilon(self, v: float) -> None:
        self._p.transformation_epsilon = v
        self._push()

    def setAcceptableMeanError(self, v: float) -> None:
        self._p.acceptable_mean_error = v
        self._push()

    def setDebugMode(self, v: bool) -> None:
        self._p.debug = 1 if v else 0
        self._push()

    def setSourceVoxelLeaf(self, leaf: float) -> None:
        """> 0: voxel-grid downsample every source scan on the device before the loop."""
        self._p.source_voxel_leaf = leaf
        self._push()

    def setMode(self, mode: int, reduce: int | None = None) -> None:
        self._p.mode = mode
        if reduce is not None:
            self._p.reduce = reduce
        self._push()

    def setInitialTransformation(self, T) -> None:
        t = _colmajor(T)
        capi.check(capi.lib().ssf_icp_set_initial(self._h, t.ctypes.data))

    def setSourcePointCloud(self, cloud) -> None:
        c = _cloud(cloud)
        capi.check(capi.lib().ssf_icp_set_source(self._h, c.ctypes.data, c.shape[0], c.strides[0]))
        self._n_source = c.shape[0]

    def setTargetPointCloud(self, cloud, normals=None) -> None:
        c = _cloud(cloud)
        if normals is not None:
            n = _cloud(normals)
            if n.shape[0] != c.shape[0]:
                raise ValueError("normals and cloud sizes differ")
            capi.check(capi.lib().ssf_icp_set_target(self._h, c.ctypes.data, c.shape[0], c.strides[0], n.ctypes.data,
                                                     n.strides[0]))
        else:
            capi.check(capi.lib().ssf_icp_set_target(self._h, c.ctypes.data, c.shape[0], c.strides[0], None, 0))

    def setTargetShard(self, shard: dict) -> None:
        """Map sharding: ``shard`` is the dict returned by ``ssf_gpu.shard.shard_map``."""
        from .shard import ShardInfo
        c = _cloud(shard["points"])
        n = _cloud(shard["normals"]) if shard.get("normals") is not None else None
        gi = np.ascontiguousarray(shard["global_index"], np.int32)
        info = ShardInfo((ctypes.c_float * 3)(*[float(v) for v in shard["origin"]]), float(shard["cell"]),
                         int(shard["own"][0]), int(shard["own"][1]))
        capi.check(capi.lib().ssf_icp_set_target_shard(
            self._h, c.ctypes.data, c.shape[0], c.strides[0], n.ctypes.data if n is not None else None,
            n.strides[0] if n is not None else 0, gi.ctypes.data, ctypes.byref(info)))

    def setAllreduce(self, hook) -> None:
        """``hook``: a ``shard.ALLREDUCE_FN`` (e.g. ``shard.torch_allreduce_hook``); kept alive here."""
        self._allreduce_hook = hook
        capi.check(capi.lib().ssf_icp_set_allreduce(self._h, ctypes.cast(hook, ctypes.c_void_p), None))

    def exchangeCreate(self, rank: int, world: int, max_scans: int) -> bytes:
        """This rank's exchange buffer for the in-kernel sum across map shards; returns its handle blob
        (CUDA IPC handle + device identity; gather the blobs of all ranks, then ``exchangeOpen``)."""
        h = (ctypes.c_ubyte * capi.XCH_HANDLE_BYTES)()
        capi.check(capi.lib().ssf_icp_exchange_create(self._h, int(rank), int(world), int(max_scans), h))
        return bytes(h)

    def exchangeOpen(self, handles) -> None:
        """``handles``: the handle blobs of all ranks, in rank order."""
        blob = b"".join(bytes(x) for x in handles)
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        capi.check(capi.lib().ssf_icp_exchange_open(self._h, buf))

    def exchangeClose(self) -> None:
        capi.check(capi.lib().ssf_icp_exchange_close(self._h))

    def ncclInit(self, unique_id: bytes, rank: int, world: int) -> None:
        """Map sharding through ``ncclAllReduce`` on the library's stream (collective call: every rank, same id)."""
        buf = (ctypes.c_ubyte * capi.NCCL_ID_BYTES).from_buffer_copy(bytes(unique_id))
        capi.check(capi.lib().ssf_icp_nccl_init(self._h, buf, int(rank), int(world)))

    def ncclClose(self) -> None:
        capi.check(capi.lib().ssf_icp_nccl_close(self._h))

    # -- calculateAlignment (icp_point_to_point.cpp:185-254) ---------------------------------------
    def calculateAlignment(self) -> ICPResult:
        r = IcpResult()
        capi.check(capi.lib().ssf_icp_align(self._h, ctypes.byref(r)))
        return ICPResult.from_c(r)

    # -- extras ----------------------------------------------------------------------------------
    def correspondences(self) -> np.ndarray:
        out = np.empty(self._n_source, np.int32)
        capi.check(capi.lib().ssf_icp_get_correspondences(self._h, out.ctypes.data, out.shape[0]))
        return out

    def trace(self):
        n = max(1, int(self._p.num_iterations))
        err = np.empty(n, np.float32)
        srch = np.empty(n, np.int32)
        capi.check(capi.lib().ssf_icp_get_trace(self._h, err.ctypes.data, srch.ctypes.data, n))
        return err, srch

    def nearest(self, queries, max_sqdist: float):
        """k=1 search + ``d2 < max_sqdist`` for already-transformed queries (cpp:64-70)."""
        q = _cloud(queries)
        idx = np.empty(q.shape[0], np.int32)
        d2 = np.empty(q.shape[0], np.float32)
        capi.check(capi.lib().ssf_nn_search(self._h, q.ctypes.data, q.shape[0], q.strides[0], max_sqdist,
                                            idx.ctypes.data, d2.ctypes.data))
        return idx, d2

    def nearest_bench(self, queries, max_sqdist: float, reps: int = 10) -> float:
        """Average device milliseconds of one search pass over ``queries`` (resident in HBM)."""
        q = _cloud(queries)
        ms = ctypes.c_float(0)
        capi.check(capi.lib().ssf_nn_search_bench(self._h, q.ctypes.data, q.shape[0], q.strides[0], max_sqdist, reps,
                                                  ctypes.byref(ms), None, None))
        return float(ms.value)

    def align_batch(self, scans: list, inits) -> list[ICPResult]:
        """Register a list of scans against the target in one call (offline reprocessing)."""
        clouds = [_cloud(s) for s in scans]
        width = clouds[0].shape[1] if clouds else 4
        if any(c.shape[1] != width for c in clouds):
            clouds = [np.ascontiguousarray(np.pad(c, ((0, 0), (0, 4 - c.shape[1])))) if c.shape[1] == 3 else c
                      for c in clouds]
            width = 4
        cat = np.ascontiguousarray(np.concatenate(clouds, axis=0)) if clouds else np.zeros((0, 4), np.float32)
        n_pts = (ctypes.c_size_t * len(clouds))(*[c.shape[0] for c in clouds])
        T = np.ascontiguousarray(np.stack([_colmajor(t) for t in inits]))
        res = (IcpResult * len(clouds))()
        capi.check(capi.lib().ssf_icp_align_batch(self._h, cat.ctypes.data, n_pts, len(clouds), 4 * width,
                                                  T.ctypes.data, res))
        return [ICPResult.from_c(r) for r in res]

    def close(self) -> None:
        if getattr(self, "_h", None):
            capi.lib().ssf_icp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    """``ncclGetUniqueId`` through the library: call on rank 0, carry the bytes to every rank, ``ncclInit``."""
    h = (ctypes.c_ubyte * capi.NCCL_ID_BYTES)()
    capi.check(capi.lib().ssf_nccl_unique_id(h))
    return bytes(h)


class Batch:
    """Scans resident in HBM, aligned together (``ssf_batch``)."""

    def __init__(self, icp: ICPPointToPoint, max_scans: int, max_total_points: int):
        self._icp = icp
        self._h = ctypes.c_void_p()
        capi.check(capi.lib().ssf_batch_create(icp._h, max_scans, max_total_points, ctypes.byref(self._h)))
        self.n_scans = 0

    def upload_ptr(self, ptr: int, n_pts, stride_bytes: int = 16, wait: bool = True) -> None:
        """wait=False: asynchronous copy -- keep the host buffer untouched until results() returns."""
        arr = (ctypes.c_size_t * len(n_pts))(*[int(n) for n in n_pts])
        fn = capi.lib().ssf_batch_upload if wait else capi.lib().ssf_batch_upload_async
        capi.check(fn(self._h, ptr, arr, len(n_pts), stride_bytes))
        self.n_scans = len(n_pts)

    def upload(self, scans: list) -> None:
        clouds = [_cloud(s) for s in scans]
        clouds = [np.ascontiguousarray(np.pad(c, ((0, 0), (0, 1)))) if c.shape[1] == 3 else c for c in clouds]
        cat = np.ascontiguousarray(np.concatenate(clouds, axis=0))
        self.upload_ptr(cat.ctypes.data, [c.shape[0] for c in clouds], 16)

    def set_initial_ptr(self, ptr: int) -> None:
        capi.check(capi.lib().ssf_batch_set_initial(self._h, ptr))

    def set_initial(self, inits) -> None:
        T = np.ascontiguousarray(np.stack([_colmajor(t) for t in inits]))
        self.set_initial_ptr(T.ctypes.data)

    def run(self) -> None:
        capi.check(capi.lib().ssf_batch_run(self._h))

    def results_into(self, res_array) -> None:
        capi.check(capi.lib().ssf_batch_results(self._h, res_array, self.n_scans))

    def results(self) -> list[ICPResult]:
        res = (IcpResult * self.n_scans)()
        self.results_into(res)
        return [ICPResult.from_c(r) for r in res]

    def search_stats(self):
        """Per search launch of the last run: (queries answered, queries that needed a walk)."""
        cap = 128
        a, w = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
        n = ctypes.c_size_t(0)
        capi.check(capi.lib().ssf_batch_search_stats(self._h, a.ctypes.data, w.ctypes.data, cap, ctypes.byref(n)))
        k = min(cap, int(n.value))
        return a[:k].copy(), w[:k].copy()

    def close(self) -> None:
        if getattr(self, "_h", None):
            capi.lib().ssf_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- Open3D-shaped surface (localization_node.py:47,233-237) ------------------------------------
@dataclass
class ICPConvergenceCriteria:
    max_iteration: int = 30
    relative_fitness: float = 1e-6
    relative_rmse: float = 1e-6


class TransformationEstimationPointToPoint:
    mode = MODE_O3D_P2P


class TransformationEstimationPointToPlane:
    mode = MODE_GN_P2PLANE


@dataclass
class RegistrationResult:
    transformation: np.ndarray
    fitness: float
    inlier_rmse: float
    correspondence_set: np.ndarray
    iterations: int = 0
    converged: bool = False


def registration_icp(source, target, max_correspondence_distance: float, init=None, estimation_method=None,
                     criteria: ICPConvergenceCriteria | None = None, *, target_normals=None,
                     context: Context | None = None) -> RegistrationResult:
    """``o3d.pipelines.registration.registration_icp`` shaped call on the GPU.

    ``source``/``target``: (N, 3) arrays (float64 accepted, cast to float32 at the boundary);
    ``target`` may also be an ``ICPPointToPoint`` whose target is already resident in HBM.
    The metric radius is squared here because the C ABI compares squared distances.
    """
    criteria = criteria or ICPConvergenceCriteria()
    estimation_method = estimation_method or TransformationEstimationPointToPoint()
    mode = estimation_method.mode
    thr = float(np.float32(max_correspondence_distance) * np.float32(max_correspondence_distance))
    if isinstance(target, ICPPointToPoint):
        icp = target
        icp._p.max_correspondence_dist = thr
        icp._p.num_iterations = criteria.max_iteration
        icp._p.mode = mode
        icp._p.acceptable_mean_error = 0.0
        icp._p.transformation_epsilon = 1e-6 if mode != MODE_O3D_P2P else 0.0
        icp._push()
    else:
        icp = ICPPointToPoint(thr, criteria.max_iteration, 0.0, 1e-6 if mode != MODE_O3D_P2P else 0.0, mode=mode,
                              context=context)
        icp.setTargetPointCloud(target, target_normals)
    icp.setSourcePointCloud(source)
    icp.setInitialTransformation(np.eye(4) if init is None else init)
    r = icp.calculateAlignment()
    corr = icp.correspondences()
    rows = np.nonzero(corr >= 0)[0]
    cs = np.stack([rows, corr[rows]], axis=1).astype(np.int32) if rows.size else np.zeros((0, 2), np.int32)
    return RegistrationResult(r.transformation.astype(np.float64), r.fitness, r.error, cs, r.iterations,
                              r.has_converged)


def voxel_down_sample(xyz, voxel_size: float, context: Context | None = None, semantics: str = "pcl"):
    """Voxel-grid downsample on the device.  ``semantics="pcl"``: pcl::VoxelGrid (the C++ node's map merge,
    global_map_frames_manager.cpp:143-146; also the scan downsample in front of the loop);
    ``semantics="open3d"``: ``pcd.voxel_down_sample(voxel_size)`` of the Python node (localization_node.py:47) --
    voxel origin ``min_bound - voxel_size / 2``, double centroids, output in ascending voxel index.
    Returns (M, 3) float32."""
    ctx = context or default_context()
    c = _cloud(xyz)
    out = np.empty((max(1, c.shape[0]), 4), np.float32)
    n_out = ctypes.c_size_t(0)
    if semantics == "open3d":
        capi.check(capi.lib().ssf_voxel_downsample_o3d(ctx._h, c.ctypes.data, c.shape[0], c.strides[0], float(voxel_size),
                                                       out.ctypes.data, ctypes.byref(n_out)))
        return out[:n_out.value, :3].copy()
    if semantics != "pcl":
        raise ValueError("semantics must be 'pcl' or 'open3d'")
    refused = ctypes.c_int(0)
    capi.check(capi.lib().ssf_voxel_downsample(ctx._h, c.ctypes.data, c.shape[0], c.strides[0], voxel_size,
                                               out.ctypes.data, ctypes.byref(n_out), ctypes.byref(refused)))
    return out[:n_out.value, :3].copy()


# ---- point_cloud_processing.hpp (reference localization/include/localization/point_cloud_processing.hpp) ----
def applyUniformSubsample(cloud, point_step: int, context: Context | None = None) -> np.ndarray:
    """Every ``point_step``-th point; unchanged when the cloud is shorter than the step (hpp:55-74)."""
    ctx = context or default_context()
    c = _cloud(cloud)
    out = np.empty((max(1, c.shape[0]), 4), np.float32)
    n_out = ctypes.c_size_t(0)
    capi.check(capi.lib().ssf_cloud_subsample(ctx._h, c.ctypes.data, c.shape[0], c.strides[0], point_step,
                                              out.ctypes.data, ctypes.byref(n_out)))
    return out[:n_out.value, :3].copy()


def removeFloor(cloud, context: Context | None = None) -> np.ndarray:
    """Keep points with z > 0, order preserved (hpp:76-92)."""
    ctx = context or default_context()
    c = _cloud(cloud)
    out = np.empty((max(1, c.shape[0]), 4), np.float32)
    n_out = ctypes.c_size_t(0)
    capi.check(capi.lib().ssf_cloud_remove_floor(ctx._h, c.ctypes.data, c.shape[0], c.strides[0], out.ctypes.data,
                                                 ctypes.byref(n_out)))
    return out[:n_out.value, :3].copy()


def cropPointCloudThroughRadius(T, radius: float, cloud, context: Context | None = None, return_indices: bool = False):
    """Points within ``radius`` of the translation of ``T``, ordered by distance like
    pcl::search::KdTree::radiusSearch (hpp:31-53)."""
    ctx = context or default_context()
    c = _cloud(cloud)
    center = np.ascontiguousarray(np.asarray(T, np.float32)[:3, 3])
    out = np.empty((max(1, c.shape[0]), 4), np.float32)
    idx = np.empty(max(1, c.shape[0]), np.int32)
    n_out = ctypes.c_size_t(0)
    capi.check(capi.lib().ssf_cloud_crop_radius(ctx._h, c.ctypes.data, c.shape[0], c.strides[0], center.ctypes.data,
                                                float(radius), out.ctypes.data, ctypes.byref(n_out), idx.ctypes.data))
    pts = out[:n_out.value, :3].copy()
    return (pts, idx[:n_out.value].copy()) if return_indices else pts


def from_pointcloud2(data, n_points: int, point_step: int, offsets=(0, 4, 8), is_bigendian: bool = False,
                     context: Context | None = None) -> np.ndarray:
    """``pcl::fromROSMsg`` for the float32 x / y / z fields of a ``sensor_msgs/PointCloud2`` byte buffer
    (localization_node.cpp:290-291): (n_points, 3) float32, extracted on the device."""
    ctx = context or default_context()
    buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    if buf.size < n_points * point_step:
        raise ValueError("PointCloud2 data shorter than width * height * point_step")
    out = np.empty((max(1, n_points), 4), np.float32)
    capi.check(capi.lib().ssf_cloud_from_pointcloud2(ctx._h, buf.ctypes.data, n_points, point_step, offsets[0], offsets[1],
                                                     offsets[2], 1 if is_bigendian else 0, out.ctypes.data))
    return out[:n_points, :3].copy()


class ResidentMap:
    """The map cloud kept in HBM (``ssf_map``): uploaded -- or merged from the recorder's PCD tiles by
    ``from_pcd_folder`` = ``GlobalMapFramesManager::getMapCloud`` (global_map_frames_manager.cpp:93-151) --
    once; the re-crop of localization_node.cpp:300-305 is then ``crop_to_target`` (a window change in HBM)."""

    def __init__(self, cloud=None, context: Context | None = None, _handle=None):
        self._ctx = context or default_context()
        self._h = ctypes.c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            c = _cloud(cloud)
            capi.check(capi.lib().ssf_map_create(self._ctx._h, c.ctypes.data, c.shape[0], c.strides[0], ctypes.byref(self._h)))

    @staticmethod
    def from_pcd_folder(data_folder: str, map_name: str, voxel_size: float, save: bool = True,
                        context: Context | None = None) -> "ResidentMap":
        ctx = context or default_context()
        h = ctypes.c_void_p()
        capi.check(capi.lib().ssf_map_from_pcd_folder(ctx._h, data_folder.encode(), map_name.encode(), voxel_size,
                                                      1 if save else 0, ctypes.byref(h)))
        return ResidentMap(context=ctx, _handle=h)

    def __len__(self) -> int:
        return int(capi.lib().ssf_map_size(self._h))

    @property
    def ingest_ms(self) -> float:
        return float(capi.lib().ssf_map_ingest_ms(self._h))

    @property
    def merge_ms(self) -> float:
        """Device ms of the voxel filter over the concatenated tiles (part of ``ingest_ms``)."""
        return float(capi.lib().ssf_map_merge_ms(self._h))

    def download(self) -> np.ndarray:
        out = np.empty((max(1, len(self)), 4), np.float32)
        capi.check(capi.lib().ssf_map_download(self._h, out.ctypes.data, out.shape[0]))
        return out[:len(self), :3].copy()

    def subsample(self, point_step: int) -> None:
        capi.check(capi.lib().ssf_map_subsample(self._h, point_step))

    def crop(self, T, radius: float, return_indices: bool = False):
        center = np.ascontiguousarray(np.asarray(T, np.float32)[:3, 3])
        n = len(self)
        out = np.empty((max(1, n), 4), np.float32)
        idx = np.empty(max(1, n), np.int32)
        n_out = ctypes.c_size_t(0)
        capi.check(capi.lib().ssf_map_crop_radius(self._h, center.ctypes.data, float(radius), out.ctypes.data, n,
                                                  ctypes.byref(n_out), idx.ctypes.data))
        pts = out[:n_out.value, :3].copy()
        return (pts, idx[:n_out.value].copy()) if return_indices else pts

    def crop_to_target(self, icp: "ICPPointToPoint", T, radius: float) -> int:
        """``cropPointCloudThroughRadius`` + ``setTargetPointCloud`` without leaving HBM; returns the target size."""
        center = np.ascontiguousarray(np.asarray(T, np.float32)[:3, 3])
        n_out = ctypes.c_size_t(0)
        capi.check(capi.lib().ssf_map_crop_to_target(self._h, icp._h, center.ctypes.data, float(radius), ctypes.byref(n_out)))
        return int(n_out.value)

    def close(self) -> None:
        if getattr(self, "_h", None):
            capi.lib().ssf_map_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- BruteForceAlignment (reference localization/include/localization/brute_force_alignment.h) ----
class BruteForceAlignment:
    """Same setters / alignClouds / getters as the reference class (brute_force_alignment.cpp), with
    the pose grid scored on the GPU.  Parameter defaults are unset like in the reference: call the
    setters (the node's values are at localization_node.cpp:38-43)."""

    def __init__(self, context: Context | None = None):
        self._p = capi.BfaParams(0.1, 0.1, 0.05, 1.5, 1.5, 0.1, float(np.float32(np.pi) / np.float32(18.0)),
                                 float(np.float32(np.pi) / np.float32(6.0)), 0.1)
        self._icp = ICPPointToPoint(1.0, 1, 0.0, 0.0, context=context)  # holds the target's map index
        self._prev = np.eye(4, dtype=np.float32)
        self._best = np.eye(4, dtype=np.float32)
        self._done = False
        self._src = None
        self.last_scores = None

    def setXYZStep(self, x, y, z):
        self._p.x_step, self._p.y_step, self._p.z_step = x, y, z

    def setXYZRange(self, x, y, z):
        self._p.x_range, self._p.y_range, self._p.z_range = x, y, z

    def setRotationStep(self, yaw):
        self._p.yaw_step = yaw

    def setRotationRange(self, yaw):
        self._p.yaw_range = yaw

    def setMeanErrorThreshold(self, v):
        self._p.mean_error_threshold = v

    def setInitialGuess(self, T):
        # only the first guess is taken (brute_force_alignment.cpp:44-51: trace() == 4.0f test)
        if float(np.trace(self._prev)) == 4.0:
            self._prev = np.asarray(T, np.float32).copy()

    def setSourceCloud(self, cloud):
        self._src = _cloud(cloud)

    def setTargetCloud(self, cloud):
        self._icp.setTargetPointCloud(cloud)

    def resetFirstAlignment(self, value: bool):
        self._done = bool(value)

    def firstAlignmentCompleted(self) -> bool:
        return self._done

    def getBestTransformation(self) -> np.ndarray:
        return (self._best if self._done else self._prev).copy()

    def alignClouds(self) -> bool:
        n_pose = capi.lib().ssf_bfa_pose_count(ctypes.byref(self._p))
        Tp = _colmajor(self._prev)
        Tb = np.empty(16, np.float32)
        score, ok = ctypes.c_float(0), ctypes.c_int(0)
        scores = np.empty(max(1, n_pose), np.float32)
        capi.check(capi.lib().ssf_bfa_align(self._icp._h, self._src.ctypes.data, self._src.shape[0],
                                            self._src.strides[0], Tp.ctypes.data, ctypes.byref(self._p), Tb.ctypes.data,
                                            ctypes.addressof(score), ctypes.addressof(ok), scores.ctypes.data))
        self.last_scores, self.best_score = scores[:n_pose], float(score.value)
        T = Tb.reshape(4, 4).T.copy()
        if ok.value:  # cpp:114-119
            self._best, self._done = T, True
            return True
        self._prev = T  # cpp:126: the best candidate is the next starting pose
        return False
