"""ctypes declarations of include/ssf/ssf.h.  No torch, no CPU fallback: a missing or
unloadable libssf_gpu.so raises, and so does every non-zero status."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSF_GPU_LIB") or os.path.join(os.path.dirname(_HERE), "csrc", "libssf_gpu.so")

SSF_OK = 0
MODE_REFERENCE, MODE_GN_P2P, MODE_GN_P2PLANE, MODE_O3D_P2P = 0, 1, 2, 3
REDUCE_STRICT, REDUCE_FAST = 0, 1
XCH_HANDLE_BYTES = 128  # SSF_XCH_HANDLE_BYTES
NCCL_ID_BYTES = 128  # SSF_NCCL_ID_BYTES

EXPORTS = [
    "ssf_last_error", "ssf_version", "ssf_ctx_create", "ssf_ctx_destroy", "ssf_ctx_synchronize", "ssf_ctx_stream", "ssf_ctx_time_searches", "ssf_ctx_search_time", "ssf_ctx_search_times",
    "ssf_icp_create", "ssf_icp_destroy", "ssf_icp_set_params", "ssf_icp_get_params", "ssf_icp_set_target",
    "ssf_icp_set_source", "ssf_icp_set_initial", "ssf_icp_align", "ssf_icp_get_correspondences", "ssf_icp_get_trace",
    "ssf_icp_target_size", "ssf_icp_set_target_shard", "ssf_icp_set_allreduce", "ssf_icp_exchange_create", "ssf_icp_exchange_open", "ssf_icp_exchange_close", "ssf_nccl_unique_id", "ssf_icp_nccl_init", "ssf_icp_nccl_close", "ssf_nn_search", "ssf_nn_search_bench", "ssf_voxel_downsample", "ssf_voxel_downsample_o3d", "ssf_pcd_read", "ssf_pcd_write_binary", "ssf_cloud_from_pointcloud2", "ssf_map_create", "ssf_map_from_pcd_folder", "ssf_map_destroy", "ssf_map_size", "ssf_map_ingest_ms", "ssf_map_merge_ms", "ssf_map_download", "ssf_map_subsample", "ssf_map_crop_radius", "ssf_map_crop_to_target", "ssf_cloud_subsample", "ssf_cloud_remove_floor", "ssf_cloud_crop_radius", "ssf_bfa_pose_count", "ssf_bfa_align", "ssf_batch_create", "ssf_batch_destroy",
    "ssf_batch_upload", "ssf_batch_upload_async", "ssf_batch_set_initial", "ssf_batch_run", "ssf_batch_results", "ssf_batch_search_stats", "ssf_icp_align_batch",
    "ssf_kernel_launches", "ssf_nn_queries",
]


class SsfError(RuntimeError):
    pass


class IcpParams(ctypes.Structure):
    _fields_ = [("max_correspondence_dist", ctypes.c_float), ("num_iterations", ctypes.c_int32),
                ("acceptable_mean_error", ctypes.c_float), ("transformation_epsilon", ctypes.c_float),
                ("mode", ctypes.c_int32), ("reduce", ctypes.c_int32), ("debug", ctypes.c_int32),
                ("source_voxel_leaf", ctypes.c_float)]


class BfaParams(ctypes.Structure):
    _fields_ = [("x_step", ctypes.c_float), ("y_step", ctypes.c_float), ("z_step", ctypes.c_float),
                ("x_range", ctypes.c_float), ("y_range", ctypes.c_float), ("z_range", ctypes.c_float),
                ("yaw_step", ctypes.c_float), ("yaw_range", ctypes.c_float), ("mean_error_threshold", ctypes.c_float)]


class IcpResult(ctypes.Structure):
    _fields_ = [("transformation", ctypes.c_float * 16), ("error", ctypes.c_float), ("iterations", ctypes.c_int32),
                ("has_converged", ctypes.c_int32), ("n_searches", ctypes.c_int32), ("k_final", ctypes.c_int32),
                ("aborted", ctypes.c_int32), ("fitness", ctypes.c_float), ("n_source", ctypes.c_int32),
                ("device_ms", ctypes.c_float)]


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SsfError(f"{LIB_PATH} is missing: build it with `make -C slam-sensor-fusion_b200` "
                       "(or __graft_entry__.build()); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, i32, f32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float
    P = ctypes.POINTER
    L.ssf_last_error.restype = ctypes.c_char_p
    L.ssf_version.restype = ctypes.c_char_p
    L.ssf_ctx_create.argtypes = [i32, P(vp)]
    L.ssf_ctx_destroy.argtypes = [vp]
    L.ssf_ctx_destroy.restype = None
    L.ssf_ctx_synchronize.argtypes = [vp]
    L.ssf_ctx_stream.argtypes = [vp]
    L.ssf_ctx_stream.restype = vp
    L.ssf_ctx_time_searches.argtypes = [vp, i32]
    L.ssf_ctx_search_time.argtypes = [vp, P(ctypes.c_double), P(ctypes.c_uint64)]
    L.ssf_ctx_search_times.argtypes = [vp, P(ctypes.c_float), ctypes.c_uint64, P(ctypes.c_uint64)]
    L.ssf_icp_create.argtypes = [vp, P(IcpParams), P(vp)]
    L.ssf_icp_destroy.argtypes = [vp]
    L.ssf_icp_destroy.restype = None
    L.ssf_icp_set_params.argtypes = [vp, P(IcpParams)]
    L.ssf_icp_get_params.argtypes = [vp, P(IcpParams)]
    L.ssf_icp_set_target.argtypes = [vp, vp, sz, sz, vp, sz]
    L.ssf_icp_set_source.argtypes = [vp, vp, sz, sz]
    L.ssf_icp_set_initial.argtypes = [vp, vp]
    L.ssf_icp_align.argtypes = [vp, P(IcpResult)]
    L.ssf_icp_get_correspondences.argtypes = [vp, vp, sz]
    L.ssf_icp_get_trace.argtypes = [vp, vp, vp, sz]
    L.ssf_icp_target_size.argtypes = [vp]
    L.ssf_icp_target_size.restype = sz
    L.ssf_icp_set_target_shard.argtypes = [vp, vp, sz, sz, vp, sz, vp, vp]
    L.ssf_icp_set_allreduce.argtypes = [vp, vp, vp]
    L.ssf_icp_exchange_create.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, vp]
    L.ssf_icp_exchange_open.argtypes = [vp, vp]
    L.ssf_icp_exchange_close.argtypes = [vp]
    L.ssf_nccl_unique_id.argtypes = [vp]
    L.ssf_icp_nccl_init.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int]
    L.ssf_icp_nccl_close.argtypes = [vp]
    L.ssf_nn_search.argtypes = [vp, vp, sz, sz, f32, vp, vp]
    L.ssf_nn_search_bench.argtypes = [vp, vp, sz, sz, f32, i32, P(f32), vp, vp]
    L.ssf_voxel_downsample.argtypes = [vp, vp, sz, sz, f32, vp, P(sz), P(i32)]
    L.ssf_voxel_downsample_o3d.argtypes = [vp, vp, sz, sz, ctypes.c_double, vp, P(sz)]
    L.ssf_pcd_read.argtypes = [ctypes.c_char_p, vp, sz, P(sz)]
    L.ssf_pcd_write_binary.argtypes = [ctypes.c_char_p, vp, sz, sz]
    L.ssf_cloud_from_pointcloud2.argtypes = [vp, vp, sz, sz, sz, sz, sz, i32, vp]
    L.ssf_map_create.argtypes = [vp, vp, sz, sz, P(vp)]
    L.ssf_map_from_pcd_folder.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, f32, i32, P(vp)]
    L.ssf_map_destroy.argtypes = [vp]
    L.ssf_map_destroy.restype = None
    L.ssf_map_size.argtypes = [vp]
    L.ssf_map_size.restype = sz
    L.ssf_map_ingest_ms.argtypes = [vp]
    L.ssf_map_ingest_ms.restype = ctypes.c_double
    L.ssf_map_merge_ms.argtypes = [vp]
    L.ssf_map_merge_ms.restype = ctypes.c_double
    L.ssf_map_download.argtypes = [vp, vp, sz]
    L.ssf_map_subsample.argtypes = [vp, sz]
    L.ssf_map_crop_radius.argtypes = [vp, vp, ctypes.c_double, vp, sz, P(sz), vp]
    L.ssf_map_crop_to_target.argtypes = [vp, vp, vp, ctypes.c_double, P(sz)]
    L.ssf_cloud_subsample.argtypes = [vp, vp, sz, sz, sz, vp, P(sz)]
    L.ssf_cloud_remove_floor.argtypes = [vp, vp, sz, sz, vp, P(sz)]
    L.ssf_cloud_crop_radius.argtypes = [vp, vp, sz, sz, vp, ctypes.c_double, vp, P(sz), vp]
    L.ssf_bfa_pose_count.argtypes = [vp]
    L.ssf_bfa_pose_count.restype = sz
    L.ssf_bfa_align.argtypes = [vp, vp, sz, sz, vp, vp, vp, vp, vp, vp]
    L.ssf_batch_create.argtypes = [vp, sz, sz, P(vp)]
    L.ssf_batch_destroy.argtypes = [vp]
    L.ssf_batch_destroy.restype = None
    L.ssf_batch_upload.argtypes = [vp, vp, vp, sz, sz]
    L.ssf_batch_upload_async.argtypes = [vp, vp, vp, sz, sz]
    L.ssf_batch_set_initial.argtypes = [vp, vp]
    L.ssf_batch_run.argtypes = [vp]
    L.ssf_batch_results.argtypes = [vp, vp, sz]
    L.ssf_batch_search_stats.argtypes = [vp, vp, vp, sz, P(sz)]
    L.ssf_icp_align_batch.argtypes = [vp, vp, vp, sz, sz, vp, vp]
    L.ssf_kernel_launches.restype = ctypes.c_uint64
    L.ssf_nn_queries.restype = ctypes.c_uint64
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != SSF_OK:
        raise SsfError(f"libssf_gpu status {rc}: {lib().ssf_last_error().decode(errors='replace')}")
