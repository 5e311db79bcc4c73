"""Map ingestion formats on either side of the registration path.

* binary / ascii PCD tiles as written by the recorder with ``pcl::io::savePCDFileBinary``
  (``mapping/src/map_data_save_node.cpp:71-80``) and read back by ``pcl::io::loadPCDFile``
  (``localization/src/global_map_frames_manager.cpp:101,129``);
* ``merge_scans_and_save`` = ``GlobalMapFramesManager::mergeScansAndSave`` (``:110-151``): every
  ``*.pcd`` of the folder in directory order, concatenated, ``pcl::VoxelGrid(voxel)`` -- run on
  the GPU through ``ssf_voxel_downsample`` -- saved as ``<map_name>.pcd``;
* ``get_map_cloud`` = ``getMapCloud`` (``:93-108``): load the cached map if it exists, else merge.

File parsing is host-side I/O (native, csrc/ingest.cu); the geometry runs on the device.
"""
from __future__ import annotations

import os

import numpy as np



def read_pcd(path: str) -> np.ndarray:
    """x, y, z of a PCD file (DATA binary or ascii; float32 / float64 fields among any others) as an (N, 3)
    float32 array -- the library's native reader (``ssf_pcd_read``, csrc/ingest.cu)."""
    import ctypes
    from . import capi
    n = ctypes.c_size_t(0)
    capi.check(capi.lib().ssf_pcd_read(path.encode(), None, 0, ctypes.byref(n)))
    out = np.empty((max(1, n.value), 3), np.float32)
    capi.check(capi.lib().ssf_pcd_read(path.encode(), out.ctypes.data, out.shape[0], ctypes.byref(n)))
    return out[:n.value].copy()


def write_pcd_binary(path: str, xyz) -> None:
    """Same layout pcl::io::savePCDFileBinary gives a PointXYZ cloud: FIELDS x y z, 12 bytes per point."""
    from . import capi
    a = np.ascontiguousarray(np.asarray(xyz, np.float32)[:, :3])
    capi.check(capi.lib().ssf_pcd_write_binary(path.encode(), a.ctypes.data, a.shape[0], 12))


def get_map_cloud(data_folder: str, map_name: str, voxel_size: float, context=None) -> np.ndarray:
    """getMapCloud (global_map_frames_manager.cpp:93-151): the cached ``<map_name>.pcd`` if present (no voxel
    filter on that branch), else every ``*.pcd`` of the folder merged, pcl::VoxelGrid(voxel_size) on the device
    and saved.  Tiles go pinned host -> HBM; see ``ssf_gpu.ResidentMap.from_pcd_folder`` to keep the result there."""
    from . import ResidentMap
    m = ResidentMap.from_pcd_folder(data_folder, map_name, voxel_size, save=True, context=context)
    out = m.download()
    m.close()
    return out


def merge_scans_and_save(data_folder: str, map_name: str, voxel_size: float, context=None) -> np.ndarray:
    """mergeScansAndSave (global_map_frames_manager.cpp:110-151): merges even when a cached map exists."""
    cached = os.path.join(data_folder, map_name + ".pcd")
    if os.path.exists(cached):
        os.remove(cached)
    return get_map_cloud(data_folder, map_name, voxel_size, context)
