"""Map ingestion formats on either side of the registration path.

* binary / ascii PCD tiles as written by the recorder with ``pcl::io::savePCDFileBinary``
  (``mapping/src/map_data_save_node.cpp:71-80``) and read back by ``pcl::io::loadPCDFile``
  (``localization/src/global_map_frames_manager.cpp:101,129``);
* ``merge_scans_and_save`` = ``GlobalMapFramesManager::mergeScansAndSave`` (``:110-151``): every
  ``*.pcd`` of the folder in directory order, concatenated, ``pcl::VoxelGrid(voxel)`` -- run on
  the GPU through ``ssf_voxel_downsample`` -- saved as ``<map_name>.pcd``;
* ``get_map_cloud`` = ``getMapCloud`` (``:93-108``): load the cached map if it exists, else merge.

File parsing is host-side I/O (numpy); the geometry runs on the device.
"""
from __future__ import annotations

import os

import numpy as np

_SIZES = {("F", 4): np.float32, ("F", 8): np.float64, ("U", 1): np.uint8, ("U", 2): np.uint16, ("U", 4): np.uint32,
          ("I", 1): np.int8, ("I", 2): np.int16, ("I", 4): np.int32}


def read_pcd(path: str) -> np.ndarray:
    """x, y, z of a PCD file (DATA binary or ascii) as an (N, 3) float32 array."""
    with open(path, "rb") as f:
        header = {}
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: no DATA line")
            text = line.decode("ascii", errors="replace").strip()
            if not text or text.startswith("#"):
                continue
            key, _, val = text.partition(" ")
            header[key.upper()] = val.split()
            if key.upper() == "DATA":
                break
        fields = header["FIELDS"]
        sizes = [int(v) for v in header["SIZE"]]
        types = header["TYPE"]
        counts = [int(v) for v in header.get("COUNT", ["1"] * len(fields))]
        n = int(header["POINTS"][0]) if "POINTS" in header else int(header["WIDTH"][0]) * int(header["HEIGHT"][0])
        kind = header["DATA"][0].lower()
        if kind == "binary":
            dt = np.dtype([(name, _SIZES[(t, s)], (c,)) if c > 1 else (name, _SIZES[(t, s)])
                           for name, s, t, c in zip(fields, sizes, types, counts)])
            raw = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
            return np.stack([raw["x"], raw["y"], raw["z"]], axis=1).astype(np.float32)
        if kind == "ascii":
            cols = np.loadtxt(f, dtype=np.float64, ndmin=2)
            offs = np.cumsum([0] + counts)
            ix, iy, iz = (int(offs[fields.index(k)]) for k in ("x", "y", "z"))
            return cols[:n, [ix, iy, iz]].astype(np.float32)
        raise ValueError(f"{path}: DATA {kind} is not supported (the recorder writes binary)")


def write_pcd_binary(path: str, xyz) -> None:
    """Same layout pcl::io::savePCDFileBinary gives a PointXYZ cloud: FIELDS x y z, 12 bytes per point."""
    a = np.ascontiguousarray(np.asarray(xyz, np.float32)[:, :3])
    n = a.shape[0]
    head = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\n"
            f"COUNT 1 1 1\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA binary\n")
    with open(path, "wb") as f:
        f.write(head.encode("ascii"))
        f.write(a.tobytes())


def merge_scans_and_save(data_folder: str, map_name: str, voxel_size: float, context=None) -> np.ndarray:
    """mergeScansAndSave (global_map_frames_manager.cpp:110-151)."""
    from . import voxel_down_sample
    clouds = []
    with os.scandir(data_folder) as it:  # directory order, like readdir
        for ent in it:
            if len(ent.name) > 4 and ent.name.endswith(".pcd"):
                clouds.append(read_pcd(ent.path))
    merged = np.concatenate(clouds, axis=0) if clouds else np.zeros((0, 3), np.float32)
    out = voxel_down_sample(merged, voxel_size, context) if merged.shape[0] else merged
    write_pcd_binary(os.path.join(data_folder, map_name + ".pcd"), out)
    return out


def get_map_cloud(data_folder: str, map_name: str, voxel_size: float, context=None) -> np.ndarray:
    """getMapCloud (global_map_frames_manager.cpp:93-108): cached map if present (no voxel filter on
    that branch), else merge + filter + save."""
    path = os.path.join(data_folder, map_name + ".pcd")
    if os.path.exists(path):
        return read_pcd(path)
    return merge_scans_and_save(data_folder, map_name, voxel_size, context)
