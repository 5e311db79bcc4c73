"""Synthetic street world: seeded maps, LiDAR scans and poses.

ctypes front end of ``synth/ssf_synth.c``.  This is workload generation for the tests
and ``bench.py``; it is neither the registration path nor the oracle.  The reference
ships no sample data (``localization/src/localization_node.cpp:7`` expects clouds under
``$HOME/Desktop/map_data``), hence the procedural world.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(os.path.dirname(_HERE), "synth", "libssf_synth.so")

SEED_WORLD = 0xB2000001
SEED_MAP = 0xB2000100
SEED_SCAN = 0xB2001000
SEED_POSE = 0xB2002000

MAP_SPACING = 0.17       # mimics voxel 0.1 + stride 3 (localization_node.cpp:19-20)
MAP_SIGMA = 0.01
RANGE_SIGMA = 0.02
SENSOR_HEIGHT = 1.8

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} missing: run `make -C slam-sensor-fusion_b200` (or __graft_entry__.build())")
        L = ctypes.CDLL(_LIB_PATH)
        L.ssf_synth_map_count.restype = ctypes.c_int64
        L.ssf_synth_map_count.argtypes = [ctypes.c_uint64, ctypes.c_double, ctypes.c_double]
        L.ssf_synth_map_fill.restype = ctypes.c_int64
        L.ssf_synth_map_fill.argtypes = [ctypes.c_uint64, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                         ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        L.ssf_synth_scan.restype = ctypes.c_int64
        L.ssf_synth_scan.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_void_p]
        L.ssf_synth_ground.restype = ctypes.c_double
        L.ssf_synth_ground.argtypes = [ctypes.c_double, ctypes.c_double]
        _lib = L
    return _lib


# map_half_extent() of the sizes the bench uses (default seed and spacing): the search costs a dozen counting
# passes over the candidate square, which matters at hundreds of millions of points
_KNOWN_HALF = {1_000_000: 71.0, 5_000_000: 156.0, 50_000_000: 483.0, 125_000_000: 774.0, 250_000_000: 1092.0,
               500_000_000: 1550.0}


def set_threads(n: int) -> None:
    """OpenMP threads of the generator (torchrun starts its workers with OMP_NUM_THREADS=1)."""
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(max(1, n)))
    except OSError:
        pass


def map_half_extent(m_points: int, spacing: float = MAP_SPACING, seed: int = SEED_WORLD) -> tuple[float, int]:
    """Smallest half-extent (multiple of 1 m) whose candidate count reaches ``m_points``."""
    L = lib()
    if spacing == MAP_SPACING and seed == SEED_WORLD and m_points in _KNOWN_HALF:
        half = _KNOWN_HALF[m_points]
        return half, int(L.ssf_synth_map_count(seed, half, spacing))
    lo, hi = 1.0, 8.0
    while L.ssf_synth_map_count(seed, hi, spacing) < m_points:
        lo, hi = hi, hi * 2.0
    while hi - lo > 1.0:
        mid = math.floor((lo + hi) / 2.0)
        if mid <= lo:
            break
        if L.ssf_synth_map_count(seed, mid, spacing) >= m_points:
            hi = mid
        else:
            lo = mid
    return hi, int(L.ssf_synth_map_count(seed, hi, spacing))


def make_map(m_points: int, normals: bool = False, spacing: float = MAP_SPACING, sigma: float = MAP_SIGMA,
             seed: int = SEED_WORLD, half: float | None = None):
    """Exactly ``m_points`` surface samples of the world as an (M, 4) float32 array (w = 1).

    Returns ``(xyz, normals_or_None, half_extent)``.
    """
    L = lib()
    if half is None:
        half, n_cand = map_half_extent(m_points, spacing, seed)
    else:
        n_cand = int(L.ssf_synth_map_count(seed, half, spacing))
    if n_cand < m_points:
        raise ValueError(f"half extent {half} holds only {n_cand} candidates < {m_points}")
    xyz = np.empty((m_points, 4), dtype=np.float32)
    nrm = np.empty((m_points, 4), dtype=np.float32) if normals else None
    r = L.ssf_synth_map_fill(seed, half, spacing, sigma, n_cand, m_points, xyz.ctypes.data,
                             nrm.ctypes.data if normals else None)
    if r != m_points:
        raise RuntimeError(f"ssf_synth_map_fill failed: {r}")
    return xyz, nrm, half


def make_scan(T_map_sensor: np.ndarray, beams: int = 32, azimuths: int = 1024, scan_id: int = 0,
              max_range: float = 100.0, fov=(-25.0, 15.0), range_sigma: float = RANGE_SIGMA,
              seed: int = SEED_WORLD) -> np.ndarray:
    """Ray-cast one scan from pose ``T_map_sensor`` (4x4); points in the sensor frame, (N, 4) float32."""
    T = np.ascontiguousarray(T_map_sensor, dtype=np.float64)
    out = np.empty((beams * azimuths, 4), dtype=np.float32)
    n = lib().ssf_synth_scan(seed, SEED_SCAN + scan_id, T.ctypes.data, beams, azimuths, fov[0], fov[1], max_range,
                             range_sigma, out.ctypes.data)
    if n < 0:
        raise RuntimeError("ssf_synth_scan failed")
    return out[:n].copy()


def rot_z(yaw: float) -> np.ndarray:
    c, s = math.cos(yaw), math.sin(yaw)
    R = np.eye(4)
    R[0, 0], R[0, 1], R[1, 0], R[1, 1] = c, -s, s, c
    return R


def street_pose(k: int, spacing: float = 0.15, half: float = 60.0) -> np.ndarray:
    """Ground-truth pose k of a route that runs back and forth along the y = 0 street
    centre-line inside [-half, half]^2, 0.1-0.2 m between poses as at reference
    stochastic_filter.cpp:11-12.  Sensor 1.8 m above the ground, heading along the route
    with a small yaw wiggle and a small lateral weave."""
    a = max(2.0, half - 15.0)
    per = 4.0 * a
    s = (k * spacing) % per
    if s < 2.0 * a:
        x, yaw = -a + s, 0.0
    else:
        x, yaw = a - (s - 2.0 * a), math.pi
    y = 1.5 * math.sin(0.05 * k * spacing)
    yaw += 0.03 * math.sin(0.37 * k)
    T = rot_z(yaw)
    T[0, 3], T[1, 3] = x, y
    T[2, 3] = SENSOR_HEIGHT + lib().ssf_synth_ground(x, y)
    return T


def perturb_pose(T_gt: np.ndarray, scan_id: int = 0, xy: float = 0.3, z: float = 0.05,
                 yaw_deg: float = 2.0) -> np.ndarray:
    """Initial guess = T_gt o delta, delta uniform in the SURVEY 8d ranges, seeded per scan."""
    rng = np.random.default_rng(SEED_POSE + scan_id)
    d = rot_z(math.radians(rng.uniform(-yaw_deg, yaw_deg)))
    d[0, 3], d[1, 3], d[2, 3] = rng.uniform(-xy, xy), rng.uniform(-xy, xy), rng.uniform(-z, z)
    return T_gt @ d
