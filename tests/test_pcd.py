"""Row N3: PCD tiles in, merged + voxel-filtered map out."""
import os

import numpy as np
import pytest

from ssf_gpu import pcd


def test_pcd_binary_roundtrip_and_ascii(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.normal(size=(1000, 3)).astype(np.float32)
    p = str(tmp_path / "cloud_10.pcd")
    pcd.write_pcd_binary(p, a)
    assert os.path.getsize(p) == len(open(p, "rb").read().split(b"DATA binary\n")[0]) + len(b"DATA binary\n") + 12000
    assert np.array_equal(pcd.read_pcd(p), a)
    # a tile with extra fields (e.g. intensity), ascii
    q = tmp_path / "ascii.pcd"
    q.write_text("# .PCD v0.7\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
                 "WIDTH 2\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 2\nDATA ascii\n1 2 3 9\n4 5 6 9\n")
    assert np.array_equal(pcd.read_pcd(str(q)), np.array([[1, 2, 3], [4, 5, 6]], np.float32))
    # binary with an extra field keeps x y z
    dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("i", "<f4")])
    raw = np.zeros(3, dt)
    raw["x"], raw["y"], raw["z"] = [1, 2, 3], [4, 5, 6], [7, 8, 9]
    r = tmp_path / "xyzi.pcd"
    r.write_bytes(b"VERSION 0.7\nFIELDS x y z i\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 3\nHEIGHT 1\n"
                  b"POINTS 3\nDATA binary\n" + raw.tobytes())
    assert np.array_equal(pcd.read_pcd(str(r)), np.array([[1, 4, 7], [2, 5, 8], [3, 6, 9]], np.float32))


@pytest.mark.gpu
def test_merge_scans_and_save_matches_oracle(tmp_path, small_world):
    from oracle import oracle
    m = small_world["map"][:, :3]
    parts = np.array_split(m[:40000], 4)
    for k, part in enumerate(parts):
        pcd.write_pcd_binary(str(tmp_path / f"cloud_{10 * (k + 1)}.pcd"), part)
    out = pcd.get_map_cloud(str(tmp_path), "map", 0.1)
    order = [e.name for e in os.scandir(tmp_path) if e.name.startswith("cloud_")]
    merged = np.concatenate([pcd.read_pcd(str(tmp_path / n)) for n in order])
    ref, refused = oracle.voxel_grid(merged, 0.1)
    assert not refused and np.array_equal(out.view(np.uint32), ref[:, :3].copy().view(np.uint32))
    # second call takes the cached map.pcd branch (no filter): identical cloud back
    assert os.path.exists(tmp_path / "map.pcd")
    assert np.array_equal(pcd.get_map_cloud(str(tmp_path), "map", 0.1), out)
