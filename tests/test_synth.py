"""The synthetic world generator is deterministic and has the shapes SURVEY.md 8d names."""
import numpy as np

from ssf_gpu import synth


def test_map_is_deterministic_and_exact_size():
    a, na, half = synth.make_map(20000, normals=True)
    b, nb, half2 = synth.make_map(20000, normals=True)
    assert a.shape == (20000, 4) and half == half2
    assert np.array_equal(a, b) and np.array_equal(na, nb)
    assert np.allclose(np.linalg.norm(na[:, :3], axis=1), 1.0, atol=1e-5)
    assert (a[:, 3] == 1.0).all() and np.abs(a[:, :2]).max() < half + 0.2


def test_scan_shapes_and_range():
    _, _, half = synth.make_map(20000)
    T = synth.street_pose(0, half=half)
    s1 = synth.make_scan(T, 8, 128, scan_id=1)
    s2 = synth.make_scan(T, 8, 128, scan_id=1)
    assert np.array_equal(s1, s2) and 0 < s1.shape[0] <= 8 * 128
    assert np.linalg.norm(s1[:, :3], axis=1).max() <= 100.5
    # the scan really is a view of the same world the map samples: most hits lie near map points
    from scipy.spatial import cKDTree
    m, _, half = synth.make_map(200000)
    T = synth.street_pose(0, half=half)
    s = synth.make_scan(T, 16, 256, scan_id=2, max_range=30.0)
    pw = s[:, :3] @ T[:3, :3].T + T[:3, 3]
    inside = (np.abs(pw[:, 0]) < half - 1) & (np.abs(pw[:, 1]) < half - 1)
    d, _ = cKDTree(m[:, :3]).query(pw[inside])
    assert np.median(d) < 0.15


def test_perturbation_ranges():
    T = synth.street_pose(10)
    for k in range(20):
        d = np.linalg.inv(T) @ synth.perturb_pose(T, k)
        assert abs(d[0, 3]) <= 0.3 and abs(d[1, 3]) <= 0.3 and abs(d[2, 3]) <= 0.05
        assert abs(np.degrees(np.arctan2(d[1, 0], d[0, 0]))) <= 2.0


def test_known_half_extents_match_the_search():
    """The cached half extents (bench sizes) are what the counting search returns."""
    from ssf_gpu import synth
    for m in (1_000_000, 5_000_000):
        cached = synth.map_half_extent(m)
        known = dict(synth._KNOWN_HALF)
        try:
            synth._KNOWN_HALF.clear()
            searched = synth.map_half_extent(m)
        finally:
            synth._KNOWN_HALF.update(known)
        assert cached == searched
