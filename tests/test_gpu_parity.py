"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Bars (BASELINE.json north_star):
  - correspondence indices and squared distances bit-exact, ties -> lowest map index;
  - REFERENCE mode with the STRICT reduction: pose, error, iteration count, search schedule
    and every correspondence bit-identical to the oracle;
  - every other mode: pose within 1e-4 m / 1e-5 rad, iteration count +-1;
  - voxel-downsampled clouds bit-exact as point sets.
"""
import numpy as np
import pytest

from conftest import pose_delta

pytestmark = pytest.mark.gpu

TOL_T, TOL_R = 1e-4, 1e-5


@pytest.fixture(scope="module")
def gpu():
    import ssf_gpu
    return ssf_gpu


@pytest.fixture(scope="module")
def ora():
    from oracle import oracle
    return oracle


def _check_nn(gpu, ora, m, q, thr):
    icp = gpu.ICPPointToPoint(thr, 1, 0.0, 0.0)
    icp.setTargetPointCloud(m)
    gi, gd = icp.nearest(q, thr)
    oi, od = ora.nn_brute(m, q) if m.shape[0] * q.shape[0] <= 4e8 else ora.KdTree(m).nn(q, threads=8)
    inside = od < np.float32(thr)
    assert np.array_equal(gi[inside], oi[inside])
    assert np.array_equal(gd[inside].view(np.uint32), od[inside].view(np.uint32))
    assert (gi[~inside] == -1).all()
    return inside.sum()


def test_nn_random(gpu, ora):
    rng = np.random.default_rng(1)
    m = rng.uniform(-20, 20, (200_000, 3)).astype(np.float32)
    q = rng.uniform(-22, 22, (20_000, 3)).astype(np.float32)
    assert _check_nn(gpu, ora, m, q, 0.5) > 1000


def test_nn_lattice_ties_and_duplicates(gpu, ora):
    # lattice points duplicated (equal distances everywhere): lowest index must win
    g = np.stack(np.meshgrid(*[np.arange(12)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.25
    m = np.concatenate([g, g[::-1], g])
    q = np.concatenate([g + 0.125, g, g + np.float32(0.25 / 2)]).astype(np.float32)
    assert _check_nn(gpu, ora, m, q, 0.5) == q.shape[0]


def test_nn_cell_borders_and_threshold_edge(gpu, ora):
    rng = np.random.default_rng(2)
    h = np.float32(np.sqrt(np.float32(0.5)) * np.float32(1.01))
    m = (rng.integers(0, 40, (30_000, 3)).astype(np.float32) * h).astype(np.float32)  # points on cell borders
    q = (m[rng.integers(0, m.shape[0], 5000)] + rng.choice([-1, 0, 1], (5000, 3)).astype(np.float32) *
         np.float32(np.sqrt(0.5 / 3))).astype(np.float32)  # distances right at the threshold
    _check_nn(gpu, ora, m, q, 0.5)


def test_nn_empty_and_far(gpu, ora):
    icp = gpu.ICPPointToPoint(0.5, 1, 0.0, 0.0)
    icp.setTargetPointCloud(np.zeros((0, 3), np.float32))
    gi, gd = icp.nearest(np.zeros((5, 3), np.float32), 0.5)
    assert (gi == -1).all()
    icp.setTargetPointCloud(np.array([[0, 0, 0], [np.nan, 0, 0], [1, 1, 1]], np.float32))
    gi, gd = icp.nearest(np.array([[0.1, 0, 0], [100, 100, 100], [np.inf, 0, 0], [0.9, 1, 1]], np.float32), 0.5)
    assert gi.tolist() == [0, -1, -1, 2]


@pytest.mark.parametrize("cell", ["0.07", "0.3", "1.5", "6.0"])
@pytest.mark.parametrize("thr", [0.01, 0.5, 5.0, 400.0])
def test_nn_any_cell_edge_any_threshold(gpu, ora, cell, thr, monkeypatch):
    """The ring walk must be exact whatever the ratio of search radius to cell edge: from a radius far
    below one cell (0.1 m vs 6 m cells) to dozens of rings (20 m vs 7 cm cells), on a surface-like
    cloud with holes, duplicates and queries outside the map."""
    monkeypatch.setenv("SSF_CELL_SIZE", cell)
    rng = np.random.default_rng(int(float(cell) * 100) + int(thr * 10))
    xy = rng.uniform(-8, 8, (30_000, 2)).astype(np.float32)
    xy = xy[(np.abs(xy[:, 0]) > 1.0) | (np.abs(xy[:, 1]) > 1.5)]  # a hole in the floor
    floor = np.c_[xy, 0.02 * rng.standard_normal(len(xy))].astype(np.float32)
    wall = np.c_[rng.uniform(-8, 8, 8000), np.full(8000, 3.0), rng.uniform(0, 4, 8000)].astype(np.float32)
    m = np.concatenate([floor, wall, floor[:500]])  # + duplicates (lowest index must win)
    q = np.concatenate([floor[::7] + rng.normal(0, 0.05, floor[::7].shape).astype(np.float32),
                        rng.uniform(-12, 12, (3000, 3)).astype(np.float32),   # in and around the map
                        np.array([[0, 0, 0.3], [0, 0, 3], [100, 100, 100], [-9, 3.0, 2]], np.float32)])
    n_in = _check_nn(gpu, ora, m, q, thr)
    assert n_in > (1000 if thr >= 0.5 else 100)


def test_nn_degenerate_maps(gpu, ora):
    """Maps the cell-size search has to survive: one point, all points identical, a line, a plane, a cloud
    far from the origin (large coordinates, small spacing) and a millimetre-scale cloud."""
    rng = np.random.default_rng(11)
    line = np.c_[np.linspace(-5, 5, 4001), np.zeros(4001), np.zeros(4001)].astype(np.float32)
    plane = np.c_[rng.uniform(-4, 4, (20000, 2)), np.zeros(20000)].astype(np.float32)
    far = (rng.uniform(-3, 3, (30000, 3)) * np.array([1, 1, 0.02]) + np.array([40000.0, -25000.0, 300.0])).astype(np.float32)
    tiny = (rng.uniform(-1, 1, (20000, 3)) * 1e-3).astype(np.float32)
    cases = [
        (np.array([[1.0, 2.0, 3.0]], np.float32), rng.uniform(0, 4, (500, 3)).astype(np.float32), 0.5),
        (np.tile(np.array([[0.5, -0.25, 2.0]], np.float32), (3000, 1)), rng.uniform(-1, 3, (500, 3)).astype(np.float32), 0.5),
        (line, np.c_[rng.uniform(-6, 6, 2000), rng.normal(0, 0.2, 2000), rng.normal(0, 0.2, 2000)].astype(np.float32), 0.5),
        (plane, np.c_[rng.uniform(-5, 5, (3000, 2)), rng.normal(0, 0.3, 3000)].astype(np.float32), 0.5),
        (far, (far[::13] + rng.normal(0, 0.05, far[::13].shape)).astype(np.float32), 0.5),
        (tiny, (tiny[::7] + rng.normal(0, 5e-5, tiny[::7].shape)).astype(np.float32), 1e-7),
    ]
    for m, q, thr in cases:
        assert _check_nn(gpu, ora, m, q, thr) > 0


def test_nn_synthetic_map(gpu, ora, c1_world):
    w = c1_world
    T0 = w["T0"].astype(np.float32)
    q = (w["scan"][:, :3] @ T0[:3, :3].T + T0[:3, 3]).astype(np.float32)
    n_in = _check_nn(gpu, ora, w["map"], q, 0.5)
    assert n_in > 0.8 * q.shape[0]


def _ref_pair(gpu, ora, w, reduce, **kw):
    prm = dict(max_correspondence_dist=0.5, num_iterations=10, acceptable_mean_error=0.05,
               transformation_epsilon=1e-5)
    prm.update(kw)
    tree = ora.KdTree(w["map"])
    ores, ocorr, otr = ora.icp_reference(tree, w["scan"], w["T0"], trace=True, **prm)
    icp = gpu.ICPPointToPoint(prm["max_correspondence_dist"], prm["num_iterations"], prm["acceptable_mean_error"],
                              prm["transformation_epsilon"], mode=gpu.MODE_REFERENCE, reduce=reduce)
    icp.setTargetPointCloud(w["map"])
    icp.setSourcePointCloud(w["scan"])
    icp.setInitialTransformation(w["T0"])
    g = icp.calculateAlignment()
    return ores, ocorr, otr, g, icp


@pytest.mark.parametrize("world", ["small_world", "c1_world"])
def test_reference_strict_bit_exact(gpu, ora, world, request):
    w = request.getfixturevalue(world)
    ores, ocorr, otr, g, icp = _ref_pair(gpu, ora, w, gpu.REDUCE_STRICT)
    assert g.iterations == ores.iterations and g.n_searches == ores.n_searches
    assert g.k_final == ores.k_final and g.has_converged == bool(ores.has_converged)
    assert np.array_equal(g.transformation.view(np.uint32), ores.T.view(np.uint32))
    assert np.float32(g.error).view(np.uint32) == np.float32(ores.error).view(np.uint32)
    assert np.array_equal(icp.correspondences(), ocorr)
    gerr, gsrch = icp.trace()
    n = ores.iterations + (1 if ores.iterations < 10 else 0)
    assert np.array_equal(gerr[:n].view(np.uint32), otr.iter_err[:n].view(np.uint32))
    assert np.array_equal(gsrch, otr.iter_searched)
    # and the registration actually worked: closer to ground truth than the initial guess
    assert pose_delta(g.transformation, w["T_gt"])[0] < pose_delta(w["T0"], w["T_gt"])[0]


def test_reference_strict_coarse_parameters(gpu, ora, small_world):
    # the "strong" parameter set of localization_node.cpp:226-229 (thr 5.0 -> radius 2.24 m, 80 passes)
    ores, ocorr, otr, g, icp = _ref_pair(gpu, ora, small_world, gpu.REDUCE_STRICT, max_correspondence_dist=5.0,
                                         num_iterations=80, acceptable_mean_error=0.4, transformation_epsilon=1e-2)
    assert g.iterations == ores.iterations and g.n_searches == ores.n_searches
    assert np.array_equal(g.transformation.view(np.uint32), ores.T.view(np.uint32))
    assert np.array_equal(icp.correspondences(), ocorr)


def test_reference_abort_too_few(gpu, ora, small_world):
    w = dict(small_world)
    far = np.eye(4)
    far[:3, 3] = [5000.0, 5000.0, 0.0]
    w["T0"] = far
    ores, ocorr, otr, g, icp = _ref_pair(gpu, ora, w, gpu.REDUCE_STRICT)
    assert ores.aborted == 1 and g.aborted
    assert g.iterations == 0 and not g.has_converged and g.error == pytest.approx(1e6)
    assert np.array_equal(g.transformation, far.astype(np.float32))


def test_reference_fast_close(gpu, ora, small_world):
    ores, ocorr, otr, g, icp = _ref_pair(gpu, ora, small_world, gpu.REDUCE_FAST)
    assert abs(g.iterations - ores.iterations) <= 1
    dt, dr = pose_delta(g.transformation, ores.T)
    # FAST sums in double; the oracle's float chains carry ~1e-4 m of summation noise and
    # the re-search schedule depends on it (DESIGN.md "reduction orders"), hence the wider bar
    assert dt < 5e-3 and dr < 5e-4


@pytest.mark.parametrize("mode", ["p2p", "p2plane"])
@pytest.mark.parametrize("world", ["small_world", "c1_world"])
def test_gn_modes(gpu, ora, world, mode, request):
    w = request.getfixturevalue(world)
    tree = ora.KdTree(w["map"])
    ores, ocorr = ora.icp_gn(tree, w["scan"], w["T0"], mode=mode, normals=w["normals"], num_iterations=10)
    icp = gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, mode=gpu.MODE_GN_P2PLANE if mode == "p2plane" else gpu.MODE_GN_P2P)
    icp.setTargetPointCloud(w["map"], w["normals"])
    icp.setSourcePointCloud(w["scan"])
    icp.setInitialTransformation(w["T0"])
    g = icp.calculateAlignment()
    assert abs(g.iterations - ores.iterations) <= 1
    dt, dr = pose_delta(g.transformation, ores.T)
    assert dt < TOL_T and dr < TOL_R, (dt, dr)
    assert g.k_final == ores.k_final
    assert (icp.correspondences() == ocorr).mean() > 0.999
    assert g.error == pytest.approx(ores.error, rel=1e-4)
    if mode == "p2plane":
        assert pose_delta(g.transformation, w["T_gt"])[0] < 0.02


def test_o3d_flow(gpu, ora, small_world):
    w = small_world
    tree = ora.KdTree(w["map"])
    ores, ofit, ocorr = ora.icp_o3d(tree, w["scan"], w["T0"], 0.5, 30)
    r = gpu.registration_icp(w["scan"][:, :3].astype(np.float64), w["map"][:, :3], 0.5, w["T0"],
                             gpu.TransformationEstimationPointToPoint(), gpu.ICPConvergenceCriteria(max_iteration=30))
    assert abs(r.iterations - ores.iterations) <= 1
    dt, dr = pose_delta(r.transformation, ores.T)
    assert dt < TOL_T and dr < TOL_R, (dt, dr)
    assert r.fitness == pytest.approx(ofit, abs=1e-4)
    assert r.inlier_rmse == pytest.approx(ores.error, rel=1e-4)
    assert r.correspondence_set.shape[0] == ores.k_final


def test_o3d_flow_python_node_on_1M_map(gpu, ora, c1_world):
    """The Python node's own call (localization_python/localization_python/localization_node.py:233-237:
    registration_icp(scan, map, threshold, identity-composed prior, PointToPoint, max_iteration=30)) on config 1's
    1M-point map, after the node's map voxel_down_sample(0.1) in Open3D semantics (:47), against the oracle's
    Open3D-flow restatement."""
    w = c1_world
    m = gpu.voxel_down_sample(w["map"], 0.1, semantics="open3d")
    om = ora.voxel_grid_o3d(w["map"], 0.1)
    assert np.array_equal(m.view(np.uint32), om.view(np.uint32))
    tree = ora.KdTree(om)
    for thr in (0.5, 0.25):
        ores, ofit, ocorr = ora.icp_o3d(tree, w["scan"], w["T0"], thr, 30)
        r = gpu.registration_icp(w["scan"][:, :3].astype(np.float64), m, thr, w["T0"],
                                 gpu.TransformationEstimationPointToPoint(), gpu.ICPConvergenceCriteria(max_iteration=30))
        assert abs(r.iterations - ores.iterations) <= 1
        dt, dr = pose_delta(r.transformation, ores.T)
        assert dt < TOL_T and dr < TOL_R, (thr, dt, dr)
        assert r.fitness == pytest.approx(ofit, abs=1e-4)
        assert r.inlier_rmse == pytest.approx(ores.error, rel=1e-4)
        assert r.correspondence_set.shape[0] == ores.k_final


def test_voxel_grid_bit_exact(gpu, ora, small_world):
    rng = np.random.default_rng(5)
    clouds = [small_world["scan"], small_world["map"][:50_000],
              rng.uniform(-3, 3, (20_000, 3)).astype(np.float32),
              np.array([[0, 0, 0], [np.nan, 1, 1], [0.01, 0.01, 0.01], [5, 5, 5]], np.float32)]
    for c, leaf in zip(clouds, [0.2, 0.1, 0.05, 0.5]):
        o, refused = ora.voxel_grid(c, leaf)
        g = gpu.voxel_down_sample(c, leaf)
        assert not refused
        assert g.shape[0] == o.shape[0]
        # same order too (ascending voxel index), which implies the same point set
        assert np.array_equal(g.view(np.uint32), o[:, :3].copy().view(np.uint32))


def test_voxel_grid_overflow_refusal(gpu, ora):
    c = np.array([[0, 0, 0], [3000, 3000, 3000], [1, 1, 1]], np.float32)
    o, refused = ora.voxel_grid(c, 0.001)
    assert refused
    g = gpu.voxel_down_sample(c, 0.001)
    assert np.array_equal(g, c)


def test_batch_matches_single(gpu, ora, small_world):
    from ssf_gpu import synth
    w = small_world
    scans, inits = [], []
    for k in range(5):
        T = synth.street_pose(3 + 7 * k, half=w["half"])
        scans.append(synth.make_scan(T, beams=16, azimuths=256 + 64 * k, scan_id=50 + k, max_range=60.0))
        inits.append(synth.perturb_pose(T, 50 + k))
    scans.append(np.zeros((0, 4), np.float32))  # an empty scan aborts like the reference
    inits.append(np.eye(4))
    for mode, reduce in [(gpu.MODE_REFERENCE, gpu.REDUCE_STRICT), (gpu.MODE_GN_P2PLANE, gpu.REDUCE_FAST)]:
        icp = gpu.ICPPointToPoint(0.5, 10, 0.05 if mode == gpu.MODE_REFERENCE else 0.0,
                                  1e-5 if mode == gpu.MODE_REFERENCE else 0.0, mode=mode, reduce=reduce)
        icp.setTargetPointCloud(w["map"], w["normals"])
        batch = icp.align_batch(scans, inits)
        assert batch[-1].aborted
        for s, T0, rb in zip(scans[:-1], inits[:-1], batch[:-1]):
            icp.setSourcePointCloud(s)
            icp.setInitialTransformation(T0)
            r1 = icp.calculateAlignment()
            assert np.array_equal(r1.transformation.view(np.uint32), rb.transformation.view(np.uint32))
            assert (r1.iterations, r1.n_searches, r1.k_final) == (rb.iterations, rb.n_searches, rb.k_final)


@pytest.mark.parametrize("mode", ["p2p", "p2plane", "o3d"])
def test_certificates_change_nothing(gpu, c1_world, small_world, mode, monkeypatch):
    """Search certificates (a query that barely moved keeps its neighbour without a walk) are an
    exactness-preserving shortcut: with them switched off every pose, error, count and
    correspondence must be bit-identical."""
    from ssf_gpu import synth
    for w, iters in ((small_world, 10), (c1_world, 30)):
        scans, inits = [w["scan"]], [w["T0"]]
        for k in range(3):
            T = synth.street_pose(5 + 11 * k, half=w["half"])
            scans.append(synth.make_scan(T, beams=16, azimuths=512, scan_id=70 + k, max_range=60.0))
            inits.append(synth.perturb_pose(T, 70 + k))
        out = []
        for no_cert in ("0", "1"):
            monkeypatch.setenv("SSF_NO_CERT", no_cert)
            m = {"p2p": gpu.MODE_GN_P2P, "p2plane": gpu.MODE_GN_P2PLANE, "o3d": gpu.MODE_O3D_P2P}[mode]
            icp = gpu.ICPPointToPoint(0.5, iters, 0.0, 0.0, mode=m)
            icp.setTargetPointCloud(w["map"], w["normals"])
            res = icp.align_batch(scans, inits)
            icp.setSourcePointCloud(scans[0])
            icp.setInitialTransformation(inits[0])
            r1 = icp.calculateAlignment()
            out.append((res, r1, icp.correspondences().copy()))
        (ra, sa, ca), (rb, sb, cb) = out
        assert np.array_equal(ca, cb)
        for x, y in list(zip(ra, rb)) + [(sa, sb)]:
            assert np.array_equal(x.transformation.view(np.uint32), y.transformation.view(np.uint32))
            assert (x.iterations, x.n_searches, x.k_final, x.has_converged) == (y.iterations, y.n_searches, y.k_final,
                                                                                y.has_converged)
            assert np.float32(x.error).view(np.uint32) == np.float32(y.error).view(np.uint32)


@pytest.mark.parametrize("mode", ["p2p", "p2plane"])
def test_reach_mask_changes_nothing(gpu, ora, c1_world, mode, monkeypatch):
    """The reach mask (one bit per map cell: "no target point within the rejection radius of any
    position in this cell") only ends walks that could not have found anything.  The target is the
    10 m crop of the map around the pose -- the reference's own usage (localization_node.cpp:302) --
    so most of the scan lies in cleared cells; with the mask switched off (SSF_NO_REACH=1) every
    pose, error, count and correspondence must be bit-identical.  Also against the oracle's NN."""
    w = c1_world
    c = w["T_gt"][:3, 3].astype(np.float32)
    keep = np.linalg.norm(w["map"][:, :3] - c, axis=1) < 10.0
    tgt, nrm = np.ascontiguousarray(w["map"][keep]), np.ascontiguousarray(w["normals"][keep])
    assert 1000 < tgt.shape[0] < w["map"].shape[0] // 2
    out = []
    for off in ("0", "1"):
        monkeypatch.setenv("SSF_NO_REACH", off)
        m = {"p2p": gpu.MODE_GN_P2P, "p2plane": gpu.MODE_GN_P2PLANE}[mode]
        icp = gpu.ICPPointToPoint(0.5, 8, 0.0, 0.0, mode=m)
        icp.setTargetPointCloud(tgt, nrm)
        icp.setSourcePointCloud(w["scan"])
        icp.setInitialTransformation(w["T0"])
        r = icp.calculateAlignment()
        out.append((r, icp.correspondences().copy()))
        # the threshold changes: the mask is rebuilt for it (coarse parameter set of the node)
        icp.setMaxCorrespondenceDist(5.0)
        r2 = icp.calculateAlignment()
        out[-1] += (r2, icp.correspondences().copy())
    (ra, ca, ra2, ca2), (rb, cb, rb2, cb2) = out
    assert np.array_equal(ca, cb) and np.array_equal(ca2, cb2)
    assert (ca >= 0).sum() > 100 and (ca < 0).sum() > 100  # both kinds of query are present
    for x, y in ((ra, rb), (ra2, rb2)):
        assert np.array_equal(x.transformation.view(np.uint32), y.transformation.view(np.uint32))
        assert (x.iterations, x.n_searches, x.k_final) == (y.iterations, y.n_searches, y.k_final)
        assert np.float32(x.error).view(np.uint32) == np.float32(y.error).view(np.uint32)
    # the parity entry point walks the same index: with the mask in place (built by the run above
    # for thr 5.0, rebuilt here for 0.5) its answers must still be the oracle's exact NN
    monkeypatch.setenv("SSF_NO_REACH", "0")
    icp.setMaxCorrespondenceDist(0.5)
    icp.calculateAlignment()
    q = (w["scan"][:, :3].astype(np.float64) @ w["T0"][:3, :3].T + w["T0"][:3, 3]).astype(np.float32)
    for thr in (0.5, 0.05):
        gi, gd = icp.nearest(q, thr)
        oi, od = ora.KdTree(tgt).nn(q, threads=8)
        inside = od < np.float32(thr)
        assert np.array_equal(gi[inside], oi[inside])
        assert np.array_equal(gd[inside].view(np.uint32), od[inside].view(np.uint32))
        assert (gi[~inside] == -1).all() and inside.sum() > 100 and (~inside).sum() > 100


def test_certificates_with_ties(gpu, monkeypatch):
    """Lattice map with every point duplicated: nearest and second-nearest distances coincide all over, so
    certificates must refuse to confirm and the walk must break the ties by index -- same bits either way."""
    g = np.stack(np.meshgrid(np.arange(-20, 21), np.arange(-20, 21), indexing="ij"), -1).reshape(-1, 2).astype(np.float32) * 0.25
    floor = np.c_[g, np.zeros(len(g), np.float32)]
    wall = np.c_[g[:, 0], np.full(len(g), 5.0, np.float32), np.abs(g[:, 1])]
    m = np.concatenate([floor, wall, floor, wall[::-1]]).astype(np.float32)
    rng = np.random.default_rng(4)
    src = np.concatenate([floor[::3], wall[::3]]) + np.float32(0.125) * np.array([1, 1, 0], np.float32)  # on cell mid-points
    src = np.concatenate([src, src[:200] + rng.normal(0, 0.01, (200, 3)).astype(np.float32)]).astype(np.float32)
    c, s_ = np.cos(0.02), np.sin(0.02)
    T0 = np.array([[c, -s_, 0, 0.06], [s_, c, 0, -0.04], [0, 0, 1, 0.03], [0, 0, 0, 1]], np.float64)
    out = []
    for no_cert in ("0", "1"):
        monkeypatch.setenv("SSF_NO_CERT", no_cert)
        for mode in (gpu.MODE_GN_P2P, gpu.MODE_O3D_P2P):
            icp = gpu.ICPPointToPoint(0.5, 15, 0.0, 0.0, mode=mode)
            icp.setTargetPointCloud(m)
            icp.setSourcePointCloud(src)
            icp.setInitialTransformation(T0)
            r = icp.calculateAlignment()
            out.append((no_cert, mode, r, icp.correspondences().copy()))
    for (_, _, ra, ca), (_, _, rb, cb) in zip(out[:2], out[2:]):
        assert np.array_equal(ca, cb)
        assert np.array_equal(ra.transformation.view(np.uint32), rb.transformation.view(np.uint32))
        assert (ra.iterations, ra.k_final) == (rb.iterations, rb.k_final)
        # duplicates: the reported neighbour is always the copy with the lower index
        assert (ca[ca >= 0] < 2 * len(floor)).all()


def test_batched_voxel_stage_matches_oracle(gpu, ora, small_world):
    """Scans of different sizes (one empty, one with NaNs, one whose voxel grid PCL refuses) go through the
    segmented voxel stage in ONE batch; every scan must enter the loop with exactly the oracle's centroids."""
    from ssf_gpu import synth
    w = small_world
    tree = ora.KdTree(w["map"])
    scans, inits = [], []
    for k, az in enumerate((256, 1024, 700)):
        T = synth.street_pose(4 + 9 * k, half=w["half"])
        scans.append(synth.make_scan(T, beams=16, azimuths=az, scan_id=90 + k, max_range=60.0))
        inits.append(synth.perturb_pose(T, 90 + k))
    bad = scans[1].copy()
    bad[::97, 1] = np.nan  # dropped by the voxel grid
    scans.append(bad)
    inits.append(inits[1])
    scans.append(np.zeros((0, 4), np.float32))
    inits.append(np.eye(4))
    icp = gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, mode=gpu.MODE_GN_P2PLANE)
    icp.setTargetPointCloud(w["map"], w["normals"])
    icp.setSourceVoxelLeaf(0.2)
    res = icp.align_batch(scans, inits)
    assert res[-1].aborted and res[-1].n_source == 0
    for s, T0, r in zip(scans[:-1], inits[:-1], res[:-1]):
        vox, refused = ora.voxel_grid(s, 0.2)
        assert not refused and r.n_source == vox.shape[0]
        o, _ = ora.icp_gn(tree, vox, T0, mode="p2plane", normals=w["normals"], num_iterations=10)
        dt, dr = pose_delta(r.transformation, o.T)
        assert dt < TOL_T and dr < TOL_R and abs(r.iterations - o.iterations) <= 1 and r.k_final == o.k_final
    # the same scans one at a time: bit-identical to the batch
    for s, T0, r in zip(scans[:-1], inits[:-1], res[:-1]):
        icp.setSourcePointCloud(s)
        icp.setInitialTransformation(T0)
        r1 = icp.calculateAlignment()
        assert np.array_equal(r1.transformation.view(np.uint32), r.transformation.view(np.uint32))
        assert (r1.n_source, r1.k_final) == (r.n_source, r.k_final)


def test_golden_fixture(gpu):
    """CUDA path against the committed golden vectors (tests/golden/c1_mini.npz)."""
    import os
    from conftest import ROOT
    g = np.load(os.path.join(ROOT, "tests", "golden", "c1_mini.npz"))
    icp = gpu.ICPPointToPoint(0.5, 10, 0.05, 1e-5)
    icp.setTargetPointCloud(g["map"], g["normals"])
    idx, d2 = icp.nearest(g["queries"], 0.5)
    inside = g["nn_d2"] < np.float32(0.5)
    assert np.array_equal(idx[inside], g["nn_idx"][inside])
    assert np.array_equal(d2[inside].view(np.uint32), g["nn_d2"][inside].view(np.uint32))
    icp.setSourcePointCloud(g["scan"])
    icp.setInitialTransformation(g["T0"])
    r = icp.calculateAlignment()
    assert np.array_equal(r.transformation.view(np.uint32), g["ref_T"].view(np.uint32))
    assert (r.iterations, r.n_searches, r.k_final) == (int(g["ref_iterations"]), int(g["ref_n_searches"]),
                                                       int(g["ref_k_final"]))
    assert np.array_equal(icp.correspondences(), g["ref_corr"])
    icp.setMode(gpu.MODE_GN_P2PLANE)
    icp.setAcceptableMeanError(0.0)
    icp.setTransformationEpsilon(0.0)
    r = icp.calculateAlignment()
    dt, dr = pose_delta(r.transformation, g["gn_p2plane_T"])
    assert dt < TOL_T and dr < TOL_R and abs(r.iterations - int(g["gn_p2plane_iterations"])) <= 1
    v = gpu.voxel_down_sample(g["scan"], 0.2)
    assert np.array_equal(v.view(np.uint32), g["vox_02"][:, :3].copy().view(np.uint32))


def test_search_timer(gpu, small_world):
    """ssf_ctx_time_searches / ssf_ctx_search_times / ssf_ctx_search_time: one event pair per K3 launch,
    per-launch read-out in launch order, the total equals their sum and clears the list."""
    w = small_world
    ctx = gpu.default_context()
    icp = gpu.ICPPointToPoint(0.5, 6, 0.0, 0.0, mode=gpu.MODE_GN_P2PLANE)
    icp.setTargetPointCloud(w["map"], w["normals"])
    b = gpu.Batch(icp, 2, 2 * w["scan"].shape[0] + 1)
    b.upload([w["scan"], w["scan"]])
    b.set_initial([w["T0"], w["T_gt"]])
    b.run()
    ref = b.results()
    ctx.time_searches(True)
    b.run()
    b.run()
    t = ctx.search_times()
    assert len(t) == 12 and all(x > 0.0 for x in t)
    assert len(ctx.search_times(cap=5)) == 5
    total, n = ctx.search_time()
    assert n == 12 and abs(total - sum(t)) < 1e-3 * max(total, 1e-3)
    assert ctx.search_time()[1] == 0
    ctx.time_searches(False)
    # the timed (non-graph) path gives the same results as the graph replay
    for x, y in zip(ref, b.results()):
        assert np.array_equal(x.transformation.view(np.uint32), y.transformation.view(np.uint32))
