"""Pins the CPU oracle to the reference's OWN text.

``oracle/_ref/libssf_ref.so`` is ``localization/src/icp_point_to_point.cpp``,
``localization/src/brute_force_alignment.cpp`` and ``point_cloud_processing.hpp`` compiled
UNMODIFIED (``oracle/Makefile`` target ``ref``) against stand-in Eigen/PCL headers
(``oracle/ref_stubs/``).  Everything the reference spells out itself -- control flow, the squared-vs-
unsquared threshold, the shrinking source, the lazy re-search rule, the abort sentinel, composition
order, float expression order, debug prints -- therefore comes from the reference's source, and the
restatement in ``ssf_oracle.c`` has to agree with it BIT FOR BIT.  (Eigen's JacobiSVD / blocked GEMM
and FLANN's tie order remain restated from their published algorithms in the stand-ins: [ext].)
"""
import numpy as np
import pytest

from oracle import oracle, ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")

FINE = dict(max_correspondence_dist=0.5, num_iterations=10, acceptable_mean_error=0.05,
            transformation_epsilon=1e-5)        # localization_node.cpp:24-27
COARSE = dict(max_correspondence_dist=5.0, num_iterations=80, acceptable_mean_error=0.4,
              transformation_epsilon=1e-2)      # localization_node.cpp:226-229


def _ref_icp(world_map, scan, T0, prm, debug=False):
    icp = ref.ICPPointToPoint(prm["max_correspondence_dist"], prm["num_iterations"], prm["acceptable_mean_error"],
                              prm["transformation_epsilon"])
    icp.setDebugMode(debug)
    icp.setTargetPointCloud(world_map)
    icp.setSourcePointCloud(scan)
    icp.setInitialTransformation(T0)
    return icp.calculateAlignment(), icp


def _same(r, o):
    assert np.array_equal(r.T.view(np.uint32), o.T.view(np.uint32)), (r.T, o.T)
    assert np.float32(r.error).view(np.uint32) == np.float32(o.error).view(np.uint32), (r.error, o.error)
    assert r.iterations == o.iterations and bool(r.has_converged) == bool(o.has_converged)


def test_default_result_sentinel():
    r = ref.default_result()   # icp_point_to_point.h:28-39
    assert np.array_equal(r.T, np.eye(4, dtype=np.float32)) and r.error == np.float32(1e6)
    assert r.iterations == 0 and not r.has_converged


@pytest.mark.parametrize("prm", [FINE, COARSE, dict(FINE, transformation_epsilon=1e9),
                                 dict(FINE, acceptable_mean_error=0.0, num_iterations=25),
                                 dict(FINE, max_correspondence_dist=0.05)],
                         ids=["fine", "coarse", "search_every_pass", "no_break_25", "tight_threshold"])
def test_icp_reference_equals_ref_small(small_world, prm):
    w = small_world
    r, _ = _ref_icp(w["map"], w["scan"], w["T0"], prm)
    o, _, _ = oracle.icp_reference(oracle.KdTree(w["map"]), w["scan"], w["T0"], **prm)
    _same(r, o)


def test_icp_reference_equals_ref_many_poses(small_world):
    """20 scans / initial guesses (including far-off ones that converge slowly or not at all)."""
    from ssf_gpu import synth
    w = small_world
    tree = oracle.KdTree(w["map"])
    icp = ref.ICPPointToPoint(**{k: FINE[k] for k in FINE})
    icp.setDebugMode(False)
    icp.setTargetPointCloud(w["map"])
    seen_iters = set()
    for k in range(20):
        T_gt = synth.street_pose(10 + k, half=w["half"])
        scan = synth.make_scan(T_gt, beams=8, azimuths=256, scan_id=10 + k, max_range=60.0)
        T0 = synth.perturb_pose(T_gt, 10 + k)
        if k % 5 == 4:
            T0 = T0.copy()
            T0[:3, 3] += np.array([0.9, -0.6, 0.1])
        prm = FINE if k % 2 == 0 else COARSE
        icp.setMaxCorrespondenceDist(prm["max_correspondence_dist"])      # re-parameterised like the node does
        icp.setNumIterations(prm["num_iterations"])
        icp.setAcceptableMeanError(prm["acceptable_mean_error"])
        icp.setTransformationEpsilon(prm["transformation_epsilon"])
        icp.setSourcePointCloud(scan)
        icp.setInitialTransformation(T0)
        r = icp.calculateAlignment()
        o, _, _ = oracle.icp_reference(tree, scan, T0, **prm)
        _same(r, o)
        seen_iters.add(r.iterations)
    assert len(seen_iters) > 1


def test_abort_sentinel_and_message(small_world):
    """< 10 correspondences on the first search (icp_point_to_point.cpp:196-200)."""
    w = small_world
    T_far = w["T0"].copy()
    T_far[:3, 3] += 500.0
    r, icp = _ref_icp(w["map"], w["scan"], T_far, FINE)
    o, _, _ = oracle.icp_reference(oracle.KdTree(w["map"]), w["scan"], T_far, **FINE)
    assert o.aborted == 1
    _same(r, o)
    assert r.error == np.float32(1e6) and r.iterations == 0 and not r.has_converged
    assert np.array_equal(r.T, T_far.astype(np.float32))
    assert icp.stderr == "[ICP ERROR] Not enough valid correspondences found. Aborting.\n"


def test_icp_reference_equals_ref_c1(c1_world):
    """Config 1 at full size: 1M-point map, 32x1024 scan, fine and coarse parameter sets."""
    w = c1_world
    tree = oracle.KdTree(w["map"])
    for prm in (FINE, COARSE):
        r, _ = _ref_icp(w["map"], w["scan"], w["T0"], prm)
        o, _, _ = oracle.icp_reference(tree, w["scan"], w["T0"], **prm)
        _same(r, o)


def test_debug_trace_matches_oracle_errors(small_world):
    """printStepDebug (cpp:172-183) and the closing block (cpp:237-246): the per-pass errors the
    reference prints are the oracle's trace, and the text has the reference's exact wording."""
    w = small_world
    r, icp = _ref_icp(w["map"], w["scan"], w["T0"], FINE, debug=True)
    o, _, tr = oracle.icp_reference(oracle.KdTree(w["map"]), w["scan"], w["T0"], trace=True, **FINE)
    lines = icp.stdout.splitlines()
    it_lines = [ln for ln in lines if ln.startswith("[ICP INFO] Iteration ")]
    passes = o.iterations + (1 if o.has_converged else 0)
    assert len(it_lines) == passes
    for i, ln in enumerate(it_lines):
        assert ln == f"[ICP INFO] Iteration {i} - Error: {float(tr.iter_err[i]):g}"
    assert f"[ICP INFO] Total iterations taken: {o.iterations}" in lines
    assert f"[ICP INFO] Final error: {float(np.float32(o.error)):g}" in lines
    k = lines.index("[ICP INFO] Final transformation matrix: ")
    # Eigen's default IOFormat: %g coefficients right-aligned to the widest one, one space between
    cells = [[f"{float(v):g}" for v in row] for row in o.T]
    wid = max(len(c) for row in cells for c in row)
    assert lines[k + 1:k + 5] == [" ".join(c.rjust(wid) for c in row) for row in cells]


# ---- BruteForceAlignment ----------------------------------------------------------------------
def _bfa_prm(**kw):
    d = dict(x_step=0.1, y_step=0.1, z_step=0.05, x_range=0.4, y_range=0.4, z_range=0.1,
             yaw_step=float(np.float32(np.pi) / np.float32(18.0)), yaw_range=float(np.float32(np.pi) / np.float32(6.0)),
             mean_error_threshold=0.1)
    d.update(kw)
    return d


@pytest.mark.parametrize("thr", [0.1, 1e-6], ids=["early_exit", "exhaustive"])
def test_bfa_equals_ref(small_world, thr):
    """alignClouds (brute_force_alignment.cpp:65-136): success flag and returned transform, with
    the early return and with the full grid (best-so-far carried to a second call)."""
    w = small_world
    scan = oracle.remove_floor(oracle.subsample(w["scan"], 16))
    tgt = oracle.remove_floor(oracle.subsample(w["map"], 4))
    d = _bfa_prm(mean_error_threshold=thr)
    ok_r, T_r = ref.bfa_align(tgt, scan, w["T0"], ref.BfaParams(**d), n_calls=2)
    tree = oracle.KdTree(tgt)
    T_prev = w["T0"].astype(np.float32)
    for call in range(2):
        ok_o, T_o, _, _ = oracle.bfa_align(tree, scan, T_prev, oracle.BfaParams(**d))
        assert bool(ok_r[call]) == ok_o
        assert np.array_equal(T_r[call], T_o), (call, T_r[call], T_o)
        if ok_o:
            break
        T_prev = T_o   # map_T_sensor_previous_ = best_T (cpp:126)


# ---- point_cloud_processing.hpp ---------------------------------------------------------------
def test_preprocessing_equals_ref(small_world):
    w = small_world
    rng = np.random.default_rng(5)
    cloud = np.concatenate([w["scan"][:, :3], rng.uniform(-12, 12, (500, 3)).astype(np.float32)])
    for step in (1, 2, 3, 15, cloud.shape[0] + 5):
        assert np.array_equal(ref.subsample(cloud, step), oracle.subsample(cloud, step))
    assert np.array_equal(ref.remove_floor(cloud), oracle.remove_floor(cloud))
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = [1.5, -2.0, 0.3]
    for radius in (10.0, 3.0, 0.0):
        assert np.array_equal(ref.crop_radius(T, radius, cloud), oracle.crop_radius(T, radius, cloud)[0])
    # the node's chain: subsample(2) -> crop(I, 10 m) (localization_node.cpp:292-296)
    a = ref.crop_radius(np.eye(4), 10.0, ref.subsample(cloud, 2))
    b = oracle.crop_radius(np.eye(4), 10.0, oracle.subsample(cloud, 2))[0]
    assert np.array_equal(a, b)


def test_golden_fixture_is_what_the_reference_sources_return():
    """tests/golden/c1_mini.npz was generated by the oracle (tests/golden/make_golden.py); its reference-loop
    entries are, bit for bit, what the reference's own unmodified icp_point_to_point.cpp returns for the
    fixture's inputs with the node's fine parameters (localization_node.cpp:24-27)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_mini.npz"))
    r, _ = _ref_icp(g["map"], g["scan"], g["T0"], FINE)
    assert np.array_equal(r.T.view(np.uint32), g["ref_T"].view(np.uint32))
    assert np.float32(r.error).view(np.uint32) == g["ref_error"].view(np.uint32)
    assert r.iterations == int(g["ref_iterations"])

