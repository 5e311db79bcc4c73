"""Pins the CPU oracle (oracle/ssf_oracle.c).  The reference has no golden vectors and cannot
be built here ("parity unpinned", see the oracle header), so the restatement is checked against
independent implementations: O(N*M) brute force, cv2.flann KDTREE_SINGLE (the FLANN lineage PCL
wraps), scipy cKDTree, numpy SVD / group-by, analytic ground-truth poses, and the committed
golden fixture tests/golden/c1_mini.npz."""
import os

import numpy as np
import pytest

from conftest import ROOT, pose_delta
from oracle import oracle


# ---- nearest neighbour -------------------------------------------------------------------------
def _f32_sqdist(a, b):
    d = (a - b).astype(np.float32)
    r = d[:, 0] * d[:, 0]
    r = r + d[:, 1] * d[:, 1]
    r = r + d[:, 2] * d[:, 2]
    return r.astype(np.float32)


def test_kdtree_equals_brute_force():
    rng = np.random.default_rng(0)
    m = rng.uniform(-5, 5, (20000, 3)).astype(np.float32)
    q = rng.uniform(-6, 6, (3000, 3)).astype(np.float32)
    i1, d1 = oracle.KdTree(m).nn(q)
    i2, d2 = oracle.nn_brute(m, q)
    assert np.array_equal(i1, i2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))
    # distance is the float32, left-to-right, non-FMA sum (flann::L2_Simple)
    assert np.array_equal(d1.view(np.uint32), _f32_sqdist(q, m[i1]).view(np.uint32))


def test_kdtree_ties_lowest_index():
    g = np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    m = np.concatenate([g, g, g[::-1]])
    q = np.concatenate([g[:300] + 0.5, g[:300]]).astype(np.float32)
    i1, d1 = oracle.KdTree(m).nn(q)
    i2, d2 = oracle.nn_brute(m, q)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    # brute force itself: first minimum in ascending index order
    d_all = ((q[:, None, :] - m[None, :, :]) ** 2).sum(-1)
    assert np.array_equal(i2, d_all.argmin(1).astype(np.int32))


def test_kdtree_multithread_identical():
    rng = np.random.default_rng(3)
    m = rng.normal(size=(50000, 3)).astype(np.float32)
    q = rng.normal(size=(10000, 3)).astype(np.float32)
    t = oracle.KdTree(m)
    i1, d1 = t.nn(q, threads=1)
    i2, d2 = t.nn(q, threads=4)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)


def test_kdtree_vs_cv2_flann_and_scipy():
    cv2 = pytest.importorskip("cv2")
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(7)
    m = rng.uniform(0, 50, (200_000, 3)).astype(np.float32)
    q = rng.uniform(0, 50, (5000, 3)).astype(np.float32)
    oi, od = oracle.KdTree(m).nn(q)
    si = cKDTree(m).query(q)[1].astype(np.int32)
    assert np.array_equal(oi, si)  # tie-free data: indices agree with an independent exact tree
    index = cv2.flann_Index(m, dict(algorithm=4, leaf_max_size=15))  # KDTREE_SINGLE, as PCL configures FLANN
    fi, fd = index.knnSearch(q, 1, params=dict(checks=-1, eps=0.0))
    assert np.array_equal(oi, fi[:, 0].astype(np.int32))
    assert np.array_equal(od.view(np.uint32), fd[:, 0].astype(np.float32).view(np.uint32))  # bit-equal d2


def test_kdtree_empty_and_single():
    i, d = oracle.KdTree(np.zeros((0, 3), np.float32)).nn(np.zeros((2, 3), np.float32))
    assert (i == -1).all()
    i, d = oracle.KdTree(np.array([[1, 2, 3]], np.float32)).nn(np.array([[1, 2, 4]], np.float32))
    assert i[0] == 0 and d[0] == 1.0


# ---- Kabsch / SVD --------------------------------------------------------------------------------
def test_svd3_reconstructs_and_matches_numpy():
    rng = np.random.default_rng(1)
    for _ in range(50):
        H = rng.normal(size=(3, 3)).astype(np.float32) * rng.choice([1e-3, 1.0, 1e3])
        U, S, V = oracle.svd3(H)
        assert np.allclose(U @ np.diag(S) @ V.T, H, rtol=0, atol=5e-6 * np.abs(H).max())
        assert np.allclose(S, np.linalg.svd(H.astype(np.float64))[1], rtol=2e-6, atol=1e-6 * np.abs(H).max())
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-5) and np.allclose(V.T @ V, np.eye(3), atol=1e-5)
        assert S[0] >= S[1] >= S[2] >= 0


def test_kabsch_recovers_rigid_motion_and_handles_reflection():
    rng = np.random.default_rng(2)
    P = rng.normal(size=(500, 3)).astype(np.float32)
    ang = 0.3
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    t = np.array([0.5, -0.2, 0.1])
    Q = (P @ R.T + t).astype(np.float32)
    T = oracle.kabsch(P, Q)
    assert np.allclose(T[:3, :3], R, atol=1e-5) and np.allclose(T[:3, 3], t, atol=1e-5)
    # numpy reference of the same closed form
    pc, qc = P - P.mean(0), Q - Q.mean(0)
    U, _, Vt = np.linalg.svd(pc.T.astype(np.float64) @ qc.astype(np.float64))
    assert np.allclose(T[:3, :3], (Vt.T @ U.T), atol=1e-5)
    # planar, mirrored data: det fix keeps a proper rotation (cpp:145-149)
    Pp = P.copy()
    Pp[:, 2] = 0
    Qp = Pp.copy()
    Qp[:, 0] *= -1
    Tm = oracle.kabsch(Pp, Qp)
    assert np.linalg.det(Tm[:3, :3].astype(np.float64)) > 0.99


# ---- voxel grid ----------------------------------------------------------------------------------
def _voxel_numpy(xyz, leaf):
    xyz = xyz[np.isfinite(xyz).all(1)]
    inv = np.float32(1.0) / np.float32(leaf)
    mn, mx = xyz.min(0), xyz.max(0)
    minb = np.floor(mn * inv).astype(np.int64)
    maxb = np.floor(mx * inv).astype(np.int64)
    div = maxb - minb + 1
    ijk = (np.floor(xyz * inv) - minb.astype(np.float32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(idx, kind="stable")
    out = []
    start = 0
    sidx = idx[order]
    while start < len(order):
        end = start
        acc = np.zeros(3, np.float32)
        while end < len(order) and sidx[end] == sidx[start]:
            acc = (acc + xyz[order[end]]).astype(np.float32)
            end += 1
        out.append(acc / np.float32(end - start))
        start = end
    return np.array(out, np.float32)


def test_voxel_grid_vs_numpy_groupby():
    rng = np.random.default_rng(4)
    xyz = rng.uniform(-2, 2, (3000, 3)).astype(np.float32)
    xyz[10] = np.nan
    o, refused = oracle.voxel_grid(xyz, 0.25)
    assert not refused
    ref = _voxel_numpy(xyz, 0.25)
    assert np.array_equal(o[:, :3].view(np.uint32), ref.view(np.uint32))
    assert (o[:, 3] == 1.0).all()


def test_voxel_grid_refuses_on_index_overflow():
    c = np.array([[0, 0, 0], [3000, 3000, 3000]], np.float32)
    o, refused = oracle.voxel_grid(c, 0.001)
    assert refused and np.array_equal(o[:, :3], c)


# ---- whole loop ----------------------------------------------------------------------------------
def test_reference_icp_moves_towards_ground_truth(small_world):
    w = small_world
    tree = oracle.KdTree(w["map"])
    res, corr, tr = oracle.icp_reference(tree, w["scan"], w["T0"], trace=True)
    assert not res.aborted and 1 <= res.iterations <= 10
    assert pose_delta(res.T, w["T_gt"])[0] < pose_delta(w["T0"], w["T_gt"])[0]
    # errors recorded for every executed pass, searched passes consistent with n_searches
    assert np.isfinite(tr.iter_err[:res.iterations]).all()
    assert 1 + tr.iter_searched.sum() == res.n_searches
    # source shrinks monotonically (cpp:77-83)
    cnt = tr.count[:res.n_searches]
    assert (np.diff(cnt) <= 0).all()
    # corr: rows alive at the end have an index, the rest are -1
    assert (corr >= 0).sum() == res.k_final


def test_reference_icp_iteration_zero_never_researches(small_world):
    w = small_world
    tree = oracle.KdTree(w["map"])
    res, corr, tr = oracle.icp_reference(tree, w["scan"], w["T0"], trace=True, transformation_epsilon=1e9,
                                         num_iterations=4)
    # last_error_ starts at FLT_MAX (cpp:205): |FLT_MAX - e| is not < eps even for huge eps < FLT_MAX
    assert tr.iter_searched[0] == 0 and tr.iter_searched[1:4].all()


def test_reference_icp_abort_and_break(small_world):
    w = small_world
    tree = oracle.KdTree(w["map"])
    far = np.eye(4)
    far[0, 3] = 1e4
    res, corr, _ = oracle.icp_reference(tree, w["scan"], far)
    assert res.aborted and res.iterations == 0 and res.error == pytest.approx(1e6) and not res.has_converged
    assert np.array_equal(res.T, far.astype(np.float32))
    # acceptable error already met at pass 0 -> break, zero iterations, converged (cpp:215-219, 252)
    res, corr, _ = oracle.icp_reference(tree, w["scan"], w["T_gt"], acceptable_mean_error=10.0)
    assert res.iterations == 0 and res.has_converged and res.error < 10.0


def test_gn_point_to_plane_recovers_pose(small_world):
    w = small_world
    tree = oracle.KdTree(w["map"])
    r, corr = oracle.icp_gn(tree, w["scan"], w["T0"], mode="p2plane", normals=w["normals"], num_iterations=15)
    dt, dr = pose_delta(r.T, w["T_gt"])
    assert dt < 0.02 and dr < 2e-3
    r2, _ = oracle.icp_gn(tree, w["scan"], w["T0"], mode="p2p", num_iterations=15)
    assert pose_delta(r2.T, w["T_gt"])[0] < pose_delta(w["T0"], w["T_gt"])[0]


def test_golden_fixture_c1_mini():
    """Frozen outputs of the oracle at fixed seeds (tests/golden/make_golden.py)."""
    path = os.path.join(ROOT, "tests", "golden", "c1_mini.npz")
    g = np.load(path)
    tree = oracle.KdTree(g["map"])
    idx, d2 = tree.nn(g["queries"])
    assert np.array_equal(idx, g["nn_idx"]) and np.array_equal(d2.view(np.uint32), g["nn_d2"].view(np.uint32))
    res, corr, _ = oracle.icp_reference(tree, g["scan"], g["T0"])
    assert np.array_equal(res.T.view(np.uint32), g["ref_T"].view(np.uint32))
    assert res.iterations == int(g["ref_iterations"]) and res.n_searches == int(g["ref_n_searches"])
    assert np.array_equal(corr, g["ref_corr"])
    vox, _ = oracle.voxel_grid(g["scan"], 0.2)
    assert np.array_equal(vox.view(np.uint32), g["vox_02"].view(np.uint32))


# ---- independent restatements of the oracle-defined solver modes (numpy / scipy, float64) -------------
def _f32_transform(T, xyz):
    """applyTransformation in the reference's float32 expression order ((T0*x + T1*y) + T2*z) + T3."""
    T = np.asarray(T, np.float32)
    x, y, z = (np.ascontiguousarray(xyz[:, k], np.float32) for k in range(3))
    return np.stack([((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3] for r in range(3)], axis=1).astype(np.float32)


def _nn_f32(tree, map_xyz, P):
    """cKDTree neighbour + flann::L2_Simple distance in float32, ((dx*dx + dy*dy) + dz*dz)."""
    _, idx = tree.query(P.astype(np.float64))
    d = P - map_xyz[idx]
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    return idx, d2.astype(np.float32)


@pytest.mark.parametrize("mode", ["p2p", "p2plane"])
def test_gn_steps_vs_numpy_restatement(small_world, mode):
    """The Gauss-Newton modes have no counterpart in the reference: the oracle defines
    them, so the oracle itself is pinned against an independent float64 numpy / scipy restatement --
    J^T J and J^T r by matrix products, numpy.linalg.solve, scipy.linalg.expm of the twist -- over three
    chained iterations.  Poses agree to float32 rounding; correspondences are identical."""
    from scipy.linalg import expm
    from scipy.spatial import cKDTree
    w = small_world
    m = np.ascontiguousarray(w["map"][:, :3], np.float32)
    nrm = np.ascontiguousarray(w["normals"][:, :3], np.float64)
    ck = cKDTree(m.astype(np.float64))
    T = np.asarray(w["T0"], np.float32).copy()
    tree = oracle.KdTree(w["map"])
    for it in range(1, 4):
        P = _f32_transform(T, w["scan"])
        idx, d2 = _nn_f32(ck, m, P)
        ok = d2 < np.float32(0.5)
        p = P[ok].astype(np.float64)
        q = m[idx[ok]].astype(np.float64)
        e = p - q
        if mode == "p2plane":
            n = nrm[idx[ok]]
            A = np.concatenate([np.cross(p, n), n], axis=1)      # K x 6
            r = np.einsum("ij,ij->i", n, e)
            H, g = A.T @ A, A.T @ r
        else:
            H, g = np.zeros((6, 6)), np.zeros(6)
            for k in range(3):                                      # the three rows of [-[p]x | I]
                ek = np.zeros(3); ek[k] = 1.0
                Jk = np.concatenate([np.cross(p, np.broadcast_to(ek, p.shape)), np.broadcast_to(ek, p.shape)], axis=1)  # row k of -[p]x is p x e_k
                H += Jk.T @ Jk
                g += Jk.T @ e[:, k]
        x = np.linalg.solve(H, -g)
        tw = np.zeros((4, 4))
        tw[:3, :3] = np.array([[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]])
        tw[:3, 3] = x[3:]
        step = expm(tw)
        step[:3, 3] = x[3:]   # the oracle's update is [exp([w]x) | t], not the full SE(3) exponential
        T = (step @ T.astype(np.float64)).astype(np.float32)
        res, corr = oracle.icp_gn(tree, w["scan"], w["T0"], mode=mode, normals=w["normals"], num_iterations=it)
        assert res.iterations == it and res.k_final == int(ok.sum())
        if it == 1:
            assert np.array_equal(corr, np.where(ok, idx, -1))
        assert np.allclose(res.T, T, rtol=0, atol=2e-6 * it), (it, np.abs(res.T - T).max())
    assert pose_delta(T, w["T_gt"])[0] < pose_delta(w["T0"], w["T_gt"])[0]


def test_o3d_flow_vs_numpy_restatement(small_world):
    """Open3D's registration_icp control flow restated with numpy: true-distance
    threshold, every source point every iteration, Kabsch/Umeyama step from numpy.linalg.svd in float64,
    T <- step * T; three chained iterations against the oracle."""
    from scipy.spatial import cKDTree
    w = small_world
    m = np.ascontiguousarray(w["map"][:, :3], np.float32)
    ck = cKDTree(m.astype(np.float64))
    T = np.asarray(w["T0"], np.float32).copy()
    tree = oracle.KdTree(w["map"])
    thr = np.float32(0.5) * np.float32(0.5)
    for it in range(1, 4):
        P = _f32_transform(T, w["scan"])
        idx, d2 = _nn_f32(ck, m, P)
        ok = d2 < thr
        p, q = P[ok].astype(np.float64), m[idx[ok]].astype(np.float64)
        pb, qb = p.mean(0), q.mean(0)
        U, S, Vt = np.linalg.svd((p - pb).T @ (q - qb))
        D = np.diag([1.0, 1.0, np.sign(np.linalg.det(Vt.T @ U.T))])
        R = Vt.T @ D @ U.T
        step = np.eye(4)
        step[:3, :3], step[:3, 3] = R, qb - R @ pb
        T = (step @ T.astype(np.float64)).astype(np.float32)
        res, fit, corr = oracle.icp_o3d(tree, w["scan"], w["T0"], 0.5, max_iteration=it)
        assert res.iterations == it
        assert np.allclose(res.T, T, rtol=0, atol=2e-6 * it), (it, np.abs(res.T - T).max())
    # fitness / rmse of the last pass are those of the pose after the last step
    P = _f32_transform(T, w["scan"])
    idx, d2 = _nn_f32(ck, m, P)
    ok = d2 < thr
    assert abs(fit - ok.mean()) < 1e-6
    e = P[ok].astype(np.float64) - m[idx[ok]]
    assert abs(res.error - np.sqrt((e * e).sum() / ok.sum())) < 1e-6


def _reference_loop_numpy(m, scan, T0, thr, n_iter, acc, eps):
    """calculateAlignment restated line by line from icp_point_to_point.cpp:185-254 (and :57-84, :99-170)
    with numpy float32 -- written against the reference file, not against oracle/ssf_oracle.c.  The 3x3
    SVD is LAPACK's (float32), the covariance a BLAS product: low-order bits differ from Eigen's."""
    from scipy.spatial import cKDTree
    f32 = np.float32
    ck = cKDTree(m.astype(np.float64))
    margins = []

    def correspondences(P):                                   # :57-84, d2 against the UNSQUARED threshold (:70)
        idx, d2 = _nn_f32(ck, m, P)
        ok = d2 < f32(thr)
        return P[ok], m[idx[ok]]

    T = np.asarray(T0, f32).copy()
    P = _f32_transform(T, scan)                               # :191-192
    P, Q = correspondences(P)                                 # :195
    searches = 1
    if P.shape[0] < 10:                                       # :196-200
        return np.asarray(T0, f32), f32(1e6), 0, False, searches, margins
    its, last = 0, np.finfo(f32).max                          # :203-205
    for _ in range(n_iter):
        d = P - Q
        norms = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(f32)
        err = f32(np.add.accumulate(norms, dtype=f32)[-1] / f32(P.shape[0]))   # :161-170, sequential float sum
        margins.append(abs(float(err) - acc))
        if err < f32(acc):                                    # :215-219
            last = err
            break
        margins.append(abs(abs(float(last) - float(err)) - eps))
        if abs(f32(last - err)) < f32(eps):                   # :221-224, lazy re-search on the SHRUNK cloud
            P, Q = correspondences(P)
            searches += 1
        K = f32(P.shape[0])
        cs = np.add.accumulate(P, axis=0, dtype=f32)[-1] / K  # :116-123
        ct = np.add.accumulate(Q, axis=0, dtype=f32)[-1] / K
        H = ((P - cs).T @ (Q - ct)).astype(f32)               # :126-134
        U, _, Vt = np.linalg.svd(H)                           # :137-139
        V = Vt.T.copy()
        R = V @ U.T
        if np.linalg.det(R.astype(np.float64)) < 0:           # :142-149
            V[:, 2] *= -1
            R = V @ U.T
        Ts = np.eye(4, dtype=f32)
        Ts[:3, :3] = R
        Ts[:3, 3] = ct - R @ cs                               # :152
        T = (Ts @ T).astype(f32)                              # :228, left-multiplied
        P = _f32_transform(Ts, P)                             # :230, points advanced by the STEP
        last = err                                            # :232
        its += 1
    return T, f32(last), its, bool(last < f32(acc)), searches, margins


@pytest.mark.parametrize("params", [(0.5, 10, 0.05, 1e-5), (5.0, 12, 0.4, 1e-2), (0.5, 6, 0.0, 1.0)])
def test_reference_loop_vs_numpy_restatement(small_world, params):
    """The oracle's restatement of calculateAlignment against a second, independent one written from the
    reference source with numpy (fine and coarse parameter sets of localization_node.cpp:24-27,226-229,
    and an eps so large that every pass re-searches).  Control flow -- iterations, searches, the
    surviving-correspondence count, has_converged -- must be identical; pose and error agree to the
    float32 noise of LAPACK-vs-Jacobi SVD and BLAS-vs-sequential covariance."""
    w = small_world
    thr, n_iter, acc, eps = params
    m = np.ascontiguousarray(w["map"][:, :3], np.float32)
    T, err, its, conv, searches, margins = _reference_loop_numpy(m, w["scan"], w["T0"], thr, n_iter, acc, eps)
    res, corr, _ = oracle.icp_reference(oracle.KdTree(w["map"]), w["scan"], w["T0"], thr, n_iter, acc, eps)
    if min(margins) < 5e-6:
        pytest.skip("a stop / re-search decision sits within float noise of its threshold")
    assert (res.iterations, res.n_searches, bool(res.has_converged)) == (its, searches, conv)
    assert abs(res.error - err) < 2e-5
    dt, dr = pose_delta(res.T, T)
    assert dt < 1e-4 and dr < 1e-5, (dt, dr)
