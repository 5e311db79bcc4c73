"""GPU parity at the sizes BASELINE.json names (-m gpu), through the C ABI, against the CPU oracle
(and, where the reference's own text defines the result, against oracle/_ref directly):

  c2  64x2048 scan -> voxel 0.2 -> point-to-plane GN vs the 5M-point map   (the benchmarked config)
  c4  32x1024 scans, the reference's own loop (REFERENCE / STRICT) vs the 5M-point map, bit-exact
  c5  128x2048 scan (40 m), voxel 0.05, 30 iterations, point-to-point GN    (map reduced to 5M points)
  c3  the map sharded over R ranks, run on ONE GPU (R handles, host-side summing hook) vs the oracle
plus the Gauss-Newton stop rules with non-zero thresholds, the debug text, the empty-source sentinel
and the cell-size search on a target with many duplicates.

Bars (north_star): correspondences bit-exact per search; pose within 1e-4 m / 1e-5 rad and the same
iteration count +-1 for the Gauss-Newton modes; everything bit-identical for the reference's loop.
"""
import threading

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


@pytest.fixture(scope="module")
def big(ora):
    """The 5M-point street map of configs 2 and 4 (with normals) and its oracle KD-tree."""
    from ssf_gpu import synth
    xyz, nrm, half = synth.make_map(5_000_000, normals=True)
    return dict(map=xyz, normals=nrm, half=half, tree=ora.KdTree(xyz))


@pytest.fixture(scope="module")
def big_icp(gpu, big):
    icp = gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, mode=gpu.MODE_GN_P2PLANE)
    icp.setTargetPointCloud(big["map"], big["normals"])
    return icp


def _pose_pair(big, k, beams, az, max_range=100.0):
    from ssf_gpu import synth
    T = synth.street_pose(k, half=big["half"])
    return synth.make_scan(T, beams, az, scan_id=k, max_range=max_range), synth.perturb_pose(T, k), T


# ---- config 2 -------------------------------------------------------------------------------------
def test_c2_full_size_point_to_plane(gpu, ora, big, big_icp):
    icp = big_icp
    icp.setMode(gpu.MODE_GN_P2PLANE)
    icp.setMaxCorrespondenceDist(0.5)
    icp.setNumIterations(10)
    icp.setAcceptableMeanError(0.0)
    icp.setTransformationEpsilon(0.0)
    icp.setSourceVoxelLeaf(0.2)
    scans, inits, gts = zip(*[_pose_pair(big, 40 * d + 3, 64, 2048) for d in range(4)])
    batch = icp.align_batch(list(scans), list(inits))
    for s, T0, T_gt, rb in zip(scans, inits, gts, batch):
        vox, refused = ora.voxel_grid(s, 0.2)
        assert not refused and rb.n_source == vox.shape[0]
        o, ocorr = ora.icp_gn(big["tree"], vox, T0, mode="p2plane", normals=big["normals"], num_iterations=10, threads=8)
        dt, dr = pose_delta(rb.transformation, o.T)
        assert dt < TOL_T and dr < TOL_R, (dt, dr)
        assert abs(rb.iterations - o.iterations) <= 1 and rb.k_final == o.k_final
        assert rb.error == pytest.approx(o.error, rel=1e-4)
        assert pose_delta(rb.transformation, T_gt)[0] < 0.05
    # one of them through the single-scan entry: every correspondence of the last search
    icp.setSourcePointCloud(scans[0])
    icp.setInitialTransformation(inits[0])
    r1 = icp.calculateAlignment()
    vox, _ = ora.voxel_grid(scans[0], 0.2)
    o, ocorr = ora.icp_gn(big["tree"], vox, inits[0], mode="p2plane", normals=big["normals"], num_iterations=10, threads=8)
    assert np.array_equal(r1.transformation.view(np.uint32), batch[0].transformation.view(np.uint32))
    gcorr = icp.correspondences()[:vox.shape[0]]
    assert (gcorr == ocorr).mean() >= 0.999
    icp.setSourceVoxelLeaf(0.0)


def test_c2_searches_bit_exact(gpu, ora, big, big_icp):
    """Kernel-level contract on the oracle's own query arrays: the queries of search 0 (initial guess)
    and of a converged search (oracle's final pose), all of a 64x2048 voxel-downsampled scan."""
    scan, T0, _ = _pose_pair(big, 91, 64, 2048)
    vox, _ = ora.voxel_grid(scan, 0.2)
    o, _ = ora.icp_gn(big["tree"], vox, T0, mode="p2plane", normals=big["normals"], num_iterations=10, threads=8)
    for T in (T0.astype(np.float32), o.T):
        q = np.empty((vox.shape[0], 3), np.float32)   # the reference's float expression ((a*x + b*y) + c*z) + d
        for r in range(3):
            q[:, r] = ((T[r, 0] * vox[:, 0] + T[r, 1] * vox[:, 1]) + T[r, 2] * vox[:, 2]) + T[r, 3]
        gi, gd = big_icp.nearest(q, 0.5)
        oi, od = big["tree"].nn(q, threads=8)
        inside = od < np.float32(0.5)
        assert inside.sum() > 0.5 * q.shape[0]
        assert np.array_equal(gi[inside], oi[inside])
        assert np.array_equal(gd[inside].view(np.uint32), od[inside].view(np.uint32))
        assert (gi[~inside] == -1).all()


# ---- config 4 -------------------------------------------------------------------------------------
def test_c4_reference_loop_bit_exact_on_5M_map(gpu, ora, big):
    icp = gpu.ICPPointToPoint(0.5, 10, 0.05, 1e-5, mode=gpu.MODE_REFERENCE, reduce=gpu.REDUCE_STRICT)
    icp.setTargetPointCloud(big["map"])
    scans, inits, _ = zip(*[_pose_pair(big, 17 * d + 1, 32, 1024) for d in range(6)])
    batch = icp.align_batch(list(scans), list(inits))
    schedules = set()
    for s, T0, rb in zip(scans, inits, batch):
        o, ocorr, _ = ora.icp_reference(big["tree"], s, T0, threads=8)
        assert np.array_equal(rb.transformation.view(np.uint32), o.T.view(np.uint32))
        assert np.float32(rb.error).view(np.uint32) == np.float32(o.error).view(np.uint32)
        assert (rb.iterations, rb.n_searches, rb.k_final, rb.has_converged) == (o.iterations, o.n_searches, o.k_final,
                                                                               bool(o.has_converged))
        schedules.add((o.iterations, o.n_searches))
    icp.setSourcePointCloud(scans[2])
    icp.setInitialTransformation(inits[2])
    r1 = icp.calculateAlignment()
    o, ocorr, _ = ora.icp_reference(big["tree"], scans[2], inits[2], threads=8)
    assert np.array_equal(r1.transformation.view(np.uint32), o.T.view(np.uint32))
    assert np.array_equal(icp.correspondences(), ocorr)


def test_offline_sequence_many_scans_bit_exact(gpu, ora, big):
    """Config 4's shape in the many: more scans than two per SM switch the reference loop to its
    4-blocks-per-SM shape (256 threads, 704-row stages).  The ordered chains are the same sums in either
    shape: every scan of a 320-scan batch equals the same scan run in batches of 64, bit for bit, and the
    oracle on a sample."""
    from ssf_gpu import synth
    icp = gpu.ICPPointToPoint(0.5, 10, 0.05, 1e-5, mode=gpu.MODE_REFERENCE, reduce=gpu.REDUCE_STRICT)
    icp.setTargetPointCloud(big["map"])
    scans, inits = [], []
    for d in range(320):
        T = synth.street_pose(9 * d + 2, half=big["half"])
        scans.append(synth.make_scan(T, 16, 256, scan_id=5000 + d, max_range=60.0))
        inits.append(synth.perturb_pose(T, 5000 + d))
    many = icp.align_batch(scans, inits)
    few = []
    for a in range(0, 320, 64):
        few += icp.align_batch(scans[a:a + 64], inits[a:a + 64])
    for x, y in zip(many, few):
        assert np.array_equal(x.transformation.view(np.uint32), y.transformation.view(np.uint32))
        assert np.float32(x.error).view(np.uint32) == np.float32(y.error).view(np.uint32)
        assert (x.iterations, x.n_searches, x.k_final, x.has_converged) == (y.iterations, y.n_searches, y.k_final,
                                                                           y.has_converged)
    for d in (0, 131, 319):
        o, _, _ = ora.icp_reference(big["tree"], scans[d], inits[d], threads=8)
        assert np.array_equal(many[d].transformation.view(np.uint32), o.T.view(np.uint32))
        assert (many[d].iterations, many[d].n_searches, many[d].k_final) == (o.iterations, o.n_searches, o.k_final)


def test_gpu_equals_reference_sources_directly(gpu, c1_world):
    """The CUDA path against oracle/_ref (the reference's own icp_point_to_point.cpp compiled unmodified),
    with no restatement in between: config 1 at full size, fine and coarse parameter sets."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libssf_ref.so not shipped")
    w = c1_world
    g = gpu.ICPPointToPoint(0.5, 10, 0.05, 1e-5)
    g.setTargetPointCloud(w["map"])
    g.setSourcePointCloud(w["scan"])
    g.setInitialTransformation(w["T0"])
    r = ref.ICPPointToPoint(0.5, 10, 0.05, 1e-5)
    r.setDebugMode(False)
    r.setTargetPointCloud(w["map"])
    r.setSourcePointCloud(w["scan"])
    r.setInitialTransformation(w["T0"])
    for prm in ((0.5, 10, 0.05, 1e-5), (5.0, 80, 0.4, 1e-2)):   # localization_node.cpp:24-27 / 226-229
        for o in (g, r):
            o.setMaxCorrespondenceDist(prm[0])
            o.setNumIterations(prm[1])
            o.setAcceptableMeanError(prm[2])
            o.setTransformationEpsilon(prm[3])
        a, b = g.calculateAlignment(), r.calculateAlignment()
        assert np.array_equal(a.transformation.view(np.uint32), b.T.view(np.uint32))
        assert np.float32(a.error).view(np.uint32) == np.float32(b.error).view(np.uint32)
        assert a.iterations == b.iterations and a.has_converged == bool(b.has_converged)


def test_debug_text_equals_reference(gpu, small_world, capfd):
    """setDebugMode(true): the library's stdout is the reference's (cpp:172-183, 237-246), line for line."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libssf_ref.so not shipped")
    w = small_world
    r = ref.ICPPointToPoint(0.5, 10, 0.05, 1e-5)
    r.setDebugMode(True)
    r.setTargetPointCloud(w["map"])
    r.setSourcePointCloud(w["scan"])
    r.setInitialTransformation(w["T0"])
    r.calculateAlignment()
    g = gpu.ICPPointToPoint(0.5, 10, 0.05, 1e-5)
    g.setDebugMode(True)
    g.setTargetPointCloud(w["map"], w["normals"])
    g.setSourcePointCloud(w["scan"])
    g.setInitialTransformation(w["T0"])
    capfd.readouterr()
    g.calculateAlignment()
    out = capfd.readouterr().out
    assert out == r.stdout, (out, r.stdout)
    # the Gauss-Newton modes print the same kind of trace
    g.setMode(gpu.MODE_GN_P2PLANE)
    g.setAcceptableMeanError(0.0)
    g.setTransformationEpsilon(0.0)
    res = g.calculateAlignment()
    lines = capfd.readouterr().out.splitlines()
    assert sum(ln.startswith("[ICP INFO] Iteration ") for ln in lines) == 10
    assert f"[ICP INFO] Total iterations taken: {res.iterations}" in lines
    assert "[ICP INFO] Final transformation matrix: " in lines


def test_empty_source_is_the_abort_sentinel(gpu, small_world, capfd):
    """setSourcePointCloud(empty cloud): 0 < 10 correspondences -> {T_init, 1e6, 0, false} and the
    reference's message (cpp:196-200), in every mode that has the rule."""
    w = small_world
    for mode in (gpu.MODE_REFERENCE, gpu.MODE_GN_P2P):
        icp = gpu.ICPPointToPoint(0.5, 10, 0.05, 1e-5, mode=mode)
        icp.setTargetPointCloud(w["map"])
        icp.setSourcePointCloud(np.zeros((0, 3), np.float32))
        icp.setInitialTransformation(w["T0"])
        capfd.readouterr()
        r = icp.calculateAlignment()
        assert r.aborted and r.iterations == 0 and not r.has_converged and r.error == np.float32(1e6)
        assert np.array_equal(r.transformation, w["T0"].astype(np.float32))
        assert "[ICP ERROR] Not enough valid correspondences found. Aborting." in capfd.readouterr().err


# ---- config 5 shape ---------------------------------------------------------------------------------
def test_c5_shape_dense_scan_30_iterations(gpu, ora, big, big_icp):
    icp = big_icp
    icp.setMode(gpu.MODE_GN_P2P)
    icp.setMaxCorrespondenceDist(0.5)
    icp.setNumIterations(30)
    icp.setAcceptableMeanError(0.0)
    icp.setTransformationEpsilon(0.0)
    icp.setSourceVoxelLeaf(0.05)
    scan, T0, T_gt = _pose_pair(big, 211, 128, 2048, max_range=40.0)
    icp.setSourcePointCloud(scan)
    icp.setInitialTransformation(T0)
    r = icp.calculateAlignment()
    vox, refused = ora.voxel_grid(scan, 0.05)
    assert not refused and r.n_source == vox.shape[0] and vox.shape[0] > 100_000
    o, ocorr = ora.icp_gn(big["tree"], vox, T0, mode="p2p", num_iterations=30, threads=8)
    dt, dr = pose_delta(r.transformation, o.T)
    assert dt < TOL_T and dr < TOL_R, (dt, dr)
    assert r.iterations == o.iterations == 30 and r.k_final == o.k_final
    assert (icp.correspondences()[:vox.shape[0]] == ocorr).mean() >= 0.999
    icp.setSourceVoxelLeaf(0.0)


# ---- Gauss-Newton stop rules ------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["p2p", "p2plane"])
def test_gn_stop_rules_fire(gpu, ora, small_world, mode):
    """err < acceptable_mean_error (ssf_oracle.c:738 region) and max|x| < transformation_epsilon (:748):
    both rules must end the loop at the oracle's pass with the oracle's convergence flag."""
    w = small_world
    tree = ora.KdTree(w["map"])
    m = gpu.MODE_GN_P2PLANE if mode == "p2plane" else gpu.MODE_GN_P2P
    # per-pass errors of the oracle: a run of k passes reports the error measured by pass k - 1
    errs = [ora.icp_gn(tree, w["scan"], w["T0"], mode=mode, normals=w["normals"], num_iterations=k)[0].error
            for k in range(1, 9)]
    cases = []
    for k in (1, 2, 4):   # a threshold between the errors of pass k - 1 and pass k: the break fires at pass k
        if errs[k] < errs[k - 1] * 0.98:
            cases.append(dict(acceptable_mean_error=float(np.sqrt(errs[k] * errs[k - 1])), transformation_epsilon=0.0))
    assert cases
    cases += [dict(acceptable_mean_error=0.0, transformation_epsilon=e) for e in (3e-2, 3e-3, 3e-4, 3e-5)]
    cases.append(dict(acceptable_mean_error=1e9, transformation_epsilon=0.0))   # stops in pass 0, nothing solved
    fired_acc = fired_eps = 0
    for c in cases:
        o, ocorr = ora.icp_gn(tree, w["scan"], w["T0"], mode=mode, normals=w["normals"], num_iterations=25, **c)
        icp = gpu.ICPPointToPoint(0.5, 25, c["acceptable_mean_error"], c["transformation_epsilon"], mode=m)
        icp.setTargetPointCloud(w["map"], w["normals"])
        icp.setSourcePointCloud(w["scan"])
        icp.setInitialTransformation(w["T0"])
        g = icp.calculateAlignment()
        assert g.has_converged == bool(o.has_converged), c
        if c["acceptable_mean_error"] > 0.0:
            assert g.iterations == o.iterations and g.n_searches == o.n_searches, c   # thresholds sit between two passes
            fired_acc += o.iterations < 25 and bool(o.has_converged)
        else:
            assert abs(g.iterations - o.iterations) <= 1, c
            fired_eps += o.iterations < 25 and bool(o.has_converged)
        if g.iterations == o.iterations:
            dt, dr = pose_delta(g.transformation, o.T)
            assert dt < TOL_T and dr < TOL_R, (c, dt, dr)
            assert g.error == pytest.approx(o.error, rel=1e-4)
    assert fired_acc >= 2 and fired_eps >= 2


# ---- cell-size search (map_build.cu) ------------------------------------------------------------------
@pytest.mark.parametrize("points_per_cell", [None, "1.0", "40"])
def test_target_with_ten_copies_of_every_point(gpu, ora, points_per_cell, monkeypatch):
    """A target whose occupancy keeps the automatic cell-size loop shrinking through all its passes: the
    view handed to the search must describe the grid the index was built with (exact NN, lowest index)."""
    if points_per_cell:
        monkeypatch.setenv("SSF_CELL_POINTS", points_per_cell)
    rng = np.random.default_rng(21)
    base = rng.uniform(-4, 4, (20_000, 3)).astype(np.float32)
    m = np.tile(base, (10, 1))
    q = np.concatenate([base[::5] + rng.normal(0, 0.03, base[::5].shape).astype(np.float32),
                        rng.uniform(-5, 5, (2000, 3)).astype(np.float32)])
    icp = gpu.ICPPointToPoint(0.5, 1, 0.0, 0.0)
    icp.setTargetPointCloud(m)
    gi, gd = icp.nearest(q, 0.5)
    oi, od = ora.KdTree(m).nn(q, threads=8)
    inside = od < np.float32(0.5)
    assert inside.sum() > 4000 and (oi[inside] < base.shape[0]).all()   # ten copies: the first one wins
    assert np.array_equal(gi[inside], oi[inside])
    assert np.array_equal(gd[inside].view(np.uint32), od[inside].view(np.uint32))
    assert (gi[~inside] == -1).all()
    # lidar-ring-like target: dense along lines
    t = np.linspace(0, 2 * np.pi, 40_000, dtype=np.float32)
    rings = np.concatenate([np.c_[r * np.cos(t), r * np.sin(t), np.full_like(t, 0.01 * r)] for r in (3.0, 5.0, 8.0, 13.0)])
    q2 = (rings[::37] + rng.normal(0, 0.05, rings[::37].shape)).astype(np.float32)
    icp.setTargetPointCloud(rings.astype(np.float32))
    gi, gd = icp.nearest(q2, 0.5)
    oi, od = ora.KdTree(rings).nn(q2, threads=8)
    inside = od < np.float32(0.5)
    assert np.array_equal(gi[inside], oi[inside]) and np.array_equal(gd[inside], od[inside])


# ---- config 3: map sharded, on one GPU ----------------------------------------------------------------
def _run_sharded_on_one_gpu(gpu, map_xyz, normals, scan, T0, world, m):
    """R shards of the map as R handles on ONE device (own stream each), driven by R host threads; the
    per-iteration hook sums the ranks' rows on the host behind a barrier (no kernel ever waits for
    another).  Returns (results per rank, correspondences per rank)."""
    import torch
    from ssf_gpu import shard
    dev = torch.device("cuda", 0)
    barrier = threading.Barrier(world)
    slots = [None] * world
    results, corrs, errors, hooks = [None] * world, [None] * world, [], []

    def make_hook(rank):
        def hook(_user, buf, count, stream):
            try:
                t = torch.as_tensor(shard._DevArray(int(buf), int(count)), device=dev)
                torch.cuda.ExternalStream(int(stream), device=dev).synchronize()
                slots[rank] = t.cpu().numpy().copy()
                barrier.wait(timeout=120)
                tot = slots[0].copy()
                for k in range(1, world):
                    tot = tot + slots[k]     # rank order on every rank
                barrier.wait(timeout=120)
                t.copy_(torch.from_numpy(tot))
                torch.cuda.synchronize(dev)
                return 0
            except Exception as e:  # pragma: no cover
                errors.append(repr(e))
                barrier.abort()
                return 1
        return shard.ALLREDUCE_FN(hook)

    def run(rank):
        try:
            ctx = gpu.Context(0)
            icp = gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, mode=m, context=ctx)
            icp.setTargetShard(shard.shard_map(map_xyz, normals, rank, world, 0.5))
            hooks.append(make_hook(rank))
            icp.setAllreduce(hooks[-1])
            icp.setSourcePointCloud(scan)
            icp.setInitialTransformation(T0)
            results[rank] = icp.calculateAlignment()
            corrs[rank] = icp.correspondences()
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))
            barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not errors, errors
    return results, corrs


def _check_sharded(results, corrs, o, ocorr):
    for r in results:
        assert np.array_equal(r.transformation.view(np.uint32), results[0].transformation.view(np.uint32))
        dt, dr = pose_delta(r.transformation, o.T)
        assert dt < TOL_T and dr < TOL_R, (dt, dr)
        assert abs(r.iterations - o.iterations) <= 1 and r.k_final == o.k_final
    stack = np.stack(corrs)                       # -2: row owned by another rank
    owners = (stack != -2).sum(0)
    assert (owners == 1).all()
    merged = stack.max(0)
    assert (merged == ocorr).mean() >= 0.999
    assert (merged >= 0).sum() == o.k_final


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("mode", ["p2plane", "p2p"])
def test_map_sharded_on_one_gpu_equals_oracle(gpu, ora, c1_world, world, mode):
    """Every rank of a map sharded R ways must report the oracle's unsharded result: pose, counts, and --
    merged over the owners -- every correspondence, in GLOBAL indices."""
    w = c1_world
    m = gpu.MODE_GN_P2PLANE if mode == "p2plane" else gpu.MODE_GN_P2P
    tree = ora.KdTree(w["map"])
    o, ocorr = ora.icp_gn(tree, w["scan"], w["T0"], mode=mode, normals=w["normals"], num_iterations=10, threads=8)
    results, corrs = _run_sharded_on_one_gpu(gpu, w["map"], w["normals"], w["scan"], w["T0"], world, m)
    _check_sharded(results, corrs, o, ocorr)


def test_c3_full_size_50M_map_unsharded_and_sharded(gpu, ora):
    """Config 3 at its full map size: the 50M-point city map with normals, a 32x1024 scan, point-to-plane GN.
    The map on one GPU as a whole, and cut into 4 column shards (run one after the other on this GPU), against
    the oracle's KD-tree over all 50M points."""
    from ssf_gpu import synth
    xyz, nrm, half = synth.make_map(50_000_000, normals=True)
    T = synth.street_pose(4321, half=half)
    scan = synth.make_scan(T, 32, 1024, scan_id=4321)
    T0 = synth.perturb_pose(T, 4321)
    tree = ora.KdTree(xyz)
    o, ocorr = ora.icp_gn(tree, scan, T0, mode="p2plane", normals=nrm, num_iterations=10, threads=8)
    del tree
    icp = gpu.ICPPointToPoint(0.5, 10, 0.0, 0.0, mode=gpu.MODE_GN_P2PLANE)
    icp.setTargetPointCloud(xyz, nrm)
    icp.setSourcePointCloud(scan)
    icp.setInitialTransformation(T0)
    r = icp.calculateAlignment()
    dt, dr = pose_delta(r.transformation, o.T)
    assert dt < TOL_T and dr < TOL_R, (dt, dr)
    assert abs(r.iterations - o.iterations) <= 1 and r.k_final == o.k_final
    assert (icp.correspondences() == ocorr).mean() >= 0.999
    icp.close()
    results, corrs = _run_sharded_on_one_gpu(gpu, xyz, nrm, scan, T0, 4, gpu.MODE_GN_P2PLANE)
    _check_sharded(results, corrs, o, ocorr)
