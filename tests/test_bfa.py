"""Row N1: BruteForceAlignment (pose-grid search) -- oracle restatement and GPU parity."""
import numpy as np
import pytest

from oracle import oracle


def _case(small_world, xy_range=0.3):
    from ssf_gpu import synth
    w = small_world
    m = oracle.remove_floor(oracle.subsample(w["map"], 15))     # localization_node.cpp:211-212
    s = oracle.remove_floor(w["scan"])                           # :213
    prm = oracle.BfaParams.node_defaults()
    prm.x_range = prm.y_range = xy_range                         # smaller grid: the CPU side stays fast
    T_prev = synth.perturb_pose(w["T_gt"], 9, xy=0.1, yaw_deg=5.0).astype(np.float32)
    return m, s, prm, T_prev


def test_pose_sequences_match_reference_counts():
    prm = oracle.BfaParams.node_defaults()
    P = oracle.bfa_poses(np.eye(4), prm)
    assert P.shape[0] == 18 * 18 * 4 * 6                         # 7 776 with the node's settings
    # breadth-first order, both signs, i = 0 twice (brute_force_alignment.cpp:160-179)
    assert np.allclose(P[:6, 0, 3], 0) and np.allclose(P[0][:3, :3], np.eye(3))
    yaw = np.arctan2(P[:6, 1, 0], P[:6, 0, 0])
    step = np.pi / 18
    assert np.allclose(yaw, [0, 0, -step, step, -2 * step, 2 * step], atol=1e-6)
    assert np.allclose(P[6, 2, 3], 0.0) and np.allclose(P[12, 2, 3], -0.05) and np.allclose(P[18, 2, 3], 0.05)


def test_oracle_bfa_early_exit_consistent(small_world):
    m, s, prm, T_prev = _case(small_world, 0.2)
    tree = oracle.KdTree(m)
    ok_all, T_all, best_all, scores = oracle.bfa_align(tree, s, T_prev, prm, no_early_exit=True, threads=4)
    prm.mean_error_threshold = float(np.sort(scores)[3])         # a threshold some candidates beat
    ok, T, best, sc = oracle.bfa_align(tree, s, T_prev, prm, threads=4)
    first = int(np.nonzero(scores < prm.mean_error_threshold)[0][0])
    assert ok and np.array_equal(T, oracle.bfa_poses(T_prev, prm)[first])
    assert np.isnan(sc[first + 1:]).all() and np.array_equal(sc[:first + 1], scores[:first + 1])


@pytest.mark.gpu
def test_bfa_gpu_matches_oracle(small_world):
    import ssf_gpu
    m, s, prm, T_prev = _case(small_world)
    tree = oracle.KdTree(m)
    ok_o, T_o, best_o, scores_o = oracle.bfa_align(tree, s, T_prev, prm, no_early_exit=True, threads=8)
    bfa = ssf_gpu.BruteForceAlignment()
    bfa.setMeanErrorThreshold(prm.mean_error_threshold)
    bfa.setXYZStep(prm.x_step, prm.y_step, prm.z_step)
    bfa.setXYZRange(prm.x_range, prm.y_range, prm.z_range)
    bfa.setRotationStep(prm.yaw_step)
    bfa.setRotationRange(prm.yaw_range)
    bfa.setInitialGuess(T_prev)
    bfa.setSourceCloud(s)
    bfa.setTargetCloud(m)
    ok_g = bfa.alignClouds()
    assert np.array_equal(bfa.last_scores.view(np.uint32), scores_o.view(np.uint32))   # every score bit-exact
    assert ok_g == ok_o and np.float32(bfa.best_score) == np.float32(best_o)
    assert np.array_equal(bfa.getBestTransformation(), T_o)
    assert not bfa.firstAlignmentCompleted()
    # second call starts from the best candidate (cpp:126) and ignores new guesses (cpp:44-51)
    bfa.setInitialGuess(np.eye(4))
    ok2_o, T2_o, best2_o, _ = oracle.bfa_align(tree, s, T_o, prm, no_early_exit=True, threads=8)
    bfa.alignClouds()
    assert np.array_equal(bfa.getBestTransformation(), T2_o)
    # with a reachable threshold: success, first candidate below it in loop order
    thr = float(np.sort(bfa.last_scores)[5])
    bfa2 = ssf_gpu.BruteForceAlignment()
    bfa2.setXYZRange(prm.x_range, prm.y_range, prm.z_range)
    bfa2.setMeanErrorThreshold(thr)
    bfa2.setInitialGuess(T_o)
    bfa2.setSourceCloud(s)
    bfa2.setTargetCloud(m)
    prm.mean_error_threshold = thr
    ok3_o, T3_o, _, _ = oracle.bfa_align(tree, s, T_o, prm, threads=8)
    assert bfa2.alignClouds() and ok3_o and bfa2.firstAlignmentCompleted()
    assert np.array_equal(bfa2.getBestTransformation(), T3_o)
