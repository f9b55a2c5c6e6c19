"""CPU checks of the oracle's post-consensus refinement / robustness restatement (oracle/ref_optimize.c) against the
properties the reference's algorithms must have (the reference holds no tests for them)."""
import numpy as np

from oracle import pyoracle as O
from tests.geom_util import perturb_pose, pnp_scene, rot_angle, rot_from_scaled_axis, three_view_scene, unit, world_homog


def _mean_residual(pose, bearings, world):
    return np.mean([O.residual_w2c(pose[0], pose[1], bearings[i], world[i]) for i in range(len(bearings))])


def test_single_view_l2_improves_the_pose_and_obeys_the_caps():
    rng = np.random.default_rng(0)
    R, t, bearings, world, _ = pnp_scene(rng, 300)
    start = perturb_pose(rng, (R, t), 2e-3, 5e-3)
    r0 = _mean_residual(start, bearings, world)
    Rr, tr, upd = O.single_view_optimize_l2(start, 1e-3, 200, bearings, world)
    assert upd == 200                                           # the iteration cap ends the loop (single_view_optimizer.rs:127-131)
    assert _mean_residual((Rr, tr), bearings, world) < r0
    Rr2, tr2, upd2 = O.single_view_optimize_l2(start, 1e-3, 100000, bearings, world)
    assert 200 < upd2 < 100000                                   # the patience rule (50 iterations without a new best) ends it
    assert _mean_residual((Rr2, tr2), bearings, world) < _mean_residual((Rr, tr), bearings, world)
    # the exact pose is a fixed point: all tangents vanish
    Re, te, _ = O.single_view_optimize_l2((R, t), 1e-3, 10, bearings, world)
    assert np.allclose(Re, R, atol=1e-12) and np.allclose(te, t, atol=1e-12)
    # no landmarks -> the pose comes back untouched
    Rn, tn, un = O.single_view_optimize_l2(start, 1e-3, 10, np.zeros((0, 3)), np.zeros((0, 4)))
    assert un == 0 and np.array_equal(Rn, start[0]) and np.array_equal(tn, start[1])


def test_three_view_gradients_vanish_on_consistent_geometry():
    rng = np.random.default_rng(1)
    poses, obs = three_view_scene(rng, 50)
    inv = [(P[0].T, -P[0].T @ P[1]) for P in poses]             # first/second -> centre
    for o in obs:
        g = O.three_view_gradients(o[0], inv[0][0] @ o[1], inv[0][1], inv[1][0] @ o[2], inv[1][1])
        assert np.abs(g).max() < 1e-9
    bad = [perturb_pose(rng, p, 1e-2, 2e-2) for p in inv]
    g = O.three_view_gradients(obs[0][0], bad[0][0] @ obs[0][1], bad[0][1], bad[1][0] @ obs[0][2], bad[1][1])
    assert np.abs(g).max() > 1e-4


def _three_view_error(poses, truth):
    return max(max(rot_angle(p[0], q[0]), np.linalg.norm(unit(p[1]) - unit(q[1]))) for p, q in zip(poses, truth))


def test_three_view_l2_and_adaptive_reduce_the_error():
    rng = np.random.default_rng(2)
    truth, obs = three_view_scene(rng, 200)
    start = [perturb_pose(rng, p, 3e-3, 5e-3) for p in truth]
    e0 = _three_view_error(start, truth)
    out, upd = O.three_view_optimize_l2(start, 1e-3, 3000, obs)
    assert 0 < upd <= 3000 and _three_view_error(out, truth) < e0
    out_a, upd_a = O.three_view_optimize_l2(start, 0.0, 300, obs, adaptive=True)
    assert upd_a == 300 and _three_view_error(out_a, truth) < e0
    same, upd0 = O.three_view_optimize_l2(truth, 1e-3, 5, obs)
    assert _three_view_error(same, truth) < 1e-9


def test_epipolar_loss_and_observation_losses():
    rng = np.random.default_rng(3)
    R = rot_from_scaled_axis([0.05, -0.1, 0.02]); t = np.array([0.5, 0.1, -0.05])
    X = np.array([0.3, -0.2, 5.0])
    a, b = unit(X), unit(R @ X + t)
    assert O.epipolar_loss(t, R @ a, b) < 1e-12                  # is_bi_landmark_robust's call, cv-sfm/src/lib.rs:1313-1316
    assert O.epipolar_loss(t, R @ a, unit(b + np.array([0, 0.05, 0]))) > 1e-3
    assert O.epipolar_loss(t, R @ a, -b) == 1.0                   # failed cheirality
    # observation_loss: 1 observation -> 2.0; 2 -> epipolar loss as a cosine distance; >= 3 -> distance to the triangulated point
    ident = (np.eye(3), np.zeros(3))
    assert O.observation_losses([ident], [a])[0] == 2.0
    two = O.observation_losses([ident, (R, t)], [a, b])
    assert two[0] == two[1] and two[0] < 1e-12
    R2 = rot_from_scaled_axis([-0.02, 0.08, 0.01]); t2 = np.array([-0.4, 0.05, 0.1])
    three = O.observation_losses([ident, (R, t), (R2, t2)], [a, b, unit(R2 @ X + t2)])
    assert np.all(three < 1e-12)
    off = O.observation_losses([ident, (R, t), (R2, t2)], [a, b, unit(R2 @ X + t2 + np.array([0.05, 0, 0]))])
    assert np.all(off > 1e-7) and np.all(off < 1e-2)
    behind = O.observation_losses([ident, (R, t), (R2, t2)], [-a, -b, -unit(R2 @ X + t2)])
    assert np.all(behind == 2.0)                                  # triangulation fails its cheirality rule


def test_tri_landmark_robust():
    rng = np.random.default_rng(4)
    poses, obs = three_view_scene(rng, 20)
    for o in obs:
        assert O.is_tri_landmark_robust(poses[0], poses[1], o[0], o[1], o[2], 1e-5, 1e-6)
        assert not O.is_tri_landmark_robust(poses[0], poses[1], o[0], o[1], o[2], 1e-5, 0.5)     # incidence angle demanded too large
        bad = unit(o[2] + np.array([0.0, 0.02, 0.0]))
        assert not O.is_tri_landmark_robust(poses[0], poses[1], o[0], o[1], bad, 1e-7, 1e-6)     # third ray misses the point
