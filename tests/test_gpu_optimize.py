"""GPU parity for the post-consensus steps: cv_b200.optimize vs the CPU port.
The per-iteration sums are added by a fixed tree on the GPU and in landmark order on the CPU, so poses agree to
rounding (bar: 1e-6 relative; observed ~1e-12) and the patience rule stops within a few iterations of each other."""
import numpy as np
import pytest

import cv_b200
from oracle import pyoracle as O
from tests.geom_util import perturb_pose, pnp_scene, rot_from_scaled_axis, three_view_scene, unit

pytestmark = pytest.mark.gpu


def _close(p, q, tol=1e-9):
    return np.allclose(p[0], q[0], atol=tol) and np.allclose(p[1], q[1], atol=tol)


def test_single_view_l2_batch_matches_cpu_port():
    rng = np.random.default_rng(0)
    poses, B, W, off, truth = [], [], [], [0], []
    for k, n in enumerate([300, 2048, 1, 0, 37, 1000]):
        R, t, bearings, world, _ = pnp_scene(rng, n, noise=2e-4) if n else (np.eye(3), np.zeros(3), np.zeros((0, 3)), np.zeros((0, 4)), None)
        poses.append(perturb_pose(rng, (R, t), 2e-3, 5e-3)); B.append(bearings); W.append(world); off.append(off[-1] + n)
    B = np.concatenate(B); W = np.concatenate(W)
    for iters in (1, 150, 4000):
        got, upd = cv_b200.single_view_simple_optimize_l2_batch(poses, 1e-3, iters, B, W, off)
        for k in range(len(poses)):
            Rw, tw, uw = O.single_view_optimize_l2(poses[k], 1e-3, iters, B[off[k]:off[k + 1]], W[off[k]:off[k + 1]])
            assert abs(int(upd[k]) - uw) <= 3, (k, iters, upd[k], uw)
            assert _close(got[k], (Rw, tw), 1e-8), (k, iters)
    # single-problem surface of the reference
    one = cv_b200.single_view_simple_optimize_l2(poses[0], 1e-3, 150, (B[:300], W[:300]))
    Rw, tw, _ = O.single_view_optimize_l2(poses[0], 1e-3, 150, B[:300], W[:300])
    assert _close(one, (Rw, tw))
    assert cv_b200.single_view_simple_optimize_l2(poses[0], 1e-3, 10, (np.zeros((0, 3)), np.zeros((0, 4)))) is poses[0]


@pytest.mark.parametrize("adaptive", [False, True])
def test_three_view_l2_batch_matches_cpu_port(adaptive):
    rng = np.random.default_rng(1)
    starts, obs, off = [], [], [0]
    for n in (200, 1024, 3, 0):
        truth, o = three_view_scene(rng, max(n, 1), noise=1e-4)
        o = o[:n]
        starts.append([perturb_pose(rng, p, 3e-3, 5e-3) for p in truth]); obs.append(o); off.append(off[-1] + n)
    obs_all = np.concatenate(obs)
    iters = 120 if adaptive else 1500
    got, upd = cv_b200.three_view_optimize_l2_batch(starts, 1e-3, iters, obs_all, off, adaptive=adaptive)
    for k in range(len(starts)):
        want, uw = O.three_view_optimize_l2(starts[k], 1e-3, iters, obs[k], adaptive=adaptive)
        assert abs(int(upd[k]) - uw) <= 3, (k, upd[k], uw)
        for v in range(2):
            assert _close(got[k][v], want[v], 1e-8), (k, v)
    if adaptive:
        single = cv_b200.three_view_adaptive_optimize_l2(starts[0], iters, obs[0])
    else:
        single = cv_b200.three_view_simple_optimize_l2(starts[0], 1e-3, iters, obs[0])
    want, _ = O.three_view_optimize_l2(starts[0], 1e-3, iters, obs[0], adaptive=adaptive)
    assert _close(single[0], want[0], 1e-8) and _close(single[1], want[1], 1e-8)


def test_observation_losses_match_cpu_port():
    rng = np.random.default_rng(2)
    poses, bearings, off = [], [], [0]
    for l in range(600):
        X = np.array([rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(3, 9)])
        k = int(rng.integers(1, 7))
        for _ in range(k):
            R = rot_from_scaled_axis(rng.uniform(-1, 1, 3) * 0.2); t = rng.uniform(-0.6, 0.6, 3)
            b = unit(R @ X + t + rng.normal(0, 2e-3, 3))
            if l % 19 == 0:
                b = -b
            poses.append((R, t)); bearings.append(b)
        off.append(len(poses))
    bearings = np.array(bearings)
    got = cv_b200.observation_losses(poses, bearings, off)
    for l in range(600):
        want = O.observation_losses(poses[off[l]:off[l + 1]], bearings[off[l]:off[l + 1]])
        assert np.allclose(got[off[l]:off[l + 1]], want, rtol=1e-9, atol=1e-14), l
    assert (got == 2.0).sum() > 50 and (got < 1e-4).sum() > 500


def test_tri_landmarks_robust_match_cpu_port():
    rng = np.random.default_rng(3)
    poses, obs = three_view_scene(rng, 800, noise=1.5e-3)
    counts = []
    for max_cos, inc in ((1e-5, 1e-6), (1e-6, 1e-3), (1e-4, 5e-3), (1e-6, 2e-2)):
        got = cv_b200.tri_landmarks_robust(poses[0], poses[1], obs, max_cos, inc)
        want = np.array([O.is_tri_landmark_robust(poses[0], poses[1], o[0], o[1], o[2], max_cos, inc) for o in obs])
        assert np.array_equal(got, want)
        counts.append(int(got.sum()))
    assert sum(0 < c < len(obs) for c in counts) >= 2, counts      # the thresholds really split the set
