"""GPU parity of the batched geometric verification against the CPU oracle (f64: 1e-6 relative per
BASELINE north_star; asserted much tighter here) and exact equality of RANSAC inlier sets."""
import os

import numpy as np
import pytest

import cv_b200
from oracle import pyoracle as O
from tests.common import GOLDEN
from tests.geom_util import pnp_scene, rot_from_euler, rot_from_scaled_axis, two_view_scene, unit, world_homog

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _pose_list(poses, cnt, h):
    return [(poses[h, k]["r"].reshape(3, 3), poses[h, k]["t"]) for k in range(cnt[h])]


def test_eight_point_batch_matches_oracle():
    rng = np.random.default_rng(0)
    R, t, a, b, _ = two_view_scene(rng, 200, noise=1e-4)
    samples = np.stack([rng.choice(200, 8, replace=False) for _ in range(300)]).astype(np.uint32)
    poses, cnt = cv_b200.EightPoint().estimate_batch(a, b, samples)
    for h in range(len(samples)):
        want = O.eight_point(a[samples[h]], b[samples[h]])
        got = _pose_list(poses, cnt, h)
        assert len(got) == len(want) == 4
        for (Rg, tg), (Rw, tw) in zip(got, want):
            assert np.allclose(Rg, Rw, atol=TOL) and np.allclose(tg, tw, atol=TOL)


def test_p3p_batch_matches_oracle():
    rng = np.random.default_rng(1)
    R, t, bearings, world, _ = pnp_scene(rng, 150)
    samples = np.stack([rng.choice(150, 3, replace=False) for _ in range(400)]).astype(np.uint32)
    poses, cnt = cv_b200.LambdaTwist().estimate_batch(bearings, world, samples)
    hit = 0
    for h in range(len(samples)):
        want = O.p3p(bearings[samples[h]], world[samples[h]])
        got = _pose_list(poses, cnt, h)
        assert len(got) == len(want)
        for (Rg, tg), (Rw, tw) in zip(got, want):
            assert np.allclose(Rg, Rw, atol=1e-8) and np.allclose(tg, tw, atol=1e-8)
        hit += any(np.allclose(Rg, R, atol=1e-6) and np.allclose(tg, t, atol=1e-6) for Rg, tg in got)
    assert hit > 380


def test_residual_matrices_match_oracle():
    rng = np.random.default_rng(2)
    R, t, a, b, _ = two_view_scene(rng, 300, outlier_frac=0.3, noise=1e-3)
    poses = [(R, t), (rot_from_euler(0.3, -0.2, 0.1), unit(rng.standard_normal(3))), (np.eye(3), np.array([1.0, 0, 0]))]
    got = cv_b200.residuals_camera_to_camera(poses, a, b)
    want = np.array([[O.residual_c2c(Rp, tp, a[i], b[i]) for i in range(300)] for Rp, tp in poses])
    assert np.allclose(got, want, rtol=1e-6, atol=1e-12)
    R, t, bearings, world, _ = pnp_scene(rng, 300, outlier_frac=0.2, noise=1e-3)
    poses = [(R, t), (rot_from_euler(0.1, 0.1, -0.3), np.array([0.3, -0.1, 0.2]))]
    got = cv_b200.residuals_world_to_camera(poses, bearings, world)
    want = np.array([[O.residual_w2c(Rp, tp, bearings[i], world[i]) for i in range(300)] for Rp, tp in poses])
    assert np.allclose(got, want, rtol=1e-9, atol=1e-15)


def test_triangulation_matches_oracle_and_doc_test():
    tri = cv_b200.LinearEigenTriangulator()
    p = np.array([0.3, 0.1, 2.0])
    R = rot_from_scaled_axis([0.1, 0.1, 0.1]); t = np.array([0.1, 0.1, 0.1])
    out = tri.triangulate_observations([((np.eye(3), np.zeros(3)), unit(p)), ((R, t), unit(R @ p + t))])
    assert out is not None and np.linalg.norm(out[:3] / out[3] - p) < 1e-6      # cv-geom/src/triangulation.rs:26-38
    rng = np.random.default_rng(3)
    poses, bearings, offsets, pts = [], [], [0], []
    for l in range(500):
        P = np.array([rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(3, 9)])
        k = int(rng.integers(1, 6))
        for _ in range(k):
            Rv = rot_from_euler(*rng.uniform(-0.2, 0.2, 3)); tv = rng.uniform(-0.5, 0.5, 3)
            bv = unit(Rv @ P + tv + rng.normal(0, 1e-3, 3))
            if l % 17 == 0:
                bv = -bv                       # behind the camera -> cheirality failure
            poses.append((Rv, tv)); bearings.append(bv)
        offsets.append(len(poses)); pts.append(P)
    got, ok = tri.triangulate_batch(poses, np.array(bearings), offsets)
    for l in range(500):
        want = O.triangulate_linear_eigen(poses[offsets[l]:offsets[l + 1]], np.array(bearings[offsets[l]:offsets[l + 1]]))
        assert ok[l] == (want is not None), l
        if want is not None:
            assert np.allclose(got[l], want, atol=1e-9)
    assert ok.sum() > 300 and (~ok).sum() > 20


def test_rng_streams_match_oracle():
    x, ox = cv_b200.Xoshiro256PlusPlus(0), O.rng_xoshiro(0)
    assert [x.next_u32() for _ in range(64)] == [O.rng_next_u32(ox) for _ in range(64)]
    p, op = cv_b200.Pcg64(bytes([1] * 32)), O.rng_pcg64(bytes([1] * 32))
    assert [p.next_u32() for _ in range(64)] == [O.rng_next_u32(op) for _ in range(64)]


def test_lambda_twist_consensus_reference_tests():
    # lambda-twist/tests/consensus.rs:18-66
    cam = np.array([[-0.228125, -0.061458334, 1.0], [0.41875, -0.58125, 2.0], [1.128125, 0.878125, 3.0],
                    [-0.528125, 0.178125, 2.5], [-0.923424, -0.235125, 2.8]])
    R = rot_from_euler(0.1, 0.2, 0.3); t = np.array([0.1, 0.2, 0.3])
    world = world_homog((cam - t) @ R)
    bearings = unit(np.concatenate([cam[:, :2] / cam[:, 2:3], np.ones((5, 1))], 1))
    ars = cv_b200.Arrsac(0.01, cv_b200.Xoshiro256PlusPlus(0))
    Rp, tp = ars.model(cv_b200.LambdaTwist(), bearings, world)
    assert np.allclose(Rp, R, atol=1e-6) and np.allclose(tp, t, atol=1e-6)


def test_kitti_estimate_pose_reference_test():
    # akaze/tests/estimate_pose.rs:63-75: Arrsac::new(0.1, Pcg64::from_seed([1; 32])) + EightPoint -> 11 inliers
    g = np.load(os.path.join(GOLDEN, "oracle_kitti_sparse.npz"))
    idx, dist = g["knn_idx"], g["knn_dist"]
    sel = np.where(dist[:, 0].astype(np.float32) < dist[:, 1].astype(np.float32) * np.float32(0.5))[0]
    fx, fy, cx, cy = 9.842439e2, 9.808141e2, 6.9e2, 2.331966e2
    a = np.array([O.calibrate(fx, fy, cx, cy, 0.0, float(g["kps0"][i]["x"]), float(g["kps0"][i]["y"])) for i in sel])
    b = np.array([O.calibrate(fx, fy, cx, cy, 0.0, float(g["kps14"][idx[i, 0]]["x"]), float(g["kps14"][idx[i, 0]]["y"])) for i in sel])
    out = cv_b200.Arrsac(0.1, cv_b200.Pcg64(bytes([1] * 32))).model_inliers(cv_b200.EightPoint(), a, b)
    assert out is not None and len(out[2]) == 11
    want = O.arrsac(O.arrsac_cfg(0.1), 0, a, b, O.rng_pcg64(bytes([1] * 32)))
    assert np.array_equal(out[2], want[2]) and np.allclose(out[0], want[0], atol=1e-9) and np.allclose(out[1], want[1], atol=1e-9)


@pytest.mark.parametrize("n,outliers,seed", [(40, 0.0, 0), (300, 0.3, 1), (1000, 0.5, 2)])
def test_arrsac_eight_point_matches_oracle(n, outliers, seed):
    rng = np.random.default_rng(10 + seed)
    R, t, a, b, good = two_view_scene(rng, n, outlier_frac=outliers, noise=2e-4)
    thr = 1e-6
    got = cv_b200.Arrsac(thr, cv_b200.Xoshiro256PlusPlus(seed)).model_inliers(cv_b200.EightPoint(), a, b)
    want = O.arrsac(O.arrsac_cfg(thr), 0, a, b, O.rng_xoshiro(seed))
    assert (got is None) == (want is None)
    assert np.array_equal(got[2], want[2])                       # identical inlier index sets
    assert np.allclose(got[0], want[0], atol=1e-9) and np.allclose(got[1], want[1], atol=1e-9)
    assert good[got[2]].mean() > 0.95 and len(got[2]) > 0.5 * good.sum()


@pytest.mark.parametrize("n,outliers,seed", [(30, 0.0, 0), (500, 0.2, 1), (2000, 0.4, 2)])
def test_arrsac_p3p_matches_oracle(n, outliers, seed):
    rng = np.random.default_rng(20 + seed)
    R, t, bearings, world, good = pnp_scene(rng, n, outlier_frac=outliers, noise=3e-4)
    thr = 1e-5                                                   # cv-sfm single-view threshold (settings.rs:352-355)
    ars = cv_b200.Arrsac(thr, cv_b200.Xoshiro256PlusPlus(seed)).initialization_hypotheses(512).max_candidate_hypotheses(128)
    got = ars.model_inliers(cv_b200.LambdaTwist(), bearings, world)
    want = O.arrsac(O.arrsac_cfg(thr, initialization_hypotheses=512, max_candidate_hypotheses=128), 1, bearings, world, O.rng_xoshiro(seed))
    assert np.array_equal(got[2], want[2])
    assert np.allclose(got[0], want[0], atol=1e-8) and np.allclose(got[1], want[1], atol=1e-8)
    assert np.allclose(got[0], R, atol=5e-3) and np.allclose(got[1], t, atol=2e-2)
    assert good[got[2]].mean() > 0.95


def test_arrsac_vslam_sandbox_configuration():
    # vslam-sandbox/src/main.rs:112-117: two-view Arrsac(1e-7).initialization_hypotheses(8192).max_candidate_hypotheses(1024)
    rng = np.random.default_rng(5)
    R, t, a, b, good = two_view_scene(rng, 1500, outlier_frac=0.3, noise=5e-5)
    ars = cv_b200.Arrsac(1e-7, cv_b200.Xoshiro256PlusPlus(0)).initialization_hypotheses(8192).max_candidate_hypotheses(1024)
    got = ars.model_inliers(cv_b200.EightPoint(), a, b)
    assert got is not None
    # properties that do not need the (slow) CPU oracle at this size: inliers are true matches, pose agrees with ground truth
    assert good[got[2]].mean() > 0.97 and len(got[2]) > 300
    from tests.geom_util import rot_angle
    assert rot_angle(got[0], R) < 5e-3 and 1.0 - abs(unit(got[1]) @ t) < 1e-3
    res = cv_b200.residuals_camera_to_camera([(got[0], got[1])], a, b)[0]
    assert np.array_equal(np.where(res < 1e-7)[0], got[2])      # model_inliers == { i : residual < threshold }


def test_estimate_pose_end_to_end_reference_test():
    """akaze/tests/estimate_pose.rs:24-76 end to end on the GPU path: extract both KITTI frames, Lowe-ratio match,
    calibrate, ARRSAC + eight-point: 399 / 343 descriptors, 11 matches, 11 inliers."""
    from tests.common import kitti_frame
    ak = cv_b200.Akaze.sparse()
    kps1, ds1 = ak.extract_from_gray_float_image(kitti_frame("0000000000"))
    kps2, ds2 = ak.extract_from_gray_float_image(kitti_frame("0000000014"))
    assert len(ds1) == 399 and len(ds2) == 343
    pairs = cv_b200.lowe_ratio_matches(ds1, ds2, 0.5)
    assert len(pairs) == 11
    K = cv_b200.CameraIntrinsics(focals=(9.842439e2, 9.808141e2), principal_point=(6.9e2, 2.331966e2), skew=0.0)
    a = K.calibrate_keypoints(kps1[[p[0] for p in pairs]])
    b = K.calibrate_keypoints(kps2[[p[1] for p in pairs]])
    assert np.allclose(a[0], O.calibrate(9.842439e2, 9.808141e2, 6.9e2, 2.331966e2, 0.0, float(kps1[pairs[0][0]]["x"]), float(kps1[pairs[0][0]]["y"])), atol=1e-15)
    out = cv_b200.Arrsac(0.1, cv_b200.Pcg64(bytes([1] * 32))).model_inliers(cv_b200.EightPoint(), a, b)
    assert out is not None and len(out[2]) == 11
    px = K.uncalibrate(a)
    assert np.abs(px[:, 0] - kps1[[p[0] for p in pairs]]["x"]).max() < 1e-6      # cv-pinhole/src/lib.rs:120-133


@pytest.mark.parametrize("corrected", [False, True])
def test_five_point_batch_matches_oracle(corrected):
    rng = np.random.default_rng(31)
    R, t, a, b, _ = two_view_scene(rng, 120, noise=0.0)
    samples = np.stack([rng.choice(120, 5, replace=False) for _ in range(200)]).astype(np.uint32)
    poses, cnt = cv_b200.NisterStewenius(corrected=corrected).estimate_batch(a, b, samples)
    O.five_point_set_row0(6 if corrected else 5)
    try:
        hit = 0
        for h in range(len(samples)):
            want = O.five_point(a[samples[h]], b[samples[h]])
            assert cnt[h] == len(want), h
            for k, (Rw, tw) in enumerate(want):
                assert np.allclose(poses[h, k]["r"].reshape(3, 3), Rw, atol=1e-7) and np.allclose(poses[h, k]["t"], tw, atol=1e-7)
            hit += any(np.allclose(poses[h, k]["r"].reshape(3, 3), R, atol=1e-6) and 1 - poses[h, k]["t"] @ t < 1e-9 for k in range(cnt[h]))
        assert (hit > 180) if corrected else (hit == 0)      # the reference's off-by-one rows never recover the pose
    finally:
        O.five_point_set_row0(5)


def test_arrsac_five_point_corrected_matches_oracle():
    rng = np.random.default_rng(41)
    R, t, a, b, good = two_view_scene(rng, 400, outlier_frac=0.3, noise=1e-4)
    thr = 1e-6
    got = cv_b200.Arrsac(thr, cv_b200.Xoshiro256PlusPlus(3)).model_inliers(cv_b200.NisterStewenius(corrected=True), a, b)
    O.five_point_set_row0(6)
    try:
        want = O.arrsac(O.arrsac_cfg(thr), 2, a, b, O.rng_xoshiro(3))
    finally:
        O.five_point_set_row0(5)
    assert got is not None and want is not None
    assert np.array_equal(got[2], want[2])
    assert np.allclose(got[0], want[0], atol=1e-7) and np.allclose(got[1], want[1], atol=1e-7)
    assert good[got[2]].mean() > 0.95 and len(got[2]) > 100
