"""Pins the geometry oracle (oracle/ref_geom.c) against the reference's own tests (CPU only).

  eight-point/tests/random.rs:14-36            >= 950 of 1000 random 16-point scenes have every essential residual < 1e-4
  cv-pinhole/src/essential.rs:93-113,197-216   pose recovery from an essential matrix (angle / translation residual < 1e-4)
  lambda-twist/tests/consensus.rs:18-66        ARRSAC + P3P on 5 exact points returns the ground-truth pose to 1e-6
  lambda-twist/tests/consensus.rs:68-134       degenerate 9-sample case terminates and returns Some
  cv-geom/src/triangulation.rs:26-38           LinearEigenTriangulator recovers (0.3, 0.1, 2.0) to 1e-6
  cv-pinhole/src/lib.rs:91-107                 calibrate / uncalibrate round trip < 0.1 px
  akaze/tests/estimate_pose.rs:63-75           ARRSAC(0.1, Pcg64 [1;32]) + eight-point on the 11 KITTI matches: 11 inliers
"""
import os

import numpy as np

from oracle import pyoracle as O
from tests.common import GOLDEN
from tests.geom_util import rot_angle, rot_from_euler, rot_from_scaled_axis, skew, two_view_scene, unit, world_homog


def test_sym_eigen_matches_numpy():
    rng = np.random.default_rng(0)
    for n in (3, 4, 9):
        for _ in range(20):
            M = rng.standard_normal((n, n)); A = M @ M.T
            ok, d, V = O.sym_eigen(A)
            assert ok
            assert np.allclose(np.sort(d), np.linalg.eigvalsh(A), rtol=1e-10, atol=1e-10)
            assert np.allclose(A @ V, V * d, atol=1e-9)
            assert np.allclose(V.T @ V, np.eye(n), atol=1e-12)


def test_round_robin_jacobi_matches_numpy():
    # the eight-point estimator's 9x9 eigensolver (tournament order, quotient-free rotation), incl. the rank-8 Gram matrices it meets
    rng = np.random.default_rng(5)
    for trial in range(40):
        M = rng.standard_normal((9, 9 if trial % 2 else 8)); A = M @ M.T
        ok, d, V = O.sym_eigen9_rr(A)
        assert ok
        assert np.allclose(np.sort(d), np.linalg.eigvalsh(A), rtol=1e-10, atol=1e-10)
        assert np.allclose(A @ V, V * d, atol=1e-9)
        assert np.allclose(V.T @ V, np.eye(9), atol=1e-12)
        ok2, d2, V2 = O.sym_eigen(A)
        i, j = np.argmin(d), np.argmin(d2)
        assert abs(abs(V[:, i] @ V2[:, j]) - 1.0) < 1e-9        # same null direction as the cyclic order


def test_eight_point_randomized_reference_test():
    # eight-point/tests/random.rs: Vector3::new_random() is uniform [0,1) per component
    rng = np.random.default_rng(1)
    successes = 0
    for _ in range(1000):
        R = rot_from_scaled_axis(rng.random(3) * np.pi * 2.0 * 0.2)
        t = rng.random(3)
        A = rng.random((16, 3)) * 2.0
        A[:, 0] -= 1.0; A[:, 1] -= 1.0; A[:, 2] += 3.0
        B = A @ R.T + t
        a, b = unit(A), unit(B)
        E = O.eight_point_essential(a[:8], b[:8])
        assert E is not None
        successes += all(abs(O.essential_residual(E, a[i], b[i])) <= 1e-4 for i in range(16))
    assert successes > 950, successes


def test_essential_pose_recovery_doc_test():
    R = rot_from_euler(0.2, 0.3, 0.4)
    t = np.array([-0.8, 0.4, 0.5])
    E = skew(t) @ R
    poses = O.essential_poses(E, 1e-6, 50)
    assert len(poses) == 4
    assert any(rot_angle(Rp, R) < 1e-4 and 1.0 - unit(tp) @ unit(t) < 1e-4 for Rp, tp in poses)
    # the four candidates are (t,Ra) (t,Rb) (-t,Ra) (-t,Rb)  (essential.rs:222-229)
    assert np.allclose(poses[0][1], -poses[2][1]) and np.allclose(poses[0][0], poses[2][0])
    assert np.allclose(poses[1][0], poses[3][0]) and np.allclose(poses[0][1], poses[1][1])
    for Rp, _ in poses:
        assert abs(np.linalg.det(Rp) - 1.0) < 1e-9


def _arrsac_manual_samples():
    cam = np.array([[-0.228125, -0.061458334, 1.0], [0.41875, -0.58125, 2.0], [1.128125, 0.878125, 3.0],
                    [-0.528125, 0.178125, 2.5], [-0.923424, -0.235125, 2.8]])
    R = rot_from_euler(0.1, 0.2, 0.3)
    t = np.array([0.1, 0.2, 0.3])
    world = (cam - t) @ R            # pose.inverse() * p
    image = cam[:, :2] / cam[:, 2:3]
    bearings = unit(np.concatenate([image, np.ones((5, 1))], 1))
    return R, t, bearings, world_homog(world)


def test_lambda_twist_arrsac_manual_reference_test():
    R, t, bearings, world = _arrsac_manual_samples()
    out = O.arrsac(O.arrsac_cfg(0.01), 1, bearings, world, O.rng_xoshiro(0))
    assert out is not None
    Rp, tp, inl = out
    assert np.allclose(Rp, R, atol=1e-6) and np.allclose(tp, t, atol=1e-6)
    assert inl.tolist() == [0, 1, 2, 3, 4]
    # P3P itself: one of the <= 4 candidates from the first three samples is the true pose
    cands = O.p3p(bearings[:3], world[:3])
    assert 1 <= len(cands) <= 4
    assert any(np.allclose(Rc, R, atol=1e-6) and np.allclose(tc, t, atol=1e-6) for Rc, tc in cands)


def test_lambda_twist_endless_loop_case_terminates():
    img = [(0.3070512144698557, 0.19317668016026052), (0.3208462966353674, 0.20741702947913013),
           (0.3070512144698557, 0.19317668016026052), (0.3208462966353674, 0.20741702947913013),
           (0.3208462966353674, 0.20741702947913013), (0.3070512144698557, 0.19317668016026052),
           (0.26619553978146293, 0.15033756455213498), (0.3494806979265859, 0.18264329458710366),
           (0.32132193890323213, 0.15408143785084824)]
    pts = [(1.0, 1.0, 0.0), (1.0, 1.5, 0.0), (3.0, 1.0, 0.0), (1.0, 2.0, 0.0), (2.0, 2.0, 0.0), (3.0, 2.0, 0.0),
           (1.0, 3.0, 0.0), (2.0, 3.0, 0.0), (3.0, 3.0, 0.0)]
    bearings = unit(np.array([[x, y, 1.0] for x, y in img]))
    out = O.arrsac(O.arrsac_cfg(0.01), 1, bearings, world_homog(np.array(pts)), O.rng_xoshiro(0))
    assert out is not None


def test_linear_eigen_triangulator_doc_test():
    p = np.array([0.3, 0.1, 2.0])
    R = rot_from_scaled_axis([0.1, 0.1, 0.1]); t = np.array([0.1, 0.1, 0.1])
    a = unit(p); b = unit(R @ p + t)
    out = O.triangulate_linear_eigen([(np.eye(3), np.zeros(3)), (R, t)], np.stack([a, b]))
    assert out is not None
    assert np.linalg.norm(out[:3] / out[3] - p) < 1e-6
    # a point behind the cameras violates cheirality -> None (triangulation.rs:120-127)
    assert O.triangulate_linear_eigen([(np.eye(3), np.zeros(3)), (R, t)], np.stack([-a, -b])) is None
    assert O.triangulate_linear_eigen([(np.eye(3), np.zeros(3))], a[None]) is None


def test_calibrate_doc_test():
    b = O.calibrate(800.0, 900.0, 500.0, 600.0, 1.7, 471.0, 322.0)
    K = np.array([[800.0, 1.7, 500.0], [0, 900.0, 600.0], [0, 0, 1]])
    u = K @ (b / b[2])
    assert np.linalg.norm(u[:2] / u[2] - np.array([471.0, 322.0])) < 0.1


def test_residuals_zero_for_exact_data_and_two_for_failure():
    rng = np.random.default_rng(5)
    R, t, a, b, _ = two_view_scene(rng, 50)
    r = [O.residual_c2c(R, t, a[i], b[i]) for i in range(50)]
    assert max(r) < 1e-12
    wrong = [O.residual_c2c(rot_from_euler(0.5, -0.4, 0.3), -t, a[i], b[i]) for i in range(50)]
    assert np.median(wrong) > 1e-4


def test_rng_streams():
    # xoshiro256++ seeded by SplitMix64(0): first outputs of the reference generator (rand_xoshiro test vectors' scheme)
    r = O.rng_xoshiro(0)
    assert list(r.s) == [0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4, 0x06c45d188009454f, 0xf88bb8a8724c81ec]
    v = [O.rng_next_u32(r) for _ in range(4)]
    assert len(set(v)) == 4
    p = O.rng_pcg64(bytes([1] * 32))
    w = [O.rng_next_u32(p) for _ in range(1000)]
    assert len(set(w)) > 990


def test_kitti_estimate_pose_reference_test():
    # akaze/tests/estimate_pose.rs:27-75 with the committed oracle vectors
    g = np.load(os.path.join(GOLDEN, "oracle_kitti_sparse.npz"))
    idx, dist = g["knn_idx"], g["knn_dist"]
    sel = np.where(dist[:, 0].astype(np.float32) < dist[:, 1].astype(np.float32) * np.float32(0.5))[0]
    assert len(sel) == 11
    fx, fy, cx, cy = 9.842439e2, 9.808141e2, 6.9e2, 2.331966e2
    a = np.array([O.calibrate(fx, fy, cx, cy, 0.0, float(g["kps0"][i]["x"]), float(g["kps0"][i]["y"])) for i in sel])
    b = np.array([O.calibrate(fx, fy, cx, cy, 0.0, float(g["kps14"][idx[i, 0]]["x"]), float(g["kps14"][idx[i, 0]]["y"])) for i in sel])
    out = O.arrsac(O.arrsac_cfg(0.1), 0, a, b, O.rng_pcg64(bytes([1] * 32)))
    assert out is not None and len(out[2]) == 11


def test_five_point_restatement_and_reference_quirk():
    """nister-stewenius has no end-to-end test upstream (tests/manual.rs is commented out).  Every step of the
    restatement is validated here with the mathematically correct eigenvector rows (6..9): the true essential matrix
    is among the solutions.  With the reference's rows 5..8 (lib.rs:229) it never is -- the quirk is reproduced."""
    rng = np.random.default_rng(7)
    for _ in range(50):
        A = rng.standard_normal((10, 10))
        ok, ev = O.real_eigenvalues10(A)
        assert ok and np.allclose(np.sort_complex(ev), np.sort_complex(np.linalg.eigvals(A)), atol=1e-9)
    scenes = [two_view_scene(rng, 5) for _ in range(100)]

    def hits():
        h = 0
        for R, t, a, b, _ in scenes:
            Et = skew(t) @ R
            Et /= np.linalg.norm(Et)
            Es = O.five_point_essentials(a, b)
            assert len(Es) <= 10
            for E in Es:                                   # every solution lies in the epipolar null space
                assert max(abs(b[k] @ E @ a[k]) for k in range(5)) < 1e-9 * np.linalg.norm(E)
            h += any(min(np.abs(E / np.linalg.norm(E) - Et).max(), np.abs(E / np.linalg.norm(E) + Et).max()) < 1e-6 for E in Es)
        return h
    try:
        O.five_point_set_row0(6)
        assert hits() >= 95
        R, t, a, b, _ = scenes[0]
        assert any(rot_angle(Rp, R) < 1e-6 and 1 - unit(tp) @ t < 1e-9 for Rp, tp in O.five_point(a, b))
        O.five_point_set_row0(5)
        assert hits() == 0
    finally:
        O.five_point_set_row0(5)


# ---- the two unit tests the reference holds for the five-point solver's polynomial products (nister-stewenius/src/lib.rs:336-417)
_BASIS = dict(XXX=0, XXY=1, XYY=2, YYY=3, XXZ=4, XYZ=5, YYZ=6, XZZ=7, YZZ=8, ZZZ=9, XX=10, XY=11, YY=12, XZ=13, YZ=14, ZZ=15, X=16, Y=17, Z=18, ONE=19)


def _eval_polynomial(p, x, y, z):
    B = _BASIS
    return (p[B["XXX"]] * x * x * x + p[B["XXY"]] * x * x * y + p[B["XXZ"]] * x * x * z + p[B["XYY"]] * x * y * y + p[B["XYZ"]] * x * y * z
            + p[B["XZZ"]] * x * z * z + p[B["YYY"]] * y * y * y + p[B["YYZ"]] * y * y * z + p[B["YZZ"]] * y * z * z + p[B["ZZZ"]] * z * z * z
            + p[B["XX"]] * x * x + p[B["XY"]] * x * y + p[B["XZ"]] * x * z + p[B["YY"]] * y * y + p[B["YZ"]] * y * z + p[B["ZZ"]] * z * z
            + p[B["X"]] * x + p[B["Y"]] * y + p[B["Z"]] * z + p[B["ONE"]])


def _vec_to_poly_basis(v):
    p = np.zeros(20)
    p[_BASIS["X"]], p[_BASIS["Y"]], p[_BASIS["Z"]], p[_BASIS["ONE"]] = v
    return p


def test_o1_manual_reference_unit_test():
    p1, p2 = np.array([0.1, 0.8, 0.3, 0.2]), np.array([0.5, 0.45, 0.82, 0.15])
    p3 = O.fp_o1(p1, p2)
    for z in range(-5, 5):
        for y in range(-5, 5):
            for x in range(-5, 5):
                want = _eval_polynomial(_vec_to_poly_basis(p1), x, y, z) * _eval_polynomial(_vec_to_poly_basis(p2), x, y, z)
                assert abs(_eval_polynomial(p3, x, y, z) - want) < 1e-6


def test_o2_manual_reference_unit_test():
    p1 = np.zeros(20)
    for name, v in (("XX", 0.2), ("XY", 0.81), ("XZ", 0.66), ("YY", 0.91), ("YZ", 0.88), ("ZZ", 0.14), ("X", 0.97), ("Y", 0.3), ("Z", 0.38), ("ONE", 0.72)):
        p1[_BASIS[name]] = v
    p2 = np.array([0.5, 0.45, 0.82, 0.15])
    p3 = O.fp_o2(p1, p2)
    for z in range(-5, 5):
        for y in range(-5, 5):
            for x in range(-5, 5):
                want = _eval_polynomial(p1, x, y, z) * _eval_polynomial(_vec_to_poly_basis(p2), x, y, z)
                assert abs(_eval_polynomial(p3, x, y, z) - want) < 1e-8
