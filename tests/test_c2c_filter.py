"""CPU: the exact-predicate filter that replaces most CameraToCamera::residual evaluations inside ARRSAC
(cv_b200/csrc/c2c_filter.cuh, compiled for the host by tests/csrc/c2c_filter_host.c) never contradicts the oracle's exact
evaluation of `residual < threshold` (cv-core/src/pose.rs:249-296), and leaves only a small fraction undecided."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as O
from tests.geom_util import two_view_scene

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def filt():
    out = os.path.join(HERE, "csrc", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libc2c_filter_host.so")
    src = os.path.join(HERE, "csrc", "c2c_filter_host.c")
    hdr = os.path.join(HERE, "..", "cv_b200", "csrc", "c2c_filter.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c", src, "-o", so, "-lm"])
    L = C.CDLL(so)
    L.c2c_filter_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_void_p]
    L.c2c_filter_batch.restype = None
    return L


def _poses(a, b, rng, nh, pool=None):
    out = []
    n = len(a) if pool is None else len(pool)
    for _ in range(nh):
        idx = rng.choice(n, 8, replace=False)
        if pool is not None:
            idx = pool[idx]
        out += O.eight_point(a[idx], b[idx])
    return out


def _check(filt, poses, a, b, thr):
    P = np.ascontiguousarray([np.concatenate([R.reshape(9), t]) for R, t in poses], np.float64)
    out = np.zeros((len(P), len(a)), np.int8)
    filt.c2c_filter_batch(P.ctypes.data, len(P), a.ctypes.data, b.ctypes.data, len(a), thr, out.ctypes.data)
    exact = np.array([[O.residual_c2c(R, t, a[i], b[i]) < thr for i in range(len(a))] for R, t in poses])
    decided = out >= 0
    assert np.array_equal(out[decided] == 1, exact[decided])
    return decided.mean(), exact.mean()


def test_filter_agrees_with_exact_predicate_on_bench_pair(filt):
    z = np.load(os.path.join(HERE, "golden", "bench_pair0.npz"))
    a, b = np.ascontiguousarray(z["ba"][:600]), np.ascontiguousarray(z["bb"][:600])
    rng = np.random.default_rng(0)
    poses = _poses(a, b, rng, 12)
    dec, inl = _check(filt, poses, a, b, 1e-7)
    assert dec > 0.98
    # hypotheses generated from the inliers of a good model: most data sit close to the threshold
    best = max(poses, key=lambda p: sum(O.residual_c2c(p[0], p[1], a[i], b[i]) < 1e-7 for i in range(len(a))))
    pool = np.array([i for i in range(len(a)) if O.residual_c2c(best[0], best[1], a[i], b[i]) < 1e-7])
    assert len(pool) > 100
    dec, inl = _check(filt, _poses(a, b, rng, 12, pool), a, b, 1e-7)
    assert dec > 0.97 and inl > 0.05


@pytest.mark.parametrize("thr,noise", [(1e-7, 5e-5), (1e-6, 2e-4), (1e-9, 0.0), (1e-5, 1e-3)])
def test_filter_agrees_on_synthetic_scenes(filt, thr, noise):
    rng = np.random.default_rng(int(-np.log10(thr)))
    R, t, a, b, good = two_view_scene(rng, 300, outlier_frac=0.3, noise=noise)
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    poses = _poses(a, b, rng, 10) + [(R, t)]
    dec, _ = _check(filt, poses, a, b, thr)
    assert dec > 0.9


def test_filter_declines_large_thresholds_and_degenerate_input(filt):
    rng = np.random.default_rng(3)
    R, t, a, b, _ = two_view_scene(rng, 50)
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    P = np.ascontiguousarray([np.concatenate([R.reshape(9), t])], np.float64)
    out = np.zeros((1, 50), np.int8)
    filt.c2c_filter_batch(P.ctypes.data, 1, a.ctypes.data, b.ctypes.data, 50, 0.1, out.ctypes.data)     # akaze/tests/estimate_pose.rs:63
    assert (out == -1).all()
    Pz = np.ascontiguousarray([np.concatenate([np.eye(3).reshape(9), np.zeros(3)])], np.float64)       # t = 0: rank-deficient design
    filt.c2c_filter_batch(Pz.ctypes.data, 1, a.ctypes.data, a.ctypes.data, 50, 1e-7, out.ctypes.data)
    exact = np.array([O.residual_c2c(np.eye(3), np.zeros(3), a[i], a[i]) < 1e-7 for i in range(50)])
    d = out[0] >= 0
    assert np.array_equal(out[0][d] == 1, exact[d])


def test_epipolar_lower_bound_of_the_residual():
    # the filter's pre-test rests on residual >= (b^T [t]x R a)^2 / (8 |t|^2) for whatever point the reference triangulates;
    # checked here against the exact evaluation on hypotheses of every quality, with matches from perfect to random
    rng = np.random.default_rng(11)
    worst = np.inf
    for trial in range(6):
        R, t, a, b, good = two_view_scene(rng, 200, outlier_frac=0.4, noise=[0.0, 1e-5, 1e-4, 1e-3, 1e-2, 0.1][trial])
        poses = _poses(a, b, rng, 6) + [(R, t), (R, 3.0 * t), (R, 1e-3 * t)]
        for Rm, tm in poses:
            E = np.cross(np.eye(3), tm).T @ Rm if False else np.array([[0, -tm[2], tm[1]], [tm[2], 0, -tm[0]], [-tm[1], tm[0], 0]]) @ Rm
            e = np.einsum("ij,jk,ik->i", b, E, a)
            for i in range(len(a)):
                res = O.residual_c2c(Rm, tm, a[i], b[i])
                bound = e[i] ** 2 / (8.0 * (tm @ tm))
                assert res >= bound * (1 - 1e-9) - 1e-15, (trial, i, res, bound)
                if bound > 1e-12:
                    worst = min(worst, res / bound)
    assert worst >= 1.0 - 1e-9


def test_cheirality_bound_of_the_residual(filt):
    # the second pre-test: residual < thr forces c1 = t.b - (u.b)(t.u) and c2 = (u.b)(t.b) - t.u (u = R a) above -5.66 |t| sqrt(thr);
    # the filter rejects below -10 |t| sqrt(thr).  Checked against the exact evaluation on the four poses of true and estimated
    # essential matrices (three of the four are wrong-cheirality poses with the same epipolar error), and on scaled translations.
    rng = np.random.default_rng(12)
    nrej = ntot = 0
    for trial, thr in enumerate([1e-9, 1e-7, 1e-6, 1e-5, 1e-4, 1e-7]):
        R, t, a, b, good = two_view_scene(rng, 150, outlier_frac=0.3, noise=[0.0, 1e-5, 1e-4, 3e-4, 1e-3, 1e-2][trial])
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        Rt = R @ (2.0 * np.outer(t, t) - np.eye(3))                 # the twisted pair of the true pose (rotation by pi about t)
        poses = _poses(a, b, rng, 5) + [(R, t), (R, -t), (Rt, t), (Rt, -t), (R, 5.0 * t), (R, -1e-2 * t)]
        for Rm, tm in poses:
            u = a @ Rm.T
            tb, ub, tu = b @ tm, np.einsum("ij,ij->i", u, b), u @ tm
            c1, c2 = tb - ub * tu, ub * tb - tu
            lim = 10.0 * np.linalg.norm(tm) * np.sqrt(thr)
            certain = (c1 < -lim) | (c2 < -lim)
            for i in np.nonzero(certain)[0]:
                assert not (O.residual_c2c(Rm, tm, a[i], b[i]) < thr), (trial, i, c1[i], c2[i])
            nrej += int(certain.sum()); ntot += len(a)
        # and the filter as a whole stays consistent with the exact predicate on these poses
        _check(filt, poses, a, b, thr)
    assert nrej > 0.3 * ntot          # the bound does decide the wrong poses (this is what makes it worth its 12 instructions)
