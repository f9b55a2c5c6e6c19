"""GPU: BASELINE configs[2] at full size and the fused two-view pair pipeline.

  * config 3: the bench's real frame pair (3 811 symmetric matches of two 1080p frames, tests/golden/bench_pair0.npz, made by
    tests/golden/make_bench_pair.py with the CPU oracle) -> Arrsac(1e-7, Xoshiro256++(0)).initialization_hypotheses(8192)
    .max_candidate_hypotheses(1024) + EightPoint (vslam-sandbox/src/main.rs:112-117): inlier set, pose and generator state equal
    to the oracle's; 5 000 x 5 000 random 486-bit descriptors (seed 2) yield no consensus on both sides.
  * the device-resident driver equals the round-1 host-driven driver (CVB_ARRSAC_HOST=1) bit for bit on odd shapes.
  * cvb_two_view_frames == the composition of the separate host API calls."""
import os

import numpy as np
import pytest

import cv_b200
from oracle import pyoracle as O
from tests.common import GOLDEN
from tests.geom_util import pnp_scene, two_view_scene
from tests.synth import random_descriptors, synth_frame, warp_frame

pytestmark = pytest.mark.gpu


def _vslam_two_view(seed=0):
    return cv_b200.Arrsac(1e-7, cv_b200.Xoshiro256PlusPlus(seed)).initialization_hypotheses(8192).max_candidate_hypotheses(1024)


def test_config3_real_pair_inlier_set_equals_oracle():
    z = np.load(os.path.join(GOLDEN, "bench_pair0.npz"))
    a, b = z["ba"], z["bb"]
    assert len(a) == 3811
    ars = _vslam_two_view()
    got = ars.model_inliers(cv_b200.EightPoint(), a, b)
    rng = O.rng_xoshiro(0)
    want = O.arrsac(O.arrsac_cfg(1e-7, initialization_hypotheses=8192, max_candidate_hypotheses=1024), 0, a, b, rng)
    assert got is not None and want is not None
    assert np.array_equal(got[2], want[2]) and len(got[2]) == 3807
    assert np.allclose(got[0], want[0], rtol=1e-6, atol=1e-12) and np.allclose(got[1], want[1], rtol=1e-6, atol=1e-12)
    assert [int(x) for x in ars.rng.state.s] == [int(x) for x in rng.s]        # same number of draws consumed
    # model_inliers == { i : residual < threshold } for the returned model
    res = cv_b200.residuals_camera_to_camera([(got[0], got[1])], a, b)[0]
    assert np.array_equal(np.where(res < 1e-7)[0], got[2])


def test_config3_shuffled_and_outlier_contaminated_pair_equals_oracle():
    """cv-sfm shuffles the matches before consensus (cv-sfm/src/lib.rs:1386); 25 % of them replaced by wrong matches so that
    the block loop generates and accepts new hypotheses for many blocks."""
    z = np.load(os.path.join(GOLDEN, "bench_pair0.npz"))
    rng = np.random.default_rng(11)
    a, b = z["ba"].copy(), z["bb"].copy()
    bad = rng.choice(len(a), len(a) // 4, replace=False)
    b[bad] = b[rng.permutation(bad)]
    perm = rng.permutation(len(a))
    a, b = np.ascontiguousarray(a[perm]), np.ascontiguousarray(b[perm])
    ars = _vslam_two_view(7)
    got = ars.model_inliers(cv_b200.EightPoint(), a, b)
    orng = O.rng_xoshiro(7)
    want = O.arrsac(O.arrsac_cfg(1e-7, initialization_hypotheses=8192, max_candidate_hypotheses=1024), 0, a, b, orng)
    assert (got is None) == (want is None)
    if got is not None:
        assert np.array_equal(got[2], want[2])
        assert np.allclose(got[0], want[0], rtol=1e-6, atol=1e-12) and np.allclose(got[1], want[1], rtol=1e-6, atol=1e-12)
    assert [int(x) for x in ars.rng.state.s] == [int(x) for x in orng.s]


def test_config3_random_descriptors_have_no_consensus():
    d1, d2 = random_descriptors(5000, 2), random_descriptors(5000, 3)
    pairs = cv_b200.symmetric_matching(d1, d2, 24)
    oi, od = O.hamming_knn(d1, d2, 2)
    ri, rd = O.hamming_knn(d2, d1, 2)
    fwd = np.where(od[:, 0] + 24 <= od[:, 1], oi[:, 0].astype(np.int64), -1)
    rev = np.where(rd[:, 0] + 24 <= rd[:, 1], ri[:, 0].astype(np.int64), -1)
    want = [(i, j) for i, j in enumerate(fwd) if j >= 0 and rev[j] == i]
    assert [tuple(p) for p in pairs.tolist()] == want
    K = cv_b200.CameraIntrinsics(focals=(1000.0, 1000.0), principal_point=(960.0, 540.0))
    px = np.random.default_rng(2).uniform(0, 1000, (max(len(pairs), 1), 2))
    a, b = K.calibrate(px)[:len(pairs)], K.calibrate(px[::-1])[:len(pairs)]
    got = _vslam_two_view().model_inliers(cv_b200.EightPoint(), a.reshape(-1, 3), b.reshape(-1, 3))
    want = O.arrsac(O.arrsac_cfg(1e-7, initialization_hypotheses=8192, max_candidate_hypotheses=1024), 0, a.reshape(-1, 3), b.reshape(-1, 3),
                    O.rng_xoshiro(0)) if len(pairs) else None
    assert (got is None) == (want is None)
    if got is not None:
        assert np.array_equal(got[2], want[2])


@pytest.mark.parametrize("n,of,noise,thr,kw", [
    (8, 0.0, 0.0, 0.1, {}),                                      # exactly MIN_SAMPLES
    (11, 0.0, 0.0, 0.1, {}),                                     # heavy rejection sampling, filter disabled (large threshold)
    (70, 0.2, 1e-4, 1e-6, {}),                                   # fewer data than initialization_blocks * block_size
    (300, 0.3, 1e-4, 1e-6, {}),
    (777, 0.5, 1e-4, 1e-6, dict(block_size=50, initialization_blocks=3, estimations_per_block=32)),   # blocks not word aligned
    (1000, 0.3, 5e-5, 1e-7, dict(initialization_hypotheses=512, max_candidate_hypotheses=128)),
    (640, 0.3, 5e-5, 1e-7, dict(estimations_per_block=0)),       # no re-estimation
    (500, 0.95, 1e-4, 1e-7, dict(initialization_hypotheses=64)), # almost no consensus
])
def test_device_driver_equals_host_driver_and_oracle(n, of, noise, thr, kw, monkeypatch):
    rng = np.random.default_rng(n)
    R, t, a, b, _ = two_view_scene(rng, n, outlier_frac=of, noise=noise)

    def run(host):
        monkeypatch.setenv("CVB_ARRSAC_HOST", "1" if host else "0")
        ars = cv_b200.Arrsac(thr, cv_b200.Xoshiro256PlusPlus(3))
        for k, v in kw.items():
            getattr(ars, k)(v)
        return ars.model_inliers(cv_b200.EightPoint(), a, b), [int(x) for x in ars.rng.state.s]
    (gd, sd), (gh, sh) = run(False), run(True)
    orng = O.rng_xoshiro(3)
    want = O.arrsac(O.arrsac_cfg(thr, **kw), 0, a, b, orng)
    assert (gd is None) == (gh is None) == (want is None)
    assert sd == sh == [int(x) for x in orng.s]
    if gd is not None:
        assert np.array_equal(gd[2], gh[2]) and np.array_equal(gd[2], want[2])
        assert np.array_equal(gd[0], gh[0]) and np.array_equal(gd[1], gh[1])
        assert np.allclose(gd[0], want[0], atol=1e-9) and np.allclose(gd[1], want[1], atol=1e-9)


def test_device_driver_p3p_single_view_configuration_equals_oracle():
    # vslam-sandbox/src/main.rs:105-111: Arrsac(1e-5).initialization_hypotheses(16384).max_candidate_hypotheses(1024).estimations_per_block(256)
    rng = np.random.default_rng(9)
    R, t, bear, world, good = pnp_scene(rng, 2000, outlier_frac=0.2, noise=1e-4)
    ars = cv_b200.Arrsac(1e-5, cv_b200.Xoshiro256PlusPlus(0)).initialization_hypotheses(16384).max_candidate_hypotheses(1024).estimations_per_block(256)
    got = ars.model_inliers(cv_b200.LambdaTwist(), bear, world)
    orng = O.rng_xoshiro(0)
    want = O.arrsac(O.arrsac_cfg(1e-5, initialization_hypotheses=16384, max_candidate_hypotheses=1024, estimations_per_block=256), 1, bear, world, orng)
    assert got is not None and want is not None
    assert np.array_equal(got[2], want[2])
    assert np.allclose(got[0], want[0], atol=1e-9) and np.allclose(got[1], want[1], atol=1e-9)
    assert [int(x) for x in ars.rng.state.s] == [int(x) for x in orng.s]
    assert good[got[2]].mean() > 0.99


def test_pcg64_generator_on_device_path():
    rng = np.random.default_rng(21)
    R, t, a, b, _ = two_view_scene(rng, 400, outlier_frac=0.3, noise=1e-4)
    ars = cv_b200.Arrsac(1e-6, cv_b200.Pcg64(bytes([1] * 32)))
    got = ars.model_inliers(cv_b200.EightPoint(), a, b)
    orng = O.rng_pcg64(bytes([1] * 32))
    want = O.arrsac(O.arrsac_cfg(1e-6), 0, a, b, orng)
    assert got is not None and np.array_equal(got[2], want[2])
    assert [int(x) for x in ars.rng.state.s] == [int(x) for x in orng.s]


def test_two_view_frames_equals_separate_calls():
    a = synth_frame(3, h=540, w=960, nblobs=2500)
    b = warp_frame(a, 1003)
    frames = np.stack([a, b])
    ak = cv_b200.Akaze(maximum_features=5000)
    cam = cv_b200.CameraIntrinsics(focals=(1000.0, 1000.0), principal_point=(480.0, 270.0))
    ars = _vslam_two_view()
    out = cv_b200.two_view_frames(ak, frames, cam, ars, better_by=24, cap=5000)
    kps, descs = ak.extract_batch(frames)
    for f in range(2):
        assert out["keypoints"][f].tobytes() == kps[f].tobytes() and np.array_equal(out["descriptors"][f], descs[f])
    pairs = cv_b200.symmetric_matching(descs[0], descs[1], 24)
    assert np.array_equal(out["matches"], pairs) and len(pairs) > 300
    ba = cam.calibrate_keypoints(kps[0][pairs[:, 0]]); bb = cam.calibrate_keypoints(kps[1][pairs[:, 1]])
    ars2 = _vslam_two_view()
    want = ars2.model_inliers(cv_b200.EightPoint(), ba, bb)
    assert (out["pose"] is None) == (want is None)
    if want is not None:
        assert np.array_equal(out["inliers"], want[2])
        assert np.array_equal(out["pose"][0], want[0]) and np.array_equal(out["pose"][1], want[1])
    assert [int(x) for x in ars.rng.state.s] == [int(x) for x in ars2.rng.state.s]
    # and the oracle, end to end on the bearings
    owant = O.arrsac(O.arrsac_cfg(1e-7, initialization_hypotheses=8192, max_candidate_hypotheses=1024), 0, ba, bb, O.rng_xoshiro(0))
    assert (owant is None) == (want is None)
    if want is not None:
        assert np.array_equal(want[2], owant[2])


def test_graph_replay_of_a_run_equals_eager_and_oracle(monkeypatch):
    # the consensus run is captured into a CUDA graph the second time the same (configuration, buffers) key is seen and replayed
    # afterwards: call 1 eager, call 2 capture + launch, calls 3-4 replay -- with a continuing generator every call draws new samples,
    # and every call must equal the oracle continuing from the same state; CVB_ARS_NO_GRAPH=1 (a fresh context) gives the same results
    rng = np.random.default_rng(77)
    R, t, a, b, good = two_view_scene(rng, 900, outlier_frac=0.3, noise=2e-4)
    cfg = dict(initialization_hypotheses=512, max_candidate_hypotheses=128)
    ars = cv_b200.Arrsac(1e-6, cv_b200.Xoshiro256PlusPlus(5)).initialization_hypotheses(512).max_candidate_hypotheses(128)
    orng = O.rng_xoshiro(5)
    results = []
    for call in range(4):
        got = ars.model_inliers(cv_b200.EightPoint(), a, b)
        want = O.arrsac(O.arrsac_cfg(1e-6, **cfg), 0, a, b, orng)
        assert (got is None) == (want is None)
        if got is not None:
            assert np.array_equal(got[2], want[2]), call
            assert np.allclose(got[0], want[0], rtol=1e-6, atol=1e-12) and np.allclose(got[1], want[1], rtol=1e-6, atol=1e-12)
        assert [int(x) for x in ars.rng.state.s] == [int(x) for x in orng.s], call
        results.append(None if got is None else got[2])
    monkeypatch.setenv("CVB_ARS_NO_GRAPH", "1")
    ctx = cv_b200.Context(0)
    ars2 = cv_b200.Arrsac(1e-6, cv_b200.Xoshiro256PlusPlus(5), ctx=ctx).initialization_hypotheses(512).max_candidate_hypotheses(128)
    for call in range(4):
        got = ars2.model_inliers(cv_b200.EightPoint(), a, b)
        assert (got is None) == (results[call] is None)
        if got is not None:
            assert np.array_equal(got[2], results[call])
    ctx.close()
