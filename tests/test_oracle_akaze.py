"""Pins the CPU oracle against the reference's own goldens (CPU only).

Reference assertions reproduced here:
  akaze/tests/estimate_pose.rs:41-42  399 / 343 descriptors at Akaze::sparse()
  akaze/tests/estimate_pose.rs:59     11 Lowe-ratio(0.5) LinearKnn matches
  akaze/src/image.rs:395-412          gaussian_kernel(3.0, 7) known answer
"""
import numpy as np
import pytest

from oracle import pyoracle as O
from tests.common import goldens, kitti_frame, lowe_matches


@pytest.fixture(scope="module")
def sparse():
    out = {}
    for name in ("0000000000", "0000000014"):
        ak = O.Akaze(detector_threshold=0.01)
        out[name] = ak.extract(kitti_frame(name))
    return out


def test_reference_golden_descriptor_counts(sparse):
    assert len(sparse["0000000000"][1]) == 399
    assert len(sparse["0000000014"][1]) == 343


def test_reference_golden_match_count(sparse):
    idx, dist = O.hamming_knn(sparse["0000000000"][1], sparse["0000000014"][1], 2)
    assert lowe_matches(dist) == 11


def test_committed_oracle_vectors_are_current(sparse):
    import os
    from tests.common import GOLDEN
    g = np.load(os.path.join(GOLDEN, "oracle_kitti_sparse.npz"))
    assert np.array_equal(g["desc0"], sparse["0000000000"][1])
    assert np.array_equal(g["desc14"], sparse["0000000014"][1])
    assert g["kps0"].tobytes() == sparse["0000000000"][0].tobytes()


def test_gaussian_kernel_known_answer():
    k = O.gaussian_kernel(3.0, 7)
    ref = np.array([0.10628852, 0.14032133, 0.16577007, 0.17524014, 0.16577007, 0.14032133, 0.10628852], np.float32)
    assert np.all(np.abs(k - ref) < 1e-4)
    assert abs(float(k.sum()) - 1.0) < 1e-6


def test_filters_replicate_border_and_correlation():
    # akaze/src/image.rs:414-432 checks against imageproc (replicate edge, no kernel flip) to 1e-4.
    rng = np.random.default_rng(0)
    img = rng.random((37, 53), dtype=np.float32)
    k = O.gaussian_kernel(3.0, 7)
    pad = np.pad(img.astype(np.float64), ((0, 0), (3, 3)), mode="edge")
    want = sum(pad[:, j:j + 53] * float(k[j]) for j in range(7))
    assert np.max(np.abs(O.horizontal_filter(img, k) - want)) < 1e-5
    pad = np.pad(img.astype(np.float64), ((3, 3), (0, 0)), mode="edge")
    want = sum(pad[j:j + 37, :] * float(k[j]) for j in range(7))
    assert np.max(np.abs(O.vertical_filter(img, k) - want)) < 1e-5
    # asymmetric kernel: correlation, not convolution
    ka = np.array([-1, 0, 1], np.float32)
    h = O.horizontal_filter(img, ka)
    assert np.array_equal(h[:, 1:-1], img[:, 2:] - img[:, :-2])


def test_half_size_odd_dims():
    rng = np.random.default_rng(1)
    img = rng.random((7, 9), dtype=np.float32)
    out = O.half_size(img)
    assert out.shape == (3, 4)
    assert out[0, 0] == np.float32((img[0, 0] + img[0, 1]) + (img[1, 0] + img[1, 1])) * np.float32(0.25)
    # bottom row / right column / corner overwrite rules (image.rs:167-196)
    assert out[2, 1] == (img[6, 2] + img[6, 3]) * np.float32(0.5)
    assert out[1, 3] == (img[2, 8] + img[3, 8]) * np.float32(0.5)
    assert out[2, 3] == img[6, 8]


def test_fed_tau_schedule_1080p():
    # n_i for 1920x1080 defaults
    ak = O.Akaze()
    img = np.zeros((1080, 1920), np.float32)
    img[::7, ::5] = 1.0
    ak.extract(img)
    n = [len(ak.evolution_info(i)["tau"]) for i in range(ak.num_evolutions())]
    assert n == [0, 3, 3, 4, 4, 5, 6, 7, 8, 10, 12, 14, 17, 20, 24, 29]
    for i in range(1, 16):
        info = ak.evolution_info(i)
        prev = ak.evolution_info(i - 1)
        ttime = 0.5 * info["esigma"] ** 2 - 0.5 * prev["esigma"] ** 2
        assert abs(sum(info["tau"]) - ttime) < 1e-9 * max(1.0, ttime)


def test_secondary_counts_default_threshold():
    g = goldens()["oracle_derived"]["default"]
    ak = O.Akaze()
    kps, d = ak.extract(kitti_frame("0000000000"))
    assert {s: len(ak.stage(s)) for s in O.STAGES} == g["stages"]["0000000000"]
    # descriptors: 486 bits -> bytes 61..63 and the top 2 bits of byte 60 are zero (descriptors.rs:60,197)
    assert not d[:, 61:].any() and not (d[:, 60] & 0xC0).any()
    # descending response order (lib.rs:326)
    assert np.all(np.diff(kps["response"]) <= 0)


def test_knn_tie_break_earlier_index_first():
    db = np.zeros((5, 64), np.uint8)
    db[0, 0] = 0b11      # d=2
    db[1, 0] = 0b1       # d=1
    db[2, 1] = 0b1       # d=1 (tie with 1, later index)
    db[3, 0] = 0b111     # d=3
    q = np.zeros((1, 64), np.uint8)
    idx, dist = O.hamming_knn(q, db, 3)
    assert idx.tolist() == [[4, 1, 2]] and dist.tolist() == [[0, 1, 1]]
    idx, dist = O.hamming_knn(q, db[:1], 2)   # fewer points than k
    assert idx[0, 0] == 0 and idx[0, 1] == 0xFFFFFFFF
