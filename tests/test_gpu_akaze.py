"""GPU parity of the CUDA AKAZE extractor against the CPU oracle, stage by stage, through the C ABI.
Bar: bit-exact planes, keypoints and descriptors (integer/index work and f32 with the reference's
rounding order)."""
import os

import numpy as np
import pytest

import cv_b200
from oracle import pyoracle as O
from tests.common import GOLDEN, kitti_frame
from tests.synth import synth_frame

pytestmark = pytest.mark.gpu


def _bits_equal(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def _first_diff(a, b):
    d = np.argwhere(a.view(np.uint32) != b.view(np.uint32))
    return None if len(d) == 0 else (tuple(d[0]), a[tuple(d[0])], b[tuple(d[0])], len(d))


def _compare_all(img, thr, **cfg):
    ora = O.Akaze(detector_threshold=thr, **cfg)
    okp, odesc = ora.extract(img)
    ak = cv_b200.Akaze(thr, **cfg)
    gkp, gdesc = ak.extract_from_gray_float_image(img)
    evs = ak.debug_evolutions()
    assert len(evs) == ora.num_evolutions()
    assert ak.debug_contrast() == ora.contrast_factor()
    for i, ev in enumerate(evs):
        info = ora.evolution_info(i)
        assert (ev["w"], ev["h"], ev["octave"], ev["n_fed_steps"]) == (info["w"], info["h"], info["octave"], len(info["tau"]))
        for name in ("Lsmooth", "Lflow", "Lt", "Lx", "Ly", "Ldet"):
            if i == 0 and name == "Lflow":
                continue
            want = ora.plane(i, name)
            got = ak.debug_plane(i, name)
            assert _bits_equal(got, want), (i, name, _first_diff(got, want))
    for st in ("candidates", "extrema", "refined", "sorted"):
        want, got = ora.stage(st), ak.debug_stage(st)
        assert len(got) == len(want), (st, len(got), len(want))
        assert got.tobytes() == want.tobytes(), st
    assert gkp.tobytes() == okp.tobytes()
    assert np.array_equal(gdesc, odesc)
    return gkp, gdesc


def test_kitti_sparse_matches_reference_goldens_and_oracle():
    kps, d = _compare_all(kitti_frame("0000000000"), 0.01)
    assert len(d) == 399                       # akaze/tests/estimate_pose.rs:41
    g = np.load(os.path.join(GOLDEN, "oracle_kitti_sparse.npz"))
    assert np.array_equal(d, g["desc0"]) and kps.tobytes() == g["kps0"].tobytes()
    kps, d = _compare_all(kitti_frame("0000000014"), 0.01)
    assert len(d) == 343                       # akaze/tests/estimate_pose.rs:42


def test_kitti_default_threshold():
    kps, d = _compare_all(kitti_frame("0000000000"), 0.001)
    assert len(d) == 3425


def test_odd_sizes_and_small_images():
    # odd dimensions exercise half_size's tail rows/columns (image.rs:167-196) and ragged tiles
    img = synth_frame(5, h=301, w=415, nblobs=600)
    _compare_all(img, 0.001)
    img = synth_frame(6, h=97, w=131, nblobs=80)
    _compare_all(img, 0.0005)


def test_too_small_image_returns_nothing():
    img = np.random.default_rng(0).random((30, 30), dtype=np.float32)
    kps, d = cv_b200.Akaze().extract_from_gray_float_image(img)
    assert len(kps) == 0 and len(d) == 0
    assert O.Akaze().extract(img)[1].shape[0] == 0


def test_non_default_config():
    img = synth_frame(7, h=240, w=320, nblobs=400)
    _compare_all(img, 0.002, num_sublevels=3, max_octave_evolution=3, descriptor_channels=2, maximum_features=40)
    _compare_all(img, 0.002, descriptor_channels=1, contrast_factor_num_bins=128, contrast_percentile=0.6)
    # derivative sigma 1 takes the un-normalised simple Scharr path (derivatives.rs:24-26); sigma 5 a wider kernel
    _compare_all(img, 0.05, derivative_factor=0.7)
    _compare_all(img, 0.0005, derivative_factor=3.0, base_scale_offset=1.2)


def test_batch_equals_single_and_is_deterministic():
    a, b = kitti_frame("0000000000"), kitti_frame("0000000014")
    ak = cv_b200.Akaze.sparse()
    kps, descs = ak.extract_batch(np.stack([a, b]))
    assert len(descs[0]) == 399 and len(descs[1]) == 343
    k0, d0 = ak.extract_from_gray_float_image(a)
    assert np.array_equal(d0, descs[0]) and k0.tobytes() == kps[0].tobytes()
    kps2, descs2 = ak.extract_batch(np.stack([a, b]))
    assert np.array_equal(descs2[1], descs[1])


def test_full_hd_frame_parity():
    img = synth_frame(0)
    kps, d = _compare_all(img, 0.001, maximum_features=5000)
    assert len(d) > 1000


def test_suppression_kernel_variants_agree():
    """The shared-memory, global-memory and serial duplicate-suppression kernels give identical output."""
    import hashlib
    import subprocess
    import sys
    code = ("import numpy as np, hashlib, cv_b200; from tests.common import kitti_frame; "
            "k, d = cv_b200.Akaze().extract_from_gray_float_image(kitti_frame('0000000000')); "
            "print(len(d), hashlib.sha1(d.tobytes() + k.tobytes()).hexdigest())")
    outs = []
    for var in (None, "CVB_SUPPRESS_GLOBAL", "CVB_SUPPRESS_SEQ", "CVB_NO_FUSE_BLUR", "CVB_NO_GRAPH"):
        env = dict(os.environ)
        if var:
            env[var] = "1"
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=600).stdout.strip())
    assert outs[0].startswith("3425 ") and all(o == outs[0] for o in outs), outs


def test_extract_from_8_and_16_bit_images_follows_from_dynamic():
    # Akaze::extract -> GrayFloatImage::from_dynamic (akaze/src/image.rs:45-69): u8 / 255f32, u16 / 65535f32, both single f32 divisions
    im8 = np.load(os.path.join(GOLDEN, "kitti_0000000000.npz"))["image"]
    assert im8.dtype == np.uint8
    ak = cv_b200.Akaze(0.01)
    kps, d = ak.extract(im8)
    assert len(d) == 399                       # the reference's own golden goes through this entry (estimate_pose.rs:28-41)
    okp, odesc = O.Akaze(detector_threshold=0.01).extract(im8.astype(np.float32) / np.float32(255))
    assert kps.tobytes() == okp.tobytes() and np.array_equal(d, odesc)
    im16 = im8.astype(np.uint16) * np.uint16(257) + np.uint16(3)          # not a multiple of the 8-bit levels
    kps16, d16 = ak.extract(im16)
    okp, odesc = O.Akaze(detector_threshold=0.01).extract(im16.astype(np.float32) / np.float32(65535))
    assert len(d16) > 0 and kps16.tobytes() == okp.tobytes() and np.array_equal(d16, odesc)
    with pytest.raises(TypeError):
        ak.extract(im8.astype(np.float64))     # image.rs:107: any other pixel type panics upstream
    with pytest.raises(ValueError):
        ak.extract(np.zeros((8, 8, 3), np.uint8))
