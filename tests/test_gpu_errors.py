"""Error behaviour of the C ABI on a GPU box: bad arguments return CVB_E* codes with a message, capacity
overflows are reported (never silently truncated), and nothing aborts the process."""
import ctypes as C

import numpy as np
import pytest

import cv_b200
from cv_b200._lib import CVB_ECAP, CVB_EINVAL, CVB_EUNSUPPORTED, KP_DTYPE
from tests.common import kitti_frame
from tests.synth import random_descriptors

pytestmark = pytest.mark.gpu


def test_invalid_arguments_are_reported_not_fatal():
    ctx = cv_b200.Context(0)
    L = ctx.lib
    cfg = cv_b200.AkazeConfig().to_c()
    n = C.c_uint32()
    img = np.zeros((64, 64), np.float32)
    kp = np.zeros(16, KP_DTYPE); d = np.zeros((16, 64), np.uint8)
    assert L.cvb_akaze_extract(ctx.handle, C.byref(cfg), None, 64, 64, kp.ctypes.data, d.ctypes.data, 16, C.byref(n)) == CVB_EINVAL
    assert L.cvb_akaze_extract(ctx.handle, C.byref(cfg), img.ctypes.data, 0, 64, kp.ctypes.data, d.ctypes.data, 16, C.byref(n)) == CVB_EINVAL
    assert b"empty" in L.cvb_last_error(ctx.handle)
    bad = cv_b200.AkazeConfig(descriptor_channels=7).to_c()
    assert L.cvb_akaze_extract(ctx.handle, C.byref(bad), img.ctypes.data, 64, 64, kp.ctypes.data, d.ctypes.data, 16, C.byref(n)) == CVB_EINVAL
    bad = cv_b200.AkazeConfig(base_scale_offset=-1.0).to_c()       # the reference asserts sigma > 0 (image.rs:384)
    assert L.cvb_akaze_extract(ctx.handle, C.byref(bad), img.ctypes.data, 64, 64, kp.ctypes.data, d.ctypes.data, 16, C.byref(n)) == CVB_EINVAL
    bad = cv_b200.AkazeConfig(descriptor_pattern_size=40).to_c()
    assert L.cvb_akaze_extract(ctx.handle, C.byref(bad), img.ctypes.data, 64, 64, kp.ctypes.data, d.ctypes.data, 16, C.byref(n)) == CVB_EUNSUPPORTED
    q = random_descriptors(4, 0)
    idx = np.zeros((4, 9), np.uint32)
    assert L.cvb_hamming_knn(ctx.handle, q.ctypes.data, 4, q.ctypes.data, 4, 9, idx.ctypes.data, idx.ctypes.data) == CVB_EINVAL   # k > 8
    # the context is still usable after errors
    k, dd = cv_b200.Akaze.sparse(ctx=ctx).extract_from_gray_float_image(kitti_frame("0000000000")[:200, :300].copy())
    assert len(k) == len(dd)


def test_output_capacity_overflow_is_an_error():
    img = kitti_frame("0000000000")
    ak = cv_b200.Akaze(0.001, max_keypoints=100)       # 3425 keypoints do not fit 100 slots
    with pytest.raises(cv_b200.CvbError) as e:
        ak.extract_from_gray_float_image(img)
    assert e.value.code == CVB_ECAP
    ak2 = cv_b200.Akaze(0.001, max_keypoints=100, maximum_features=100)   # Akaze::maximum_features truncation is fine
    k, d = ak2.extract_from_gray_float_image(img)
    assert len(d) == 100


def test_python_layer_rejects_bad_shapes():
    with pytest.raises(ValueError):
        cv_b200.hamming_knn(np.zeros((3, 32), np.uint8), np.zeros((3, 64), np.uint8))
    with pytest.raises(TypeError):
        cv_b200.Akaze().extract(np.zeros((8, 8), np.float64))
    with pytest.raises(ValueError):
        cv_b200.Akaze().extract_batch(np.zeros((8, 8), np.float32))
    assert cv_b200.symmetric_matching(random_descriptors(1, 0), random_descriptors(5, 1)).shape == (0, 2)   # < 2 features: no matches


def test_failed_workspace_build_is_not_cached():
    """A configuration that fails while the workspace is being built (unsupported pattern size: the error comes AFTER the planes
    are allocated) must fail identically when the same call is repeated, and must not poison the context."""
    ctx = cv_b200.Context(0)
    L = ctx.lib
    n = C.c_uint32()
    img = kitti_frame("0000000000")[:128, :160].copy()
    kp = np.zeros(4096, KP_DTYPE); d = np.zeros((4096, 64), np.uint8)
    bad = cv_b200.AkazeConfig(descriptor_pattern_size=40).to_c()
    for _ in range(3):
        assert L.cvb_akaze_extract(ctx.handle, C.byref(bad), img.ctypes.data, 160, 128, kp.ctypes.data, d.ctypes.data, 4096, C.byref(n)) == CVB_EUNSUPPORTED
    good = cv_b200.AkazeConfig().to_c()
    assert L.cvb_akaze_extract(ctx.handle, C.byref(good), img.ctypes.data, 160, 128, kp.ctypes.data, d.ctypes.data, 4096, C.byref(n)) == 0
    assert n.value > 0


def test_geometry_wrappers_validate_lengths():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((20, 3)); b = rng.standard_normal((19, 3))
    with pytest.raises(ValueError):
        cv_b200.residuals_camera_to_camera([(np.eye(3), np.ones(3))], a, b)
    with pytest.raises(ValueError):
        cv_b200.Arrsac(1e-6, cv_b200.Xoshiro256PlusPlus(0)).model_inliers(cv_b200.EightPoint(), a, b)
    with pytest.raises(ValueError):
        cv_b200.LinearEigenTriangulator().triangulate_batch([(np.eye(3), np.zeros(3))] * 3, rng.standard_normal((3, 3)), [0, 2, 5])   # offsets past the data
    with pytest.raises(ValueError):
        cv_b200.LinearEigenTriangulator().triangulate_batch([(np.eye(3), np.zeros(3))] * 3, rng.standard_normal((3, 3)), [0, 2, 1])   # not monotone


def test_device_resident_extract_clamps_and_flags_overflow():
    """cvb_akaze_extract_batch_dev with too small an output capacity: the count is clamped (downstream kernels index by it) and the
    truncation is reported by cvb_akaze_dev_overflow, once."""
    import torch
    ctx = cv_b200.Context(0)
    L = ctx.lib
    dev = torch.device("cuda", 0)
    img = torch.from_numpy(kitti_frame("0000000000")[None].copy()).to(dev)
    cap = 100
    kp = torch.zeros(cap * KP_DTYPE.itemsize, dtype=torch.uint8, device=dev); d = torch.zeros(cap * 64, dtype=torch.uint8, device=dev)
    n = torch.zeros(1, dtype=torch.int32, device=dev)
    cfg = cv_b200.AkazeConfig().to_c()
    ctx.check(L.cvb_akaze_extract_batch_dev(ctx.handle, C.byref(cfg), img.data_ptr(), 1, img.shape[2], img.shape[1], kp.data_ptr(), d.data_ptr(), cap, n.data_ptr()))
    flag = C.c_uint32()
    ctx.check(L.cvb_akaze_dev_overflow(ctx.handle, C.byref(flag)))
    assert int(n.cpu()[0]) == cap and flag.value == 3
    ctx.check(L.cvb_akaze_dev_overflow(ctx.handle, C.byref(flag)))
    assert flag.value == 0
