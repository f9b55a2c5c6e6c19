"""Pins the restated glibc float routines (oracle/ref_libm.h) against the host libm.

Rust's f32::{sin,cos,atan2} call the platform libm on linux-gnu, so the host libm IS the
reference for akaze/src/descriptors.rs:70-71 and scale_space_extrema.rs:242.  The exhaustive
sweep over every float in [0, 120) was run when the restatement was written (0 mismatches);
this test keeps a large sampled sweep so it stays in the minutes-budget CPU suite.
"""
import ctypes as C

import numpy as np

from oracle import pyoracle as O

libm = C.CDLL("libm.so.6")
for f in ("sinf", "cosf"):
    getattr(libm, f).restype = C.c_float
    getattr(libm, f).argtypes = [C.c_float]
libm.atan2f.restype = C.c_float
libm.atan2f.argtypes = [C.c_float, C.c_float]


def test_sin_cos_match_host_libm_on_angle_range():
    L = O.lib()
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.uniform(0, 2 * np.pi, 40000), rng.uniform(-8, 120, 5000),
                         np.array([0.0, 1e-5, np.pi / 4, np.pi / 2, np.pi, 2 * np.pi])]).astype(np.float32)
    for x in xs:
        x = float(x)
        assert np.float32(L.ref_sinf(x)).tobytes() == np.float32(libm.sinf(x)).tobytes(), x
        assert np.float32(L.ref_cosf(x)).tobytes() == np.float32(libm.cosf(x)).tobytes(), x


def test_atan2_matches_host_libm():
    L = O.lib()
    rng = np.random.default_rng(4)
    ys = (rng.standard_normal(40000) * 10.0 ** rng.integers(-6, 3, 40000)).astype(np.float32)
    xs = (rng.standard_normal(40000) * 10.0 ** rng.integers(-6, 3, 40000)).astype(np.float32)
    ys[:10] = 0.0
    xs[10:20] = 0.0
    xs[20:30] = 1.0
    for y, x in zip(ys, xs):
        y, x = float(y), float(x)
        assert np.float32(L.ref_atan2f(y, x)).tobytes() == np.float32(libm.atan2f(y, x)).tobytes(), (y, x)
    v = L.ref_fast_atan2_equiv(-1.0, 1.0)
    assert 0.0 <= v < 2 * np.pi
