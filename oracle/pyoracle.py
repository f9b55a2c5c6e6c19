"""ctypes binding of the CPU oracle (oracle/_build/libcvb_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (cv_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcvb_oracle.so")


class AkazeCfg(C.Structure):
    """mirrors akaze::Akaze (akaze/src/lib.rs:109-142)"""
    _fields_ = [
        ("maximum_features", C.c_int64),
        ("num_sublevels", C.c_uint32),
        ("max_octave_evolution", C.c_uint32),
        ("base_scale_offset", C.c_double),
        ("initial_contrast", C.c_double),
        ("contrast_percentile", C.c_double),
        ("contrast_factor_num_bins", C.c_uint64),
        ("derivative_factor", C.c_double),
        ("detector_threshold", C.c_double),
        ("descriptor_channels", C.c_uint64),
        ("descriptor_pattern_size", C.c_uint64),
    ]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("response", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("octave", "<u4"), ("class_id", "<u4")])

PLANES = {"Lt": 0, "Lsmooth": 1, "Lx": 2, "Ly": 3, "Lflow": 4, "Ldet": 5, "Lxx": 6, "Lyy": 7, "Lxy": 8}
STAGES = {"candidates": 0, "extrema": 1, "refined": 2, "sorted": 3, "final": 4}


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        L.ref_akaze_default_cfg.argtypes = [C.POINTER(AkazeCfg)]
        L.ref_akaze_create.restype = C.c_void_p
        L.ref_akaze_create.argtypes = [C.POINTER(AkazeCfg)]
        L.ref_akaze_destroy.argtypes = [C.c_void_p]
        L.ref_akaze_extract.argtypes = [C.c_void_p, fp, C.c_int, C.c_int]
        L.ref_akaze_num_evolutions.argtypes = [C.c_void_p]
        L.ref_akaze_evolution_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                               C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.POINTER(C.c_int),
                                               C.POINTER(C.c_double)]
        L.ref_akaze_contrast_factor.restype = C.c_double
        L.ref_akaze_contrast_factor.argtypes = [C.c_void_p]
        L.ref_akaze_plane.restype = fp
        L.ref_akaze_plane.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_akaze_stage.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.ref_akaze_descriptors.restype = C.POINTER(C.c_uint8)
        L.ref_akaze_descriptors.argtypes = [C.c_void_p]
        L.ref_horizontal_filter.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, fp]
        L.ref_vertical_filter.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, fp]
        L.ref_gaussian_kernel.argtypes = [C.c_float, C.c_int, fp]
        L.ref_half_size.argtypes = [fp, C.c_int, C.c_int, fp]
        L.ref_fed_tau.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double), C.c_int]
        for f in ("ref_sinf", "ref_cosf"):
            getattr(L, f).restype = C.c_float
            getattr(L, f).argtypes = [C.c_float]
        for f in ("ref_atan2f", "ref_fast_atan2_equiv"):
            getattr(L, f).restype = C.c_float
            getattr(L, f).argtypes = [C.c_float, C.c_float]
        L.ref_hamming_knn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                      C.c_void_p]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def set_num_threads(n):
    L = lib()
    L.ref_set_num_threads.argtypes = [C.c_int]
    L.ref_set_num_threads(int(n))


def default_cfg(**kw):
    c = AkazeCfg()
    lib().ref_akaze_default_cfg(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


class Akaze:
    """Oracle counterpart of akaze::Akaze::extract_from_gray_float_image (akaze/src/lib.rs:309-339)."""

    def __init__(self, cfg=None, **kw):
        self.cfg = cfg if cfg is not None else default_cfg(**kw)
        self._h = lib().ref_akaze_create(C.byref(self.cfg))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_akaze_destroy(self._h)
            self._h = None

    def extract(self, image):
        image = np.ascontiguousarray(image, dtype=np.float32)
        h, w = image.shape
        self._shape = (h, w)
        n = lib().ref_akaze_extract(self._h, _fp(image), w, h)
        kps = self.stage("final")
        if n <= 0:
            return kps, np.zeros((0, 64), np.uint8)
        d = np.ctypeslib.as_array(lib().ref_akaze_descriptors(self._h), shape=(n, 64)).copy()
        return kps, d

    def num_evolutions(self):
        return lib().ref_akaze_num_evolutions(self._h)

    def evolution_info(self, i):
        w, h, nt = C.c_int(), C.c_int(), C.c_int()
        o, es = C.c_uint32(), C.c_double()
        tau = (C.c_double * 256)()
        r = lib().ref_akaze_evolution_info(self._h, i, C.byref(w), C.byref(h), C.byref(o), C.byref(es), C.byref(nt), tau)
        assert r == 0
        return dict(w=w.value, h=h.value, octave=o.value, esigma=es.value, tau=[tau[j] for j in range(nt.value)])

    def contrast_factor(self):
        return lib().ref_akaze_contrast_factor(self._h)

    def plane(self, i, name):
        info = self.evolution_info(i)
        p = lib().ref_akaze_plane(self._h, i, PLANES[name])
        if not p:
            return None
        return np.ctypeslib.as_array(p, shape=(info["h"], info["w"])).copy()

    def stage(self, name):
        ptr = C.c_void_p()
        n = lib().ref_akaze_stage(self._h, STAGES[name], C.byref(ptr))
        if n <= 0 or not ptr.value:
            return np.zeros(0, dtype=KP_DTYPE)
        buf = (C.c_uint8 * (n * KP_DTYPE.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=KP_DTYPE, count=n).copy()


def horizontal_filter(img, kernel):
    img = np.ascontiguousarray(img, np.float32); k = np.ascontiguousarray(kernel, np.float32)
    out = np.empty_like(img)
    lib().ref_horizontal_filter(_fp(img), img.shape[1], img.shape[0], _fp(k), len(k), _fp(out))
    return out


def vertical_filter(img, kernel):
    img = np.ascontiguousarray(img, np.float32); k = np.ascontiguousarray(kernel, np.float32)
    out = np.empty_like(img)
    lib().ref_vertical_filter(_fp(img), img.shape[1], img.shape[0], _fp(k), len(k), _fp(out))
    return out


def gaussian_kernel(r, ks):
    out = np.empty(ks, np.float32)
    lib().ref_gaussian_kernel(r, ks, _fp(out))
    return out


def half_size(img):
    img = np.ascontiguousarray(img, np.float32)
    out = np.zeros((img.shape[0] // 2, img.shape[1] // 2), np.float32)
    lib().ref_half_size(_fp(img), img.shape[1], img.shape[0], _fp(out))
    return out


def fed_tau(T, tau_max=0.25):
    out = (C.c_double * 256)()
    n = lib().ref_fed_tau(T, tau_max, out, 256)
    return [out[i] for i in range(n)]


def hamming_knn(q, db, k=2):
    """space::LinearKnn{Hamming}.knn for every query: (idx[n,k], dist[n,k]) uint32."""
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 64); db = np.ascontiguousarray(db, np.uint8).reshape(-1, 64)
    idx = np.empty((len(q), k), np.uint32); dist = np.empty((len(q), k), np.uint32)
    lib().ref_hamming_knn(q.ctypes.data, len(q), db.ctypes.data, len(db), k, idx.ctypes.data, dist.ctypes.data)
    return idx, dist


# ---------------------------------------------------------------------------------------------
# geometry oracle (oracle/ref_geom.c)
class Pose(C.Structure):
    _fields_ = [("R", C.c_double * 9), ("t", C.c_double * 3)]

    def numpy(self):
        return np.array(self.R, dtype=np.float64).reshape(3, 3), np.array(self.t, dtype=np.float64)


class Rng(C.Structure):
    _fields_ = [("kind", C.c_int), ("s", C.c_uint64 * 4)]


class ArrsacCfg(C.Structure):
    _fields_ = [("inlier_threshold", C.c_double), ("initialization_hypotheses", C.c_uint32), ("initialization_blocks", C.c_uint32),
                ("max_candidate_hypotheses", C.c_uint32), ("estimations_per_block", C.c_uint32), ("block_size", C.c_uint32),
                ("likelihood_ratio_threshold", C.c_float), ("initial_epsilon", C.c_float), ("initial_delta", C.c_float)]


_geom_ready = False


def _geom():
    global _geom_ready
    L = lib()
    if not _geom_ready:
        dp = C.POINTER(C.c_double)
        L.ref_sym_eigen.argtypes = [C.c_int, dp, C.c_double, C.c_int, dp, dp]
        L.ref_sym_eigen9_rr.argtypes = [dp, C.c_double, C.c_int, dp, dp]
        L.ref_eight_point_essential.argtypes = [dp, dp, C.c_double, C.c_int, dp]
        L.ref_essential_poses.argtypes = [dp, C.c_double, C.c_int, C.POINTER(Pose)]
        L.ref_eight_point.argtypes = [dp, dp, C.POINTER(Pose)]
        L.ref_essential_residual.restype = C.c_double
        L.ref_essential_residual.argtypes = [dp, dp, dp]
        L.ref_residual_c2c.restype = C.c_double
        L.ref_residual_c2c.argtypes = [C.POINTER(Pose), dp, dp]
        L.ref_residual_w2c.restype = C.c_double
        L.ref_residual_w2c.argtypes = [C.POINTER(Pose), dp, dp]
        L.ref_p3p.argtypes = [dp, dp, C.POINTER(Pose)]
        L.ref_real_eigenvalues10.argtypes = [dp, dp, dp]
        L.ref_five_point_essentials.argtypes = [dp, dp, dp]
        L.ref_five_point.argtypes = [dp, dp, C.POINTER(Pose)]
        L.ref_five_point_set_row0.argtypes = [C.c_int]
        L.ref_triangulate_linear_eigen.argtypes = [C.POINTER(Pose), dp, C.c_int, dp]
        L.ref_calibrate.argtypes = [C.c_double] * 7 + [dp]
        L.ref_rng_seed_xoshiro.argtypes = [C.POINTER(Rng), C.c_uint64]
        L.ref_rng_seed_pcg64.argtypes = [C.POINTER(Rng), C.c_char_p]
        L.ref_rng_next_u32.restype = C.c_uint32
        L.ref_rng_next_u32.argtypes = [C.POINTER(Rng)]
        L.ref_arrsac_default_cfg.argtypes = [C.POINTER(ArrsacCfg), C.c_double]
        L.ref_arrsac.argtypes = [C.POINTER(ArrsacCfg), C.c_int, dp, dp, C.c_uint32, C.POINTER(Rng), C.POINTER(Pose),
                                 C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        _geom_ready = True
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_pose(R, t):
    p = Pose()
    p.R[:] = list(np.asarray(R, np.float64).reshape(9))
    p.t[:] = list(np.asarray(t, np.float64).reshape(3))
    return p


def sym_eigen(A, eps=1e-12, iters=1000):
    A = np.ascontiguousarray(A, np.float64); n = A.shape[0]
    d = np.zeros(n); V = np.zeros((n, n))
    ok = _geom().ref_sym_eigen(n, _dp(A), eps, iters, _dp(d), _dp(V))
    return ok, d, V


def sym_eigen9_rr(A, eps=1e-12, iters=1000):
    """the round-robin Jacobi of the eight-point estimator (ref_geom.c)"""
    A = np.ascontiguousarray(A, np.float64)
    assert A.shape == (9, 9)
    d = np.zeros(9); V = np.zeros((9, 9))
    ok = _geom().ref_sym_eigen9_rr(_dp(A), eps, iters, _dp(d), _dp(V))
    return ok, d, V


def eight_point_essential(a, b, eps=1e-12, iters=1000):
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    E = np.zeros((3, 3))
    ok = _geom().ref_eight_point_essential(_dp(a), _dp(b), eps, iters, _dp(E))
    return E if ok else None


def essential_poses(E, eps=1e-12, iters=1000):
    E = np.ascontiguousarray(E, np.float64)
    out = (Pose * 4)()
    n = _geom().ref_essential_poses(_dp(E), eps, iters, out)
    return [out[i].numpy() for i in range(n)]


def eight_point(a, b):
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    out = (Pose * 4)()
    n = _geom().ref_eight_point(_dp(a), _dp(b), out)
    return [out[i].numpy() for i in range(n)]


def essential_residual(E, a, b):
    E = np.ascontiguousarray(E, np.float64); a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    return _geom().ref_essential_residual(_dp(E), _dp(a), _dp(b))


def residual_c2c(R, t, a, b):
    p = make_pose(R, t); a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    return _geom().ref_residual_c2c(C.byref(p), _dp(a), _dp(b))


def residual_w2c(R, t, bearing, world):
    p = make_pose(R, t); a = np.ascontiguousarray(bearing, np.float64); b = np.ascontiguousarray(world, np.float64)
    return _geom().ref_residual_w2c(C.byref(p), _dp(a), _dp(b))


def p3p(bearings, world):
    a = np.ascontiguousarray(bearings, np.float64); b = np.ascontiguousarray(world, np.float64)
    out = (Pose * 4)()
    n = _geom().ref_p3p(_dp(a), _dp(b), out)
    return [out[i].numpy() for i in range(n)]


def real_eigenvalues10(A):
    A = np.ascontiguousarray(A, np.float64); wr = np.zeros(10); wi = np.zeros(10)
    ok = _geom().ref_real_eigenvalues10(_dp(A), _dp(wr), _dp(wi))
    return ok, wr + 1j * wi


def five_point_set_row0(r0):
    """5 = the reference's eigenvector rows (nister-stewenius/src/lib.rs:229), 6 = mathematically correct rows."""
    _geom().ref_five_point_set_row0(r0)


def five_point_essentials(a, b):
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    Es = np.zeros(90)
    n = _geom().ref_five_point_essentials(_dp(a), _dp(b), _dp(Es))
    return Es[:9 * n].reshape(n, 3, 3).copy()


def five_point(a, b):
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    out = (Pose * 40)()
    n = _geom().ref_five_point(_dp(a), _dp(b), out)
    return [out[i].numpy() for i in range(n)]


def triangulate_linear_eigen(poses, bearings):
    arr = (Pose * len(poses))(*[make_pose(R, t) for R, t in poses])
    b = np.ascontiguousarray(bearings, np.float64)
    out = np.zeros(4)
    ok = _geom().ref_triangulate_linear_eigen(arr, _dp(b), len(poses), _dp(out))
    return out if ok else None


def calibrate(fx, fy, cx, cy, skew, px, py):
    out = np.zeros(3)
    _geom().ref_calibrate(fx, fy, cx, cy, skew, px, py, _dp(out))
    return out


def rng_xoshiro(seed):
    r = Rng(); _geom().ref_rng_seed_xoshiro(C.byref(r), seed); return r


def rng_pcg64(seed_bytes):
    r = Rng(); _geom().ref_rng_seed_pcg64(C.byref(r), bytes(seed_bytes)); return r


def rng_next_u32(r):
    return _geom().ref_rng_next_u32(C.byref(r))


def arrsac_cfg(threshold, **kw):
    c = ArrsacCfg(); _geom().ref_arrsac_default_cfg(C.byref(c), threshold)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def arrsac(cfg, kind, a, b, rng):
    """kind 0: EightPoint over FeatureMatch(a[n,3], b[n,3]); kind 1: LambdaTwist over (bearing[n,3], world[n,4])."""
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    n = len(a)
    model = Pose(); inl = np.zeros(max(n, 1), np.uint32); cnt = C.c_uint32()
    ok = _geom().ref_arrsac(C.byref(cfg), kind, _dp(a), _dp(b), n, C.byref(rng), C.byref(model),
                            inl.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(cnt))
    if not ok:
        return None
    R, t = model.numpy()
    return R, t, inl[:cnt.value].copy()


# ---------------------------------------------------------------------------------------------
# post-consensus refinement and robustness checks (oracle/ref_optimize.c)
_opt_ready = False


def _opt():
    global _opt_ready
    L = _geom()
    if not _opt_ready:
        dp = C.POINTER(C.c_double)
        L.ref_world_pose_gradient.argtypes = [dp, dp, dp, dp]
        L.ref_single_view_optimize_l2.restype = C.c_uint32
        L.ref_single_view_optimize_l2.argtypes = [C.POINTER(Pose), C.c_double, C.c_uint32, dp, dp, C.c_uint32]
        L.ref_three_view_gradients.argtypes = [dp] * 6
        L.ref_three_view_optimize_l2.restype = C.c_uint32
        L.ref_three_view_optimize_l2.argtypes = [C.POINTER(Pose), C.c_int, C.c_double, C.c_uint32, dp, C.c_uint32]
        L.ref_epipolar_loss.restype = C.c_double
        L.ref_epipolar_loss.argtypes = [dp, dp, dp]
        L.ref_observation_losses.argtypes = [C.POINTER(Pose), dp, C.c_uint32, dp]
        L.ref_is_tri_landmark_robust.argtypes = [C.POINTER(Pose), C.POINTER(Pose), dp, dp, dp, C.c_double, C.c_double]
        _opt_ready = True
    return L


def single_view_optimize_l2(pose, rate, iterations, bearings, world):
    """single_view_simple_optimize_l2 -> (R, t, pose updates applied)"""
    p = make_pose(*pose)
    b = np.ascontiguousarray(bearings, np.float64); w = np.ascontiguousarray(world, np.float64)
    it = _opt().ref_single_view_optimize_l2(C.byref(p), rate, iterations, _dp(b), _dp(w), len(b))
    R, t = p.numpy()
    return R, t, it


def three_view_gradients(c, f, ftoc, s, stoc):
    out = np.zeros(12)
    a = [np.ascontiguousarray(x, np.float64) for x in (c, f, ftoc, s, stoc)]
    _opt().ref_three_view_gradients(*[_dp(x) for x in a], _dp(out))
    return out


def three_view_optimize_l2(poses, rate, iterations, obs, adaptive=False):
    """three_view_simple_optimize_l2 / three_view_adaptive_optimize_l2; obs[n, 3, 3] = (centre, first, second) bearings"""
    arr = (Pose * 2)(make_pose(*poses[0]), make_pose(*poses[1]))
    o = np.ascontiguousarray(obs, np.float64).reshape(-1, 9)
    it = _opt().ref_three_view_optimize_l2(arr, int(adaptive), rate, iterations, _dp(o), len(o))
    return [arr[0].numpy(), arr[1].numpy()], it


def epipolar_loss(t, a, b):
    t, a, b = [np.ascontiguousarray(x, np.float64) for x in (t, a, b)]
    return _opt().ref_epipolar_loss(_dp(t), _dp(a), _dp(b))


def observation_losses(poses, bearings):
    arr = (Pose * len(poses))(*[make_pose(R, t) for R, t in poses])
    b = np.ascontiguousarray(bearings, np.float64)
    out = np.zeros(len(poses))
    _opt().ref_observation_losses(arr, _dp(b), len(poses), _dp(out))
    return out


def is_tri_landmark_robust(first, second, c, f, s, max_cos, inc_min_cos):
    c, f, s = [np.ascontiguousarray(x, np.float64) for x in (c, f, s)]
    return bool(_opt().ref_is_tri_landmark_robust(C.byref(make_pose(*first)), C.byref(make_pose(*second)), _dp(c), _dp(f), _dp(s),
                                                  max_cos, inc_min_cos))


def landmark_matches_ref(new_descriptors, views, better_by=24, landmark_views=None, landmark_observations=None):
    """Loop-for-loop restatement of cv-sfm/src/lib.rs:1468-1576 (matching part of register_frame_subset) over the exact
    LinearKnn; ties among equal distances broken by landmark id (the reference's HashMap order is unspecified).  Small cases only."""
    original = []
    for f in range(len(new_descriptors)):
        raw = []
        for desc, landmarks in views:
            idx, dist = hamming_knn(new_descriptors[f:f + 1], desc, 3)
            raw += [(int(landmarks[i]), int(d)) for i, d in zip(idx[0], dist[0]) if i != 0xFFFFFFFF]
        best_of = {}
        for l, d in raw:
            if l not in best_of or best_of[l] > d:
                best_of[l] = d
        items = sorted(best_of.items(), key=lambda ld: (ld[1], ld[0]))
        best = items[:3]
        assert len(best) == 3
        if best[0][1] + better_by <= best[1][1]:
            original.append(((best[0][0],), f))
        elif best[1][1] + better_by <= best[2][1]:
            a, b = best[0][0], best[1][0]
            if landmark_views is None or not (set(landmark_views[a]) & set(landmark_views[b])):
                original.append(((a, b), f))
    counts = {}
    for ls, _ in original:
        for l in ls:
            counts[l] = counts.get(l, 0) + 1
    kept = [m for m in original if all(counts[l] == 1 for l in m[0])]
    if landmark_observations is not None:
        kept = sorted(kept, key=lambda m: -sum(landmark_observations[l] for l in m[0]))
    return kept


def fp_o1(a, b):
    """nister-stewenius o1: product of two linear polynomials (x, y, z, 1 coefficients) in the 20-term basis"""
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64); r = np.zeros(20)
    L = _geom(); L.ref_fp_o1.argtypes = [C.POINTER(C.c_double)] * 3
    L.ref_fp_o1(_dp(a), _dp(b), _dp(r))
    return r


def fp_o2(a, b):
    """nister-stewenius o2: (degree <= 2 polynomial in the 20-term layout) x (linear polynomial)"""
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64); r = np.zeros(20)
    L = _geom(); L.ref_fp_o2.argtypes = [C.POINTER(C.c_double)] * 3
    L.ref_fp_o2(_dp(a), _dp(b), _dp(r))
    return r
