"""ctypes binding of cv_b200/libcvb200.so (the C ABI declared in include/cvb200.h)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CVB_OK, CVB_EINVAL, CVB_ENODEV, CVB_ECUDA, CVB_ENOMEM, CVB_ECAP, CVB_EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
_ERRNAMES = {-1: "EINVAL", -2: "ENODEV", -3: "ECUDA", -4: "ENOMEM", -5: "ECAP", -6: "EUNSUPPORTED"}


class CvbError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"cvb200 error {_ERRNAMES.get(code, code)}: {msg}")
        self.code = code


class AkazeCfg(C.Structure):
    """cvb_akaze_cfg == akaze::Akaze (akaze/src/lib.rs:109-142)"""
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


# cvb_keypoint == akaze::KeyPoint (akaze/src/lib.rs:71-93)
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("response", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("octave", "<u4"), ("class_id", "<u4")])

# every symbol include/cvb200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "cvb_ctx_create", "cvb_ctx_create_on_stream", "cvb_ctx_destroy", "cvb_ctx_sync", "cvb_last_error", "cvb_version",
    "cvb_ctx_launch_count", "cvb_ctx_timer_begin", "cvb_ctx_timer_end", "cvb_ctx_profile", "cvb_ctx_profile_report",
    "cvb_akaze_default_cfg", "cvb_akaze_extract", "cvb_akaze_extract_batch", "cvb_akaze_extract_batch_dev", "cvb_akaze_dev_overflow",
    "cvb_akaze_debug_num_evolutions", "cvb_akaze_debug_evolution", "cvb_akaze_debug_plane", "cvb_akaze_debug_contrast",
    "cvb_akaze_debug_stage",
    "cvb_hamming_knn", "cvb_hamming_knn_dev", "cvb_hamming_knn_dev_counts", "cvb_match_symmetric", "cvb_match_symmetric_dev",
    "cvb_arrsac_default_cfg", "cvb_rng_seed_xoshiro256pp", "cvb_rng_seed_pcg64", "cvb_rng_next_u32",
    "cvb_eight_point_batch", "cvb_p3p_batch", "cvb_five_point_batch", "cvb_arrsac_five_point", "cvb_residuals_camera_to_camera", "cvb_residuals_world_to_camera",
    "cvb_triangulate_linear_eigen", "cvb_arrsac_eight_point", "cvb_arrsac_p3p",
    "cvb_hash_bag", "cvb_hash_bag_dev", "cvb_match_symmetric_pairs_dev", "cvb_pair_bearings_dev", "cvb_arrsac_eight_point_dev", "cvb_arrsac_p3p_dev", "cvb_arrsac_commit_rng",
    "cvb_two_view_pair_dev", "cvb_two_view_frames",
    "cvb_single_view_optimize_l2", "cvb_three_view_optimize_l2", "cvb_observation_losses", "cvb_tri_landmarks_robust",
]


def lib_path():
    return os.path.join(_HERE, "libcvb200.so")


def load_library():
    """Loads the CUDA extension. Fails loudly when it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise CvbError(CVB_ENODEV, f"{p} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                   f"or `make -C cv_b200/csrc`")
    # one hardware queue per stream when many contexts are pipelined (effective only if CUDA is not initialised yet)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    L = C.CDLL(p)
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    L.cvb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.cvb_ctx_create_on_stream.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.cvb_ctx_destroy.argtypes = [vp]
    L.cvb_ctx_destroy.restype = None
    L.cvb_ctx_sync.argtypes = [vp]
    L.cvb_last_error.argtypes = [vp]
    L.cvb_last_error.restype = C.c_char_p
    L.cvb_version.restype = C.c_char_p
    L.cvb_ctx_launch_count.argtypes = [vp]
    L.cvb_ctx_launch_count.restype = u64
    L.cvb_ctx_timer_begin.argtypes = [vp]
    L.cvb_ctx_timer_end.argtypes = [vp, C.POINTER(C.c_float)]
    L.cvb_ctx_profile.argtypes = [vp, C.c_int]
    L.cvb_ctx_profile_report.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.cvb_akaze_default_cfg.argtypes = [C.POINTER(AkazeCfg)]
    L.cvb_akaze_default_cfg.restype = None
    L.cvb_akaze_extract.argtypes = [vp, C.POINTER(AkazeCfg), vp, u32, u32, vp, vp, u32, C.POINTER(u32)]
    L.cvb_akaze_extract_batch.argtypes = [vp, C.POINTER(AkazeCfg), vp, u32, u32, u32, vp, vp, u32, vp]
    L.cvb_akaze_extract_batch_dev.argtypes = [vp, C.POINTER(AkazeCfg), vp, u32, u32, u32, vp, vp, u32, vp]
    L.cvb_akaze_dev_overflow.argtypes = [vp, C.POINTER(u32)]
    L.cvb_akaze_debug_num_evolutions.argtypes = [vp, C.POINTER(u32)]
    L.cvb_akaze_debug_evolution.argtypes = [vp, u32] + [C.POINTER(u32)] * 5
    L.cvb_akaze_debug_plane.argtypes = [vp, u32, u32, u32, vp]
    L.cvb_akaze_debug_contrast.argtypes = [vp, u32, C.POINTER(C.c_double)]
    L.cvb_akaze_debug_stage.argtypes = [vp, u32, u32, vp, u32, C.POINTER(u32)]
    L.cvb_hamming_knn.argtypes = [vp, vp, u32, vp, u32, u32, vp, vp]
    L.cvb_hamming_knn_dev.argtypes = [vp, vp, u32, vp, u32, u32, vp, vp]
    L.cvb_hamming_knn_dev_counts.argtypes = [vp, vp, vp, u32, vp, vp, u32, u32, vp, vp]
    L.cvb_match_symmetric.argtypes = [vp, vp, u32, vp, u32, u32, vp, u32, C.POINTER(u32)]
    L.cvb_match_symmetric_dev.argtypes = [vp, vp, u32, vp, u32, u32, vp]
    L.cvb_hash_bag.argtypes = [vp, vp, u32, vp, u32, vp]
    L.cvb_hash_bag_dev.argtypes = [vp, vp, vp, u32, vp, u32, vp]
    _LIB = L
    return L


class Context:
    """A cvb_ctx: one CUDA stream + workspaces on one device. Not thread-safe; use one per thread."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.cvb_ctx_create_on_stream(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise CvbError(rc, "cvb_ctx_create failed (a Blackwell-class CUDA device is required; there is no CPU fallback)")
        self.handle = h
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.lib.cvb_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise CvbError(rc, self.lib.cvb_last_error(self.handle).decode())

    def sync(self):
        self.check(self.lib.cvb_ctx_sync(self.handle))

    def launch_count(self):
        return int(self.lib.cvb_ctx_launch_count(self.handle))

    def profile(self, enable=True):
        self.check(self.lib.cvb_ctx_profile(self.handle, 1 if enable else 0))

    def profile_report(self):
        """{kernel: dict(launches, ms, bytes)} accumulated since profile(True)."""
        buf = C.create_string_buffer(1 << 16)
        self.check(self.lib.cvb_ctx_profile_report(self.handle, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms, by = line.split()
            out[name] = dict(launches=int(n), ms=float(ms), bytes=float(by))
        return out

    def timer_begin(self):
        self.check(self.lib.cvb_ctx_timer_begin(self.handle))

    def timer_end(self):
        ms = C.c_float()
        self.check(self.lib.cvb_ctx_timer_end(self.handle, C.byref(ms)))
        return ms.value


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
