"""Host-side mirror of akaze::Akaze (akaze/src/lib.rs:109-185, 295-366) over the C ABI."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import KP_DTYPE, AkazeCfg, Context, CvbError, default_context

PLANES = {"Lt": 0, "Lsmooth": 1, "Lx": 2, "Ly": 3, "Lflow": 4, "Ldet": 5}
STAGES = {"candidates": 0, "extrema": 1, "refined": 2, "sorted": 3}


@dataclass
class AkazeConfig:
    """The 11 public fields of akaze::Akaze with Akaze::default() values (lib.rs:169-185)."""
    maximum_features: int = -1          # usize::MAX
    num_sublevels: int = 4
    max_octave_evolution: int = 4
    base_scale_offset: float = 1.6
    initial_contrast: float = 0.001
    contrast_percentile: float = 0.7
    contrast_factor_num_bins: int = 300
    derivative_factor: float = 1.5
    detector_threshold: float = 0.001
    descriptor_channels: int = 3
    descriptor_pattern_size: int = 10

    def to_c(self):
        c = AkazeCfg()
        for f, _ in AkazeCfg._fields_:
            setattr(c, f, getattr(self, f))
        return c


class Akaze:
    """akaze::Akaze.  `Akaze(threshold)` == Akaze::new, plus `sparse()` / `dense()` (lib.rs:147-166)."""

    def __init__(self, detector_threshold=None, ctx=None, max_keypoints=32768, **fields):
        self.config = AkazeConfig(**fields)
        if detector_threshold is not None:
            self.config.detector_threshold = float(detector_threshold)
        self.ctx = ctx
        self.max_keypoints = int(max_keypoints)

    @classmethod
    def sparse(cls, **kw):
        return cls(0.01, **kw)

    @classmethod
    def dense(cls, **kw):
        return cls(0.0001, **kw)

    def _ctx(self):
        if self.ctx is None:
            self.ctx = default_context(0)
        return self.ctx

    # -- Akaze::extract (lib.rs:295): DynamicImage -> GrayFloatImage::from_dynamic (image.rs:45-109)
    def extract(self, image):
        image = np.asarray(image)
        if image.dtype == np.uint8:
            f = image.astype(np.float32) / np.float32(255)
        elif image.dtype == np.uint16:
            f = image.astype(np.float32) / np.float32(65535)
        elif image.dtype == np.float32:
            f = image
        else:
            raise TypeError("DynamicImage::grayscale() returned unexpected type")   # image.rs:107
        if f.ndim != 2:
            raise ValueError("expected a single-channel (luma) image")
        return self.extract_from_gray_float_image(f)

    # -- Akaze::extract_from_gray_float_image (lib.rs:309-339)
    def extract_from_gray_float_image(self, float_image):
        kps, descs = self.extract_batch(np.asarray(float_image, dtype=np.float32)[None])
        return kps[0], descs[0]

    def extract_batch(self, images):
        """B independent frames of one size in a single pass. Returns lists of (keypoints, descriptors)."""
        images = np.ascontiguousarray(images, dtype=np.float32)
        if images.ndim != 3:
            raise ValueError("images must be [B, H, W] float32")
        B, H, W = images.shape
        ctx = self._ctx()
        cap = self.max_keypoints
        kp = np.zeros((B, cap), dtype=KP_DTYPE)
        desc = np.zeros((B, cap, 64), dtype=np.uint8)
        n = np.zeros(B, dtype=np.uint32)
        cfg = self.config.to_c()
        rc = ctx.lib.cvb_akaze_extract_batch(ctx.handle, C.byref(cfg), images.ctypes.data, B, W, H, kp.ctypes.data,
                                             desc.ctypes.data, cap, n.ctypes.data)
        ctx.check(rc)
        return [kp[b, :n[b]].copy() for b in range(B)], [desc[b, :n[b]].copy() for b in range(B)]

    # -- introspection used by the parity tests (no reference counterpart)
    def debug_evolutions(self):
        ctx = self._ctx()
        n = C.c_uint32()
        ctx.check(ctx.lib.cvb_akaze_debug_num_evolutions(ctx.handle, C.byref(n)))
        out = []
        for i in range(n.value):
            v = [C.c_uint32() for _ in range(5)]
            ctx.check(ctx.lib.cvb_akaze_debug_evolution(ctx.handle, i, *[C.byref(x) for x in v]))
            out.append(dict(w=v[0].value, h=v[1].value, octave=v[2].value, sigma_size=v[3].value, n_fed_steps=v[4].value))
        return out

    def debug_plane(self, evolution, name, frame=0):
        ctx = self._ctx()
        info = self.debug_evolutions()[evolution]
        out = np.empty((info["h"], info["w"]), np.float32)
        ctx.check(ctx.lib.cvb_akaze_debug_plane(ctx.handle, frame, evolution, PLANES[name], out.ctypes.data))
        return out

    def debug_contrast(self, frame=0):
        ctx = self._ctx()
        k = C.c_double()
        ctx.check(ctx.lib.cvb_akaze_debug_contrast(ctx.handle, frame, C.byref(k)))
        return k.value

    def debug_stage(self, name, frame=0, cap=1 << 20):
        ctx = self._ctx()
        n = C.c_uint32()
        ctx.check(ctx.lib.cvb_akaze_debug_stage(ctx.handle, frame, STAGES[name], None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=KP_DTYPE)
        ctx.check(ctx.lib.cvb_akaze_debug_stage(ctx.handle, frame, STAGES[name], out.ctypes.data, n.value, C.byref(n)))
        return out[:n.value]


__all__ = ["Akaze", "AkazeConfig", "CvbError", "Context"]
