"""cv_b200 -- B200-native (sm_100a) drop-in for rust-cv's AKAZE -> Hamming match -> RANSAC hot path.

Host-side mirror of the reference's interfaces for this path, over the C ABI in include/cvb200.h:

  Akaze                 <- akaze::Akaze                       (akaze/src/lib.rs:109-185, 295-366)
  KeyPoint dtype        <- akaze::KeyPoint                    (akaze/src/lib.rs:71-93)
  LinearKnn / hamming_knn <- space::LinearKnn + bitarray::Hamming (call sites akaze/tests/estimate_pose.rs:78-97)
  matching / symmetric_matching <- cv-sfm/src/lib.rs:3097-3133, tutorial-code chapter4 main.rs:91-137

There is no CPU fallback: every call runs CUDA kernels from cv_b200/libcvb200.so and raises
CvbError when the library or a Blackwell GPU is missing.
"""
from ._lib import CvbError, Context, KP_DTYPE, lib_path, load_library  # noqa: F401
from .akaze import Akaze, AkazeConfig  # noqa: F401
from .knn import HammingHasher, LinearKnn, hamming_knn, lowe_ratio_matches, matching, symmetric_matching  # noqa: F401
from .pinhole import CameraIntrinsics  # noqa: F401
from .geom import (Arrsac, EightPoint, LambdaTwist, LinearEigenTriangulator, NisterStewenius, Pcg64, Xoshiro256PlusPlus,  # noqa: F401
                   residuals_camera_to_camera, residuals_world_to_camera)
from .optimize import (observation_losses, single_view_simple_optimize_l2, single_view_simple_optimize_l2_batch,  # noqa: F401
                       three_view_adaptive_optimize_l2, three_view_optimize_l2_batch, three_view_simple_optimize_l2,
                       tri_landmarks_robust)
from .sfm_match import landmark_matches  # noqa: F401
from . import checkpoint  # noqa: F401  (bincode record images of the VSlamData checkpoint)
from .pair import Intrinsics, TwoViewBuffers, two_view_frames  # noqa: F401

__version__ = "0.1.0"
