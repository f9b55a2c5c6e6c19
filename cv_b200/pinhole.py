"""cv_pinhole::CameraIntrinsics (no distortion): pixel <-> unit bearing (cv-pinhole/src/lib.rs:32-148).

Scalar per-keypoint host work in the reference ("negligible; keep on host"); it is
vectorised here over all keypoints in f64 with the reference's operation order.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class CameraIntrinsics:
    focals: tuple            # (fx, fy)
    principal_point: tuple   # (cx, cy)
    skew: float = 0.0

    def calibrate(self, points):
        """CameraModel::calibrate (cv-pinhole/src/lib.rs:108-116): [N, 2] pixel coordinates -> [N, 3] unit bearings."""
        p = np.asarray(points, np.float64).reshape(-1, 2)
        y = (p[:, 1] - self.principal_point[1]) / self.focals[1]
        x = (p[:, 0] - self.principal_point[0] - self.skew * y) / self.focals[0]
        n = np.sqrt(x * x + y * y + 1.0)
        return np.stack([x / n, y / n, 1.0 / n], 1)

    def uncalibrate(self, bearings):
        """CameraModel::uncalibrate (cv-pinhole/src/lib.rs:134-141): unit bearings -> pixel coordinates
        (NaN where the reference returns None: z not sign-positive)."""
        b = np.asarray(bearings, np.float64).reshape(-1, 3)
        with np.errstate(divide="ignore", invalid="ignore"):
            x, y = b[:, 0] / b[:, 2], b[:, 1] / b[:, 2]
            px = x * self.focals[0] + self.skew * y + self.principal_point[0]
            py = y * self.focals[1] + self.principal_point[1]
        out = np.stack([px, py], 1)
        out[np.signbit(b[:, 2])] = np.nan
        return out

    def calibrate_keypoints(self, kps):
        """akaze::KeyPoint implements ImagePoint via (point.0 as f64, point.1 as f64) (akaze/src/lib.rs:95-99)."""
        return self.calibrate(np.stack([kps["x"].astype(np.float64), kps["y"].astype(np.float64)], 1))
