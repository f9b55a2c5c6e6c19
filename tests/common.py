"""Shared helpers for the test-suite: golden fixtures and seeded synthetic inputs."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def kitti_frame(name):
    """res/<name>.png as GrayFloatImage::from_dynamic would load it (akaze/src/image.rs:53-55)."""
    im = np.load(os.path.join(GOLDEN, f"kitti_{name}.npz"))["image"]
    return im.astype(np.float32) / np.float32(255)


def goldens():
    with open(os.path.join(GOLDEN, "akaze_goldens.json")) as f:
        return json.load(f)


def lowe_matches(dist, ratio=0.5):
    """akaze/tests/estimate_pose.rs:91-93"""
    return int((dist[:, 0].astype(np.float32) < dist[:, 1].astype(np.float32) * np.float32(ratio)).sum())
