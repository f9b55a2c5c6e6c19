"""Small extraction with TMA-staged tiles (CVB_TMA=1) for compute-sanitizer: python scripts/tma_probe.py [w h batch]"""
import os, sys
os.environ.setdefault("CVB_TMA", "1")
os.environ.setdefault("CVB_NO_GRAPH", "1")
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv_b200
from tests.synth import synth_frame
w, h, B = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (320, 240, 2)
frames = np.stack([synth_frame(7 + i, h=h, w=w, nblobs=max(50, w * h // 200)) for i in range(B)])
ak = cv_b200.Akaze(0.002)
k, d = ak.extract_batch(frames)
print("mask", os.environ.get("CVB_TMA_MASK"), "keypoints", [len(x) for x in d])
