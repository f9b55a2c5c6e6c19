"""One bench step (2 x 1080p extract + symmetric 2-NN) for ncu captures: python scripts/prof_step.py [steps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv_b200  # noqa: E402
from cv_b200._lib import KP_DTYPE  # noqa: E402
from tests.synth import synth_frame, warp_frame  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
a = synth_frame(0)
pair = np.stack([a, warp_frame(a, 1000)])
dev = torch.device("cuda", 0)
ctx = cv_b200.Context(0)
cfg = cv_b200.AkazeConfig(maximum_features=5000).to_c()
cap = 8192
img = torch.from_numpy(pair).to(dev)
d_kp = torch.empty(2 * cap * KP_DTYPE.itemsize, dtype=torch.uint8, device=dev)
d_desc = torch.zeros(2 * cap * 64, dtype=torch.uint8, device=dev)
d_n = torch.zeros(2, dtype=torch.int32, device=dev)
d_i = torch.empty(cap * 2, dtype=torch.int32, device=dev)
d_d = torch.empty_like(d_i)
torch.cuda.synchronize()
for s in range(steps):
    ctx.check(ctx.lib.cvb_akaze_extract_batch_dev(ctx.handle, C.byref(cfg), img.data_ptr(), 2, 1920, 1080, d_kp.data_ptr(),
                                                  d_desc.data_ptr(), cap, d_n.data_ptr()))
    ctx.check(ctx.lib.cvb_hamming_knn_dev_counts(ctx.handle, d_desc.data_ptr(), d_n.data_ptr(), 5000, d_desc.data_ptr() + cap * 64,
                                                 d_n.data_ptr() + 4, 5000, 2, d_i.data_ptr(), d_d.data_ptr()))
    ctx.check(ctx.lib.cvb_hamming_knn_dev_counts(ctx.handle, d_desc.data_ptr() + cap * 64, d_n.data_ptr() + 4, 5000, d_desc.data_ptr(),
                                                 d_n.data_ptr(), 5000, 2, d_i.data_ptr(), d_d.data_ptr()))
ctx.sync()
print("keypoints", d_n.cpu().tolist())
