"""Multi-frame path on one GPU (world size 1: every pair is local) against the oracle; the NCCL exchange itself is
exercised by scripts/multi_gpu_check.py under torchrun on >= 2 GPUs, and its host logic by tests/test_dist_gloo.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_extract_and_match_all_pairs_single_rank():
    import torch
    import cv_b200
    from cv_b200 import multi
    from oracle import pyoracle as O
    from tests.synth import synth_frame, warp_frame
    base = synth_frame(3, h=360, w=480, nblobs=900)
    frames = np.stack([base, warp_frame(base, 1, shift=(2.2, 1.1)), warp_frame(base, 2, shift=(-3.1, 0.7)), synth_frame(4, h=360, w=480, nblobs=900)])
    ctx = multi.make_context(0)          # library kernels and torch ops (packing, NCCL) share this stream
    cfg = cv_b200.AkazeConfig(detector_threshold=0.001)
    counts, res = multi.extract_and_match_all_pairs(ctx, cfg, torch.from_numpy(frames).cuda(), num_frames=4, cap=4096)
    descs = []
    for f in frames:
        kps, d = O.Akaze(detector_threshold=0.001).extract(f)
        descs.append(d)
    assert [counts[g] for g in range(4)] == [len(d) for d in descs]
    assert sorted(res.keys()) == [(i, j) for i in range(4) for j in range(i + 1, 4)]
    for (i, j), pairs in res.items():
        fi, fd = O.hamming_knn(descs[i], descs[j], 2)
        ri, rd = O.hamming_knn(descs[j], descs[i], 2)
        fwd = np.where(fd[:, 0] + 24 <= fd[:, 1], fi[:, 0].astype(np.int64), -1)
        rev = np.where(rd[:, 0] + 24 <= rd[:, 1], ri[:, 0].astype(np.int64), -1)
        want = [[a, int(b)] for a, b in enumerate(fwd) if b >= 0 and rev[b] == a]
        assert pairs.tolist() == want, (i, j)
    assert len(res[(0, 1)]) > 5 * max(len(res[(0, 3)]), 1)       # warped views match, unrelated frames do not
