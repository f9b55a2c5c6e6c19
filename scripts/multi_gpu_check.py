"""torchrun --nproc-per-node N scripts/multi_gpu_check.py : config-4 style step over NCCL, checked against a
single-rank recomputation (rank 0 re-extracts every frame locally and compares descriptor counts and pair results)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv_b200  # noqa: E402
from cv_b200 import dist as D, multi  # noqa: E402
from tests.synth import synth_frame, warp_frame  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
F = 2 * world + 1          # uneven shards: the last rank(s) hold one frame less (padded)
base = synth_frame(11, h=360, w=480, nblobs=900)
frames = [base] + [warp_frame(base, g, shift=(1.0 + g, 0.5 * g)) for g in range(1, F)]
mine = D.shard_frames(F, rank, world)
ctx = multi.make_context(lr)
cfg = cv_b200.AkazeConfig(detector_threshold=0.001)
tm = {}
counts, res = multi.extract_and_match_all_pairs(ctx, cfg, torch.from_numpy(np.stack([frames[g] for g in mine])).cuda(), num_frames=F, cap=4096, timing=tm)
gathered = [None] * world
dist.all_gather_object(gathered, {k: v.tolist() for k, v in res.items()})
if rank == 0:
    allres = {}
    for d in gathered:
        allres.update(d)
    assert sorted(allres) == [(i, j) for i in range(F) for j in range(i + 1, F)], sorted(allres)
    ak = cv_b200.Akaze(0.001, ctx=ctx)
    descs = [ak.extract_from_gray_float_image(f)[1] for f in frames]
    assert [counts[g] for g in range(F)] == [len(d) for d in descs]
    for (i, j), pairs in allres.items():
        assert cv_b200.symmetric_matching(descs[i], descs[j], 24, ctx=ctx).tolist() == pairs, (i, j)
    print(f"multi-GPU check ok: world {world}, {F} frames, {len(allres)} pairs, one NCCL all-gather of {tm['gather_bytes']} B in {tm['gather_ms']:.3f} ms; rank 0 timing {tm}")
dist.destroy_process_group()
