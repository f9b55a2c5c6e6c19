"""Multi-frame workload (BASELINE configs[3]): every rank extracts its own frames, ONE
all-gather moves the packed descriptors over NVLink, then the frame pairs are partitioned over the ranks and
matched with the brute-force symmetric matcher.  torch is plumbing here (device buffers, NCCL); all compute is
in libcvb200.so through device pointers.
"""
import ctypes as C

import numpy as np
import torch

from . import dist as D
from ._lib import KP_DTYPE, Context
from .akaze import AkazeConfig


def extract_frames_dev(ctx, config, images_dev, cap=8192):
    """images_dev: float32 cuda tensor [F, H, W].  Returns (kp uint8 [F, cap, 28], desc uint8 [F, cap, 64], n int32 [F]),
    all device tensors; asynchronous on the context stream (call ctx.sync() before reading them on another stream)."""
    assert images_dev.is_cuda and images_dev.dtype == torch.float32 and images_dev.dim() == 3 and images_dev.is_contiguous()
    F, H, W = images_dev.shape
    dev = images_dev.device
    kp = torch.empty((F, cap, KP_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    desc = torch.zeros((F, cap, 64), dtype=torch.uint8, device=dev)
    n = torch.zeros(F, dtype=torch.int32, device=dev)
    cfg = config.to_c() if isinstance(config, AkazeConfig) else config
    torch.cuda.current_stream(dev).synchronize()
    ctx.check(ctx.lib.cvb_akaze_extract_batch_dev(ctx.handle, C.byref(cfg), images_dev.data_ptr(), F, W, H, kp.data_ptr(),
                                                  desc.data_ptr(), cap, n.data_ptr()))
    return kp, desc, n


def match_pairs_dev(ctx, desc_all, counts_all, pairs, better_by=24):
    """Symmetric matches (cv-sfm rule d0 + better_by <= d1 + cross-check) for the given (i, j) frame pairs.
    desc_all: uint8 cuda [G, cap, 64]; counts_all: host list/array of G counts.  Returns {(i, j): int64 [K, 2]}."""
    dev = desc_all.device
    cap = desc_all.shape[1]
    out = {}
    flags = torch.empty(cap, dtype=torch.int32, device=dev)
    torch.cuda.current_stream(dev).synchronize()
    for (i, j) in pairs:
        ni, nj = int(counts_all[i]), int(counts_all[j])
        ctx.check(ctx.lib.cvb_match_symmetric_dev(ctx.handle, desc_all[i].data_ptr(), ni, desc_all[j].data_ptr(), nj, better_by,
                                                  flags.data_ptr()))
        ctx.sync()
        f = flags[:ni].cpu().numpy().view(np.uint32)
        a = np.where(f != 0xFFFFFFFF)[0]
        out[(i, j)] = np.stack([a.astype(np.int64), f[a].astype(np.int64)], 1) if len(a) else np.zeros((0, 2), np.int64)
    return out


def extract_and_match_all_pairs(ctx, config, my_images_dev, num_frames, cap=8192, better_by=24):
    """Whole config-4 step on this rank: extract local frames (round-robin shard of `num_frames`), all-gather the
    descriptors, match the pairs this rank owns.  Returns (counts_all, {(gi, gj): pairs}) in GLOBAL frame indices."""
    world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    kp, desc, n = extract_frames_dev(ctx, config, my_images_dev, cap)
    ctx.sync()
    desc_all, cnt_all = D.all_gather_descriptors(desc, n, cap)          # the single collective of the workload
    per = my_images_dev.shape[0]
    counts = cnt_all.cpu().numpy()
    # position in the gathered array of global frame g (owner rank g % world, local slot g // world)
    slot = lambda g: (g % world) * per + (g // world)
    mine = D.my_pairs(num_frames, rank, world)
    res = match_pairs_dev(ctx, desc_all, {slot(g): counts[slot(g)] for g in range(num_frames)}, [(slot(i), slot(j)) for i, j in mine], better_by)
    inv = {slot(g): g for g in range(num_frames)}
    return {g: int(counts[slot(g)]) for g in range(num_frames)}, {(inv[a], inv[b]): v for (a, b), v in res.items()}
