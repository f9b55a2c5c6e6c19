"""Multi-frame workload (BASELINE configs[3]): every rank extracts its own frames, ONE all-gather moves
the PACKED descriptors (count-sized, not capacity-sized) over NVLink, then the frame pairs are partitioned over the ranks and
matched with the brute-force symmetric matcher.  Pairs whose two frames are local are matched while the all-gather is in flight;
no pair synchronises with the host -- every pair's matches are compacted on the device into one result block that is read once.
torch is plumbing here (device buffers, streams, NCCL); all compute is in libcvb200.so through device pointers.
"""
import ctypes as C

import numpy as np
import torch

from . import dist as D
from ._lib import KP_DTYPE
from .akaze import AkazeConfig
from .pair import bind as _bind_pair


def _stream_of(ctx, dev):
    """the torch view of the context's CUDA stream (contexts made with Context(device, stream=...) share it with torch)"""
    s = getattr(ctx, "torch_stream", None)
    return s if s is not None else torch.cuda.current_stream(dev)


def make_context(device_index):
    """A cvb context that runs on a torch-visible stream, so that torch ops (NCCL, copies) and library kernels order by stream."""
    from ._lib import Context
    s = torch.cuda.Stream(device=device_index)
    ctx = Context(device_index, stream=s.cuda_stream)
    ctx.torch_stream = s
    return ctx


def extract_frames_dev(ctx, config, images_dev, cap=8192, batch=8):
    """images_dev: float32 cuda tensor [F, H, W].  Returns (kp uint8 [F, cap, 28], desc uint8 [F, cap, 64], n int32 [F]),
    all device tensors; asynchronous on the context stream."""
    assert images_dev.is_cuda and images_dev.dtype == torch.float32 and images_dev.dim() == 3 and images_dev.is_contiguous()
    F, H, W = images_dev.shape
    dev = images_dev.device
    with torch.cuda.stream(_stream_of(ctx, dev)):
        kp = torch.empty((F, cap, KP_DTYPE.itemsize), dtype=torch.uint8, device=dev)
        desc = torch.zeros((F, cap, 64), dtype=torch.uint8, device=dev)
        n = torch.zeros(F, dtype=torch.int32, device=dev)
    if getattr(ctx, "torch_stream", None) is None:
        torch.cuda.current_stream(dev).synchronize()
    cfg = config.to_c() if isinstance(config, AkazeConfig) else config
    for f0 in range(0, F, batch):          # one batched pass per `batch` frames (bounds the pyramid workspace: ~0.35 GB per 1080p frame)
        fb = min(batch, F - f0)
        ctx.check(ctx.lib.cvb_akaze_extract_batch_dev(ctx.handle, C.byref(cfg), images_dev[f0:].data_ptr(), fb, W, H, kp[f0:].data_ptr(),
                                                      desc[f0:].data_ptr(), cap, n[f0:].data_ptr()))
    return kp, desc, n


class PairResults:
    """Device-side result block of many symmetric matches: pairs[p, :count[p]] = (a, b) index pairs of frame pair p."""

    def __init__(self, npairs, cap, dev):
        self.pairs = torch.zeros((max(npairs, 1), cap, 2), dtype=torch.int32, device=dev)
        self.count = torch.zeros(max(npairs, 1), dtype=torch.int32, device=dev)
        self.cap = cap

    def host(self, keys):
        cnt = self.count.cpu().numpy()
        pr = self.pairs.cpu().numpy()
        return {k: pr[p, :cnt[p]].astype(np.int64) for p, k in enumerate(keys)}


def match_pairs_into(ctx, results, first, frames, pairs, better_by=24):
    """Enqueue the symmetric match (cv-sfm/src/lib.rs:3097-3133) of every (i, j) in `pairs` on the context stream, no host
    synchronisation.  frames[g] = (descriptor device pointer, count device pointer, capacity)."""
    L = ctx.lib
    _bind_pair(L)
    for k, (i, j) in enumerate(pairs):
        pi, ci, ni = frames[i]
        pj, cj, nj = frames[j]
        p = first + k
        ctx.check(L.cvb_match_symmetric_pairs_dev(ctx.handle, pi, ci, ni, pj, cj, nj, better_by,
                                                  results.pairs[p].data_ptr(), results.cap, results.count[p:].data_ptr()))


def pack_layout(counts):
    """counts: int array [world, per] (descriptors per frame, 0 for padding frames).  Returns (row offset of every frame inside its
    rank's pack [world, per + 1], rows of the largest pack): the all-gather moves world x maxtot x 64 bytes."""
    counts = np.asarray(counts, np.int64)
    off = np.concatenate([np.zeros((counts.shape[0], 1), np.int64), np.cumsum(counts, 1)], 1)
    return off, int(max(int(counts.sum(1).max()), 1))


def extract_and_match_all_pairs(ctx, config, my_images_dev, num_frames, cap=8192, better_by=24, timing=None):
    """Whole config-4 step on this rank: extract local frames (round-robin shard of `num_frames`), all-gather the packed
    descriptors, match the pairs this rank owns.  Returns (counts per global frame, {(gi, gj): pairs}) in GLOBAL frame indices.
    `timing` (dict, optional) receives CUDA-event milliseconds: extract, gather (the collective alone), match, total."""
    dist = torch.distributed
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = my_images_dev.device
    assert getattr(ctx, "torch_stream", None) is not None, "use cv_b200.multi.make_context (library kernels and torch ops must share a stream)"
    S = _stream_of(ctx, dev)
    ev = {k: torch.cuda.Event(enable_timing=True) for k in ("t0", "ext", "g0", "g1", "loc", "end")}
    per = -(-num_frames // world)                      # shards are padded to ceil(F / world) frames (count 0)
    mine = D.shard_frames(num_frames, rank, world)
    assert my_images_dev.shape[0] == len(mine)
    with torch.cuda.stream(S):
        ev["t0"].record(S)
        kp, desc, n = extract_frames_dev(ctx, config, my_images_dev, cap)
        ev["ext"].record(S)
        n_pad = torch.zeros(per, dtype=torch.int32, device=dev)
        n_pad[:len(mine)] = n
        # counts of every rank (tiny) -> host: packing offsets and launch bounds
        cnt_all = torch.empty(world * per, dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_gather_into_tensor(cnt_all, n_pad) if dist.get_backend() == "nccl" else dist.all_gather(list(cnt_all.view(world, per).unbind(0)), n_pad)
        else:
            cnt_all.copy_(n_pad)
        counts = cnt_all.cpu().numpy().reshape(world, per)          # the one host synchronisation before matching
        off, maxtot = pack_layout(counts)                              # row offsets inside a rank's pack
        # pack the local descriptors (count-sized)
        pack = torch.zeros((maxtot, 64), dtype=torch.uint8, device=dev)
        for l in range(len(mine)):
            c = int(counts[rank, l])
            if c:
                pack[int(off[rank, l]):int(off[rank, l]) + c] = desc[l, :c]
        gathered = torch.empty((world, maxtot, 64), dtype=torch.uint8, device=dev)
        ev["g0"].record(S)
    # the collective runs on its own stream; local-vs-local pairs are matched on the context stream meanwhile
    comm = torch.cuda.Stream(device=dev)
    comm.wait_stream(S)
    gdone = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(comm):
        g_start = torch.cuda.Event(enable_timing=True)
        g_start.record(comm)
        if world > 1:
            dist.all_gather_into_tensor(gathered, pack)            # THE collective of the workload: one all-gather over NVLink
        else:
            gathered[0].copy_(pack)
        gdone.record(comm)
    owner = lambda g: g % world
    local = lambda g: g // world
    my_pairs = D.my_pairs(num_frames, rank, world)
    loc_pairs = [(i, j) for i, j in my_pairs if owner(i) == rank and owner(j) == rank]
    rem_pairs = [(i, j) for i, j in my_pairs if not (owner(i) == rank and owner(j) == rank)]
    with torch.cuda.stream(S):
        cnt_dev = torch.from_numpy(counts.reshape(-1).astype(np.int32)).to(dev)      # device copy of every frame's count
        results = PairResults(len(my_pairs), cap, dev)
        f_local = {g: (desc[local(g)].data_ptr(), cnt_dev[owner(g) * per + local(g):].data_ptr(), max(int(counts[owner(g), local(g)]), 1))
                   for g in range(num_frames) if owner(g) == rank}
        match_pairs_into(ctx, results, 0, f_local, loc_pairs, better_by)
        ev["loc"].record(S)
        S.wait_event(gdone)
        f_all = {g: (gathered[owner(g), int(off[owner(g), local(g)]):].data_ptr(), cnt_dev[owner(g) * per + local(g):].data_ptr(),
                     max(int(counts[owner(g), local(g)]), 1)) for g in range(num_frames)}
        match_pairs_into(ctx, results, len(loc_pairs), f_all, rem_pairs, better_by)
        ev["end"].record(S)
    ctx.sync()
    comm.synchronize()
    out = results.host(loc_pairs + rem_pairs)
    if timing is not None:
        timing.update(extract_ms=ev["t0"].elapsed_time(ev["ext"]), gather_ms=g_start.elapsed_time(gdone),
                      match_ms=ev["g0"].elapsed_time(ev["end"]), total_ms=ev["t0"].elapsed_time(ev["end"]),
                      gather_bytes=int(world * maxtot * 64), pairs=len(my_pairs), local_pairs=len(loc_pairs),
                      comparisons=float(sum(2.0 * counts[owner(i), local(i)] * counts[owner(j), local(j)] for i, j in my_pairs)))
    return {g: int(counts[owner(g), local(g)]) for g in range(num_frames)}, out
