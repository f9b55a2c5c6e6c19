"""Multi-GPU plumbing for the hot path (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

The path shards by FRAME: extraction needs no communication; all-pairs matching over a
frame set needs exactly one exchange -- an all-gather of each rank's packed descriptors -- after which frame
pairs are partitioned over ranks.  Nothing here touches pixels or descriptors arithmetic; it is the host logic
that bench.py and a multi-frame matcher share, and it is covered by world_size-2 gloo tests.
"""
import torch
import torch.distributed as dist


def shard_frames(num_frames, rank, world):
    """One-frame-per-GPU round robin: the frame indices owned by `rank`."""
    return list(range(rank, num_frames, world))


def pair_owner(i, j, num_frames, world):
    """Owner rank of the (i, j) frame-pair tile: (i * F + j) mod world."""
    return (i * num_frames + j) % world


def my_pairs(num_frames, rank, world):
    """All unordered frame pairs i < j owned by `rank`; the union over ranks is every pair exactly once."""
    return [(i, j) for i in range(num_frames) for j in range(i + 1, num_frames) if pair_owner(i, j, num_frames, world) == rank]


def all_gather_descriptors(desc, counts, cap):
    """Single all-gather of packed descriptors.

    desc:   uint8 [F_local, cap, 64]  (device of the backend: cuda for nccl, cpu for gloo)
    counts: int32 [F_local]           valid descriptors per local frame
    Returns (desc_all [world * F_local, cap, 64], counts_all [world * F_local]) ordered by (rank, local frame),
    i.e. global frame index g = r + world * l for round-robin sharding is recovered by `global_frame_index`.
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return desc, counts
    desc = desc.contiguous(); counts = counts.contiguous()
    out = torch.empty((world,) + tuple(desc.shape), dtype=desc.dtype, device=desc.device)
    cnt = torch.empty((world,) + tuple(counts.shape), dtype=counts.dtype, device=counts.device)
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(out, desc)          # one NCCL all-gather over NVLink
        dist.all_gather_into_tensor(cnt, counts)
    else:                                                # gloo (CPU tests): same semantics through the list API
        dist.all_gather(list(out.unbind(0)), desc)
        dist.all_gather(list(cnt.unbind(0)), counts)
    return out.reshape((-1,) + tuple(desc.shape[1:])), cnt.reshape(-1)


def global_frame_index(rank, local, world):
    return rank + world * local


def max_over_ranks_ms(ms, device):
    """Device time of a multi-GPU step is the max over ranks (never wall clock of one rank)."""
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank, ms_local, device):
    """Whole-job throughput: units all ranks processed / max-over-ranks time (weak scaling: fixed work per rank)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    ms = max_over_ranks_ms(ms_local, device)
    return world * units_per_rank / (ms * 1e-3), ms
