"""world_size-2 gloo test of the multi-GPU host logic (frame sharding, the descriptor all-gather, pair
partitioning, max-over-ranks timing) -- runs on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.synth import random_descriptors


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, num_frames, cap, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cv_b200 import dist as D
    mine = D.shard_frames(num_frames, rank, world)
    counts = torch.tensor([100 + 7 * g for g in mine], dtype=torch.int32)
    desc = torch.zeros((len(mine), cap, 64), dtype=torch.uint8)
    for l, g in enumerate(mine):
        desc[l, :counts[l]] = torch.from_numpy(random_descriptors(int(counts[l]), seed=g))
    desc_all, cnt_all = D.all_gather_descriptors(desc, counts, cap)
    # every rank now holds every frame's descriptors, addressable by global frame index
    ok = True
    per = len(mine)
    for r in range(world):
        for l in range(per):
            g = D.global_frame_index(r, l, world)
            n = 100 + 7 * g
            ok &= int(cnt_all[r * per + l]) == n
            ok &= bool(np.array_equal(desc_all[r * per + l, :n].numpy(), random_descriptors(n, seed=g)))
    pairs = D.my_pairs(num_frames, rank, world)
    value, ms = D.aggregate_throughput(2, 10.0 * (rank + 1), torch.device("cpu"))
    out[rank] = (ok, pairs, value, ms)
    dist.destroy_process_group()


def test_sharding_allgather_and_pair_partition_world2():
    world, num_frames, cap = 2, 8, 256
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), num_frames, cap, out), nprocs=world, join=True)
    assert all(out[r][0] for r in range(world))
    allpairs = sorted(p for r in range(world) for p in out[r][1])
    assert allpairs == [(i, j) for i in range(num_frames) for j in range(i + 1, num_frames)]      # each pair exactly once
    assert abs(len(out[0][1]) - len(out[1][1])) <= num_frames                                       # balanced
    # max-over-ranks timing: rank 1 took 20 ms -> 2 ranks * 2 units / 0.020 s
    assert all(abs(out[r][3] - 20.0) < 1e-9 and abs(out[r][2] - 200.0) < 1e-6 for r in range(world))


def test_shard_frames_round_robin():
    from cv_b200 import dist as D
    assert D.shard_frames(64, 3, 8) == list(range(3, 64, 8))
    assert sorted(sum((D.shard_frames(10, r, 4) for r in range(4)), [])) == list(range(10))
