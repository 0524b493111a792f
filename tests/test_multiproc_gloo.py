"""CPU, world_size 2, gloo: the data-parallel host logic of the path -- contiguous batch sharding, the SUM
all-reduce of the pose-gradient check-sum, the MIN all-reduce for the batch-global normal shift and the
max-over-ranks timing rule (bench.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from self6dpp_b200 import dist_utils as du


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        gR = torch.randn(total, 3, 3, generator=g)          # the "single-process" gradients, same on every rank
        gt = torch.randn(total, 3, generator=g)
        lo, hi = du.shard_range(total, rank, world)
        vec = du.pose_grad_checksum(gR[lo:hi], gt[lo:hi])
        du.allreduce_sum(vec)
        full = du.pose_grad_checksum(gR, gt)
        ok_sum = torch.allclose(vec, full, atol=1e-5)
        mn = torch.tensor([float(gR[lo:hi].min())])
        du.allreduce_min_scalar(mn)
        ok_min = abs(float(mn) - float(gR.min())) < 1e-12
        t = du.max_over_ranks(10.0 + rank, torch.device("cpu"))
        out[rank] = (lo, hi, bool(ok_sum), bool(ok_min), t)
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_everything():
    for total in (0, 1, 31, 32, 512, 513):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = du.shard_range(total, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(total))
            sizes = [du.shard_range(total, r, world)[1] - du.shard_range(total, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_allreduce_matches_single_process():
    world, total = 2, 37
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
    assert sorted(out.keys()) == [0, 1]
    assert out[0][:2] == (0, 19) and out[1][:2] == (19, 37)
    for r in range(world):
        assert out[r][2] and out[r][3]
        assert out[r][4] == 11.0            # max over ranks of (10 + rank)
