"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py (pair sharding + the single all_gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nmrf_b200.sharding import gather_stats, local_indices


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = list(local_indices(total, world, rank))
    vec = torch.tensor([len(idx), float(sum(idx)), 1.0 + rank], dtype=torch.float64)
    allv = gather_stats(vec)
    assert allv.shape == (world, 3)
    assert int(allv[:, 0].sum()) == total                              # every pair processed exactly once
    assert float(allv[:, 1].sum()) == float(sum(range(total)))
    assert float(allv[:, 2].max()) == float(world)                      # max-over-ranks timing reduction
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_gather_world2():
    mp.spawn(_worker, args=(2, _free_port(), 11), nprocs=2, join=True)


def test_local_indices_partition():
    for total in (0, 1, 7, 8, 33):
        for world in (1, 2, 4, 8):
            seen = [i for r in range(world) for i in local_indices(total, world, r)]
            assert seen == list(range(total))
            sizes = [len(local_indices(total, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_gather_stats_single_process():
    assert gather_stats(torch.tensor([1.0, 2.0], dtype=torch.float64)).tolist() == [[1.0, 2.0]]
