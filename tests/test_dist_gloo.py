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


def _eval_worker(rank, world, port):
    """DispEvaluator.evaluate over 2 ranks: the per-image statistics of each rank (filled in directly -- the device kernel
    is covered by the GPU tests) must reduce to the mean over ALL images, skipping images without valid pixels"""
    from nmrf_b200.evaluation import DispEvaluator
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ev = DispEvaluator(["1.0"], True, 192)
    # rows: [n_valid, sum |e|, n_d1, n_bad1]; rank 0 has images A, B(no valid pixel); rank 1 has image C
    if rank == 0:
        ev._acc = [torch.tensor([[100.0, 50.0, 10.0, 20.0], [0.0, 0.0, 0.0, 0.0]], dtype=torch.float64)]
    else:
        ev._acc = [torch.tensor([[200.0, 300.0, 100.0, 50.0]], dtype=torch.float64)]
    res = ev.evaluate()["disp"]
    assert abs(res["epe"] - (0.5 + 1.5) / 2) < 1e-12
    assert abs(res["d1"] - 100 * (0.1 + 0.5) / 2) < 1e-9
    assert abs(res["bad 1.0"] - 100 * (0.2 + 0.25) / 2) < 1e-9
    dist.barrier()
    dist.destroy_process_group()


def test_evaluator_reduction_world2():
    mp.spawn(_eval_worker, args=(2, _free_port()), nprocs=2, join=True)
