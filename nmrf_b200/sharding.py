"""Multi-GPU: independent stereo pairs sharded over ranks (one process per GPU, full weight replica
each); the only collective is one all_gather of a small per-rank fp64 vector AFTER the timed region.

Mirrors the reference's evaluation sharding (InferenceSampler._get_local_indices,
nmrf/utils/evaluation.py:62-69) and replaces its gloo gather_object (nmrf/utils/dist_utils.py:142-171).
"""
import torch
import torch.distributed as dist


def local_indices(total, world_size, rank):
    """contiguous ranges, the first (total % world) ranks get one extra item (evaluation.py:62-69)"""
    base, extra = divmod(total, world_size)
    sizes = [base + (r < extra) for r in range(world_size)]
    begin = sum(sizes[:rank])
    return range(begin, min(begin + sizes[rank], total))


def gather_stats(vec, device=None):
    """vec: 1-D float64 tensor (same length on every rank) -> [world, len] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return vec.reshape(1, -1).clone()
    world = dist.get_world_size()
    v = vec.to(device if device is not None else vec.device, torch.float64).contiguous()
    out = [torch.empty_like(v) for _ in range(world)]
    dist.all_gather(out, v)
    return torch.stack(out).cpu()
