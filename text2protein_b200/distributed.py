"""Batch sharding across GPUs (one process per GPU) for the sampler.

Sampling chains are independent (up to the batch-mean Langevin step size), so the batch is split into
contiguous shards, each rank runs ``pc_sampler`` on its shard with ``sample_offset`` = its first global sample
index (Philox noise is keyed by GLOBAL sample index), and the only collective is one all-gather of the final maps
(NCCL over NVLink on GPUs; gloo in the CPU tests).  Replaces ``nn.DataParallel`` (score_sde_pytorch/utils.py:8).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous shard [start, stop) of ``total`` samples for ``rank``; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_samples(local, total):
    """All-gathers per-rank sample tensors [b_r, C, N, N] into [total, C, N, N] in global sample order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(total, r, world) for r in range(world)]
    maxb = max(b - a for a, b in sizes)
    pad = torch.zeros((maxb,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, sizes)], dim=0)
