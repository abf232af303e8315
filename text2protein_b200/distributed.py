"""Batch sharding across GPUs (one process per GPU) for the sampler.

Sampling chains are independent (up to the batch-mean Langevin step size), so the batch is split into
contiguous shards, each rank runs ``pc_sampler`` on its shard with ``sample_offset`` = its first global sample
index (Philox noise is keyed by GLOBAL sample index), and the only collective is one all-gather of the final maps
(NCCL over NVLink on GPUs; gloo in the CPU tests).  Replaces ``nn.DataParallel`` (score_sde_pytorch/utils.py:8).

Opt-in: ``StepSizeSync`` makes the Langevin step size the mean over the GLOBAL batch (what one reference run of the
whole batch computes, sampling.py:193-195) -- two sums per corrector step, exchanged inside the fused corrector
kernel through NVLink peer memory (csrc/pc_step.cu ``exchange_sums``), no collective launch and no host round trip.
"""
import ctypes as C

import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous shard [start, stop) of ``total`` samples for ``rank``; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_samples(local, total, group=None):
    """All-gathers per-rank sample tensors [b_r, C, N, N] into [total, C, N, N] in global sample order: one
    ``all_gather_into_tensor`` straight into the result when the shards are equal, a padded gather otherwise."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    maxb = max(b - a for a, b in sizes)
    local = local.contiguous()
    if all(b - a == maxb for a, b in sizes):
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    pad = torch.zeros((maxb,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    buf = torch.empty((world * maxb,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * maxb: r * maxb + (b - a)] for r, (a, b) in enumerate(sizes)], dim=0)


def exchange_handles(handle_bytes, group=None):
    """Every rank's opaque IPC handle, in rank order, as one bytes object (host-side exchange; any backend)."""
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, bytes(handle_bytes), group=group)
    if any(len(p) != len(handle_bytes) for p in parts):
        raise RuntimeError("IPC handle size differs between ranks")
    return b"".join(parts)


class StepSizeSync:
    """Peer mailboxes of the ranks of one node for the global-batch Langevin step size.  Create one per rank after
    ``init_process_group`` (all ranks, collectively), pass it to ``get_pc_sampler(..., sync_step_size=obj)``; all
    ranks must then make the same sequence of sampler calls.  ``global_batch`` = total samples over all ranks."""

    def __init__(self, global_batch, group=None):
        from text2protein_b200 import _lib
        if not dist.is_initialized() or dist.get_world_size(group) < 2:
            raise ValueError("StepSizeSync needs an initialised process group with at least two ranks")
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        L = _lib.lib()
        box = C.c_void_p(0)
        mine = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        _lib.check(L.t2p_peer_mailbox_create(self.world, C.byref(box), mine))
        every = exchange_handles(mine.raw, group)
        h = C.c_void_p(0)
        _lib.check(L.t2p_peer_group_open(box, every, self.world, self.rank, int(global_batch), C.byref(h)))
        self.handle = h
        dist.barrier(group=group)  # every mailbox is mapped everywhere before the first kernel writes to a peer

    def close(self):
        from text2protein_b200 import _lib
        if self.handle:
            _lib.lib().t2p_peer_group_close(self.handle)
            self.handle = C.c_void_p(0)
