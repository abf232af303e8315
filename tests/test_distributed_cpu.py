"""world_size-2 gloo test of the multi-GPU host logic: shard bookkeeping, global-index noise keys, all-gather order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.philox_ref import philox_normal
from text2protein_b200.distributed import exchange_handles, gather_samples, shard_range


def test_shard_ranges_cover_batch():
    for total in (1, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(total, rank, world)
    per = 5 * 8 * 8
    # what a rank's sampler draws: noise of ITS samples under the GLOBAL element indexing (sample_offset = a)
    local = torch.from_numpy(philox_normal(11, 3, a * per, (b - a) * per)).reshape(b - a, 5, 8, 8)
    full = gather_samples(local, total)
    # the opaque per-rank handles of StepSizeSync travel in rank order, whatever the backend
    handles = exchange_handles(bytes([rank]) * 64)
    assert handles == b"".join(bytes([r]) * 64 for r in range(world))
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("total", [5, 6])  # uneven shards (3 + 2: padded gather) and equal shards (one all_gather_into_tensor)
def test_two_rank_gather_equals_single_process_noise(total):
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = philox_normal(11, 3, 0, total * 5 * 8 * 8).reshape(total, 5, 8, 8)
    assert np.array_equal(got, want)  # sharding does not change which normals a sample sees
