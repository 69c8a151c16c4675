"""Host-side multi-rank logic (problem sharding, max-over-ranks timing, final gather) on CPU with the gloo backend,
world_size = 2.  The data path itself has no collective (independent problems)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyneuralempc_b200.sharding import gather_solutions, max_over_ranks, shard_range


@pytest.mark.parametrize("total,world", [(4096, 2), (4096, 8), (7, 2), (5, 8), (0, 4), (1, 1)])
def test_shard_range_partitions_the_batch(total, world):
    spans = [shard_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (a0, a1), (b0, b1) in zip(spans[:-1], spans[1:]):
        assert a1 == b0 and a1 >= a0
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    # each rank "solves" its own problems: a deterministic function of the global problem index
    local = torch.arange(lo, hi, dtype=torch.float64).unsqueeze(1) * torch.tensor([[1.0, -2.0, 0.5]], dtype=torch.float64)
    slow = max_over_ranks(10.0 + rank)                       # timed region = slowest rank
    dist.barrier()
    full = gather_solutions(local, dst=0)
    q.put((rank, slow, None if full is None else full.numpy()))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world, total = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(abs(r[1] - 11.0) < 1e-12 for r in res)          # max over ranks, seen by every rank
    assert res[1][2] is None
    expect = np.arange(total)[:, None] * np.array([[1.0, -2.0, 0.5]])
    np.testing.assert_array_equal(res[0][2], expect)          # rank order == problem order, ragged shards
