"""World-size-2 gloo test of the batch sharding + final gather (host logic of the multi-GPU path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200edit.distributed import gather_images, sample_seeds, shard_range


def test_shard_ranges_cover_batch():
    for n in (1, 7, 8, 64, 255):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert sample_seeds(100, 3, 6) == [103, 104, 105]


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, world, rank)
    # each sample is produced from its own seed, independent of the sharding
    local = torch.stack([torch.randn(3, 4, 4, generator=torch.Generator().manual_seed(s))
                         for s in sample_seeds(100, lo, hi)])
    full = gather_images(local, n_total)
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_total = 5   # ragged: 3 + 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = torch.stack([torch.randn(3, 4, 4, generator=torch.Generator().manual_seed(100 + i)) for i in range(n_total)])
    for r in range(2):
        assert torch.equal(results[r], expect)
