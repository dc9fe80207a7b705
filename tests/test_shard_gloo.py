"""CPU: the N>1 path's host logic with world_size = 2 over gloo (the GPU path uses NCCL with the same
code): instance sharding, the single all-gather of commitments (equal and ragged shards), and the
max-over-ranks timing reduction."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b2rsa import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_commitment(i: int, w: int) -> torch.Tensor:
    return torch.arange(w, dtype=torch.int64) * 1000003 + i * 7919


def _worker(rank, world, port, total, w, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.instance_range(rank, world, total)
        local = torch.stack([_fake_commitment(i, w) for i in mine]) if len(mine) else torch.zeros((0, w), dtype=torch.int64)
        full = shard.gather_commitments(local, total)
        want = torch.stack([_fake_commitment(i, w) for i in range(total)])
        ok = bool(torch.equal(full, want))
        t = shard.max_over_ranks(10.0 + rank, torch.device("cpu"))
        q.put((rank, ok, t))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7, 2])
def test_allgather_commitments_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, 40, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [11.0, 11.0]


def test_instance_range_partitions():
    for world in (1, 2, 3, 8):
        for total in (0, 1, 7, 64, 512):
            got = [i for r in range(world) for i in shard.instance_range(r, world, total)]
            assert got == list(range(total))
            sizes = [len(shard.instance_range(r, world, total)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert list(shard.instance_range(3, 8, 512)) == list(range(192, 256))   # BASELINE config 4: 64 per GPU
    with pytest.raises(ValueError):
        shard.instance_range(2, 2, 10)


def test_single_process_passthrough():
    x = torch.arange(12, dtype=torch.int64).reshape(3, 4)
    assert shard.gather_commitments(x) is x
    assert shard.max_over_ranks(3.5, torch.device("cpu")) == 3.5
