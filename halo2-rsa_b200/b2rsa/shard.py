"""Multi-GPU plumbing for the proving hot path: proofs are independent (one create_proof call per
instance, reference benches/bench.rs:319-331), so a batch is sharded across ranks with no data-path
collective; the only exchange is ONE all-gather of the per-proof commitment block at the end
(SURVEY.md 8e).  One process per GPU, torch.distributed (NCCL on CUDA tensors; gloo in CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def instance_range(rank: int, world: int, total: int) -> range:
    """contiguous shard of [0, total): the first (total % world) ranks take one extra instance"""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_commitments(local: torch.Tensor, total: int | None = None, group=None) -> torch.Tensor:
    """local: int64[m_local, w] (w = 8 words per affine point x points per proof) -> int64[total, w]
    in instance order on every rank.  Equal shards use one all_gather_into_tensor; ragged shards
    (total % world != 0) are padded to the largest shard first."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    m_local, w = local.shape
    total = world * m_local if total is None else total
    m_max = max(len(instance_range(r, world, total)) for r in range(world))
    if len(instance_range(rank, world, total)) != m_local:
        raise ValueError("local shard size does not match instance_range(rank, world, total)")
    send = local
    if m_local != m_max:
        send = torch.zeros((m_max, w), dtype=local.dtype, device=local.device)
        send[:m_local] = local
    out = torch.empty((world * m_max, w), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    if total == world * m_max:
        return out
    parts = [out[r * m_max: r * m_max + len(instance_range(r, world, total))] for r in range(world)]
    return torch.cat(parts, dim=0)


def max_over_ranks(ms: float, device, group=None) -> float:
    """device-timed milliseconds -> max over ranks (the number every multi-GPU figure reports)"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
