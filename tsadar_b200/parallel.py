"""Multi-GPU sharding of the hot path.  Lineouts are independent (every parameter is per-lineout,
ts_params.py:93-99), so ranks own contiguous blocks of lineouts with all static tables replicated and NO data-path
collective; the only exchange is the all-reduce of the scalar loss (the reference's loss is a nanmean over the whole
batch, loss_function.py:371, so per-rank partial sums carry the global 1/B_total already)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous block of lineouts owned by `rank` (sizes differ by at most one; empty blocks allowed)."""
    base, rem = divmod(int(n_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_loss(loss: torch.Tensor) -> torch.Tensor:
    """Sum the per-rank partial losses (each already scaled by 1/B_total).  NCCL on GPUs, gloo in the CPU tests."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(loss, op=dist.ReduceOp.SUM)
    return loss


def gather_rows(x: torch.Tensor, n_total: int) -> torch.Tensor:
    """Debug/verification helper: all-gather variable-size row blocks back into the global [n_total, ...] array."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    world = dist.get_world_size()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    pad = max(e - s for s, e in sizes)
    buf = torch.zeros((pad,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    buf[: x.shape[0]] = x
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[: e - s] for o, (s, e) in zip(out, sizes)], dim=0)
