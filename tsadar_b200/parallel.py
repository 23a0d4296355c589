"""Multi-GPU sharding of the hot path.  Lineouts are independent (every parameter is per-lineout,
ts_params.py:93-99), so ranks own contiguous blocks of lineouts with all static tables replicated and NO data-path
collective; the only exchange is the all-reduce of the scalar loss (the reference's loss is a nanmean over the whole
batch, loss_function.py:371, so per-rank partial sums carry the global 1/B_total already).

ARTS ("angular_full": ONE parameter set for one image) shards the wavelength axis instead, mirroring
the reference's own device split of the pole grid (form_factor.py:431-447): every rank evaluates the form factor on W/N
wavelengths x all angles and applies the angular weight matrix locally; the [1024, W/N] slabs are all-gathered in front
of the instrument stage, which is replicated.  The pair of graph operators below keeps autograd exact:
    copy_to_shards   forward identity,   backward all-reduce(sum)   -- on the operands entering the sharded region
    gather_columns   forward all-gather, backward "take my slab"    -- on its output
so the partial cotangents of the shared parameters / f table are summed across ranks exactly once, while everything
computed redundantly downstream (amp1/amp2/lam in the instrument stage, the loss) is left alone."""
from __future__ import annotations

import numpy as np
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


def agreed_step_count(ms_per_step: float, seconds: float, at_least: int, device=None) -> int:
    """Steps that fill `seconds` at `ms_per_step`, THE SAME on every rank: the slowest rank's step time (all-reduce MAX) decides.
    A loop whose body holds a collective (the per-step loss all-reduce) must run equally often everywhere; a count rounded from each
    rank's own clock can differ by one between ranks and leaves them waiting for each other in mismatched collectives."""
    t = torch.tensor([float(ms_per_step)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return max(int(at_least), int(np.ceil(seconds * 1e3 / float(t.item()))))


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


# ---- ARTS: wavelength-axis sharding ------------------------------------------------------------------------------
def _active(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


class _CopyToShards(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


class _GatherColumns(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sizes, rank, group):
        ctx.sizes, ctx.rank = sizes, rank
        pad = max(sizes)
        buf = x.new_zeros(tuple(x.shape[:-1]) + (pad,))
        buf[..., : x.shape[-1]] = x
        out = [torch.empty_like(buf) for _ in sizes]
        dist.all_gather(out, buf.contiguous(), group=group)
        return torch.cat([o[..., :n] for o, n in zip(out, sizes)], dim=-1)

    @staticmethod
    def backward(ctx, g):
        o = sum(ctx.sizes[: ctx.rank])
        return g[..., o:o + ctx.sizes[ctx.rank]].contiguous(), None, None, None


def copy_to_shards(x: torch.Tensor, group=None) -> torch.Tensor:
    return _CopyToShards.apply(x, group) if _active(group) and x.requires_grad else x


def gather_columns(x_local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather the last axis: rank r holds columns shard_range(n_total, r, world)."""
    if not _active(group):
        return x_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [e - s for s, e in (shard_range(n_total, r, world) for r in range(world))]
    assert x_local.shape[-1] == sizes[rank], (x_local.shape, sizes, rank)
    return _GatherColumns.apply(x_local, sizes, rank, group)


class WShard:
    """This rank's slice of a wavelength axis of `npts` points.  [j0, j1) are the points it owns; in table mode the
    forward difference along omega (form_factor.py:258-261) needs the next point, so the evaluated slice [j0, j1e) carries
    one halo point whose output is dropped (`keep` = j1 - j0 outputs are kept)."""

    def __init__(self, npts, rank=None, world=None, group=None, halo=True):
        self.group = group
        self.world = dist.get_world_size(group) if world is None else int(world)
        self.rank = dist.get_rank(group) if rank is None else int(rank)
        self.npts = int(npts)
        self.j0, self.j1 = shard_range(self.npts, self.rank, self.world)
        if self.j1 <= self.j0:
            raise ValueError(f"rank {self.rank} of {self.world} owns no wavelength of {npts}: use fewer ranks")
        self.j1e = min(self.j1 + (1 if halo else 0), self.npts)
        self.keep = self.j1 - self.j0


# ---- multiplexed shots: one shot per half of the ranks ---------------------------------------------------
def shot_split(rank: int, world: int):
    """(shot index, ranks of my half): ranks [0, world/2) evaluate shot 0, the rest shot 1; world must be even (or 1)."""
    if world < 2:
        return None, [0]
    if world % 2:
        raise ValueError(f"multiplexed shots need an even number of ranks, got {world}")
    half = world // 2
    shot = 0 if rank < half else 1
    return shot, list(range(shot * half, (shot + 1) * half))


_shot_groups = {}


def shot_assignment():
    """-> (shot or None, ranks per half, process group of my half or None).  Every rank must call it (new_group is
    collective); the groups are cached per world size."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
        return None, 1, None
    world, rank = dist.get_world_size(), dist.get_rank()
    shot, mine = shot_split(rank, world)
    half = world // 2
    if half == 1:
        return shot, 1, None
    if world not in _shot_groups:
        _shot_groups[world] = [dist.new_group(list(range(k * half, (k + 1) * half))) for k in range(2)]
    return shot, half, _shot_groups[world][shot]


class _AllReduceSumIdentityGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = x.detach().clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM)
        return y

    @staticmethod
    def backward(ctx, g):
        return g


def allreduce_sum_identity_grad(x: torch.Tensor) -> torch.Tensor:
    """sum over ranks of per-rank partial losses; d total / d my partial = 1 (the other ranks' partials do not depend on my
    graph)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    return _AllReduceSumIdentityGrad.apply(x)


def bind_to_gpu_numa_node(device_index):
    """Restrict the calling process to the CPUs NVML reports as local to this GPU (nvmlDeviceSetCpuAffinity), so that the
    pinned staging buffers it allocates afterwards are first-touched on the GPU's own NUMA node and the H2D copies of the
    ranks of a node do not cross the socket interconnect.  One process per GPU, call it before allocating pinned memory.
    Returns the number of CPUs in the new affinity mask, or None when NVML or the topology information is unavailable
    (the process is then left as it was)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        try:
            bus = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(props.uuid))
        before = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = os.sched_getaffinity(0)
        if not after:
            os.sched_setaffinity(0, before)
            return None
        return len(after)
    except Exception:
        return None
