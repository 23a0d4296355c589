"""N>1 host logic on CPU with the gloo backend (world_size 2 and 3): lineout sharding covers every lineout exactly
once, and the all-reduced partial losses (each rank scaling by 1/B_total over its own block) equal the single-process
nanmean-style loss.  No kernels are involved: this is the plumbing around them."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tsadar_b200.parallel import shard_range, allreduce_loss, gather_rows
    rng = np.random.default_rng(0)
    theory = torch.tensor(rng.normal(size=(n_total, 64)))
    data = torch.tensor(rng.normal(size=(n_total, 64)))
    w = torch.tensor((rng.uniform(size=64) > 0.5).astype(np.float64))
    w = w / w.sum()
    s, e = shard_range(n_total, rank, world)
    part = (w * (data[s:e] - theory[s:e]) ** 2).sum() / n_total      # what tsff_loss_fwd_bwd computes with scale=1/B_total
    total = allreduce_loss(part.reshape(1).clone())
    rows = gather_rows(theory[s:e].clone(), n_total)
    full = (w * (data - theory) ** 2).sum() / n_total
    ok = bool(torch.allclose(total, full.reshape(1), rtol=1e-13)) and bool(torch.equal(rows, theory))
    q.put((rank, s, e, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 10), (3, 7), (2, 1)])
def test_sharding_and_loss_allreduce(world, n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    covered = []
    for rank, s, e, ok in res:
        assert ok, f"rank {rank} mismatch"
        covered += list(range(s, e))
    assert covered == list(range(n_total))
