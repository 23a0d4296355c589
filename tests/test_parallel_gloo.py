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


# ---- ARTS wavelength-axis sharding: the copy_to_shards / gather_columns operator pair ------------------------------
def _stand_in_ff(theta, fe, lam):
    """A differentiable stand-in for the form factor on a slice of the wavelength axis: [len(lam), A]."""
    ang = torch.linspace(0.3, 2.5, 5, dtype=torch.float64)
    return torch.exp(-((lam[:, None] - 500.0 - 30 * theta[0]) / (20.0 + 5 * theta[1])) ** 2) * torch.cos(ang)[None, :] ** 2 \
        + (fe[: lam.numel(), None] * torch.sin(ang)[None, :] * theta[1])


def _chain(theta, fe, lam_all, wm, sl=None, group=None, npts=None):
    """weights @ ff^T (sharded or not) followed by a replicated 'instrument stage' that uses theta again."""
    from tsadar_b200.parallel import copy_to_shards, gather_columns
    if sl is None:
        modl = wm @ _stand_in_ff(theta, fe, lam_all).t()
    else:
        th, f = copy_to_shards(theta, group), copy_to_shards(fe, group)
        ff = _stand_in_ff(th, f[sl[0]:], lam_all[sl[0]:sl[1]])
        modl = gather_columns((wm @ ff.t()).contiguous(), npts, group)
    return (modl * torch.linspace(1, 2, modl.shape[1], dtype=torch.float64)).sum() * theta[0] + (modl ** 2).sum() * theta[1]


def _worker_wshard(rank, world, port, npts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tsadar_b200.parallel import WShard
    rng = np.random.default_rng(5)
    lam_all = torch.linspace(400, 700, npts, dtype=torch.float64)
    wm = torch.tensor(rng.uniform(size=(7, 5)))
    t0, f0 = rng.uniform(0.5, 1.5, size=2), rng.uniform(size=npts)
    theta = torch.tensor(t0, requires_grad=True)
    fe = torch.tensor(f0, requires_grad=True)
    ref = _chain(theta, fe, lam_all, wm)
    ref.backward()
    gt_ref, gf_ref = theta.grad.clone(), fe.grad.clone()
    theta2 = torch.tensor(t0, requires_grad=True)
    fe2 = torch.tensor(f0, requires_grad=True)
    sh = WShard(npts, halo=False)
    out = _chain(theta2, fe2, lam_all, wm, sl=(sh.j0, sh.j1), npts=npts)
    out.backward()
    ok = bool(torch.allclose(out, ref, rtol=1e-13)) and bool(torch.allclose(theta2.grad, gt_ref, rtol=1e-12)) \
        and bool(torch.allclose(fe2.grad, gf_ref, rtol=1e-12, atol=1e-15))
    shh = WShard(npts, halo=True)
    ok = ok and shh.keep == sh.j1 - sh.j0 and shh.j1e == min(sh.j1 + 1, npts)
    q.put((rank, sh.j0, sh.j1, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,npts", [(2, 16), (3, 17), (2, 3)])
def test_wavelength_sharding_operators(world, npts):
    """Sharded evaluation + all-gather == single-process result, and the gradients of parameters used BOTH inside the
    sharded region and in the replicated stage behind it are identical on every rank and equal to the unsharded ones."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_wshard, args=(r, world, port, npts, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    covered = []
    for rank, s, e, ok in res:
        assert ok, f"rank {rank} mismatch"
        covered += list(range(s, e))
    assert covered == list(range(npts))


def test_numa_binding_is_a_no_op_without_gpu_topology():
    """bind_to_gpu_numa_node must never break a rank: without a CUDA device / NVML it leaves the affinity alone."""
    import os
    import torch
    from tsadar_b200.parallel import bind_to_gpu_numa_node
    if torch.cuda.is_available():
        pytest.skip("CPU-side behaviour")
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before


# ---- multiplexed shots: one shot per half of the ranks -----------------------------------------------------
def test_shot_split():
    from tsadar_b200.parallel import shot_split
    assert shot_split(0, 1) == (None, [0])
    assert [shot_split(r, 2)[0] for r in range(2)] == [0, 1]
    assert [shot_split(r, 8) for r in (0, 3, 4, 7)] == [(0, [0, 1, 2, 3]), (0, [0, 1, 2, 3]), (1, [4, 5, 6, 7]), (1, [4, 5, 6, 7])]
    with pytest.raises(ValueError):
        shot_split(0, 3)


def _worker_shots(rank, world, port, q):
    """The recipe of LossFunction's shot sharding on a toy model: loss = sum over two shots of g_k(theta); half of the ranks
    evaluate each shot, the total and the gradient of the shared leaf must equal the single-process ones on every rank."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tsadar_b200.parallel import shot_assignment, allreduce_sum_identity_grad
    shot, half, grp = shot_assignment()
    theta = torch.tensor([0.3, -1.1], dtype=torch.float64, requires_grad=True)
    g = [lambda t: (t ** 2).sum() * 1.5, lambda t: torch.sin(t).sum() + t[0] * t[1]]
    local = g[shot](theta)
    total = allreduce_sum_identity_grad(local) / half
    total.backward()
    dist.all_reduce(theta.grad)
    t0 = torch.tensor([0.3, -1.1], dtype=torch.float64, requires_grad=True)
    ref = g[0](t0) + g[1](t0)
    ref.backward()
    ok = bool(torch.allclose(total, ref, rtol=1e-14)) and bool(torch.allclose(theta.grad, t0.grad, rtol=1e-14))
    ok = ok and ((grp is None) == (half == 1)) and shot == (0 if rank < world // 2 else 1)
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_shot_sharding_recipe(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_shots, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_pixel_rotation_matches_the_oracle():
    """tsadar_b200.vector_tools.rotate (torch) vs oracle.rotate_pixels (NumPy restatement of vector_tools.py:94-138), several
    angles, and the properties that follow from the source: rotation by 0 is the identity, by pi/2 a quarter turn."""
    import numpy as np
    from oracle import np_oracle as O
    from tsadar_b200.vector_tools import rotate
    rng = np.random.default_rng(0)
    A = rng.normal(size=(24, 24))
    for th in (0.0, 0.3, -1.2, np.pi / 2, 2.5, 33.0 * np.pi / 180):
        a, b = O.rotate_pixels(A, th), rotate(torch.tensor(A), th).numpy()
        assert np.abs(a - b).max() <= 1e-13
    assert np.array_equal(O.rotate_pixels(A, 0.0), A)
    At = torch.tensor(A, requires_grad=True)
    rotate(At, 0.4).sum().backward()
    assert torch.isfinite(At.grad).all() and float(At.grad.abs().sum()) > 0


# ---- loops that contain a collective run equally often on every rank ----------------------------------------------------
def _worker_steps(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tsadar_b200.parallel import agreed_step_count
    # step times either side of the 297 / 298 boundary of ceil(2000 / ms) (2000 / 297 = 6.734): each rank's own clock would say
    # 298, 297, 297, 298 steps; the agreed count is the slowest rank's
    ms = [6.730, 6.740, 6.7345, 6.7339][rank]
    own = int(np.ceil(2000.0 / ms))
    n = agreed_step_count(ms, 2.0, 20)
    # the loop the count is for: one all-reduce per step, then a barrier -- with per-rank counts this never returns
    acc = torch.zeros(1, dtype=torch.float64)
    for _ in range(n):
        t = torch.ones(1, dtype=torch.float64)
        dist.all_reduce(t)
        acc += t
    dist.barrier()
    q.put((rank, own, n, float(acc)))
    dist.destroy_process_group()


def test_step_count_is_agreed_across_ranks():
    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_steps, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert len({own for _, own, _, _ in res}) == 2          # the ranks' own clocks disagree ...
    assert {n for _, _, n, _ in res} == {297}               # ... the agreed count does not (slowest rank: 6.740 ms -> 297)
    assert all(acc == 297.0 * world for _, _, _, acc in res)
