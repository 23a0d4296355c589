"""Fit drivers over the CUDA path -- the two optimiser loops of the reference, kept thin:

* scipy_fit  -- tsadar/inverse/loops.py:43-60 + loss_function.py:149-168 (`vg_loss` for SciPy: flat float64 vector of
  the active normalised leaves in, (loss, flat gradient) out, L-BFGS-B by default) and the identical pattern of
  tests/test_inverse/test_1d_random.py:133-145;
* adam_fit   -- loops.py:75-106 / 225-250 (optax.adam on the same leaves; optax defaults b1=0.9, b2=0.999, eps=1e-8,
  eps_root=0); every lineout of the batch advances in the same launch, nothing leaves the device between steps.

Both take a closure `loss_closure(ts_params) -> torch scalar` whose graph runs through the hand-written adjoint kernels
(torch.autograd.Function wrappers in form_factor.py / irf.py / loss_function.py).  No CPU fallback: the closures only
work on CUDA tensors."""
from __future__ import annotations

import numpy as np
import torch


def batch_to_device(batch, device=None):
    """The data batch (`e_data`, `i_data`, `e_amps`, `i_amps`, `noise_e`, `noise_i`; loops.py:135-142) as float64 CUDA
    tensors.  Needed for the graphed drivers: a CUDA graph cannot capture copies from pageable host arrays."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    return {k: (v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v), dtype=torch.float64)).to(dev) for k, v in batch.items()}


def ravel_leaves(leaves):
    """Flat float64 host vector of the active leaves (jax.flatten_util.ravel_pytree in the reference)."""
    return np.concatenate([t.detach().reshape(-1).cpu().numpy() for t in leaves]) if leaves else np.zeros(0)


def unravel_into(leaves, flat):
    o = 0
    with torch.no_grad():
        for t in leaves:
            n = t.numel()
            t.copy_(torch.as_tensor(flat[o:o + n], dtype=t.dtype).reshape(t.shape))
            o += n


def value_and_grad(loss_closure, ts_params):
    """loss (python float) and flat gradient with respect to the active normalised leaves."""
    leaves = ts_params.parameters()
    for t in leaves:
        t.grad = None
    loss = loss_closure(ts_params)
    loss.backward()
    g = np.concatenate([(t.grad if t.grad is not None else torch.zeros_like(t)).reshape(-1).cpu().numpy() for t in leaves])
    return float(loss.detach()), g


class GraphedValueAndGrad:
    """loss and gradient as ONE captured CUDA graph: a call copies the flat parameter vector into the leaves (one H2D),
    replays the graph (transforms -> form factor -> IRF -> loss -> hand-written adjoints) and reads loss + flat gradient
    back (one D2H).  For the SciPy drivers, whose every function evaluation is otherwise a few dozen small launches.
    The closure must work on device-resident data only (batch_to_device): host arrays cannot be copied inside a capture."""

    def __init__(self, loss_closure, ts_params):
        self.leaves = ts_params.parameters()
        dev = self.leaves[0].device
        n = sum(t.numel() for t in self.leaves)
        self.x_dev = torch.zeros(n, dtype=torch.float64, device=dev)
        self.out_dev = torch.zeros(n + 1, dtype=torch.float64, device=dev)          # [loss, flat gradient]
        self.x_pin = torch.zeros(n, dtype=torch.float64).pin_memory()
        self.out_pin = torch.zeros(n + 1, dtype=torch.float64).pin_memory()

        def body():
            o = 0
            with torch.no_grad():
                for t in self.leaves:
                    t.copy_(self.x_dev[o:o + t.numel()].reshape(t.shape))
                    o += t.numel()
            loss = loss_closure(ts_params)
            g = torch.autograd.grad(loss, self.leaves, allow_unused=True)
            with torch.no_grad():
                self.out_dev[0] = loss.detach()
                o = 1
                for t, gt in zip(self.leaves, g):
                    self.out_dev[o:o + t.numel()] = 0.0 if gt is None else gt.reshape(-1)
                    o += t.numel()

        self.x_dev.copy_(torch.as_tensor(ravel_leaves(self.leaves)))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            body()

    def __call__(self, x):
        self.x_pin.copy_(torch.as_tensor(np.asarray(x, dtype=np.float64)))
        self.x_dev.copy_(self.x_pin, non_blocking=True)
        self.graph.replay()
        self.out_pin.copy_(self.out_dev, non_blocking=True)
        torch.cuda.current_stream(self.x_dev.device).synchronize()
        out = self.out_pin.numpy()
        return float(out[0]), out[1:].copy()


def scipy_fit(loss_closure, ts_params, method="L-BFGS-B", options=None, bounds=None, cuda_graph=False):
    """scipy.optimize.minimize(jac=True) over the active leaves of `ts_params` (updated in place).  Returns the
    OptimizeResult.  NaN losses propagate to SciPy unchanged, as in the reference.  cuda_graph=True evaluates loss and
    gradient through GraphedValueAndGrad."""
    from scipy.optimize import minimize
    leaves = ts_params.parameters()
    if cuda_graph:
        x0 = ravel_leaves(leaves)
        fun = GraphedValueAndGrad(loss_closure, ts_params)
        res = minimize(fun, x0, method=method, jac=True, bounds=bounds, options=options or {})
        unravel_into(leaves, res["x"])
        return res

    def fun(x):
        unravel_into(leaves, x)
        return value_and_grad(loss_closure, ts_params)

    res = minimize(fun, ravel_leaves(leaves), method=method, jac=True, bounds=bounds, options=options or {})
    unravel_into(leaves, res["x"])
    return res


def adam_fit(loss_closure, ts_params, learning_rate, num_steps, b1=0.9, b2=0.999, eps=1e-8, callback=None, cuda_graph=False):
    """optax.adam(learning_rate) on the active leaves; the moment updates are one fused foreach call per step and the
    loss history stays on the device until the end (one D2H copy).

    cuda_graph=True captures ONE whole step (ThomsonParams transforms -> form factor -> IRF -> loss -> hand-written
    adjoints -> adam update) as a CUDA graph and replays it: the reference's fits run with 2-6 lineouts per batch
    (loops.py:132-152), where a step is a few dozen small launches and the launch/host overhead, not the kernels, sets the
    pace.  The bias corrections enter through a device-side step counter, so every replay is the
    same graph."""
    if cuda_graph:
        return _adam_fit_graphed(loss_closure, ts_params, learning_rate, num_steps, b1, b2, eps)
    leaves = ts_params.parameters()
    mu = [torch.zeros_like(t) for t in leaves]
    nu = [torch.zeros_like(t) for t in leaves]
    hist = torch.zeros(num_steps, dtype=torch.float64, device=leaves[0].device)
    for k in range(num_steps):
        for t in leaves:
            t.grad = None
        loss = loss_closure(ts_params)
        loss.backward()
        hist[k] = loss.detach()
        g = [t.grad for t in leaves]
        with torch.no_grad():
            torch._foreach_mul_(mu, b1)
            torch._foreach_add_(mu, g, alpha=1 - b1)
            torch._foreach_mul_(nu, b2)
            torch._foreach_addcmul_(nu, g, g, value=1 - b2)
            c1, c2 = 1 - b1 ** (k + 1), 1 - b2 ** (k + 1)
            den = torch._foreach_div(nu, c2)
            torch._foreach_sqrt_(den)
            torch._foreach_add_(den, eps)
            torch._foreach_addcdiv_(leaves, mu, den, value=-learning_rate / c1)
        if callback is not None:
            callback(k, hist[k])
    return hist.cpu().numpy()


def _adam_fit_graphed(loss_closure, ts_params, learning_rate, num_steps, b1, b2, eps):
    leaves = ts_params.parameters()
    dev = leaves[0].device
    mu = [torch.zeros_like(t) for t in leaves]
    nu = [torch.zeros_like(t) for t in leaves]
    n_warm = 3
    hist = torch.zeros(max(num_steps, n_warm), dtype=torch.float64, device=dev)   # the warm-up steps write slots 0..2 too
    step = torch.zeros((), dtype=torch.float64, device=dev)          # device-side iteration counter
    slot = torch.zeros((), dtype=torch.long, device=dev)

    def one_step():
        for t in leaves:
            t.grad = None
        loss = loss_closure(ts_params)
        g = torch.autograd.grad(loss, leaves)
        with torch.no_grad():
            step.add_(1.0)
            hist.index_put_((slot,), loss.detach())
            slot.add_(1)
            torch._foreach_mul_(mu, b1)
            torch._foreach_add_(mu, g, alpha=1 - b1)
            torch._foreach_mul_(nu, b2)
            torch._foreach_addcmul_(nu, g, g, value=1 - b2)
            c1 = 1.0 - torch.pow(torch.full_like(step, b1), step)
            c2 = 1.0 - torch.pow(torch.full_like(step, b2), step)
            for t, m, v in zip(leaves, mu, nu):
                t.sub_(learning_rate * (m / c1) / (torch.sqrt(v / c2) + eps))

    keep = [t.detach().clone() for t in leaves]

    def reset():
        with torch.no_grad():
            for t, k0 in zip(leaves, keep):
                t.copy_(k0)
            for m in mu + nu:
                m.zero_()
            step.zero_(); slot.zero_(); hist.zero_()

    # warm-up on a side stream (lazy context creation, scratch buffers, shared-memory opt-ins).  The warm-up advances the live
    # leaves and the moments: whatever happens in it (or in the capture), the caller's parameters are restored.
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    try:
        with torch.cuda.stream(side):
            for _ in range(n_warm):
                one_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
    finally:
        reset()
    graph = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(graph):
            one_step()
    finally:
        reset()                               # capture does not execute; make the state pristine anyway
    for _ in range(num_steps):
        graph.replay()
    return hist[:num_steps].cpu().numpy()


def fused_adam_fit(loss_closure, ts_params, learning_rate, num_steps, b1=0.9, b2=0.999, eps=1e-8, cuda_graph=True):
    """adam_fit for a FusedThomsonParams: per step ONE launch turns the normalised leaves into the
    parameter block and the f tables (tsff_params_fwd), the form factor / IRF / loss kernels and their adjoints run, ONE
    launch reverses the transforms (tsff_params_bwd) and ONE launch applies optax.adam to every active leaf of every lineout
    (tsff_adam_step, step counters on the device) -- captured as one CUDA graph.  -> loss history [num_steps] (numpy)."""
    from . import _ffi
    x = ts_params.x
    B, n = x.shape
    dev = x.device
    mu, nu = torch.zeros_like(x), torch.zeros_like(x)
    count = torch.zeros(B, dtype=torch.float64, device=dev)
    n_warm = 3
    hist = torch.zeros(max(num_steps, n_warm), dtype=torch.float64, device=dev)
    slot = torch.zeros((), dtype=torch.long, device=dev)

    def one_step():
        loss = loss_closure(ts_params)
        (g,) = torch.autograd.grad(loss, [x])
        with torch.no_grad():
            hist.index_put_((slot,), loss.detach())
            slot.add_(1)
            st = torch.cuda.current_stream(dev).cuda_stream
            _ffi.check(_ffi.lib().tsff_adam_step(B, n, x.data_ptr(), g.contiguous().data_ptr(), mu.data_ptr(), nu.data_ptr(), count.data_ptr(),
                                                 float(learning_rate), b1, b2, eps, st))

    if not cuda_graph:
        for _ in range(num_steps):
            one_step()
        return hist[:num_steps].cpu().numpy()
    keep = x.detach().clone()

    def reset():
        with torch.no_grad():
            x.copy_(keep)
            mu.zero_(); nu.zero_(); count.zero_(); slot.zero_(); hist.zero_()

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    try:
        with torch.cuda.stream(side):
            for _ in range(n_warm):
                one_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
    finally:
        reset()
    graph = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(graph):
            one_step()
    finally:
        reset()
    for _ in range(num_steps):
        graph.replay()
    return hist[:num_steps].cpu().numpy()
