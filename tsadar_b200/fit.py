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


def scipy_fit(loss_closure, ts_params, method="L-BFGS-B", options=None, bounds=None):
    """scipy.optimize.minimize(jac=True) over the active leaves of `ts_params` (updated in place).  Returns the
    OptimizeResult.  NaN losses propagate to SciPy unchanged, as in the reference."""
    from scipy.optimize import minimize
    leaves = ts_params.parameters()

    def fun(x):
        unravel_into(leaves, x)
        return value_and_grad(loss_closure, ts_params)

    res = minimize(fun, ravel_leaves(leaves), method=method, jac=True, bounds=bounds, options=options or {})
    unravel_into(leaves, res["x"])
    return res


def adam_fit(loss_closure, ts_params, learning_rate, num_steps, b1=0.9, b2=0.999, eps=1e-8, callback=None):
    """optax.adam(learning_rate) on the active leaves; the moment updates are one fused foreach call per step and the
    loss history stays on the device until the end (one D2H copy)."""
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
