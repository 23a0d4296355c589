"""ratintn -- host mirror of tsadar.core.physics.ratintn.ratintn (ratintn.py:4-23), boundary B1 of SURVEY.md 8(b).

Same name and arguments: `ratintn(f, g, z)` integrates f / g dz with f and g piecewise linear on the nodes z.  The CUDA
kernels (tsff_pv_fwd / tsff_pv_bwd) cover the form every call site of the reference uses (form_factor.py:266-268 under
`vmap(ratintn, (None, 0, None))`, and :385-386 per pole): z a uniform real grid and g = z - pole.  Anything else is refused
with ValueError -- there is no CPU fallback.

    f [N] or [1, N]   float64 CUDA tensor (differentiable)
    g [N]             -> result [1]        (one pole, what the reference returns for 1-D g)
    g [P, N]          -> result [P, 1]     (the reference's vmap over poles: `chiERratprim[:, 0]` at form_factor.py:270)
    z [N]             uniform grid (numpy array or tensor; static in the reference)

Gradients flow to f and g (through the pole: d/dpole = -sum_i d/dg_i, since g_i = z_i - pole) via tsff_pv_bwd."""
from __future__ import annotations

import numpy as np
import torch

from .engine import pv_integral, pv_integral_vjp


class _PV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, pole, z0, h, precision):
        out, _ = pv_integral(f, z0, h, pole, precision=precision, want_grad=False)
        ctx.save_for_backward(f, pole)
        ctx.grid = (z0, h)
        return out

    @staticmethod
    def backward(ctx, out_bar):
        f, pole = ctx.saved_tensors
        f_bar, pole_bar = pv_integral_vjp(f, ctx.grid[0], ctx.grid[1], pole, out_bar.contiguous())
        return f_bar, pole_bar, None, None, None


def ratintn(f, g, z, precision="fp64", check=True):
    """precision: "fp64" (validation path of the kernel, ~1e-14 of the reference) or "fp32" (the fast PV inner loop,
    ~1e-7).  check=True verifies on the host that z is uniform and that g - z is constant along the nodes (one D2H
    copy); pass False inside captured graphs."""
    zh = z.detach().cpu().numpy() if isinstance(z, torch.Tensor) else np.asarray(z)
    zh = np.asarray(zh, dtype=np.float64).reshape(-1)
    N = zh.size
    if N < 4:
        raise ValueError("ratintn: at least 4 nodes")
    z0, h = float(zh[0]), float((zh[-1] - zh[0]) / (N - 1))
    if check and np.max(np.abs(np.diff(zh) - h)) > 1e-9 * abs(h):
        raise ValueError("ratintn: the CUDA path needs a uniform grid z (every call site of the reference has one)")
    dev = torch.device("cuda", torch.cuda.current_device())
    f = f if isinstance(f, torch.Tensor) else torch.as_tensor(np.asarray(f), dtype=torch.float64)
    g = g if isinstance(g, torch.Tensor) else torch.as_tensor(np.asarray(g), dtype=torch.float64)
    f = f.to(device=dev, dtype=torch.float64)
    g = g.to(device=dev, dtype=torch.float64)
    if f.dim() == 2 and f.shape[0] != 1:
        raise ValueError("ratintn: f is one table [N] or [1, N]")
    f = f.reshape(1, N)
    single = g.dim() == 1
    g2 = g.reshape(-1, N)
    zt = torch.as_tensor(zh, device=dev)
    if check:
        d = g2.detach() - zt[None, :]
        if float((d - d[:, :1]).abs().max()) > 1e-9 * max(1.0, float(zt.abs().max())):
            raise ValueError("ratintn: the CUDA path needs g = z - pole (constant offset along the nodes)")
    pole = (zt[0] - g2[:, 0]).reshape(1, -1)          # differentiable in g through its first column ...
    if g.requires_grad:
        # ... but the cotangent belongs to all of g: g_i = z_i - pole for every i, and the reference's result depends on g only
        # through the pole on such inputs; spread it evenly so that any re-parametrisation g(pole) receives d/dpole exactly
        pole = (zt[None, :] - g2).mean(dim=1).reshape(1, -1)
    out = _PV.apply(f.contiguous(), pole.contiguous(), z0, h, precision)      # [1, P]
    return out.reshape(1) if single else out.reshape(-1, 1)
