"""vector_tools.rotate -- torch mirror of tsadar.utils.vector_tools.rotate (vector_tools.py:94-138): rotation of a square
table about its centre on the pixel grid with bilinear weights.  Used by the multiplexed-shot loss
(loss_function.py:287-317), where the second shot sees the electron distribution rotated by data.shot_rot.

Producer-side glue on a [V, V] table (16 k values): plain torch gathers, differentiable with respect to A.
Quirks kept from the reference: indices are TRUNCATED toward zero (`jnp.asarray(., dtype=int)`) while the weights use the
floor-based fractional part (`% 1`); indices are clipped to [0, n] and the gather clamps n to n - 1."""
from __future__ import annotations

import math

import torch


def rotate(A: torch.Tensor, theta: float) -> torch.Tensor:
    n0, n1 = A.shape
    dev, dt = A.device, A.dtype
    rp0, rp1 = n0 / 2, n1 / 2
    x = torch.arange(n0, dtype=dt, device=dev).reshape(1, -1).expand(n1, n0)      # meshgrid(arange(n0), arange(n1)), 'xy'
    y = torch.arange(n1, dtype=dt, device=dev).reshape(-1, 1).expand(n1, n0)
    c, s = math.cos(theta), math.sin(theta)
    or_x = c * (x - rp0) + s * (y - rp1)                                           # R(-theta) applied to (x, y) - centre
    or_y = -s * (x - rp0) + c * (y - rp1)
    fx, fy = torch.remainder(or_x, 1.0), torch.remainder(or_y, 1.0)
    w11, w12, w21, w22 = (1 - fx) * (1 - fy), (1 - fx) * fy, fx * (1 - fy), fx * fy

    def q(dy, dx):
        r = torch.clamp((or_y + rp1 + dy).to(torch.long), 0, n1).clamp(max=A.shape[0] - 1)
        cidx = torch.clamp((or_x + rp0 + dx).to(torch.long), 0, n0).clamp(max=A.shape[1] - 1)
        return A[r, cidx]

    return w11 * q(0, 0) + w12 * q(1, 0) + w21 * q(0, 1) + w22 * q(1, 1)
