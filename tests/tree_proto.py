"""NumPy prototype of the block-multipole far field used by the PV sweeps (tsff_tree.cuh).  Development aid: checks
the expansion algebra, the truncation order K and FP32 evaluation against the oracle's literal ratintn.

I(xi) = sum_{i=1..M-1} p_i W(g_i) + p_0 E_0(g_0) + p_M E_M(g_M),  g_i = z_i - xi,  x = h/g:
    W(g)   = sum_j x^(2j+1) / ((2j+1)(j+1))
    E_0(g) = sum_{k>=1} (-1)^(k+1) x^k / (k(k+1)),     E_M(g) = sum_{k>=1} x^k / (k(k+1))
Block of S nodes with centre c (index units), offsets e_i = i - c, y = h/g_c, t = s*y (s = S/2):
    sum_{i in blk} p_i W(g_i) = sum_m At_m t^(m+1),
    At_m = (1/s) sum_j C(m,2j) / ((2j+1)(j+1)) s^(-2j) mu_(m-2j),     mu_k = sum_i p_i (-e_i/s)^k
"""
import sys, os
import numpy as np
from math import comb

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import np_oracle as O


def moments(p, M, S, K):
    """p: node values [M+1]; returns At [NB][K], with end-node terms folded in (float64)."""
    NB = (M + 1 + S - 1) // S
    s = S / 2.0
    At = np.zeros((NB, K))
    for b in range(NB):
        c = S * b + (S - 1) / 2.0
        idx = np.arange(S * b, min(S * (b + 1), M + 1))
        w = np.where((idx >= 1) & (idx <= M - 1), p[idx], 0.0)
        e = (idx - c) / s
        mu = np.array([np.sum(w * (-e) ** k) for k in range(K)])
        for m in range(K):
            a = 0.0
            for j in range(m // 2 + 1):
                a += comb(m, 2 * j) / ((2 * j + 1) * (j + 1)) * s ** (-2 * j) * mu[m - 2 * j]
            At[b, m] = a / s
        # end nodes: coefficient of y^(m+1) is  sum_{k=1..m+1} sgn_k /(k(k+1)) C(m, m+1-k) (-e0)^(m+1-k)
        for node, sign in ((0, -1.0), (M, 1.0)):
            if S * b <= node < S * (b + 1):
                e0 = node - c
                for m in range(K):
                    a = 0.0
                    for k in range(1, m + 2):
                        sg = 1.0 if sign > 0 else (-1.0) ** (k + 1)
                        a += sg / (k * (k + 1)) * comb(m, m + 1 - k) * (-e0) ** (m + 1 - k)
                    At[b, m] += p[node] * a / s ** (m + 1)
    return At


def eval_tree(p, z0, h, M, xi, S=64, K=16, dtype=np.float64):
    At = moments(p, M, S, K).astype(dtype)
    NB = At.shape[0]
    s = S / 2.0
    n = np.clip(np.rint((xi - z0) / h), 0, M).astype(int)
    delta = xi - (z0 + n * h)
    bn = n // S
    out = np.zeros(xi.shape)
    dout = np.zeros(xi.shape)
    for ip in range(xi.size):
        accI = dtype(0)
        accJ = dtype(0)
        for b in range(NB):
            if abs(b - bn[ip]) < 2:
                continue
            c = S * b + (S - 1) / 2.0
            gc = dtype(dtype(c - n[ip]) * dtype(h) - dtype(delta[ip]))
            t = dtype(s * h) / gc
            hA = dtype(0)
            hB = dtype(0)
            for m in range(K - 1, -1, -1):
                hA = hA * t + At[b, m]
                hB = hB * t + dtype(m + 1) * At[b, m]
            accI += t * hA
            accJ += t * t * hB / dtype(s * h)
        # near window, exact in float64 (reference form)
        lo, hi = max(0, S * (bn[ip] - 1)), min(M, S * (bn[ip] + 2) - 1)
        g = z0 + h * np.arange(M + 1) - xi[ip]
        phi = lambda x: x * np.log(np.abs(x))
        I = 0.0
        J = 0.0
        for i in range(lo, hi + 1):
            if 1 <= i <= M - 1:
                I += p[i] * (phi(g[i] + h) - 2 * phi(g[i]) + phi(g[i] - h)) / h
                J += -p[i] * (np.log(abs(g[i] + h)) - 2 * np.log(abs(g[i])) + np.log(abs(g[i] - h))) / h
            elif i == 0:
                I += p[0] * ((phi(g[0] + h) - phi(g[0])) / h - 1 - np.log(abs(g[0])))
                J += p[0] * (-(np.log(abs(g[0] + h)) - np.log(abs(g[0]))) / h + 1 / g[0])
            elif i == M:
                I += p[M] * ((phi(g[M] - h) - phi(g[M])) / h + 1 + np.log(abs(g[M])))
                J += p[M] * (-(np.log(abs(g[M] - h)) - np.log(abs(g[M]))) / h - 1 / g[M])
        out[ip] = I + float(accI)
        dout[ip] = J + float(accJ)
    return out, dout


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N, P in ((4096, 200), (1024, 200), (130, 50)):
        h = 12.0 / N
        z0 = -6 + h / 2
        z = z0 + h * np.arange(N)
        f = -z * np.exp(-0.5 * z**2) * 0.4 + 0.01 * np.sin(3 * z)
        xi = rng.uniform(-7.5, 7.5, P)
        ref = O.ratintn(f[None, :], z[None, :] - xi[:, None], z)
        e = 1e-6
        fd = (O.ratintn(f[None, :], z[None, :] - (xi + e)[:, None], z) - O.ratintn(f[None, :], z[None, :] - (xi - e)[:, None], z)) / (2 * e)
        M = N - 2
        for K in (8, 12, 16):
            for dt in (np.float64, np.float32):
                out, dout = eval_tree(f[: M + 1], z0, h, M, xi, 64, K, dt)
                print(f"N={N} K={K} {dt.__name__}: I err {np.abs(out - ref.ravel()).max() / np.abs(ref).max():.2e}  "
                      f"dI err {np.abs(dout - fd.ravel()).max() / np.abs(fd).max():.2e}")
