"""f(v) producers -- torch mirrors of tsadar/core/modules/distribution_functions (base.py, spherical_harmonics.py):
Arbitrary1V (base.py:154-204), Arbitrary2V (base.py:335-426), SphericalHarmonics (spherical_harmonics.py:150-318)
with its radial models FLM_MY (:59-117) and ArbitraryVr (:119-147).  They sit directly upstream of the form-factor
kernels: the output table is the `fe` operand of tsff_ff_fwd and torch autograd carries the
kernels' fe_bar back to the trainable leaves.

Everything static is folded at construction time so that one call is a handful of dense device ops:
* the forward-backward second-order Butterworth filter of Arbitrary1V (a serial `lax.scan` of V steps in the reference,
  base.py:41-96) is LINEAR in the signal, start-up convention included, so it is applied as one [V, V] matrix
  (built once on the host by pushing the identity through the recurrence);
* the Hann smoothing of ArbitraryVr (`jnp.convolve(..., "same")`) likewise;
* `jnp.interp(vr_vxvy, vr, .)` onto the Cartesian grid becomes a fixed gather + lerp (indices and weights precomputed);
* the real spherical harmonics Re Y_l^m on the (vx, vy) mesh are constants.
FLM_NN (two equinox MLPs, spherical_harmonics.py:14-49) is mirrored with the same architecture and forward arithmetic; its
INITIAL weights in the reference come from jax.random.PRNGKey(0) / PRNGKey(42) through equinox's Linear initialiser --
JAX's threefry stream is not available here, so the weights are either loaded (`params.nn_weights`: the arrays exported from
a JAX run) or drawn from the same distribution, U(-1/sqrt(fan_in), 1/sqrt(fan_in)), with torch's generator seeded alike."""
from __future__ import annotations

import math

import numpy as np
import torch
from scipy.special import gamma as _gamma, lpmv

DT = torch.float64


def _inv_act(x):
    return np.log(1e-2 + x / (1 - x + 1e-2))  # the reference's "inverse" (not the logit)


def vgrid(nvx):
    vmax = 6.0
    dv = 2 * vmax / nvx
    return np.linspace(-vmax + dv / 2, vmax - dv / 2, nvx)


def _tgamma(x):
    return torch.exp(torch.lgamma(x))          # arguments are positive here (3/m, 5/m, 6/m, 8/m with m in [2, 5])


# ---- linear operators built once on the host ----------------------------------------------------------------------
def butterworth_matrix(n, f_sampling=100, f_cutoff=6):
    """[n, n] matrix S with second_order_butterworth(x, method="forward_backward") == S @ x (base.py:41-96)."""
    ff = f_cutoff / f_sampling
    ita = 1.0 / np.tan(np.pi * ff)
    q = np.sqrt(2.0)
    b0 = 1.0 / (1.0 + q * ita + ita**2)
    b1, b2 = 2 * b0, b0
    a1 = 2.0 * (ita**2 - 1.0) * b0
    a2 = -(1.0 - q * ita + ita**2) * b0

    def one_pass(X):                      # X [n, k]: k signals as columns, filtered along axis 0
        x1, x2, y1, y2 = X[1], X[0], X[1], X[0]
        rows = []
        for i in range(2, X.shape[0]):
            y = b0 * X[i] + b1 * x1 + b2 * x2 + a1 * y1 + a2 * y2
            x1, x2, y1, y2 = X[i], x1, y, y1
            rows.append(y)
        rows = np.stack(rows)
        return np.concatenate((rows[0:1], rows[0:1], rows))

    S = one_pass(np.eye(n))
    return one_pass(S[::-1])[::-1].copy()


def hann_same_matrix(n, window_size):
    """[n, n] matrix H with smooth1d(x, window_size) == H @ x   (base.py:17-38: np.convolve(x, hanning/sum, "same"))."""
    w = np.hanning(window_size)
    w = w / w.sum()
    return np.stack([np.convolve(e, w, mode="same") for e in np.eye(n)], axis=1)


def interp_plan(xq, xp):
    """Static part of jnp.interp(xq, xp, fp): cell index i (fp[i], fp[i+1]), weight t, masks of the clamped sides."""
    xq = np.asarray(xq, dtype=np.float64).ravel()
    i = np.clip(np.searchsorted(xp, xq, side="right"), 1, len(xp) - 1) - 1
    t = (xq - xp[i]) / (xp[i + 1] - xp[i])
    return i, t, xq < xp[0], xq > xp[-1]


# ---- 1V -------------------------------------------------------------------------------------------------------------
class Arbitrary1V:
    """Learned 1-D table: f = 10^(-(7 S fval)^2), normalised (base.py:201-204); fval starts from a super-Gaussian of
    order init_m (base.py:188-198).  One row of `fval` per lineout (the reference keeps a list of modules)."""

    def __init__(self, dist_cfg, batch_size, device, trainable):
        self.vx = vgrid(dist_cfg["nvx"])
        self.dv = self.vx[1] - self.vx[0]
        m = float(dist_cfg["params"]["init_m"])
        alpha = np.sqrt(3.0 * _gamma(3.0 / m) / 2.0 / _gamma(5.0 / m))
        cst = m / (4.0 * np.pi * alpha**3.0 * _gamma(3.0 / m))
        fdlm = cst * np.exp(-(np.abs(self.vx / alpha) ** m))
        fdlm = fdlm / np.sum(fdlm) / self.dv
        f0 = np.sqrt(-np.log10(fdlm)) / 7.0
        self.fval = torch.tensor(np.tile(f0, (batch_size, 1)), dtype=DT, device=device, requires_grad=bool(trainable))
        self.S_t = torch.tensor(butterworth_matrix(self.vx.size).T.copy(), dtype=DT, device=device)

    def leaves(self):
        return {"fval": self.fval}

    def __call__(self):
        f = torch.pow(10.0, -((7.0 * (self.fval @ self.S_t)) ** 2))
        return f / f.sum(dim=1, keepdim=True) / self.dv


# ---- 2V -------------------------------------------------------------------------------------------------------------
class Arbitrary2V:
    """Learned 2-D table (base.py:335-426): f = fval^2 (or 10^(-fval^2) with learn_log), normalised on the dvx^2 mesh."""

    def __init__(self, dist_cfg, device, trainable):
        self.vx = vgrid(dist_cfg["nvx"])
        self.dv = self.vx[1] - self.vx[0]
        self.learn_log = bool(dist_cfg["params"]["learn_log"])
        m = float(dist_cfg["params"]["init_m"])
        vth_x = np.sqrt(2.0)
        alpha = np.sqrt(3.0 * _gamma(3.0 / m) / 2.0 / _gamma(5.0 / m))
        cst = m / (4.0 * np.pi * alpha**3.0 * _gamma(3.0 / m))
        f = cst / vth_x**3.0 * np.exp(-((np.sqrt(self.vx[:, None] ** 2.0 + self.vx[None, :] ** 2.0) / alpha / vth_x) ** m))
        f = f / np.sum(f) / self.dv**2.0
        if self.learn_log:
            f = -np.log10(f)
        self.fval = torch.tensor(np.sqrt(f), dtype=DT, device=device, requires_grad=bool(trainable))

    def leaves(self):
        return {"fval": self.fval}

    def __call__(self):
        f = self.fval**2.0
        if self.learn_log:
            f = torch.pow(10.0, -f)
        return f / f.sum() / self.dv**2.0


class _FlmMoraYahi:
    """FLM_MY (spherical_harmonics.py:59-117): Mora & Yahi (1982) eq. 3 radial profile, scaled by 10^-log_10_LT f00."""

    def __init__(self, vr, LT, device, trainable):
        self.vr = vr
        self.log_10_LT = torch.tensor(math.log10(LT), dtype=DT, device=device, requires_grad=bool(trainable))

    def leaves(self):
        return {"log_10_LT": self.log_10_LT}

    def __call__(self, m_f0, f00):
        ve = _tgamma(5.0 / m_f0) / 3 / _tgamma(3.0 / m_f0)          # :107 (no square root in the reference)
        uu = self.vr
        lambda_v = (self.vr / ve) ** 4.0
        coeff = (m_f0 / 2 * uu**m_f0 - 5 * m_f0 / 12 * _tgamma(8 / m_f0) / _tgamma(6 / m_f0) * uu ** (m_f0 - 2) - 1.5) * lambda_v
        return coeff / 10**self.log_10_LT * f00


class _FlmArbitraryVr:
    """ArbitraryVr (spherical_harmonics.py:119-147): 10^(-10 sigmoid(H mag)) * tanh(H sign), H = Hann smoothing."""

    def __init__(self, nvr, device, trainable):
        self.flm_sign = torch.zeros(nvr, dtype=DT, device=device, requires_grad=bool(trainable))
        self.flm_mag = torch.zeros(nvr, dtype=DT, device=device, requires_grad=bool(trainable))
        self.H_t = torch.tensor(hann_same_matrix(nvr, nvr // 4).T.copy(), dtype=DT, device=device)

    def leaves(self):
        return {"flm_mag": self.flm_mag, "flm_sign": self.flm_sign}

    def __call__(self, m_f0, f00):
        sign = torch.tanh(self.flm_sign @ self.H_t)
        mag = -torch.sigmoid(self.flm_mag @ self.H_t) * 10
        return 10**mag * sign


class _FlmNN:
    """FLM_NN (spherical_harmonics.py:14-49): flm(vr) = 10^(-MLP_mag(vr)) f00 tanh-MLP_sign(vr), both MLPs = eqx.nn.MLP(in 1,
    out 1, width 32, depth 3): Linear(1,32) relu Linear(32,32) relu Linear(32,32) relu Linear(32,1), final activation relu
    (magnitude) / tanh (sign).  All weights trainable when the distribution is."""

    def __init__(self, vr, device, trainable, weights=None, seeds=(0, 42)):
        self.vr = vr
        sizes = [(32, 1), (32, 32), (32, 32), (1, 32)]
        self.nets = {}
        for name, seed in zip(("mag", "sign"), seeds):
            layers = []
            g = torch.Generator().manual_seed(seed)
            for k, (o, i) in enumerate(sizes):
                if weights is not None:
                    W, b = weights[name][k]
                    W, b = torch.as_tensor(np.asarray(W), dtype=DT).reshape(o, i), torch.as_tensor(np.asarray(b), dtype=DT).reshape(o)
                else:        # equinox Linear: weight and bias ~ U(-1/sqrt(in_features), 1/sqrt(in_features))
                    lim = 1.0 / math.sqrt(i)
                    W = (torch.rand((o, i), dtype=DT, generator=g) * 2 - 1) * lim
                    b = (torch.rand((o,), dtype=DT, generator=g) * 2 - 1) * lim
                layers.append((W.to(device).requires_grad_(bool(trainable)), b.to(device).requires_grad_(bool(trainable))))
            self.nets[name] = layers

    def leaves(self):
        return {f"flm_{n}.layers[{k}].{w}": t for n, ls in self.nets.items() for k, (W, b) in enumerate(ls) for w, t in (("weight", W), ("bias", b))}

    def _mlp(self, name, final):
        h = self.vr[:, None]
        ls = self.nets[name]
        for k, (W, b) in enumerate(ls):
            h = h @ W.t() + b
            if k < len(ls) - 1:
                h = torch.relu(h)
        return final(h[:, 0])

    def __call__(self, m_f0, f00):
        mag = -self._mlp("mag", torch.relu)                      # from minus inf to 0
        return torch.pow(10.0, mag) * f00 * self._mlp("sign", torch.tanh)


class SphericalHarmonics:
    """f(vx, vy) = f00(|v|) + sum_{l,m} flm(|v|) Re Y_l^m, floored at 1e-32 and normalised (spherical_harmonics.py:287-318).
    f00 = super-Gaussian of order m = 2 + 3 sigmoid(normed_m) on the radial grid vr (:267-285); the radial functions are
    lerped from vr onto |v| of the Cartesian mesh (`right=` fill beyond the last radial node)."""

    def __init__(self, dist_cfg, device, trainable):
        p = dist_cfg["params"]
        self.vx = vgrid(dist_cfg["nvx"])
        self.dv = self.vx[1] - self.vx[0]
        vmax = 6.0 * 1.05 * np.sqrt(2.0)
        nvr = int(p["nvr"])
        dvr = vmax / nvr
        vr = np.linspace(dvr / 2, vmax - dvr / 2, nvr)
        self.dvr = vr[1] - vr[0]
        VX, VY = np.meshgrid(self.vx, self.vx)                      # "xy" indexing, as jnp.meshgrid (:204)
        th = np.arctan2(VY, VX)
        with np.errstate(invalid="ignore", divide="ignore"):
            phi = np.arccos(VY / np.abs(VY))
        r = np.sqrt(VX**2 + VY**2)
        self.shape = r.shape
        self.Nl = int(p["Nl"])
        self.flm_type = p["flm_type"].casefold()
        self.vr = torch.tensor(vr, dtype=DT, device=device)
        i, t, lo, hi = interp_plan(r, vr)
        self._i = torch.tensor(i, dtype=torch.long, device=device)
        self._t = torch.tensor(t, dtype=DT, device=device)
        self._lo = torch.tensor(lo, device=device)
        self._hi = torch.tensor(hi, device=device)
        # trainable order m of f00: always stored through the "inverse" activation and read through the sigmoid (:219-223)
        self.m_scale, self.m_shift = 3.0, 2.0
        self.normed_m = torch.tensor(_inv_act((float(p["init_m"]) - self.m_shift) / self.m_scale), dtype=DT, device=device,
                                     requires_grad=bool(trainable))
        self.flm, self._ylm = {}, {}
        for l in range(1, self.Nl + 1):
            for m in range(l + 1):
                if self.flm_type == "mora-yahi":
                    if (l, m) == (1, 0):
                        self.flm[(l, m)] = _FlmMoraYahi(self.vr, p["LTx"], device, trainable)
                    elif (l, m) == (1, 1):
                        self.flm[(l, m)] = _FlmMoraYahi(self.vr, p["LTy"], device, trainable)
                    else:
                        raise NotImplementedError("Mora-Yahi only supports l=1, m=0 and l=1, m=1")
                elif self.flm_type == "arbitrary":
                    self.flm[(l, m)] = _FlmArbitraryVr(nvr, device, trainable and l == 1)   # base.py:487-499: only l=1 is trained
                elif self.flm_type == "nn":
                    self.flm[(l, m)] = _FlmNN(self.vr, device, trainable, weights=(p.get("nn_weights") or {}).get((l, m)))
                else:
                    raise NotImplementedError(f"Unknown flm_type: {p['flm_type']}")
                # Re Y_l^m(azimuth = phi, polar = th)  (jax.scipy.special.sph_harm(m, n, theta, phi), :310-312):
                # sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) P_l^m(cos th) cos(m phi), Condon-Shortley phase inside P_l^m
                norm = math.sqrt((2 * l + 1) / (4 * math.pi) * math.factorial(l - m) / math.factorial(l + m))
                y = norm * lpmv(m, l, np.cos(th)) * np.cos(m * phi)
                self._ylm[(l, m)] = torch.tensor(y.reshape(-1), dtype=DT, device=device)

    def leaves(self):
        out = {"normed_m": self.normed_m}
        for (l, m), f in self.flm.items():
            for k, v in f.leaves().items():
                out[f"flm[{l}][{m}].{k}"] = v
        return out

    def get_unnormed_m(self):
        return torch.sigmoid(self.normed_m) * self.m_scale + self.m_shift

    def get_f00(self):
        m = self.get_unnormed_m()
        v0 = 1.0 / torch.sqrt(_tgamma(5.0 / m) / 3.0 / _tgamma(3.0 / m))
        cst = m / (4 * math.pi * _tgamma(3.0 / m))
        f00 = cst / v0**3.0 * torch.exp(-((self.vr / v0) ** m))
        return f00 / (torch.sum(f00 * 4 * math.pi * self.vr**2.0) * self.dvr)

    def _to_mesh(self, fr, right):
        v = fr[self._i] * (1.0 - self._t) + fr[self._i + 1] * self._t
        v = torch.where(self._lo, fr[0], v)
        return torch.where(self._hi, torch.full_like(v, right), v)

    def __call__(self):
        f00 = self.get_f00()
        f = self._to_mesh(f00, 1e-16)
        m_f0 = self.get_unnormed_m()
        for key, flm in self.flm.items():
            f = f + self._to_mesh(flm(m_f0, f00), 1e-32) * self._ylm[key]
        f = torch.clamp(f, min=1e-32)
        f = f / (f.sum() * self.dv * self.dv)
        return f.reshape(self.shape)

    def get_unnormed_params(self):
        f00 = self.get_f00()
        m_f0 = self.get_unnormed_m()
        return {"flm": {0: {0: f00}, **{l: {m: self.flm[(l, m)](m_f0, f00) for m in range(l + 1)} for l in range(1, self.Nl + 1)}}}
