"""NumPy float64 restatement of the parameter producer that feeds the hot path
(tsadar/core/modules/ts_params.py and distribution_functions/base.py).  TEST INFRASTRUCTURE ONLY.

The reference keeps this stage in JAX (out of scope for the kernels) but its
exact arithmetic decides the *inputs* of the golden vector, so the oracle mirrors it.
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammaincc, gamma


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def act_pair(param_cfg, activate):
    """get_act_and_inv_act (ts_params.py:329-350).  NB the 'inverse' is not the logit:
    stored = log(0.01 + x/(1-x+0.01)) so sigmoid(stored) != x."""
    if param_cfg["active"] and activate:
        return sigmoid, (lambda x: np.log(1e-2 + x / (1 - x + 1e-2)))
    return (lambda x: x), (lambda x: x)


def scalar_physical(param_cfg, activate):
    """normalise -> store -> de-normalise for one scalar (ts_params.py:93-104, 202-205, 434-495)."""
    act, inv = act_pair(param_cfg, activate)
    scale = param_cfg["ub"] - param_cfg["lb"]
    shift = param_cfg["lb"]
    stored = inv((param_cfg["val"] - shift) / scale)
    return act(stored) * scale + shift


def vgrid(nvx):
    """DistributionFunction1V grid (base.py:149-151)."""
    vmax = 6.0
    dv = 2 * vmax / nvx
    return np.linspace(-vmax + dv / 2, vmax - dv / 2, nvx)


_DLM_CACHE = {}


def dlm_table():
    """Regenerated stand-in for the missing blob external/numDistFuncs/DLM_x_-3_-10_10_m_-1_2_5.mat
    (base.py:266-268): IT[20001, 31] on vx_ax = linspace(-10,10,20001), m_ax = linspace(2,5,31).

    IT(v,m) = projection of exp(-(|v|/(alpha*sqrt2))^m) on one axis = Gamma(2/m) Q(2/m, (|v|/(alpha sqrt2))^m),
    alpha = sqrt(3 Gamma(3/m) / (2 Gamma(5/m))).  Normalisation is irrelevant
    (base.py:294 renormalises)."""
    if "IT" not in _DLM_CACHE:
        vx_ax = np.linspace(-10, 10, 20001)
        m_ax = np.linspace(2, 5, 31)
        IT = np.empty((vx_ax.size, m_ax.size))
        for j, m in enumerate(m_ax):
            alpha = np.sqrt(3.0 * gamma(3.0 / m) / 2.0 / gamma(5.0 / m))
            IT[:, j] = gamma(2.0 / m) * gammaincc(2.0 / m, (np.abs(vx_ax) / (alpha * np.sqrt(2.0))) ** m)
        _DLM_CACHE["IT"] = (vx_ax, m_ax, IT)
    return _DLM_CACHE["IT"]


def dlm1v(nvx, m):
    """DLM1V.__call__ (base.py:269-272, 277-294): lerp the table in v onto vx, lerp in m, normalise."""
    vx = vgrid(nvx)
    vx_ax, m_ax, IT = dlm_table()
    f_vx_m = np.stack([np.interp(vx, vx_ax, IT[:, j]) for j in range(m_ax.size)], axis=1)
    fdlm = np.array([np.interp(m, m_ax, f_vx_m[i, :]) for i in range(vx.size)])
    return vx, fdlm / np.sum(fdlm) / (vx[1] - vx[0])


def dlm_m_physical(fe_cfg, activate):
    """DLM1V.__init__ / get_unnormed_params (base.py:255-265, 274-275): scale 3, shift 2."""
    if activate and fe_cfg["active"]:
        inv = lambda x: np.log(1e-2 + x / (1 - x + 1e-2))
        act = sigmoid
    else:
        inv = act = lambda x: x
    return act(inv((fe_cfg["params"]["m"]["val"] - 2.0) / 3.0)) * 3.0 + 2.0


def maxwellian1v(nvx):
    """'mx' type (ts_params.py:135-143)."""
    vx = vgrid(nvx)
    f = np.exp(-(vx**2 / 2))
    return vx, f / np.sum(f) / (vx[1] - vx[0])


def super_gaussian_projected(vx, m):
    """Analytic projected super-Gaussian of order m on an arbitrary grid (SURVEY.md §8d synthetic
    sweep): f ∝ Gamma(2/m) Q(2/m, (|v|/(alpha sqrt2))^m), normalised sum(f) dv = 1."""
    alpha = np.sqrt(3.0 * gamma(3.0 / m) / 2.0 / gamma(5.0 / m))
    f = gamma(2.0 / m) * gammaincc(2.0 / m, (np.abs(vx) / (alpha * np.sqrt(2.0))) ** m)
    return f / np.sum(f) / (vx[1] - vx[0])


def second_order_butterworth(signal, f_sampling=100, f_cutoff=15, method="forward_backward"):
    """base.py:41-96 (direct-form IIR with the reference's start-up convention)."""
    if method == "forward_backward":
        signal = second_order_butterworth(signal, f_sampling, f_cutoff, "forward")
        return second_order_butterworth(signal, f_sampling, f_cutoff, "backward")
    if method == "backward":
        signal = signal[::-1]
    ff = f_cutoff / f_sampling
    ita = 1.0 / np.tan(np.pi * ff)
    q = np.sqrt(2.0)
    b0 = 1.0 / (1.0 + q * ita + ita**2)
    b1, b2 = 2 * b0, b0
    a1 = 2.0 * (ita**2 - 1.0) * b0
    a2 = -(1.0 - q * ita + ita**2) * b0
    x1, x2, y1, y2 = signal[1], signal[0], signal[1], signal[0]
    out = []
    for x in signal[2:]:
        y = b0 * x + b1 * x1 + b2 * x2 + a1 * y1 + a2 * y2
        x1, x2, y1, y2 = x, x1, y, y1
        out.append(y)
    out = np.array(out)
    out = np.concatenate((out[0:1], out[0:1], out))
    if method == "backward":
        out = out[::-1]
    return out


def arbitrary1v_init(nvx, init_m):
    """Arbitrary1V.init_dlm (base.py:188-196)."""
    vx = vgrid(nvx)
    alpha = np.sqrt(3.0 * gamma(3.0 / init_m) / 2.0 / gamma(5.0 / init_m))
    cst = init_m / (4.0 * np.pi * alpha**3.0 * gamma(3.0 / init_m))
    fdlm = cst * np.exp(-(np.abs(vx / alpha) ** init_m))
    fdlm = fdlm / np.sum(fdlm) / (vx[1] - vx[0])
    return vx, np.sqrt(-np.log10(fdlm)) / 7.0


def arbitrary1v_call(vx, fval):
    """Arbitrary1V.__call__ (base.py:201-204)."""
    f = (7.0 * second_order_butterworth(fval, 100, 6, "forward_backward")) ** 2.0
    f = np.power(10.0, -f)
    return f / np.sum(f) / (vx[1] - vx[0])


# ---- 2V producers (base.py:335-426, spherical_harmonics.py:59-318) ---------------------------------------------------
def smooth1d(array, window_size):
    """base.py:17-38."""
    window = np.hanning(window_size)
    window = window / window.sum()
    return np.convolve(array, window, mode="same")


def arbitrary2v_init(nvx, init_m, learn_log):
    """Arbitrary2V.init_dlm (base.py:374-405)."""
    vx = vgrid(nvx)
    vth_x = np.sqrt(2.0)
    alpha = np.sqrt(3.0 * gamma(3.0 / init_m) / 2.0 / gamma(5.0 / init_m))
    cst = init_m / (4.0 * np.pi * alpha**3.0 * gamma(3.0 / init_m))
    fdlm = cst / vth_x**3.0 * np.exp(-((np.sqrt(vx[:, None] ** 2.0 + vx[None, :] ** 2.0) / alpha / vth_x) ** init_m))
    fdlm = fdlm / np.sum(fdlm) / (vx[1] - vx[0]) ** 2.0
    if learn_log:
        fdlm = -np.log10(fdlm)
    return vx, np.sqrt(fdlm)


def arbitrary2v_call(vx, fval, learn_log):
    """Arbitrary2V.__call__ (base.py:410-426)."""
    f = fval**2.0
    if learn_log:
        f = np.power(10.0, -f)
    return f / np.sum(f) / (vx[1] - vx[0]) ** 2.0


def flm_mora_yahi(vr, log_10_LT, m_f0, f00):
    """FLM_MY.__call__ (spherical_harmonics.py:94-117)."""
    v0 = 1.0
    lambda_e = 1.0
    ve = gamma(5.0 / m_f0) / 3 / gamma(3.0 / m_f0) * v0
    uu = vr / v0
    lambda_v = lambda_e * (vr / ve) ** 4.0
    coeff = (m_f0 / 2 * uu**m_f0 - 5 * m_f0 / 12 * gamma(8 / m_f0) / gamma(6 / m_f0) * uu ** (m_f0 - 2) - 1.5) * lambda_v
    return coeff / 10**log_10_LT * f00


def flm_arbitrary_vr(flm_sign, flm_mag):
    """ArbitraryVr.__call__ (spherical_harmonics.py:142-147)."""
    nvr = flm_sign.size
    sign = np.tanh(smooth1d(flm_sign, nvr // 4))
    mag = -sigmoid(smooth1d(flm_mag, nvr // 4)) * 10
    return 10**mag * sign


def flm_nn(vr, f00, mag_weights, sign_weights):
    """FLM_NN.__call__ (spherical_harmonics.py:41-49): two eqx.nn.MLP(in 1, out 1, width 32, depth 3, relu between the layers)
    evaluated at every vr: flm = 10^(-relu-MLP(vr)) * f00 * tanh-MLP(vr).  weights: list of (W [out, in], b [out]) per layer."""
    def mlp(ws, x, final):
        h = x[:, None]
        for k, (W, b) in enumerate(ws):
            h = h @ np.asarray(W).T + np.asarray(b)
            if k < len(ws) - 1:
                h = np.maximum(h, 0.0)
        return final(h[:, 0])
    mag = -mlp(mag_weights, np.asarray(vr, dtype=np.float64), lambda z: np.maximum(z, 0.0))
    return np.power(10.0, mag) * f00 * mlp(sign_weights, np.asarray(vr, dtype=np.float64), np.tanh)


def spherical_harmonics_fe(dist_cfg, normed_m=None, flm_leaves=None):
    """SphericalHarmonics.__init__ + __call__ (spherical_harmonics.py:199-247, 267-318) -> vx, f[V, V].
    `normed_m` / `flm_leaves[(l, m)]` override the initial trainable values ({"log_10_LT": x} for mora-yahi,
    {"flm_sign": a, "flm_mag": b} for arbitrary).  jax.scipy.special.sph_harm(m, n, theta=azimuth, phi=polar) is restated
    with scipy.special.sph_harm_y(n, m, polar, azimuth); JAX builds P_l^m from sqrt(1 - cos^2(polar)) >= 0, i.e. the polar
    angle folded into [0, pi], whatever the sign of arctan2(vy, vx)."""
    from scipy.special import sph_harm_y
    p = dist_cfg["params"]
    vx = vgrid(dist_cfg["nvx"])
    vmax = 6.0 * 1.05 * np.sqrt(2.0)
    dvr = vmax / p["nvr"]
    vr = np.linspace(dvr / 2, vmax - dvr / 2, p["nvr"])
    VX, VY = np.meshgrid(vx, vx)
    th = np.arctan2(VY, VX)
    phi = np.arccos(VY / np.abs(VY))
    vr_vxvy = np.sqrt(VX**2 + VY**2)
    if normed_m is None:
        normed_m = np.log(1e-2 + ((p["init_m"] - 2.0) / 3.0) / (1 - (p["init_m"] - 2.0) / 3.0 + 1e-2))
    m_f0 = sigmoid(normed_m) * 3.0 + 2.0
    v0 = 1.0 / np.sqrt(gamma(5.0 / m_f0) / 3.0 / gamma(3.0 / m_f0))
    cst = m_f0 / (4 * np.pi * gamma(3.0 / m_f0))
    f00 = cst / v0**3.0 * np.exp(-((vr / v0) ** m_f0))
    f00 = f00 / (np.sum(f00 * 4 * np.pi * vr**2.0) * (vr[1] - vr[0]))
    f = np.interp(vr_vxvy, vr, f00, right=1e-16)
    typ = p["flm_type"].casefold()
    for l in range(1, p["Nl"] + 1):
        for m in range(l + 1):
            lv = (flm_leaves or {}).get((l, m), {})
            if typ == "mora-yahi":
                LT = {(1, 0): p["LTx"], (1, 1): p["LTy"]}[(l, m)]
                flm = flm_mora_yahi(vr, lv.get("log_10_LT", np.log10(LT)), m_f0, f00)
            elif typ == "arbitrary":
                flm = flm_arbitrary_vr(lv.get("flm_sign", np.zeros(p["nvr"])), lv.get("flm_mag", np.zeros(p["nvr"])))
            elif typ == "nn":      # the MLP weights must be given: JAX's PRNG stream (PRNGKey(0) / PRNGKey(42)) is not available here
                flm = flm_nn(vr, f00, lv["mag_weights"], lv["sign_weights"])
            else:
                raise NotImplementedError(typ)
            flm_xy = np.interp(vr_vxvy, vr, flm, right=1e-32)
            ylm = sph_harm_y(l, m, np.arccos(np.cos(th.reshape(-1))), phi.reshape(-1)).reshape(vr_vxvy.shape)
            f = f + flm_xy * np.real(ylm)
    f = np.maximum(f, 1e-32)
    f = f / (np.sum(f) * (vx[1] - vx[0]) * (vx[1] - vx[0]))
    return vx, f


def thomson_params(param_cfg, activate=True, dlm_m_offset=0.0):
    """ThomsonParams(...)() for ONE lineout (ts_params.py:583-603): nested dict of physical scalars +
    fe/v.  ``dlm_m_offset`` shifts the DLM order m (used only to quantify the effect of the missing
    table blob)."""
    ecfg = param_cfg["electron"]
    out = {"electron": {}, "general": {}}
    for p in ["Te", "ne"]:
        out["electron"][p] = scalar_physical(ecfg[p], activate)
    fe_cfg = ecfg["fe"]
    typ = fe_cfg["type"].casefold()
    if fe_cfg["dim"] == 2:
        if "sph" in typ:
            vx, fe = spherical_harmonics_fe(fe_cfg)
        elif typ == "arbitrary":
            vx, fval = arbitrary2v_init(fe_cfg["nvx"], fe_cfg["params"]["init_m"], fe_cfg["params"]["learn_log"])
            fe = arbitrary2v_call(vx, fval, fe_cfg["params"]["learn_log"])
        else:
            raise NotImplementedError(typ)
    elif typ == "dlm":
        m = dlm_m_physical(fe_cfg, activate) + dlm_m_offset
        vx, fe = dlm1v(fe_cfg["nvx"], m)
        out["electron"]["m"] = m
    elif typ == "mx":
        vx, fe = maxwellian1v(fe_cfg["nvx"])
    elif typ == "arbitrary":
        vx, fval = arbitrary1v_init(fe_cfg["nvx"], fe_cfg["params"]["init_m"])
        fe = arbitrary1v_call(vx, fval)
    else:
        raise NotImplementedError(typ)
    out["electron"]["fe"], out["electron"]["v"] = fe, vx
    for p in ["lam", "amp1", "amp2", "amp3", "ne_gradient", "Te_gradient", "ud", "Va"]:
        out["general"][p] = scalar_physical(param_cfg["general"][p], activate)
    ions = sorted([k for k in param_cfg if k.startswith("ion-")], key=lambda s: int(s.split("-")[1]))
    for k in ions:
        icfg = param_cfg[k]
        act_f, inv_f = act_pair(icfg["fract"], activate)
        out[k] = {
            "A": icfg["A"]["val"],
            "fract": act_f(inv_f(icfg["fract"]["val"])),  # ts_params.py:288,298,323 (raw val through act)
            "Ti": scalar_physical(icfg["Ti"], activate),
            "Z": scalar_physical(icfg["Z"], activate),
        }
    # renormalize_ions (ts_params.py:543-563)
    fsum = 0.0
    for n, k in enumerate(ions):
        if n > 0 and param_cfg[k]["Ti"].get("same", False):
            out[k]["Ti"] = out["ion-1"]["Ti"]
        fsum += out[k]["fract"]
    for k in ions:
        out[k]["fract"] = out[k]["fract"] / fsum
    return out
