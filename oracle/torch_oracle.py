"""torch float64 twin of oracle/np_oracle.py, used for AUTOGRAD gradients (the reference differentiates the
same graph with equinox.filter_value_and_grad, loss_function.py:107-108).  TEST INFRASTRUCTURE ONLY.

The forward arithmetic follows np_oracle function by function (tests assert the two agree to ~1e-13);
gradients are additionally checked against central finite differences of the NumPy oracle.
"""
from __future__ import annotations

import math
import numpy as np
import torch

from . import np_oracle as O

DT = torch.float64
C, ME, MP, RE = O.C, O.ME, O.MP, O.RE


def T(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=DT)


def t_interp(xq, xp, fp):
    """np.interp (edge clamped) for tensors; xp ascending 1-D."""
    xp, fp = T(xp), T(fp)
    i = torch.clamp(torch.searchsorted(xp, xq.detach().contiguous(), right=True), 1, xp.numel() - 1)
    x0, x1 = xp[i - 1], xp[i]
    t = (xq - x0) / (x1 - x0)
    out = fp[i - 1] + t * (fp[i] - fp[i - 1])
    out = torch.where(xq <= xp[0], fp[0].expand_as(out), out)
    out = torch.where(xq >= xp[-1], fp[-1].expand_as(out), out)
    return out


def t_gradient(f, h):
    """np.gradient(f, h) along the last axis."""
    inner = (f[..., 2:] - f[..., :-2]) / (2 * h)
    return torch.cat([((f[..., 1] - f[..., 0]) / h).unsqueeze(-1), inner, ((f[..., -1] - f[..., -2]) / h).unsqueeze(-1)], -1)


def t_cubic(xq, x, f, extrap):
    """interpax cubic (np_oracle.interp1d_cubic)."""
    x = T(x)
    s = (f[1:] - f[:-1]) / (x[1:] - x[:-1])
    fx = torch.cat([s[:1], 0.5 * (s[:-1] + s[1:]), s[-1:]])
    i = torch.clamp(torch.searchsorted(x, xq.detach().contiguous(), right=True), 1, x.numel() - 1)
    dx = x[i] - x[i - 1]
    t = (xq - x[i - 1]) / dx
    f0, f1 = f[i - 1], f[i]
    m0, m1 = fx[i - 1] * dx, fx[i] * dx
    c2 = -3 * f0 + 3 * f1 - 2 * m0 - m1
    c3 = 2 * f0 - 2 * f1 + m0 + m1
    fq = f0 + m0 * t + c2 * t**2 + c3 * t**3
    lo, hi = extrap
    fq = torch.where(xq < x[0], torch.full_like(fq, lo), fq)
    fq = torch.where(xq > x[-1], torch.full_like(fq, hi), fq)
    return fq


def t_ratintn(f, g, z):
    """np_oracle.ratcen/ratintn with real(log(complex ratio)) = log|ratio|.  f [N], g [..., N]."""
    fdif = f[..., 1:-1] - f[..., 0:-2]
    gdif = g[..., 1:-1] - g[..., 0:-2]
    fav = 0.5 * (f[..., 1:-1] + f[..., 0:-2])
    gav = 0.5 * (g[..., 1:-1] + g[..., 0:-2])
    tmp = fav * gdif - gav * fdif
    rfn = fdif / gdif + tmp * torch.log(torch.abs((gav + 0.5 * gdif) / (gav - 0.5 * gdif))) / gdif**2
    # the Taylor branch |gdif| < 1e-4 |gav| is unreachable on these grids; assert instead of branching
    assert not bool((torch.abs(gdif) < 1.0e-4 * torch.abs(gav)).any())
    zdif = T(z)[1:-1] - T(z)[0:-2]
    return torch.sum(rfn * zdif, -1)


def _linspace_factor(grad, G):
    lo, hi = 1 - grad / 200, 1 + grad / 200
    if G == 1:
        return lo.reshape(1)
    w = torch.linspace(0, 1, G, dtype=DT)
    return lo + (hi - lo) * w


def kinematics(p, grids, sa_deg, G, lam_shift):
    """np_oracle._kinematics; p = dict of torch scalars: Te, ne, lam, Va, ud, ne_gradient, Te_gradient,
    ions = list of dicts(A, Z, Ti, fract)."""
    ne = 1.0e20 * p["ne"] * _linspace_factor(p["ne_gradient"], G)
    Te = p["Te"] * _linspace_factor(p["Te_gradient"], G)
    lam = p["lam"] + lam_shift
    A = torch.stack([T(i["A"]) for i in p["ions"]])
    Z = torch.stack([T(i["Z"]) for i in p["ions"]])
    Ti = torch.stack([T(i["Ti"]) for i in p["ions"]])
    fract = torch.stack([T(i["fract"]) for i in p["ions"]])
    Va, ud = p["Va"] * 1e6, p["ud"] * 1e6
    Mi = A * MP
    constants = math.sqrt(4 * math.pi * (ME * C**2 * RE) / ME)
    sarad = (T(np.asarray(sa_deg, dtype=np.float64)) * math.pi / 180).reshape(1, 1, -1)
    omgL = grids.omgL_num / lam
    omgpe = constants * torch.sqrt(ne[:, None, None])
    omgs = T(grids.omgs)
    omg = omgs - omgL
    ks = torch.sqrt(omgs**2 - omgpe**2) / C
    kL = torch.sqrt(omgL**2 - omgpe**2) / C
    k = torch.sqrt(ks**2 + kL**2 - 2 * ks * kL * torch.cos(sarad))
    omgdop = omg - k * Va
    vTe = torch.sqrt(Te[:, None, None] / ME)
    klde = (vTe / omgpe) * k
    Z4, Mi4, fr4 = Z.reshape(1, 1, 1, -1), Mi.reshape(1, 1, 1, -1), fract.reshape(1, 1, 1, -1)
    Zbar = torch.sum(Z4 * fr4)
    ni = fr4 * ne[:, None, None, None] / Zbar
    omgpi = constants * Z4 * torch.sqrt(ni * ME / Mi4)
    vTi = torch.sqrt(Ti / Mi4)
    kldi = (vTi / omgpi) * k[..., None]
    xii = 1.0 / (math.sqrt(2.0) * vTi) * ((omgdop / k)[..., None])
    xie = omgdop / (k * vTe) - ud / vTe
    return dict(ne=ne, omgL=omgL, omgs=omgs, k=k, omgdop=omgdop, vTe=vTe, klde=klde, Z=Z4, fract=fr4, Zbar=Zbar,
                vTi=vTi, kldi=kldi, xii=xii, xie=xie)


def chi_ion(kin, grids):
    xii = kin["xii"]
    xi2 = T(grids.xi2)
    zr = t_interp(xii, xi2, T(grids.Zpi[0]))
    zi = t_interp(xii, xi2, T(grids.Zpi[1]))
    outside = (xii < xi2[0]) | (xii > xi2[-1])
    zr = torch.where(outside, xii**-2.0, zr)
    zi = torch.where(outside, torch.zeros_like(zi), zi)
    c = -0.5 / kin["kldi"] ** 2
    return torch.sum(c * zr, 3), torch.sum(c * zi, 3)


def assemble(kin, chiEr, chiEi, chiIr, chiIi, fe_vphi, grids):
    k, vTe, vTi, xii = kin["k"], kin["vTe"], kin["vTi"], kin["xii"]
    eps2 = (1.0 + chiEr + chiIr) ** 2 + (chiEi + chiIi) ** 2
    ce2 = chiEr**2 + chiEi**2
    ion_comp_fact = kin["fract"] * kin["Z"] ** 2 / kin["Zbar"] / vTi
    ion_comp = ion_comp_fact * (ce2[..., None] * torch.exp(-(xii**2)) / math.sqrt(2 * math.pi))
    ele_comp = ((1.0 + chiIr) ** 2 + chiIi**2) * fe_vphi / vTe
    SKW_ion = torch.sum(1.0 / k[..., None] * ion_comp / eps2[..., None], 3)
    SKW_ele = 1.0 / k * ele_comp / eps2
    PsOmg = (SKW_ion + SKW_ele) * (1 + 2 * kin["omgdop"] / kin["omgL"]) * RE**2.0 * kin["ne"][:, None, None]
    lams = 2 * math.pi * C / kin["omgs"]
    return PsOmg * 2 * math.pi * C / lams**2


def form_factor_1v(p, fe, vx, grids, sa_deg, G=1, lam_shift=0.0):
    """np_oracle.form_factor_1v (table mode).  fe: torch [V]."""
    kin = kinematics(p, grids, sa_deg, G, lam_shift)
    xie, klde = kin["xie"], kin["klde"]
    chiIr, chiIi = chi_ion(kin, grids)
    logf = torch.log(fe)
    fe_vphi = torch.exp(t_cubic(xie, vx, logf, (-50.0, -50.0)))
    df = (fe_vphi[:, 1:, :] - fe_vphi[:, :-1, :]) / (xie[:, 1:, :] - xie[:, :-1, :])
    df = torch.cat([df, torch.zeros_like(df[:, :1, :])], 1)
    chiEi = math.pi / klde**2 * df
    xi1, xi2 = T(grids.xi1), T(grids.xi2)
    ratmod = torch.exp(t_cubic(xi1, vx, logf, (-50.0, -50.0)))
    ratdf = t_gradient(ratmod, grids.xi1[1] - grids.xi1[0])
    prim = t_ratintn(ratdf[None, :], xi1[None, :] - xi2[:, None], xi1)
    chiEr = -1.0 / klde**2 * t_interp(xie, xi2, prim)
    return assemble(kin, chiEr, chiEi, chiIr, chiIi, fe_vphi, grids)


def form_factor_direct(p, fe, vx, grids, sa_deg, G=1, lam_shift=0.0):
    """np_oracle.form_factor_direct."""
    kin = kinematics(p, grids, sa_deg, G, lam_shift)
    xie, klde = kin["xie"], kin["klde"]
    chiIr, chiIi = chi_ion(kin, grids)
    vxt = T(vx)
    dvx = vx[1] - vx[0]
    df = t_gradient(fe, dvx)
    fe_vphi = t_interp(xie, vxt, fe)
    dfe = t_interp(xie, vxt, df)
    chiEi = math.pi / klde**2 * dfe
    flat = xie.reshape(-1)
    rat = torch.cat([t_ratintn(df[None, :], vxt[None, :] - flat[s : s + 256, None], vxt) for s in range(0, flat.numel(), 256)])
    chiEr = -1.0 / klde**2 * rat.reshape(xie.shape)
    return assemble(kin, chiEr, chiEi, chiIr, chiIi, fe_vphi, grids)


def params_from_block(row, nI, requires_grad=True):
    """Build the torch parameter dict from one row of the C-ABI parameter block (include/tsff.h)."""
    leaves = torch.tensor(np.asarray(row, dtype=np.float64), dtype=DT, requires_grad=requires_grad)
    p = dict(Te=leaves[0], ne=leaves[1], lam=leaves[2], Va=leaves[3], ud=leaves[4], ne_gradient=leaves[5],
             Te_gradient=leaves[6], amp1=leaves[7], amp2=leaves[8], amp3=leaves[9], ions=[])
    for i in range(nI):
        o = 10 + 4 * i
        p["ions"].append(dict(A=leaves[o], Z=leaves[o + 1], Ti=leaves[o + 2], fract=leaves[o + 3]))
    return leaves, p


def modl_from_ff(ff, weights, jmul=None):
    """mean over G, weighted angle sum, static multiplier (generate_spectra.py:164-165,193,197,210-216)."""
    m = torch.sum(torch.mean(ff, 0) * T(weights), 1)
    return m if jmul is None else m * T(jmul)


def add_electron_irf(lamAxisE, modlE, amps, lam, amp1, amp2, stddevE):
    """np_oracle.add_electron_irf (irf.py:90-132), norm == 0."""
    lamAxisE = T(lamAxisE)
    origin = (lamAxisE.max() + lamAxisE.min()) / 2.0
    inst = (1.0 / (stddevE * math.sqrt(2.0 * math.pi))) * torch.exp(-((lamAxisE - origin) ** 2.0) / (2.0 * stddevE**2.0))
    n = modlE.numel()
    full = torch.nn.functional.conv1d(modlE.reshape(1, 1, -1), inst.flip(0).reshape(1, 1, -1), padding=n - 1).reshape(-1)
    y = full[(n - 1) // 2 : (n - 1) // 2 + n]          # np.convolve(..., "same")
    y = (modlE.max() / y.max()) * y
    y = y.reshape(1024, -1).mean(1)
    lamb = lamAxisE.reshape(1024, -1).mean(1)
    y = amps * y / y.max()
    return lamb, torch.where(lamb < lam, amp1 * y, amp2 * y)


def loss_electron(ThryE, lamb, e_data, cfg, e_norm=1.0):
    """np_oracle.calc_ei_error electron part with nanmean (loss_function.py:234-264), l2."""
    fr, ex = cfg["data"]["fit_rng"], cfg["other"]["extraoptions"]
    err = (T(e_data) - ThryE) ** 2 / e_norm**2
    tot = 0.0
    if ex["fit_EPWb"]:
        m = (lamb > fr["blue_min"]) & (lamb < fr["blue_max"])
        tot = tot + err[..., m].mean()
    if ex["fit_EPWr"]:
        m = (lamb > fr["red_min"]) & (lamb < fr["red_max"])
        tot = tot + err[..., m].mean()
        if ex["fit_EPWb"]:
            tot = tot * 0.5
    return tot


def _conv_same(x, v):
    """np.convolve(x, v, "same") for equal lengths, along the last axis of x [rows, n]."""
    n = x.shape[-1]
    full = torch.nn.functional.conv1d(x.reshape(-1, 1, n), v.flip(0).reshape(1, 1, -1), padding=n - 1)
    return full[:, 0, (n - 1) // 2:(n - 1) // 2 + n].reshape(x.shape)


def ats_chain(modlE, lamAxisE, angAxis, spect_fwhm, ang_fwhm, lam, amp1, amp2, e_amps, n_lam_data, ccd0, row_start, row_end):
    """np_oracle.add_ats_irf + reduce_ats_to_resunit (irf.py:5-47, thomson_diagnostic.py:78-107) with autograd."""
    from . import np_oracle as O
    inst_lam, inst_ang = T(O.ats_taps(np.asarray(lamAxisE), spect_fwhm)), T(O.ats_taps(np.asarray(angAxis), ang_fwhm))
    y = _conv_same(modlE.t().contiguous(), inst_ang).t()      # along the angle axis
    y = _conv_same(y.contiguous(), inst_lam)                  # along wavelength
    y = modlE.max(dim=1, keepdim=True).values / y.max(dim=1, keepdim=True).values * y
    NA, W = y.shape
    lam_step, ang_step = round(W / n_lam_data), round(NA / ccd0)
    y = torch.stack([y[:, i:i + lam_step].mean(dim=1) for i in range(0, W, lam_step)])            # [nl, NA]
    y = torch.stack([y[:, i:i + ang_step].mean(dim=1) for i in range(0, NA, ang_step)])           # [na, nl]
    lamb = T(np.array([np.average(np.asarray(lamAxisE)[i:i + lam_step]) for i in range(0, W, lam_step)]))
    y = y[row_start:row_end]
    y = T(e_amps).reshape(-1, 1) * y / y.max(dim=1, keepdim=True).values
    return torch.where(lamb < lam, amp1 * y, amp2 * y), lamb


# ---- 2V path (np_oracle.form_factor_2d) with autograd: bicubic rotate/project per pole ---------------------------------
def _t_hermite_weights(xq, x):
    """torch twin of np_oracle._hermite_node_weights on a UNIFORM grid x: idx [n,4] (clamped), w [n,4] differentiable in xq."""
    x = T(x)
    n = x.numel()
    h = x[1] - x[0]
    i = torch.clamp(torch.searchsorted(x, xq.detach().contiguous(), right=True), 1, n - 1)
    t = (xq - x[i - 1]) / h
    h00, h01 = 2 * t**3 - 3 * t**2 + 1, -2 * t**3 + 3 * t**2
    h10, h11 = t**3 - 2 * t**2 + t, t**3 - t**2          # in units of h (slopes are multiplied by h)
    z = torch.zeros_like(t)
    left, right = (i == 1), (i == n - 1)
    w0 = torch.where(left, z, -0.5 * h10)
    w1 = torch.where(left, h00 - h10 - 0.5 * h11, torch.where(right, h00 - h11, h00 - 0.5 * h11))
    w2 = torch.where(left, h01 + h10, torch.where(right, h01 + 0.5 * h10 + h11, h01 + 0.5 * h10))
    w3 = torch.where(right, z, 0.5 * h11)
    idx = torch.stack([i - 2, i - 1, i, i + 1], dim=1).clamp(0, n - 1)
    return idx, torch.stack([w0, w1, w2, w3], dim=1)


def t_interp2d_cubic(xq, yq, x, y, f):
    ix, wx = _t_hermite_weights(xq, x)
    iy, wy = _t_hermite_weights(yq, y)
    out = 0.0
    for a in range(4):
        for b in range(4):
            out = out + wx[:, a] * wy[:, b] * f[ix[:, a], iy[:, b]]
    return out


def form_factor_2d(p, fe2d, vx, grids, sa_deg, G=1, lam_shift=0.0, ud_ang=0.0, va_ang=0.0):
    """torch twin of np_oracle.form_factor_2d (form_factor.py:449-587).  p as in `kinematics`; fe2d [V,V] tensor."""
    ne = 1.0e20 * p["ne"] * _linspace_factor(p["ne_gradient"], G)
    Te = p["Te"] * _linspace_factor(p["Te_gradient"], G)
    lam = p["lam"] + lam_shift
    A = torch.stack([T(i["A"]) for i in p["ions"]])
    Z = torch.stack([T(i["Z"]) for i in p["ions"]])
    Ti = torch.stack([T(i["Ti"]) for i in p["ions"]])
    fract = torch.stack([T(i["fract"]) for i in p["ions"]])
    Va, ud = p["Va"] * 1e6, p["ud"] * 1e6
    Vax, Vay = Va * math.cos(va_ang * math.pi / 180), Va * math.sin(va_ang * math.pi / 180)
    udx, udy = ud * math.cos(ud_ang * math.pi / 180), ud * math.sin(ud_ang * math.pi / 180)
    Mi = A * MP
    constants = math.sqrt(4 * math.pi * (ME * C**2 * RE) / ME)
    sarad = (T(np.asarray(sa_deg, dtype=np.float64)) * math.pi / 180).reshape(1, 1, -1)
    omgL = grids.omgL_num / lam
    omgpe = constants * torch.sqrt(ne[:, None, None])
    omgs = T(grids.omgs)
    omg = omgs - omgL
    kLx = torch.sqrt(omgL**2 - omgpe**2) / C
    ks_mag = torch.sqrt(omgs**2 - omgpe**2) / C
    kx, ky = torch.cos(sarad) * ks_mag - kLx, torch.sin(sarad) * ks_mag + 0.0 * kLx
    k = torch.sqrt(kx * kx + ky * ky)
    omgdop = omg - (kx * Vax + ky * Vay)
    vTe = torch.sqrt(Te[:, None, None] / ME)
    klde = (vTe / omgpe) * k
    Z4, Mi4, fr4 = Z.reshape(1, 1, 1, -1), Mi.reshape(1, 1, 1, -1), fract.reshape(1, 1, 1, -1)
    Zbar = torch.sum(Z4 * fr4)
    ni = fr4 * ne[:, None, None, None] / Zbar
    omgpi = constants * Z4 * torch.sqrt(ni * ME / Mi4)
    vTi = torch.sqrt(Ti / Mi4)
    kldi = (vTi / omgpi) * k[..., None]
    xii = 1.0 / (math.sqrt(2.0) * vTi) * ((omgdop / k)[..., None])
    kin = dict(ne=ne, omgL=omgL, omgs=omgs, k=k, omgdop=omgdop, vTe=vTe, klde=klde, Z=Z4, fract=fr4, Zbar=Zbar, vTi=vTi,
               kldi=kldi, xii=xii)
    chiIr, chiIi = chi_ion(kin, grids)
    xiex = ((omgdop / k**2) * kx - udx) / vTe
    xiey = ((omgdop / k**2) * ky - udy) / vTe
    xmag = torch.sqrt(xiex**2 + xiey**2)
    beta = torch.atan(xiey / xiex) + math.pi * (xiex < 0).to(DT)
    vxt = T(vx)
    dv = float(vx[1] - vx[0])
    V = len(vx)
    shp = beta.shape
    fphi_l, chiEi_l, chiEr_l = [], [], []
    va, vb = torch.meshgrid(vxt, vxt, indexing="ij")          # F[a][b]: x = vx[a], y = vx[b]
    for bq, xm, kl in zip(beta.reshape(-1), xmag.reshape(-1), (klde * torch.ones_like(beta)).reshape(-1)):
        cb, sb = torch.cos(bq), torch.sin(bq)
        xq = (cb * va - sb * vb).reshape(-1)
        yq = (sb * va + cb * vb).reshape(-1)
        F = t_interp2d_cubic(xq, yq, vx, vx, fe2d).reshape(V, V)
        f1 = F.sum(dim=0) * dv
        df = t_gradient(f1, dv)
        fphi_l.append(t_interp(xm.reshape(1), vx, f1)[0])
        dfe = t_interp(xm.reshape(1), vx, df)[0]
        chiEi_l.append(math.pi / kl**2 * dfe)
        chiEr_l.append(-1.0 / kl**2 * t_ratintn(df, vxt - xm, vx))
    fphi = torch.stack(fphi_l).reshape(shp)
    chiEi = torch.stack(chiEi_l).reshape(shp)
    chiEr = torch.stack(chiEr_l).reshape(shp)
    return assemble(kin, chiEr, chiEi, chiIr, chiIi, fphi, grids)
