"""NumPy float64 restatement of the TSADAR form-factor hot path.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py).  Every function cites the reference file:line it follows; the
arithmetic is kept in the reference's own form (e.g. ``ratcen`` uses the reference's complex-log
expression, *not* the summation-by-parts form the CUDA kernels use) so that the oracle is an
independent statement of the algorithm.

All paths are relative to the reference root (ergodicio/tsadar).
"""
from __future__ import annotations

import os
import numpy as np

# --------------------------------------------------------------------------------------
# constants (form_factor.py:123-125, 207-209)
# --------------------------------------------------------------------------------------
C = 2.99792458e10  # cm/s
ME = 510.9896 / C**2  # keV s^2/cm^2
MP = ME * 1836.1
RE = 2.8179e-13  # cm

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tsadar_b200", "data")


# --------------------------------------------------------------------------------------
# third-party arithmetic restated (SURVEY.md §8c, A-note 1)
# --------------------------------------------------------------------------------------
def cubic_slopes(x, f):
    """interpax ``approx_df(method="cubic")``: node slopes = mean of adjacent secants,
    one-sided at the ends (call sites form_factor.py:256,263)."""
    s = np.diff(f) / np.diff(x)
    return np.concatenate([s[:1], 0.5 * (s[:-1] + s[1:]), s[-1:]])


def interp1d_cubic(xq, x, f, extrap):
    """interpax.interp1d(xq, x, f, method="cubic", extrap=[lo, hi]) (form_factor.py:256,263).

    C1 cubic Hermite; cell index clip(searchsorted(x, xq, 'right'), 1, n-1); outside
    [x0, x_{n-1}] the result is replaced by lo / hi."""
    xq = np.asarray(xq, dtype=np.float64)
    fx = cubic_slopes(x, f)
    i = np.clip(np.searchsorted(x, xq, side="right"), 1, len(x) - 1)
    dx = x[i] - x[i - 1]
    t = (xq - x[i - 1]) / dx
    f0, f1 = f[i - 1], f[i]
    m0, m1 = fx[i - 1] * dx, fx[i] * dx
    c2 = -3 * f0 + 3 * f1 - 2 * m0 - m1
    c3 = 2 * f0 - 2 * f1 + m0 + m1
    fq = f0 + m0 * t + c2 * t**2 + c3 * t**3
    lo, hi = extrap
    fq = np.where(xq < x[0], lo, fq)
    fq = np.where(xq > x[-1], hi, fq)
    return fq


def interp_lr(xq, xp, fp, left=None, right=None):
    """jnp.interp with array-valued left/right fills (form_factor.py:247-248)."""
    out = np.interp(xq, xp, fp)
    if left is not None:
        out = np.where(xq < xp[0], left, out)
    if right is not None:
        out = np.where(xq > xp[-1], right, out)
    return out


# --------------------------------------------------------------------------------------
# a1: Z' table (form_factor.py:20-45)
# --------------------------------------------------------------------------------------
def zprime_maxw(xi):
    """zprimeMaxw: linear interpolation of the 2001-point rdWT/idWT tables; xi^-2 / 0 outside
    +-10 (never reached for xi2)."""
    tab = np.load(os.path.join(_DATA, "zprime_table.npz"))
    x, re, im = tab["x"], tab["re"], tab["im"]
    ai, bi = xi < -10, xi > 10
    mid = ~(ai | bi)
    rz = np.concatenate((xi[ai] ** -2.0, np.interp(xi[mid], x, re), xi[bi] ** -2.0))
    iz = np.concatenate((0 * xi[ai], np.interp(xi[mid], x, im), 0 * xi[bi]))
    return np.vstack((rz, iz))


# --------------------------------------------------------------------------------------
# a2: static grids (form_factor.py:120-140)
# --------------------------------------------------------------------------------------
class Grids:
    def __init__(self, lambda_range, npts):
        self.npts = int(npts)
        self.lam_axis = np.linspace(lambda_range[0], lambda_range[1], self.npts)
        self.omgL_num = 2 * np.pi * 1e7 * C
        self.omgs = (2e7 * np.pi * C / self.lam_axis)[None, :, None]
        minmax, h1 = 8.2, 1024
        self.xi1 = np.linspace(-minmax - np.sqrt(2.0) / h1, minmax + np.sqrt(2.0) / h1, h1)
        self.xi2 = np.arange(-minmax, minmax, 0.01)
        self.Zpi = zprime_maxw(self.xi2)


# --------------------------------------------------------------------------------------
# a4: ratintn / ratcen (ratintn.py:4-52)
# --------------------------------------------------------------------------------------
def ratcen(f, g):
    """ratintn.py:26-52.  f [N], g [..., N] -> [..., N-2].  NB the ``[1:-1]-[0:-2]`` slices
    drop the last interval (N-2 intervals)."""
    fdif = f[..., 1:-1] - f[..., 0:-2]
    gdif = g[..., 1:-1] - g[..., 0:-2]
    fav = 0.5 * (f[..., 1:-1] + f[..., 0:-2])
    gav = 0.5 * (g[..., 1:-1] + g[..., 0:-2])
    tmp = fav * gdif - gav * fdif
    rf = fav / gav + tmp * gdif / (12.0 * gav**3)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = ((gav + 0.5 * gdif) / (gav - 0.5 * gdif)).astype(np.complex128)
        rfn = fdif / gdif + tmp * np.log(ratio) / gdif**2
    out = np.where(np.abs(gdif) < 1.0e-4 * np.abs(gav), rf, rfn)
    return np.real(out)


def ratintn(f, g, z):
    """ratintn.py:4-23: sum(ratcen(f,g) * (z[1:-1]-z[0:-2]))."""
    zdif = z[1:-1] - z[0:-2]
    return np.sum(ratcen(f, g) * zdif, axis=-1)


def pv_table(ratdf, xi1, xi2, chunk=256):
    """form_factor.py:266-268: vmap(ratintn)(ratdf, xi1[None]-xi2[:,None], xi1) -> [len(xi2)]."""
    out = np.empty(len(xi2))
    for s in range(0, len(xi2), chunk):
        g = xi1[None, :] - xi2[s : s + chunk, None]
        out[s : s + chunk] = ratintn(ratdf[None, :], g, xi1)
    return out


# --------------------------------------------------------------------------------------
# kinematics shared by a3 / direct mode (form_factor.py:182-243)
# --------------------------------------------------------------------------------------
def _ion_lists(params):
    keys = [k for k in params.keys() if "ion" in k]
    A = np.array([float(params[k]["A"]) for k in keys])
    Z = np.array([float(params[k]["Z"]) for k in keys])
    Ti = np.array([float(params[k]["Ti"]) for k in keys])
    fr = np.array([float(params[k]["fract"]) for k in keys])
    return A, Z, Ti, fr


def _kinematics(params, grids, sa_deg, num_grad_points, lam_shift):
    G = num_grad_points
    gen, ele = params["general"], params["electron"]
    ne = 1.0e20 * float(ele["ne"]) * np.linspace(1 - float(gen["ne_gradient"]) / 200, 1 + float(gen["ne_gradient"]) / 200, G)
    Te = float(ele["Te"]) * np.linspace(1 - float(gen["Te_gradient"]) / 200, 1 + float(gen["Te_gradient"]) / 200, G)
    lam = float(gen["lam"]) + lam_shift
    A, Z, Ti, fract = _ion_lists(params)
    Va = float(gen["Va"]) * 1e6
    ud = float(gen["ud"]) * 1e6
    Mi = A * MP
    Esq = ME * C**2 * RE
    constants = np.sqrt(4 * np.pi * Esq / ME)
    sarad = (np.asarray(sa_deg, dtype=np.float64) * np.pi / 180).reshape(1, 1, -1)
    omgL = grids.omgL_num / lam
    omgpe = constants * np.sqrt(ne[:, None, None])
    omgs = grids.omgs
    omg = omgs - omgL
    ks = np.sqrt(omgs**2 - omgpe**2) / C
    kL = np.sqrt(omgL**2 - omgpe**2) / C
    k = np.sqrt(ks**2 + kL**2 - 2 * ks * kL * np.cos(sarad))
    omgdop = omg - k * Va
    vTe = np.sqrt(Te[:, None, None] / ME)
    klde = (vTe / omgpe) * k
    Z4 = Z.reshape(1, 1, 1, -1)
    Mi4 = Mi.reshape(1, 1, 1, -1)
    fr4 = fract.reshape(1, 1, 1, -1)
    Zbar = np.sum(Z4 * fr4)
    ni = fr4 * ne[:, None, None, None] / Zbar
    omgpi = constants * Z4 * np.sqrt(ni * ME / Mi4)
    vTi = np.sqrt(Ti / Mi4)
    kldi = (vTi / omgpi) * k[..., None]
    xii = 1.0 / (np.sqrt(2.0) * vTi) * ((omgdop / k)[..., None])
    xie = omgdop / (k * vTe) - ud / vTe
    return dict(
        ne=ne, Te=Te, omgL=omgL, omgpe=omgpe, omgs=omgs, k=k, omgdop=omgdop, vTe=vTe, klde=klde,
        Z=Z4, fract=fr4, Zbar=Zbar, vTi=vTi, kldi=kldi, xii=xii, xie=xie,
    )


def _chi_ion(kin, grids):
    """form_factor.py:247-249."""
    xii = kin["xii"]
    ZpiR = interp_lr(xii, grids.xi2, grids.Zpi[0], left=xii**-2.0, right=xii**-2.0)
    ZpiI = interp_lr(xii, grids.xi2, grids.Zpi[1], left=0.0, right=0.0)
    return np.sum(-0.5 / (kin["kldi"] ** 2) * (ZpiR + 1j * ZpiI), 3)


def _assemble(kin, chiE, chiI, fe_vphi, grids):
    """form_factor.py:273-296 (identical at :562-584)."""
    k, vTe, vTi, xii = kin["k"], kin["vTe"], kin["vTi"], kin["xii"]
    epsilon = 1.0 + chiE + chiI
    ion_comp_fact = kin["fract"] * kin["Z"] ** 2 / kin["Zbar"] / vTi
    ion_comp = ion_comp_fact * ((np.abs(chiE[..., None])) ** 2.0 * np.exp(-(xii**2)) / np.sqrt(2 * np.pi))
    ele_comp = (np.abs(1.0 + chiI)) ** 2.0 * fe_vphi / vTe
    SKW_ion = np.sum(1.0 / k[..., None] * ion_comp / ((np.abs(epsilon[..., None])) ** 2), 3)
    SKW_ele = 1.0 / k * ele_comp / ((np.abs(epsilon)) ** 2)
    PsOmg = (SKW_ion + SKW_ele) * (1 + 2 * kin["omgdop"] / kin["omgL"]) * RE**2.0 * kin["ne"][:, None, None]
    lams = 2 * np.pi * C / grids.omgs
    PsLam = PsOmg * 2 * np.pi * C / lams**2
    return PsLam, lams


# --------------------------------------------------------------------------------------
# a3: FormFactor.__call__ (form_factor.py:163-298), "table mode"
# --------------------------------------------------------------------------------------
def form_factor_1v(params, grids, sa_deg, num_grad_points=1, lam_shift=0.0, return_parts=False):
    """params: {"electron": {Te, ne, fe[V], v[V]}, "general": {...}, "ion-k": {A, Z, Ti, fract}}
    (scalars = one lineout, as seen under the reference's vmap).  Returns formfactor [G,W,A],
    lams [1,W,1]."""
    kin = _kinematics(params, grids, sa_deg, num_grad_points, lam_shift)
    fe = np.asarray(params["electron"]["fe"], dtype=np.float64).reshape(-1)
    vx = np.asarray(params["electron"]["v"], dtype=np.float64).reshape(-1)
    xie, klde = kin["xie"], kin["klde"]
    chiI = _chi_ion(kin, grids)
    logf = np.log(fe)
    fe_vphi = np.exp(interp1d_cubic(xie, vx, logf, extrap=[-50, -50]))  # :256
    df = np.diff(fe_vphi, 1, 1) / np.diff(xie, 1, 1)  # :258
    df = np.append(df, np.zeros((df.shape[0], 1, df.shape[2])), 1)  # :259
    chiEI = np.pi / (klde**2) * 1j * df  # :261
    ratmod = np.exp(interp1d_cubic(grids.xi1, vx, logf, extrap=[-50, -50]))  # :263
    ratdf = np.gradient(ratmod, grids.xi1[1] - grids.xi1[0])  # :264
    prim = pv_table(ratdf, grids.xi1, grids.xi2)  # :266-268
    chiERrat = np.interp(xie.flatten(), grids.xi2, prim).reshape(xie.shape)  # :270
    chiERrat = -1.0 / (klde**2) * chiERrat  # :271
    chiE = chiERrat + chiEI
    out, lams = _assemble(kin, chiE, chiI, fe_vphi, grids)
    if return_parts:
        return out, lams, dict(kin=kin, chiE=chiE, chiI=chiI, fe_vphi=fe_vphi, prim=prim, ratdf=ratdf)
    return out, lams


# --------------------------------------------------------------------------------------
# direct-pole mode: 1V kinematics with calc_chi_vals-style susceptibility
# (form_factor.py:349-388 with the projected 1-D table given; SURVEY.md §8d synthetic sweep)
# --------------------------------------------------------------------------------------
def chi_vals_1d(vx, f1d, xie, klde):
    """calc_chi_vals (form_factor.py:369-388) after the rotate/project step, vectorised over
    poles: df = gradient(f1d); fe_vphi = lerp(xie); dfe = lerp(xie; df);
    chiEI = pi/klde^2 dfe; chiERrat = -1/klde^2 ratintn(df, vx - xie, vx)."""
    dvx = vx[1] - vx[0]
    df = np.gradient(f1d, dvx)  # :372
    shp = np.shape(xie)
    x = np.reshape(xie, -1)
    fe_vphi = np.interp(x, vx, f1d)  # :376
    dfe = np.interp(x, vx, df)  # :377
    kl = np.reshape(klde, -1)
    chiEI = np.pi / kl**2 * dfe  # :381
    rat = np.empty_like(x)
    for s in range(0, len(x), 256):
        g = vx[None, :] - x[s : s + 256, None]
        rat[s : s + 256] = ratintn(df[None, :], g, vx)
    chiERrat = -1.0 / kl**2 * rat  # :385-386
    return fe_vphi.reshape(shp), chiEI.reshape(shp), chiERrat.reshape(shp)


def form_factor_direct(params, grids, sa_deg, num_grad_points=1, lam_shift=0.0, return_parts=False):
    kin = _kinematics(params, grids, sa_deg, num_grad_points, lam_shift)
    fe = np.asarray(params["electron"]["fe"], dtype=np.float64).reshape(-1)
    vx = np.asarray(params["electron"]["v"], dtype=np.float64).reshape(-1)
    chiI = _chi_ion(kin, grids)
    fe_vphi, chiEI, chiERrat = chi_vals_1d(vx, fe, kin["xie"], kin["klde"] * np.ones_like(kin["xie"]))
    chiE = chiERrat + 1j * chiEI
    out, lams = _assemble(kin, chiE, chiI, fe_vphi, grids)
    if return_parts:
        return out, lams, dict(kin=kin, chiE=chiE, chiI=chiI, fe_vphi=fe_vphi)
    return out, lams


# --------------------------------------------------------------------------------------
# a5/a6: FormFactor.calc_in_2D (form_factor.py:449-587), rotate (:300-324), calc_chi_vals (:349-388)
# PARITY UNPINNED: the reference golden for this path (ThryE-arts2v.npy) is a missing blob and interpax's bicubic
# interp2d(..., extrap=True) is restated from its published algorithm (tensor-product cubic Hermite, node derivatives
# fx, fy, fxy from the same 3-point rule as the 1-D "cubic" method, edge cell evaluated outside the grid).
# --------------------------------------------------------------------------------------
def _hermite_node_weights(xq, x):
    """1-D cubic (interpax "cubic") as explicit weights on 4 nodes: returns idx [n,4], w [n,4] such that
    interp1d_cubic(xq, x, f, extrap=True) == sum_k w[:,k] f[idx[:,k]] (uniform or non-uniform grid)."""
    xq = np.asarray(xq, dtype=np.float64)
    n = len(x)
    i = np.clip(np.searchsorted(x, xq, side="right"), 1, n - 1)
    dx = x[i] - x[i - 1]
    t = (xq - x[i - 1]) / dx
    h00, h01 = 2 * t**3 - 3 * t**2 + 1, -2 * t**3 + 3 * t**2
    h10, h11 = (t**3 - 2 * t**2 + t) * dx, (t**3 - t**2) * dx
    idx = np.stack([i - 2, i - 1, i, i + 1], axis=1)
    w = np.zeros((len(xq), 4))
    w[:, 1] += h00
    w[:, 2] += h01
    # slope at node k = i-1 (weights h10) and k = i (weights h11): mean of adjacent secants, one-sided at the ends
    for col, k, h in ((1, i - 1, h10), (2, i, h11)):
        left, right = k == 0, k == n - 1
        interior = ~(left | right)
        kk = np.clip(k, 1, n - 2)
        sl = 1.0 / (x[kk] - x[kk - 1])
        sr = 1.0 / (x[kk + 1] - x[kk])
        # m_k = 0.5 * ((f_k - f_{k-1}) sl + (f_{k+1} - f_k) sr)
        w[:, col - 1] += np.where(interior, -0.5 * sl * h, 0.0)
        w[:, col] += np.where(interior, 0.5 * (sl - sr) * h, 0.0)
        w[:, col + 1] += np.where(interior, 0.5 * sr * h, 0.0)
        # left end: m_0 = (f_1 - f_0)/(x_1 - x_0)   (k = i-1 = 0 -> columns 1, 2)
        s0 = 1.0 / (x[1] - x[0])
        w[:, col] += np.where(left, -s0 * h, 0.0)
        w[:, col + 1] += np.where(left, s0 * h, 0.0)
        # right end: m_{n-1} = (f_{n-1} - f_{n-2})/(...)   (k = i = n-1 -> columns 1, 2)
        s1 = 1.0 / (x[n - 1] - x[n - 2])
        w[:, col - 1] += np.where(right, -s1 * h, 0.0)
        w[:, col] += np.where(right, s1 * h, 0.0)
    return np.clip(idx, 0, n - 1), w


def interp2d_cubic(xq, yq, x, y, f):
    """interpax.interp2d(xq, yq, x, y, f, method="cubic", extrap=True) (call site form_factor.py:324): f[i, j] = f(x_i, y_j)."""
    ix, wx = _hermite_node_weights(xq, x)
    iy, wy = _hermite_node_weights(yq, y)
    out = np.zeros(len(xq))
    for a in range(4):
        for b in range(4):
            out += wx[:, a] * wy[:, b] * f[ix[:, a], iy[:, b]]
    return out


def rotate(vx, df, angle_deg):
    """FormFactor.rotate (form_factor.py:300-324)."""
    rad = np.deg2rad(-angle_deg)
    c, s = np.cos(rad), np.sin(rad)
    R = np.array([[c, -s], [s, c]])
    _vx, _vy = np.meshgrid(vx, vx)
    coords = np.stack((_vx.flatten(), _vy.flatten()))
    rc = np.einsum("ij, ik->kj", R, coords)
    return interp2d_cubic(rc[:, 0], rc[:, 1], vx, vx, df).reshape((vx.size, vx.size), order="F")


def calc_chi_vals_2d(vx, DF, beta, xie_mag, klde_mag):
    """calc_chi_vals (form_factor.py:349-388) for one pole."""
    dvx = vx[1] - vx[0]
    fe_2D_k = rotate(vx, DF, beta * 180 / np.pi)       # :370
    fe_1D_k = np.sum(fe_2D_k, axis=0) * dvx             # :371
    df = np.gradient(fe_1D_k, dvx)                      # :372
    fe_vphi = np.interp(xie_mag, vx, fe_1D_k)           # :376
    dfe = np.interp(xie_mag, vx, df)                    # :377
    chiEI = np.pi / (klde_mag**2) * dfe                 # :381
    chiERrat = -1.0 / (klde_mag**2) * ratintn(df[None, :], (vx - xie_mag)[None, :], vx)[0]   # :385-386
    return fe_vphi, chiEI, chiERrat, fe_1D_k


def form_factor_2d(params, grids, sa_deg, num_grad_points=1, lam_shift=0.0, ud_ang=0.0, va_ang=0.0, return_parts=False):
    """FormFactor.calc_in_2D (form_factor.py:449-587).  params["electron"]["fe"] is the 2-D table DF[V, V] on vx x vx."""
    G = num_grad_points
    gen, ele = params["general"], params["electron"]
    ne = 1.0e20 * float(ele["ne"]) * np.linspace(1 - float(gen["ne_gradient"]) / 200, 1 + float(gen["ne_gradient"]) / 200, G)
    Te = float(ele["Te"]) * np.linspace(1 - float(gen["Te_gradient"]) / 200, 1 + float(gen["Te_gradient"]) / 200, G)
    lam = float(gen["lam"]) + lam_shift
    A, Z, Ti, fract = _ion_lists(params)
    Va = float(gen["Va"]) * 1e6
    ud = float(gen["ud"]) * 1e6
    DF = np.squeeze(np.asarray(ele["fe"], dtype=np.float64))
    vx = np.asarray(ele["v"], dtype=np.float64).reshape(-1)
    Mi = A * MP
    Esq = ME * C**2 * RE
    constants = np.sqrt(4 * np.pi * Esq / ME)
    sarad = (np.asarray(sa_deg, dtype=np.float64) * np.pi / 180).reshape(1, 1, -1)
    Va = (Va * np.cos(va_ang * np.pi / 180), Va * np.sin(va_ang * np.pi / 180))      # :501
    ud = (ud * np.cos(ud_ang * np.pi / 180), ud * np.sin(ud_ang * np.pi / 180))      # :504
    omgL = grids.omgL_num / lam
    omgpe = constants * np.sqrt(ne[:, None, None])
    omgs = grids.omgs
    omg = omgs - omgL
    kLx = np.sqrt(omgL**2 - omgpe**2) / C                                              # :512 (y component 0)
    ks_mag = np.sqrt(omgs**2 - omgpe**2) / C
    kx, ky = np.cos(sarad) * ks_mag - kLx, np.sin(sarad) * ks_mag + 0.0 * kLx          # :513-515
    k_mag = np.sqrt(kx * kx + ky * ky)
    omgdop = omg - (kx * Va[0] + ky * Va[1])                                           # :519
    vTe = np.sqrt(Te[:, None, None] / ME)
    klde_mag = (vTe / omgpe) * k_mag
    Z4, Mi4, fr4 = Z.reshape(1, 1, 1, -1), Mi.reshape(1, 1, 1, -1), fract.reshape(1, 1, 1, -1)
    Zbar = np.sum(Z4 * fr4)
    ni = fr4 * ne[:, None, None, None] / Zbar
    omgpi = constants * Z4 * np.sqrt(ni * ME / Mi4)
    vTi = np.sqrt(Ti / Mi4)
    kldi = (vTi / omgpi) * k_mag[..., None]
    xii = 1.0 / (np.sqrt(2.0) * vTi) * ((omgdop / k_mag)[..., None])
    kin = dict(ne=ne, Te=Te, omgL=omgL, omgpe=omgpe, omgs=omgs, k=k_mag, omgdop=omgdop, vTe=vTe, klde=klde_mag,
               Z=Z4, fract=fr4, Zbar=Zbar, vTi=vTi, kldi=kldi, xii=xii)
    chiI = _chi_ion(kin, grids)
    xiex = ((omgdop / k_mag**2) * kx - ud[0]) / vTe                                    # :552
    xiey = ((omgdop / k_mag**2) * ky - ud[1]) / vTe
    xie_mag = np.sqrt(xiex**2 + xiey**2)
    with np.errstate(divide="ignore", invalid="ignore"):
        beta = np.arctan(xiey / xiex) + np.pi * (-np.heaviside(xiex, 1) + 1)           # :558
    fe_vphi, chiEI, chiER = np.empty(beta.shape), np.empty(beta.shape), np.empty(beta.shape)
    for idx in np.ndindex(beta.shape):
        fe_vphi[idx], chiEI[idx], chiER[idx], _ = calc_chi_vals_2d(vx, DF, beta[idx], xie_mag[idx], klde_mag[idx])
    chiE = chiER + 1j * chiEI
    out, lams = _assemble(kin, chiE, chiI, fe_vphi, grids)
    if return_parts:
        return out, lams, dict(kin=kin, chiE=chiE, chiI=chiI, fe_vphi=fe_vphi, beta=beta, xie_mag=xie_mag)
    return out, lams


# --------------------------------------------------------------------------------------
# a7: FitModel (generate_spectra.py:139-220), temporal / 1d spectype
# --------------------------------------------------------------------------------------
def fit_model_electron(params, grids, sa, cfg_other, num_grad_points=1, lam_shift=0.0, mode="table",
                       angular_full=False, ud_ang=0.0, va_ang=0.0):
    """electron_spectrum (generate_spectra.py:171-220).  ``sa`` = {"sa": deg[A], "weights": ...}.
    cfg_other needs: iawoff, iawfilter, lamrangE.  A 2-D params["electron"]["fe"] takes calc_in_2D (:185-188)."""
    if np.ndim(params["electron"]["fe"]) == 2:
        ThryE, lamAxisE = form_factor_2d(params, grids, sa["sa"], num_grad_points, lam_shift, ud_ang, va_ang)
    else:
        ff = form_factor_1v if mode == "table" else form_factor_direct
        ThryE, lamAxisE = ff(params, grids, sa["sa"], num_grad_points, lam_shift)
    lamAxisE = np.squeeze(lamAxisE) * 1e7  # :191
    ThryE = np.mean(ThryE, axis=0)  # :193  [W,A]
    if angular_full:
        modlE = np.matmul(sa["weights"], ThryE.transpose())  # :194-195 -> [1024, W]
    else:
        modlE = np.sum(ThryE * sa["weights"][0], axis=1)  # :197
    lam = float(params["general"]["lam"])
    if cfg_other.get("iawoff", 0) and (cfg_other["lamrangE"][0] < lam < cfg_other["lamrangE"][1]):
        # :199-208 "set the ion feature to 0".  As written, lamlocb = argmin|lam_axis - lam - 3| is the index nearest lam + 3 and
        # lamlocr the one nearest lam - 3, so on an ascending axis lamlocb > lamlocr and jnp.zeros(lamlocr - lamlocb) has a
        # negative size: the branch cannot run (every checked-in deck has iawoff: 0).  What it evidently means -- the model
        # zeroed between the samples nearest lam - 3 nm and lam + 3 nm -- is what is restated here.
        i_lo = int(np.argmin(np.abs(lamAxisE - (lam - 3.0))))
        i_hi = int(np.argmin(np.abs(lamAxisE - (lam + 3.0))))
        modlE = np.array(modlE, dtype=np.float64, copy=True)
        modlE[..., i_lo:i_hi] = 0.0
    iawf = cfg_other.get("iawfilter", [0])
    if iawf[0]:
        filterb = iawf[3] - iawf[2] / 2
        filterr = iawf[3] + iawf[2] / 2
        if cfg_other["lamrangE"][0] < filterr and cfg_other["lamrangE"][1] > filterb:
            indices = (filterb < lamAxisE) & (filterr > lamAxisE)
            modlE = np.where(indices, modlE * 10.0 ** (-iawf[1]), modlE)
    return lamAxisE, modlE


def fit_model_ion(params, grids, sa, num_grad_points=1, mode="table"):
    """ion_spectrum (generate_spectra.py:139-169): lam_shift = 0 for the ion window (:98-106)."""
    ff = form_factor_1v if mode == "table" else form_factor_direct
    ThryI, lamAxisI = ff(params, grids, sa["sa"], num_grad_points, 0.0)
    lamAxisI = np.squeeze(lamAxisI) * 1e7
    ThryI = np.mean(ThryI, axis=0)
    modlI = np.sum(ThryI * sa["weights"][0], axis=1)
    return lamAxisI, modlI


# --------------------------------------------------------------------------------------
# a8: IRF (irf.py:50-132)
# --------------------------------------------------------------------------------------
def _gauss(lam_axis, stddev):
    origin = (np.amax(lam_axis) + np.amin(lam_axis)) / 2.0
    return (1.0 / (stddev * np.sqrt(2.0 * np.pi))) * np.exp(-((lam_axis - origin) ** 2.0) / (2.0 * stddev**2.0))


def add_electron_irf(lamAxisE, modlE, amps, lam, amp1, amp2, stddevE, norm=0):
    """irf.add_electron_IRF (irf.py:90-132)."""
    inst = _gauss(lamAxisE, stddevE)
    ThryE = np.convolve(modlE, inst, "same")  # :114
    ThryE = (np.amax(modlE) / np.amax(ThryE)) * ThryE  # :115
    if norm > 0:  # :117-122
        ThryE = np.where(
            lamAxisE < lam,
            amp1 * (ThryE / np.amax(ThryE[lamAxisE < lam])),
            amp2 * (ThryE / np.amax(ThryE[lamAxisE > lam])),
        )
    ThryE = np.average(ThryE.reshape(1024, -1), axis=1)  # :124
    if norm == 0:  # :125-130
        lamAxisE = np.average(lamAxisE.reshape(1024, -1), axis=1)
        ThryE = amps * ThryE / np.amax(ThryE)
        ThryE = np.where(lamAxisE < lam, amp1 * ThryE, amp2 * ThryE)
    return lamAxisE, ThryE


def add_ion_irf(lamAxisI, modlI, amps, amp3, stddevI, norm=0):
    """irf.add_ion_IRF (irf.py:50-87)."""
    if stddevI:
        inst = _gauss(lamAxisI, stddevI)
        ThryI = np.convolve(modlI, inst, "same")
        ThryI = (np.amax(modlI) / np.amax(ThryI)) * ThryI
        ThryI = np.average(ThryI.reshape(1024, -1), axis=1)
        if norm == 0:
            lamAxisI = np.average(lamAxisI.reshape(1024, -1), axis=1)
            ThryI = amp3 * amps * ThryI / np.amax(ThryI)
    else:
        ThryI = modlI
    return lamAxisI, ThryI


# --------------------------------------------------------------------------------------
# a9: ThomsonScatteringDiagnostic.__call__ for temporal/imaging/1d (thomson_diagnostic.py:109-142)
# --------------------------------------------------------------------------------------
def diagnostic_1d(params_list, cfg, sa, batch, mode="table"):
    """vmapped diagnostic: params_list = one physical-param dict per lineout.
    cfg = merged deck (dict) with cfg["other"]["lamrangE"/"lamrangI"/"npts"] filled in.
    batch = {e_amps[B], i_amps[B], noise_e, noise_i}.  Returns ThryE[B,1024], ThryI, lamE, lamI."""
    oth = cfg["other"]
    G = cfg["parameters"]["general"]["Te_gradient"]["num_grad_points"]
    gE = Grids(oth["lamrangE"], oth["npts"])
    gI = Grids(oth["lamrangI"], oth["npts"])
    ThryE, ThryI, lamE, lamI = [], [], [], []
    for b, p in enumerate(params_list):
        gen = p["general"]
        if oth["extraoptions"]["load_ion_spec"]:
            lI, mI = fit_model_ion(p, gI, sa, G, mode)
            lI, tI = add_ion_irf(lI, mI, np.asarray(batch["i_amps"]).reshape(-1)[b], float(gen["amp3"]),
                                 oth["PhysParams"]["widIRF"]["spect_stddev_ion"], oth["PhysParams"]["norm"])
        else:
            lI, tI = np.zeros(1), 0.0
        if oth["extraoptions"]["load_ele_spec"]:
            lE, mE = fit_model_electron(p, gE, sa, oth, G, cfg["data"]["ele_lam_shift"], mode)
            lE, tE = add_electron_irf(lE, mE, np.asarray(batch["e_amps"]).reshape(-1)[b], float(gen["lam"]),
                                      float(gen["amp1"]), float(gen["amp2"]),
                                      oth["PhysParams"]["widIRF"]["spect_stddev_ele"], oth["PhysParams"]["norm"])
        else:
            lE, tE = np.zeros(1), 0.0
        ThryE.append(tE); ThryI.append(tI); lamE.append(lE); lamI.append(lI)
    ThryE = np.array(ThryE) + np.asarray(batch["noise_e"])  # :139
    ThryI = np.array(ThryI) + np.asarray(batch["noise_i"])  # :140
    return ThryE, ThryI, np.array(lamE), np.array(lamI)


# --------------------------------------------------------------------------------------
# a8/a9 (ARTS): irf.add_ATS_IRF (irf.py:5-47) and reduce_ATS_to_resunit (thomson_diagnostic.py:78-107)
# --------------------------------------------------------------------------------------
def ats_taps(axis, fwhm):
    """Gaussian instrument function sampled on `axis`, centred at mid-axis (irf.py:22-33)."""
    stddev = fwhm / 2.3548
    origin = (np.amax(axis) + np.amin(axis)) / 2.0
    return np.squeeze((1.0 / (stddev * np.sqrt(2.0 * np.pi))) * np.exp(-((axis - origin) ** 2.0) / (2.0 * stddev**2.0)))


def add_ats_irf(lamAxisE, angAxis, modlE, spect_fwhm, ang_fwhm, norm=0):
    """add_ATS_IRF (irf.py:5-47): modlE [NA, W] -> ThryE [NA, W]."""
    inst_lam = ats_taps(lamAxisE, spect_fwhm)
    inst_ang = ats_taps(angAxis, ang_fwhm)
    ThryE = np.array([np.convolve(modlE[:, i], inst_ang, "same") for i in range(modlE.shape[1])])  # :34 -> [W, NA]
    ThryE = np.array([np.convolve(ThryE[:, i], inst_lam, "same") for i in range(ThryE.shape[1])])  # :36 -> [NA, W]
    ThryE = np.amax(modlE, axis=1, keepdims=True) / np.amax(ThryE, axis=1, keepdims=True) * ThryE   # :39
    if norm > 0:
        raise NotImplementedError("PhysParams.norm > 0: not used by any reference deck")
    return lamAxisE, ThryE


def reduce_ats_to_resunit(ThryE, lamAxisE, lam, amp1, amp2, e_amps, n_lam_data, ccd0, row_start, row_end):
    """reduce_ATS_to_resunit (thomson_diagnostic.py:78-107)."""
    lam_step = round(ThryE.shape[1] / n_lam_data)   # :93
    ang_step = round(ThryE.shape[0] / ccd0)          # :94
    ThryE = np.array([np.average(ThryE[:, i:i + lam_step], axis=1) for i in range(0, ThryE.shape[1], lam_step)])
    ThryE = np.array([np.average(ThryE[:, i:i + ang_step], axis=1) for i in range(0, ThryE.shape[1], ang_step)])
    lamAxisE = np.array([np.average(lamAxisE[i:i + lam_step], axis=0) for i in range(0, lamAxisE.shape[0], lam_step)])
    ThryE = ThryE[row_start:row_end, :]
    ThryE = e_amps * ThryE / np.amax(ThryE, axis=1, keepdims=True)
    ThryE = np.where(lamAxisE < lam, amp1 * ThryE, amp2 * ThryE)
    return ThryE, lamAxisE


def diagnostic_arts(params, cfg, sa, batch, mode="table"):
    """ThomsonScatteringDiagnostic.__call__ for spectype "angular_full" (thomson_diagnostic.py:109-142): one parameter
    set, one image.  sa = {"sa": deg[241], "weights": [1024, 241], "angAxis": [1024]}."""
    oth = cfg["other"]
    G = cfg["parameters"]["general"]["Te_gradient"]["num_grad_points"]
    gE = Grids(oth["lamrangE"], oth["npts"])
    gen_cfg = cfg["parameters"]["general"]
    lamE, modlE = fit_model_electron(params, gE, sa, oth, G, cfg["data"]["ele_lam_shift"], mode, angular_full=True,
                                     ud_ang=gen_cfg["ud"].get("angle", 0.0), va_ang=gen_cfg["Va"].get("angle", 0.0))
    wid = oth["PhysParams"]["widIRF"]
    lamE, ThryE = add_ats_irf(lamE, sa["angAxis"], modlE, wid["spect_FWHM_ele"], wid["ang_FWHM_ele"], oth["PhysParams"]["norm"])
    gen = params["general"]
    ThryE, lamE = reduce_ats_to_resunit(ThryE, lamE, float(gen["lam"]), float(gen["amp1"]), float(gen["amp2"]),
                                        np.asarray(batch["e_amps"]), np.asarray(batch["e_data"]).shape[1], oth["CCDsize"][0],
                                        cfg["data"]["lineouts"]["start"], cfg["data"]["lineouts"]["end"])
    return ThryE + np.asarray(batch["noise_e"]), lamE, modlE


def extract_lineouts(image, pixels, dpixel, gain, window=None):
    """The lineout extraction of get_lineouts (tsadar/utils/process/lineouts.py:85-93, 125-152) for one spectrometer image
    [NY, NX]: column sums over [a - dpixel, a + dpixel), boxcar smoothing with span 2 dpixel + 1, / gain, amplitude = max over
    the fit-window mask.  -> (data [L, NY], amps [L])."""
    image = np.asarray(image, dtype=np.float64)
    span = 2 * dpixel + 1
    raw = [np.sum(image[:, a - dpixel: a + dpixel], axis=1) for a in pixels]
    smooth = [np.convolve(r, np.ones(span) / span, "same") for r in raw]
    data = np.array([s / gain for s in smooth])
    m = np.ones(image.shape[0], dtype=bool) if window is None else np.asarray(window, dtype=bool)
    return data, np.amax(data[:, m], axis=1)


def rotate_pixels(A, theta):
    """vector_tools.rotate (tsadar/utils/vector_tools.py:94-138): bilinear rotation of a table about its centre on the pixel
    grid, as the multiplexed-shot loss applies to f(vx, vy) (loss_function.py:291-293).  `jnp.asarray(., dtype=int)` truncates
    toward zero, `% 1` is the floor-based fraction, the clip upper bound n is clamped to n - 1 by the gather."""
    A = np.asarray(A, dtype=np.float64)
    n0, n1 = A.shape
    rp = [n0 / 2, n1 / 2]
    x, y = np.meshgrid(np.arange(n0), np.arange(n1))
    R = np.array([[np.cos(-theta), -np.sin(-theta)], [np.sin(-theta), np.cos(-theta)]])
    o = R @ np.array([x.flatten() - rp[0], y.flatten() - rp[1]])
    or_x, or_y = o[0].reshape(x.shape), o[1].reshape(y.shape)
    w11 = (1 - or_x % 1) * (1 - or_y % 1)
    w12 = (1 - or_x % 1) * (or_y % 1)
    w21 = (or_x % 1) * (1 - or_y % 1)
    w22 = (or_x % 1) * (or_y % 1)

    def q(dy, dx):
        r = np.minimum(np.clip(np.trunc(or_y + rp[1] + dy).astype(int), 0, n1), A.shape[0] - 1)
        c = np.minimum(np.clip(np.trunc(or_x + rp[0] + dx).astype(int), 0, n0), A.shape[1] - 1)
        return A[r, c]

    return w11 * q(0, 0) + w12 * q(1, 0) + w21 * q(0, 1) + w22 * q(1, 1)


# --------------------------------------------------------------------------------------
# a10: loss (loss_function.py:190-267, 269-342, 364-373, 386-418)
# --------------------------------------------------------------------------------------
def loss_functionals(d, t, uncert, method="l2"):
    if method == "l1":
        return np.abs(d - t) / uncert
    if method == "l2":
        return np.square(d - t) / uncert
    if method == "log-cosh":
        return np.log(np.cosh(d - t))
    if method == "poisson":
        return t - d * np.log(t)
    raise NotImplementedError(method)


def calc_ei_error(cfg, batch, ThryI, lamAxisI, ThryE, lamAxisE, uncert, reduce_func=np.nanmean):
    """loss_function.py:190-267."""
    i_error, e_error = 0.0, 0.0
    fr = cfg["data"]["fit_rng"]
    ex = cfg["other"]["extraoptions"]
    method = cfg["optimizer"]["loss_method"]
    if ex["fit_IAW"]:
        err = loss_functionals(batch["i_data"], ThryI, uncert[0], method)
        m = ((lamAxisI > fr["iaw_min"]) & (lamAxisI < fr["iaw_cf_min"])) | ((lamAxisI > fr["iaw_cf_max"]) & (lamAxisI < fr["iaw_max"]))
        i_error += reduce_func(np.where(m, err, np.nan))
    if ex["fit_EPWb"]:
        err = loss_functionals(batch["e_data"], ThryE, uncert[1], method)
        m = (lamAxisE > fr["blue_min"]) & (lamAxisE < fr["blue_max"])
        e_error += reduce_func(np.where(m, err, np.nan))
    if ex["fit_EPWr"]:
        err = loss_functionals(batch["e_data"], ThryE, uncert[1], method)
        m = (lamAxisE > fr["red_min"]) & (lamAxisE < fr["red_max"])
        e_error += reduce_func(np.where(m, err, np.nan))
        if ex["fit_EPWb"]:
            e_error *= 1.0 / 2.0
    return i_error, e_error


def loss_1d(params_list, cfg, sa, batch, i_norm=1.0, e_norm=1.0, mode="table"):
    """LossFunction.__loss__ (loss_function.py:364-373) -> calc_loss (:269-342), non-multiplexed."""
    ThryE, ThryI, lamE, lamI = diagnostic_1d(params_list, cfg, sa, batch, mode)
    i_err, e_err = calc_ei_error(cfg, batch, ThryI, lamI, ThryE, lamE, [np.square(i_norm), np.square(e_norm)])
    return cfg["data"]["ion_loss_scale"] * i_err + e_err, ThryE, ThryI
