"""CPU checks of the oracle's 2-D pieces (rows a5/a6: rotate / calc_chi_vals), which have no golden vector of their own:
the bicubic restatement (oracle.np_oracle.interp2d_cubic) is tied here to an INDEPENDENT construction of the same published
scheme -- scipy.interpolate.CubicHermiteSpline with mean-of-secants node slopes, applied as a tensor product -- and to the
golden-pinned 1-D routine through separability; `rotate` is checked through the rotations whose answer is known in closed
form (0 deg, 90 deg, isotropic tables)."""
import numpy as np
import pytest
from scipy.interpolate import CubicHermiteSpline

from oracle import np_oracle as O


def _slopes(x, f):
    s = np.diff(f) / np.diff(x)
    return np.concatenate([s[:1], 0.5 * (s[:-1] + s[1:]), s[-1:]])


def _tensor_hermite(xq, yq, x, y, f):
    """interp along y for every x-row with scipy's Hermite spline, then along x, one query at a time (extrapolating)."""
    out = np.zeros(len(xq))
    rows = np.stack([CubicHermiteSpline(y, f[i], _slopes(y, f[i]), extrapolate=True)(yq) for i in range(len(x))])  # [nx, nq]
    for q in range(len(xq)):
        col = rows[:, q]
        out[q] = CubicHermiteSpline(x, col, _slopes(x, col), extrapolate=True)(xq[q])
    return out


@pytest.mark.parametrize("uniform", [True, False])
def test_interp1d_matches_scipy_hermite(uniform):
    rng = np.random.default_rng(3)
    x = np.linspace(-3, 3, 41) if uniform else np.sort(rng.uniform(-3, 3, 41))
    f = np.exp(-x**2 / 2) * (1 + 0.3 * np.sin(3 * x))
    xq = rng.uniform(x[0], x[-1], 500)
    ref = CubicHermiteSpline(x, f, _slopes(x, f))(xq)
    got = O.interp1d_cubic(xq, x, f, extrap=(0.0, 0.0))
    assert np.max(np.abs(got - ref)) < 1e-14
    # outside the grid: the fills, not the polynomial
    out = O.interp1d_cubic(np.array([x[0] - 1.0, x[-1] + 1.0]), x, f, extrap=(7.0, -7.0))
    assert out.tolist() == [7.0, -7.0]


@pytest.mark.parametrize("uniform", [True, False])
def test_interp2d_matches_tensor_product_of_scipy_hermite(uniform):
    rng = np.random.default_rng(5)
    n = 33
    x = np.linspace(-4, 4, n) if uniform else np.sort(rng.uniform(-4, 4, n))
    y = np.linspace(-4, 4, n) if uniform else np.sort(rng.uniform(-4, 4, n))
    X, Y = np.meshgrid(x, y, indexing="ij")
    f = np.exp(-(X**2 + 0.5 * Y**2 + 0.4 * X * Y) / 2) * (1 + 0.2 * X)
    # inside and OUTSIDE the grid (extrap=True: the edge cell's polynomial continues)
    xq = rng.uniform(x[0] - 0.7, x[-1] + 0.7, 400)
    yq = rng.uniform(y[0] - 0.7, y[-1] + 0.7, 400)
    got = O.interp2d_cubic(xq, yq, x, y, f)
    ref = _tensor_hermite(xq, yq, x, y, f)
    assert np.max(np.abs(got - ref)) < 1e-13 * max(1.0, np.max(np.abs(ref)))


def test_interp2d_separable_ties_to_golden_pinned_1d():
    """f(x, y) = g(x) h(y)  =>  bicubic(f)(xq, yq) = cubic(g)(xq) * cubic(h)(yq): the interpolant is a tensor product and
    interp1d_cubic is the routine the reference's golden spectra pin (tests/test_oracle_twin.py)."""
    rng = np.random.default_rng(7)
    x = np.linspace(-5, 5, 64)
    g, h = np.exp(-x**2 / 2), 1.0 / (1.0 + x**2)
    xq, yq = rng.uniform(-5, 5, 300), rng.uniform(-5, 5, 300)
    got = O.interp2d_cubic(xq, yq, x, x, np.outer(g, h))
    ref = O.interp1d_cubic(xq, x, g, (0.0, 0.0)) * O.interp1d_cubic(yq, x, h, (0.0, 0.0))
    assert np.max(np.abs(got - ref)) < 1e-14


def test_interp2d_exact_on_nodes_and_quadratics():
    x = np.linspace(-2, 2, 21)
    X, Y = np.meshgrid(x, x, indexing="ij")
    f = (1 + X + 0.5 * X**2) * (2 - Y + 0.25 * Y**2)
    # nodes are reproduced exactly
    got = O.interp2d_cubic(X.ravel(), Y.ravel(), x, x, f)
    assert np.max(np.abs(got - f.ravel())) < 1e-13
    # central secant slopes are exact for quadratics on a uniform grid: the interior interpolant reproduces them
    rng = np.random.default_rng(11)
    xq, yq = rng.uniform(x[1], x[-2], 300), rng.uniform(x[1], x[-2], 300)
    exact = (1 + xq + 0.5 * xq**2) * (2 - yq + 0.25 * yq**2)
    assert np.max(np.abs(O.interp2d_cubic(xq, yq, x, x, f) - exact)) < 1e-12


def test_rotate_identity_quarter_turn_and_isotropic():
    vx = np.linspace(-6, 6, 65)[:-1] + 6.0 / 64          # the reference's cell-centred grid (symmetric about 0)
    X, Y = np.meshgrid(vx, vx, indexing="ij")
    f = np.exp(-(X**2 / 1.0 + Y**2 / 2.5) / 2) * (1 + 0.3 * np.tanh(X))
    # 0 deg: every query is a node
    r0 = O.rotate(vx, f, 0.0)
    assert np.max(np.abs(r0 - f)) < 1e-14
    # 90 deg maps the symmetric grid onto itself, so the answer is exact at nodes.  By hand from form_factor.py:300-324:
    # rad = -pi/2 -> R = [[0, 1], [-1, 0]]; meshgrid('xy') point k = a V + b is (vx[b], vx[a]); einsum "ij,ik->kj" gives the
    # query (-vx[a], vx[b]) = node [V-1-a, b]; the order="F" reshape puts k at [p, q] = [b, a]  =>  out[p, q] = f[V-1-q, p]
    r90 = O.rotate(vx, f, 90.0)
    assert np.max(np.abs(r90 - f.T[:, ::-1])) < 1e-12
    assert np.max(np.abs(O.rotate(vx, f, -90.0) - f.T[::-1, :])) < 1e-12
    # four quarter turns compose to the identity
    r = f
    for _ in range(4):
        r = O.rotate(vx, r, 90.0)
    assert np.max(np.abs(r - f)) < 1e-12
    # an isotropic table is unchanged by any rotation up to the interpolation error, which is third order in dv for
    # this scheme (second-order node slopes): halving dv must cut it by ~8
    def iso_err(V, ang):
        v = np.linspace(-6, 6, V + 1)[:-1] + 6.0 / V
        A, B = np.meshgrid(v, v, indexing="ij")
        iso = np.exp(-(A**2 + B**2) / 2)
        inner = (np.abs(A) < 4.0) & (np.abs(B) < 4.0)     # corners rotate out of the box and extrapolate
        return np.max(np.abs(O.rotate(v, iso, ang) - iso)[inner])

    for ang in (17.0, 45.0, -123.0):
        e64, e128 = iso_err(64, ang), iso_err(128, ang)
        assert e64 < 5e-4 and e128 < e64 / 5.0, (ang, e64, e128)


def test_calc_chi_vals_isotropic_equals_1d_projection():
    """For an isotropic Maxwellian table the rotated projection is the 1-D Maxwellian for every beta, so calc_chi_vals_2d
    must agree with the 1-D susceptibility pieces (lerp of f and df, ratintn of df) evaluated on that projection."""
    V = 128
    vx = np.linspace(-6, 6, V + 1)[:-1] + 6.0 / V
    dv = vx[1] - vx[0]
    X, Y = np.meshgrid(vx, vx, indexing="ij")
    DF = np.exp(-(X**2 + Y**2) / 2) / (2 * np.pi)
    f1 = np.sum(DF, axis=0) * dv
    df1 = np.gradient(f1, dv)
    for beta, xi, klde in [(0.3, 0.7, 0.31), (-1.1, -2.2, 0.5), (2.0, 3.9, 0.27)]:
        fe_vphi, chiEI, chiERrat, proj = O.calc_chi_vals_2d(vx, DF, beta, xi, klde)
        assert np.max(np.abs(proj - f1)) < 5e-5
        assert abs(fe_vphi - np.interp(xi, vx, f1)) < 5e-5
        assert abs(chiEI - np.pi / klde**2 * np.interp(xi, vx, df1)) < 2e-3 * np.pi / klde**2
        ref = -1.0 / klde**2 * O.ratintn(df1[None, :], (vx - xi)[None, :], vx)[0]
        assert abs(chiERrat - ref) < 2e-3 / klde**2
