"""The reference's physics known-answer tests run on the CUDA path, written like the originals
(tests/test_form_factor/test_epw.py:17-74 and test_iaw.py:14-71): same deck (epw_defaults + epw_inputs), same
FormFactor / ThomsonParams calls, same peak picking, same assertions."""
from copy import deepcopy

import numpy as np
import pytest
from numpy.testing import assert_allclose
from scipy.signal import find_peaks

from tests.common import load_cfg

pytestmark = pytest.mark.gpu


def test_epw():
    """Bohm-Gross: the electron-plasma-wave resonances of the computed spectrum sit on omega^2 = omega_pe^2 + 3 k^2 vTe^2."""
    from tsadar_b200.form_factor import FormFactor
    from tsadar_b200.ts_params import ThomsonParams
    config = load_cfg("cfg_epw")
    npts = 8192
    ts_params = ThomsonParams(config["parameters"], num_params=1, batch=False)
    electron_form_factor = FormFactor(
        [400, 700],
        npts=npts,
        lam_shift=config["data"]["ele_lam_shift"],
        scattering_angles={"sa": np.array([60])},
        num_grad_points=config["parameters"]["general"]["ne_gradient"]["num_grad_points"],
        ud_ang=None,
        va_ang=None,
    )
    sa = np.array([60])
    physical_params = ts_params()
    ThryE, lamAxisE = electron_form_factor(physical_params)
    ThryE, lamAxisE = ThryE.cpu().numpy(), lamAxisE.cpu().numpy()
    ThryE = np.squeeze(ThryE)
    test = deepcopy(np.asarray(ThryE))
    peaks, peak_props = find_peaks(test, height=(0.01, 0.5), prominence=0.02)
    highest_peak_index = peaks[np.argmax(peak_props["peak_heights"])]
    second_highest_peak_index = peaks[np.argsort(peak_props["peak_heights"])[0]]

    C = 2.99792458e10
    Me = 510.9896 / C**2  # electron mass keV/C^2
    re = 2.8179e-13  # classical electron radius cm
    Esq = Me * C**2 * re  # sq of the electron charge keV cm
    constants = np.sqrt(4 * np.pi * Esq / Me)

    lams = lamAxisE[0, [highest_peak_index, second_highest_peak_index], 0]
    model_omegas = 2 * np.pi * C / lams  # peak frequencies
    omgpe = constants * np.sqrt(0.2 * 1e20)
    omgL = 2 * np.pi * 1e7 * C / config["parameters"]["general"]["lam"]["val"]  # laser frequency Rad / s
    ks = np.sqrt(model_omegas**2 - omgpe**2) / C
    kL = np.sqrt(omgL**2 - omgpe**2) / C
    k = np.sqrt(ks**2 + kL**2 - 2 * ks * kL * np.cos(sa * np.pi / 180))
    vTe = np.sqrt(0.5 / Me)
    omg = np.sqrt(omgpe**2 + 3 * k**2 * vTe**2)
    theory_omegas = [omgL + omg[0], omgL - omg[1]]
    assert_allclose(model_omegas, theory_omegas, rtol=1e-2)


def test_iaw():
    """Ion-acoustic resonances at omega_L +- 2 kL sqrt((Te + 3 Ti) / Mp)."""
    from tsadar_b200.form_factor import FormFactor
    from tsadar_b200.ts_params import ThomsonParams
    config = load_cfg("cfg_epw")
    C = 2.99792458e10
    Me = 510.9896 / C**2  # electron mass keV/C^2
    Mp = Me * 1836.1  # proton mass keV/C^2
    re = 2.8179e-13  # classical electron radius cm
    Esq = Me * C**2 * re  # sq of the electron charge keV cm
    ion_form_factor = FormFactor(
        [525, 528],
        npts=8192,
        lam_shift=0.0,
        scattering_angles={"sa": np.array([60])},
        num_grad_points=config["parameters"]["general"]["ne_gradient"]["num_grad_points"],
        ud_ang=None,
        va_ang=None,
    )
    constants = np.sqrt(4 * np.pi * Esq / Me)
    ts_params = ThomsonParams(config["parameters"], num_params=1, batch=False)
    physical_params = ts_params()
    ThryI, lamAxisI = ion_form_factor(physical_params)
    ThryI, lamAxisI = ThryI.cpu().numpy(), lamAxisI.cpu().numpy()
    ThryI = np.mean(ThryI, axis=0)
    ThryI = np.squeeze(ThryI)
    test = deepcopy(np.asarray(ThryI))
    peaks, peak_props = find_peaks(test, height=0.1, prominence=0.2)
    highest_peak_index = peaks[np.argmax(peak_props["peak_heights"])]
    second_highest_peak_index = peaks[np.argpartition(peak_props["peak_heights"], -2)[-2]]

    lams = lamAxisI[0, [highest_peak_index, second_highest_peak_index], 0]
    omgpe = constants * np.sqrt(0.2 * 1e20)
    omgL = 2 * np.pi * 1e7 * C / config["parameters"]["general"]["lam"]["val"]  # laser frequency Rad / s
    kL = np.sqrt(omgL**2 - omgpe**2) / C
    model_omegas = 2 * np.pi * C / lams  # peak frequencies
    omg = 2 * kL * np.sqrt((0.5 + 3 * 0.2) / Mp)
    theory_omegas = [omgL + omg, omgL - omg]
    assert_allclose(np.sort(theory_omegas), np.sort(model_omegas), rtol=1e-2)
