"""lineout extraction (tsadar/utils/process/lineouts.py:85-165) through tsff_lineouts_fwd against the NumPy
restatement (oracle.extract_lineouts): synthetic CCD counts, several lineouts, the electron fit windows of the 1d deck."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from tests.common import load_cfg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("NY,NX,dpixel", [(1024, 1024, 5), (1024, 700, 0), (517, 300, 12)])
def test_lineout_extraction_matches_oracle(NY, NX, dpixel):
    from tsadar_b200.lineouts import extract_lineouts
    rng = np.random.default_rng(NY + dpixel)
    image = rng.poisson(40.0, size=(NY, NX)).astype(np.float64) + 300.0 * np.exp(-0.5 * ((np.arange(NY)[:, None] - NY / 3) / 20.0) ** 2)
    pixels = [dpixel, NX // 3, NX // 2 + 7, NX - dpixel]          # both image edges included
    cfg = load_cfg("cfg_1d")
    fr = cfg["data"]["fit_rng"]
    axis_y = np.linspace(319.7, 739.6, NY)
    windows = [(fr["blue_min"], fr["blue_max"]), (fr["red_min"], fr["red_max"])]
    mask = ((fr["blue_min"] < axis_y) & (axis_y < fr["blue_max"])) | ((fr["red_min"] < axis_y) & (axis_y < fr["red_max"]))
    gain = 4.2
    data, amps = extract_lineouts(torch.tensor(image, device="cuda"), pixels, dpixel, gain, axis_y, windows)
    ref_d, ref_a = O.extract_lineouts(image, pixels, dpixel, gain, mask)
    np.testing.assert_allclose(data.cpu().numpy(), ref_d, rtol=1e-13, atol=1e-12)
    np.testing.assert_allclose(amps.cpu().numpy(), ref_a, rtol=1e-13)
    d2, a2 = extract_lineouts(torch.tensor(image, device="cuda"), pixels, dpixel, gain)
    np.testing.assert_allclose(a2.cpu().numpy(), ref_d.max(axis=1), rtol=1e-13)


def test_lineout_outside_the_image_is_refused():
    from tsadar_b200.lineouts import extract_lineouts
    img = torch.zeros((64, 64), dtype=torch.float64, device="cuda")
    with pytest.raises(RuntimeError, match="leaves the image"):
        extract_lineouts(img, [2], 5, 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        extract_lineouts(torch.zeros((4, 4), dtype=torch.float64), [2], 1, 1.0)
