"""SURVEY.md 8f rows N1 / N3 as kernels: tsff_params_fwd / _bwd (ThomsonParams.__call__ + DLM1V.__call__, ts_params.py:583-603,
distribution_functions/base.py:277-294) and tsff_adam_step (optax.adam, inverse/loops.py:87-89) against the torch-eager mirror
(tsadar_b200.ts_params.ThomsonParams, itself checked against oracle/params_oracle.py on the CPU, tests/test_params_producers.py)."""
import copy

import numpy as np
import pytest
import torch

from tests.common import SA_P9, load_cfg

pytestmark = pytest.mark.gpu


def _pair(par, B, fe_dtype=torch.float64):
    from tsadar_b200.ts_params import ThomsonParams, FusedThomsonParams
    tp = ThomsonParams(copy.deepcopy(par), num_params=B, batch=True, activate=True)
    rng = np.random.default_rng(3)
    with torch.no_grad():                                     # make the lineouts differ
        for s in tp.leaves.values():
            if s.active:
                s.value += torch.tensor(rng.normal(size=B) * 0.4, device=s.value.device)
    fz = FusedThomsonParams(copy.deepcopy(par), num_params=B, batch=True, activate=True, fe_dtype=fe_dtype)
    with torch.no_grad():
        for k, name in enumerate(fz.active_names):
            fz.x[:, k] = tp.leaves[name].value
    return tp, fz


def _two_ion_deck():
    par = load_cfg("cfg_1d")["parameters"]
    par["ion-1"]["fract"]["val"] = 0.6
    par["ion-2"] = copy.deepcopy(par["ion-1"])
    par["ion-2"]["A"]["val"], par["ion-2"]["Z"]["val"], par["ion-2"]["fract"]["val"] = 1.0, 1.0, 0.9
    par["ion-2"]["Ti"]["val"] = 0.35
    par["ion-2"]["Ti"]["same"] = True
    par["ion-1"]["Ti"]["active"] = True
    par["ion-1"]["Z"]["active"] = True
    par["general"]["amp3"]["active"] = True
    return par


@pytest.mark.parametrize("deck", ["1d", "two_ions"])
def test_params_kernels_match_the_eager_mirror(deck):
    from tsadar_b200.form_factor import pack_params
    par = load_cfg("cfg_1d")["parameters"] if deck == "1d" else _two_ion_deck()
    B = 5
    tp, fz = _pair(par, B)
    block_ref, fe_ref, _, _, nI = pack_params(tp(), tp.device)
    out = fz()
    block, fe = out["_packed"][0], out["_packed"][1]
    assert block.shape == block_ref.shape and fe.shape == fe_ref.shape
    assert float((block - block_ref).abs().max()) <= 1e-13 * float(block_ref.abs().max())
    assert float((fe - fe_ref).abs().max()) <= 1e-13 * float(fe_ref.abs().max())
    assert torch.equal(out["electron"]["Te"], block[:, 0]) and torch.equal(out["ion-1"]["fract"], block[:, 13])
    # VJP: random cotangents on both outputs
    g = torch.Generator(device="cuda").manual_seed(1)
    cb = torch.randn(block.shape, dtype=torch.float64, device="cuda", generator=g)
    cf = torch.randn(fe.shape, dtype=torch.float64, device="cuda", generator=g)
    ((block * cb).sum() + (fe * cf).sum()).backward()
    ((block_ref * cb).sum() + (fe_ref * cf).sum()).backward()
    assert len(fz.active_names) == fz.x.shape[1] == len(tp.parameters())
    for k, name in enumerate(fz.active_names):
        ref = tp.leaves[name].value.grad
        got = fz.x.grad[:, k]
        assert float((got - ref).abs().max()) <= 1e-11 * max(float(ref.abs().max()), 1e-300), (name, got, ref)


def test_params_kernel_float32_tables_and_m_clamping():
    """fe as float32 (the benchmark's table type); an inactive m beyond the table's m axis takes the edge table, as jnp.interp
    does (base.py:292)."""
    from tsadar_b200.ts_params import FusedThomsonParams
    par = load_cfg("cfg_1d")["parameters"]
    tp, fz = _pair(par, 4, fe_dtype=torch.float32)
    _, fe = fz.physical()
    fe64 = tp()["electron"]["fe"]
    assert fe.dtype == torch.float32
    assert float((fe.double() - fe64).abs().max()) <= 1e-7 * float(fe64.abs().max())
    par2 = copy.deepcopy(par)
    par2["electron"]["fe"]["active"] = False
    par2["electron"]["fe"]["params"]["m"]["val"] = 5.5
    fz2 = FusedThomsonParams(par2, num_params=2, batch=True, activate=True)
    _, fe2 = fz2.physical()
    par3 = copy.deepcopy(par2)
    par3["electron"]["fe"]["params"]["m"]["val"] = 5.0
    _, fe3 = FusedThomsonParams(par3, num_params=2, batch=True, activate=True).physical()
    assert float((fe2 - fe3).abs().max()) <= 1e-12 * float(fe3.abs().max())      # (5 - 2) / 0.1 rounds just below 30
    assert abs(float(fe2[0].sum()) * fz2.dv - 1.0) < 1e-12


def test_adam_kernel_matches_optax_semantics():
    """tsff_adam_step against the textbook update (optax.adam defaults), five steps, every lineout its own counter."""
    from tsadar_b200 import _ffi
    B, n = 37, 6
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn((B, n), dtype=torch.float64, device="cuda", generator=g)
    mu, nu, cnt = torch.zeros_like(x), torch.zeros_like(x), torch.zeros(B, dtype=torch.float64, device="cuda")
    xr, mr, vr = x.clone(), torch.zeros_like(x), torch.zeros_like(x)
    lr, b1, b2, eps = 0.01, 0.9, 0.999, 1e-8
    for t in range(1, 6):
        grad = torch.randn((B, n), dtype=torch.float64, device="cuda", generator=g)
        _ffi.check(_ffi.lib().tsff_adam_step(B, n, x.data_ptr(), grad.data_ptr(), mu.data_ptr(), nu.data_ptr(), cnt.data_ptr(), lr, b1, b2, eps,
                                             torch.cuda.current_stream().cuda_stream))
        mr = b1 * mr + (1 - b1) * grad
        vr = b2 * vr + (1 - b2) * grad * grad
        xr = xr - lr * (mr / (1 - b1**t)) / (torch.sqrt(vr / (1 - b2**t)) + eps)
        assert float((x - xr).abs().max()) <= 1e-14 * float(xr.abs().max())
    assert torch.equal(cnt, torch.full_like(cnt, 5.0))


@pytest.mark.parametrize("graph", [False, True])
def test_fused_fit_step_equals_the_eager_fit(graph):
    """fit.fused_adam_fit (params kernel -> form factor -> IRF -> loss -> adjoints -> params VJP kernel -> adam kernel, one CUDA
    graph) against fit.adam_fit on the torch-eager ThomsonParams: same loss history, same final parameters."""
    from tsadar_b200.fit import adam_fit, fused_adam_fit, batch_to_device
    from tsadar_b200.loss_function import LossFunction
    cfg = load_cfg("cfg_1d")
    cfg["other"]["points_per_pixel"] = 1
    cfg["other"]["npts"] = 1024
    cfg["parameters"]["electron"]["fe"]["nvx"] = 64
    B = 2
    lamb = np.linspace(400, 700, 1024)
    e_data = 0.6 * np.exp(-0.5 * ((lamb - 470) / 12.0) ** 2) + 0.5 * np.exp(-0.5 * ((lamb - 590) / 15.0) ** 2) + 0.01
    batch = dict(e_data=np.stack([e_data, 1.2 * e_data]), i_data=np.ones((B, 1024)), e_amps=np.array([1.0, 1.2]),
                 i_amps=np.ones(B), noise_e=np.zeros((B, 1024)), noise_i=np.zeros((B, 1024)))
    bt = batch_to_device(batch)
    lf = LossFunction(cfg, SA_P9, batch)
    tp, fz = _pair(cfg["parameters"], B)
    closure = lambda p: lf.calc_loss(p, bt)[0]
    h_ref = adam_fit(closure, tp, 0.01, 6, cuda_graph=False)
    h = fused_adam_fit(closure, fz, 0.01, 6, cuda_graph=graph)
    np.testing.assert_allclose(h, h_ref, rtol=1e-8)
    for k, name in enumerate(fz.active_names):
        # the adjoint's scatter-adds are atomics (their order varies run to run) and feed FP32 sweeps: gradients repeat to ~1e-7
        # relative, so six adam steps of 0.01 repeat to ~1e-8 absolute
        assert float((fz.x[:, k] - tp.leaves[name].value).abs().max()) <= 2e-7
