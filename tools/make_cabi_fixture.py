#!/usr/bin/env python
"""Writes tests/golden/cabi_fixture.bin: inputs and ORACLE outputs of one small direct-mode step
(tsff_ctx_create -> tsff_ff_fwd -> tsff_loss_fwd_bwd -> tsff_ff_bwd) for the torch-free C client tests/cabi_client/client.c.
Runs on the CPU (oracle/np_oracle.py for the spectrum, oracle/torch_oracle.py autograd for the gradients of the l2 loss).

Layout (little endian): char magic[8] = "TSFFFIX1"; int32 B, W, V, NP, zp_n, pad[3]; double lam_min, lam_max, v0, dv, sa_deg,
weight, uncert, scale; then zp_x, zp_re, zp_im [zp_n] f64; params [B][NP] f64; fe [B][V] f32; target [B][W] f64; wq [W] f64;
expected modl [B][W] f64; loss f64; params_bar [B][NP] f64; fe_bar [B][V] f64."""
import os, sys, struct
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from oracle import np_oracle as O, torch_oracle as TO
from tests.common import row_to_params
from tsadar_b200.synthetic import make_lineouts, SA_SYN, LAM_RANGE

B, W, V = 3, 256, 512
params, fe, vx, _ = make_lineouts(B, seed=7, nvx=V)            # float32 tables
NP = params.shape[1]
grids = O.Grids(list(LAM_RANGE), W)
pert = params.copy(); pert[:, 0] *= 1.05; pert[:, 1] *= 0.95
target = np.stack([O.form_factor_direct(row_to_params(pert[b], fe[b], vx, 1), grids, SA_SYN, 1, 0.0)[0][0, :, 0] for b in range(B)])
wq = np.full(W, 1.0 / W)
wq[:16] = 0.0                                                   # a fit window: the first points carry no weight
wq /= wq.sum()
uncert, scale = 1.3, 1.0 / B
modl, pbar, fbar, loss = [], [], [], 0.0
for b in range(B):
    leaves, p = TO.params_from_block(params[b], 1)
    fet = torch.tensor(fe[b].astype(np.float64), requires_grad=True)
    ff = TO.form_factor_direct(p, fet, vx, grids, SA_SYN, 1, 0.0)
    m = TO.modl_from_ff(ff, np.array([1.0]))
    lb = scale * torch.sum(torch.tensor(wq) * (torch.tensor(target[b]) - m) ** 2 / uncert)
    lb.backward()
    loss += float(lb)
    modl.append(m.detach().numpy()); pbar.append(leaves.grad.numpy().copy()); fbar.append(fet.grad.numpy().copy())
zt = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tsadar_b200", "data", "zprime_table.npz"))
zx, zr, zi = (np.ascontiguousarray(zt[k], dtype=np.float64) for k in ("x", "re", "im"))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "cabi_fixture.bin")
with open(out, "wb") as fo:
    fo.write(b"TSFFFIX1")
    fo.write(struct.pack("<8i", B, W, V, NP, zx.size, 0, 0, 0))
    fo.write(struct.pack("<8d", LAM_RANGE[0], LAM_RANGE[1], vx[0], vx[1] - vx[0], float(SA_SYN[0]), 1.0, uncert, scale))
    for a in (zx, zr, zi, params.astype(np.float64), fe.astype(np.float32), target, wq, np.array(modl), np.array([loss]), np.array(pbar), np.array(fbar)):
        fo.write(np.ascontiguousarray(a).tobytes())
print(out, os.path.getsize(out), "bytes; loss", loss)
