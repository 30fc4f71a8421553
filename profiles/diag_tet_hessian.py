"""diagnostic: per-tet elastic Hessian blocks, GPU vs oracle, on C3: which tets differ and what their sigma-space quantities look like"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc
from ipc_b200 import lib as L, scenes

m, info = scenes.ball_on_mat_c3(nx=200)
ctx = L.Context(0)
ctx.set_mesh(m.V_rest_soa, m.T_soa, m.restTriInv, m.vol, m.mu, m.lam, m.mass, m.dbc, m.energy)
ia, ja = m.csr_pattern(1)
ctx.set_csr(ia, ja, 1)
ctx.set_state(m.V_soa)
coef = 0.025 ** 2
for spd in (1, 0):
    ctx.elastic_hessian(coef, 1, spd, 1, None)
    h78 = L.untile_hessians(ctx.download(L.BUF_TET_HESSIANS, 78 * 64 * ((m.nT + 63) // 64)), m.nT)
    H_ref = orc.Elastic(m).hessian_blocks(coef, spd, nthreads=64)
    err = np.empty(m.nT)
    for t in range(m.nT):
        err[t] = np.abs(orc.blocks78_to_dense(h78[t], m.T[t]) - H_ref[t]).max() / np.abs(H_ref[t]).max()
    bad = np.nonzero(err > 1e-9)[0]
    print("projectSPD", spd, "tets with rel err > 1e-9:", len(bad), "of", m.nT, "max", err.max(), "mat tets among them:", int((bad < 240000).sum()))
    x = m.V[m.T]; X = m.V_rest[m.T]
    for t in bad[:6]:
        Ds = np.stack([x[t, 1] - x[t, 0], x[t, 2] - x[t, 0], x[t, 3] - x[t, 0]], axis=1)
        Dm = np.stack([X[t, 1] - X[t, 0], X[t, 2] - X[t, 0], X[t, 3] - X[t, 0]], axis=1)
        F = Ds @ np.linalg.inv(Dm)
        sig = np.linalg.svd(F, compute_uv=False)
        mu, lam = m.mu[t], m.lam[t]
        dE = mu * (sig - 1 / sig) + lam * np.log(np.prod(sig)) / sig
        print(" tet", t, "err", err[t], "sigma", sig, "dPsi/dsigma sums (01,12,20)", dE[0] + dE[1], dE[1] + dE[2], dE[2] + dE[0])
