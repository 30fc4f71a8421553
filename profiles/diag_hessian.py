"""diagnostic: which CSR entries of the barrier Hessian differ between the GPU path and the oracle on a scene (default C3)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc
from ipc_b200 import lib as L, scenes
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
m, info = scenes.ball_on_mat_c3(nx=200) if which == "c3" else scenes.squeeze_out_tiled()
dHat, kappa = info["dHat"], 1e8
ctx = L.Context(0)
ctx.set_mesh(m.V_rest_soa, m.T_soa, m.restTriInv, m.vol, m.mu, m.lam, m.mass, m.dbc, m.energy)
ctx.set_surface(m.SVI, m.SFEdges, m.SF_soa, m.vCoDim)
ctx.set_state(m.V_soa)
s = orc.Surf(m)
mm, pa, pe, cand = ctx.constraint_set(dHat, 1)
ia, ja = m.csr_pattern(1, extra_pairs=bench.contact_pattern_pairs(m, mm, pa, pe))
ctx.set_csr(ia, ja, 1)
a = np.zeros(ja.size); ctx.barrier_hessian(dHat, kappa, 1, a)
a_r = s.barrier_hessian_csr(mm, pa, pe, dHat, kappa, ia, ja, 1, 1, nthreads=64)
print("barrier-only rel err", np.linalg.norm(a - a_r) / np.linalg.norm(a_r), "max abs", np.abs(a - a_r).max(), "max ref", np.abs(a_r).max())
e = np.zeros(ja.size); ctx.elastic_hessian(0.025 ** 2, 1, 1, 1, e)
e_r = orc.Elastic(m).hessian_csr(0.025 ** 2, ia, ja, 1, 1, 1, nthreads=64)
print("elastic-only rel err", np.linalg.norm(e - e_r) / np.linalg.norm(e_r), "max abs", np.abs(e - e_r).max(), "max ref", np.abs(e_r).max())
# per pair: upload one pair at a time for the worst rows
bad = np.argsort(-np.abs(a - a_r))[:8]
rows = np.searchsorted(ia - 1, bad, side="right") - 1
print("worst entries (row, col, gpu, ref):", [(int(r), int(ja[k] - 1), float(a[k]), float(a_r[k])) for r, k in zip(rows, bad)])
vbad = set(int(r) // 3 for r in rows)
for c, row in enumerate(mm):
    vs = [(-row[0] - 1) if row[0] < 0 else row[0]] + [x for x in row[1:] if x >= 0]
    if vbad & set(int(v) for v in vs):
        ctx.set_constraint_set(row[None, :], pa[:0], pe[:0])
        a1 = np.zeros(ja.size); ctx.barrier_hessian(dHat, kappa, 1, a1)
        a1r = s.barrier_hessian_csr(row[None, :], pa[:0], pe[:0], dHat, kappa, ia, ja, 1, 1)
        print("pair", row, "rel", np.linalg.norm(a1 - a1r) / max(np.linalg.norm(a1r), 1e-300), "norm", np.linalg.norm(a1r))
