"""CPU tests of the oracle's elastic restatement: golden vectors (mpmath), FD self-consistency, invariants."""
import json
import os

import numpy as np
import pytest

import oracle as orc
from ipc_b200 import mesh as M

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "elastic_golden.json")))
MU, LAM = GOLD["mu"], GOLD["lam"]


def rand_F(rng, scale=0.3):
    return np.eye(3) + scale * rng.standard_normal((3, 3))


def test_svd_convention_random():
    rng = np.random.default_rng(0)
    for k in range(500):
        F = rand_F(rng, 0.6) if k % 3 else rng.standard_normal((3, 3))
        U, S, V = orc.svd3(F)
        assert np.allclose(U @ np.diag(S) @ V.T, F, atol=1e-13 * max(1, abs(F).max()))
        assert np.allclose(U @ U.T, np.eye(3), atol=1e-13) and np.allclose(V @ V.T, np.eye(3), atol=1e-13)
        assert abs(np.linalg.det(U) - 1) < 1e-12 and abs(np.linalg.det(V) - 1) < 1e-12
        assert abs(S[0]) >= abs(S[1]) >= abs(S[2]) and S[0] >= 0 and S[1] >= 0
        assert np.sign(S[2]) == np.sign(np.linalg.det(F)) or abs(S[2]) < 1e-14
        assert np.allclose(np.abs(S), np.linalg.svd(F, compute_uv=False), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("F", [np.eye(3), np.diag([2.0, 2.0, 2.0]), np.diag([1.0, 1.0, -1.0]), np.zeros((3, 3)),
                               np.array([[1, 1, 0], [0, 1, 1], [0, 0, 1.0]]), np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0.0]]),
                               np.diag([1e-8, 1.0, 1e8]), np.array([[1, 2, 3], [2, 4, 6], [3, 6, 9.0]])])
def test_svd_special(F):
    U, S, V = orc.svd3(F)
    assert np.allclose(U @ np.diag(S) @ V.T, F, atol=1e-12 * max(1.0, abs(F).max()))
    assert abs(np.linalg.det(U) - 1) < 1e-12 and abs(np.linalg.det(V) - 1) < 1e-12


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"{c['name']}-e{c['energy']}")
def test_golden_psi_P_dPdF(case):
    et, F = case["energy"], np.array(case["F"])
    U, S, V = orc.svd3(F)
    assert np.allclose(np.abs(S), case["sigma"], rtol=1e-13, atol=1e-15)
    scale = max(1.0, abs(case["psi"]))
    assert abs(orc.psi(et, S, MU, LAM) - case["psi"]) <= 1e-12 * scale
    Pg = np.array(case["P"])
    assert np.allclose(orc.pk1(et, F, MU, LAM), Pg, rtol=0, atol=1e-12 * max(1.0, abs(Pg).max()))
    Hg = np.array(case["dPdF"])
    H = orc.dPdF(et, F, MU, LAM, 1.0, 0)
    tol = 1e-10 if case["name"] != "near_degenerate" else 1e-6  # (psi'_i+psi'_j)/(s_i+s_j) stays smooth, FD golden is the weak side
    assert np.allclose(H, Hg, rtol=0, atol=tol * abs(Hg).max())
    assert np.allclose(H, H.T, atol=1e-12 * abs(H).max())


def test_rest_state_zero():  # NeoHookeanEnergy.cpp:156-170 checkEnergyVal: sigma = 1 -> E = 0, P = 0
    for et in (0, 1):
        assert orc.psi(et, np.ones(3), MU, LAM) == 0.0
        assert np.allclose(orc.pk1(et, np.eye(3), MU, LAM), 0.0, atol=1e-13)
        assert np.allclose(orc.dpsi(et, np.ones(3), MU, LAM), 0.0, atol=1e-13)


@pytest.mark.parametrize("et", [0, 1])
def test_sigma_derivatives_fd(et):  # Energy.cpp:584-700 unit tests, h = 1e-6
    rng = np.random.default_rng(1)
    for _ in range(20):
        S = 1.0 + 0.3 * rng.standard_normal(3)
        S = np.abs(S) + 0.2
        h = 1e-6
        g = orc.dpsi(et, S, MU, LAM)
        H = orc.d2psi(et, S, MU, LAM)
        for i in range(3):
            e = np.zeros(3); e[i] = h
            fd = (orc.psi(et, S + e, MU, LAM) - orc.psi(et, S - e, MU, LAM)) / (2 * h)
            assert abs(fd - g[i]) < 1e-6 * max(1, abs(g).max())
            fdH = (orc.dpsi(et, S + e, MU, LAM) - orc.dpsi(et, S - e, MU, LAM)) / (2 * h)
            assert np.allclose(fdH, H[:, i], atol=1e-6 * max(1, abs(H).max()))


def test_makePD():
    rng = np.random.default_rng(2)
    for n in (3, 6, 9, 12):
        for _ in range(10):
            A = rng.standard_normal((n, n)); A = A + A.T
            P = orc.makePD(A)
            w, Q = np.linalg.eigh(A)
            ref = (Q * np.maximum(w, 0)) @ Q.T
            assert np.allclose(P, ref, atol=1e-12 * abs(A).max())
        A = rng.standard_normal((n, n)); A = A @ A.T + np.eye(n)
        assert np.array_equal(orc.makePD(A), A)  # lambda_min >= 0: returned unchanged (IglUtils.hpp:123-125)


def small_mesh(et, seed=3, inverted=False):
    V, T = M.grid_tets(3, 3, 2)
    m = M.Mesh(V, T, energy=et)
    M.deform(m, seed, twist=0.6, amp=0.03, noise=0.05, require_positive=not inverted)
    if inverted:
        rng = np.random.default_rng(seed)
        m.V[rng.integers(0, m.nV, 4)] += 0.4 * rng.standard_normal((4, 3))
    return m


@pytest.mark.parametrize("et", [0, 1])
def test_mesh_gradient_fd_and_invariants(et):
    m = small_mesh(et)
    coef = 0.025 ** 2
    o = orc.Elastic(m)
    E, per = o.energy(coef)
    assert abs(E - coef * per.sum()) <= 1e-15 * abs(E)
    g = o.gradient(coef, projectDBC=0)
    assert np.allclose(g.reshape(-1, 3).sum(0), 0, atol=1e-12 * abs(g).max())  # translation invariance
    x = m.V
    torque = np.cross(x, g.reshape(-1, 3)).sum(0)
    assert np.allclose(torque, 0, atol=1e-10 * abs(g).max())  # rotation invariance
    rng = np.random.default_rng(5)
    for _ in range(6):
        dx = rng.standard_normal(m.V.shape)
        h = 1e-6
        Ep, _ = orc.Elastic(m, V=m.V + h * dx).energy(coef)
        Em, _ = orc.Elastic(m, V=m.V - h * dx).energy(coef)
        fd = (Ep - Em) / (2 * h)
        assert abs(fd - g @ dx.ravel()) <= 1e-6 * np.linalg.norm(g) * np.linalg.norm(dx)


@pytest.mark.parametrize("et", [0, 1])
def test_mesh_hessian_fd_psd_and_csr(et):
    m = small_mesh(et)
    coef = 0.025 ** 2
    o = orc.Elastic(m)
    H0 = o.hessian_blocks(coef, projectSPD=0)
    # FD of the gradient against the unprojected Hessian (Energy::checkHessian, Energy.cpp:120-192)
    dense = np.zeros((3 * m.nV, 3 * m.nV))
    for t in range(m.nT):
        idx = (3 * m.T[t][:, None] + np.arange(3)).ravel()
        dense[np.ix_(idx, idx)] += H0[t]
    rng = np.random.default_rng(6)
    dx = rng.standard_normal(m.V.shape)
    h = 1e-6
    gp = orc.Elastic(m, V=m.V + h * dx).gradient(coef, 0)
    gm = orc.Elastic(m, V=m.V - h * dx).gradient(coef, 0)
    fd = (gp - gm) / (2 * h)
    assert np.allclose(fd, dense @ dx.ravel(), atol=2e-6 * abs(dense @ dx.ravel()).max())
    # projected blocks are PSD and symmetric
    H1 = o.hessian_blocks(coef, projectSPD=1)
    for t in range(m.nT):
        assert np.allclose(H1[t], H1[t].T, atol=1e-12 * abs(H1[t]).max())
        assert np.linalg.eigvalsh(H1[t]).min() >= -1e-10 * abs(H1[t]).max()
    # CSR sink == dense upper triangle, with Dirichlet rows -> identity (IglUtils.hpp:44-53)
    dbc = np.zeros(m.nV, dtype=np.uint8); dbc[[0, 5]] = 1; dbc[7] = 2
    ia, ja = m.csr_pattern(index_base=1)
    for projectDBC in (0, 1):
        a = orc.Elastic(m, dbc=dbc).hessian_csr(coef, ia, ja, 1, 1, projectDBC)
        dense = np.zeros((3 * m.nV, 3 * m.nV))
        fixed = (dbc == 1) | ((dbc == 2) & bool(projectDBC))
        for t in range(m.nT):
            for a_ in range(4):
                for b_ in range(4):
                    va, vb = m.T[t][a_], m.T[t][b_]
                    if fixed[va] or fixed[vb]:
                        continue
                    dense[3 * va:3 * va + 3, 3 * vb:3 * vb + 3] += H1[t][3 * a_:3 * a_ + 3, 3 * b_:3 * b_ + 3]
        for v in np.nonzero(fixed)[0]:
            dense[3 * v:3 * v + 3, 3 * v:3 * v + 3] = np.eye(3)
        for r in range(3 * m.nV):
            cols = ja[ia[r] - 1:ia[r + 1] - 1] - 1
            assert np.all(cols >= r)
            assert np.allclose(a[ia[r] - 1:ia[r + 1] - 1], dense[r, cols], atol=1e-12 * abs(dense).max())


def test_fcr_inverted_tets_finite():  # FixedCoRotEnergy.cpp:173-176: inversion allowed
    m = small_mesh(1, inverted=True)
    o = orc.Elastic(m)
    x = m.V[m.T]
    det = np.linalg.det(np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0], x[:, 3] - x[:, 0]], axis=2))
    assert det.min() < 0
    E, _ = o.energy(1.0)
    assert np.isfinite(E) and np.all(np.isfinite(o.gradient(1.0))) and np.all(np.isfinite(o.hessian_blocks(1.0)))


def test_inversion_step():
    m = small_mesh(0)
    rng = np.random.default_rng(8)
    p = rng.standard_normal(3 * m.nV) * 0.5
    a, per = orc.Elastic(m).inversion_step(p, 0.2, 1.0)
    assert 0 < a <= 1.0
    P = p.reshape(-1, 3)

    def dets(t):
        x = (m.V + t * P)[m.T]
        return np.linalg.det(np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0], x[:, 3] - x[:, 0]], axis=2))

    d0 = dets(0.0)
    for t in range(m.nT):
        if per[t] < 1e19:
            x = (m.V + per[t] * P)[m.T[t]]
            dt = np.linalg.det(np.stack([x[1] - x[0], x[2] - x[0], x[3] - x[0]], axis=1))
            assert abs(dt - 0.2 * d0[t]) <= 1e-9 * abs(d0[t])  # volume shrinks to slack * current volume
    assert a == min(per.min(), 1.0)
    assert dets(a * 0.999).min() > 0


def test_csr_pattern_matches_reference_layout():
    m = small_mesh(0)
    ia, ja = m.csr_pattern(index_base=1)
    lo, hi = m.neighbor_pairs()
    ptr = np.zeros(m.nV + 1, dtype=np.int32)
    nbr = []
    adj = [[] for _ in range(m.nV)]
    for a_, b_ in zip(lo, hi):
        adj[a_].append(b_); adj[b_].append(a_)
    for v in range(m.nV):
        adj[v].sort(); nbr += adj[v]; ptr[v + 1] = len(nbr)
    nbr = np.array(nbr, dtype=np.int32)
    import ctypes as C
    nnz = orc.lib().orc_csr_pattern(m.nV, orc.i(ptr), orc.i(nbr), 1, None, None)
    ia2 = np.empty(3 * m.nV + 1, dtype=np.int32); ja2 = np.empty(nnz, dtype=np.int32)
    orc.lib().orc_csr_pattern(m.nV, orc.i(ptr), orc.i(nbr), 1, orc.i(ia2), orc.i(ja2))
    assert np.array_equal(ia, ia2) and np.array_equal(ja, ja2)


def test_makePD2d_is_the_reference_formula_not_the_eigenvalue_clamp():
    """IglUtils::makePD2d (IglUtils.hpp:138-177) returns v v^T / L1 with v = (L1 - d, b) where the projection onto the positive eigen-direction would be
    L1 v v^T / |v|^2.  The oracle (and the kernels) restate the formula as it stands -- DESIGN.md 3.4; this test records the difference so that nobody
    "fixes" one side only."""
    import ctypes as C
    M = np.array([1.0, 2.0, 2.0, -1.0])
    out = M.copy()
    orc.lib().orc_makePD2d(orc.d(out))
    L1 = np.sqrt(5.0)
    v = np.array([L1 + 1.0, 2.0])
    assert np.allclose(out.reshape(2, 2), np.outer(v, v) / L1, rtol=1e-14)
    exact = L1 * np.outer(v, v) / (v @ v)
    assert np.abs(out.reshape(2, 2) - exact).max() > 1.0  # far from the eigenvalue clamp, yet positive semi-definite
    assert np.linalg.eigvalsh(out.reshape(2, 2)).min() >= -1e-12
    # the twist block at rest, [[mu, mu], [mu, mu]] with L2 = -0: comes back halved, whereas L2 = +0 leaves it untouched (the discontinuity)
    mu = 3.0
    for eps, factor in ((+1e-12, 0.5), (-1e-12, 1.0)):  # off-diagonal a hair above / below the diagonal: L2 = -+1e-12
        B = np.array([mu, mu + eps, mu + eps, mu])
        orc.lib().orc_makePD2d(orc.d(B))
        assert np.allclose(B, factor * mu, rtol=1e-9)
    # through dP/dF: whenever a block is projected the result differs from the exact PSD projection of the 9 x 9 matrix
    rng = np.random.default_rng(4)
    worst = 0.0
    for _ in range(200):
        F = np.eye(3) + 0.6 * rng.standard_normal((3, 3))
        if np.linalg.det(F) <= 0.05:
            continue
        H = orc.dPdF(0, F, 3.6e4, 1.4e5, 1.0, 0)
        Hp = orc.dPdF(0, F, 3.6e4, 1.4e5, 1.0, 1)
        w, Q = np.linalg.eigh((H + H.T) / 2)
        assert np.linalg.eigvalsh((Hp + Hp.T) / 2).min() >= -1e-9 * np.abs(H).max()  # it IS positive semi-definite
        worst = max(worst, np.abs(Hp - (Q * np.maximum(w, 0)) @ Q.T).max() / np.abs(H).max())
    assert worst > 0.05


def test_inversion_step_is_the_smallest_positive_root():
    """Energy::filterStepSize / getSmallestPositiveRealCubicRoot (Energy.cpp:565-581, get_feasible_steps.cpp): per tet the step is the SMALLEST positive
    root of the cubic  det(X + t P) = slack * det(X)  -- checked against numpy's companion-matrix roots of the cubic interpolated through four
    exact determinants (an independent route: the oracle and the kernel use Cardano's formulas in complex arithmetic)."""
    m = small_mesh(0)
    rng = np.random.default_rng(9)
    p = rng.standard_normal(3 * m.nV) * 0.8
    slack = 0.2
    a, per = orc.Elastic(m).inversion_step(p, slack, 1e30)
    P = p.reshape(-1, 3)

    def dets(t):
        x = (m.V + t * P)[m.T]
        return np.linalg.det(np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0], x[:, 3] - x[:, 0]], axis=2))

    ts = np.array([0.0, 1.0, 2.0, 3.0])
    D = np.stack([dets(t) for t in ts])  # (4, nT): a cubic in t per tet, interpolated exactly by four samples
    coef = np.linalg.solve(np.vander(ts, 4), D)  # rows: t^3, t^2, t, 1
    n_roots = 0
    for t in range(m.nT):
        c = coef[:, t].copy()
        c[3] -= slack * D[0, t]
        r = np.roots(c)
        pos = sorted(x.real for x in r if abs(x.imag) <= 1e-9 * max(1.0, abs(x)) and x.real > 0)
        if pos:
            n_roots += 1
            assert abs(per[t] - pos[0]) <= 1e-8 * pos[0], (t, per[t], pos)
        else:
            assert per[t] > 1e19  # no positive root: no bound from this tet
    assert n_roots > m.nT // 4 and a == per.min()
