"""CPU tests of the friction oracle (oracle/friction.cpp): the reference holds no expected outputs for these terms
(SelfCollisionHandler.cpp:2481-2987, FrictionUtils.hpp), so the restatement is pinned by finite differences (E -> g -> unprojected H),
by the closed form H = T^T S T, and by invariants."""
import numpy as np
import pytest

import oracle as orc
from ipc_b200 import scenes

KAPPA = 1e8
COEF = 0.3  # animConfig.selfFric


def friction_scene(seed=0, scale=1.0):
    """4-ball pile: every pair kind (PT, EE, PE, PP incl. multiplicities); V_prev = V - a displacement with a rigid sliding part per ball
    plus noise, sized so that both the static (|u| <= eps) and the sliding branch occur."""
    m, info = scenes.ball_pile(4, res=8, seed=5, height=4)
    s = orc.Surf(m)
    mm, _, _, _ = s.constraint_set(info["dHat"], nthreads=8)
    rng = np.random.default_rng(seed)
    h = m.avgEdgeLen
    ball = np.argmin(np.linalg.norm(m.V[:, None, :] - info["centers"][None, :, :], axis=2), axis=1)
    slide = rng.normal(size=(len(info["centers"]), 3)) * 2e-3 * h * scale
    disp = slide[ball] + rng.normal(size=m.V.shape) * 2e-4 * h * scale
    Vt = m.V - disp
    lam, co, ba = s.friction_lag(mm, info["dHat"], KAPPA)
    # eps2 = median of the tangential slip: half the pairs on each side of the clamp
    u2 = slip2(m.V, Vt, mm, co, ba)
    return m, info, s, mm, Vt, lam, co, ba, float(np.median(u2))


def weights(mm4, c):
    if mm4[0] >= 0:
        return [mm4[0], mm4[1], mm4[2], mm4[3]], [1 - c[0], c[0], c[1] - 1, -c[1]]
    v0 = -mm4[0] - 1
    if mm4[2] < 0:
        return [v0, mm4[1]], [1.0, -1.0]
    if mm4[3] < 0:
        return [v0, mm4[1], mm4[2]], [1.0, c[0] - 1, -c[0]]
    return [v0, mm4[1], mm4[2], mm4[3]], [1.0, -1 + c[0] + c[1], -c[0], -c[1]]


def slip2(V, Vt, mm, co, ba):
    out = np.empty(len(mm))
    for k, (m4, c, b) in enumerate(zip(mm, co, ba)):
        vs, w = weights(m4, c)
        r = sum(wi * (V[v] - Vt[v]) for v, wi in zip(vs, w))
        B = b.reshape(2, 3)
        out[k] = np.sum((B @ r) ** 2)
    return out


def test_lagged_data_is_consistent():
    m, info, s, mm, Vt, lam, co, ba, eps2 = friction_scene()
    assert len(mm) > 100 and (lam > 0).all()  # repulsive normal force: -kappa b'(d) 2 sqrt(d) > 0 for d < dHat
    B = ba.reshape(-1, 2, 3)
    assert np.allclose(np.einsum("nij,nkj->nik", B, B), np.eye(2)[None], atol=1e-12)  # orthonormal tangent frame
    for m4, c, b in zip(mm, co, B):
        vs, w = weights(m4, c)
        assert abs(sum(w)) < 1e-12  # relative displacement: translation invariant
        r = sum(wi * m.V[v] for v, wi in zip(vs, w))  # closest-point difference = the contact normal direction (times the distance)
        assert np.linalg.norm(r) > 0
        assert abs(b[0] @ r) <= 1e-9 * np.linalg.norm(r) and abs(b[1] @ r) <= 1e-9 * np.linalg.norm(r)
    # multiplicities scale lambda (Optimizer.cpp:1588-1591)
    dup = mm[:, 3] < -1
    assert dup.any()
    mm1 = mm.copy()
    mm1[dup, 3] = -1
    lam1, _, _ = s.friction_lag(mm1, info["dHat"], KAPPA)
    assert np.allclose(lam[dup], lam1[dup] * (-mm[dup, 3]), rtol=1e-15)


def test_energy_matches_closed_form_and_is_c1_at_the_clamp():
    m, info, s, mm, Vt, lam, co, ba, eps2 = friction_scene()
    u2 = slip2(m.V, Vt, mm, co, ba)
    eps = np.sqrt(eps2)
    f0 = u2 * (-np.sqrt(u2) / 3 + eps) / eps ** 2 + eps / 3
    E_ref = COEF * np.sum(lam * np.where(u2 > eps2, np.sqrt(u2), f0))
    E = s.friction_energy(Vt, mm, lam, co, ba, eps2, COEF)
    assert abs(E - E_ref) <= 1e-13 * abs(E_ref)
    assert (u2 > eps2).sum() > 20 and (u2 <= eps2).sum() > 20
    # f0(eps^2) = eps and f0' matches the slope of |u| there (FrictionUtils.hpp:278-287)
    assert abs((eps2 * (-eps / 3 + eps) / eps ** 2 + eps / 3) - eps) < 1e-15 * eps
    assert abs((-eps + 2 * eps) / eps ** 2 * eps - 1.0) < 1e-14


def test_gradient_is_the_derivative_of_the_energy():
    m, info, s, mm, Vt, lam, co, ba, eps2 = friction_scene()
    g = s.friction_gradient(Vt, mm, lam, co, ba, eps2, COEF)
    verts = np.unique(np.concatenate([weights(m4, c)[0] for m4, c in zip(mm[:40], co[:40])]))
    rng = np.random.default_rng(1)
    h = 1e-7 * m.avgEdgeLen
    for v in rng.choice(verts, 12, replace=False):
        for q in range(3):
            Vp, Vm = m.V.copy(), m.V.copy()
            Vp[v, q] += h
            Vm[v, q] -= h
            Ep = orc.Surf(m, V=Vp).friction_energy(Vt, mm, lam, co, ba, eps2, COEF)
            Em = orc.Surf(m, V=Vm).friction_energy(Vt, mm, lam, co, ba, eps2, COEF)
            fd = (Ep - Em) / (2 * h)
            assert abs(fd - g[3 * v + q]) <= 2e-5 * max(abs(g[3 * v + q]), np.abs(g).max() * 1e-3), (v, q, fd, g[3 * v + q])
    assert np.abs(g.reshape(-1, 3).sum(axis=0)).max() <= 1e-9 * np.abs(g).max()  # internal forces sum to zero


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_pair_hessian_fd_structure_and_projection(kind):
    m, info, s, mm, Vt, lam, co, ba, eps2 = friction_scene()
    kinds = np.where(mm[:, 0] >= 0, 1, np.where(mm[:, 2] < 0, 3, np.where(mm[:, 3] < 0, 2, 0)))
    u2 = slip2(m.V, Vt, mm, co, ba)
    idx = np.nonzero(kinds == kind)[0]
    picked = [i for i in idx if u2[i] > eps2][:2] + [i for i in idx if u2[i] <= eps2][:2]
    assert picked
    for c in picked:
        H0, nv = s.friction_pair_hessian(Vt, mm[c], lam[c], co[c], ba[c], eps2, COEF, project=0)
        H1, _ = s.friction_pair_hessian(Vt, mm[c], lam[c], co[c], ba[c], eps2, COEF, project=1)
        n = 3 * nv
        vs, w = weights(mm[c], co[c])
        # finite differences of this pair's gradient
        one = lambda V: orc.Surf(m, V=V).friction_gradient(Vt, mm[c:c + 1], lam[c:c + 1], co[c:c + 1], ba[c:c + 1], eps2, COEF)
        h = 1e-6 * np.sqrt(max(u2[c], eps2))
        Hfd = np.zeros((n, n))
        for j, v in enumerate(vs):
            for q in range(3):
                Vp, Vm = m.V.copy(), m.V.copy()
                Vp[v, q] += h
                Vm[v, q] -= h
                dg = (one(Vp) - one(Vm)) / (2 * h)
                Hfd[:, 3 * j + q] = np.concatenate([dg[3 * vi:3 * vi + 3] for vi in vs])
        assert np.abs(Hfd - H0[:n, :n]).max() <= 1e-5 * np.abs(H0).max(), (kind, c)
        # closed form: H = T^T S T, T = [w_k B^T], S = c lam (a I + b u u^T)
        B = ba[c].reshape(2, 3)
        T = np.concatenate([wk * B for wk in w], axis=1)
        r = sum(wi * (m.V[v] - Vt[v]) for v, wi in zip(vs, w))
        u = B @ r
        x2 = u @ u
        eps = np.sqrt(eps2)
        if x2 > eps2:
            S = COEF * lam[c] * (np.eye(2) / np.sqrt(x2) - np.outer(u, u) / x2 ** 1.5)
        else:
            f1, f2 = (-np.sqrt(x2) + 2 * eps) / eps2, 2 * (eps - np.sqrt(x2)) / eps2
            S = COEF * lam[c] * (f1 * np.eye(2) + (f2 - f1) / x2 * np.outer(u, u))
        assert np.abs(T.T @ S @ T - H0[:n, :n]).max() <= 1e-12 * np.abs(H0).max()
        # the block is PSD already: makePD changes it by rounding only; translations are in its null space
        assert np.abs(H1 - H0).max() <= 1e-12 * np.abs(H0).max()
        assert np.linalg.eigvalsh(H1[:n, :n]).min() >= -1e-12 * np.abs(H1).max()
        for q in range(3):
            t = np.zeros(n)
            t[q::3] = 1.0
            assert np.abs(H1[:n, :n] @ t).max() <= 1e-12 * np.abs(H1).max()
        assert n == 12 or (np.abs(H1[n:, :]).max() == 0 and np.abs(H1[:, n:]).max() == 0)


def test_csr_hessian_equals_the_sum_of_the_pair_blocks():
    m, info, s, mm, Vt, lam, co, ba, eps2 = friction_scene()
    import bench
    ia, ja = m.csr_pattern(1, extra_pairs=bench.contact_pattern_pairs(m, mm, np.zeros((0, 4), np.int32), np.zeros((0, 2), np.int32)))
    a = s.friction_hessian_csr(Vt, mm, lam, co, ba, eps2, COEF, ia, ja, 1, projectDBC=1, nthreads=4)
    import scipy.sparse as sp
    A = sp.csr_matrix((a, ja - 1, ia - 1), shape=(3 * m.nV, 3 * m.nV)).toarray()
    D = np.zeros_like(A)
    for c in range(len(mm)):
        H, nv = s.friction_pair_hessian(Vt, mm[c], lam[c], co[c], ba[c], eps2, COEF)
        vs, _ = weights(mm[c], co[c])
        for i_, vi in enumerate(vs):
            for j_, vj in enumerate(vs):
                D[3 * vi:3 * vi + 3, 3 * vj:3 * vj + 3] += H[3 * i_:3 * i_ + 3, 3 * j_:3 * j_ + 3]
    assert np.abs(A - np.triu(D)).max() <= 1e-12 * np.abs(D).max()
