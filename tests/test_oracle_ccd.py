"""CPU tests of the oracle's CCD restatement (Tight-Inclusion is "parity unpinned": no reference vectors exist; the
checks below are the reference's own unit-test geometries with their analytic hit criteria, conservativeness and
consistency properties)."""
import numpy as np
import pytest

import oracle as orc
from ipc_b200 import mesh as M
from ipc_b200 import scenes

ERR = np.full(3, 1e-13)


@pytest.mark.parametrize("u0y", [-1.1, 0.0, 1.1])
@pytest.mark.parametrize("u1y", [-1.1, 0.0, 1.1])
def test_reference_pt_geometry(u0y, u1y):  # tests/Collisions/CollisionConstraintTests.cpp:18-35
    v0 = np.array([0, 1, -0.5]); v1 = np.array([-1, 0, 1.0]); v2 = np.array([1, 0, 1.0]); v3 = np.array([0, 0, -1.0])
    u0 = np.array([0, u0y, 0]); u1 = np.array([0, u1y, 0])
    x0 = np.array([v0, v1, v3, v2]); x1 = np.array([v0 + u0, v1 + u1, v3 + u1, v2 + u1])
    hit, toi, _ = orc.ti("vf", x0, x1, ERR, 0.0)
    assert hit == (-u0y + u1y >= 1)
    if hit:
        exact = 1.0 / (u1y - u0y)
        assert exact - 1e-4 <= toi <= exact  # conservative: never later than the true time of impact


@pytest.mark.parametrize("ydisp", [-2.0, 0.0, 2.0])
def test_reference_ee_geometry(ydisp):  # CollisionConstraintTests.cpp:83-99
    v0 = np.array([-1, -1, 0.0]); v1 = np.array([1, -1, 0.0]); v2 = np.array([0, 1, -1.0]); v3 = np.array([0, 1, 1.0])
    u0 = np.array([0, ydisp, 0]); u1 = np.array([0, -ydisp, 0])
    x0 = np.array([v0, v1, v2, v3]); x1 = np.array([v0 + u0, v1 + u0, v2 + u1, v3 + u1])
    hit, toi, _ = orc.ti("ee", x0, x1, ERR, 0.0)
    assert hit == (ydisp >= 1.0)
    if hit:
        assert 0.5 - 1e-4 <= toi <= 0.5


def test_minimum_separation_and_max_t():
    v0 = np.array([0.2, 0.2, 1.0]); tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0.0]])
    x0 = np.vstack([v0, tri]); x1 = x0.copy(); x1[0, 2] = -1.0  # point falls through the triangle at t = 0.5
    hit, toi, _ = orc.ti("vf", x0, x1, ERR, 0.0)
    assert hit and 0.5 - 5e-4 <= toi <= 0.5
    hit, toi_ms, _ = orc.ti("vf", x0, x1, ERR, 0.1)  # stop 0.1 before contact: toi ~ (1-0.1)/2
    assert hit and 0.45 - 5e-4 <= toi_ms <= 0.45
    hit, _, _ = orc.ti("vf", x0, x1, ERR, 0.0, max_t=0.4)  # impact lies beyond max_t
    assert not hit
    hit, toi2, _ = orc.ti("vf", x0, x1, ERR, 0.0, max_t=0.6)
    assert hit and abs(toi2 - toi) < 1e-4
    x1m = x0.copy(); x1m[0, 0] += 5.0  # tangential motion: never hits
    assert not orc.ti("vf", x0, x1m, ERR, 0.0)[0]
    xs = x0.copy(); xs[0, 2] = 1e-9  # starts (almost) in contact while approaching: no_zero_toi must return a positive time
    x1s = xs.copy(); x1s[0, 2] = -1.0
    hit, toi0, _ = orc.ti("vf", xs, x1s, ERR, 0.0)
    assert hit and 0.0 < toi0 <= 1e-9 * 1.0000001


def test_conservative_on_random_pairs():
    """toi returned by TI is never later than the first time the exact distance drops below ms (sampled)."""
    rng = np.random.default_rng(3)
    hits = 0
    for _ in range(200):
        x0 = rng.standard_normal((4, 3)); x1 = x0 + 0.8 * rng.standard_normal((4, 3))
        for kind in ("vf", "ee"):
            hit, toi, _ = orc.ti(kind, x0, x1, ERR, 1e-3)
            ts = np.linspace(0, 1, 4001)
            f = orc.point_tri_d if kind == "vf" else orc.edge_edge_d
            dmin = np.array([f(x0 + t * (x1 - x0)) for t in ts[::20]])
            if hit:
                hits += 1
                before = ts[::20] < toi - 1e-3
                assert np.all(dmin[before] > 0.0)
            else:
                assert dmin.min() > 1e-8  # no contact anywhere on the sampled trajectory
    assert hits > 20


def test_ti_error_filter():
    V = np.array([[0, 0, 0], [2, 1, 0.5]], dtype=float)
    evf, eee = orc.ti_error(np.ascontiguousarray(V.T).ravel(), 2)
    r = 0.5 * np.sqrt(4 + 1 + 0.25)
    c = np.array([1, 0.5, 0.25])
    mx = np.maximum(np.abs(c + 10 * r / np.sqrt(3)), 1.0)
    assert np.allclose(eee, mx ** 3 * 7.105427357601002e-15, rtol=1e-14) and np.allclose(evf, mx ** 3 * 7.549516567451064e-15, rtol=1e-14)


def test_partial_and_full_ccd_on_stacked_balls():
    m, info = scenes.ball_pile(2, res=6, seed=11, height=2)
    s = orc.Surf(m)
    p = info["p"]
    evf, eee = orc.ti_error(s.V, m.nV, p)
    mm, pa, pe, cand = s.constraint_set(info["dHat"])
    assert len(cand) > 0
    a_part, z = orc.ccd_partial(s, p, cand, 1e-6, evf, eee, 1.0)
    assert z == 0 and 0 < a_part < 1.0
    g, a_grid = orc.grid_swept(s, p, 1.0, m.avgEdgeLen / 3)
    assert a_grid <= 1.0
    a_full, z, npairs = orc.ccd_full(s, p, g, a_grid, 1e-6, evf, eee, a_grid)
    assert z == 0 and npairs > len(cand)
    assert a_full <= min(a_part, a_grid) + 1e-15  # the full sweep sees at least the active candidates
    # moving by 0.999*alpha keeps every active distance positive (intersection free)
    V2 = m.V + 0.999 * a_full * p.reshape(-1, 3)
    s2 = orc.Surf(m, V=V2)
    for c in cand:
        if c[0] < 0:
            v = [m.SVI[-c[0] - 1]] + list(m.SF[c[1]])
            assert orc.point_tri_d(V2[v]) > 0
        else:
            v = list(m.SFEdges[c[0]]) + list(m.SFEdges[c[1]])
            assert orc.edge_edge_d(V2[v]) > 0
    # separating motion: no bound
    a_sep, _ = orc.ccd_partial(s, -p, cand, 1e-6, evf, eee, 1.0)
    assert a_sep == 1.0


def test_hashed_drivers_equal_brute_force():
    """The reference-style hash (oracle/hash.cpp) must not change any result: same sets, same step bound."""
    m, info = scenes.ball_pile(3, res=6, seed=12, height=3)
    s = orc.Surf(m)
    p = info["p"]
    h = m.avgEdgeLen / 3
    mm, pa, pe, cand = s.constraint_set(info["dHat"])
    mm2, pa2, pe2, cand2 = s.constraint_set_hashed(info["dHat"], h, nthreads=4)
    assert len(mm) > 0
    assert np.array_equal(mm, mm2) and np.array_equal(pa, pa2) and np.array_equal(pe, pe2) and np.array_equal(cand, cand2)
    evf, eee = orc.ti_error(s.V, m.nV, p)
    g, a_grid = orc.grid_swept(s, p, 1.0, h)
    a_bf, z, n_bf = orc.ccd_full(s, p, g, a_grid, 1e-6, evf, eee, a_grid, nthreads=4)
    a_h, z2, n_h = orc.ccd_full_hashed(s, p, 1.0, h, 1e-6, evf, eee, nthreads=4)
    assert z == z2 == 0 and a_bf == a_h
    assert n_h >= n_bf  # voxel-id aliasing in the hash can only add (harmless) candidates


def test_ball_on_mat_config_c3():
    """BASELINE config C3 (ball on a one-cell-thick mat, barrier contact + CCD) at a size the brute-force enumeration sweeps: the
    reference-style hashed drivers equal brute force, the barrier is finite, and the TI step bound is conservative."""
    from ipc_b200 import scenes
    m, info = scenes.ball_on_mat(24, res=5)
    s = orc.Surf(m)
    dHat, p = info["dHat"], info["p"]
    hv = m.avgEdgeLen / 3.0
    mm, pa, pe, cand = s.constraint_set(dHat, nthreads=4)
    mm2, pa2, pe2, cand2 = s.constraint_set_hashed(dHat, hv, 4)
    assert len(mm) >= 10 and len(cand) >= len(mm)
    assert np.array_equal(mm, mm2) and np.array_equal(pa, pa2) and np.array_equal(pe, pe2) and np.array_equal(cand, cand2)
    E, bad = s.barrier_energy(mm, pa, pe, dHat, 1e8)
    assert bad == 0 and np.isfinite(E) and E > 0
    evf, eee = orc.ti_error(s.V, m.nV, p)
    a_part, _ = orc.ccd_partial(s, p, cand, 1e-6, evf, eee, 1.0, 4)
    g, ag = orc.grid_swept(s, p, a_part, hv)
    a_bf, z, npairs = orc.ccd_full(s, p, g, ag, 1e-6, evf, eee, ag, nthreads=4)
    a_h, z2, npairs2 = orc.ccd_full_hashed(s, p, a_part, hv, 1e-6, evf, eee, 4)
    assert z == 0 and z2 == 0 and npairs == npairs2 and a_bf == a_h and 0.0 < a_h <= a_part < 1.0  # the ball does reach the mat within the step
    V2 = m.V + a_h * p.reshape(-1, 3)
    for c in cand:
        if c[0] < 0:
            assert orc.point_tri_d(V2[[m.SVI[-c[0] - 1]] + list(m.SF[c[1]])]) > 0
        else:
            assert orc.edge_edge_d(V2[list(m.SFEdges[c[0]]) + list(m.SFEdges[c[1]])]) > 0


def test_tie_order_of_equal_times_does_not_change_the_results():
    """Tight-Inclusion leaves the order of boxes with equal t_lo inside a level to its heap; the oracle and the GPU fix it (ascending u_lo, v_lo).
    A result that depended on that choice would make "bit-exact against the oracle" weaker than "bit-exact against the library".  Measured here:
    with the opposite admissible order (descending) every time of impact is the same bit pattern -- on random pairs, on grazing / symmetric
    configurations built to create ties, and on every CCD candidate of a ball pile and of an obstacle scene."""
    rng = np.random.default_rng(11)
    cases = []
    for _ in range(400):
        x0 = rng.standard_normal((4, 3))
        cases.append((x0, x0 + 0.8 * rng.standard_normal((4, 3))))
    for _ in range(200):  # symmetric set-ups: a point above the centroid / an edge across the middle of another, moving straight in (ties in u, v)
        h = rng.uniform(0.01, 0.5)
        tri = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
        pnt = np.array([[1 / 3, 1 / 3, h]])
        x0 = np.concatenate([pnt, tri])
        x1 = x0.copy()
        x1[0, 2] -= rng.uniform(0.5, 2.0) * h * 2
        cases.append((x0, x1))
        e = np.array([[-1.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, -1.0, h], [0.0, 1.0, h]])
        e1 = e.copy()
        e1[2:, 2] -= rng.uniform(0.5, 2.0) * h * 2
        cases.append((e, e1))

    def run_all():
        out = []
        for x0, x1 in cases:
            for kind in ("vf", "ee"):
                hit, toi, tol = orc.ti(kind, x0, x1, ERR, 1e-6)
                out.append((hit, np.float64(toi).view(np.uint64) if hit else 0))
        return out

    try:
        orc.lib().orc_ti_debug_tie_order(0)
        canonical = run_all()
        m, info = scenes.ball_pile(4, res=8, seed=5, height=4)
        s = orc.Surf(m)
        evf, eee = orc.ti_error(s.V, m.nV, info["p"])
        _, _, _, cand = s.constraint_set(info["dHat"], 8)
        a_pile = orc.ccd_partial(s, info["p"], cand, 1e-6, evf, eee, 1.0, 8)
        g, ag = orc.grid_swept(s, info["p"], a_pile[0], m.avgEdgeLen / 3)
        f_pile = orc.ccd_full(s, info["p"], g, ag, 1e-6, evf, eee, ag, nthreads=8)
        orc.lib().orc_ti_debug_tie_order(1)
        reversed_ = run_all()
        a_pile_r = orc.ccd_partial(s, info["p"], cand, 1e-6, evf, eee, 1.0, 8)
        f_pile_r = orc.ccd_full(s, info["p"], g, ag, 1e-6, evf, eee, ag, nthreads=8)
    finally:
        orc.lib().orc_ti_debug_tie_order(0)
    assert sum(h for h, _ in canonical) > 300
    differing = [k for k, (a, b) in enumerate(zip(canonical, reversed_)) if a != b]
    assert not differing, (len(differing), differing[:5])
    assert a_pile == a_pile_r and f_pile == f_pile_r


def test_time_of_impact_is_tight_as_well_as_conservative():
    """Quantitative pin of the restated search against the geometry itself (no library needed): with minimum separation ms and tolerance delta the
    returned time lies between the first instant the exact distance reaches sqrt(3) (ms + O(delta)) and the first instant it reaches ms: the
    inclusion test asks whether the difference vector enters the CUBE of half-width ms (+ error bound) around the origin, whose corners are
    sqrt(3) ms away.  (Conservative: never later than contact at separation ms.  Tight: not earlier than cube and co-domain tolerance allow.)"""
    from scipy.optimize import brentq
    rng = np.random.default_rng(21)
    ms, delta = 1e-3, 1e-6
    checked = 0
    for _ in range(300):
        x0 = rng.standard_normal((4, 3))
        x1 = x0 + 0.8 * rng.standard_normal((4, 3))
        disp = np.abs(x1 - x0).max()
        for kind in ("vf", "ee"):
            f = orc.point_tri_d if kind == "vf" else orc.edge_edge_d
            dist = lambda t: np.sqrt(f(x0 + t * (x1 - x0)))

            def first_time(thr):
                ts = np.linspace(0.0, 1.0, 1501)
                v = np.array([dist(t) - thr for t in ts])
                k = np.nonzero(v <= 0)[0]
                if len(k) == 0:
                    return None
                if k[0] == 0:
                    return 0.0
                return brentq(lambda t: dist(t) - thr, ts[k[0] - 1], ts[k[0]], xtol=1e-13)

            if dist(0.0) <= ms + 1e-4:
                continue  # starts in contact: not the case under test
            hit, toi, _ = orc.ti(kind, x0, x1, ERR, ms, tol=delta)
            t_contact = first_time(ms)
            t_early = first_time(np.sqrt(3.0) * (ms + 4.0 * delta))  # the eps-cube's corner, widened by the co-domain tolerance and the error bound
            if t_contact is not None:
                assert hit, "a trajectory that reaches the minimum separation must be reported"
                assert toi <= t_contact + 1e-12
            if hit:
                assert t_early is not None and toi >= t_early - 2.0 * delta / max(disp, 1e-3), (kind, toi, t_early, t_contact)
                checked += 1
    assert checked > 40
