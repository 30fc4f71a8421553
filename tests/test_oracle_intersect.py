"""Pins the oracle's restatement of igl::predicates::orient3d (exact predicate) against exact RATIONAL arithmetic, and the
segment-triangle / mesh intersection checks against brute force and hand-built cases (SelfCollisionHandler.cpp:3254-3296)."""
from fractions import Fraction

import numpy as np

import oracle as orc
from ipc_b200 import mesh as M


def orient3d_rational(a, b, c, d):
    F = Fraction
    ad, bd, cd = ([F(float(x)) - F(float(y)) for x, y in zip(p, d)] for p in (a, b, c))
    det = ad[2] * (bd[0] * cd[1] - cd[0] * bd[1]) + bd[2] * (cd[0] * ad[1] - ad[0] * cd[1]) + cd[2] * (ad[0] * bd[1] - bd[0] * ad[1])
    return (det > 0) - (det < 0)


def test_orient3d_matches_exact_rational_arithmetic():
    rng = np.random.default_rng(0)
    n_exact_zero = n_tiny = 0
    for it in range(600):
        kind = it % 4
        # (kinds 2, 3 round to multiples of 1/64: keep them O(1) or larger so that no coordinate collapses to 0 -- the expansion
        #  arithmetic, like Shewchuk's, is exact only in the absence of underflow)
        a, b, c = rng.standard_normal((3, 3)) * 10.0 ** rng.integers(-3 if kind < 2 else 0, 4) + (0.0 if kind < 2 else 3.0)
        if kind == 0:      # generic
            d = rng.standard_normal(3)
        elif kind == 1:    # in the plane up to rounding: the filter must hand over to the exact stage
            u, v = rng.uniform(-1, 2, 2)
            d = a + u * (b - a) + v * (c - a)
        elif kind == 2:    # exactly coplanar: dyadic coordinates
            a, b, c = np.round(a * 64) / 64, np.round(b * 64) / 64, np.round(c * 64) / 64
            d = a + 0.5 * (b - a) + 0.25 * (c - a)
        else:              # one ulp off the plane
            a, b, c = np.round(a * 64) / 64, np.round(b * 64) / 64, np.round(c * 64) / 64
            d = a + 0.5 * (b - a) + 0.25 * (c - a)
            k = rng.integers(0, 3)
            d[k] = np.nextafter(d[k], np.inf if rng.random() < 0.5 else -np.inf)
        ref = orient3d_rational(a, b, c, d)
        assert orc.orient3d(a, b, c, d) == ref
        assert orc.orient3d(a, b, c, d, exact=True) == ref
        n_exact_zero += ref == 0
        n_tiny += kind in (1, 3)
    assert n_exact_zero >= 100 and n_tiny >= 200
    # antisymmetry under a swap, invariance under a cyclic shift
    a, b, c, d = rng.standard_normal((4, 3))
    assert orc.orient3d(a, b, c, d) == -orc.orient3d(b, a, c, d) == orc.orient3d(b, c, a, d)


def test_segment_triangle_cases():
    t0, t1, t2 = np.array([0.0, 0, 0]), np.array([1.0, 0, 0]), np.array([0.0, 1, 0])
    assert orc.seg_tri_intersect([0.2, 0.2, -1], [0.2, 0.2, 1], t0, t1, t2) == 1   # pierces the interior
    assert orc.seg_tri_intersect([0.2, 0.2, 0.1], [0.2, 0.2, 1], t0, t1, t2) == 0  # stops above
    assert orc.seg_tri_intersect([0.9, 0.9, -1], [0.9, 0.9, 1], t0, t1, t2) == 0   # crosses the plane outside
    assert orc.seg_tri_intersect([0.2, 0.2, 0.0], [0.2, 0.2, 1], t0, t1, t2) == 0  # endpoint exactly in the plane: "coplanar" -> d(PT) = 0 catches it
    assert orc.seg_tri_intersect([0.1, 0.1, 0.0], [0.5, 0.2, 0.0], t0, t1, t2) == 0  # coplanar segment
    rng = np.random.default_rng(1)
    for _ in range(300):  # against the parametric solution in exact rationals (away from the boundary of the test)
        tri = rng.standard_normal((3, 3))
        e0, e1 = rng.standard_normal((2, 3)) * 1.5
        n = np.cross(tri[1] - tri[0], tri[2] - tri[0])
        s0, s1 = n @ (e0 - tri[0]), n @ (e1 - tri[0])
        hit = False
        if s0 * s1 < 0:
            x = e0 + s0 / (s0 - s1) * (e1 - e0)
            M_ = np.stack([tri[1] - tri[0], tri[2] - tri[0], n], axis=1)
            u, v, _w = np.linalg.solve(M_, x - tri[0])
            if min(u, v, 1 - u - v) < -1e-9:
                hit = False
            elif min(u, v, 1 - u - v) > 1e-9:
                hit = True
            else:
                continue
        assert orc.seg_tri_intersect(e0, e1, tri[0], tri[1], tri[2]) == int(hit)


def two_cubes(dz):
    V1, T1 = M.grid_tets(2, 2, 2, h=0.5)
    V2, T2 = M.grid_tets(2, 2, 2, h=0.5, origin=(0.13, 0.21, dz))
    return M.merge_meshes([(V1, T1), (V2, T2)])


def test_mesh_intersection_check():
    s_free = orc.Surf(two_cubes(1.05))
    assert s_free.intersection_free(nthreads=2) == (True, 0)
    m = two_cubes(0.8)  # the upper cube dips 0.2 into the lower one
    ok, hits = orc.Surf(m).intersection_free(nthreads=2)
    assert not ok and hits > 0
    # brute force over all (triangle, edge) pairs agrees on the count, whatever the grid cell
    V = m.V
    brute = 0
    for f in m.SF:
        hit = False
        for e in m.SFEdges:
            if set(e) & set(f):
                continue
            if orc.seg_tri_intersect(V[e[0]], V[e[1]], V[f[0]], V[f[1]], V[f[2]]):
                hit = True
                break
        brute += hit
    assert hits == brute
    assert orc.Surf(m).intersection_free(cell=0.05, nthreads=2) == (False, brute) and orc.Surf(m).intersection_free(cell=3.0) == (False, brute)
    # all-Dirichlet pairs are skipped (:3282): freeze both bodies -> nothing is reported
    m.dbc[:] = 1
    assert orc.Surf(m).intersection_free() == (True, 0)


def test_count_inverted():
    V, T = M.grid_tets(2, 2, 2)
    m = M.Mesh(V, T)
    assert orc.Elastic(m).count_inverted() == 0
    m.V = m.V_rest.copy()
    m.V[13] += np.array([0.9, 0.9, 0.9])  # drag the centre vertex through its neighbours
    x = m.V[m.T]
    det = np.linalg.det(np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0], x[:, 3] - x[:, 0]], axis=2))
    assert (det < 0).sum() > 0 and orc.Elastic(m).count_inverted() == int((det < 0).sum())
    m.mu[:] = 0.0  # kinematic parts (mu = lambda = 0) are not checked (Mesh.cpp:749)
    assert orc.Elastic(m).count_inverted() == 0
