"""CPU: the MeshCO restatement (oracle/meshco.cpp) checked against itself and against the independent self-contact restatement.

A kinematic obstacle is, mathematically, a Dirichlet body of the merged surface whose rows never reach the system.  MeshCO.cpp restates
that with its own loops, encoding and counting rules; the tests below run the MeshCO oracle next to the self-contact oracle on the merged
surface and require the same pairs, energies, gradients, Hessians and step bounds -- plus finite differences and the host-side translation
of ipc_b200/obstacle.py (the mirror of the C++ adapter)."""
import numpy as np
import pytest

import oracle as orc
from ipc_b200 import obstacle as OB
from ipc_b200 import scenes

KAPPA = 1e8
NTH = 8


def scene(plate_angle=0.37, **kw):
    m, info = scenes.balls_on_obstacle(plate_angle=plate_angle, **kw)
    ob = info["obstacle"]
    s = orc.Surf(m)
    return m, info, ob, s, orc.Obstacle(s, ob["V"], ob["E"], ob["F"])


def canonical_cross(mm, nV):
    """what the library does to the merged self-contact list: an obstacle point against a mesh vertex is renamed mesh-vertex-first and merged
    with its twin (MeshCO.cpp:1831, :1919, :2168-2190)"""
    out = {}
    for q in mm:
        q = [int(x) for x in q]
        if q[0] < 0 and q[2] < 0 and (-q[0] - 1) >= nV:
            q = [-q[1] - 1, -q[0] - 1, -1, q[3]]
        key = tuple(q[:3]) if (q[0] < 0 and q[3] < 0) else tuple(q)
        if q[0] < 0 and q[3] < 0:
            out[key] = out.get(key, 0) + (-q[3])
        else:
            out[key] = None
    rows = [list(k) + [-v] if v is not None and len(k) == 3 else list(k) for k, v in out.items()]
    return np.array(sorted(rows), dtype=np.int32).reshape(-1, 4)


def contact_pairs(mm, pa, pe, SFEdges, nV_dof):
    pairs = []
    for r in list(mm) + list(pa):
        vs = [(-r[0] - 1) if r[0] < 0 else r[0]] + [x for x in r[1:] if x >= 0]
        pairs += [(a, b) for a in vs for b in vs if a < b and b < nV_dof]
    for e in pe:
        if e[0] >= 0:
            vs = list(SFEdges[e[0]]) + list(SFEdges[e[1]])
            pairs += [(a, b) for a in vs for b in vs if a != b and a < nV_dof and b < nV_dof]
    return pairs


@pytest.mark.parametrize("angle", [0.37, 0.0])
def test_sets_equal_the_self_contact_sets_of_the_merged_surface(angle):
    m, info, ob, s, o = scene(angle)
    dHat = info["dHat"]
    mm, pa, pe, cand = o.constraint_set(dHat, NTH)
    assert len(mm) > 10 and len(cand) >= len(mm)
    kinds = set()
    for q in mm:
        kinds.add("EE" if q[0] >= 0 else (("PP" if q[2] < 0 else ("PE" if q[3] < 0 else "PT")) if q[1] >= 0 else ("TP" if q[2] < 0 else "EP")))
    assert {"EE", "PT", "TP"} <= kinds
    if angle == 0.0:
        assert len(pa) > 0  # the plate's edges are parallel to the flat pole's: mollified entries
    # the self-contact oracle on the merged surface finds the union of the mesh's own pairs and the cross pairs
    M2 = OB.with_obstacle(m, ob["V"], ob["E"], ob["F"])
    s2 = orc.Surf(M2)
    mm2, pa2, pe2, cand2 = s2.constraint_set(dHat, NTH)
    mm2 = canonical_cross(mm2, m.nV)
    (smm, spa, spe), (cmm, cpa, cpe) = OB.split_sets(mm2, pa2, pe2, m.nV, len(m.SFEdges))
    mm_s, pa_s, pe_s, cand_s = s.constraint_set(dHat, NTH)
    assert np.array_equal(smm, mm_s) and np.array_equal(spa, pa_s) and np.array_equal(spe, pe_s)
    order = np.lexsort(cmm.T[::-1])
    assert np.array_equal(cmm[order], mm)
    if len(pa):
        k = np.lexsort(np.concatenate([cpa, cpe], axis=1).T[::-1])
        assert np.array_equal(cpa[k], pa) and np.array_equal(cpe[k], pe)
    sc, cc = OB.split_candidates(cand2, len(m.SVI), len(m.SF), len(m.SFEdges))
    assert np.array_equal(sc[np.lexsort(sc.T[::-1])], cand_s)
    assert np.array_equal(cc[np.lexsort(cc.T[::-1])], cand)
    # translation round trip: oracle's MeshCO -> merged equals the inverse of obstacle.merged_to_meshco
    merged, pe_m = o.to_merged(mm, pe)
    back = np.array([OB.merged_to_meshco(q, m.nV) for q in merged], dtype=np.int32)
    assert np.array_equal(back, mm)


def test_energy_gradient_hessian_against_the_merged_self_contact_oracle_and_finite_differences():
    m, info, ob, s, o = scene(0.0)
    dHat = info["dHat"]
    mm, pa, pe, _ = o.constraint_set(dHat, NTH)
    E, bad = o.energy(mm, pa, pe, dHat, KAPPA)
    assert not bad and E > 0
    g = o.gradient(mm, pa, pe, dHat, KAPPA)
    assert np.abs(g).max() > 0
    # merged self-contact oracle with the translated entries
    M2 = OB.with_obstacle(m, ob["V"], ob["E"], ob["F"])
    s2 = orc.Surf(M2)
    mg, pg = o.to_merged(mm, pe)
    pag, _ = o.to_merged(pa, pe)
    E2, _ = s2.barrier_energy(mg, pag, pg, dHat, KAPPA)
    assert abs(E - E2) <= 1e-14 * abs(E2)
    g2 = s2.barrier_gradient(mg, pag, pg, dHat, KAPPA)
    assert np.array_equal(g, g2[: 3 * m.nV])
    # finite differences of the energy over the fixed set (mesh vertices only; the obstacle has no degrees of freedom)
    rng = np.random.default_rng(0)
    touched = np.nonzero(np.abs(g.reshape(-1, 3)).sum(1))[0]
    h = 1e-7 * np.sqrt(dHat)
    for v in rng.choice(touched, 6, replace=False):
        for c in range(3):
            Vp, Vm = m.V.copy(), m.V.copy()
            Vp[v, c] += h
            Vm[v, c] -= h
            Ep = orc.Obstacle(orc.Surf(m, Vp), ob["V"], ob["E"], ob["F"]).energy(mm, pa, pe, dHat, KAPPA)[0]
            Em = orc.Obstacle(orc.Surf(m, Vm), ob["V"], ob["E"], ob["F"]).energy(mm, pa, pe, dHat, KAPPA)[0]
            fd = (Ep - Em) / (2 * h)
            assert abs(fd - g[3 * v + c]) <= 2e-5 * max(abs(g[3 * v + c]), np.abs(g).max() * 1e-3), (v, c, fd, g[3 * v + c])
    # Hessian: the mesh's own pattern (no column of the obstacle); equals the merged oracle's leading block, symmetric positive semi-definite
    ia, ja = m.csr_pattern(1, extra_pairs=contact_pairs(mg, pag, pg, M2.SFEdges, m.nV))
    a = o.hessian_csr(mm, pa, pe, dHat, KAPPA, ia, ja, 1, 1, nthreads=NTH)
    ia2, ja2 = M2.csr_pattern(1, extra_pairs=contact_pairs(mg, pag, pg, M2.SFEdges, m.nV))
    a2 = s2.barrier_hessian_csr(mg, pag, pg, dHat, KAPPA, ia2, ja2, 1, 1, nthreads=NTH)
    assert len(a) == ia[3 * m.nV] - 1 and np.array_equal(a, a2[: len(a)])  # the mesh's values are a prefix of the merged value array
    assert np.abs(a).max() > 0
    import scipy.sparse as sp
    U = sp.csr_matrix((a, ja - 1, ia - 1), shape=(3 * m.nV, 3 * m.nV))
    H = (U + sp.triu(U, 1).T).toarray()
    rows = np.nonzero(np.abs(H).sum(1))[0]
    w = np.linalg.eigvalsh(H[np.ix_(rows, rows)])
    assert w.min() >= -1e-9 * w.max()


def test_step_bounds_against_the_merged_self_contact_oracle():
    m, info, ob, s, o = scene(0.37)
    dHat, p = info["dHat"], info["p"]
    mm, pa, pe, cand = o.constraint_set(dHat, NTH)
    evf, eee = orc.ti_error(s.V, m.nV, None)
    # with the edge-edge routine the obstacle pairs are ordinary pairs of the merged surface: min(self, obstacle) bit for bit
    a_part, z = o.ccd_partial(p, cand, 1e-6, evf, eee, 1.0, ee_as_vf=0, nthreads=NTH)
    assert not z and 0.0 < a_part < 1.0
    a_full, z, npairs = o.ccd_full(p, 1e-6, evf, eee, 1.0, ee_as_vf=0, nthreads=NTH)
    assert not z and npairs >= len(cand) and a_full <= a_part
    M2 = OB.with_obstacle(m, ob["V"], ob["E"], ob["F"])
    s2 = orc.Surf(M2)
    p2 = OB.pad_direction(p, M2.nV)
    _, _, _, cand2 = s2.constraint_set(dHat, NTH)
    a2_part, _ = orc.ccd_partial(s2, p2, cand2, 1e-6, evf, eee, 1.0, NTH)
    _, _, _, cand_s = s.constraint_set(dHat, NTH)
    a_self, _ = orc.ccd_partial(s, p, cand_s, 1e-6, evf, eee, 1.0, NTH)
    assert a2_part == min(a_self, a_part)
    g = orc.grid_swept(s2, p2, 1.0, m.avgEdgeLen / 3.0)
    a2_full, _, _ = orc.ccd_full(s2, p2, g[0], g[1], 1e-6, evf, eee, g[1], nthreads=NTH)
    gs = orc.grid_swept(s, p, 1.0, m.avgEdgeLen / 3.0)
    a_self_full, _, _ = orc.ccd_full(s, p, gs[0], gs[1], 1e-6, evf, eee, gs[1], nthreads=NTH)
    # (the swept grid rescales the step to its own alpha first; compare at the common entry step)
    a_co_full, _, _ = o.ccd_full(p, 1e-6, evf, eee, g[1], ee_as_vf=0, nthreads=NTH)
    assert a2_full == min(a_self_full, a_co_full) or g[1] != gs[1]
    # no stencil of a candidate pair degenerates at 0.999 of the bound
    V2 = m.V + 0.999 * a_full * p.reshape(-1, 3)
    for c in cand:
        if c[0] < 0 and c[1] < 0:
            x = np.concatenate([ob["V"][[-c[1] - 1]], V2[m.SF[-c[0] - 1]]])
            assert orc.point_tri_d(x) > 0
        elif c[0] < 0:
            x = np.concatenate([V2[[m.SVI[-c[0] - 1]]], ob["V"][ob["F"][c[1]]]])
            assert orc.point_tri_d(x) > 0
        else:
            x = np.concatenate([V2[m.SFEdges[c[0]]], ob["V"][ob["E"][c[1]]]])
            assert orc.edge_edge_d(x) > 0
    # the reference's call (vertex-face routine on edge pairs) is a different function of the same inputs: it still returns a step in (0, 1]
    a_ref, z = o.ccd_partial(p, cand, 1e-6, evf, eee, 1.0, ee_as_vf=1, nthreads=NTH)
    assert not z and 0.0 < a_ref <= 1.0
    # zero initial distance -> step 0 (MeshCO.cpp:771-779)
    Vz = ob["V"].copy()
    c = next(c for c in cand if c[0] < 0 and c[1] >= 0)
    Vz[ob["F"][c[1]][0]] = m.V[m.SVI[-c[0] - 1]]
    oz = orc.Obstacle(s, Vz, ob["E"], ob["F"])
    a_z, z = oz.ccd_partial(p, cand, 1e-6, evf, eee, 1.0, nthreads=NTH)
    assert z and a_z == 0.0


def test_encoding_translation_of_every_kind_and_empty_lists():
    """ipc_b200/obstacle.py (host mirror of adapters/IpcGpuMeshCO.hpp) against the oracle's translation, on hand-built entries of all six kinds"""
    import ctypes as C
    nV, nSE = 100, 40
    co = np.array([[3, 7, 5, 9],          # EE  (mesh edge 3-7, obstacle edge 5-9)
                   [-12, 4, -1, -3],      # PP  (mesh vertex 11, obstacle vertex 4, found three times)
                   [-12, 4, 6, -2],       # PE  (mesh vertex 11, obstacle edge 4-6, twice)
                   [-12, 4, 6, 8],        # PT
                   [-2, -5, -9, 17],      # TP  (obstacle vertex 17, mesh triangle 1-4-8)
                   [-2, -5, 17, -1]],     # EP  (obstacle vertex 17, mesh edge 1-4)
                  dtype=np.int32)
    pe = np.array([[2, 5], [-1, -1]], dtype=np.int32)
    out = np.empty_like(co)
    pe_out = np.empty_like(pe)
    orc.lib().orc_meshco_to_merged(nV, nSE, orc.i(co), len(co), orc.i(out), orc.i(pe), len(pe), orc.i(pe_out))
    expect = np.array([[3, 7, 105, 109], [-12, 104, -1, -3], [-12, 104, 106, -2], [-12, 104, 106, 108], [-118, 1, 4, 8], [-118, 1, 4, -1]], dtype=np.int32)
    assert np.array_equal(out, expect) and np.array_equal(pe_out, [[2, 45], [-1, -1]])
    assert all(OB.involves_obstacle(q, nV) for q in out)
    assert np.array_equal(np.array([OB.merged_to_meshco(q, nV) for q in out], dtype=np.int32), co)
    assert not OB.involves_obstacle([-12, 13, 14, 15], nV) and not OB.involves_obstacle([1, 2, 3, 4], nV)
    # splitting: the mesh's own entries stay as they are, the obstacle's are translated; mollified pairs by their edge pair as well
    mm = np.concatenate([out, [[-12, 13, 14, 15], [1, 2, 3, 4]]]).astype(np.int32)
    pa = np.array([[-4, 105, -1, -1], [5, 6, 7, 8]], dtype=np.int32)
    pe2 = np.array([[2, 45], [-1, -1]], dtype=np.int32)
    (smm, spa, spe), (cmm, cpa, cpe) = OB.split_sets(mm, pa, pe2, nV, nSE)
    assert np.array_equal(smm, [[-12, 13, 14, 15], [1, 2, 3, 4]]) and np.array_equal(cmm, co)
    assert np.array_equal(spa, [[5, 6, 7, 8]]) and np.array_equal(spe, [[-1, -1]]) and np.array_equal(cpa, [[-4, 5, -1, -1]]) and np.array_equal(cpe, [[2, 5]])
    sc, cc = OB.split_candidates([[-3, 4], [-3, 64], [-53, 4], [7, 9], [7, 49]], n_mesh_sv=50, n_mesh_tris=60, n_mesh_edges=40)
    assert np.array_equal(sc, [[-3, 4], [7, 9]]) and np.array_equal(cc, [[-3, 4], [-5, -3], [7, 9]])
    e = np.zeros((0, 4), np.int32)
    (a, b, c), (d, f, g) = OB.split_sets(e, e, np.zeros((0, 2), np.int32), nV, nSE)
    assert all(len(x) == 0 for x in (a, b, c, d, f, g))
    # an obstacle out of reach: empty sets, the step stays what it was
    m, info, ob, s, o = scene(0.37)
    far = orc.Obstacle(s, ob["V"] + np.array([0.0, 0.0, -10.0]), ob["E"], ob["F"])
    mm0, pa0, pe0, cand0 = far.constraint_set(info["dHat"], NTH)
    assert len(mm0) == len(pa0) == len(cand0) == 0
    evf, eee = orc.ti_error(s.V, m.nV, None)
    assert far.ccd_partial(info["p"], cand0, 1e-6, evf, eee, 0.7, nthreads=NTH) == (0.7, 0)
    a_full, z, npairs = far.ccd_full(info["p"], 1e-6, evf, eee, 0.7, nthreads=NTH)
    assert (a_full, z, npairs) == (0.7, 0, 0)
    assert far.energy(mm0, pa0, pe0, info["dHat"], KAPPA) == (0.0, 0)


def test_row_owner_partition_with_an_obstacle_tail():
    """ipc_b200/partition.py (mirror of the library's rule) on a mesh with an obstacle at the tail: the tail's vertices touch no tetrahedron, so they
    all fall to the last rank, every tetrahedron is still assembled, and the solver's own values (the CSR prefix) split between the ranks as before"""
    from ipc_b200 import partition as P
    m, info, ob, s, o = scene(0.37)
    M2 = OB.with_obstacle(m, ob["V"], ob["E"], ob["F"])
    ia, ja = M2.csr_pattern(1)
    ia0, _ = m.csr_pattern(1)
    for world in (2, 3, 8):
        b = P.vertex_boundaries(M2.T, M2.nV, world)
        b0 = P.vertex_boundaries(m.T, m.nV, world)
        assert b[:-1] == b0[:-1] and b[-1] == M2.nV and b[-2] <= m.nV  # same cuts; only the last range grows by the tail
        seen = np.zeros(m.nT, dtype=int)
        for r in range(world):
            seen[P.assembled_tets(M2.T, b[r], b[r + 1])] += 1
        assert seen.min() >= 1
        lo, hi = P.owned_value_range(ia, 1, b[world - 1], b[world])
        lo0, hi0 = P.owned_value_range(ia0, 1, b0[world - 1], b0[world])
        assert lo == lo0 and hi == hi0 + 6 * len(ob["V"])  # the tail's identity blocks: 3 + 2 + 1 values per vertex, after the mesh's values
