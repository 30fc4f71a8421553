"""Host-side mirror of the obstacle hand-off (include/ipcgpu.h: ipcgpu_set_obstacle_tail) -- what adapters/IpcGpuAdapters.hpp does in C++.

A kinematic mesh obstacle (the reference's MeshCO<3>: Base::V, edges, Base::F, src/CollisionObject/MeshCO.hpp:39-233) rides at the tail of
the mesh's arrays; the library reports contact entries in the self-contact encoding over that merged numbering.  The functions here build
the merged arrays and translate entries between the merged form and MeshCO's own MMCVID encoding (MeshCO.cpp:83-120):

    EE (m0, m1, o0, o1)   PP (-m-1, o, -1, -mult)   PE (-m-1, o0, o1, -mult)   PT (-m-1, o0, o1, o2)
    TP (-m0-1, -m1-1, -m2-1, o)   EP (-m0-1, -m1-1, o, -mult)          (negative = mesh vertex, non-negative = obstacle vertex)
"""
import copy

import numpy as np


def with_obstacle(m, Vo, Eo, Fo):
    """The mesh `m` with the obstacle at the tail of its vertex and surface arrays (a shallow copy; `m` is not modified).
    Obstacle vertex k becomes vertex m.nV + k: rest = current position, Dirichlet flag 1, mass 0, codimension 3; EVERY obstacle vertex is a
    surface vertex (MeshCO's point loops run over Base::V.rows(), MeshCO.cpp:1899); edges and triangles are appended re-indexed.
    The copy carries nV_dof = m.nV (the argument of ipcgpu_set_obstacle_tail) and n_mesh_edges / n_mesh_tris."""
    Vo = np.ascontiguousarray(Vo, dtype=np.float64)
    Eo = np.asarray(Eo, dtype=np.int32).reshape(-1, 2)
    Fo = np.asarray(Fo, dtype=np.int32).reshape(-1, 3)
    mm = copy.copy(m)
    nV, nVo = m.nV, Vo.shape[0]
    mm.V_rest = np.concatenate([m.V_rest, Vo])
    mm.V = np.concatenate([m.V, Vo])
    mm.mass = np.concatenate([m.mass, np.zeros(nVo)])
    mm.dbc = np.concatenate([m.dbc, np.ones(nVo, dtype=np.uint8)]).astype(np.uint8)
    mm.vCoDim = np.concatenate([m.vCoDim, np.full(nVo, 3, dtype=np.int32)]).astype(np.int32)
    mm.SVI = np.concatenate([m.SVI, nV + np.arange(nVo, dtype=np.int32)]).astype(np.int32)
    mm.SFEdges = np.concatenate([m.SFEdges, Eo + nV]).astype(np.int32)
    mm.SF = np.concatenate([m.SF, Fo + nV]).astype(np.int32)
    mm.nV = nV + nVo
    mm.nV_dof, mm.n_mesh_edges, mm.n_mesh_tris, mm.n_mesh_sv = nV, len(m.SFEdges), len(m.SF), len(m.SVI)
    mm._nbr = None
    return mm


def pad_direction(p, nV_all):
    """search direction of the merged arrays: the obstacle does not move during a line search (MeshCO.cpp:790-800 passes Base::V twice)"""
    out = np.zeros(3 * nV_all)
    out[: len(p)] = np.asarray(p, dtype=np.float64).ravel()
    return out


def involves_obstacle(q, nV):
    """does a merged self-contact entry (or a mollified one) touch an obstacle vertex"""
    vs = [(-q[0] - 1) if q[0] < 0 else q[0]] + [int(x) for x in q[1:] if x >= 0]
    return any(v >= nV for v in vs)


def merged_to_meshco(q, nV):
    """one merged entry with an obstacle vertex -> MeshCO's encoding (slot 3 keeps a multiplicity / mollifier marker < 0 as it is)"""
    q = [int(x) for x in q]
    if q[0] >= 0:  # edge-edge: mesh edge first (its sorted edge index is the smaller one)
        assert q[0] < nV and q[1] < nV and q[2] >= nV
        return [q[0], q[1], q[2] - nV, q[3] - nV if q[3] >= 0 else q[3]]
    p = -q[0] - 1
    if p < nV:  # mesh point against obstacle vertex / edge / triangle
        assert q[1] >= nV
        return [q[0], q[1] - nV, q[2] - nV if q[2] >= 0 else q[2], q[3] - nV if q[3] >= 0 else q[3]]
    assert q[2] >= 0, "a point-point entry always names the mesh vertex first"
    if q[3] < 0:  # obstacle point against mesh edge
        return [-q[1] - 1, -q[2] - 1, p - nV, q[3]]
    return [-q[1] - 1, -q[2] - 1, -q[3] - 1, p - nV]  # obstacle point against mesh triangle


def split_sets(mm, pa, pe, nV, n_mesh_edges):
    """(active, mollified, mollified edge pairs) of the merged run -> the SelfCollisionHandler's lists and MeshCO's lists"""
    mm, pa, pe = (np.asarray(x, dtype=np.int64).reshape(-1, k) for x, k in ((mm, 4), (pa, 4), (pe, 2)))
    co = np.array([involves_obstacle(q, nV) for q in mm], dtype=bool)
    self_mm = mm[~co].astype(np.int32)
    co_mm = np.array([merged_to_meshco(q, nV) for q in mm[co]], dtype=np.int32).reshape(-1, 4)
    pco = np.array([involves_obstacle(q, nV) or (e[1] >= n_mesh_edges) for q, e in zip(pa, pe)], dtype=bool) if len(pa) else np.zeros(0, dtype=bool)
    self_pa, self_pe = pa[~pco].astype(np.int32), pe[~pco].astype(np.int32)
    co_pa = np.array([merged_to_meshco(q, nV) for q in pa[pco]], dtype=np.int32).reshape(-1, 4)
    co_pe = np.array([[e[0], e[1] - n_mesh_edges if e[1] >= 0 else -1] for e in pe[pco]], dtype=np.int32).reshape(-1, 2)
    return (self_mm, self_pa, self_pe), (co_mm, co_pa, co_pe)


def split_candidates(cand, n_mesh_sv, n_mesh_tris, n_mesh_edges):
    """cs_PTEE of the merged run -> (self-contact candidates, MeshCO candidates in MeshCO's form: PT (-svI-1, sfI obstacle),
    TP (-sfI-1, -vI-1), EE (eI mesh, eJ obstacle); MeshCO.cpp:2144-2161)"""
    self_c, co_c = [], []
    for a, b in np.asarray(cand, dtype=np.int64).reshape(-1, 2):
        if a < 0:
            sv, sf = -a - 1, b
            if sv < n_mesh_sv and sf < n_mesh_tris:
                self_c.append((a, b))
            elif sv < n_mesh_sv:
                co_c.append((a, sf - n_mesh_tris))
            else:
                co_c.append((-sf - 1, -(sv - n_mesh_sv) - 1))
        elif b < n_mesh_edges:
            self_c.append((a, b))
        else:
            co_c.append((a, b - n_mesh_edges))
    return np.array(self_c, dtype=np.int32).reshape(-1, 2), np.array(co_c, dtype=np.int32).reshape(-1, 2)
