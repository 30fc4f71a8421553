// MeshCOEncoding.hpp -- translation between the two MMCVID encodings of a mesh / obstacle pair (no dependencies: plain ints).
//
//   merged : the self-contact encoding over the merged vertex numbering the device uses -- obstacle vertex k is vertex nV + k
//   MeshCO : the reference's own encoding for MeshCO<3> (src/CollisionObject/MeshCO.cpp:83-120): negative entries -v-1 are MESH vertices,
//            non-negative entries are OBSTACLE vertices
//              EE (m0, m1, o0, o1)   PP (-m-1, o, -1, -mult)   PE (-m-1, o0, o1, -mult)   PT (-m-1, o0, o1, o2)
//              TP (-m0-1, -m1-1, -m2-1, o)   EP (-m0-1, -m1-1, o, -mult)
// Slot 3 keeps a multiplicity / mollifier marker (< 0) as it is.  Used by adapters/IpcGpuMeshCO.hpp; compiled and run against the oracle's
// translation (oracle/meshco.cpp: co_to_merged) by tests/test_adapters_compile.py.
#pragma once

namespace meshco_encoding {

// does a merged entry name an obstacle vertex (then it belongs to MeshCO's list, otherwise to the SelfCollisionHandler's)
inline bool touches_obstacle(const int* q, int nV)
{
    if ((q[0] < 0 ? -q[0] - 1 : q[0]) >= nV) return true;
    for (int k = 1; k < 4; ++k)
        if (q[k] >= nV) return true;
    return false;
}

inline void to_meshco(const int* q, int nV, int* m)
{
    if (q[0] >= 0) { // EE: the mesh edge comes first (its sorted edge index is the smaller one)
        m[0] = q[0]; m[1] = q[1]; m[2] = q[2] - nV; m[3] = q[3] >= 0 ? q[3] - nV : q[3];
        return;
    }
    const int p = -q[0] - 1;
    if (p < nV) { // PP / PE / PT: mesh point against obstacle vertex / edge / triangle
        m[0] = q[0]; m[1] = q[1] - nV; m[2] = q[2] >= 0 ? q[2] - nV : q[2]; m[3] = q[3] >= 0 ? q[3] - nV : q[3];
    }
    else if (q[3] < 0) { // EP: obstacle point against mesh edge (a point-point entry always names the mesh vertex first, so q[2] >= 0 here)
        m[0] = -q[1] - 1; m[1] = -q[2] - 1; m[2] = p - nV; m[3] = q[3];
    }
    else { // TP: obstacle point against mesh triangle
        m[0] = -q[1] - 1; m[1] = -q[2] - 1; m[2] = -q[3] - 1; m[3] = p - nV;
    }
}

inline void to_merged(const int* m, int nV, int* q)
{
    if (m[0] >= 0) { q[0] = m[0]; q[1] = m[1]; q[2] = nV + m[2]; q[3] = m[3] >= 0 ? nV + m[3] : m[3]; }
    else if (m[1] >= 0) { q[0] = m[0]; q[1] = nV + m[1]; q[2] = m[2] >= 0 ? nV + m[2] : m[2]; q[3] = m[3] >= 0 ? nV + m[3] : m[3]; }
    else if (m[2] < 0) { q[0] = -(nV + m[3]) - 1; q[1] = -m[0] - 1; q[2] = -m[1] - 1; q[3] = -m[2] - 1; }
    else { q[0] = -(nV + m[2]) - 1; q[1] = -m[0] - 1; q[2] = -m[1] - 1; q[3] = m[3]; }
}

} // namespace meshco_encoding
