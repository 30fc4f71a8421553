// broadphase.cuh -- sort-based uniform grid shared by the constraint-set build and the CCD broad phase.
//
// Replaces SpatialHash<3> (src/Utils/SpatialHash.hpp:46-201 static build, :589-750 swept build, and the queries
// :203-229, :375-421, :752-832), whose std::unordered_map<int, std::vector<int>> and serial inserts do not map to a GPU.
// Design: every edge / triangle registers its AABB in at most 8 cells of a grid whose cell edge is >= the largest
// (inflated) primitive extent of this frame, the (cell,id) entries are radix-sorted by cell (CUB), and a query walks
// the <= 8 cells of its own box with a binary search per cell.  A pair is reported exactly once, from the
// lexicographically smallest common cell of the two boxes.  The grid is only an accelerator: what decides membership
// is the exact test applied afterwards (d < dHat for the constraint set; the reference's own voxel-AABB overlap for CCD
// candidates), so results do not depend on the cell size chosen here.
#pragma once
#include "common.cuh"
#include "contact.cuh"
#include "broadphase_types.h"

namespace ipcgpu {


// monotone map double -> uint64 (handles negatives) for atomicMin/Max on coordinates
DEV unsigned long long flip_ord(double x)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
DEV double unflip_ord(unsigned long long u)
{
    u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)u);
}

DEV void cell_range(const Grid& g, const Box& b, int* c0, int* c1)
{
    const double o[3] = { g.ox, g.oy, g.oz };
    const int n[3] = { g.nx, g.ny, g.nz };
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int lo = (int)floor((b.lo[a] - o[a]) * g.inv_h), hi = (int)floor((b.hi[a] - o[a]) * g.inv_h);
        c0[a] = min(max(lo, 0), n[a] - 1);
        c1[a] = min(max(hi, 0), n[a] - 1);
    }
}
DEV unsigned long long cell_key(const Grid& g, int ix, int iy, int iz) { return ((unsigned long long)iz * g.ny + iy) * g.nx + ix; }

DEV int lower_bound_u64(const unsigned long long* __restrict__ keys, int n, unsigned long long k)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (keys[mid] < k) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

} // namespace ipcgpu
