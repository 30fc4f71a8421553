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

DEV int upper_bound_u64(const unsigned long long* __restrict__ keys, int n, unsigned long long k)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (keys[mid] <= k) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Grid v2: every primitive is registered ONCE, in the cell of its box's lower corner.  With cell edge >= (largest box extent +
// 2*radius), a box starting in cell c ends in c or c+1, and a query box inflated by `radius` covers cells q0..q1 (q1 <= q0+1), so
// every partner starts in [q0-1, q1] per axis: at most 3x3 rows of <= 3 consecutive cells.  Entries are sorted by cell key; an
// open-addressing table (cell key -> index of the first entry of that cell) replaces binary searches: the <= 27 cell lookups of a
// query run on 27 lanes at once.
struct SortedGrid {
    const unsigned long long* keys; // sorted cell keys
    const int* ids;                 // primitive id per entry
    const Box* boxes;               // primitive boxes gathered in sorted order (coalesced candidate scan)
    int n;
    const unsigned* tab_key;        // hash table: cell key (0xffffffff = empty)
    const int* tab_start;           //             first entry of that cell in the sorted arrays
    unsigned tab_mask;              // table size - 1 (power of two)
};

DEV unsigned cell_hash(unsigned key) { return key * 2654435761u; }
DEV int cell_lookup(const SortedGrid& sg, unsigned key)
{
    unsigned h = cell_hash(key) & sg.tab_mask;
    for (;;) {
        const unsigned k = sg.tab_key[h];
        if (k == key) return sg.tab_start[h];
        if (k == 0xffffffffu) return -1;
        h = (h + 1) & sg.tab_mask;
    }
}

DEV bool boxes_overlap(const Box& a, const Box& b)
{
    return !(a.lo[0] > b.hi[0] || b.lo[0] > a.hi[0] || a.lo[1] > b.hi[1] || b.lo[1] > a.hi[1] || a.lo[2] > b.hi[2] || b.lo[2] > a.hi[2]);
}

// warp-cooperative scan: calls f(id, box) on every registered primitive whose box overlaps the (already inflated) query box.
// All 32 lanes must call this together; f runs on the lane that found the candidate.
template <typename F>
DEV void warp_scan_candidates(const Grid& g, const SortedGrid& sg, const Box& qb, int lane, F f)
{
    int c0[3], c1[3];
    cell_range(g, qb, c0, c1);
    const int x0 = max(c0[0] - 1, 0), x1 = c1[0];
    const int y0 = max(c0[1] - 1, 0), z0 = max(c0[2] - 1, 0);
    const int ny = c1[1] - y0 + 1, nz = c1[2] - z0 + 1; // <= 3 each
    const int nrows = ny * nz;
    // lane 3r+dx looks up cell (x0+dx) of row r
    int mine = -1;
    {
        const int r = lane / 3, dx = lane - 3 * r;
        if (r < nrows && x0 + dx <= x1) mine = cell_lookup(sg, (unsigned)cell_key(g, x0 + dx, y0 + r % ny, z0 + r / ny));
    }
    for (int r = 0; r < nrows; ++r) {
        const int s0 = __shfl_sync(0xffffffffu, mine, 3 * r), s1 = __shfl_sync(0xffffffffu, mine, 3 * r + 1), s2 = __shfl_sync(0xffffffffu, mine, 3 * r + 2);
        const int start = s0 >= 0 ? s0 : (s1 >= 0 ? s1 : s2); // cells of one row are consecutive keys => consecutive runs
        if (start < 0) continue;
        const unsigned long long key_hi = cell_key(g, x1, y0 + r % ny, z0 + r / ny);
        for (int k = start + lane;; k += 32) {
            const bool in = k < sg.n && sg.keys[k] <= key_hi;
            if (in) {
                const Box b = sg.boxes[k];
                if (boxes_overlap(qb, b)) f(sg.ids[k], b);
            }
            if (!__any_sync(0xffffffffu, in)) break;
        }
    }
}

} // namespace ipcgpu
