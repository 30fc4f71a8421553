// ccd.cu -- CCD line-search step bound: swept broad phase + Tight-Inclusion narrow phase + device-wide min (sm_100a).
//   *** this translation unit is compiled with --fmad=false ***  (bit-exact step bound: every product/sum is rounded
//   exactly like the CPU oracle built with -ffp-contract=off; see __graft_entry__.py NOFMA_FILES)
//
// Reference being replaced:
//   SpatialHash<3>::build(mesh, searchDir, curMaxStepSize, voxelSize)          src/Utils/SpatialHash.hpp:589-750
//   SpatialHash queries  queryPointForPrimitives / queryEdgeForEdgesWithBBoxCheck  :752-773, :803-832
//   SelfCollisionHandler::largestFeasibleStepSize_TightInclusion (partial)      src/CollisionObject/SelfCollisionHandler.cpp:690-866
//   SelfCollisionHandler::largestFeasibleStepSize_CCD_TightInclusion (full)     :1370-1630
//   inclusion_ccd::vertexFaceCCD_double / edgeEdgeCCD_double (un-vendored Tight-Inclusion; restated from the published algorithm)
//
// Structure
//   broad phase : the same sort-based coarse grid as the constraint set, over SWEPT boxes; a pair becomes a candidate iff
//                 the reference's own voxel ranges (cell = avgEdgeLen/3, origin = swept bbox corner) overlap -- i.e. exactly
//                 the pairs the reference's hash query returns -- plus its swept-AABB test for edge pairs.
//   narrow, stage 1 (thread per candidate, HBM-bound: 8 B pair + 4 x 48 B gather): current distance, ms, and the ROOT box
//                 of the interval search; a candidate whose root box excludes the origin cannot collide and dies here.
//   narrow, stage 2 (warp per surviving pair, persistent CTAs + atomic work counter): level-synchronous breadth-first
//                 interval bisection.  The library's (level, t_lo) priority order is realised without sorting: per level,
//                 two warp min-reductions over the lexicographic key (t_lo,u_lo,v_lo) give the first box containing the
//                 origin and the first "terminal" box, which is all the sequential semantics depend on.
//   reduction   : atomicMin on the order-preserving uint64 image of the (non-negative) time of impact.
#include "broadphase.cuh"
#include "context.h"
#include "../../include/ipcgpu.h"
#include <cub/cub.cuh>

namespace ipcgpu {

// ------------------------------------------------------------------------------------------------------------------
// reference voxel ranges of every surface vertex on the swept grid (SpatialHash.hpp:642-662, :841-845)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ref_ranges(SurfArgs s, const double* __restrict__ dir, const IterState* __restrict__ st, int* __restrict__ vmin, int* __restrict__ vmax)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.nSV) return;
    const double alpha = st->alpha_grid, inv_h = st->ref_inv_h;
    const int v = s.SVI[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double x = s.V[(size_t)c * s.nV + v];
        const double xt = x + alpha * dir[3 * (size_t)v + c];
        const int a = (int)floor((x - st->ref_lo[c]) * inv_h), b = (int)floor((xt - st->ref_lo[c]) * inv_h);
        vmin[3 * (size_t)v + c] = min(a, b);
        vmax[3 * (size_t)v + c] = max(a, b);
    }
}

// bbox of all vertices (V) and of the displaced surface vertices: bounds[0..2] min, [3..5] max (flipped-order uint64)
__global__ void __launch_bounds__(256) k_swept_bounds(SurfArgs s, const double* __restrict__ dir, const IterState* __restrict__ st, unsigned long long* __restrict__ bounds)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double alpha = st->alpha_grid;
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    if (i < s.nV) {
#pragma unroll
        for (int c = 0; c < 3; ++c) lo[c] = hi[c] = s.V[(size_t)c * s.nV + i];
    }
    if (i < s.nSV) {
        const int v = s.SVI[i];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double xt = s.V[(size_t)c * s.nV + v] + alpha * dir[3 * (size_t)v + c];
            lo[c] = fmin(lo[c], xt);
            hi[c] = fmax(hi[c], xt);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double a = lo[c], b = hi[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a = fmin(a, __shfl_xor_sync(0xffffffffu, a, o));
            b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(bounds + c, flip_ord(a));
            atomicMax(bounds + 3 + c, flip_ord(b));
        }
    }
}

// SpatialHash.hpp:603-618 on the device-resident step: spanSize = alpha * mean|p| / h ; if (spanSize > 1) alpha /= spanSize.
// (pSize is the reference's serial host sum over the surface vertices, computed when the search direction is uploaded.)
__global__ void k_swept_alpha(IterState* st, const double* __restrict__ pSize_ptr, double h, unsigned long long* __restrict__ bounds)
{
    if (threadIdx.x != 0) return;
    const double pSize = *pSize_ptr;
    double alpha = ord_to_dbl(st->step_ord);
    const double span = alpha * pSize / h;
    if (span > 1) alpha /= span;
    st->step_ord = dbl_to_ord(alpha);
    st->alpha_grid = alpha;
    st->alpha_stage[2] = alpha;
    for (int c = 0; c < 3; ++c) { bounds[c] = ~0ull; bounds[3 + c] = 0ull; }
}
// reference grid geometry from the swept bbox (SpatialHash.hpp:627-640), incl. the cast-overflow fallback (:632-636)
__global__ void k_refgrid_params(IterState* st, const unsigned long long* __restrict__ bounds, double h)
{
    if (threadIdx.x != 0) return;
    double lo[3], hi[3], rmax = 0.0;
    double inv_h = 1.0 / h;
    bool bad = false;
    for (int c = 0; c < 3; ++c) {
        lo[c] = unflip_ord(bounds[c]);
        hi[c] = unflip_ord(bounds[3 + c]);
        st->ref_lo[c] = lo[c];
        st->ref_count[c] = (int)ceil((hi[c] - lo[c]) * inv_h);
        rmax = fmax(rmax, hi[c] - lo[c]);
        if (st->ref_count[c] <= 0) bad = true;
    }
    if (bad) { // cast overflow due to a huge search direction
        inv_h = 1.0 / (rmax * 1.01);
        st->ref_count[0] = st->ref_count[1] = st->ref_count[2] = 1;
    }
    st->ref_inv_h = inv_h;
    st->radius = 1.0 / inv_h; // pairs sharing a reference voxel are at most one voxel apart per axis
}

// ------------------------------------------------------------------------------------------------------------------
// broad-phase queries on the coarse grid with the reference's exact candidate filter
// ------------------------------------------------------------------------------------------------------------------
DEV bool ranges_overlap(const int* alo, const int* ahi, const int* blo, const int* bhi)
{
    return !(blo[0] > ahi[0] || bhi[0] < alo[0] || blo[1] > ahi[1] || bhi[1] < alo[1] || blo[2] > ahi[2] || bhi[2] < alo[2]);
}
DEV void prim_range(const int* __restrict__ vmin, const int* __restrict__ vmax, const int* vs, int n, int* lo, int* hi)
{
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        lo[c] = vmin[3 * (size_t)vs[0] + c];
        hi[c] = vmax[3 * (size_t)vs[0] + c];
        for (int k = 1; k < n; ++k) {
            lo[c] = min(lo[c], vmin[3 * (size_t)vs[k] + c]);
            hi[c] = max(hi[c], vmax[3 * (size_t)vs[k] + c]);
        }
    }
}
DEV bool dbc_v(const SurfArgs& s, int v) { return s.dbc && s.dbc[v] != 0; }
DEV int cod_v(const SurfArgs& s, int v) { return s.vCoDim ? s.vCoDim[v] : 3; }

struct CandOut {
    int2* cand;
    unsigned long long* n;
    unsigned long long cap;
    int* overflow;
};
DEV void push_cand(const CandOut& o, int2 c)
{
    // aggregated over the lanes that are active here: one atomic per warp
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(o.n, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    const unsigned long long i = base + __popc(m & ((1u << lane) - 1u));
    if (i < o.cap) o.cand[i] = c;
    else atomicExch(o.overflow, 1);
}

// ---- phase 1: one WARP per query primitive, boxes only: (query, partner) pairs whose swept boxes are within one reference voxel
__global__ void __launch_bounds__(256) k_ccd_pairs_pt(const Grid* __restrict__ gp, const Box* __restrict__ vboxes, SortedGrid tg, const IterState* __restrict__ st, int first, int last, PairOut out)
{
    __shared__ PairStage stage;
    pair_stage_init(stage);
    const double radius = st->radius;
    const int lane = threadIdx.x & 31;
    const Grid g = *gp;
    const int q0 = first + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kPairQueriesPerWarp;
    for (int svI = q0; svI < min(q0 + kPairQueriesPerWarp, last); ++svI) {
        Box qb = vboxes[svI];
        for (int k = 0; k < 3; ++k) { qb.lo[k] -= radius; qb.hi[k] += radius; }
        warp_scan_candidates(g, tg, qb, lane, [&](bool hit, int sfI) { warp_push_pair(stage, out, hit, svI, sfI, lane); });
    }
    pair_stage_flush(stage, out);
}
// queries are the entries of the sorted swept-edge grid itself ([first, last) = sorted positions); each walks only the entries behind it
__global__ void __launch_bounds__(256) k_ccd_pairs_ee(const Grid* __restrict__ gp, SortedGrid eg, const Box* __restrict__ eboxes, const IterState* __restrict__ st, int first, int last, PairOut out)
{
    __shared__ PairStage stage;
    pair_stage_init(stage);
    const double radius = st->radius;
    const int lane = threadIdx.x & 31;
    const Grid g = *gp;
    const int q0 = first + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kPairQueriesPerWarp;
    for (int i = q0; i < min(q0 + kPairQueriesPerWarp, last); ++i) {
        const int eI = eg.ids[i];
        Box qb = eboxes[eI];
        for (int k = 0; k < 3; ++k) { qb.lo[k] -= radius; qb.hi[k] += radius; }
        // (the exact swept-AABB test of queryEdgeForEdgesWithBBoxCheck is applied by k_ccd_filter_ee on the double boxes)
        warp_scan_candidates(g, eg, qb, lane, [&](bool hit, int eJ) { warp_push_pair(stage, out, hit, min(eI, eJ), max(eI, eJ), lane); }, i);
    }
    pair_stage_flush(stage, out);
}
// ---- phase 2: one THREAD per pair: the reference's voxel-range overlap (its hash query) and the index filters
__global__ void __launch_bounds__(256) k_ccd_filter_pt(SurfArgs s, const int2* __restrict__ pairs, const unsigned* __restrict__ nPairs, unsigned cap, const int* __restrict__ vmin,
    const int* __restrict__ vmax, CandOut out)
{
    const unsigned nP = min(*nPairs, cap);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nP; i += gridDim.x * blockDim.x) { // grid-stride over the device-resident count
    const int svI = pairs[i].x, sfI = pairs[i].y;
    const int vI = s.SVI[svI];
    int tv[3] = { s.SF[sfI], s.SF[(size_t)s.nSF + sfI], s.SF[(size_t)2 * s.nSF + sfI] };
    if (vI == tv[0] || vI == tv[1] || vI == tv[2]) continue;
    int lo[3], hi[3];
    prim_range(vmin, vmax, tv, 3, lo, hi);
    if (!ranges_overlap(vmin + 3 * (size_t)vI, vmax + 3 * (size_t)vI, lo, hi)) continue; // the reference's hash would not pair them
    const bool oP = obstacle_vertex(s, vI), oT = obstacle_vertex(s, tv[0]); // mesh-obstacle pairs are not filtered (MeshCO.cpp:1396-1570)
    if (oP && oT) continue;
    if (!oP && !oT && ((cod_v(s, vI) < 3 && cod_v(s, tv[0]) < 3) || (dbc_v(s, vI) && dbc_v(s, tv[0]) && dbc_v(s, tv[1]) && dbc_v(s, tv[2])))) continue;
    push_cand(out, make_int2(-svI - 1, sfI));
    }
}
__global__ void __launch_bounds__(256) k_ccd_filter_ee(SurfArgs s, const int2* __restrict__ pairs, const unsigned* __restrict__ nPairs, unsigned cap, const int* __restrict__ vmin,
    const int* __restrict__ vmax, const Box* __restrict__ eboxes, CandOut out)
{
    const unsigned nP = min(*nPairs, cap);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nP; i += gridDim.x * blockDim.x) {
    const int eI = pairs[i].x, eJ = pairs[i].y;
    const int a[2] = { s.SE[2 * eI], s.SE[2 * eI + 1] }, b[2] = { s.SE[2 * eJ], s.SE[2 * eJ + 1] };
    if (a[0] == b[0] || a[0] == b[1] || a[1] == b[0] || a[1] == b[1]) continue;
    int qlo[3], qhi[3], lo[3], hi[3];
    prim_range(vmin, vmax, a, 2, qlo, qhi);
    prim_range(vmin, vmax, b, 2, lo, hi);
    if (!ranges_overlap(qlo, qhi, lo, hi)) continue;
    {   // swept-AABB test of queryEdgeForEdgesWithBBoxCheck (SpatialHash.hpp:819-828) on the exact boxes {x, x + alpha p} of both edges
        const Box eb = eboxes[eI], jb = eboxes[eJ];
        bool sep = false;
#pragma unroll
        for (int c = 0; c < 3; ++c) sep = sep || (jb.lo[c] - eb.hi[c] > 0.0) || (eb.lo[c] - jb.hi[c] > 0.0);
        if (sep) continue;
    }
    const bool oA = obstacle_vertex(s, a[0]), oB = obstacle_vertex(s, b[0]); // (MeshCO.cpp:1576-1660)
    if (oA && oB) continue;
    if (!oA && !oB && ((cod_v(s, a[0]) < 3 && cod_v(s, b[0]) < 3) || (dbc_v(s, a[0]) && dbc_v(s, a[1]) && dbc_v(s, b[0]) && dbc_v(s, b[1])))) continue;
    push_cand(out, make_int2(eI, eJ));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Tight-Inclusion pieces shared by both narrow-phase stages
// ------------------------------------------------------------------------------------------------------------------
struct TiPair {
    double x0[12], x1[12]; // 4 vertices at t=0 / t=1 ; VF: (p,t0,t1,t2)  EE: (a0,a1,b0,b1)
    // 1: an edge pair of the mesh and the obstacle that runs through the vertex-face ROUTINE on its four points in edge order (MeshCO.cpp:900-940,
    // :1609-1655 call vertexFaceCCD_double there) while its initial distance and its error bound stay the edge-edge ones
    int ee_metric;
};
// the routine flag `vf` says which inclusion function runs; distance and error bound follow the kind of the PAIR
DEV bool vf_metric(bool vf, const TiPair& P) { return vf && !P.ee_metric; }
struct DBox { // parameter box: [n/2^k, (n+1)/2^k] per axis; kk = tk | uk<<8 | vk<<16 | flags<<24
    unsigned long long tn, un, vn;
    unsigned kk;
    unsigned pad;
};
DEV double pow2neg(int k) { return __longlong_as_double((long long)(1023 - k) << 52); }
DEV double dy_lo(unsigned long long n, int k) { return (double)n * pow2neg(k); }
DEV double dy_hi(unsigned long long n, int k) { return (double)(n + 1) * pow2neg(k); }

// co-domain test of one box: returns zero_in; sets box_in and max/each true_tol   (oracle: origin_in_box)
DEV bool origin_in_box(bool VF, const TiPair& P, const DBox& b, const double* err, double ms, bool& box_in, double* true_tol)
{
    const int tk = b.kk & 0xff, uk = (b.kk >> 8) & 0xff, vk = (b.kk >> 16) & 0xff;
    const double tv[2] = { dy_lo(b.tn, tk), dy_hi(b.tn, tk) }, uv[2] = { dy_lo(b.un, uk), dy_hi(b.un, uk) }, vv[2] = { dy_lo(b.vn, vk), dy_hi(b.vn, vk) };
    box_in = true;
    bool zero_in = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double mn = 1e300, mx = -1e300;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            double p[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) p[k] = (P.x1[3 * k + c] - P.x0[3 * k + c]) * tv[i] + P.x0[3 * k + c];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int l = 0; l < 2; ++l) {
                    double f;
                    if (VF) {
                        const double pt = ((p[2] - p[1]) * uv[j] + (p[3] - p[1]) * vv[l]) + p[1];
                        f = p[0] - pt;
                    }
                    else {
                        const double pa = (p[1] - p[0]) * uv[j] + p[0];
                        const double pb = (p[3] - p[2]) * vv[l] + p[2];
                        f = pa - pb;
                    }
                    mn = fmin(mn, f);
                    mx = fmax(mx, f);
                }
        }
        true_tol[c] = mx - mn;
        const double eps = err[c] + ms;
        if (mn > eps || mx < -eps) zero_in = false;
        if (!(mn >= -eps && mx <= eps)) box_in = false;
    }
    return zero_in;
}

DEV double linf3(const double* a, const double* b) { return fmax(fmax(fabs(a[0] - b[0]), fabs(a[1] - b[1])), fabs(a[2] - b[2])); }

DEV void width_tolerances(bool VF, const TiPair& P, double tolerance, double* tol)
{
    double ps[4][3], pe[4][3];
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const double* x = side ? P.x1 : P.x0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double q0, q1, q2, q3;
            if (VF) {
                q0 = x[c] - x[3 + c];
                q1 = x[c] - x[9 + c];
                q2 = x[c] - (x[6 + c] + x[9 + c] - x[3 + c]);
                q3 = x[c] - x[6 + c];
            }
            else {
                q0 = x[c] - x[6 + c];
                q1 = x[c] - x[9 + c];
                q2 = x[3 + c] - x[9 + c];
                q3 = x[3 + c] - x[6 + c];
            }
            if (side) { pe[0][c] = q0; pe[1][c] = q1; pe[2][c] = q2; pe[3][c] = q3; }
            else { ps[0][c] = q0; ps[1][c] = q1; ps[2][c] = q2; ps[3][c] = q3; }
        }
    }
    double dl = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) dl = fmax(dl, linf3(pe[q], ps[q]));
    const double e0 = fmax(fmax(linf3(ps[3], ps[0]), linf3(pe[3], pe[0])), fmax(linf3(pe[2], pe[1]), linf3(ps[2], ps[1])));
    const double e1 = fmax(fmax(linf3(ps[1], ps[0]), linf3(pe[1], pe[0])), fmax(linf3(pe[2], pe[3]), linf3(ps[2], ps[3])));
    tol[0] = tolerance / (3.0 * dl);
    tol[1] = tolerance / (3.0 * e0);
    tol[2] = tolerance / (3.0 * e1);
}

DEV void load_pair(const SurfArgs& s, const double* __restrict__ dir, int2 c, bool& vf, int* v, TiPair& P)
{
    vf = c.x < 0;
    if (vf) {
        const int svI = -c.x - 1, sfI = c.y;
        v[0] = s.SVI[svI]; v[1] = s.SF[sfI]; v[2] = s.SF[(size_t)s.nSF + sfI]; v[3] = s.SF[(size_t)2 * s.nSF + sfI];
    }
    else {
        v[0] = s.SE[2 * c.x]; v[1] = s.SE[2 * c.x + 1]; v[2] = s.SE[2 * c.y]; v[3] = s.SE[2 * c.y + 1];
    }
    P.ee_metric = 0;
    if (!vf && s.ee_as_vf && obstacle_vertex(s, v[0]) != obstacle_vertex(s, v[2])) { // (mesh edge first: its sorted index is the smaller one)
        vf = true;
        P.ee_metric = 1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double x = __ldg(s.V + (size_t)q * s.nV + v[k]);
            P.x0[3 * k + q] = x;
            P.x1[3 * k + q] = x + __ldg(dir + 3 * (size_t)v[k] + q);
        }
}

DEV double pair_distance_sqrt(bool vf, const TiPair& P)
{
    const V3 a = { P.x0[0], P.x0[1], P.x0[2] }, b = { P.x0[3], P.x0[4], P.x0[5] }, c = { P.x0[6], P.x0[7], P.x0[8] }, d = { P.x0[9], P.x0[10], P.x0[11] };
    return sqrt(vf_metric(vf, P) ? point_tri_d(a, b, c, d) : edge_edge_d(a, b, c, d));
}

// ------------------------------------------------------------------------------------------------------------------
// stage 1: one thread per candidate
// ------------------------------------------------------------------------------------------------------------------
struct NarrowArgs {
    SurfArgs s;
    const double* dir;
    const int2* cand;
    double err_vf[3], err_ee[3];
    double tol;
    int max_itr;
    // device-resident: cand_range = the slice of `cand` this rank walks, max_t = the step on entry (every pair's max_t), ccd_ord = the
    // running device-wide minimum (ordered-uint image): boxes starting at or after it cannot lower the result
    const IterState* st;
};

__global__ void __launch_bounds__(128) k_ti_stage1(NarrowArgs a, unsigned* __restrict__ survivors, unsigned* __restrict__ nSurv, int* __restrict__ zero_flag)
{
    const unsigned long long begin = a.st->cand_range[0], end = a.st->cand_range[1];
    // grid-stride over whole warps (the warp-aggregated append below needs all 32 lanes)
    for (unsigned long long base = begin + (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < end; base += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long i = base + (threadIdx.x & 31);
    bool alive = false;
    if (i < end) {
        bool vf;
        int v[4];
        TiPair P;
        load_pair(a.s, a.dir, a.cand[i], vf, v, P);
        const double d = pair_distance_sqrt(vf, P);
        if (d == 0.0) atomicExch(zero_flag, 1); // "Initial CCD distance is zero! Returning 0 stepSize." (:730-737)
        else {
            const double ms = fmin(0.2 * d, 1e-6);
            DBox root = { 0ull, 0ull, 0ull, 0u, 0u };
            bool box_in;
            double tt[3];
            alive = origin_in_box(vf, P, root, vf_metric(vf, P) ? a.err_vf : a.err_ee, ms, box_in, tt);
        }
    }
    // warp-aggregated append of the survivors
    const unsigned m = __ballot_sync(0xffffffffu, alive);
    if (m) {
        const int lane = threadIdx.x & 31;
        unsigned base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(nSurv, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (alive) survivors[base + __popc(m & ((1u << lane) - 1))] = (unsigned)i;
    }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// stage 2: warp per surviving pair
// ------------------------------------------------------------------------------------------------------------------
struct Key3 {
    double t, u, v;
};
DEV bool key_less(const Key3& a, const Key3& b)
{
    if (a.t != b.t) return a.t < b.t;
    if (a.u != b.u) return a.u < b.u;
    return a.v < b.v;
}
// warp min of (key, payload)
DEV void warp_min_key(Key3& k, unsigned& pay, double& aux)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Key3 r;
        r.t = __shfl_xor_sync(0xffffffffu, k.t, o);
        r.u = __shfl_xor_sync(0xffffffffu, k.u, o);
        r.v = __shfl_xor_sync(0xffffffffu, k.v, o);
        const unsigned rp = __shfl_xor_sync(0xffffffffu, pay, o);
        const double ra = __shfl_xor_sync(0xffffffffu, aux, o);
        if (key_less(r, k)) {
            k = r;
            pay = rp;
            aux = ra;
        }
    }
}
DEV bool sum_le_1(unsigned long long an, int ak, unsigned long long bn, int bk)
{
    // lo(a) + lo(b) <= 1 exactly; k <= 60 so the aligned sum fits in 64 bits when the larger exponent is used carefully
    const int k = max(ak, bk);
    // n < 2^ak: n << (k-ak) < 2^k <= 2^60 -> the sum is < 2^61
    const unsigned long long s = (an << (k - ak)) + (bn << (k - bk));
    return s <= (1ull << k);
}

constexpr unsigned F_ZERO = 1u << 24; // zero_in flag stored in DBox::kk

constexpr int kStage2WarpsPerCtaDev = 4;
constexpr int kSmemLevel = 184;
constexpr int kWideLevel = 10; // levels with at least this many boxes are evaluated box-parallel, narrower ones corner-parallel
__constant__ int c_wide_level = kWideLevel; // (IPCGPU_TI_WIDE_LEVEL overrides it for tuning runs)    // boxes per level buffer kept in shared memory by the warp-level pass

// group helpers: W = 32 (one warp per pair) or W = 1 (one thread per pair, no cross-lane traffic)
template <int W> DEV void group_min_key(Key3& k, unsigned& pay, double& aux) { if (W == 32) warp_min_key(k, pay, aux); }
template <int W> DEV int group_count(bool b) { return (W == 32) ? __popc(__ballot_sync(0xffffffffu, b)) : (b ? 1 : 0); }
template <int W> DEV bool group_any(bool b) { return (W == 32) ? (__any_sync(0xffffffffu, b) != 0) : b; }
template <int W> DEV void group_sync() { if (W == 32) __syncwarp(); }
template <int W> DEV void group_scan(int v, int lane, int& incl, int& total)
{
    incl = v;
    total = v;
    if (W == 32) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        total = __shfl_sync(0xffffffffu, incl, 31);
    }
}

// result codes of the root finder: 0 no collision, 1 collision (toi set), 2 deferred (W = 1 only: level buffer too small)
template <int W>
__device__ int ti_root_finder(bool VF, const TiPair& P, const double* tol, const double* inv_tol, double co_tol, double max_t, const double* err, double ms, int max_itr, DBox* bufA,
    DBox* bufB, int gcap, int lane, double& toi, double& out_tol, int* __restrict__ warn, DBox* sA = nullptr, DBox* sB = nullptr, long long thread_budget = 0,
    const unsigned long long* best = nullptr)
{
    const bool check_t = (max_t != 1.0);
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    // levels start in shared memory (when provided) and migrate to the global buffers if they outgrow it
    bool in_smem = (sA != nullptr);
    DBox* cur = in_smem ? sA : bufA;
    DBox* nxt = in_smem ? sB : bufB;
    int cap = in_smem ? kSmemLevel : gcap;
    if (lane == 0) cur[0] = DBox{ 0ull, 0ull, 0ull, 0u, 0u };
    group_sync<W>();
    int n = 1;
    double toi_skip = INF;
    bool use_skip = false;
    long long refine = 0;
    double temp_toi = INF, temp_out_tol = co_tol;
    DBox* gA = bufA; // global (or local) backing store used when a level outgrows the shared-memory buffers
    DBox* gB = bufB;
    out_tol = co_tol;
    toi = INF;
    unsigned long long bo_cur = best ? *reinterpret_cast<const volatile unsigned long long*>(best) : 0ull;
    while (n > 0) {
        // exact pruning against the running device-wide minimum: a box that starts at t_lo >= max(best, 1e-6) can only yield a time of
        // impact >= best (and no 0.8-rescaled retry, which needs toi < 1e-6), so dropping it cannot change the final min over pairs
        // The bound used here was requested one level ago (stale = prunes less, never wrong): its L2 round trip overlaps the level.
        double t_prune = INF;
        if (best) {
            t_prune = fmax(ord_to_dbl(bo_cur), 1e-6);
            bo_cur = *reinterpret_cast<const volatile unsigned long long*>(best); // W == 32: same address on all lanes, one value
        }
        // ---- pass 1: evaluate, find K1 (first box containing the origin) and K2 (first terminal box) ----------
        Key3 k1 = { INF, INF, INF }, k2 = { INF, INF, INF };
        unsigned p1 = 0, p2 = 0; // payload bit0: flagged (K1) / cond1 (K2)
        double a1 = 0.0, a2 = 0.0;
        int visited = 0;
        for (int base = 0; base < n; base += W) {
            const int i = base + lane;
            Key3 mk1 = { INF, INF, INF }, mk2 = { INF, INF, INF };
            unsigned mp1 = 0, mp2 = 0;
            double ma1 = 0.0, ma2 = 0.0;
            bool vis = false;
            if (i < n) {
                DBox b = cur[i];
                const int tk = b.kk & 0xff, uk = (b.kk >> 8) & 0xff, vk = (b.kk >> 16) & 0xff;
                const double tlo = dy_lo(b.tn, tk);
                unsigned flags = 0;
                if (tlo < toi_skip && tlo < t_prune) {
                    vis = true;
                    bool box_in;
                    double tt[3];
                    if (origin_in_box(VF, P, b, err, ms, box_in, tt)) {
                        flags = F_ZERO;
                        const bool tol_cond = tt[0] <= co_tol && tt[1] <= co_tol && tt[2] <= co_tol;
                        const bool cond1 = pow2neg(tk) <= tol[0] && pow2neg(uk) <= tol[1] && pow2neg(vk) <= tol[2];
                        const Key3 key = { tlo, dy_lo(b.un, uk), dy_lo(b.vn, vk) };
                        mk1 = key;
                        mp1 = (tol_cond || box_in || cond1) ? 1u : 0u;
                        ma1 = fmax(fmax(tt[0], tt[1]), tt[2]);
                        if (mp1) {
                            mk2 = key;
                            mp2 = cond1 ? 1u : 0u;
                        }
                    }
                }
                cur[i].kk = (b.kk & 0x00ffffffu) | flags;
            }
            visited += group_count<W>(vis);
            group_min_key<W>(mk1, mp1, ma1);
            group_min_key<W>(mk2, mp2, ma2);
            if (key_less(mk1, k1)) { k1 = mk1; p1 = mp1; a1 = ma1; }
            if (key_less(mk2, k2)) { k2 = mk2; p2 = mp2; a2 = ma2; }
        }
        group_sync<W>();
        const bool any_zero = k1.t != INF;
        if (!any_zero) { // nothing at this level contains the origin: the search space is exhausted
            n = 0;
            break;
        }
        if (p1 & 1u) { // the first box containing the origin is terminal: conditions 1/2/3 of the library
            toi = k1.t;
            return 1;
        }
        const bool has_k2 = k2.t != INF;
        if (has_k2 && (p2 & 1u)) { // a later box already below the width tolerances (condition 1)
            // boxes between K1 and K2 do not matter: the library returns here
            toi = k2.t;
            return 1;
        }
        if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(warn + (W == 1 ? 5 : 7)), (unsigned long long)visited); // diagnostics: boxes evaluated
        if (W == 1 && refine + visited > thread_budget) return 2; // over budget: handed to the next pass
        if (max_itr > 0) {
            temp_toi = k1.t;
            temp_out_tol = fmax(a1, co_tol);
            refine += visited;
            if (refine > max_itr) { // conservative early-out (the library returns its per-level estimate here as well)
                if (lane == 0) atomicAdd(warn, 1);
                toi = temp_toi;
                out_tol = temp_out_tol;
                return 1;
            }
        }
        if (has_k2) {
            if (k2.t < toi_skip) toi_skip = k2.t;
            use_skip = true;
        }
        // ---- pass 2: split every box that contains the origin and precedes K2 ------------------------------------------
        int nn = 0;
        bool over = false;
        for (int base = 0; base < n; base += W) {
            const int i = base + lane;
            int nchild = 0;
            DBox c0, c1;
            if (i < n) {
                const DBox b = cur[i];
                if (b.kk & F_ZERO) {
                    const int tk = b.kk & 0xff, uk = (b.kk >> 8) & 0xff, vk = (b.kk >> 16) & 0xff;
                    const Key3 key = { dy_lo(b.tn, tk), dy_lo(b.un, uk), dy_lo(b.vn, vk) };
                    if (!has_k2 || key_less(key, k2)) {
                        const double w[3] = { pow2neg(tk), pow2neg(uk), pow2neg(vk) };
                        int split = -1;
                        double best = -1.0;
#pragma unroll
                        for (int d = 0; d < 3; ++d)
                            if (w[d] > tol[d]) {
                                const double r = inv_tol[d] * w[d]; // = w[d] / tol[d] bit for bit: w[d] is a power of two (see ti_ccd)
                                if (r > best) { best = r; split = d; }
                            }
                        const int pk = split == 0 ? tk : (split == 1 ? uk : vk);
                        if (split < 0 || pk >= 60) over = true; // bisection overflow: handled like the iteration overflow
                        else {
                            const unsigned long long pn = split == 0 ? b.tn : (split == 1 ? b.un : b.vn);
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                const unsigned long long hn = 2 * pn + half;
                                const int hk = pk + 1;
                                bool keep = true;
                                if (split == 0) { if (check_t) keep = !(dy_hi(hn, hk) < 0.0 || dy_lo(hn, hk) > max_t); }
                                else if (VF) keep = (split == 1) ? sum_le_1(hn, hk, b.vn, vk) : sum_le_1(hn, hk, b.un, uk);
                                if (keep) {
                                    DBox c = b;
                                    c.kk &= 0x00ffffffu;
                                    if (split == 0) { c.tn = hn; c.kk = (c.kk & ~0xffu) | (unsigned)hk; }
                                    else if (split == 1) { c.un = hn; c.kk = (c.kk & ~0xff00u) | ((unsigned)hk << 8); }
                                    else { c.vn = hn; c.kk = (c.kk & ~0xff0000u) | ((unsigned)hk << 16); }
                                    if (nchild == 0) c0 = c;
                                    else c1 = c;
                                    ++nchild;
                                }
                            }
                        }
                    }
                }
            }
            // group exclusive scan of nchild
            int incl, total;
            group_scan<W>(nchild, lane, incl, total);
            const int off = nn + incl - nchild;
            if (W == 32 && in_smem && nn + total > cap) { // this level no longer fits in shared memory: continue it in global memory
                for (int q = lane; q < nn; q += 32) gB[q] = nxt[q];
                group_sync<W>();
                nxt = gB;
                cap = gcap;
                in_smem = false;
            }
            if (nn + total > cap) over = true;
            else {
                if (nchild > 0) nxt[off] = c0;
                if (nchild > 1) nxt[off + 1] = c1;
            }
            nn += total;
            if (group_any<W>(over)) { over = true; break; }
        }
        if (group_any<W>(over)) {
            if (W == 1) return 2; // the thread-level pass hands the pair to the warp-level pass
            // level buffer / bisection depth exhausted: return the conservative per-level estimate (earliest box containing the origin)
            if (lane == 0) atomicAdd(warn, 1);
            toi = temp_toi;
            out_tol = temp_out_tol;
            return 1;
        }
        group_sync<W>();
        if (W == 32 && !in_smem && nxt == gB && cur != gA) { // first level after the migration: from now on ping-pong between the global buffers
            cur = gB;
            nxt = gA;
        }
        else {
            DBox* t = cur; cur = nxt; nxt = t;
        }
        n = nn;
    }
    if (use_skip) {
        toi = toi_skip;
        return 1;
    }
    return 0;
}

// ---- corner-parallel variant of the root finder for the warp-level pass ------------------------------------------------------
// Deep searches are narrow (a handful of boxes per level), so there is little parallelism over boxes; what can be parallelised is
// the box itself.  A narrow level is evaluated four boxes at a time: each 8-lane group takes one box, a lane is one of its 8 corners
// and evaluates the 3 coordinates of F in turn; the co-domain interval of a coordinate is a 3-step shuffle min/max inside the group,
// after which all 8 lanes know every inclusion flag of their box (no ballots).  Each group keeps its own K1/K2 candidates, merged by
// one 2-step exchange per level.  Levels of >= kWideLevel boxes switch to one box per lane.  The split pass is always one box per
// lane with a warp scan.  Level buffers live in shared memory.  Same arithmetic per corner, same decisions => same result as the
// box-parallel variant.  Returns -1 when a level outgrows the shared-memory buffer (the caller restarts with the box-parallel variant).
__device__ int ti_root_finder_cp(bool VF, const TiPair& P, const double* tol, const double* inv_tol, double co_tol, double max_t, const double* err, double ms, int max_itr, DBox* sA, DBox* sB,
    int lane, double& toi, double& out_tol, int* __restrict__ warn, const unsigned long long* best)
{
    const bool check_t = (max_t != 1.0);
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    const int corner = lane & 7;
    const int ci = corner >> 2, cj = (corner >> 1) & 1, cl = corner & 1;
    DBox* cur = sA;
    DBox* nxt = sB;
    if (lane == 0) cur[0] = DBox{ 0ull, 0ull, 0ull, 0u, 0u };
    __syncwarp();
    int n = 1;
    double toi_skip = INF;
    bool use_skip = false;
    long long refine = 0;
    double temp_toi = INF, temp_out_tol = co_tol;
    out_tol = co_tol;
    toi = INF;
    unsigned long long bo_cur = best ? *reinterpret_cast<const volatile unsigned long long*>(best) : 0ull;
    while (n > 0) {
        Key3 k1 = { INF, INF, INF }, k2 = { INF, INF, INF };
        unsigned p1 = 0, p2 = 0;
        double a1max = 0.0;
        int visited = 0;
        // exact pruning against the running device-wide minimum (see ti_root_finder).  The value used at this level was requested
        // one level ago (a stale bound only prunes less), so its L2 round trip is off the per-level critical path.
        double t_prune = INF;
        if (best) {
            t_prune = fmax(ord_to_dbl(bo_cur), 1e-6);
            bo_cur = *reinterpret_cast<const volatile unsigned long long*>(best); // same address on all lanes: one broadcast request
        }
        if (n >= c_wide_level) {
            // wide level: one box per lane (box-parallel), K1/K2 by warp min-reduction over the keys
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                Key3 mk1 = { INF, INF, INF }, mk2 = { INF, INF, INF };
                unsigned mp1 = 0, mp2 = 0;
                double ma1 = 0.0, ma2 = 0.0;
                bool vis = false;
                if (i < n) {
                    const DBox b = cur[i];
                    const int tk = b.kk & 0xff, uk = (b.kk >> 8) & 0xff, vk = (b.kk >> 16) & 0xff;
                    const double tlo = dy_lo(b.tn, tk);
                    unsigned flags = 0;
                    if (tlo < toi_skip && tlo < t_prune) {
                        vis = true;
                        bool box_in;
                        double tt[3];
                        if (origin_in_box(VF, P, b, err, ms, box_in, tt)) {
                            flags = F_ZERO;
                            const bool tol_cond = tt[0] <= co_tol && tt[1] <= co_tol && tt[2] <= co_tol;
                            const bool cond1 = pow2neg(tk) <= tol[0] && pow2neg(uk) <= tol[1] && pow2neg(vk) <= tol[2];
                            const Key3 key = { tlo, dy_lo(b.un, uk), dy_lo(b.vn, vk) };
                            mk1 = key;
                            mp1 = (tol_cond || box_in || cond1) ? 1u : 0u;
                            ma1 = fmax(fmax(tt[0], tt[1]), tt[2]);
                            if (mp1) { mk2 = key; mp2 = cond1 ? 1u : 0u; }
                        }
                    }
                    cur[i].kk = (b.kk & 0x00ffffffu) | flags;
                }
                visited += __popc(__ballot_sync(0xffffffffu, vis));
                warp_min_key(mk1, mp1, ma1);
                warp_min_key(mk2, mp2, ma2);
                if (key_less(mk1, k1)) { k1 = mk1; p1 = mp1; a1max = ma1; }
                if (key_less(mk2, k2)) { k2 = mk2; p2 = mp2; }
            }
        }
        else {
            // narrow level: FOUR boxes at a time, one per 8-lane group; a lane is one corner and evaluates the three coordinates in
            // turn, so the co-domain interval of a coordinate is a 3-step shuffle min/max inside the group and every inclusion flag
            // is known to all 8 lanes without a ballot.  Each group tracks its own K1/K2; one 2-step exchange merges them.
            Key3 gk1 = { INF, INF, INF }, gk2 = { INF, INF, INF };
            unsigned gp1 = 0, gp2 = 0;
            double ga1 = 0.0, ga2 = 0.0;
            const int grp = lane >> 3;
            for (int base = 0; base < n; base += 4) {
                const int bI = base + grp;
                DBox b = DBox{ 0ull, 0ull, 0ull, 0u, 0u };
                if (bI < n) b = cur[bI];
                const int tk = b.kk & 0xff, uk = (b.kk >> 8) & 0xff, vk = (b.kk >> 16) & 0xff;
                const double tlo = dy_lo(b.tn, tk);
                const bool act = bI < n && tlo < toi_skip && tlo < t_prune;
                const double tv = ci ? dy_hi(b.tn, tk) : tlo;
                const double uv = cj ? dy_hi(b.un, uk) : dy_lo(b.un, uk);
                const double vv = cl ? dy_hi(b.vn, vk) : dy_lo(b.vn, vk);
                bool excl = false, inside = true, tolc = true;
                double tmax = 0.0;
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    double pp[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) pp[k] = (P.x1[3 * k + cc] - P.x0[3 * k + cc]) * tv + P.x0[3 * k + cc];
                    double f;
                    if (VF) {
                        const double pt = ((pp[2] - pp[1]) * uv + (pp[3] - pp[1]) * vv) + pp[1];
                        f = pp[0] - pt;
                    }
                    else {
                        const double pa = (pp[1] - pp[0]) * uv + pp[0];
                        const double pb = (pp[3] - pp[2]) * vv + pp[2];
                        f = pa - pb;
                    }
                    double mn = f, mx = f;
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    }
                    const double e = err[cc] + ms;
                    const double tt = mx - mn;
                    excl = excl || (mn > e || mx < -e);
                    inside = inside && (mn >= -e && mx <= e);
                    tolc = tolc && !(tt > co_tol);
                    tmax = (cc == 0) ? tt : fmax(tmax, tt);
                }
                unsigned flags = 0;
                if (act && !excl) { // the co-domain box contains the origin
                    flags = F_ZERO;
                    const bool cond1 = pow2neg(tk) <= tol[0] && pow2neg(uk) <= tol[1] && pow2neg(vk) <= tol[2];
                    const Key3 key = { tlo, dy_lo(b.un, uk), dy_lo(b.vn, vk) };
                    const bool flagged = tolc || inside || cond1;
                    if (key_less(key, gk1)) { gk1 = key; gp1 = flagged ? 1u : 0u; ga1 = tmax; }
                    if (flagged && key_less(key, gk2)) { gk2 = key; gp2 = cond1 ? 1u : 0u; }
                }
                visited += __popc(__ballot_sync(0xffffffffu, act && corner == 0));
                if (bI < n && corner == 0) cur[bI].kk = (b.kk & 0x00ffffffu) | flags;
            }
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) { // merge the four groups (all lanes of a group hold the same candidates)
                Key3 r1, r2;
                r1.t = __shfl_xor_sync(0xffffffffu, gk1.t, o); r1.u = __shfl_xor_sync(0xffffffffu, gk1.u, o); r1.v = __shfl_xor_sync(0xffffffffu, gk1.v, o);
                r2.t = __shfl_xor_sync(0xffffffffu, gk2.t, o); r2.u = __shfl_xor_sync(0xffffffffu, gk2.u, o); r2.v = __shfl_xor_sync(0xffffffffu, gk2.v, o);
                const unsigned rp1 = __shfl_xor_sync(0xffffffffu, gp1, o), rp2 = __shfl_xor_sync(0xffffffffu, gp2, o);
                const double ra1 = __shfl_xor_sync(0xffffffffu, ga1, o);
                if (key_less(r1, gk1)) { gk1 = r1; gp1 = rp1; ga1 = ra1; }
                if (key_less(r2, gk2)) { gk2 = r2; gp2 = rp2; }
            }
            (void)ga2;
            k1 = gk1; p1 = gp1; a1max = ga1;
            k2 = gk2; p2 = gp2;
        }
        __syncwarp();
        if (k1.t == INF) break; // search space exhausted
        if (p1 & 1u) { toi = k1.t; return 1; }
        const bool has_k2 = k2.t != INF;
        if (has_k2 && (p2 & 1u)) { toi = k2.t; return 1; }
        if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(warn + 7), (unsigned long long)visited);
        if (max_itr > 0) {
            temp_toi = k1.t;
            temp_out_tol = fmax(a1max, co_tol);
            refine += visited;
            if (refine > max_itr) {
                if (lane == 0) atomicAdd(warn, 1);
                toi = temp_toi;
                out_tol = temp_out_tol;
                return 1;
            }
        }
        if (has_k2) {
            if (k2.t < toi_skip) toi_skip = k2.t;
            use_skip = true;
        }
        int nn = 0;
        { // split pass, one box per lane (levels of the warp-level pass never exceed a few dozen boxes before the box-parallel switch)
            bool over = false, deep = false;
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                int nchild = 0;
                DBox c0, c1;
                if (i < n) {
                    const DBox b = cur[i];
                    if (b.kk & F_ZERO) {
                        const int tk = b.kk & 0xff, uk = (b.kk >> 8) & 0xff, vk = (b.kk >> 16) & 0xff;
                        const Key3 key = { dy_lo(b.tn, tk), dy_lo(b.un, uk), dy_lo(b.vn, vk) };
                        if (!has_k2 || key_less(key, k2)) {
                            const double w[3] = { pow2neg(tk), pow2neg(uk), pow2neg(vk) };
                            int split = -1;
                            double best = -1.0;
#pragma unroll
                            for (int d = 0; d < 3; ++d)
                                if (w[d] > tol[d]) {
                                    const double r = inv_tol[d] * w[d]; // = w[d] / tol[d] bit for bit: w[d] is a power of two (see ti_ccd)
                                    if (r > best) { best = r; split = d; }
                                }
                            const int pk = split == 0 ? tk : (split == 1 ? uk : vk);
                            if (split < 0 || pk >= 60) deep = true;
                            else {
                                const unsigned long long pn = split == 0 ? b.tn : (split == 1 ? b.un : b.vn);
#pragma unroll
                                for (int half = 0; half < 2; ++half) {
                                    const unsigned long long hn = 2 * pn + half;
                                    const int hk = pk + 1;
                                    bool keep = true;
                                    if (split == 0) { if (check_t) keep = !(dy_hi(hn, hk) < 0.0 || dy_lo(hn, hk) > max_t); }
                                    else if (VF) keep = (split == 1) ? sum_le_1(hn, hk, b.vn, vk) : sum_le_1(hn, hk, b.un, uk);
                                    if (keep) {
                                        DBox ch = b;
                                        ch.kk &= 0x00ffffffu;
                                        if (split == 0) { ch.tn = hn; ch.kk = (ch.kk & ~0xffu) | (unsigned)hk; }
                                        else if (split == 1) { ch.un = hn; ch.kk = (ch.kk & ~0xff00u) | ((unsigned)hk << 8); }
                                        else { ch.vn = hn; ch.kk = (ch.kk & ~0xff0000u) | ((unsigned)hk << 16); }
                                        if (nchild == 0) c0 = ch;
                                        else c1 = ch;
                                        ++nchild;
                                    }
                                }
                            }
                        }
                    }
                }
                int incl, total;
                group_scan<32>(nchild, lane, incl, total);
                const int off = nn + incl - nchild;
                if (nn + total > kSmemLevel) over = true;
                else {
                    if (nchild > 0) nxt[off] = c0;
                    if (nchild > 1) nxt[off + 1] = c1;
                }
                nn += total;
                if (__any_sync(0xffffffffu, over || deep)) break;
            }
            if (__any_sync(0xffffffffu, deep)) {
                if (lane == 0) atomicAdd(warn, 1);
                toi = temp_toi;
                out_tol = temp_out_tol;
                return 1;
            }
            if (__any_sync(0xffffffffu, over)) return -1;
        }
        __syncwarp();
        DBox* t = cur; cur = nxt; nxt = t;
        n = nn;
    }
    if (use_skip) { toi = toi_skip; return 1; }
    return 0;
}

// out-of-line instances for the warp-level pass.  Round 2, second half: ncu on k_ti_stage2 showed "no instruction" as the largest stall
// (3.4 per issued instruction): 16 K SASS instructions -- VF and EE template copies of both root finders, each inlined at two call sites
// (first run and ms = 0 retry) -- with four warps per CTA at unrelated program counters.  VF / EE is now a run-time flag tested inside the
// corner evaluation (a few instructions differ), the retry is a second trip through ONE call site, and the split rule multiplies by exact
// reciprocals instead of dividing (below): one copy of each root finder.
__device__ __noinline__ int ti_root_finder_cp_call(bool vf, const TiPair& P, const double* tol, const double* inv_tol, double co_tol, double max_t, const double* err, double ms,
    int max_itr, DBox* sA, DBox* sB, int lane, double& toi, double& out_tol, int* __restrict__ warn, const unsigned long long* best)
{
    return ti_root_finder_cp(vf, P, tol, inv_tol, co_tol, max_t, err, ms, max_itr, sA, sB, lane, toi, out_tol, warn, best);
}
__device__ __noinline__ int ti_root_finder_warp_call(bool vf, const TiPair& P, const double* tol, const double* inv_tol, double co_tol, double max_t, const double* err, double ms,
    int max_itr, DBox* bufA, DBox* bufB, int cap, int lane, double& toi, double& out_tol, int* __restrict__ warn, const unsigned long long* best)
{
    return ti_root_finder<32>(vf, P, tol, inv_tol, co_tol, max_t, err, ms, max_itr, bufA, bufB, cap, lane, toi, out_tol, warn, nullptr, nullptr, 0, best);
}

// vertexFaceCCD_double / edgeEdgeCCD_double including the no_zero_toi refinement loop; returns 0 / 1 / 2 (deferred)
template <int W>
__device__ int ti_ccd(bool vf, const TiPair& P, const double* err, double ms, double tolerance, double t_max, int max_itr, DBox* bufA, DBox* bufB, int cap, int lane,
    double& toi, int* __restrict__ warn, DBox* sA = nullptr, DBox* sB = nullptr, long long thread_budget = 0, const unsigned long long* best = nullptr)
{
    double tolerance_in = tolerance, ms_in = ms, out_tol = tolerance;
    bool is_impacting = false, tmp = false;
    unsigned iter = 0;
    do {
        double tol[3], inv_tol[3];
        width_tolerances(vf, P, tolerance_in, tol);
        // The split rule compares width / tolerance per axis.  Every width is a power of two, and x -> x * 2^-k commutes with rounding
        // (no overflow / underflow here: tolerances are ~1e-9 .. 1e-3 or inf for a static axis), so 2^-k / tol == (1 / tol) * 2^-k bit
        // for bit: one division per axis and search instead of three per box.
#pragma unroll
        for (int d = 0; d < 3; ++d) inv_tol[d] = 1.0 / tol[d];
        int rc;
        if (W == 32 && sA) {
            rc = ti_root_finder_cp_call(vf, P, tol, inv_tol, tolerance_in, t_max, err, ms_in, max_itr, sA, sB, lane, toi, out_tol, warn, best);
            if (rc == -1) rc = ti_root_finder_warp_call(vf, P, tol, inv_tol, tolerance_in, t_max, err, ms_in, max_itr, bufA, bufB, cap, lane, toi, out_tol, warn, best);
        }
        else rc = ti_root_finder<W>(vf, P, tol, inv_tol, tolerance_in, t_max, err, ms_in, max_itr, bufA, bufB, cap, lane, toi, out_tol, warn, sA, sB, thread_budget, best);
        if (rc == 2) return 2;
        tmp = rc == 1;
        if (iter == 0) is_impacting = tmp;
        else toi = tmp ? toi : t_max;
        if (tmp && toi == 0.0) {
            if (out_tol > tolerance_in) t_max *= 0.9;
            else if (10 * tolerance_in < ms_in) ms_in *= 0.5;
            else tolerance_in *= 0.1;
        }
        ++iter;
    } while (iter < 0x7fffffffu && tmp && toi == 0.0);
    return is_impacting ? 1 : 0;
}

// one candidate end to end (SelfCollisionHandler.cpp:740-790): 0 / 1 (toi set) / 2 deferred
template <int W>
__device__ int pair_ccd(bool vf, const TiPair& P, const NarrowArgs& a, DBox* bufA, DBox* bufB, int cap, int lane, double& toi, int* __restrict__ warn,
    DBox* sA = nullptr, DBox* sB = nullptr, long long thread_budget = 0)
{
    const double d = pair_distance_sqrt(vf, P);
    const double max_t = a.st->max_t;                 // canonical semantics: every pair sees the step on entry
    const double* err = vf_metric(vf, P) ? a.err_vf : a.err_ee;
    int hit = 0;
    // first trip: ms = min(0.2 d, 1e-6), pruned against the running device-wide minimum; second trip (:759-781, only after a hit with
    // toi < 1e-6): ms = 0, result scaled by 0.8 and NOT pruned -- a box starting in [best, 1.25 best) can still lower the global step
    // (the exactness argument of ti_root_finder only covers unscaled results).  One call site for both.
#pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
        const double ms = attempt ? 0.0 : fmin(0.2 * d, 1e-6);
        const unsigned long long* best = attempt ? nullptr : &a.st->ccd_ord;
        hit = ti_ccd<W>(vf, P, err, ms, a.tol, max_t, a.max_itr, bufA, bufB, cap, lane, toi, warn, sA, sB, thread_budget, best);
        if (hit == 2) return 2;
        if (attempt == 1) {
            if (hit) toi *= 0.8;
            break;
        }
        if (!(hit && toi < 1e-6)) break;
    }
    return hit;
}

// ------------------------------------------------------------------------------------------------------------------
// group pass: EIGHT LANES per surviving pair (four pairs per warp).  A lane is one corner of the parameter box being evaluated, the
// co-domain interval of a coordinate is a 3-step shuffle min/max inside the group -- the same corner-parallel evaluation as the narrow
// levels of the warp-level pass, with the same arithmetic per corner, hence the same decisions and the same time of impact.  What
// changes is the concurrency: the interval search is a chain of short dependent steps (evaluate a level, pick K1/K2, split), and one
// warp per pair left three quarters of its lanes idle on the 1-4 box levels that make up almost every search, at 24 pairs in flight
// per SM.  Four pairs per warp quadruple the searches in flight for the same registers and shared memory.  Levels are walked one box
// at a time per group; the split pass handles eight boxes per round.  A pair whose level outgrows the group's buffer (kGrpCap boxes)
// is handed to the warp-level pass, which restarts it.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGrpCap = 44;       // boxes per level buffer per group (two buffers per group: 4 x 2 x 44 x 32 B = 11.3 KB per warp; static smem stays < 48 KB)
constexpr int kGrpWarpsPerCta = 4;

__device__ int ti_root_finder_grp(bool vf, const TiPair& P, const double* tol, const double* inv_tol, double co_tol, double max_t, const double* err, double ms, int max_itr, DBox* sA, DBox* sB,
    int gl, unsigned gmask, double& toi, double& out_tol, int* __restrict__ warn, const unsigned long long* best)
{
    const bool check_t = (max_t != 1.0);
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    const int ci = gl >> 2, cj = (gl >> 1) & 1, cl = gl & 1;
    DBox* cur = sA;
    DBox* nxt = sB;
    if (gl == 0) cur[0] = DBox{ 0ull, 0ull, 0ull, 0u, 0u };
    __syncwarp(gmask);
    int n = 1;
    double toi_skip = INF;
    bool use_skip = false;
    long long refine = 0;
    double temp_toi = INF, temp_out_tol = co_tol;
    out_tol = co_tol;
    toi = INF;
    unsigned long long bo_cur = best ? *reinterpret_cast<const volatile unsigned long long*>(best) : 0ull;
    while (n > 0) {
        Key3 k1 = { INF, INF, INF }, k2 = { INF, INF, INF };
        unsigned p1 = 0, p2 = 0;
        double a1max = 0.0;
        int visited = 0;
        // exact pruning against the running device-wide minimum (see ti_root_finder); the value used here was requested one level ago
        double t_prune = INF;
        if (best) {
            t_prune = fmax(ord_to_dbl(bo_cur), 1e-6);
            bo_cur = *reinterpret_cast<const volatile unsigned long long*>(best);
        }
        for (int bI = 0; bI < n; ++bI) { // one box at a time, its 8 corners on the 8 lanes (every branch below is uniform over the group)
            const DBox b = cur[bI];
            const int tk = b.kk & 0xff, uk = (b.kk >> 8) & 0xff, vk = (b.kk >> 16) & 0xff;
            const double tlo = dy_lo(b.tn, tk);
            unsigned flags = 0;
            if (tlo < toi_skip && tlo < t_prune) {
                ++visited;
                const double tv = ci ? dy_hi(b.tn, tk) : tlo;
                const double uv = cj ? dy_hi(b.un, uk) : dy_lo(b.un, uk);
                const double vv = cl ? dy_hi(b.vn, vk) : dy_lo(b.vn, vk);
                bool excl = false, inside = true, tolc = true;
                double tmax = 0.0;
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    double pp[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) pp[k] = (P.x1[3 * k + cc] - P.x0[3 * k + cc]) * tv + P.x0[3 * k + cc];
                    double f;
                    if (vf) {
                        const double pt = ((pp[2] - pp[1]) * uv + (pp[3] - pp[1]) * vv) + pp[1];
                        f = pp[0] - pt;
                    }
                    else {
                        const double pa = (pp[1] - pp[0]) * uv + pp[0];
                        const double pb = (pp[3] - pp[2]) * vv + pp[2];
                        f = pa - pb;
                    }
                    double mn = f, mx = f;
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        mn = fmin(mn, __shfl_xor_sync(gmask, mn, o, 8));
                        mx = fmax(mx, __shfl_xor_sync(gmask, mx, o, 8));
                    }
                    const double e = err[cc] + ms;
                    const double tt = mx - mn;
                    excl = excl || (mn > e || mx < -e);
                    inside = inside && (mn >= -e && mx <= e);
                    tolc = tolc && !(tt > co_tol);
                    tmax = (cc == 0) ? tt : fmax(tmax, tt);
                }
                if (!excl) { // the co-domain box contains the origin
                    flags = F_ZERO;
                    const bool cond1 = pow2neg(tk) <= tol[0] && pow2neg(uk) <= tol[1] && pow2neg(vk) <= tol[2];
                    const Key3 key = { tlo, dy_lo(b.un, uk), dy_lo(b.vn, vk) };
                    const bool flagged = tolc || inside || cond1;
                    if (key_less(key, k1)) { k1 = key; p1 = flagged ? 1u : 0u; a1max = tmax; }
                    if (flagged && key_less(key, k2)) { k2 = key; p2 = cond1 ? 1u : 0u; }
                }
            }
            if (gl == 0) cur[bI].kk = (b.kk & 0x00ffffffu) | flags;
        }
        __syncwarp(gmask);
        if (k1.t == INF) break; // search space exhausted
        if (p1 & 1u) { toi = k1.t; return 1; }
        const bool has_k2 = k2.t != INF;
        if (has_k2 && (p2 & 1u)) { toi = k2.t; return 1; }
        if (gl == 0) atomicAdd(reinterpret_cast<unsigned long long*>(warn + 5), (unsigned long long)visited); // diagnostics: boxes of this pass
        if (max_itr > 0) {
            temp_toi = k1.t;
            temp_out_tol = fmax(a1max, co_tol);
            refine += visited;
            if (refine > max_itr) {
                if (gl == 0) atomicAdd(warn, 1);
                toi = temp_toi;
                out_tol = temp_out_tol;
                return 1;
            }
        }
        if (has_k2) {
            if (k2.t < toi_skip) toi_skip = k2.t;
            use_skip = true;
        }
        int nn = 0;
        bool over = false, deep = false;
        for (int base = 0; base < n; base += 8) { // split pass: one box per lane of the group
            const int i = base + gl;
            int nchild = 0;
            DBox c0, c1;
            if (i < n) {
                const DBox b = cur[i];
                if (b.kk & F_ZERO) {
                    const int tk = b.kk & 0xff, uk = (b.kk >> 8) & 0xff, vk = (b.kk >> 16) & 0xff;
                    const Key3 key = { dy_lo(b.tn, tk), dy_lo(b.un, uk), dy_lo(b.vn, vk) };
                    if (!has_k2 || key_less(key, k2)) {
                        const double w[3] = { pow2neg(tk), pow2neg(uk), pow2neg(vk) };
                        int split = -1;
                        double bestr = -1.0;
#pragma unroll
                        for (int d = 0; d < 3; ++d)
                            if (w[d] > tol[d]) {
                                const double r = inv_tol[d] * w[d]; // = w[d] / tol[d] bit for bit: w[d] is a power of two (see ti_ccd)
                                if (r > bestr) { bestr = r; split = d; }
                            }
                        const int pk = split == 0 ? tk : (split == 1 ? uk : vk);
                        if (split < 0 || pk >= 60) deep = true;
                        else {
                            const unsigned long long pn = split == 0 ? b.tn : (split == 1 ? b.un : b.vn);
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                const unsigned long long hn = 2 * pn + half;
                                const int hk = pk + 1;
                                bool keep = true;
                                if (split == 0) { if (check_t) keep = !(dy_hi(hn, hk) < 0.0 || dy_lo(hn, hk) > max_t); }
                                else if (vf) keep = (split == 1) ? sum_le_1(hn, hk, b.vn, vk) : sum_le_1(hn, hk, b.un, uk);
                                if (keep) {
                                    DBox ch = b;
                                    ch.kk &= 0x00ffffffu;
                                    if (split == 0) { ch.tn = hn; ch.kk = (ch.kk & ~0xffu) | (unsigned)hk; }
                                    else if (split == 1) { ch.un = hn; ch.kk = (ch.kk & ~0xff00u) | ((unsigned)hk << 8); }
                                    else { ch.vn = hn; ch.kk = (ch.kk & ~0xff0000u) | ((unsigned)hk << 16); }
                                    if (nchild == 0) c0 = ch;
                                    else c1 = ch;
                                    ++nchild;
                                }
                            }
                        }
                    }
                }
            }
            int incl = nchild; // inclusive scan over the 8 lanes of the group
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const int y = __shfl_up_sync(gmask, incl, o, 8);
                if (gl >= o) incl += y;
            }
            const int total = __shfl_sync(gmask, incl, 7, 8);
            const int off = nn + incl - nchild;
            if (nn + total > kGrpCap) over = true;
            else {
                if (nchild > 0) nxt[off] = c0;
                if (nchild > 1) nxt[off + 1] = c1;
            }
            nn += total;
            if (__any_sync(gmask, over || deep)) break;
        }
        if (__any_sync(gmask, deep)) { // bisection depth exhausted: the conservative per-level estimate, like the other passes
            if (gl == 0) atomicAdd(warn, 1);
            toi = temp_toi;
            out_tol = temp_out_tol;
            return 1;
        }
        if (__any_sync(gmask, over)) return 2; // the level outgrew the group's buffer: hand the pair to the warp-level pass
        __syncwarp(gmask);
        DBox* tsw = cur; cur = nxt; nxt = tsw;
        n = nn;
    }
    if (use_skip) { toi = toi_skip; return 1; }
    return 0;
}

// vertexFaceCCD_double / edgeEdgeCCD_double incl. the no_zero_toi refinement loop, for one 8-lane group; 0 / 1 / 2 (deferred)
__device__ int ti_ccd_grp(bool vf, const TiPair& P, const double* err, double ms, double tolerance, double t_max, int max_itr, DBox* sA, DBox* sB, int gl, unsigned gmask,
    double& toi, int* __restrict__ warn, const unsigned long long* best)
{
    double tolerance_in = tolerance, ms_in = ms, out_tol = tolerance;
    bool is_impacting = false, tmp = false;
    unsigned iter = 0;
    do {
        double tol[3];
        width_tolerances(vf, P, tolerance_in, tol);
        double inv_tol[3];
        for (int d = 0; d < 3; ++d) inv_tol[d] = 1.0 / tol[d];
        const int rc = ti_root_finder_grp(vf, P, tol, inv_tol, tolerance_in, t_max, err, ms_in, max_itr, sA, sB, gl, gmask, toi, out_tol, warn, best);
        if (rc == 2) return 2;
        tmp = rc == 1;
        if (iter == 0) is_impacting = tmp;
        else toi = tmp ? toi : t_max;
        if (tmp && toi == 0.0) {
            if (out_tol > tolerance_in) t_max *= 0.9;
            else if (10 * tolerance_in < ms_in) ms_in *= 0.5;
            else tolerance_in *= 0.1;
        }
        ++iter;
    } while (iter < 0x7fffffffu && tmp && toi == 0.0);
    return is_impacting ? 1 : 0;
}

__global__ void __launch_bounds__(32 * kGrpWarpsPerCta, 4) k_ti_groups(NarrowArgs a, const unsigned* __restrict__ survivors, const unsigned* __restrict__ nSurvPtr,
    unsigned* __restrict__ work, unsigned* __restrict__ deferred, unsigned* __restrict__ nDeferred, unsigned long long* __restrict__ min_ord, int* __restrict__ warn)
{
    __shared__ DBox sLevels[kGrpWarpsPerCta][4][2][kGrpCap];
    __shared__ TiPair sPair[kGrpWarpsPerCta][4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane >> 3, gl = lane & 7;
    const unsigned gmask = 0xffu << (8 * grp);
    DBox* sA = sLevels[wib][grp][0];
    DBox* sB = sLevels[wib][grp][1];
    TiPair& Ps = sPair[wib][grp];
    const unsigned nSurv = *nSurvPtr;
    const double max_t = a.st->max_t;
    const unsigned long long* best = &a.st->ccd_ord;
    for (;;) {
        unsigned w = 0;
        if (gl == 0) w = atomicAdd(work, 1u);
        w = __shfl_sync(gmask, w, 0, 8);
        if (w >= nSurv) break;
        const unsigned idx = survivors[w];
        bool vf;
        {
            int v[4];
            TiPair Pl;
            load_pair(a.s, a.dir, a.cand[idx], vf, v, Pl);
            if (gl == 0) Ps = Pl;
            __syncwarp(gmask);
        }
        const TiPair& P = Ps;
        const double d = pair_distance_sqrt(vf, P);
        const double ms = fmin(0.2 * d, 1e-6);
        const double* err = vf_metric(vf, P) ? a.err_vf : a.err_ee;
        double toi;
        int hit = ti_ccd_grp(vf, P, err, ms, a.tol, max_t, a.max_itr, sA, sB, gl, gmask, toi, warn, best);
        if (hit == 1 && toi < 1e-6) { // :759-781 (no pruning here: the result is rescaled by 0.8, see pair_ccd)
            hit = ti_ccd_grp(vf, P, err, 0.0, a.tol, max_t, a.max_itr, sA, sB, gl, gmask, toi, warn, nullptr);
            if (hit == 1) toi *= 0.8;
        }
        if (gl == 0) {
            if (hit == 2) deferred[atomicAdd(nDeferred, 1u)] = idx;
            else if (hit == 1) atomicMin(min_ord, dbl_to_ord(toi));
        }
        __syncwarp(gmask);
    }
}

// stage 1.5: one THREAD per surviving pair with a small private level buffer; pairs whose search outgrows it are deferred
constexpr int kThreadCap = 12;
template <int CAP, int OCC, bool SMEM = false>
__global__ void __launch_bounds__(128, OCC) k_ti_stage15(NarrowArgs a, const unsigned* __restrict__ survivors, const unsigned* __restrict__ nSurvPtr,
    unsigned* __restrict__ deferred, unsigned* __restrict__ nDeferred, long long budget, unsigned long long* __restrict__ min_ord, int* __restrict__ warn)
{
    const unsigned nSurv = *nSurvPtr;
    const unsigned stride = gridDim.x * blockDim.x;
    // grid-stride over whole warps so that the warp-aggregated append below always has all 32 lanes present
    for (unsigned base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < nSurv; base += stride) {
        const unsigned i = base + (threadIdx.x & 31);
        bool defer = false;
        unsigned idx = 0;
        if (i < nSurv) {
            idx = survivors[i];
            bool vf;
            int v[4];
            TiPair P;
            load_pair(a.s, a.dir, a.cand[idx], vf, v, P);
            // level buffers: in LOCAL memory they cost ~20 % of this kernel's stall samples (ncu source view: the flag test and the
            // integer->double conversions right behind the box loads; 100 KB of stack per CTA thrashes the L1).  SMEM = true keeps them in
            // shared memory, one padded slab per thread ((CAP * 32 + 8)-byte stride: the lanes' 64-bit words fall into distinct bank pairs).
            DBox lA[SMEM ? 1 : CAP], lB[SMEM ? 1 : CAP];
            DBox* bufA = lA;
            DBox* bufB = lB;
            if (SMEM) {
                extern __shared__ __align__(16) unsigned char s_lvl[];
                constexpr int kSlab = CAP * (int)sizeof(DBox) + 8;
                bufA = reinterpret_cast<DBox*>(s_lvl + (size_t)threadIdx.x * kSlab);
                bufB = reinterpret_cast<DBox*>(s_lvl + (size_t)(blockDim.x + threadIdx.x) * kSlab);
            }
            double toi;
            const int hit = pair_ccd<1>(vf, P, a, bufA, bufB, CAP, 0, toi, warn, nullptr, nullptr, budget);
            if (hit == 2) defer = true;
            else if (hit == 1) atomicMin(min_ord, dbl_to_ord(toi));
        }
        const unsigned m = __ballot_sync(0xffffffffu, defer);
        if (m) {
            const int lane = threadIdx.x & 31;
            unsigned b0 = 0;
            if (lane == __ffs(m) - 1) b0 = atomicAdd(nDeferred, __popc(m));
            b0 = __shfl_sync(0xffffffffu, b0, __ffs(m) - 1);
            if (defer) deferred[b0 + __popc(m & ((1u << lane) - 1))] = idx;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Thread pass with LANE-LEVEL REFILL (round 2, second half).  ncu on k_ti_stage15: 7.4 of 32 lanes active per issued instruction -- a lane whose
// search ended after two boxes waited for the lane of its warp that used its whole 24-box budget.  Here the search is a resumable state
// machine: one loop iteration = ONE LEVEL of one lane's breadth-first search (the body of ti_root_finder's level loop, same arithmetic, same
// decisions); a lane whose pair is finished takes the next survivor from a work counter in the same iteration.  The no_zero_toi loop of
// ti_ccd and the ms = 0 retry of pair_ccd are states of the same machine.
// ------------------------------------------------------------------------------------------------------------------
struct Bfs1 {
    DBox* cur;
    DBox* nxt;
    int n;
    double toi_skip;
    bool use_skip;
    long long refine;
    double temp_toi, temp_out_tol;
    unsigned long long bo_cur;
};
DEV void bfs1_init(Bfs1& b, DBox* bufA, DBox* bufB, double co_tol, const unsigned long long* best)
{
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    b.cur = bufA; b.nxt = bufB;
    b.cur[0] = DBox{ 0ull, 0ull, 0ull, 0u, 0u };
    b.n = 1;
    b.toi_skip = INF; b.use_skip = false; b.refine = 0;
    b.temp_toi = INF; b.temp_out_tol = co_tol;
    b.bo_cur = best ? *reinterpret_cast<const volatile unsigned long long*>(best) : 0ull;
}
// one level of ti_root_finder<1>: 0 no collision, 1 collision (toi set), 2 deferred, 3 go on with the next level
DEV int bfs1_level(bool VF, const TiPair& P, const double* tol, const double* inv_tol, double co_tol, double max_t, const double* err, double ms, int max_itr,
    int cap, long long thread_budget, const unsigned long long* best, Bfs1& b, double& toi, double& out_tol, int* __restrict__ warn)
{
    const bool check_t = (max_t != 1.0);
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    double t_prune = INF;
    if (best) {
        t_prune = fmax(ord_to_dbl(b.bo_cur), 1e-6);
        b.bo_cur = *reinterpret_cast<const volatile unsigned long long*>(best);
    }
    Key3 k1 = { INF, INF, INF }, k2 = { INF, INF, INF };
    unsigned p1 = 0, p2 = 0;
    double a1 = 0.0;
    int visited = 0;
    for (int i = 0; i < b.n; ++i) {
        DBox bx = b.cur[i];
        const int tk = bx.kk & 0xff, uk = (bx.kk >> 8) & 0xff, vk = (bx.kk >> 16) & 0xff;
        const double tlo = dy_lo(bx.tn, tk);
        unsigned flags = 0;
        if (tlo < b.toi_skip && tlo < t_prune) {
            ++visited;
            bool box_in;
            double tt[3];
            if (origin_in_box(VF, P, bx, err, ms, box_in, tt)) {
                flags = F_ZERO;
                const bool tol_cond = tt[0] <= co_tol && tt[1] <= co_tol && tt[2] <= co_tol;
                const bool cond1 = pow2neg(tk) <= tol[0] && pow2neg(uk) <= tol[1] && pow2neg(vk) <= tol[2];
                const Key3 key = { tlo, dy_lo(bx.un, uk), dy_lo(bx.vn, vk) };
                const unsigned mp1 = (tol_cond || box_in || cond1) ? 1u : 0u;
                if (key_less(key, k1)) { k1 = key; p1 = mp1; a1 = fmax(fmax(tt[0], tt[1]), tt[2]); }
                if (mp1 && key_less(key, k2)) { k2 = key; p2 = cond1 ? 1u : 0u; }
            }
        }
        b.cur[i].kk = (bx.kk & 0x00ffffffu) | flags;
    }
    if (k1.t == INF) { // nothing at this level contains the origin: the search space is exhausted
        if (b.use_skip) { toi = b.toi_skip; return 1; }
        return 0;
    }
    if (p1 & 1u) { toi = k1.t; return 1; }
    const bool has_k2 = k2.t != INF;
    if (has_k2 && (p2 & 1u)) { toi = k2.t; return 1; }
    atomicAdd(reinterpret_cast<unsigned long long*>(warn + 5), (unsigned long long)visited); // diagnostics: boxes evaluated by the thread pass
    if (b.refine + visited > thread_budget) return 2; // over budget: handed to the warp pass
    if (max_itr > 0) {
        b.temp_toi = k1.t;
        b.temp_out_tol = fmax(a1, co_tol);
        b.refine += visited;
        if (b.refine > max_itr) {
            atomicAdd(warn, 1);
            toi = b.temp_toi;
            out_tol = b.temp_out_tol;
            return 1;
        }
    }
    if (has_k2) {
        if (k2.t < b.toi_skip) b.toi_skip = k2.t;
        b.use_skip = true;
    }
    int nn = 0;
    for (int i = 0; i < b.n; ++i) {
        const DBox bx = b.cur[i];
        if (!(bx.kk & F_ZERO)) continue;
        const int tk = bx.kk & 0xff, uk = (bx.kk >> 8) & 0xff, vk = (bx.kk >> 16) & 0xff;
        const Key3 key = { dy_lo(bx.tn, tk), dy_lo(bx.un, uk), dy_lo(bx.vn, vk) };
        if (has_k2 && !key_less(key, k2)) continue;
        const double w[3] = { pow2neg(tk), pow2neg(uk), pow2neg(vk) };
        int split = -1;
        double bestr = -1.0;
#pragma unroll
        for (int d = 0; d < 3; ++d)
            if (w[d] > tol[d]) {
                const double r = inv_tol[d] * w[d]; // = w[d] / tol[d] bit for bit (see ti_ccd)
                if (r > bestr) { bestr = r; split = d; }
            }
        const int pk = split == 0 ? tk : (split == 1 ? uk : vk);
        if (split < 0 || pk >= 60) return 2; // bisection overflow: the warp pass handles it like the iteration overflow
        const unsigned long long pn = split == 0 ? bx.tn : (split == 1 ? bx.un : bx.vn);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const unsigned long long hn = 2 * pn + half;
            const int hk = pk + 1;
            bool keep = true;
            if (split == 0) { if (check_t) keep = !(dy_hi(hn, hk) < 0.0 || dy_lo(hn, hk) > max_t); }
            else if (VF) keep = (split == 1) ? sum_le_1(hn, hk, bx.vn, vk) : sum_le_1(hn, hk, bx.un, uk);
            if (keep) {
                if (nn >= cap) return 2; // the level outgrew the buffer
                DBox c = bx;
                c.kk &= 0x00ffffffu;
                if (split == 0) { c.tn = hn; c.kk = (c.kk & ~0xffu) | (unsigned)hk; }
                else if (split == 1) { c.un = hn; c.kk = (c.kk & ~0xff00u) | ((unsigned)hk << 8); }
                else { c.vn = hn; c.kk = (c.kk & ~0xff0000u) | ((unsigned)hk << 16); }
                b.nxt[nn++] = c;
            }
        }
    }
    DBox* t = b.cur; b.cur = b.nxt; b.nxt = t;
    b.n = nn;
    if (nn == 0) {
        if (b.use_skip) { toi = b.toi_skip; return 1; }
        return 0;
    }
    return 3;
}

template <int CAP>
__global__ void __launch_bounds__(128, 2) k_ti_stage15_refill(NarrowArgs a, const unsigned* __restrict__ survivors, const unsigned* __restrict__ nSurvPtr, unsigned* __restrict__ work,
    unsigned* __restrict__ deferred, unsigned* __restrict__ nDeferred, long long budget, unsigned long long* __restrict__ min_ord, int* __restrict__ warn, int refill_batch)
{
    extern __shared__ __align__(16) unsigned char s_lvl[];
    constexpr int kSlab = CAP * (int)sizeof(DBox) + 8;
    DBox* bufA = reinterpret_cast<DBox*>(s_lvl + (size_t)threadIdx.x * kSlab);
    DBox* bufB = reinterpret_cast<DBox*>(s_lvl + (size_t)(blockDim.x + threadIdx.x) * kSlab);
    const unsigned nSurv = *nSurvPtr;
    const double max_t0 = a.st->max_t;
    const unsigned long long* best0 = &a.st->ccd_ord;
    const int lane = threadIdx.x & 31;
    // per-lane state of pair_ccd / ti_ccd / the root finder
    bool busy = false, vf = false;
    unsigned idx = 0;
    TiPair P;
    Bfs1 bf;
    int attempt = 0;
    unsigned iter = 0;
    bool is_impacting = false;
    double dist = 0.0, ms_in = 0.0, tolerance_in = 0.0, t_max = 0.0, out_tol = 0.0, toi = 0.0;
    double tol[3] = { 0, 0, 0 }, inv_tol[3] = { 0, 0, 0 };
    const double* err = a.err_vf;
    const unsigned long long* best = nullptr;
    bool drained = false;
    for (;;) {
        // Refill in batches: fetching a pair is two dependent rounds of global loads (candidate -> vertex ids -> positions, directions), and the
        // whole warp waits for them -- refilling whenever any lane is idle made every iteration pay that latency (ncu: long scoreboard 3.0 per
        // issued instruction).  Lanes wait until kRefillBatch of them are idle (or nothing is running).
        __syncwarp();
        const unsigned idle_m = __ballot_sync(0xffffffffu, !busy && !drained);
        const bool refill_now = __popc(idle_m) >= refill_batch || __ballot_sync(0xffffffffu, busy) == 0u;
        if (!busy && !drained && refill_now) { // take the next survivor (one atomic per warp and round)
            const unsigned m = __activemask();
            const int leader = __ffs(m) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(work, (unsigned)__popc(m));
            base = __shfl_sync(m, base, leader);
            const unsigned w = base + __popc(m & ((1u << lane) - 1u));
            if (w < nSurv) {
                idx = survivors[w];
                int v[4];
                load_pair(a.s, a.dir, a.cand[idx], vf, v, P);
                err = vf_metric(vf, P) ? a.err_vf : a.err_ee;
                dist = pair_distance_sqrt(vf, P);
                attempt = 0;
                // ti_ccd entry state of the first trip (pair_ccd): ms = min(0.2 d, 1e-6), pruned against the running minimum
                ms_in = fmin(0.2 * dist, 1e-6); tolerance_in = a.tol; t_max = max_t0; out_tol = a.tol; iter = 0; is_impacting = false; best = best0;
                width_tolerances(vf, P, tolerance_in, tol);
#pragma unroll
                for (int d = 0; d < 3; ++d) inv_tol[d] = 1.0 / tol[d];
                bfs1_init(bf, bufA, bufB, tolerance_in, best);
                out_tol = tolerance_in;
                busy = true;
            }
            else drained = true;
        }
        __syncwarp();
        if (__all_sync(0xffffffffu, !busy)) break; // the lanes of a warp leave together: all idle and the list exhausted
        if (busy) {
            const int rc = bfs1_level(vf, P, tol, inv_tol, tolerance_in, t_max, err, ms_in, a.max_itr, CAP, budget, best, bf, toi, out_tol, warn);
            if (rc != 3) {
                bool restart = false, finished = false;
                int hit = 0;
                if (rc == 2) { deferred[atomicAdd(nDeferred, 1u)] = idx; finished = true; }
                else { // ti_ccd: the no_zero_toi refinement loop around the root finder
                    const bool tmp = rc == 1;
                    if (iter == 0) is_impacting = tmp;
                    else toi = tmp ? toi : t_max;
                    if (tmp && toi == 0.0) {
                        if (out_tol > tolerance_in) t_max *= 0.9;
                        else if (10 * tolerance_in < ms_in) ms_in *= 0.5;
                        else tolerance_in *= 0.1;
                        ++iter;
                        restart = true;
                    }
                    else {
                        hit = is_impacting ? 1 : 0;
                        // pair_ccd: second trip (:759-781) only after a hit with toi < 1e-6: ms = 0, not pruned, result scaled by 0.8
                        if (attempt == 0 && hit && toi < 1e-6) {
                            attempt = 1;
                            ms_in = 0.0; tolerance_in = a.tol; t_max = max_t0; iter = 0; is_impacting = false; best = nullptr;
                            restart = true;
                        }
                        else {
                            if (attempt == 1 && hit) toi *= 0.8;
                            finished = true;
                        }
                    }
                }
                if (restart) {
                    width_tolerances(vf, P, tolerance_in, tol);
#pragma unroll
                    for (int d = 0; d < 3; ++d) inv_tol[d] = 1.0 / tol[d];
                    bfs1_init(bf, bufA, bufB, tolerance_in, best);
                    out_tol = tolerance_in;
                }
                if (finished) {
                    if (hit == 1) atomicMin(min_ord, dbl_to_ord(toi));
                    busy = false;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(128, 3) k_ti_stage2(NarrowArgs a, const unsigned* __restrict__ survivors, const unsigned* __restrict__ nSurvPtr, unsigned* __restrict__ work,
    DBox* __restrict__ scratch, int cap, unsigned long long* __restrict__ min_ord, int* __restrict__ warn)
{
    __shared__ DBox sLevels[kStage2WarpsPerCtaDev][2][kSmemLevel];
    __shared__ TiPair sPair[kStage2WarpsPerCtaDev];
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    DBox* sA = sLevels[threadIdx.x >> 5][0];
    DBox* sB = sLevels[threadIdx.x >> 5][1];
    DBox* bufA = scratch + (size_t)warp_global * 2 * cap;
    DBox* bufB = bufA + cap;
    const unsigned nSurv = *nSurvPtr;
    for (;;) {
        unsigned w = 0;
        if (lane == 0) w = atomicAdd(work, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= nSurv) break;
        bool vf;
        {
            // the pair's 24 coordinates live in shared memory: the warp-level search only needs them 8 at a time per lane
            int v[4];
            TiPair Pl;
            load_pair(a.s, a.dir, a.cand[survivors[w]], vf, v, Pl);
            if (lane == 0) sPair[threadIdx.x >> 5] = Pl;
            __syncwarp();
        }
        const TiPair& P = sPair[threadIdx.x >> 5];
        double toi;
        const long long t0 = clock64();
        const int hit = pair_ccd<32>(vf, P, a, bufA, bufB, cap, lane, toi, warn, sA, sB);
        if (hit && lane == 0) atomicMin(min_ord, dbl_to_ord(toi));
        if (lane == 0) { // diagnostics: the longest pair bounds this pass from below
            const unsigned dt = (unsigned)((clock64() - t0) >> 6);
            atomicMax(reinterpret_cast<unsigned*>(warn + 3), dt);
            atomicAdd(reinterpret_cast<unsigned*>(warn + 4), dt);
        }
        __syncwarp();
    }
}

// start of a narrow phase: snapshot the device-resident step as max_t and as the initial running minimum, fix this rank's slice of the
// candidate list (n32 / n64: the list size, wherever the producing stage left it), clear the per-phase counters
__global__ void k_ccd_init(IterState* st, const int* __restrict__ n32, const unsigned long long* __restrict__ n64, unsigned long long cap, int rank, int nranks, int share,
    double seed, unsigned* nSurv, unsigned* work, int* flags)
{
    if (threadIdx.x != 0) return;
    const double alpha = ord_to_dbl(st->step_ord);
    st->max_t = alpha;
    st->ccd_ord = (seed >= 0.0 && seed < alpha) ? dbl_to_ord(seed) : st->step_ord; // (seed: test hook, ipcgpu_ccd_debug_seed_bound)
    unsigned long long n = n32 ? (unsigned long long)max(*n32, 0) : *n64;
    if (n > cap) n = cap; // an overflowing producer raised its capacity flag; stay inside the list
    st->cand_range[0] = share ? n * rank / nranks : 0ull;
    st->cand_range[1] = share ? n * (rank + 1) / nranks : n;
    *nSurv = 0;
    *work = 0;
    for (int q = 0; q < 12; ++q) flags[q] = 0; // zero distance, warnings, deferred count, pad, boxes(thread pass) x2, boxes(warp pass) x2
}
// end of a narrow phase (this rank): a zero initial distance forces the step to 0 (:730-737); diagnostics go to the iteration state
__global__ void k_ccd_finish(IterState* st, const unsigned* __restrict__ nSurv, const int* __restrict__ flags, const int* __restrict__ overflow, int is_full)
{
    if (threadIdx.x != 0) return;
    if (flags[0]) {
        st->ccd_ord = 0ull;
        st->flags[FLAG_ZERO_CCD_DISTANCE] = 1;
    }
    if (overflow && *overflow) st->flags[FLAG_CCD_CAPACITY] = 1;
    st->flags[FLAG_TI_WARNINGS] += flags[1];
    st->ccd_stats[0] = *nSurv;
    st->ccd_stats[1] = (unsigned long long)flags[1];
    st->ccd_stats[2] = (unsigned)flags[2];
    st->ccd_stats[3] = (unsigned long long)(unsigned)flags[4] << 6;
    st->ccd_stats[4] = (unsigned long long)(unsigned)flags[5] << 6;
    st->ccd_stats[5] = *reinterpret_cast<const unsigned long long*>(flags + 6);
    st->ccd_stats[6] = *reinterpret_cast<const unsigned long long*>(flags + 8);
    st->ccd_stats[7] = st->cand_range[1] - st->cand_range[0];
    if (is_full) st->n_full_cand = st->ccd_stats[7];
}
// after the cross-rank min: the narrow phase's minimum becomes the step
__global__ void k_ccd_commit(IterState* st, int stage)
{
    if (threadIdx.x != 0) return;
    st->step_ord = st->ccd_ord;
    st->alpha_stage[stage] = ord_to_dbl(st->ccd_ord);
}

} // namespace ipcgpu

using namespace ipcgpu;

#define CKD(call)                                                      \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) {                                       \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); \
            return IPCGPU_ERR_CUDA;                                    \
        }                                                              \
    } while (0)

static inline int nblk(long long n, int b) { return (int)((n + b - 1) / b); }
SurfArgs surf_args(const ipcgpu_ctx* ctx); // constraint.cu
SortedGrid tri_grid(const ipcgpu_ctx* ctx);
SortedGrid edge_grid(const ipcgpu_ctx* ctx);
int boxes_and_grid(ipcgpu_ctx* ctx, const double* dir, const double* alpha_ptr, double radius, const double* radius_ptr, bool with_vertex_boxes, const int* vmin,
    const int* vmax); // constraint.cu
int pairs_mode();                                                                                                                                  // constraint.cu
void cell_pairs_ee(const ipcgpu::Grid* gp, const ipcgpu::SortedGrid& eg, const ipcgpu::SurfArgs& s, double radius_val, const double* radius_ptr, const ipcgpu::IterState* vox,
    int first, int last, const ipcgpu::PairOut& out, cudaStream_t st);
void cell_pairs_pt(const ipcgpu::Grid* gp, const ipcgpu::SortedGrid& vg, const ipcgpu::SortedGrid& tg, const ipcgpu::SurfArgs& s, double radius_val, const double* radius_ptr,
    const ipcgpu::IterState* vox, int first, int last, const ipcgpu::PairOut& out, cudaStream_t st);
SortedGrid vertex_grid(const ipcgpu_ctx* ctx);

static bool refill_mode() // thread pass with lane-level refill (IPCGPU_TI_REFILL=0: one pair per lane and grid-stride round)
{
    static const bool v = [] { const char* e = std::getenv("IPCGPU_TI_REFILL"); return e ? std::atoi(e) != 0 : true; }();
    return v;
}
static bool lvl_smem() // level buffers of the thread pass in shared memory (IPCGPU_TI_LVL_SMEM=0: thread-local memory)
{
    static const bool v = [] { const char* e = std::getenv("IPCGPU_TI_LVL_SMEM"); return e ? std::atoi(e) != 0 : true; }();
    return v;
}
constexpr int kStage2WarpsPerCta = 4;
constexpr int kStage2Ctas = 148 * 6; // persistent: 6 CTAs x 4 warps per SM
constexpr int kLevelCap = 4096;      // boxes per BFS level buffer (2 buffers per warp)

int ccd_alloc(ipcgpu_ctx* ctx)
{
    CcdWork& w = ctx->ccd;
    const size_t warps = (size_t)kStage2Ctas * kStage2WarpsPerCta;
    bool ok = w.vmin.reserve((size_t)3 * ctx->nV) && w.vmax.reserve((size_t)3 * ctx->nV) && w.cand.reserve(ctx->ccd_capacity) && w.surv.reserve(ctx->ccd_capacity) && w.surv2.reserve(ctx->ccd_capacity)
        && w.scratch.reserve(warps * 2 * kLevelCap * sizeof(DBox)) && w.counters.reserve(16) && w.ncand.reserve(2) && w.bounds.reserve(8);
    if (!ok) {
        ctx->err = "CCD workspace allocation failed";
        return IPCGPU_ERR_CUDA;
    }
    return 0;
}

// cross-rank min of the running minimum (api.cu; no-op on one rank)
int nccl_min_u64(ipcgpu_ctx* ctx, unsigned long long* word);
int fetch_iter_state(ipcgpu_ctx* ctx); // api.cu: one D2H copy of the iteration state + stream synchronisation

// Narrow phase over a device-resident candidate list.  The step is read from and written back to the device-resident iteration state
// (IterState::step_ord): nothing is read back here.  n32 / n64: device-resident list size; share: walk only this rank's contiguous
// slice of the list (replicated lists) or all of it (lists that are already this rank's own); stage: 1 = partial, 3 = full CCD.
int ccd_narrow(ipcgpu_ctx* ctx, const int2* cand, const int* n32, const unsigned long long* n64, unsigned long long cap, int share, double tol, const double* err_vf,
    const double* err_ee, int stage, const int* overflow)
{
    CcdWork& w = ctx->ccd;
    cudaStream_t st = ctx->stream;
    NarrowArgs a;
    a.s = surf_args(ctx);
    a.dir = ctx->dir.p;
    a.cand = cand;
    for (int c = 0; c < 3; ++c) { a.err_vf[c] = err_vf[c]; a.err_ee[c] = err_ee[c]; }
    a.tol = tol;
    a.max_itr = 1000000;    // TIGHT_INCLUSION_MAX_ITER (CCDUtils.hpp:14)
    a.st = ctx->iter.p;
    IterState* ist = ctx->iter.p;
    unsigned* nSurv = reinterpret_cast<unsigned*>(w.counters.p);
    unsigned* work = nSurv + 1;
    int* flags = w.counters.p + 2; // [0] zero distance, [1] warnings, [2] deferred, [4] longest, [5] total cycles, [6..7] / [8..9] boxes
    static const int wide_env = [] { const char* e = std::getenv("IPCGPU_TI_WIDE_LEVEL"); return e ? std::atoi(e) : 0; }();
    if (wide_env > 0 && !w.wide_level_set) {
        CKD(cudaMemcpyToSymbolAsync(c_wide_level, &wide_env, sizeof(int), 0, cudaMemcpyHostToDevice, st));
        w.wide_level_set = true;
    }
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_CCD_NARROW);
    k_ccd_init<<<1, 32, 0, st>>>(ist, n32, n64, cap, ctx->rank, ctx->nranks, share, ctx->debug_prune_seed, nSurv, work, flags);
    {
        // the candidate count lives on the device: every pass is a grid-stride / persistent kernel that reads it there
        // pass 1 (thread per candidate): the root box only;
        // pass A (thread per survivor, 10-box budget): the shallow majority dies here without holding its warp hostage;
        // pass B (warp per pair, corner-parallel box evaluation): the deep searches, compacted.
        // (a three-tier split -- 3-box, then 12-box thread passes -- was measured slower: the per-pair set-up dominates pass A)
        cudaEvent_t pe1 = ctx->prof_begin(IPCGPU_STAGE_CCD_ROOT_FILTER);
        k_ti_stage1<<<kSMs * 16, 128, 0, st>>>(a, w.surv.p, nSurv, flags);
        ctx->prof_end(pe1);
        unsigned* nDefA = reinterpret_cast<unsigned*>(flags + 2);
        // pass A (thread per survivor, 24-box budget): the shallow majority (2-3 boxes) at 32 pairs per warp;
        // pass G (8 lanes per pair, 4 pairs per warp): the searches pass A gave up on, unless a level outgrows the group's 44-box buffer;
        // pass B (warp per pair): those wide searches.  IPCGPU_TI_MODE: 0 = A + B (default), 1 = G + B, 2 = A + G + B, 3 = A + A(48-box levels) + B.
        // Measured on C5 (narrow phase per iteration): A + B 1.62 ms, G + B 2.02 ms, A + G + B 2.12 ms -- the four groups of a warp diverge and
        // the hardware runs divergent paths of one warp one after the other, so pass G buys no latency hiding (kept for the record)
        static const int ti_mode = [] { const char* e = std::getenv("IPCGPU_TI_MODE"); return e ? std::atoi(e) : 0; }();
        unsigned* grp_work = reinterpret_cast<unsigned*>(flags + 3);
        unsigned* nDefB = reinterpret_cast<unsigned*>(flags + 10);
        if (ti_mode == 1) {
            k_ti_groups<<<kSMs * 4, 32 * kGrpWarpsPerCta, 0, st>>>(a, w.surv.p, nSurv, grp_work, w.surv2.p, nDefA, &ist->ccd_ord, flags + 1);
            k_ti_stage2<<<kStage2Ctas, 32 * kStage2WarpsPerCta, 0, st>>>(a, w.surv2.p, nDefA, work, reinterpret_cast<DBox*>(w.scratch.p), kLevelCap, &ist->ccd_ord, flags + 1);
        }
        else {
            static const long long budgetA = [] { const char* e = std::getenv("IPCGPU_TI_BUDGET"); return e ? std::atoll(e) : 24ll; }(); // boxes a thread may evaluate before it hands its pair on
            // (C5 after the code-size refactor, narrow phase per iteration: 10 -> 1.33 ms, 16 -> 1.11, 24 -> 1.05, 32 -> 1.09, 64 -> 1.23, 128 -> 1.42)
            static const int occA = [] { const char* e = std::getenv("IPCGPU_TI_OCC"); return e ? std::atoi(e) : 2; }(); // CTAs/SM the thread pass is compiled for (2: 255 regs, 3: 168 regs + spills)
            static const int capA = [] { const char* e = std::getenv("IPCGPU_TI_CAP"); return e ? std::atoi(e) : 8; }(); // boxes per level buffer of the thread pass
            // (C5, narrow phase per iteration: local memory 12 boxes 1.044 ms, 8 boxes 1.035, 6 boxes 1.024; shared memory 12 boxes 1.029, 8 boxes 1.006)
            if (occA == 3) k_ti_stage15<kThreadCap, 3><<<kSMs * 16, 128, 0, st>>>(a, w.surv.p, nSurv, w.surv2.p, nDefA, budgetA, &ist->ccd_ord, flags + 1);
            else if (refill_mode()) {
                constexpr int bytes = 2 * 128 * (8 * (int)sizeof(DBox) + 8);
                static bool attr = false;
                if (!attr) { cudaFuncSetAttribute(k_ti_stage15_refill<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); attr = true; }
                static const int batch = [] { const char* e = std::getenv("IPCGPU_TI_REFILL_BATCH"); return e ? std::atoi(e) : 16; }(); // (1 .. 32 measured within 1.5 % of each other)
                k_ti_stage15_refill<8><<<kSMs * 2, 128, bytes, st>>>(a, w.surv.p, nSurv, grp_work, w.surv2.p, nDefA, budgetA, &ist->ccd_ord, flags + 1, batch);
            }
            else if (lvl_smem() && capA == 8) {
                constexpr int bytes = 2 * 128 * (8 * (int)sizeof(DBox) + 8);
                static bool attr = false;
                if (!attr) { cudaFuncSetAttribute(k_ti_stage15<8, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); attr = true; }
                k_ti_stage15<8, 2, true><<<kSMs * 16, 128, bytes, st>>>(a, w.surv.p, nSurv, w.surv2.p, nDefA, budgetA, &ist->ccd_ord, flags + 1);
            }
            else if (lvl_smem() && capA == kThreadCap) {
                constexpr int bytes = 2 * 128 * (kThreadCap * (int)sizeof(DBox) + 8);
                static bool attr = false;
                if (!attr) { cudaFuncSetAttribute(k_ti_stage15<kThreadCap, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); attr = true; }
                k_ti_stage15<kThreadCap, 2, true><<<kSMs * 16, 128, bytes, st>>>(a, w.surv.p, nSurv, w.surv2.p, nDefA, budgetA, &ist->ccd_ord, flags + 1);
            }
            else if (capA == 8) k_ti_stage15<8, 2><<<kSMs * 16, 128, 0, st>>>(a, w.surv.p, nSurv, w.surv2.p, nDefA, budgetA, &ist->ccd_ord, flags + 1);
            else if (capA == 6) k_ti_stage15<6, 2><<<kSMs * 16, 128, 0, st>>>(a, w.surv.p, nSurv, w.surv2.p, nDefA, budgetA, &ist->ccd_ord, flags + 1);
            else if (capA == 16) k_ti_stage15<16, 2><<<kSMs * 16, 128, 0, st>>>(a, w.surv.p, nSurv, w.surv2.p, nDefA, budgetA, &ist->ccd_ord, flags + 1);
            else k_ti_stage15<kThreadCap, 2><<<kSMs * 16, 128, 0, st>>>(a, w.surv.p, nSurv, w.surv2.p, nDefA, budgetA, &ist->ccd_ord, flags + 1);
            if (ti_mode == 2) { // the survivor list is dead after pass A: pass G's own deferrals go there
                k_ti_groups<<<kSMs * 4, 32 * kGrpWarpsPerCta, 0, st>>>(a, w.surv2.p, nDefA, grp_work, w.surv.p, nDefB, &ist->ccd_ord, flags + 1);
                k_ti_stage2<<<kStage2Ctas, 32 * kStage2WarpsPerCta, 0, st>>>(a, w.surv.p, nDefB, work, reinterpret_cast<DBox*>(w.scratch.p), kLevelCap, &ist->ccd_ord, flags + 1);
                ++ctx->launches;
            }
            else if (ti_mode == 3) { // second thread pass over the compacted deferrals: 48-box levels in local memory, 256-box budget
                static const long long budgetA2 = [] { const char* e = std::getenv("IPCGPU_TI_BUDGET2"); return e ? std::atoll(e) : 256ll; }();
                k_ti_stage15<48, 2><<<kSMs * 8, 128, 0, st>>>(a, w.surv2.p, nDefA, w.surv.p, nDefB, budgetA2, &ist->ccd_ord, flags + 1);
                k_ti_stage2<<<kStage2Ctas, 32 * kStage2WarpsPerCta, 0, st>>>(a, w.surv.p, nDefB, work, reinterpret_cast<DBox*>(w.scratch.p), kLevelCap, &ist->ccd_ord, flags + 1);
                ++ctx->launches;
            }
            else
                k_ti_stage2<<<kStage2Ctas, 32 * kStage2WarpsPerCta, 0, st>>>(a, w.surv2.p, nDefA, work, reinterpret_cast<DBox*>(w.scratch.p), kLevelCap, &ist->ccd_ord, flags + 1);
        }
    }
    k_ccd_finish<<<1, 32, 0, st>>>(ist, nSurv, flags, overflow, stage == 3);
    ctx->prof_end(pe);
    ctx->launches += 5;
    CKD(cudaGetLastError());
    int rc = nccl_min_u64(ctx, &ist->ccd_ord); // min over ranks of the step; a zero-distance flag travels as step 0
    if (rc) return rc;
    k_ccd_commit<<<1, 32, 0, st>>>(ist, stage);
    ++ctx->launches;
    return 0;
}

// sync-mode helper shared by the three step-bound entry points: copy the iteration state back and refresh the host-side diagnostics
int ccd_read_back(ipcgpu_ctx* ctx, double* alpha_out)
{
    int rc = fetch_iter_state(ctx);
    if (rc) return rc;
    CcdWork& w = ctx->ccd;
    const IterState& h = *ctx->h_iter;
    w.last_survivors = (unsigned)h.ccd_stats[0];
    w.last_warnings = (int)h.ccd_stats[1];
    w.last_deferred = h.ccd_stats[2];
    w.last_longest_cycles = h.ccd_stats[3];
    w.last_total_cycles = h.ccd_stats[4];
    w.last_boxes_thread = h.ccd_stats[5];
    w.last_boxes_warp = h.ccd_stats[6];
    w.last_candidates = h.ccd_stats[7];
    if (alpha_out) std::memcpy(alpha_out, &h.step_ord, sizeof(double));
    return 0;
}

// ---- swept hash build: SpatialHash::build(mesh, searchDir, curMaxStepSize, voxelSize) ------------------------------------
// Entirely on the stream: the span rescale of the step (:603-618), the swept bbox, the reference grid geometry and the coarse
// accelerator grid all take their inputs from the device-resident iteration state.
int ccd_build_swept(ipcgpu_ctx* ctx, double h)
{
    CcdWork& w = ctx->ccd;
    cudaStream_t st = ctx->stream;
    const SurfArgs s = surf_args(ctx);
    IterState* ist = ctx->iter.p;
    if (!ctx->dir_valid) {
        ctx->err = "no search direction: pass p (or call ipcgpu_set_search_dir first)";
        return IPCGPU_ERR_STATE;
    }
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_CCD_BROAD);
    k_swept_alpha<<<1, 32, 0, st>>>(ist, ctx->pSize_dev.p, h, w.bounds.p);
    // bbox of V and of the displaced surface vertices (:627-628)
    k_swept_bounds<<<nblk(std::max(s.nV, s.nSV), 256), 256, 0, st>>>(s, ctx->dir.p, ist, w.bounds.p);
    k_refgrid_params<<<1, 32, 0, st>>>(ist, w.bounds.p, h);
    if (s.nSV > 0) k_ref_ranges<<<nblk(s.nSV, 256), 256, 0, st>>>(s, ctx->dir.p, ist, w.vmin.p, w.vmax.p);
    ctx->launches += 4;
    // coarse accelerator grid over the swept boxes; pairs sharing a reference voxel are at most one voxel apart per axis
    int rc = boxes_and_grid(ctx, ctx->dir.p, &ist->alpha_grid, 0.0, &ist->radius, true, w.vmin.p, w.vmax.p);
    ctx->prof_end(pe);
    if (rc) return rc;
    CKD(cudaGetLastError());
    w.swept_ready = true;
    return 0;
}

int ccd_full(ipcgpu_ctx* ctx, double tol, const double* err_vf, const double* err_ee)
{
    CcdWork& w = ctx->ccd;
    ContactWork& cw = ctx->cw;
    cudaStream_t st = ctx->stream;
    const SurfArgs s = surf_args(ctx);
    IterState* ist = ctx->iter.p;
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_CCD_BROAD);
    CKD(cudaMemsetAsync(w.ncand.p, 0, 2 * sizeof(unsigned long long), st));
    CKD(cudaMemsetAsync(w.counters.p + 14, 0, sizeof(int), st));
    CandOut out{ w.cand.p, w.ncand.p, (unsigned long long)ctx->ccd_capacity, w.counters.p + 14 };
    // multi-GPU: every rank sweeps a contiguous share of the query primitives (the reference's own loop decomposition, :1385, :1498)
    const int v0 = (int)((long long)s.nSV * ctx->rank / ctx->nranks), v1 = (int)((long long)s.nSV * (ctx->rank + 1) / ctx->nranks);
    const int e0 = (int)((long long)s.nSE * ctx->rank / ctx->nranks), e1 = (int)((long long)s.nSE * (ctx->rank + 1) / ctx->nranks);
    const SortedGrid tg = tri_grid(ctx), eg = edge_grid(ctx);
    CKD(cudaMemsetAsync(cw.counters.p + 8, 0, 2 * sizeof(int), st));
    unsigned* nPairs = reinterpret_cast<unsigned*>(cw.counters.p + 8);
    PairOut ppt{ cw.bp_pairs.p, nPairs, (unsigned)cw.bp_cap, w.counters.p + 14 }, pee{ cw.bp_pairs.p + cw.bp_cap, nPairs + 1, (unsigned)cw.bp_cap, w.counters.p + 14 };
    if (v1 > v0 && s.nSF > 0) {
        if (pairs_mode() == 0 || cw.built_vertices != s.nSV) k_ccd_pairs_pt<<<nblk(v1 - v0, 8 * kPairQueriesPerWarp), 256, 0, st>>>(cw.grid.p, cw.vbox.p, tg, ist, v0, v1, ppt);
        else cell_pairs_pt(cw.grid.p, vertex_grid(ctx), tg, s, 0.0, &ist->radius, cw.built_voxel_entries ? ist : nullptr, s.nSF + s.nSE + v0, s.nSF + s.nSE + v1, ppt, st);
        k_ccd_filter_pt<<<kSMs * 8, 256, 0, st>>>(s, ppt.pairs, ppt.n, ppt.cap, w.vmin.p, w.vmax.p, out);
        ctx->launches += 2;
    }
    if (e1 > e0 && s.nSE > 1) {
        if (pairs_mode() == 0) k_ccd_pairs_ee<<<nblk(e1 - e0, 8 * kPairQueriesPerWarp), 256, 0, st>>>(cw.grid.p, eg, cw.ebox.p, ist, s.nSF + e0, s.nSF + e1, pee); // edge entries sit behind the triangles
        else cell_pairs_ee(cw.grid.p, eg, s, 0.0, &ist->radius, cw.built_voxel_entries ? ist : nullptr, s.nSF + e0, s.nSF + e1, pee, st);
        k_ccd_filter_ee<<<kSMs * 8, 256, 0, st>>>(s, pee.pairs, pee.n, pee.cap, w.vmin.p, w.vmax.p, cw.ebox.p, out);
        ctx->launches += 2;
    }
    ctx->prof_end(pe);
    // the candidate list is this rank's own (its share of the queries produced it): no further slicing; the list size and the
    // capacity flag stay on the device (checked at ipcgpu_fetch_iteration / by the synchronous wrapper)
    return ccd_narrow(ctx, w.cand.p, nullptr, w.ncand.p, (unsigned long long)ctx->ccd_capacity, 0, tol, err_vf, err_ee, 3, w.counters.p + 14);
}
