// safeguard.cu -- line-search safeguards on the device (sm_100a): element inversion check and edge-triangle intersection check.
//   *** compiled with --fmad=false (NOFMA_FILES): the floating-point decisions (filter, full-pivot solve) round like the CPU oracle ***
//
// Reference being replaced ("a whole line-search trial is one stream with no host round trip"):
//   Mesh<3>::checkInversion(bool mute)                         src/Mesh.cpp:715-763
//   SelfCollisionHandler<3>::checkEdgeTriIntersectionIfAny     src/CollisionObject/SelfCollisionHandler.cpp:3254-3296
//   IglUtils::segTriIntersect (USE_PREDICATES, CMakeLists:139) src/Utils/IglUtils.hpp:214-265
//   SpatialHash::queryTriangleForEdges                         src/Utils/SpatialHash.hpp:526-556   (here: the sort-based grid of broadphase.cuh)
//
// igl::predicates::orient3d is an exact predicate (Shewchuk).  Device version: the usual floating-point filter; inputs it cannot decide
// (a handful per million tests) go through an exact evaluation with floating-point expansions held in thread-local memory
// (two-sum, two-product via the explicit fma intrinsic, grow-expansion, scale-expansion -- Shewchuk 1997, figs. 6, 7, 13).
#include "broadphase.cuh"
#include "context.h"
#include "../../include/ipcgpu.h"

namespace ipcgpu {

// ---- exact arithmetic on expansions ------------------------------------------------------------------------------------------
DEV void two_sum(double a, double b, double& x, double& y)
{
    x = a + b;
    const double bv = x - a, av = x - bv;
    y = (a - av) + (b - bv);
}
DEV void two_prod(double a, double b, double& x, double& y)
{
    x = a * b;
    y = __fma_rn(a, b, -x); // exact rounding error of the product
}
// h (n components, nonoverlapping, increasing magnitude) += b, in place; returns the new length (<= n + 1); zeros are dropped
__device__ __noinline__ int grow_inplace(int n, double* h, double b)
{
    double q = b;
    int m = 0;
    for (int i = 0; i < n; ++i) {
        double s, r;
        two_sum(q, h[i], s, r); // h[i] is read before h[m] (m <= i) is written
        if (r != 0.0) h[m++] = r;
        q = s;
    }
    if (q != 0.0 || m == 0) h[m++] = q;
    return m;
}
// h = e + f (h must not alias f; capacity ne + nf)
DEV int expansion_sum(int ne, const double* e, int nf, const double* f, double* h)
{
    int n = ne;
    for (int i = 0; i < ne; ++i) h[i] = e[i];
    for (int j = 0; j < nf; ++j) n = grow_inplace(n, h, f[j]);
    return n;
}
// h = e * b (h must not alias e; capacity 2 n)
__device__ __noinline__ int scale_expansion(int n, const double* e, double b, double* h)
{
    double q, hh;
    two_prod(e[0], b, q, hh);
    int m = 0;
    if (hh != 0.0) h[m++] = hh;
    for (int i = 1; i < n; ++i) {
        double t1, t0, s, r;
        two_prod(e[i], b, t1, t0);
        two_sum(q, t0, s, r);
        if (r != 0.0) h[m++] = r;
        two_sum(t1, s, q, r);
        if (r != 0.0) h[m++] = r;
    }
    if (q != 0.0 || m == 0) h[m++] = q;
    return m;
}
// x*y - z*w exactly (<= 4 components)
DEV int prod_diff(double x, double y, double z, double w, double* h)
{
    double a1, a0, b1, b0;
    two_prod(x, y, a1, a0);
    two_prod(z, w, b1, b0);
    h[0] = a0;
    h[1] = a1;
    int n = grow_inplace(2, h, -b0);
    return grow_inplace(n, h, -b1);
}
// exact sign of the 4x4 orientation determinant on the untranslated coordinates (cofactor expansion along z)
__device__ __noinline__ int orient3d_exact(const double* pa, const double* pb, const double* pc, const double* pd)
{
    double ab[4], bc[4], cd[4], da[4], ac[4], bd[4];
    const int nab = prod_diff(pa[0], pb[1], pb[0], pa[1], ab), nbc = prod_diff(pb[0], pc[1], pc[0], pb[1], bc), ncd = prod_diff(pc[0], pd[1], pd[0], pc[1], cd);
    const int nda = prod_diff(pd[0], pa[1], pa[0], pd[1], da), nac = prod_diff(pa[0], pc[1], pc[0], pa[1], ac), nbd = prod_diff(pb[0], pd[1], pd[0], pb[1], bd);
    double nac_[4], nbd_[4], t[12], cda[12], dab[12], abc[12], bcd[12];
    for (int i = 0; i < nac; ++i) nac_[i] = -ac[i];
    for (int i = 0; i < nbd; ++i) nbd_[i] = -bd[i];
    int nt = expansion_sum(ncd, cd, nda, da, t);
    const int ncda = expansion_sum(nt, t, nac, ac, cda);
    nt = expansion_sum(nda, da, nab, ab, t);
    const int ndab = expansion_sum(nt, t, nbd, bd, dab);
    nt = expansion_sum(nab, ab, nbc, bc, t);
    const int nabc = expansion_sum(nt, t, nac, nac_, abc);
    nt = expansion_sum(nbc, bc, ncd, cd, t);
    const int nbcd = expansion_sum(nt, t, nbd, nbd_, bcd);
    double adet[24], bdet[24], cdet[24], ddet[24], det[96];
    const int na = scale_expansion(nbcd, bcd, pa[2], adet), nb = scale_expansion(ncda, cda, -pb[2], bdet);
    const int nc = scale_expansion(ndab, dab, pc[2], cdet), nd = scale_expansion(nabc, abc, -pd[2], ddet);
    int n = expansion_sum(na, adet, nb, bdet, det);
    for (int j = 0; j < nc; ++j) n = grow_inplace(n, det, cdet[j]);
    for (int j = 0; j < nd; ++j) n = grow_inplace(n, det, ddet[j]);
    const double top = det[n - 1]; // the largest-magnitude component carries the sign
    return (top > 0.0) - (top < 0.0);
}
// sign of orient3d(pa, pb, pc, pd): +1 / 0 / -1
DEV int orient3d(const double* pa, const double* pb, const double* pc, const double* pd)
{
    const double adx = pa[0] - pd[0], bdx = pb[0] - pd[0], cdx = pc[0] - pd[0];
    const double ady = pa[1] - pd[1], bdy = pb[1] - pd[1], cdy = pc[1] - pd[1];
    const double adz = pa[2] - pd[2], bdz = pb[2] - pd[2], cdz = pc[2] - pd[2];
    const double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, cdxady = cdx * ady, adxcdy = adx * cdy, adxbdy = adx * bdy, bdxady = bdx * ady;
    const double det = adz * (bdxcdy - cdxbdy) + bdz * (cdxady - adxcdy) + cdz * (adxbdy - bdxady);
    const double permanent = (fabs(bdxcdy) + fabs(cdxbdy)) * fabs(adz) + (fabs(cdxady) + fabs(adxcdy)) * fabs(bdz) + (fabs(adxbdy) + fabs(bdxady)) * fabs(cdz);
    const double errbound = 1.7763568394002505e-15 * permanent; // 16 eps: wider than Shewchuk's (7 + 56 eps) eps, see the oracle
    if (det > errbound) return 1;
    if (-det > errbound) return -1;
    return orient3d_exact(pa, pb, pc, pd);
}

// x = M^-1 b, Gaussian elimination with full pivoting in Eigen::FullPivLU's order (first maximum in column-major order of the
// remaining corner), rank threshold eps * 3 * max pivot; m is the 3x3 in rows
DEV void full_piv_solve3(double (&m)[3][3], const double* b, double* x)
{
    int cq[3] = { 0, 1, 2 };
    double rhs[3] = { b[0], b[1], b[2] };
    double maxpiv = 0.0;
    int nz = 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k >= nz) break;
        int pr = k, pc = k;
        double big = -1.0;
        for (int j = k; j < 3; ++j)
            for (int i = k; i < 3; ++i)
                if (fabs(m[i][j]) > big) { big = fabs(m[i][j]); pr = i; pc = j; }
        if (big == 0.0) { nz = k; break; }
        maxpiv = fmax(maxpiv, big);
        if (pr != k) {
            for (int j = 0; j < 3; ++j) { const double t = m[k][j]; m[k][j] = m[pr][j]; m[pr][j] = t; }
            const double t = rhs[k]; rhs[k] = rhs[pr]; rhs[pr] = t;
        }
        if (pc != k) {
            for (int i = 0; i < 3; ++i) { const double t = m[i][k]; m[i][k] = m[i][pc]; m[i][pc] = t; }
            const int t = cq[k]; cq[k] = cq[pc]; cq[pc] = t;
        }
        for (int i = k + 1; i < 3; ++i) m[i][k] /= m[k][k];
        for (int i = k + 1; i < 3; ++i)
            for (int j = k + 1; j < 3; ++j) m[i][j] -= m[i][k] * m[k][j];
    }
    int rank = 0;
    for (int k = 0; k < nz; ++k)
        if (fabs(m[k][k]) > maxpiv * 2.220446049250313e-16 * 3.0) ++rank;
    for (int i = 1; i < 3; ++i)
        for (int j = 0; j < i; ++j) rhs[i] -= m[i][j] * rhs[j];
    double y[3] = { 0.0, 0.0, 0.0 };
    for (int i = rank - 1; i >= 0; --i) {
        double s = rhs[i];
        for (int j = i + 1; j < rank; ++j) s -= m[i][j] * y[j];
        y[i] = s / m[i][i];
    }
    x[0] = x[1] = x[2] = 0.0;
    for (int k = 0; k < 3; ++k) x[cq[k]] = y[k];
}

// IglUtils::segTriIntersect (IglUtils.hpp:214-265)
DEV bool seg_tri_intersect(const double* ve0, const double* ve1, const double* vt0, const double* vt1, const double* vt2)
{
    const int o1 = orient3d(vt0, vt1, vt2, ve0), o2 = orient3d(vt0, vt1, vt2, ve1);
    if (o1 == 0 || o2 == 0) return false; // coplanar: caught through d(EE) = 0 or d(PT) = 0
    if (o1 == o2) return false;           // the edge is on one side of the triangle's plane
    double m[3][3], b[3], uvt[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        m[i][0] = vt1[i] - vt0[i];
        m[i][1] = vt2[i] - vt0[i];
        m[i][2] = ve0[i] - ve1[i];
        b[i] = ve0[i] - vt0[i];
    }
    full_piv_solve3(m, b, uvt);
    return uvt[0] >= 0.0 && uvt[1] >= 0.0 && uvt[0] + uvt[1] <= 1.0 && uvt[2] >= 0.0 && uvt[2] <= 1.0;
}

// ---- kernels ---------------------------------------------------------------------------------------------------------------
// one warp per query triangle (kPairQueriesPerWarp in a row): scan the edge grid for boxes that overlap the triangle's box, run the
// exact test on the lanes that hold a candidate, count the triangles with at least one intersecting edge (the reference stops at
// the first per triangle: same count)
__global__ void __launch_bounds__(256) k_tri_edge_intersect(SurfArgs s, const Grid* __restrict__ gp, SortedGrid eg, const Box* __restrict__ tboxes, int first, int last,
    int* __restrict__ n_hit)
{
    const int lane = threadIdx.x & 31;
    const Grid g = *gp;
    const int q0 = first + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kPairQueriesPerWarp;
    for (int f = q0; f < min(q0 + kPairQueriesPerWarp, last); ++f) {
        const int tv[3] = { s.SF[f], s.SF[(size_t)s.nSF + f], s.SF[(size_t)2 * s.nSF + f] };
        double t[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) t[k][c] = __ldg(s.V + (size_t)c * s.nV + tv[k]);
        const int cod_f = s.vCoDim ? s.vCoDim[tv[0]] : 3;
        const bool tri_dbc = s.dbc && s.dbc[tv[0]] && s.dbc[tv[1]] && s.dbc[tv[2]];
        bool any = false;
        warp_scan_candidates(g, eg, tboxes[f], lane, [&](bool hit, int e) {
            if (!hit || any) return;
            const int e0 = s.SE[2 * e], e1 = s.SE[2 * e + 1];
            if (e0 == tv[0] || e0 == tv[1] || e0 == tv[2] || e1 == tv[0] || e1 == tv[1] || e1 == tv[2]) return;
            const int cod_e = s.vCoDim ? s.vCoDim[e0] : 3;
            // mesh against itself: :3281-3284; mesh against the obstacle: every pair (MeshCO.cpp:2611-2678); inside the obstacle: none
            const bool oT = obstacle_vertex(s, tv[0]), oE = obstacle_vertex(s, e0);
            if (oT && oE) return;
            if (!oT && !oE && ((cod_f < 3 && cod_e < 3) || (tri_dbc && s.dbc[e0] && s.dbc[e1]))) return;
            double p0[3], p1[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                p0[c] = __ldg(s.V + (size_t)c * s.nV + e0);
                p1[c] = __ldg(s.V + (size_t)c * s.nV + e1);
            }
            if (seg_tri_intersect(p0, p1, t[0], t[1], t[2])) any = true;
        });
        if (__any_sync(0xffffffffu, any) && lane == 0) atomicAdd(n_hit, 1);
    }
}

// Mesh::checkInversion: tets with mu, lambda != 0 whose current edge matrix has a negative determinant
__global__ void __launch_bounds__(256) k_count_inverted(ElasticArgs p, int* __restrict__ n_inv)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool inv = false;
    if (t < p.t_end - p.t_begin) {
        const int tt = p.t_begin + t;
        if (__ldg(p.mu + tt) != 0.0 && __ldg(p.lam + tt) != 0.0) {
            int v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = __ldg(p.T + (size_t)k * p.nT + tt);
            double e[3][3];
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int c = 0; c < 3; ++c) e[k][c] = __ldg(p.V + (size_t)c * p.nV + v[k + 1]) - __ldg(p.V + (size_t)c * p.nV + v[0]);
            // determinant of the matrix whose COLUMNS are e[0], e[1], e[2] (cofactors along the first row)
            const double det = e[0][0] * (e[1][1] * e[2][2] - e[2][1] * e[1][2]) - e[1][0] * (e[0][1] * e[2][2] - e[2][1] * e[0][2])
                + e[2][0] * (e[0][1] * e[1][2] - e[1][1] * e[0][2]);
            inv = det < 0.0;
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, inv);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(n_inv, __popc(m));
}

} // namespace ipcgpu

using namespace ipcgpu;

SurfArgs surf_args(const ipcgpu_ctx* ctx); // constraint.cu
SortedGrid edge_grid(const ipcgpu_ctx* ctx);
int boxes_and_grid(ipcgpu_ctx* ctx, const double* dir, const double* alpha_ptr, double radius, const double* radius_ptr, bool with_vertex_boxes, const int* vmin,
    const int* vmax); // constraint.cu

static inline int nblk(long long n, int b) { return (int)((n + b - 1) / b); }

// enqueue the inversion count of this rank's tets into IterState::checks[0]
int safeguard_inversion(ipcgpu_ctx* ctx)
{
    cudaStream_t st = ctx->stream;
    int* cnt = &ctx->iter.p->checks[0];
    if (cudaMemsetAsync(cnt, 0, sizeof(int), st) != cudaSuccess) return IPCGPU_ERR_CUDA;
    const ElasticArgs p = ctx->eargs();
    const int n = p.t_end - p.t_begin;
    if (n > 0) k_count_inverted<<<nblk(n, 256), 256, 0, st>>>(p, cnt);
    ++ctx->launches;
    return cudaGetLastError() == cudaSuccess ? 0 : IPCGPU_ERR_CUDA;
}

// enqueue the edge-triangle intersection count (this rank's share of the triangles) into IterState::checks[1].  The static grid is
// rebuilt at the current positions with zero inflation (the reference rebuilds its hash before the check as well, Optimizer.cpp:2720).
int safeguard_intersections(ipcgpu_ctx* ctx)
{
    cudaStream_t st = ctx->stream;
    ContactWork& w = ctx->cw;
    const SurfArgs s = surf_args(ctx);
    int* cnt = &ctx->iter.p->checks[1];
    if (cudaMemsetAsync(cnt, 0, sizeof(int), st) != cudaSuccess) return IPCGPU_ERR_CUDA;
    if (s.nSF == 0 || s.nSE == 0) return 0;
    int rc = boxes_and_grid(ctx, nullptr, nullptr, 0.0, nullptr, false, nullptr, nullptr);
    if (rc) return rc;
    const SortedGrid eg = edge_grid(ctx);
    const int f0 = (int)((long long)s.nSF * ctx->rank / ctx->nranks), f1 = (int)((long long)s.nSF * (ctx->rank + 1) / ctx->nranks);
    if (f1 > f0) k_tri_edge_intersect<<<nblk(f1 - f0, 8 * kPairQueriesPerWarp), 256, 0, st>>>(s, w.grid.p, eg, w.tbox.p, f0, f1, cnt);
    ++ctx->launches;
    return cudaGetLastError() == cudaSuccess ? 0 : IPCGPU_ERR_CUDA;
}
