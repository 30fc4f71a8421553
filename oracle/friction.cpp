/*
 * friction.cpp -- CPU ORACLE (test infrastructure, NOT product code): lagged smoothed-static-friction terms of the self-contact pairs.
 *
 * Restates, statement for statement and Eigen-free:
 *   FrictionUtils.hpp:24-347              tangent bases, closest-point coordinates, relative displacement, lifts, T^T T, f0/f1/f2 (C1 clamping:
 *                                         SFCLAMPING_ORDER = 1, Types.hpp:42)
 *   SelfCollisionHandler.cpp:2481-2527    computeDistCoordAndTanBasis
 *   SelfCollisionHandler.cpp:2529-2596    computeFrictionEnergy
 *   SelfCollisionHandler.cpp:2598-2735    augmentFrictionGradient
 *   SelfCollisionHandler.cpp:2745-2987    augmentFrictionHessian (makePD on the 6/9/12-square block, serial CSR add)
 *   Optimizer.cpp:1582-1595               lagged normal force  lambda_c = -kappa * b'(d_c) * 2 sqrt(d_c) * multiplicity
 *
 * Pinning: none of this has expected outputs in the reference tree; the tests pin it by finite differences (E -> g -> unprojected H), by the
 * closed-form structure (H = T^T S T with a 2x2 S) and by invariants (translation null space, PSD).  Eigen's normalized()/ldlt() are restated.
 */
#include "oracle.h"
#include <algorithm>
#include <cmath>
#include <vector>

extern "C" void orc_makePD(int n, double* M);

namespace {

struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator*(double s, V3 a) { return { s * a.x, s * a.y, s * a.z }; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline V3 normalized(V3 a) /* Eigen: a / sqrt(squaredNorm) when the squared norm is positive */
{
    const double z = dot(a, a);
    return z > 0.0 ? (1.0 / std::sqrt(z)) * V3{ a.x, a.y, a.z } : a;
}
inline V3 vert(const double* V, int nV, int v) { return { V[v], V[(size_t)nV + v], V[(size_t)2 * nV + v] }; }

/* Eigen's 2x2 ldlt().solve (pivot on the larger diagonal), A = [[a,b],[b,c]] */
inline void ldlt2_solve(double a, double b, double c, double r0, double r1, double& x0, double& x1)
{
    if (std::fabs(c) > std::fabs(a)) {
        std::swap(a, c);
        std::swap(r0, r1);
        const double l = b / a, dd = c - l * b;
        const double y0 = r0, y1 = r1 - l * y0;
        const double z1 = (dd != 0.0) ? y1 / dd : 0.0, z0 = (a != 0.0) ? y0 / a : 0.0;
        x1 = z0 - l * z1;
        x0 = z1;
    }
    else {
        const double l = (a != 0.0) ? b / a : 0.0, dd = c - l * b;
        const double y0 = r0, y1 = r1 - l * y0;
        const double z1 = (dd != 0.0) ? y1 / dd : 0.0, z0 = (a != 0.0) ? y0 / a : 0.0;
        x1 = z1;
        x0 = z0 - l * x1;
    }
}

/* pair kind from the MMCVID encoding (SelfCollisionHandler.cpp:2491-2523): 1 EE, 3 PP, 2 PE, 0 PT; v = vertex ids */
inline int decode(const int mm[4], int v[4], int& nv)
{
    if (mm[0] >= 0) { v[0] = mm[0]; v[1] = mm[1]; v[2] = mm[2]; v[3] = mm[3]; nv = 4; return 1; }
    v[0] = -mm[0] - 1; v[1] = mm[1]; v[2] = mm[2]; v[3] = mm[3];
    if (mm[2] < 0) { nv = 2; return 3; }
    if (mm[3] < 0) { nv = 3; return 2; }
    nv = 4;
    return 0;
}

/* FrictionUtils.hpp:24-34, 95-105, 163-172, 227-244: basis as (col0, col1) */
inline void tangent_basis(int kind, const V3* x, V3& b0, V3& b1)
{
    if (kind == 0) { const V3 v12 = x[2] - x[1]; b0 = normalized(v12); b1 = normalized(cross(cross(v12, x[3] - x[1]), v12)); }
    else if (kind == 1) { const V3 v01 = x[1] - x[0]; b0 = normalized(v01); b1 = normalized(cross(cross(v01, x[3] - x[2]), v01)); }
    else if (kind == 2) { const V3 v12 = x[2] - x[1]; b0 = normalized(v12); b1 = normalized(cross(v12, x[0] - x[1])); }
    else {
        const V3 v01 = x[1] - x[0];
        const V3 xc = cross(V3{ 1.0, 0.0, 0.0 }, v01), yc = cross(V3{ 0.0, 1.0, 0.0 }, v01);
        if (dot(xc, xc) > dot(yc, yc)) { b0 = normalized(xc); b1 = normalized(cross(v01, xc)); }
        else { b0 = normalized(yc); b1 = normalized(cross(v01, yc)); }
    }
}
/* FrictionUtils.hpp:36-46, 107-129, 174-181 */
inline void closest_point(int kind, const V3* x, double& c0, double& c1)
{
    c0 = c1 = 0.0;
    if (kind == 0) {
        const V3 r0 = x[2] - x[1], r1 = x[3] - x[1], rel = x[0] - x[1];
        ldlt2_solve(dot(r0, r0), dot(r0, r1), dot(r1, r1), dot(r0, rel), dot(r1, rel), c0, c1);
    }
    else if (kind == 1) {
        const V3 e20 = x[0] - x[2], e01 = x[1] - x[0], e23 = x[3] - x[2];
        ldlt2_solve(dot(e01, e01), -dot(e23, e01), dot(e23, e23), -dot(e20, e01), dot(e20, e23), c0, c1);
    }
    else if (kind == 2) {
        const V3 e12 = x[2] - x[1];
        c0 = dot(x[0] - x[1], e12) / dot(e12, e12);
    }
}
/* stencil weights w_k with relDX = sum_k w_k dx_k  (FrictionUtils.hpp:48-57, 131-140, 183-191, 246-252); T = [w_0 B^T ... w_3 B^T] */
inline void weights(int kind, double c0, double c1, double w[4])
{
    if (kind == 0) { w[0] = 1.0; w[1] = -1.0 + c0 + c1; w[2] = -c0; w[3] = -c1; }
    else if (kind == 1) { w[0] = 1.0 - c0; w[1] = c0; w[2] = c1 - 1.0; w[3] = -c1; }
    else if (kind == 2) { w[0] = 1.0; w[1] = c0 - 1.0; w[2] = -c0; w[3] = 0.0; }
    else { w[0] = 1.0; w[1] = -1.0; w[2] = w[3] = 0.0; }
}
/* relDX3D exactly as the reference writes it (operation order matters to the last bit only) */
inline V3 rel_dx(int kind, const V3* dx, double c0, double c1)
{
    if (kind == 0) return dx[0] - (dx[1] + c0 * (dx[2] - dx[1]) + c1 * (dx[3] - dx[1]));
    if (kind == 1) return dx[0] + c0 * (dx[1] - dx[0]) - (dx[2] + c1 * (dx[3] - dx[2]));
    if (kind == 2) return dx[0] - (dx[1] + c0 * (dx[2] - dx[1]));
    return dx[0] - dx[1];
}
/* C1 clamping (FrictionUtils.hpp:278-292) */
inline double f0_SF(double x2, double eps) { return x2 * (-std::sqrt(x2) / 3.0 + eps) / (eps * eps) + eps / 3.0; }
inline double f1_SF_div(double x2, double eps) { return (-std::sqrt(x2) + 2.0 * eps) / (eps * eps); }
inline double f2_SF(double x2, double eps) { return 2.0 * (eps - std::sqrt(x2)) / (eps * eps); }

struct PairData {
    int kind, nv, v[4];
    double w[4];
    V3 b0, b1;
    double u0, u1; /* relDX in the tangent plane */
};
inline PairData pair_data(const double* V, const double* Vt, int nV, const int mm[4], const double* coord, const double* basis)
{
    PairData p;
    p.kind = decode(mm, p.v, p.nv);
    weights(p.kind, coord[0], coord[1], p.w);
    p.b0 = { basis[0], basis[1], basis[2] };
    p.b1 = { basis[3], basis[4], basis[5] };
    V3 dx[4];
    for (int k = 0; k < p.nv; ++k) dx[k] = vert(V, nV, p.v[k]) - vert(Vt, nV, p.v[k]);
    const V3 r = rel_dx(p.kind, dx, coord[0], coord[1]);
    p.u0 = dot(r, p.b0);
    p.u1 = dot(r, p.b1);
    return p;
}
/* liftRelDXTanToMesh_*: TTTDX = T^T u */
inline void lift(const PairData& p, double u0, double u1, double* out /* 3 nv */)
{
    const V3 t = u0 * p.b0 + u1 * p.b1;
    for (int k = 0; k < p.nv; ++k) {
        out[3 * k] = p.w[k] * t.x;
        out[3 * k + 1] = p.w[k] * t.y;
        out[3 * k + 2] = p.w[k] * t.z;
    }
}

struct CsrSink { /* LinSysSolver::addCoeff: upper triangle only, silently ignores entries outside the pattern's triangle */
    const int* ia;
    const int* ja;
    int base;
    double* a;
    void add(int r, int c, double v) const
    {
        if (r > c) return;
        const int lo = ia[r] - base, hi = ia[r + 1] - base;
        const int* q = std::lower_bound(ja + lo, ja + hi, c + base);
        if (q != ja + hi && *q == c + base) a[q - ja] += v;
    }
};

} // namespace

extern "C" {

/* Optimizer.cpp:1582-1595 + SelfCollisionHandler.cpp:2481-2527.  basis: 6 per pair, Eigen column-major 3x2 (col0 | col1). */
void orc_friction_lag(const orc_surf* s, const int* mmcvid, int nC, double dHat, double kappa, double* lambda, double* coord, double* basis)
{
    for (int c = 0; c < nC; ++c) {
        const int* mm = mmcvid + 4 * c;
        int v[4], nv;
        const int kind = decode(mm, v, nv);
        V3 x[4];
        double xs[12];
        for (int k = 0; k < nv; ++k) {
            x[k] = vert(s->V, s->nV, v[k]);
            xs[3 * k] = x[k].x; xs[3 * k + 1] = x[k].y; xs[3 * k + 2] = x[k].z;
        }
        double d;
        if (kind == 0) orc_d_PT(xs, &d);
        else if (kind == 1) orc_d_EE(xs, &d);
        else if (kind == 2) orc_d_PE(xs, &d);
        else orc_d_PP(xs, &d);
        double b, db, d2b;
        orc_barrier(d, dHat, &b, &db, &d2b);
        double lam = db;
        lam *= -kappa * 2.0 * std::sqrt(d);
        if (mm[3] < -1) lam *= -mm[3]; /* PP or PE duplication */
        lambda[c] = lam;
        closest_point(kind, x, coord[2 * c], coord[2 * c + 1]);
        V3 b0, b1;
        tangent_basis(kind, x, b0, b1);
        double* B = basis + 6 * (size_t)c;
        B[0] = b0.x; B[1] = b0.y; B[2] = b0.z; B[3] = b1.x; B[4] = b1.y; B[5] = b1.z;
    }
}

/* SelfCollisionHandler.cpp:2529-2596.  s->V = current positions, Vt = positions at the start of the time step (result.V_prev), both SoA. */
void orc_friction_energy(const orc_surf* s, const double* Vt, const int* mmcvid, int nC, const double* lambda, const double* coord, const double* basis,
    double eps2, double coef, double* E)
{
    const double eps = std::sqrt(eps2);
    double sum = 0.0;
    for (int c = 0; c < nC; ++c) {
        const PairData p = pair_data(s->V, Vt, s->nV, mmcvid + 4 * c, coord + 2 * c, basis + 6 * (size_t)c);
        const double x2 = p.u0 * p.u0 + p.u1 * p.u1;
        sum += (x2 > eps2) ? lambda[c] * std::sqrt(x2) : lambda[c] * f0_SF(x2, eps);
    }
    *E = sum * coef;
}

/* SelfCollisionHandler.cpp:2598-2735 ; g += (interleaved) */
void orc_friction_gradient(const orc_surf* s, const double* Vt, const int* mmcvid, int nC, const double* lambda, const double* coord, const double* basis,
    double eps2, double coef, double* g)
{
    const double eps = std::sqrt(eps2);
    for (int c = 0; c < nC; ++c) {
        const PairData p = pair_data(s->V, Vt, s->nV, mmcvid + 4 * c, coord + 2 * c, basis + 6 * (size_t)c);
        const double x2 = p.u0 * p.u0 + p.u1 * p.u1;
        double u0 = p.u0, u1 = p.u1;
        if (x2 > eps2) { const double n = std::sqrt(x2); u0 /= n; u1 /= n; }
        else { const double f = f1_SF_div(x2, eps); u0 *= f; u1 *= f; }
        double t[12];
        lift(p, u0, u1, t);
        const double w = coef * lambda[c];
        for (int k = 0; k < p.nv; ++k)
            for (int q = 0; q < 3; ++q) g[3 * (size_t)p.v[k] + q] += t[3 * k + q] * w;
    }
}

/* one pair's projected block (n = 3 nv, row-major n x n into H[144] with leading dimension 12), :2759-2960 */
void orc_friction_pair_hessian(const orc_surf* s, const double* Vt, const int mm[4], double lambda, const double coord[2], const double basis[6], double eps2,
    double coef, int project, double* H144, int* nvert)
{
    const double eps = std::sqrt(eps2);
    const PairData p = pair_data(s->V, Vt, s->nV, mm, coord, basis);
    const int n = 3 * p.nv;
    std::vector<double> Hn((size_t)n * n);
    /* TTT = T^T T */
    const double B[2][3] = { { p.b0.x, p.b0.y, p.b0.z }, { p.b1.x, p.b1.y, p.b1.z } };
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double v = 0.0;
            for (int r = 0; r < 2; ++r) v += (p.w[i / 3] * B[r][i % 3]) * (p.w[j / 3] * B[r][j % 3]);
            Hn[(size_t)i * n + j] = v;
        }
    const double x2 = p.u0 * p.u0 + p.u1 * p.u1, xn = std::sqrt(x2);
    double t[12];
    lift(p, p.u0, p.u1, t);
    bool proj = false;
    if (x2 > eps2) {
        const double a = coef * lambda / xn, b = coef * lambda / (x2 * xn);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) Hn[(size_t)i * n + j] = Hn[(size_t)i * n + j] * a - (t[i] * b) * t[j];
        proj = true;
    }
    else {
        const double f1 = f1_SF_div(x2, eps), f2 = f2_SF(x2, eps);
        const double a = coef * lambda * f1;
        for (int i = 0; i < n * n; ++i) Hn[i] *= a;
        if (f2 != f1 && x2) {
            const double b = coef * lambda * (f2 - f1) / x2;
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) Hn[(size_t)i * n + j] += (t[i] * b) * t[j];
            proj = true;
        }
    }
    if (proj && project) orc_makePD(n, Hn.data());
    for (int i = 0; i < 144; ++i) H144[i] = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) H144[i * 12 + j] = Hn[(size_t)i * n + j];
    *nvert = p.nv;
}

/* SelfCollisionHandler.cpp:2745-2987 ; a += */
void orc_friction_hessian_csr(const orc_surf* s, const double* Vt, const int* mmcvid, int nC, const double* lambda, const double* coord, const double* basis,
    double eps2, double coef, int projectDBC, const int* ia, const int* ja, int index_base, double* a, int nthreads)
{
    CsrSink sink{ ia, ja, index_base, a };
    std::vector<double> blocks((size_t)144 * std::max(nC, 1));
    std::vector<int> nvs(std::max(nC, 1));
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 64)
    for (int c = 0; c < nC; ++c)
        orc_friction_pair_hessian(s, Vt, mmcvid + 4 * c, lambda[c], coord + 2 * c, basis + 6 * (size_t)c, eps2, coef, 1, &blocks[(size_t)144 * c], &nvs[c]);
    for (int c = 0; c < nC; ++c) {
        int v[4], nv;
        decode(mmcvid + 4 * c, v, nv);
        const double* H = &blocks[(size_t)144 * c];
        for (int i = 0; i < nv; ++i) {
            if (s->dbc && (s->dbc[v[i]] == 1 || (s->dbc[v[i]] == 2 && projectDBC))) continue;
            for (int j = 0; j < nv; ++j) {
                if (s->dbc && (s->dbc[v[j]] == 1 || (s->dbc[v[j]] == 2 && projectDBC))) continue;
                for (int r = 0; r < 3; ++r)
                    for (int q = 0; q < 3; ++q) sink.add(3 * v[i] + r, 3 * v[j] + q, H[(3 * i + r) * 12 + 3 * j + q]);
            }
        }
    }
}

} // extern "C"
