// contact.cuh -- device pair math for the barrier-contact stage: squared distances, closest-feature
// classification, gradients/Hessians, C2 barrier, EE mollifier.
//
// Functions being replaced (ipc-sim/IPC): src/CollisionObject/MeshCollisionUtils.hpp d_PP/d_PE/d_PT/d_EE
// (:156, :227, :685, :1287), g_*/H_* (MATLAB codegen, :163-2002), dType_PT/EE (:2073-2210),
// computePointTriD/EdgeEdgeD (:2279-2383), EE cross-norm + mollifier (:2409-2912), compute_eps_x (:2969),
// src/Utils/BarrierFunctions.hpp:56-83.
//
// The derivatives are NOT the codegen: they are evaluated in "difference space".  Every distance depends on
// the vertices only through y = (m, e1, e2) (three difference vectors), where
//     d = N/L,   N = s^2 (s = m.(e1 x e2)) or |w x u|^2,   L = |e1 x e2|^2 or |u|^2,
//     grad d = (grad N - d grad L)/L,   hess d = (hess N - grad d grad L^T - grad L grad d^T - d hess L)/L,
// and the constant +-I Jacobian maps (9-vector, 9x9) to the 3/4-vertex stencil.
#pragma once
#include "common.cuh"

namespace ipcgpu {

struct V3 {
    double x, y, z;
};
DEV V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
DEV V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
DEV V3 operator*(double s, V3 a) { return { s * a.x, s * a.y, s * a.z }; }
DEV double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DEV V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
DEV double norm2(V3 a) { return dot(a, a); }
DEV double comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
DEV V3 load_vertex(const double* __restrict__ V, int nV, int v) { return { __ldg(V + v), __ldg(V + (size_t)nV + v), __ldg(V + (size_t)2 * nV + v) }; }

// ---- squared distances --------------------------------------------------------------------------------
DEV double d_PP(V3 a, V3 b) { return norm2(a - b); }
DEV double d_PE(V3 p, V3 e0, V3 e1) { return norm2(cross(e0 - p, e1 - p)) / norm2(e1 - e0); }
DEV double d_PT(V3 p, V3 t0, V3 t1, V3 t2)
{
    V3 b = cross(t1 - t0, t2 - t0);
    double aTb = dot(p - t0, b);
    return aTb * aTb / norm2(b);
}
DEV double d_EE(V3 a0, V3 a1, V3 b0, V3 b1)
{
    V3 b = cross(a1 - a0, b1 - b0);
    double aTb = dot(b0 - a0, b);
    return aTb * aTb / norm2(b);
}

// ---- closest-feature classification -------------------------------------------------------------------
// 2x2 LDL^T solve with diagonal pivoting (what Eigen's ldlt() does for the reference, :2174)
DEV void ldlt2(double a, double b, double c, double r0, double r1, double& x0, double& x1)
{
    const bool sw = fabs(c) > fabs(a);
    if (sw) {
        double t = a; a = c; c = t;
        t = r0; r0 = r1; r1 = t;
    }
    const double l = (a != 0.0) ? b / a : 0.0, dd = c - l * b;
    const double y1 = r1 - l * r0;
    const double z1 = (dd != 0.0) ? y1 / dd : 0.0, z0 = (a != 0.0) ? r0 / a : 0.0;
    const double s1 = z1, s0 = z0 - l * s1;
    x0 = sw ? s1 : s0;
    x1 = sw ? s0 : s1;
}
DEV void edge_param(V3 e, V3 n, V3 rel, double& p0, double& p1)
{
    V3 b1 = cross(e, n);
    ldlt2(dot(e, e), dot(e, b1), dot(b1, b1), dot(e, rel), dot(b1, rel), p0, p1);
}
// 0,1,2: PP with t0,t1,t2 ; 3,4,5: PE with (t0,t1),(t1,t2),(t2,t0) ; 6: PT      [:2160-2210]
DEV int dType_PT(V3 p, V3 t0, V3 t1, V3 t2)
{
    V3 n = cross(t1 - t0, t2 - t0);
    double p00, p10, p01, p11, p02, p12;
    edge_param(t1 - t0, n, p - t0, p00, p10);
    if (p00 > 0.0 && p00 < 1.0 && p10 >= 0.0) return 3;
    edge_param(t2 - t1, n, p - t1, p01, p11);
    if (p01 > 0.0 && p01 < 1.0 && p11 >= 0.0) return 4;
    edge_param(t0 - t2, n, p - t2, p02, p12);
    if (p02 > 0.0 && p02 < 1.0 && p12 >= 0.0) return 5;
    if (p00 <= 0.0 && p02 >= 1.0) return 0;
    if (p01 <= 0.0 && p00 >= 1.0) return 1;
    if (p02 <= 0.0 && p01 >= 1.0) return 2;
    return 6;
}
// [:2073-2158]
DEV int dType_EE(V3 v0, V3 v1, V3 v2, V3 v3)
{
    V3 u = v1 - v0, v = v3 - v2, w = v0 - v2;
    const double a = norm2(u), b = dot(u, v), c = norm2(v), d = dot(u, w), e = dot(v, w);
    const double D = a * c - b * b;
    double tD = D, tN;
    int def = 8;
    const double sN = (b * e - c * d);
    if (sN <= 0.0) { tN = e; tD = c; def = 2; }
    else if (sN >= D) { tN = e + b; tD = c; def = 5; }
    else {
        tN = (a * e - b * d);
        V3 uxv = cross(u, v);
        if (tN > 0.0 && tN < tD && (dot(uxv, w) == 0.0 || norm2(uxv) < 1.0e-20 * a * c)) {
            if (sN < D / 2) { tN = e; tD = c; def = 2; }
            else { tN = e + b; tD = c; def = 5; }
        }
    }
    if (tN <= 0.0) {
        if (-d <= 0.0) return 0;
        else if (-d >= a) return 3;
        else return 6;
    }
    else if (tN >= tD) {
        if ((-d + b) <= 0.0) return 1;
        else if ((-d + b) >= a) return 4;
        else return 7;
    }
    return def;
}
// unsigned squared distances used by the CCD drivers  [:2279-2383]
DEV double point_tri_d(V3 p, V3 t0, V3 t1, V3 t2)
{
    switch (dType_PT(p, t0, t1, t2)) {
    case 0: return d_PP(p, t0);
    case 1: return d_PP(p, t1);
    case 2: return d_PP(p, t2);
    case 3: return d_PE(p, t0, t1);
    case 4: return d_PE(p, t1, t2);
    case 5: return d_PE(p, t2, t0);
    default: return d_PT(p, t0, t1, t2);
    }
}
DEV double edge_edge_d(V3 a0, V3 a1, V3 b0, V3 b1)
{
    switch (dType_EE(a0, a1, b0, b1)) {
    case 0: return d_PP(a0, b0);
    case 1: return d_PP(a0, b1);
    case 2: return d_PE(a0, b0, b1);
    case 3: return d_PP(a1, b0);
    case 4: return d_PP(a1, b1);
    case 5: return d_PE(a1, b0, b1);
    case 6: return d_PE(b0, a0, a1);
    case 7: return d_PE(b1, a0, a1);
    default: return d_EE(a0, a1, b0, b1);
    }
}

// ---- C2 clamped log barrier on squared distance [BarrierFunctions.hpp:56-83] -----------------------------
DEV void barrier_all(double d, double dHat, double& b, double& db, double& d2b)
{
    const double t2 = d - dHat;
    const double lg = log(d / dHat);
    b = -(d - dHat) * (d - dHat) * lg;
    db = t2 * lg * -2.0 - (t2 * t2) / d;
    d2b = (lg * -2.0 - t2 * 4.0 / d) + 1.0 / (d * d) * (t2 * t2);
}

// ---- difference-space derivatives ------------------------------------------------------------------------
struct M33 {
    double a[9];
};
DEV M33 m_zero() { M33 m; for (int i = 0; i < 9; ++i) m.a[i] = 0.0; return m; }
DEV M33 m_ident(double s) { M33 m = m_zero(); m.a[0] = m.a[4] = m.a[8] = s; return m; }
DEV M33 m_outer(V3 u, V3 v)
{
    M33 m;
    m.a[0] = u.x * v.x; m.a[1] = u.x * v.y; m.a[2] = u.x * v.z;
    m.a[3] = u.y * v.x; m.a[4] = u.y * v.y; m.a[5] = u.y * v.z;
    m.a[6] = u.z * v.x; m.a[7] = u.z * v.y; m.a[8] = u.z * v.z;
    return m;
}
DEV M33 m_skew(V3 v)
{
    M33 m = m_zero();
    m.a[1] = -v.z; m.a[2] = v.y; m.a[3] = v.z; m.a[5] = -v.x; m.a[6] = -v.y; m.a[7] = v.x;
    return m;
}
DEV M33 m_axpy(double s, const M33& x, const M33& y) { M33 m; for (int i = 0; i < 9; ++i) m.a[i] = s * x.a[i] + y.a[i]; return m; }
DEV M33 m_scale(double s, const M33& x) { M33 m; for (int i = 0; i < 9; ++i) m.a[i] = s * x.a[i]; return m; }
DEV M33 m_T(const M33& x) { M33 m; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m.a[3 * i + j] = x.a[3 * j + i]; return m; }

// value + gradient (3 blocks) + Hessian (3x3 blocks, only k<=l stored; H[l][k] = H[k][l]^T)
struct Diff {
    double val;
    V3 g[3];
    M33 H[6]; // (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
};
__device__ __forceinline__ int hidx(int k, int l) { return k == 0 ? l : (k == 1 ? 2 + l : 5); } // k<=l

// c = |e1 x e2|^2 on slots (1,2)
__device__ inline void cross_norm(V3 e1, V3 e2, Diff& r)
{
    V3 n = cross(e1, e2);
    r.val = norm2(n);
    r.g[0] = { 0, 0, 0 };
    r.g[1] = 2.0 * cross(e2, n);
    r.g[2] = 2.0 * cross(n, e1);
    r.H[0] = r.H[1] = r.H[2] = m_zero();
    r.H[3] = m_scale(2.0, m_axpy(-1.0, m_outer(e2, e2), m_ident(norm2(e2))));
    r.H[5] = m_scale(2.0, m_axpy(-1.0, m_outer(e1, e1), m_ident(norm2(e1))));
    r.H[4] = m_axpy(4.0, m_outer(e1, e2), m_axpy(-2.0, m_outer(e2, e1), m_ident(-2.0 * dot(e1, e2))));
}

// d = (m.(e1 x e2))^2 / |e1 x e2|^2
__device__ inline void plane_dist(V3 m, V3 e1, V3 e2, Diff& r)
{
    Diff L;
    cross_norm(e1, e2, L);
    V3 n = cross(e1, e2);
    const double s = dot(m, n);
    V3 gs[3] = { n, cross(e2, m), cross(m, e1) };
    r.val = s * s / L.val;
    const double invL = 1.0 / L.val;
#pragma unroll
    for (int i = 0; i < 3; ++i) r.g[i] = invL * ((2.0 * s) * gs[i] - r.val * L.g[i]);
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int l = k; l < 3; ++l) {
            M33 Hs = m_zero();
            if (k == 0 && l == 1) Hs = m_scale(-1.0, m_skew(e2));
            if (k == 0 && l == 2) Hs = m_skew(e1);
            if (k == 1 && l == 2) Hs = m_scale(-1.0, m_skew(m));
            M33 t = m_axpy(2.0, m_outer(gs[k], gs[l]), m_scale(2.0 * s, Hs));
            t = m_axpy(-1.0, m_outer(r.g[k], L.g[l]), t);
            t = m_axpy(-1.0, m_outer(L.g[k], r.g[l]), t);
            t = m_axpy(-r.val, L.H[hidx(k, l)], t);
            r.H[hidx(k, l)] = m_scale(invL, t);
        }
}

// d = |w x u|^2 / |u|^2 on slots (1,2)
__device__ inline void line_dist(V3 w, V3 u, Diff& r)
{
    Diff N;
    cross_norm(w, u, N);
    const double L = norm2(u), invL = 1.0 / L;
    V3 gL[3] = { { 0, 0, 0 }, { 0, 0, 0 }, 2.0 * u };
    r.val = N.val * invL;
#pragma unroll
    for (int i = 0; i < 3; ++i) r.g[i] = invL * (N.g[i] - r.val * gL[i]);
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int l = k; l < 3; ++l) {
            M33 t = N.H[hidx(k, l)];
            t = m_axpy(-1.0, m_outer(r.g[k], gL[l]), t);
            t = m_axpy(-1.0, m_outer(gL[k], r.g[l]), t);
            if (k == 2 && l == 2) t = m_axpy(-r.val, m_ident(2.0), t);
            r.H[hidx(k, l)] = m_scale(invL, t);
        }
}

// Jacobians y_k = sum_v J[k][v] x_v, packed 2 bits per entry (0 -> 0, 1 -> +1, 2 -> -1), index 4*k+v
//   PT (p,t0,t1,t2): m=p-t0, e1=t1-t0, e2=t2-t0      EE (a0,a1,b0,b1): m=b0-a0, e1=a1-a0, e2=b1-b0
//   PE (p,e0,e1):   w=p-e0 (slot 1), u=e1-e0 (slot 2) CR (a0,a1,b0,b1): e1=a1-a0 (slot 1), e2=b1-b0 (slot 2)
__device__ __forceinline__ int jac(int kind, int k, int v)
{
    // kind: 0 PT, 1 EE, 2 PE, 3 CR
    const signed char J[4][3][4] = {
        { { 1, -1, 0, 0 }, { 0, -1, 1, 0 }, { 0, -1, 0, 1 } },
        { { -1, 0, 1, 0 }, { -1, 1, 0, 0 }, { 0, 0, -1, 1 } },
        { { 0, 0, 0, 0 }, { 1, -1, 0, 0 }, { 0, -1, 1, 0 } },
        { { 0, 0, 0, 0 }, { -1, 1, 0, 0 }, { 0, 0, -1, 1 } } };
    return J[kind][k][v];
}

// scatter a Diff into vertex space. g: 3*nv entries, H: (3nv)x(3nv) accessed through the functor put(i,j,val)
template <typename PutG, typename PutH>
__device__ inline void diff_to_vertices(const Diff& D, int kind, int nv, bool want_g, bool want_h, PutG putg, PutH puth)
{
    if (want_g) {
        for (int v = 0; v < nv; ++v) {
            V3 acc = { 0, 0, 0 };
            for (int k = 0; k < 3; ++k) {
                int j = jac(kind, k, v);
                if (j) acc = acc + (double)j * D.g[k];
            }
            putg(3 * v, acc.x);
            putg(3 * v + 1, acc.y);
            putg(3 * v + 2, acc.z);
        }
    }
    if (want_h) {
        for (int a = 0; a < nv; ++a)
            for (int b = 0; b < nv; ++b) {
                M33 acc = m_zero();
                for (int k = 0; k < 3; ++k)
                    for (int l = 0; l < 3; ++l) {
                        int j = jac(kind, k, a) * jac(kind, l, b);
                        if (!j) continue;
                        if (k <= l) acc = m_axpy((double)j, D.H[hidx(k, l)], acc);
                        else acc = m_axpy((double)j, m_T(D.H[hidx(l, k)]), acc);
                    }
                for (int i = 0; i < 3; ++i)
                    for (int jx = 0; jx < 3; ++jx) puth(3 * a + i, 3 * b + jx, acc.a[3 * i + jx]);
            }
    }
}

// eps_x of the EE mollifier from rest positions [:2969-2974]
DEV double eps_x_rest(const double* __restrict__ Vr, int nV, int a0, int a1, int b0, int b1)
{
    return 1.0e-3 * norm2(load_vertex(Vr, nV, a0) - load_vertex(Vr, nV, a1)) * norm2(load_vertex(Vr, nV, b0) - load_vertex(Vr, nV, b1));
}

} // namespace ipcgpu
