/*
 * contact.cpp -- CPU ORACLE (test infrastructure only; see oracle.h header).
 * Restates the barrier-contact pair math and constraint-set semantics of ipc-sim/IPC:
 *   src/CollisionObject/MeshCollisionUtils.hpp  (d_*, g_*, H_*, dType_*, mollifier)
 *   src/Utils/BarrierFunctions.hpp:56-83
 *   src/CollisionObject/SelfCollisionHandler.cpp:38-148, 418-561, 2149-2478, 2990-3201
 *   src/TimeStepper/Optimizer.cpp:3290-3353
 *
 * The reference's gradients/Hessians are MATLAB-codegen bodies (hundreds of temporaries); they are NOT
 * reproduced here.  The same functions are written from their definitions in "difference space":
 * every squared distance depends on the vertices only through difference vectors y = (m, e1, e2), so
 *   d = N/L,  grad d = (grad N - d grad L)/L,  hess d = (hess N - grad d grad L^T - grad L grad d^T - d hess L)/L
 * with N, L built from the triple product s = m.(e1 x e2) and the cross norm |e1 x e2|^2, then mapped to the
 * vertices by the constant +-I Jacobian.  tests/test_oracle_contact.py pins these against the reference's own
 * codegen (oracle/_ref/libref_pairs.so, compiled from /root/reference) and against committed goldens.
 */
#include "oracle.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <array>
#include <vector>

namespace {

struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator*(double s, V3 a) { return { s * a.x, s * a.y, s * a.z }; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline double norm2(V3 a) { return dot(a, a); }
inline V3 ld(const double* p) { return { p[0], p[1], p[2] }; }
inline void st(double* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

/* 3x3 helpers, row-major */
struct M33 {
    double a[9];
};
inline M33 zero33() { M33 m; for (double& v : m.a) v = 0.0; return m; }
inline M33 ident33(double s) { M33 m = zero33(); m.a[0] = m.a[4] = m.a[8] = s; return m; }
inline M33 outer(V3 u, V3 v)
{
    M33 m;
    const double U[3] = { u.x, u.y, u.z }, W[3] = { v.x, v.y, v.z };
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m.a[3 * i + j] = U[i] * W[j];
    return m;
}
inline M33 skew(V3 v) /* [v]x w = v x w */
{
    M33 m = zero33();
    m.a[1] = -v.z; m.a[2] = v.y;
    m.a[3] = v.z; m.a[5] = -v.x;
    m.a[6] = -v.y; m.a[7] = v.x;
    return m;
}
inline M33 add(M33 a, M33 b) { M33 m; for (int i = 0; i < 9; ++i) m.a[i] = a.a[i] + b.a[i]; return m; }
inline M33 scale(double s, M33 a) { M33 m; for (int i = 0; i < 9; ++i) m.a[i] = s * a.a[i]; return m; }
inline M33 transpose(M33 a) { M33 m; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m.a[3 * i + j] = a.a[3 * j + i]; return m; }

/* difference-space quantities: gradient as 3 blocks, Hessian as 3x3 blocks (full, symmetric) */
struct Diff {
    double val;
    V3 g[3];
    M33 H[3][3];
};

/* cross norm c = |e1 x e2|^2 as a function of y=(m,e1,e2) (m unused) */
Diff cross_norm(V3 e1, V3 e2)
{
    Diff r;
    V3 n = cross(e1, e2);
    r.val = norm2(n);
    r.g[0] = { 0, 0, 0 };
    r.g[1] = 2.0 * cross(e2, n);
    r.g[2] = 2.0 * cross(n, e1);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.H[i][j] = zero33();
    r.H[1][1] = scale(2.0, add(ident33(norm2(e2)), scale(-1.0, outer(e2, e2))));
    r.H[2][2] = scale(2.0, add(ident33(norm2(e1)), scale(-1.0, outer(e1, e1))));
    /* d/de2 of (2|e2|^2 e1 - 2(e1.e2) e2) = 4 e1 e2^T - 2 e2 e1^T - 2 (e1.e2) I */
    r.H[1][2] = add(add(scale(4.0, outer(e1, e2)), scale(-2.0, outer(e2, e1))), ident33(-2.0 * dot(e1, e2)));
    r.H[2][1] = transpose(r.H[1][2]);
    return r;
}

/* d = (m.(e1 x e2))^2 / |e1 x e2|^2 */
Diff plane_dist(V3 m, V3 e1, V3 e2)
{
    V3 n = cross(e1, e2);
    double s = dot(m, n);
    V3 gs[3] = { n, cross(e2, m), cross(m, e1) };
    M33 Hs[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Hs[i][j] = zero33();
    Hs[0][1] = scale(-1.0, skew(e2)); Hs[1][0] = skew(e2);
    Hs[0][2] = skew(e1); Hs[2][0] = scale(-1.0, skew(e1));
    Hs[1][2] = scale(-1.0, skew(m)); Hs[2][1] = skew(m);
    Diff L = cross_norm(e1, e2);
    Diff r;
    double N = s * s;
    r.val = N / L.val;
    V3 gN[3];
    for (int i = 0; i < 3; ++i) gN[i] = (2.0 * s) * gs[i];
    for (int i = 0; i < 3; ++i) r.g[i] = (1.0 / L.val) * (gN[i] - r.val * L.g[i]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            M33 HN = add(scale(2.0, outer(gs[i], gs[j])), scale(2.0 * s, Hs[i][j]));
            M33 t = add(HN, scale(-1.0, add(add(outer(r.g[i], L.g[j]), outer(L.g[i], r.g[j])), scale(r.val, L.H[i][j]))));
            r.H[i][j] = scale(1.0 / L.val, t);
        }
    return r;
}

/* d = |w x u|^2 / |u|^2 as a function of y=(unused, w, u) */
Diff line_dist(V3 w, V3 u)
{
    Diff Nn = cross_norm(w, u);
    double L = norm2(u);
    V3 gL[3] = { { 0, 0, 0 }, { 0, 0, 0 }, 2.0 * u };
    Diff r;
    r.val = Nn.val / L;
    for (int i = 0; i < 3; ++i) r.g[i] = (1.0 / L) * (Nn.g[i] - r.val * gL[i]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            M33 HL = (i == 2 && j == 2) ? ident33(2.0) : zero33();
            M33 t = add(Nn.H[i][j], scale(-1.0, add(add(outer(r.g[i], gL[j]), outer(gL[i], r.g[j])), scale(r.val, HL))));
            r.H[i][j] = scale(1.0 / L, t);
        }
    return r;
}

/* map difference space -> vertex space. J[k][v] in {-1,0,1}: y_k = sum_v J[k][v] x_v */
template <int NV>
void to_vertices(const Diff& D, const int J[3][4], double* g, double* H)
{
    const int n = 3 * NV;
    if (g) {
        for (int v = 0; v < NV; ++v) {
            V3 acc = { 0, 0, 0 };
            for (int k = 0; k < 3; ++k)
                if (J[k][v]) acc = acc + (double)J[k][v] * D.g[k];
            st(g + 3 * v, acc);
        }
    }
    if (H) {
        for (int a = 0; a < NV; ++a)
            for (int b = 0; b < NV; ++b) {
                M33 acc = zero33();
                for (int k = 0; k < 3; ++k)
                    for (int l = 0; l < 3; ++l)
                        if (J[k][a] && J[l][b]) acc = add(acc, scale((double)(J[k][a] * J[l][b]), D.H[k][l]));
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) H[(3 * a + i) * n + 3 * b + j] = acc.a[3 * i + j];
            }
    }
}

/* vertex orderings: PT (p,t0,t1,t2): m=p-t0, e1=t1-t0, e2=t2-t0 ; EE (a0,a1,b0,b1): m=b0-a0, e1=a1-a0, e2=b1-b0 ;
 * PE (p,e0,e1): w=p-e0 (slot 1), u=e1-e0 (slot 2) */
const int J_PT[3][4] = { { 1, -1, 0, 0 }, { 0, -1, 1, 0 }, { 0, -1, 0, 1 } };
const int J_EE[3][4] = { { -1, 0, 1, 0 }, { -1, 1, 0, 0 }, { 0, 0, -1, 1 } };
const int J_PE[3][4] = { { 0, 0, 0, 0 }, { 1, -1, 0, 0 }, { 0, -1, 1, 0 } };
const int J_CR[3][4] = { { 0, 0, 0, 0 }, { -1, 1, 0, 0 }, { 0, 0, -1, 1 } }; /* EE cross norm: e1=a1-a0, e2=b1-b0 */

/* Eigen's 2x2 ldlt().solve with its diagonal pivoting (MeshCollisionUtils.hpp:2174): A = [[a,b],[b,c]] */
inline void ldlt2_solve(double a, double b, double c, double r0, double r1, double& x0, double& x1)
{
    if (std::fabs(c) > std::fabs(a)) { /* pivot on the larger diagonal */
        std::swap(a, c);
        std::swap(r0, r1);
        double l = b / a, dd = c - l * b;
        double y0 = r0, y1 = r1 - l * y0;
        double z1 = (dd != 0.0) ? y1 / dd : 0.0, z0 = (a != 0.0) ? y0 / a : 0.0;
        double s1 = z1, s0 = z0 - l * s1;
        x1 = s0;
        x0 = s1;
    }
    else {
        double l = (a != 0.0) ? b / a : 0.0, dd = c - l * b;
        double y0 = r0, y1 = r1 - l * y0;
        double z1 = (dd != 0.0) ? y1 / dd : 0.0, z0 = (a != 0.0) ? y0 / a : 0.0;
        x1 = z1;
        x0 = z0 - l * x1;
    }
}
/* param = (B B^T)^-1 B (p - o) with B = [e; e x n] (MeshCollisionUtils.hpp:2172-2175) */
inline void edge_param(V3 e, V3 nVec, V3 rel, double& p0, double& p1)
{
    V3 b1 = cross(e, nVec);
    ldlt2_solve(dot(e, e), dot(e, b1), dot(b1, b1), dot(e, rel), dot(b1, rel), p0, p1);
}

inline V3 vert(const orc_surf* s, int v) { return { s->V[v], s->V[(size_t)s->nV + v], s->V[(size_t)2 * s->nV + v] }; }
inline V3 vert_rest(const orc_surf* s, int v) { return { s->Vrest[v], s->Vrest[(size_t)s->nV + v], s->Vrest[(size_t)2 * s->nV + v] }; }
inline bool is_dbc(const orc_surf* s, int v) { return s->dbc && s->dbc[v] != 0; } /* Mesh::isDBCVertex */
inline bool is_proj_dbc(const orc_surf* s, int v, int projectDBC) { return s->dbc && (s->dbc[v] == 1 || (s->dbc[v] == 2 && projectDBC)); }
inline int codim(const orc_surf* s, int v) { return s->vCoDim ? s->vCoDim[v] : 3; }

struct CsrSink {
    const int* ia;
    const int* ja;
    int base;
    double* a;
    void add(int r, int c, double v) const
    {
        if (r > c) return;
        int lo = ia[r] - base, hi = ia[r + 1] - base;
        const int* p = std::lower_bound(ja + lo, ja + hi, c + base);
        if (p != ja + hi && *p == c + base) a[p - ja] += v;
    }
};

/* stencil evaluation for an active-set entry: vertex ids, n verts, d, g, H (n = 3*nv) */
int pair_verts(const int mm[4], int v[4])
{
    if (mm[0] >= 0) { v[0] = mm[0]; v[1] = mm[1]; v[2] = mm[2]; v[3] = mm[3]; return 4; }
    v[0] = -mm[0] - 1; v[1] = mm[1];
    if (mm[2] < 0) return 2;
    v[2] = mm[2];
    if (mm[3] < 0) return 3;
    v[3] = mm[3];
    return 4;
}

void pair_dgH(const orc_surf* s, const int mm[4], int v[4], int& nv, double& d, double* g, double* H)
{
    nv = pair_verts(mm, v);
    double x[12];
    for (int k = 0; k < nv; ++k) st(x + 3 * k, vert(s, v[k]));
    if (mm[0] >= 0) {
        orc_d_EE(x, &d); if (g) orc_g_EE(x, g); if (H) orc_H_EE(x, H);
    }
    else if (nv == 2) {
        orc_d_PP(x, &d); if (g) orc_g_PP(x, g); if (H) orc_H_PP(x, H);
    }
    else if (nv == 3) {
        orc_d_PE(x, &d); if (g) orc_g_PE(x, g); if (H) orc_H_PE(x, H);
    }
    else {
        orc_d_PT(x, &d); if (g) orc_g_PT(x, g); if (H) orc_H_PT(x, H);
    }
}

inline double eps_x_of(const orc_surf* s, int a0, int a1, int b0, int b1) /* MeshCollisionUtils.hpp:2969-2974 */
{
    return 1.0e-3 * norm2(vert_rest(s, a0) - vert_rest(s, a1)) * norm2(vert_rest(s, b0) - vert_rest(s, b1));
}

/* para-EE entry -> 4-vertex stencil of the two edges + mollifier */
void para_stencil(const orc_surf* s, const int mm[4], const int eIeJ[2], int ev[4])
{
    if (mm[3] >= 0) { ev[0] = mm[0]; ev[1] = mm[1]; ev[2] = mm[2]; ev[3] = mm[3]; }
    else {
        ev[0] = s->SE[2 * eIeJ[0]]; ev[1] = s->SE[2 * eIeJ[0] + 1];
        ev[2] = s->SE[2 * eIeJ[1]]; ev[3] = s->SE[2 * eIeJ[1] + 1];
    }
}

} // namespace

extern "C" {

void orc_makePD(int n, double* M); /* elastic.cpp */

/* MeshCollisionUtils.hpp:156-161, 227-233, 685-694, 1287-1296 */
void orc_d_PP(const double* v, double* d) { *d = norm2(ld(v) - ld(v + 3)); }
void orc_d_PE(const double* v, double* d)
{
    V3 v0 = ld(v), v1 = ld(v + 3), v2 = ld(v + 6);
    *d = norm2(cross(v1 - v0, v2 - v0)) / norm2(v2 - v1);
}
void orc_d_PT(const double* v, double* d)
{
    V3 v0 = ld(v), v1 = ld(v + 3), v2 = ld(v + 6), v3 = ld(v + 9);
    V3 b = cross(v2 - v1, v3 - v1);
    double aTb = dot(v0 - v1, b);
    *d = aTb * aTb / norm2(b);
}
void orc_d_EE(const double* v, double* d)
{
    V3 v0 = ld(v), v1 = ld(v + 3), v2 = ld(v + 6), v3 = ld(v + 9);
    V3 b = cross(v1 - v0, v3 - v2);
    double aTb = dot(v2 - v0, b);
    *d = aTb * aTb / norm2(b);
}
/* :163-176 */
void orc_g_PP(const double* v, double* g)
{
    V3 r = 2.0 * (ld(v) - ld(v + 3));
    st(g, r);
    st(g + 3, -1.0 * r);
}
void orc_H_PP(const double*, double* H)
{
    for (int i = 0; i < 36; ++i) H[i] = 0.0;
    for (int i = 0; i < 3; ++i) {
        H[i * 6 + i] = H[(i + 3) * 6 + i + 3] = 2.0;
        H[i * 6 + i + 3] = H[(i + 3) * 6 + i] = -2.0;
    }
}
/* functions of :235-287 / :300-620 (PE), :696-757 / :772-1217 (PT), :1298-1377 / :1392-2002 (EE) */
void orc_g_PE(const double* v, double* g) { Diff D = line_dist(ld(v) - ld(v + 3), ld(v + 6) - ld(v + 3)); to_vertices<3>(D, J_PE, g, nullptr); }
void orc_H_PE(const double* v, double* H) { Diff D = line_dist(ld(v) - ld(v + 3), ld(v + 6) - ld(v + 3)); to_vertices<3>(D, J_PE, nullptr, H); }
void orc_g_PT(const double* v, double* g) { Diff D = plane_dist(ld(v) - ld(v + 3), ld(v + 6) - ld(v + 3), ld(v + 9) - ld(v + 3)); to_vertices<4>(D, J_PT, g, nullptr); }
void orc_H_PT(const double* v, double* H) { Diff D = plane_dist(ld(v) - ld(v + 3), ld(v + 6) - ld(v + 3), ld(v + 9) - ld(v + 3)); to_vertices<4>(D, J_PT, nullptr, H); }
void orc_g_EE(const double* v, double* g) { Diff D = plane_dist(ld(v + 6) - ld(v), ld(v + 3) - ld(v), ld(v + 9) - ld(v + 6)); to_vertices<4>(D, J_EE, g, nullptr); }
void orc_H_EE(const double* v, double* H) { Diff D = plane_dist(ld(v + 6) - ld(v), ld(v + 3) - ld(v), ld(v + 9) - ld(v + 6)); to_vertices<4>(D, J_EE, nullptr, H); }

/* MeshCollisionUtils.hpp:2160-2210 */
int orc_dType_PT(const double* v)
{
    V3 v0 = ld(v), v1 = ld(v + 3), v2 = ld(v + 6), v3 = ld(v + 9);
    V3 nVec = cross(v2 - v1, v3 - v1);
    double p00, p10, p01, p11, p02, p12;
    edge_param(v2 - v1, nVec, v0 - v1, p00, p10);
    if (p00 > 0.0 && p00 < 1.0 && p10 >= 0.0) return 3;
    edge_param(v3 - v2, nVec, v0 - v2, p01, p11);
    if (p01 > 0.0 && p01 < 1.0 && p11 >= 0.0) return 4;
    edge_param(v1 - v3, nVec, v0 - v3, p02, p12);
    if (p02 > 0.0 && p02 < 1.0 && p12 >= 0.0) return 5;
    if (p00 <= 0.0 && p02 >= 1.0) return 0;
    if (p01 <= 0.0 && p00 >= 1.0) return 1;
    if (p02 <= 0.0 && p01 >= 1.0) return 2;
    return 6;
}
/* MeshCollisionUtils.hpp:2073-2158 */
int orc_dType_EE(const double* vv)
{
    V3 v0 = ld(vv), v1 = ld(vv + 3), v2 = ld(vv + 6), v3 = ld(vv + 9);
    V3 u = v1 - v0, v = v3 - v2, w = v0 - v2;
    double a = norm2(u), b = dot(u, v), c = norm2(v), d = dot(u, w), e = dot(v, w);
    double D = a * c - b * b, tD = D, sN, tN;
    int defaultCase = 8;
    sN = (b * e - c * d);
    if (sN <= 0.0) { tN = e; tD = c; defaultCase = 2; }
    else if (sN >= D) { tN = e + b; tD = c; defaultCase = 5; }
    else {
        tN = (a * e - b * d);
        V3 uxv = cross(u, v);
        if (tN > 0.0 && tN < tD && (dot(uxv, w) == 0.0 || norm2(uxv) < 1.0e-20 * a * c)) {
            if (sN < D / 2) { tN = e; tD = c; defaultCase = 2; }
            else { tN = e + b; tD = c; defaultCase = 5; }
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
    return defaultCase;
}
/* :2279-2325 */
void orc_point_tri_d(const double* v, double* d)
{
    double x[9];
    switch (orc_dType_PT(v)) {
    case 0: std::memcpy(x, v, 24); std::memcpy(x + 3, v + 3, 24); orc_d_PP(x, d); break;
    case 1: std::memcpy(x, v, 24); std::memcpy(x + 3, v + 6, 24); orc_d_PP(x, d); break;
    case 2: std::memcpy(x, v, 24); std::memcpy(x + 3, v + 9, 24); orc_d_PP(x, d); break;
    case 3: std::memcpy(x, v, 24); std::memcpy(x + 3, v + 3, 24); std::memcpy(x + 6, v + 6, 24); orc_d_PE(x, d); break;
    case 4: std::memcpy(x, v, 24); std::memcpy(x + 3, v + 6, 24); std::memcpy(x + 6, v + 9, 24); orc_d_PE(x, d); break;
    case 5: std::memcpy(x, v, 24); std::memcpy(x + 3, v + 9, 24); std::memcpy(x + 6, v + 3, 24); orc_d_PE(x, d); break;
    case 6: orc_d_PT(v, d); break;
    default: *d = -1.0;
    }
}
/* :2327-2383 */
void orc_edge_edge_d(const double* v, double* d)
{
    double x[9];
    auto pp = [&](int a, int b) { std::memcpy(x, v + 3 * a, 24); std::memcpy(x + 3, v + 3 * b, 24); orc_d_PP(x, d); };
    auto pe = [&](int a, int b, int c) { std::memcpy(x, v + 3 * a, 24); std::memcpy(x + 3, v + 3 * b, 24); std::memcpy(x + 6, v + 3 * c, 24); orc_d_PE(x, d); };
    switch (orc_dType_EE(v)) {
    case 0: pp(0, 2); break;
    case 1: pp(0, 3); break;
    case 2: pe(0, 2, 3); break;
    case 3: pp(1, 2); break;
    case 4: pp(1, 3); break;
    case 5: pe(1, 2, 3); break;
    case 6: pe(2, 0, 1); break;
    case 7: pe(3, 0, 1); break;
    case 8: orc_d_EE(v, d); break;
    default: *d = -1.0;
    }
}

/* BarrierFunctions.hpp:56-83 */
void orc_barrier(double d, double dHat, double* b, double* db, double* d2b)
{
    double t2 = d - dHat;
    double lg = std::log(d / dHat);
    if (b) *b = -(d - dHat) * (d - dHat) * lg;
    if (db) *db = t2 * lg * -2.0 - (t2 * t2) / d;
    if (d2b) *d2b = (lg * -2.0 - t2 * 4.0 / d) + 1.0 / (d * d) * (t2 * t2);
}

/* MeshCollisionUtils.hpp:2409-2417 (value), :2419-2477 (gradient), :2494-2762 (Hessian) */
void orc_ee_cross(const double* v, double* c, double* g, double* H)
{
    Diff D = cross_norm(ld(v + 3) - ld(v), ld(v + 9) - ld(v + 6));
    if (c) *c = D.val;
    to_vertices<4>(D, J_CR, g, H);
}
/* :2834-2912 */
void orc_mollifier(const double* v, double eps_x, double* e, double* g, double* H)
{
    double c, cg[12], cH[144];
    orc_ee_cross(v, &c, cg, cH);
    if (c < eps_x) {
        double r = c / eps_x;
        if (e) *e = (-r + 2.0) * r;
        double inv = 1.0 / eps_x;
        double qg = 2.0 * inv * (-inv * c + 1.0);
        double qH = -2.0 / (eps_x * eps_x);
        if (g) for (int i = 0; i < 12; ++i) g[i] = cg[i] * qg;
        if (H) for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) H[i * 12 + j] = cH[i * 12 + j] * qg + (qH * cg[i]) * cg[j];
    }
    else {
        if (e) *e = 1.0;
        if (g) for (int i = 0; i < 12; ++i) g[i] = 0.0;
        if (H) for (int i = 0; i < 144; ++i) H[i] = 0.0;
    }
}

/* SelfCollisionHandler.cpp:2149-2478.  Brute force over all PT / EE pairs: every pair with d < dHat is found
 * by the reference's hash query (the query box has radius sqrt(dHat)), so the active set does not depend on
 * the broad phase.  Output order is canonical (sorted), the reference's is thread/hash dependent. */
int orc_constraint_set(const orc_surf* s, double dHat, int cap, int* mmcvid, int* nC, int capP, int* para, int* para_eIeJ, int* nPara,
    int capK, int* cand, int* nCand, int nthreads)
{
    typedef std::array<int, 4> Q;
    std::vector<std::vector<Q>> csPT(s->nSV), csEE(s->nSE);
    std::vector<std::vector<int>> candPT(s->nSV), candEE(s->nSE);
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 16)
    for (int svI = 0; svI < s->nSV; ++svI) {
        int vI = s->SVI[svI];
        for (int sfI = 0; sfI < s->nSF; ++sfI) {
            int t[3] = { s->SF[sfI], s->SF[(size_t)s->nSF + sfI], s->SF[(size_t)2 * s->nSF + sfI] };
            if (vI == t[0] || vI == t[1] || vI == t[2]) continue;
            if ((codim(s, vI) < 3 && codim(s, t[0]) < 3) || (is_dbc(s, vI) && is_dbc(s, t[0]) && is_dbc(s, t[1]) && is_dbc(s, t[2]))) continue;
            double x[12];
            st(x, vert(s, vI)); st(x + 3, vert(s, t[0])); st(x + 6, vert(s, t[1])); st(x + 9, vert(s, t[2]));
            /* cheap exact reject: the distance to the triangle is >= the distance to its bounding box */
            int dtype = orc_dType_PT(x);
            double d;
            double y[9];
            Q q;
            switch (dtype) {
            case 0: case 1: case 2:
                std::memcpy(y, x, 24); std::memcpy(y + 3, x + 3 * (dtype + 1), 24); orc_d_PP(y, &d);
                q = { -vI - 1, t[dtype], -1, -1 };
                break;
            case 3: case 4: case 5: {
                int a = dtype - 3, b = (dtype - 2) % 3;
                std::memcpy(y, x, 24); std::memcpy(y + 3, x + 3 * (a + 1), 24); std::memcpy(y + 6, x + 3 * (b + 1), 24); orc_d_PE(y, &d);
                q = { -vI - 1, t[a], t[b], -1 };
                break;
            }
            default:
                orc_d_PT(x, &d);
                q = { -vI - 1, t[0], t[1], t[2] };
            }
            if (d < dHat) {
                csPT[svI].push_back(q);
                candPT[svI].push_back(sfI);
            }
        }
    }
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 16)
    for (int eI = 0; eI < s->nSE; ++eI) {
        int a0 = s->SE[2 * eI], a1 = s->SE[2 * eI + 1];
        for (int eJ = eI + 1; eJ < s->nSE; ++eJ) { /* eI > eJ pairs are skipped (:2294) */
            int b0 = s->SE[2 * eJ], b1 = s->SE[2 * eJ + 1];
            if (a0 == b0 || a0 == b1 || a1 == b0 || a1 == b1) continue;
            if ((codim(s, a0) < 3 && codim(s, b0) < 3) || (is_dbc(s, a0) && is_dbc(s, a1) && is_dbc(s, b0) && is_dbc(s, b1))) continue;
            double x[12];
            st(x, vert(s, a0)); st(x + 3, vert(s, a1)); st(x + 6, vert(s, b0)); st(x + 9, vert(s, b1));
            int dtype = orc_dType_EE(x);
            double cr;
            orc_ee_cross(x, &cr, nullptr, nullptr);
            int add_e = (cr < eps_x_of(s, a0, a1, b0, b1)) ? -eJ - 2 : -1;
            double d, y[9];
            Q q;
            auto PP = [&](int va, int ia_, int vb, int ib_) { std::memcpy(y, x + 3 * ia_, 24); std::memcpy(y + 3, x + 3 * ib_, 24); orc_d_PP(y, &d); q = { -va - 1, vb, -1, add_e }; };
            auto PE = [&](int vp, int ip, int ve0, int i0, int ve1, int i1) { std::memcpy(y, x + 3 * ip, 24); std::memcpy(y + 3, x + 3 * i0, 24); std::memcpy(y + 6, x + 3 * i1, 24); orc_d_PE(y, &d); q = { -vp - 1, ve0, ve1, add_e }; };
            switch (dtype) {
            case 0: PP(a0, 0, b0, 2); break;
            case 1: PP(a0, 0, b1, 3); break;
            case 2: PE(a0, 0, b0, 2, b1, 3); break;
            case 3: PP(a1, 1, b0, 2); break;
            case 4: PP(a1, 1, b1, 3); break;
            case 5: PE(a1, 1, b0, 2, b1, 3); break;
            case 6: PE(b0, 2, a0, 0, a1, 1); break;
            case 7: PE(b1, 3, a0, 0, a1, 1); break;
            default:
                orc_d_EE(x, &d);
                q = (add_e <= -2) ? Q{ a0, a1, b0, -b1 - s->nSE - 2 } : Q{ a0, a1, b0, b1 };
            }
            if (d < dHat) {
                csEE[eI].push_back(q);
                candEE[eI].push_back(eJ);
            }
        }
    }
    /* merge (:2411-2476) */
    std::vector<Q> act, par;
    std::vector<std::array<int, 2>> parE, cnd;
    std::map<Q, int> counter;
    for (int svI = 0; svI < s->nSV; ++svI) {
        for (int sfI : candPT[svI]) cnd.push_back({ -svI - 1, sfI });
        for (const Q& c : csPT[svI]) {
            if (c[3] < 0) ++counter[c];
            else act.push_back(c);
        }
    }
    for (int eI = 0; eI < s->nSE; ++eI) {
        for (int eJ : candEE[eI]) cnd.push_back({ eI, eJ });
        for (const Q& c : csEE[eI]) {
            if (c[3] >= 0) act.push_back(c);
            else if (c[3] == -1) ++counter[c];
            else if (c[3] >= -s->nSE - 1) {
                par.push_back({ c[0], c[1], c[2], -1 });
                parE.push_back({ eI, -c[3] - 2 });
            }
            else {
                par.push_back({ c[0], c[1], c[2], -c[3] - s->nSE - 2 });
                parE.push_back({ -1, -1 });
            }
        }
    }
    for (const auto& kv : counter) act.push_back({ kv.first[0], kv.first[1], kv.first[2], -kv.second });
    std::sort(act.begin(), act.end());
    std::vector<int> order(par.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return par[a] < par[b] || (par[a] == par[b] && parE[a] < parE[b]); });
    std::sort(cnd.begin(), cnd.end());
    *nC = (int)act.size();
    *nPara = (int)par.size();
    *nCand = (int)cnd.size();
    if ((int)act.size() > cap || (int)par.size() > capP || (int)cnd.size() > capK) return -1;
    for (size_t i = 0; i < act.size(); ++i) std::memcpy(mmcvid + 4 * i, act[i].data(), 16);
    for (size_t i = 0; i < par.size(); ++i) {
        std::memcpy(para + 4 * i, par[order[i]].data(), 16);
        std::memcpy(para_eIeJ + 2 * i, parE[order[i]].data(), 8);
    }
    for (size_t i = 0; i < cnd.size(); ++i) std::memcpy(cand + 2 * i, cnd[i].data(), 8);
    return 0;
}

/* SelfCollisionHandler.cpp:64-81 + Optimizer.cpp:3290-3353. Returns 1 if some d <= 0 (the reference exits). */
int orc_barrier_energy(const orc_surf* s, const int* mmcvid, int nC, const int* para, const int* para_eIeJ, int nPara, double dHat, double kappa, double* E)
{
    double sum = 0.0;
    int bad = 0;
    for (int c = 0; c < nC; ++c) {
        const int* mm = mmcvid + 4 * c;
        int v[4], nv;
        double d;
        pair_dgH(s, mm, v, nv, d, nullptr, nullptr);
        if (d <= 0.0) { bad = 1; continue; }
        double b;
        orc_barrier(d, dHat, &b, nullptr, nullptr);
        if (mm[3] < -1) b *= -mm[3];
        sum += b;
    }
    for (int c = 0; c < nPara; ++c) {
        const int* mm = para + 4 * c;
        int v[4], nv, ev[4];
        double d;
        pair_dgH(s, mm, v, nv, d, nullptr, nullptr);
        if (d <= 0.0) { bad = 1; continue; }
        para_stencil(s, mm, para_eIeJ + 2 * c, ev);
        double x[12], e, b;
        for (int k = 0; k < 4; ++k) st(x + 3 * k, vert(s, ev[k]));
        orc_mollifier(x, eps_x_of(s, ev[0], ev[1], ev[2], ev[3]), &e, nullptr, nullptr);
        orc_barrier(d, dHat, &b, nullptr, nullptr);
        sum += b * e;
    }
    *E = kappa * sum;
    return bad;
}

/* SelfCollisionHandler.cpp:84-148 (+ Optimizer.cpp:3492-3499) and :2990-3045 */
void orc_barrier_gradient(const orc_surf* s, const int* mmcvid, int nC, const int* para, const int* para_eIeJ, int nPara, double dHat, double kappa, int projectDBC, double* g)
{
    (void)projectDBC; /* Optimizer zeroes projected rows afterwards (Optimizer.cpp:3512-3516) */
    for (int c = 0; c < nC; ++c) {
        const int* mm = mmcvid + 4 * c;
        int v[4], nv;
        double d, gd[12], db;
        pair_dgH(s, mm, v, nv, d, gd, nullptr);
        orc_barrier(d, dHat, nullptr, &db, nullptr);
        double mult = (mm[0] < 0 && nv < 4) ? (double)(-mm[3]) : 1.0;
        double w = kappa * mult * db;
        for (int k = 0; k < nv; ++k)
            for (int i = 0; i < 3; ++i) g[3 * (size_t)v[k] + i] += w * gd[3 * k + i];
    }
    for (int c = 0; c < nPara; ++c) {
        const int* mm = para + 4 * c;
        int v[4], nv, ev[4];
        double d, gd[12], b, db, e, eg[12], x[12];
        pair_dgH(s, mm, v, nv, d, gd, nullptr);
        orc_barrier(d, dHat, &b, &db, nullptr);
        para_stencil(s, mm, para_eIeJ + 2 * c, ev);
        for (int k = 0; k < 4; ++k) st(x + 3 * k, vert(s, ev[k]));
        orc_mollifier(x, eps_x_of(s, ev[0], ev[1], ev[2], ev[3]), &e, eg, nullptr);
        for (int k = 0; k < 4; ++k)
            for (int i = 0; i < 3; ++i) g[3 * (size_t)ev[k] + i] += kappa * b * eg[3 * k + i];
        double w = kappa * e * db; /* slot3 is -1 or a vertex id => multiplicity 1 */
        for (int k = 0; k < nv; ++k)
            for (int i = 0; i < 3; ++i) g[3 * (size_t)v[k] + i] += w * gd[3 * k + i];
    }
}

void orc_barrier_pair_hessian(const orc_surf* s, const int mm[4], double dHat, double kappa, double* H144, int* nvert)
{
    int v[4], nv;
    double d, gd[12], Hd[144], db, d2b;
    pair_dgH(s, mm, v, nv, d, gd, Hd);
    orc_barrier(d, dHat, nullptr, &db, &d2b);
    const int n = 3 * nv;
    double mult = (mm[0] < 0 && nv < 4) ? (double)(-mm[3]) : 1.0;
    double coef = kappa * mult;
    std::vector<double> Hb((size_t)n * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Hb[i * n + j] = ((coef * d2b) * gd[i]) * gd[j] + (coef * db) * Hd[i * n + j];
    orc_makePD(n, Hb.data());
    for (int i = 0; i < 144; ++i) H144[i] = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) H144[i * 12 + j] = Hb[i * n + j];
    *nvert = nv;
}

/* SelfCollisionHandler.cpp:418-561 and :3049-3201 */
void orc_barrier_hessian_csr(const orc_surf* s, const int* mmcvid, int nC, const int* para, const int* para_eIeJ, int nPara, double dHat, double kappa,
    int projectDBC, const int* ia, const int* ja, int index_base, double* a, int nthreads)
{
    CsrSink sink{ ia, ja, index_base, a };
    std::vector<double> blocks((size_t)144 * std::max(nC, 1));
    std::vector<int> nvs(std::max(nC, 1));
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 64)
    for (int c = 0; c < nC; ++c) orc_barrier_pair_hessian(s, mmcvid + 4 * c, dHat, kappa, &blocks[(size_t)144 * c], &nvs[c]);
    for (int c = 0; c < nC; ++c) {
        int v[4];
        int nv = pair_verts(mmcvid + 4 * c, v);
        const double* H = &blocks[(size_t)144 * c];
        for (int i = 0; i < nv; ++i) {
            if (is_proj_dbc(s, v[i], projectDBC)) continue;
            for (int j = 0; j < nv; ++j) {
                if (is_proj_dbc(s, v[j], projectDBC)) continue;
                for (int r = 0; r < 3; ++r)
                    for (int q = 0; q < 3; ++q) sink.add(3 * v[i] + r, 3 * v[j] + q, H[(3 * i + r) * 12 + 3 * j + q]);
            }
        }
    }
    for (int c = 0; c < nPara; ++c) {
        const int* mm = para + 4 * c;
        int v[4], nv, ev[4];
        double d, gd0[12], Hd0[144], b, db, d2b, e, eg[12], eH[144], x[12];
        pair_dgH(s, mm, v, nv, d, gd0, Hd0);
        orc_barrier(d, dHat, &b, &db, &d2b);
        para_stencil(s, mm, para_eIeJ + 2 * c, ev);
        for (int k = 0; k < 4; ++k) st(x + 3 * k, vert(s, ev[k]));
        orc_mollifier(x, eps_x_of(s, ev[0], ev[1], ev[2], ev[3]), &e, eg, eH);
        /* grad_d / H_d embedded in the 4-vertex edge stencil (:3108-3160) */
        double gd[12] = { 0 }, Hd[144] = { 0 };
        int map[4];
        for (int k = 0; k < nv; ++k) {
            map[k] = -1;
            for (int i = 0; i < 4; ++i)
                if (ev[i] == v[k]) map[k] = i;
        }
        const int n0 = 3 * nv;
        for (int k = 0; k < nv; ++k) {
            if (map[k] < 0) continue;
            for (int i = 0; i < 3; ++i) gd[3 * map[k] + i] = gd0[3 * k + i];
            for (int l = 0; l < nv; ++l) {
                if (map[l] < 0) continue;
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) Hd[(3 * map[k] + i) * 12 + 3 * map[l] + j] = Hd0[(3 * k + i) * n0 + 3 * l + j];
            }
        }
        double H[144];
        for (int i = 0; i < 12; ++i)
            for (int j = 0; j < 12; ++j) {
                double cross_ij = ((kappa * db) * gd[i]) * eg[j];
                double cross_ji = ((kappa * db) * gd[j]) * eg[i];
                H[i * 12 + j] = cross_ij + cross_ji + (kappa * b) * eH[i * 12 + j] + ((kappa * e * d2b) * gd[i]) * gd[j] + (kappa * e * db) * Hd[i * 12 + j];
            }
        orc_makePD(12, H);
        for (int i = 0; i < 4; ++i) {
            if (is_proj_dbc(s, ev[i], projectDBC)) continue;
            for (int j = 0; j < 4; ++j) {
                if (is_proj_dbc(s, ev[j], projectDBC)) continue;
                for (int r = 0; r < 3; ++r)
                    for (int q = 0; q < 3; ++q) sink.add(3 * ev[i] + r, 3 * ev[j] + q, H[(3 * i + r) * 12 + 3 * j + q]);
            }
        }
    }
}

} // extern "C"
