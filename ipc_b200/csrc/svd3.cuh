// svd3.cuh -- per-thread 3x3 rotation-variant SVD, all state in registers.
//
// Same convention the reference's elastic path relies on (AutoFlipSVD.hpp:35-54 ->
// JIXIE::singularValueDecomposition, ImplicitQRSVD.h:687-850): F = U diag(s) V^T with
// det U = det V = +1, |s0| >= |s1| >= |s2| and only s2 allowed negative.  The algorithm is the
// implicit-shift bidiagonal QR with the same stopping rule (tol = 1024 eps max(||B||/2, 1)),
// written here with compile-time row/column indices so that nothing is ever indexed dynamically
// (no local memory) -- every Givens rotation touches named registers only.
#pragma once
#include "common.cuh"

namespace ipcgpu {

struct Giv {
    double c, s;
};

// (c -s; s c)(a;b) = (*;0)     [ImplicitQRSVD.h:137-150]
DEV Giv giv(double a, double b)
{
    Giv g{ 1.0, 0.0 };
    double d = a * a + b * b;
    if (d != 0.0) {
        double t = 1.0 / sqrt(d);
        g.c = a * t;
        g.s = -b * t;
    }
    return g;
}
// (c -s; s c)(a;b) = (0;*)     [ImplicitQRSVD.h:157-170]
DEV Giv giv_unconv(double a, double b)
{
    Giv g{ 0.0, 1.0 };
    double d = a * a + b * b;
    if (d != 0.0) {
        double t = 1.0 / sqrt(d);
        g.s = a * t;
        g.c = b * t;
    }
    return g;
}
template <int I, int K>
DEV void row_rot(const Giv& g, M3& A)
{
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double t1 = A(I, j), t2 = A(K, j);
        A(I, j) = g.c * t1 - g.s * t2;
        A(K, j) = g.s * t1 + g.c * t2;
    }
}
template <int I, int K>
DEV void col_rot(const Giv& g, M3& A)
{
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double t1 = A(j, I), t2 = A(j, K);
        A(j, I) = g.c * t1 - g.s * t2;
        A(j, K) = g.s * t1 + g.c * t2;
    }
}

template <bool UV>
DEV void zero_chase(M3& H, M3& U, M3& V) // ImplicitQRSVD.h:252-301
{
    Giv r1 = giv(H(0, 0), H(1, 0));
    Giv r2;
    if (H(1, 0) != 0.0)
        r2 = giv(H(0, 0) * H(0, 1) + H(1, 0) * H(1, 1), H(0, 0) * H(0, 2) + H(1, 0) * H(1, 2));
    else
        r2 = giv(H(0, 1), H(0, 2));
    row_rot<0, 1>(r1, H);
    col_rot<1, 2>(r2, H);
    if (UV) col_rot<1, 2>(r2, V);
    Giv r3 = giv(H(1, 1), H(2, 1));
    row_rot<1, 2>(r3, H);
    if (UV) {
        col_rot<0, 1>(r1, U);
        col_rot<1, 2>(r3, U);
    }
}

// 2x2 polar + SVD of the block B(T..T+1, T..T+1)   [ImplicitQRSVD.h:398-418, 454-518, 570-585]
template <int T, bool UV>
DEV void process2(const M3& B, M3& U, double* sg /* sigma[3] */, M3& V)
{
    constexpr int other = (T == 1) ? 0 : 2;
    sg[other] = B(other, other);
    const double a00 = B(T, T), a01 = B(T, T + 1), a10 = B(T + 1, T), a11 = B(T + 1, T + 1);
    double x0 = a00 + a11, x1 = a10 - a01;
    double den = sqrt(x0 * x0 + x1 * x1);
    Giv u{ 1.0, 0.0 };
    if (den != 0.0) {
        u.c = x0 / den;
        u.s = -x1 / den;
    }
    const double x = u.c * a00 - u.s * a10;
    const double y = u.c * a01 - u.s * a11;
    const double z = u.s * a01 + u.c * a11;
    double cosine, sine, s0, s1;
    if (y == 0.0) {
        cosine = 1.0;
        sine = 0.0;
        s0 = x;
        s1 = z;
    }
    else {
        double tau = 0.5 * (x - z);
        double w = sqrt(tau * tau + y * y);
        double t = (tau > 0.0) ? y / (tau + w) : y / (tau - w);
        cosine = 1.0 / sqrt(t * t + 1.0);
        sine = -t * cosine;
        double c2 = cosine * cosine, csy = 2.0 * cosine * sine * y, s2 = sine * sine;
        s0 = c2 * x - csy + s2 * z;
        s1 = s2 * x + csy + c2 * z;
    }
    Giv v;
    if (s0 < s1) {
        double tmp = s0;
        s0 = s1;
        s1 = tmp;
        v.c = -sine;
        v.s = cosine;
    }
    else {
        v.c = cosine;
        v.s = sine;
    }
    sg[T] = s0;
    sg[T + 1] = s1;
    if (UV) {
        Giv uu{ u.c * v.c - u.s * v.s, u.s * v.c + u.c * v.s };
        col_rot<T, T + 1>(uu, U);
        col_rot<T, T + 1>(v, V);
    }
}

template <int A, int B_>
DEV void swap_col(M3& M)
{
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double t = M(r, A);
        M(r, A) = M(r, B_);
        M(r, B_) = t;
    }
}
template <int A>
DEV void neg_col(M3& M)
{
#pragma unroll
    for (int r = 0; r < 3; ++r) M(r, A) = -M(r, A);
}
DEV void dswap(double& a, double& b)
{
    double t = a;
    a = b;
    b = t;
}

template <bool UV>
DEV void sort_t0(M3& U, double* s, M3& V) // ImplicitQRSVD.h:599-636
{
    if (fabs(s[1]) >= fabs(s[2])) {
        if (s[1] < 0) {
            s[1] = -s[1];
            s[2] = -s[2];
            if (UV) { neg_col<1>(U); neg_col<2>(U); }
        }
        return;
    }
    if (s[2] < 0) {
        s[1] = -s[1];
        s[2] = -s[2];
        if (UV) { neg_col<1>(U); neg_col<2>(U); }
    }
    dswap(s[1], s[2]);
    if (UV) { swap_col<1, 2>(U); swap_col<1, 2>(V); }
    if (s[1] > s[0]) {
        dswap(s[0], s[1]);
        if (UV) { swap_col<0, 1>(U); swap_col<0, 1>(V); }
    }
    else if (UV) {
        neg_col<2>(U);
        neg_col<2>(V);
    }
}
template <bool UV>
DEV void sort_t1(M3& U, double* s, M3& V) // ImplicitQRSVD.h:641-678
{
    if (fabs(s[0]) >= s[1]) {
        if (s[0] < 0) {
            s[0] = -s[0];
            s[2] = -s[2];
            if (UV) { neg_col<0>(U); neg_col<2>(U); }
        }
        return;
    }
    dswap(s[0], s[1]);
    if (UV) { swap_col<0, 1>(U); swap_col<0, 1>(V); }
    if (fabs(s[1]) < fabs(s[2])) {
        dswap(s[1], s[2]);
        if (UV) { swap_col<1, 2>(U); swap_col<1, 2>(V); }
    }
    else if (UV) {
        neg_col<1>(U);
        neg_col<1>(V);
    }
    if (s[1] < 0) {
        s[1] = -s[1];
        s[2] = -s[2];
        if (UV) { neg_col<1>(U); neg_col<2>(U); }
    }
}

// UV=false computes singular values only (energy evaluation needs nothing else).
template <bool UV>
DEV void svd3(const M3& F, M3& U, double* sg, M3& V)
{
    M3 B = F;
    if (UV) {
#pragma unroll
        for (int q = 0; q < 9; ++q) U.m[q] = V.m[q] = (q % 4 == 0) ? 1.0 : 0.0;
    }
    { // makeUpperBidiag, ImplicitQRSVD.h:314-328
        Giv r = giv(B(1, 0), B(2, 0));
        row_rot<1, 2>(r, B);
        if (UV) col_rot<1, 2>(r, U);
        zero_chase<UV>(B, U, V);
    }
    double a1 = B(0, 0), b1 = B(0, 1), a2 = B(1, 1), a3 = B(2, 2), b2 = B(1, 2);
    double g1 = a1 * b1, g2 = a2 * b2;
    double tol = 1024.0 * 2.220446049250313e-16;
    tol *= fmax(0.5 * sqrt(a1 * a1 + a2 * a2 + a3 * a3 + b1 * b1 + b2 * b2), 1.0);
    int guard = 0;
    while (fabs(b2) > tol && fabs(b1) > tol && fabs(a1) > tol && fabs(a2) > tol && fabs(a3) > tol && guard < 64) {
        // Wilkinson shift, ImplicitQRSVD.h:552-565
        double wa = a2 * a2 + b1 * b1, wc = a3 * a3 + b2 * b2;
        double d = 0.5 * (wa - wc);
        double bs = g2 * g2;
        double mu = wc - copysign(bs / (fabs(d) + sqrt(d * d + bs)), d);
        Giv r = giv(a1 * a1 - mu, g1);
        col_rot<0, 1>(r, B);
        if (UV) col_rot<0, 1>(r, V);
        zero_chase<UV>(B, U, V);
        a1 = B(0, 0); b1 = B(0, 1); a2 = B(1, 1); a3 = B(2, 2); b2 = B(1, 2);
        g1 = a1 * b1;
        g2 = a2 * b2;
        ++guard;
    }
    if (fabs(b2) <= tol) {
        process2<0, UV>(B, U, sg, V);
        sort_t0<UV>(U, sg, V);
    }
    else if (fabs(b1) <= tol) {
        process2<1, UV>(B, U, sg, V);
        sort_t1<UV>(U, sg, V);
    }
    else if (fabs(a2) <= tol) {
        Giv r1 = giv_unconv(B(1, 2), B(2, 2));
        row_rot<1, 2>(r1, B);
        if (UV) col_rot<1, 2>(r1, U);
        process2<0, UV>(B, U, sg, V);
        sort_t0<UV>(U, sg, V);
    }
    else if (fabs(a3) <= tol) {
        Giv r1 = giv(B(1, 1), B(1, 2));
        col_rot<1, 2>(r1, B);
        if (UV) col_rot<1, 2>(r1, V);
        Giv r2 = giv(B(0, 0), B(0, 2));
        col_rot<0, 2>(r2, B);
        if (UV) col_rot<0, 2>(r2, V);
        process2<0, UV>(B, U, sg, V);
        sort_t0<UV>(U, sg, V);
    }
    else {
        Giv r1 = giv_unconv(B(0, 1), B(1, 1));
        row_rot<0, 1>(r1, B);
        if (UV) col_rot<0, 1>(r1, U);
        Giv r2 = giv_unconv(B(0, 2), B(2, 2));
        row_rot<0, 2>(r2, B);
        if (UV) col_rot<0, 2>(r2, U);
        process2<1, UV>(B, U, sg, V);
        sort_t1<UV>(U, sg, V);
    }
}

} // namespace ipcgpu
