// elastic.cuh -- per-tet material math on registers: psi(sigma), dpsi, d2psi, B-block coefficients,
// first Piola-Kirchhoff stress, sigma-space PSD projection.  NeoHookean (ENERGY=0) and
// FixedCoRot (ENERGY=1); formulas follow the reference's
//   src/Energy/Physics_Elasticity/NeoHookeanEnergy.cpp:55-153
//   src/Energy/Physics_Elasticity/FixedCoRotEnergy.cpp:62-153
//   src/Energy/Energy.cpp:448-529 (A / B blocks, eps floor 1e-6, projection before rotation)
//   src/Utils/IglUtils.hpp:119-177 (makePD / makePD2d)
#pragma once
#include "common.cuh"
#include "svd3.cuh"

namespace ipcgpu {

template <int ENERGY>
DEV double psi(const double* s, double mu, double lam)
{
    if (ENERGY == 0) {
        if (mu == 0.0 && lam == 0.0) return 0.0;
        double s2 = s[0] * s[0] + s[1] * s[1] + s[2] * s[2];
        double lJ = log(s[0] * s[1] * s[2]);
        return mu / 2.0 * (s2 - 3.0) - (mu - lam / 2.0 * lJ) * lJ;
    }
    else {
        double a = s[0] - 1.0, b = s[1] - 1.0, c = s[2] - 1.0;
        double Jm1 = s[0] * s[1] * s[2] - 1.0;
        return mu * (a * a + b * b + c * c) + lam / 2.0 * Jm1 * Jm1;
    }
}

// sigma-space derivative bundle used by gradient-free Hessian assembly
struct SigmaDerivs {
    double dE[3];     // dpsi/dsigma
    double A[6];      // d2psi/dsigma2 : a00 a01 a02 a11 a12 a22
    double BL[3];     // B-block "left" coefficients for pairs (0,1) (1,2) (2,0)
};

template <int ENERGY>
DEV void sigma_derivs(const double* s, double mu, double lam, SigmaDerivs& d)
{
    if (ENERGY == 0) {
        if (mu == 0.0 && lam == 0.0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) d.dE[i] = d.BL[i] = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) d.A[i] = 0.0;
            return;
        }
        const double J = s[0] * s[1] * s[2];
        const double lJ = log(J);
        const double i0 = 1.0 / s[0], i1 = 1.0 / s[1], i2 = 1.0 / s[2];
        d.dE[0] = mu * (s[0] - i0) + lam * i0 * lJ;
        d.dE[1] = mu * (s[1] - i1) + lam * i1 * lJ;
        d.dE[2] = mu * (s[2] - i2) + lam * i2 * lJ;
        const double k = lam * (lJ - 1.0);
        d.A[0] = mu * (1.0 + i0 * i0) - k * (i0 * i0);
        d.A[3] = mu * (1.0 + i1 * i1) - k * (i1 * i1);
        d.A[5] = mu * (1.0 + i2 * i2) - k * (i2 * i2);
        d.A[1] = lam * i0 * i1;
        d.A[4] = lam * i1 * i2;
        d.A[2] = lam * i2 * i0;
        const double mid = mu - lam * lJ;
        d.BL[0] = (mu + mid * i0 * i1) / 2.0;
        d.BL[1] = (mu + mid * i1 * i2) / 2.0;
        d.BL[2] = (mu + mid * i2 * i0) / 2.0;
    }
    else {
        const double J = s[0] * s[1] * s[2];
        const double n0 = s[1] * s[2], n1 = s[2] * s[0], n2 = s[0] * s[1];
        const double k = lam * (J - 1.0);
        const double m2 = 2.0 * mu;
        d.dE[0] = m2 * (s[0] - 1.0) + n0 * k;
        d.dE[1] = m2 * (s[1] - 1.0) + n1 * k;
        d.dE[2] = m2 * (s[2] - 1.0) + n2 * k;
        d.A[0] = m2 + lam * n0 * n0;
        d.A[3] = m2 + lam * n1 * n1;
        d.A[5] = m2 + lam * n2 * n2;
        d.A[1] = lam * (s[2] * (J - 1.0) + n0 * n1);
        d.A[2] = lam * (s[1] * (J - 1.0) + n0 * n2);
        d.A[4] = lam * (s[0] * (J - 1.0) + n2 * n1);
        const double hl = lam / 2.0;
        d.BL[0] = mu - hl * s[2] * (J - 1.0);
        d.BL[1] = mu - hl * s[0] * (J - 1.0);
        d.BL[2] = mu - hl * s[1] * (J - 1.0);
    }
}

// P = dpsi/dF   [NeoHookeanEnergy.cpp:138-153, FixedCoRotEnergy.cpp:145-153, IglUtils.hpp:436-464]
template <int ENERGY>
DEV void pk1(const M3& F, const M3& U, const double* s, const M3& V, double mu, double lam, M3& P)
{
    M3 C;
    C(0, 0) = F(1, 1) * F(2, 2) - F(1, 2) * F(2, 1);
    C(0, 1) = F(1, 2) * F(2, 0) - F(1, 0) * F(2, 2);
    C(0, 2) = F(1, 0) * F(2, 1) - F(1, 1) * F(2, 0);
    C(1, 0) = F(0, 2) * F(2, 1) - F(0, 1) * F(2, 2);
    C(1, 1) = F(0, 0) * F(2, 2) - F(0, 2) * F(2, 0);
    C(1, 2) = F(0, 1) * F(2, 0) - F(0, 0) * F(2, 1);
    C(2, 0) = F(0, 1) * F(1, 2) - F(0, 2) * F(1, 1);
    C(2, 1) = F(0, 2) * F(1, 0) - F(0, 0) * F(1, 2);
    C(2, 2) = F(0, 0) * F(1, 1) - F(0, 1) * F(1, 0);
    const double J = s[0] * s[1] * s[2];
    if (ENERGY == 0) {
        if (mu == 0.0 && lam == 0.0) {
#pragma unroll
            for (int q = 0; q < 9; ++q) P.m[q] = 0.0;
            return;
        }
        const double lJ = log(J);
        const double invJ = 1.0 / J;
        const double k = lam * lJ - mu;
#pragma unroll
        for (int q = 0; q < 9; ++q) P.m[q] = mu * F.m[q] + k * (C.m[q] * invJ);
    }
    else {
        const double k = lam * (J - 1.0);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double R = U(i, 0) * V(j, 0) + U(i, 1) * V(j, 1) + U(i, 2) * V(j, 2);
                P(i, j) = 2.0 * mu * (F(i, j) - R) + k * C(i, j);
            }
    }
}

// ---- makePD for the 3x3 sigma-space block (IglUtils.hpp:119-137) ---------------------------------
// The reference returns the input untouched when the smallest eigenvalue is >= 0; we test positive
// semidefiniteness through the principal minors and only run the Jacobi eigen-solver when it fails.
template <int P, int Q>
DEV void jacobi_rot3(double (&a)[3][3], double (&v)[3][3])
{
    double apq = a[P][Q];
    if (apq == 0.0) return;
    double theta = (a[Q][Q] - a[P][P]) / (2.0 * apq);
    double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
    double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double akp = a[k][P], akq = a[k][Q];
        a[k][P] = c * akp - s * akq;
        a[k][Q] = s * akp + c * akq;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double apk = a[P][k], aqk = a[Q][k];
        a[P][k] = c * apk - s * aqk;
        a[Q][k] = s * apk + c * aqk;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double vkp = v[k][P], vkq = v[k][Q];
        v[k][P] = c * vkp - s * vkq;
        v[k][Q] = s * vkp + c * vkq;
    }
}

DEV void make_pd3(double* A /* a00 a01 a02 a11 a12 a22 */)
{
    const double a00 = A[0], a01 = A[1], a02 = A[2], a11 = A[3], a12 = A[4], a22 = A[5];
    const double m01 = a00 * a11 - a01 * a01, m02 = a00 * a22 - a02 * a02, m12 = a11 * a22 - a12 * a12;
    const double det = a00 * m12 - a01 * (a01 * a22 - a12 * a02) + a02 * (a01 * a12 - a11 * a02);
    if (a00 >= 0.0 && a11 >= 0.0 && a22 >= 0.0 && m01 >= 0.0 && m02 >= 0.0 && m12 >= 0.0 && det >= 0.0) return;
    double a[3][3] = { { a00, a01, a02 }, { a01, a11, a12 }, { a02, a12, a22 } };
    double v[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        double dg = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-26 * dg || off <= 1e-300) break;
        jacobi_rot3<0, 1>(a, v);
        jacobi_rot3<0, 2>(a, v);
        jacobi_rot3<1, 2>(a, v);
    }
    const double l0 = a[0][0], l1 = a[1][1], l2 = a[2][2];
    if (l0 >= 0.0 && l1 >= 0.0 && l2 >= 0.0) return; // eigenvalues()[0] >= 0 : unchanged
    const double e0 = fmax(l0, 0.0), e1 = fmax(l1, 0.0), e2 = fmax(l2, 0.0);
    A[0] = v[0][0] * e0 * v[0][0] + v[0][1] * e1 * v[0][1] + v[0][2] * e2 * v[0][2];
    A[1] = v[0][0] * e0 * v[1][0] + v[0][1] * e1 * v[1][1] + v[0][2] * e2 * v[1][2];
    A[2] = v[0][0] * e0 * v[2][0] + v[0][1] * e1 * v[2][1] + v[0][2] * e2 * v[2][2];
    A[3] = v[1][0] * e0 * v[1][0] + v[1][1] * e1 * v[1][1] + v[1][2] * e2 * v[1][2];
    A[4] = v[1][0] * e0 * v[2][0] + v[1][1] * e1 * v[2][1] + v[1][2] * e2 * v[2][2];
    A[5] = v[2][0] * e0 * v[2][0] + v[2][1] * e1 * v[2][1] + v[2][2] * e2 * v[2][2];
}

// makePD2d on the symmetric block [[p, q],[q, r]]   (IglUtils.hpp:138-177)
DEV void make_pd2(double& p, double& q, double& r)
{
    const double a = p, b = q, d = r;
    const double b2 = b * b;
    const double D = a * d - b2;
    const double Th = (a + d) / 2.0;
    const double sq = sqrt(Th * Th - D);
    const double L2 = Th - sq;
    if (L2 < 0.0) {
        const double L1 = Th + sq;
        if (L1 <= 0.0) {
            p = q = r = 0.0;
        }
        else if (b2 == 0.0) {
            p = L1;
            q = r = 0.0;
        }
        else {
            const double L1md = L1 - d;
            const double ratio = L1md / L1;
            p = ratio * L1md;
            q = b * ratio;
            r = b2 / L1;
        }
    }
}

} // namespace ipcgpu
