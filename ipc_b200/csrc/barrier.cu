// barrier.cu -- per-pair barrier energy / gradient / PSD-projected Hessian over the active constraint set (sm_100a).
//
// Reference being replaced: SelfCollisionHandler<3>::evaluateConstraints (:64-81), leftMultiplyConstraintJacobianT
// (:84-148, serial "TODO: parallelize"), augmentIPHessian (:418-561, serial CSR add), augmentParaEEGradient/Hessian
// (:2990-3201) and the composition in Optimizer.cpp:3290-3353, 3492-3502, 3693-3695.
//
//  * one thread per pair; the squared distance, its gradient and Hessian come from contact.cuh;
//  * makePD on the 6/9/12-square block (IglUtils.hpp:119-137) is a cyclic Jacobi eigen-solver whose three
//    12x12 work matrices live in shared memory, element-major / lane-minor (element e of lane t at e*32+t), so every
//    access of a warp is conflict-free even though p,q are data dependent per sweep position only;
//  * gradients and CSR values are accumulated with FP64 red.global.add (pair counts are orders of magnitude below the
//    tet count; the elastic path stays deterministic, the barrier scatter is order-free to ~1 ulp of the sum).
#include "contact.cuh"
#include "kernels.h"

namespace ipcgpu {

struct PairStencil {
    int v[4];
    int nv;   // 2 PP, 3 PE, 4 PT/EE
    int kind; // 0 PT, 1 EE, 2 PE, 3 PP
    double mult;
};

DEV PairStencil decode(int4 mm)
{
    PairStencil s;
    s.mult = 1.0;
    if (mm.x >= 0) {
        s.v[0] = mm.x; s.v[1] = mm.y; s.v[2] = mm.z; s.v[3] = mm.w;
        s.nv = 4; s.kind = 1;
    }
    else {
        s.v[0] = -mm.x - 1; s.v[1] = mm.y; s.v[2] = mm.z; s.v[3] = mm.w;
        if (mm.z < 0) { s.nv = 2; s.kind = 3; s.mult = (double)(-mm.w); }
        else if (mm.w < 0) { s.nv = 3; s.kind = 2; s.mult = (double)(-mm.w); }
        else { s.nv = 4; s.kind = 0; }
    }
    return s;
}

DEV double pair_distance(const PairStencil& s, const V3* x)
{
    switch (s.kind) {
    case 0: return d_PT(x[0], x[1], x[2], x[3]);
    case 1: return d_EE(x[0], x[1], x[2], x[3]);
    case 2: return d_PE(x[0], x[1], x[2]);
    default: return d_PP(x[0], x[1]);
    }
}

// distance + gradient (12) [+ Hessian through puth]
template <typename PutH>
__device__ inline double pair_derivs(const PairStencil& s, const V3* x, double* g, bool want_h, PutH puth)
{
    if (s.kind == 3) {
        V3 r = 2.0 * (x[0] - x[1]);
        g[0] = r.x; g[1] = r.y; g[2] = r.z; g[3] = -r.x; g[4] = -r.y; g[5] = -r.z;
        if (want_h)
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j) puth(i, j, (i == j) ? 2.0 : ((i % 3 == j % 3) ? -2.0 : 0.0));
        return norm2(x[0] - x[1]);
    }
    Diff D;
    if (s.kind == 0) plane_dist(x[0] - x[1], x[2] - x[1], x[3] - x[1], D);
    else if (s.kind == 1) plane_dist(x[2] - x[0], x[1] - x[0], x[3] - x[2], D);
    else line_dist(x[0] - x[1], x[2] - x[1], D);
    diff_to_vertices(D, s.kind, s.nv, true, want_h, [&](int i, double v) { g[i] = v; }, puth);
    return D.val;
}

DEV int csr_find(const int* __restrict__ ia, const int* __restrict__ ja, int base, int row, int col)
{
    int lo = ia[row] - base, hi = ia[row + 1] - base;
    const int end = hi, target = col + base;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (ja[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    return (lo < end && ja[lo] == target) ? lo : -1;
}

DEV bool proj_dbc(const uint8_t* dbc, int v, int projectDBC) { return dbc && (dbc[v] == 1 || (dbc[v] == 2 && projectDBC)); }

// mollifier e(x) on the 4-vertex edge stencil: value, gradient (12), Hessian via puth   [:2851-2912]
template <typename PutH>
__device__ inline double mollifier(const V3* ex, double eps_x, double* eg, bool want_h, PutH puth)
{
    Diff C;
    cross_norm(ex[1] - ex[0], ex[3] - ex[2], C);
    if (!(C.val < eps_x)) {
        for (int i = 0; i < 12; ++i) eg[i] = 0.0;
        if (want_h)
            for (int i = 0; i < 12; ++i)
                for (int j = 0; j < 12; ++j) puth(i, j, 0.0);
        return 1.0;
    }
    const double inv = 1.0 / eps_x;
    const double qg = 2.0 * inv * (-inv * C.val + 1.0);
    const double qH = -2.0 / (eps_x * eps_x);
    double cg[12];
    diff_to_vertices(C, 3, 4, true, false, [&](int i, double v) { cg[i] = v; }, [](int, int, double) {});
    for (int i = 0; i < 12; ++i) eg[i] = cg[i] * qg;
    if (want_h)
        diff_to_vertices(C, 3, 4, false, true, [](int, double) {}, [&](int i, int j, double v) { puth(i, j, v * qg + (qH * cg[i]) * cg[j]); });
    const double r = C.val / eps_x;
    return (-r + 2.0) * r;
}

DEV void para_edge_stencil(int4 mm, int2 e, const int* __restrict__ SE, int* ev)
{
    if (mm.w >= 0 && mm.x >= 0) { ev[0] = mm.x; ev[1] = mm.y; ev[2] = mm.z; ev[3] = mm.w; }
    else {
        ev[0] = SE[2 * e.x]; ev[1] = SE[2 * e.x + 1];
        ev[2] = SE[2 * e.y]; ev[3] = SE[2 * e.y + 1];
    }
}

// -----------------------------------------------------------------------------------------------------------
// energy: kappa * sum( mult*b(d) ) + kappa * sum( e*b(d) )      (Optimizer.cpp:3290-3353)
// -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_barrier_energy(BarrierArgs p, double* __restrict__ partials, int* __restrict__ bad)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    double val = 0.0;
    if (c < p.nC + p.nP) {
        const bool is_para = c >= p.nC;
        const int4 mm = is_para ? p.para[c - p.nC] : p.cs[c];
        PairStencil s = decode(mm);
        V3 x[4];
        for (int k = 0; k < s.nv; ++k) x[k] = load_vertex(p.V, p.nV, s.v[k]);
        const double d = pair_distance(s, x);
        if (!(d > 0.0)) atomicExch(bad, 1);
        else {
            double b, db, d2b;
            barrier_all(d, p.dHat, b, db, d2b);
            if (!is_para) val = (mm.x < 0 && mm.w < -1) ? b * (double)(-mm.w) : b;
            else {
                int ev[4];
                para_edge_stencil(mm, p.para_e[c - p.nC], p.SE, ev);
                V3 ex[4];
                for (int k = 0; k < 4; ++k) ex[k] = load_vertex(p.V, p.nV, ev[k]);
                double eg[12];
                const double e = mollifier(ex, eps_x_rest(p.Vrest, p.nV, ev[0], ev[1], ev[2], ev[3]), eg, false, [](int, int, double) {});
                val = b * e;
            }
        }
    }
    __shared__ double sm[8];
    double w = warp_sum(val);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += sm[i];
        partials[blockIdx.x] = s;
    }
}

// -----------------------------------------------------------------------------------------------------------
// gradient: g += kappa*mult*b'(d) grad d   (+ para-EE: kappa*b*grad e + kappa*e*b' grad d)
// -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_barrier_gradient(BarrierArgs p, double* __restrict__ g)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.nC + p.nP) return;
    const bool is_para = c >= p.nC;
    const int4 mm = is_para ? p.para[c - p.nC] : p.cs[c];
    PairStencil s = decode(mm);
    V3 x[4];
    for (int k = 0; k < s.nv; ++k) x[k] = load_vertex(p.V, p.nV, s.v[k]);
    double gd[12];
    const double d = pair_derivs(s, x, gd, false, [](int, int, double) {});
    double b, db, d2b;
    barrier_all(d, p.dHat, b, db, d2b);
    double w;
    if (!is_para) w = p.kappa * s.mult * db;
    else {
        int ev[4];
        para_edge_stencil(mm, p.para_e[c - p.nC], p.SE, ev);
        V3 ex[4];
        for (int k = 0; k < 4; ++k) ex[k] = load_vertex(p.V, p.nV, ev[k]);
        double eg[12];
        const double e = mollifier(ex, eps_x_rest(p.Vrest, p.nV, ev[0], ev[1], ev[2], ev[3]), eg, false, [](int, int, double) {});
        for (int k = 0; k < 4; ++k)
            for (int i = 0; i < 3; ++i) atomicAdd(g + 3 * (size_t)ev[k] + i, p.kappa * b * eg[3 * k + i]);
        w = p.kappa * e * db; // slot 3 is -1 (or a vertex id): multiplicity 1
    }
    for (int k = 0; k < s.nv; ++k)
        for (int i = 0; i < 3; ++i) atomicAdd(g + 3 * (size_t)s.v[k] + i, w * gd[3 * k + i]);
}

// -----------------------------------------------------------------------------------------------------------
// Hessian: makePD( kappa*mult*(b'' g g^T + b' H_d) ) scattered into the CSR (upper triangle)
// -----------------------------------------------------------------------------------------------------------
constexpr int kHW = 32; // pairs per CTA (one warp); 3 matrices x 144 x 32 lanes x 8 B = 110,592 B of shared memory

__device__ inline void jacobi_psd(int n, double* __restrict__ A, double* __restrict__ Vv, const double* __restrict__ H0, double* __restrict__ out, int lane)
{
    // A: work copy (destroyed), Vv: eigenvectors, H0: original, out: result (may alias H0). element (i,j) at (i*12+j)*32+lane
#define EL(M, i, j) M[((i) * 12 + (j)) * kHW + lane]
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            EL(A, i, j) = EL(H0, i, j);
            EL(Vv, i, j) = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int i = 0; i < n; ++i) {
            double t = EL(A, i, i);
            dg += t * t;
            for (int j = i + 1; j < n; ++j) {
                double u = EL(A, i, j);
                off += u * u;
            }
        }
        if (off <= 1e-26 * dg || off <= 1e-300) break; // off-diagonal mass at rounding level (eigenvalue error is second order in it)
        for (int pI = 0; pI < n; ++pI)
            for (int q = pI + 1; q < n; ++q) {
                const double apq = EL(A, pI, q);
                if (apq == 0.0) continue;
                const double theta = (EL(A, q, q) - EL(A, pI, pI)) / (2.0 * apq);
                const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = EL(A, k, pI), akq = EL(A, k, q);
                    EL(A, k, pI) = c * akp - s * akq;
                    EL(A, k, q) = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = EL(A, pI, k), aqk = EL(A, q, k);
                    EL(A, pI, k) = c * apk - s * aqk;
                    EL(A, q, k) = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = EL(Vv, k, pI), vkq = EL(Vv, k, q);
                    EL(Vv, k, pI) = c * vkp - s * vkq;
                    EL(Vv, k, q) = s * vkp + c * vkq;
                }
            }
    }
    bool neg = false;
    for (int i = 0; i < n; ++i) neg = neg || (EL(A, i, i) < 0.0);
    if (!neg) return; // lambda_min >= 0: the reference returns the matrix unchanged
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) {
            double sacc = 0.0;
            for (int k = 0; k < n; ++k) {
                const double lam = EL(A, k, k);
                if (lam > 0.0) sacc += EL(Vv, i, k) * lam * EL(Vv, j, k);
            }
            EL(out, i, j) = sacc;
            EL(out, j, i) = sacc;
        }
#undef EL
}

__global__ void __launch_bounds__(kHW) k_barrier_hessian(BarrierArgs p, double* __restrict__ a, int* __restrict__ err)
{
    extern __shared__ double sm[];
    double* H = sm;
    double* A = sm + 144 * kHW;
    double* Vv = sm + 288 * kHW;
    const int lane = threadIdx.x;
    const int c = blockIdx.x * kHW + lane;
    if (c >= p.nC + p.nP) return;
#define HE(i, j) H[((i) * 12 + (j)) * kHW + lane]
    const bool is_para = c >= p.nC;
    const int4 mm = is_para ? p.para[c - p.nC] : p.cs[c];
    PairStencil s = decode(mm);
    V3 x[4];
    for (int k = 0; k < s.nv; ++k) x[k] = load_vertex(p.V, p.nV, s.v[k]);
    int n, rows[4];
    if (!is_para) {
        n = 3 * s.nv;
        double gd[12];
        const double d = pair_derivs(s, x, gd, true, [&](int i, int j, double v) { HE(i, j) = v; });
        double b, db, d2b;
        barrier_all(d, p.dHat, b, db, d2b);
        const double coef = p.kappa * s.mult;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) HE(i, j) = ((coef * d2b) * gd[i]) * gd[j] + (coef * db) * HE(i, j);
        for (int k = 0; k < 4; ++k) rows[k] = (k < s.nv) ? s.v[k] : -1;
    }
    else {
        // mollified pair on the two-edge stencil (:3049-3173)
        n = 12;
        int ev[4];
        para_edge_stencil(mm, p.para_e[c - p.nC], p.SE, ev);
        V3 ex[4];
        for (int k = 0; k < 4; ++k) ex[k] = load_vertex(p.V, p.nV, ev[k]);
        double gd0[12], gd[12], eg[12];
        // distance derivatives first into A (scratch), then embedded into the edge stencil in Vv (scratch)
#define AE(i, j) A[((i) * 12 + (j)) * kHW + lane]
#define VE(i, j) Vv[((i) * 12 + (j)) * kHW + lane]
        const double d = pair_derivs(s, x, gd0, true, [&](int i, int j, double v) { AE(i, j) = v; });
        int map[4];
        for (int k = 0; k < s.nv; ++k) {
            map[k] = -1;
            for (int i = 0; i < 4; ++i)
                if (ev[i] == s.v[k]) map[k] = i;
        }
        for (int i = 0; i < 12; ++i) {
            gd[i] = 0.0;
            for (int j = 0; j < 12; ++j) VE(i, j) = 0.0;
        }
        for (int k = 0; k < s.nv; ++k) {
            if (map[k] < 0) continue;
            for (int i = 0; i < 3; ++i) gd[3 * map[k] + i] = gd0[3 * k + i];
            for (int l = 0; l < s.nv; ++l) {
                if (map[l] < 0) continue;
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) VE(3 * map[k] + i, 3 * map[l] + j) = AE(3 * k + i, 3 * l + j);
            }
        }
        double b, db, d2b;
        barrier_all(d, p.dHat, b, db, d2b);
        const double e = mollifier(ex, eps_x_rest(p.Vrest, p.nV, ev[0], ev[1], ev[2], ev[3]), eg, true, [&](int i, int j, double v) { HE(i, j) = v; });
        const double k = p.kappa;
        for (int i = 0; i < 12; ++i)
            for (int j = 0; j < 12; ++j)
                HE(i, j) = ((k * db) * gd[i]) * eg[j] + ((k * db) * gd[j]) * eg[i] + (k * b) * HE(i, j) + ((k * e * d2b) * gd[i]) * gd[j] + (k * e * db) * VE(i, j);
        for (int q = 0; q < 4; ++q) rows[q] = ev[q];
#undef AE
#undef VE
    }
    jacobi_psd(n, A, Vv, H, H, lane);
    // scatter (upper triangle only; projected Dirichlet rows/cols dropped)   [:533-556]
    const int nvb = n / 3;
    for (int i = 0; i < nvb; ++i) {
        if (proj_dbc(p.dbc, rows[i], p.projectDBC)) continue;
        for (int j = 0; j < nvb; ++j) {
            if (proj_dbc(p.dbc, rows[j], p.projectDBC)) continue;
            if (rows[i] > rows[j]) continue; // lower-triangular block: its transpose is added by (j,i)
            for (int r = 0; r < 3; ++r) {
                const int row = 3 * rows[i] + r;
                const int c0 = (rows[i] == rows[j]) ? r : 0;
                const int pos = csr_find(p.ia, p.ja, p.base, row, 3 * rows[j] + c0);
                if (pos < 0) {
                    atomicExch(err, 1);
                    continue;
                }
                for (int q = c0; q < 3; ++q) atomicAdd(a + pos + (q - c0), HE(3 * i + r, 3 * j + q));
            }
        }
    }
#undef HE
}

// -----------------------------------------------------------------------------------------------------------
void barrier_energy(const BarrierArgs& p, double* partials, int* bad, cudaStream_t st)
{
    const int n = p.nC + p.nP;
    if (n > 0) k_barrier_energy<<<(n + 255) / 256, 256, 0, st>>>(p, partials, bad);
}
int barrier_energy_blocks(int n) { return (n + 255) / 256; }
void barrier_gradient(const BarrierArgs& p, double* g, cudaStream_t st)
{
    const int n = p.nC + p.nP;
    if (n > 0) k_barrier_gradient<<<(n + 127) / 128, 128, 0, st>>>(p, g);
}
void barrier_hessian(const BarrierArgs& p, double* a, int* err, cudaStream_t st)
{
    const int n = p.nC + p.nP;
    if (n <= 0) return;
    const size_t smem = (size_t)3 * 144 * kHW * sizeof(double);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_barrier_hessian, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    k_barrier_hessian<<<(n + kHW - 1) / kHW, kHW, smem, st>>>(p, a, err);
}

} // namespace ipcgpu
