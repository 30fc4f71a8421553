// solve.cu -- device-resident linear-solve hand-off: the Hessian never leaves HBM.
//
// Reference being stood in for: LinSysSolver::{factorize, solve} (src/LinSysSolver/LinSysSolver.hpp:230-236) as used by
// Optimizer::computeSearchDir (src/TimeStepper/Optimizer.cpp:2324-2355), i.e. CHOLMODSolver.cpp:123-154.  The north star keeps the sparse
// Cholesky a black box (CHOLMOD / cuDSS); cuDSS is not in this image, so the production binding is documented in INTEGRATION.md
// (ipcgpu_device_ptr hands cuDSS the device-resident ia / ja / a) and what is BUILT here is the hand-off itself plus a reference solver
// that runs entirely on the device: a block-Jacobi preconditioned conjugate gradient on the upper-triangular CSR the assembly stages fill.
// It takes its right-hand side from the device-resident gradient and leaves the search direction where the step-bound stages read it,
// so that a whole Newton iteration (assembly -> solve -> CCD) needs no host transfer of any vertex- or matrix-sized array.
//
// SpMV on a symmetric matrix stored by its upper triangle: at ipcgpu_set_csr time the host builds the FULL row structure once
// (col index + position of the value inside the upper-triangular array for every entry of both triangles), so the product is a plain
// deterministic row-parallel CSR SpMV that gathers a[] through that position map -- no atomics, no transposed pass.
#include "common.cuh"
#include "context.h"
#include "../../include/ipcgpu.h"
#include <algorithm>
#include <cmath>
#include <vector>

namespace ipcgpu {

// y = A x over full rows; dot(x, y) accumulated into scal[slot] (one warp per row)
__global__ void __launch_bounds__(256) k_spmv_dot(int n, const int* __restrict__ fia, const int* __restrict__ fja, const int* __restrict__ fpos, const double* __restrict__ a,
    const double* __restrict__ x, double* __restrict__ y, double* __restrict__ scal, int slot)
{
    const int lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double acc = 0.0;
    for (int row = row0; row < n; row += (gridDim.x * blockDim.x) >> 5) {
        double s = 0.0;
        for (int k = fia[row] + lane; k < fia[row + 1]; k += 32) s += __ldg(a + fpos[k]) * __ldg(x + fja[k]);
        s = warp_sum(s);
        if (lane == 0) {
            y[row] = s;
            acc += x[row] * s;
        }
    }
    if (lane == 0 && acc != 0.0) atomicAdd(scal + slot, acc);
}

// block-Jacobi preconditioner: inverse of the 3x3 diagonal block of every vertex (row 3v: [d00 d01 d02], row 3v+1: [d11 d12], row 3v+2: [d22])
__global__ void __launch_bounds__(256) k_block_jacobi(int nV, const int* __restrict__ ia, int base, const double* __restrict__ a, double* __restrict__ Minv /* 6 per vertex */)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const int o0 = ia[3 * v] - base, o1 = ia[3 * v + 1] - base, o2 = ia[3 * v + 2] - base;
    const double d00 = a[o0], d01 = a[o0 + 1], d02 = a[o0 + 2], d11 = a[o1], d12 = a[o1 + 1], d22 = a[o2];
    const double c00 = d11 * d22 - d12 * d12, c01 = d02 * d12 - d01 * d22, c02 = d01 * d12 - d02 * d11;
    const double det = d00 * c00 + d01 * c01 + d02 * c02;
    double* m = Minv + 6 * (size_t)v;
    if (!(fabs(det) > 0.0)) { // singular block (should not happen for an SPD matrix): fall back to the scalar diagonal
        m[0] = d00 != 0.0 ? 1.0 / d00 : 1.0; m[3] = d11 != 0.0 ? 1.0 / d11 : 1.0; m[5] = d22 != 0.0 ? 1.0 / d22 : 1.0;
        m[1] = m[2] = m[4] = 0.0;
        return;
    }
    const double id = 1.0 / det;
    m[0] = c00 * id; m[1] = c01 * id; m[2] = c02 * id;
    m[3] = (d00 * d22 - d02 * d02) * id; m[4] = (d01 * d02 - d00 * d12) * id;
    m[5] = (d00 * d11 - d01 * d01) * id;
}

// scal: [0] rz, [1] pAp, [2] rz_new, [3] |r|^2, [4] |b|^2
// init: r = b (= sign * src), x = 0, z = Minv r, p = z, rz = r.z
__global__ void __launch_bounds__(256) k_pcg_init(int nV, const double* __restrict__ src, double sign, const double* __restrict__ Minv, double* __restrict__ x, double* __restrict__ r,
    double* __restrict__ p, double* __restrict__ scal)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    double rz = 0.0, bb = 0.0;
    if (v < nV) {
        const double r0 = sign * src[3 * (size_t)v], r1 = sign * src[3 * (size_t)v + 1], r2 = sign * src[3 * (size_t)v + 2];
        const double* m = Minv + 6 * (size_t)v;
        const double z0 = m[0] * r0 + m[1] * r1 + m[2] * r2, z1 = m[1] * r0 + m[3] * r1 + m[4] * r2, z2 = m[2] * r0 + m[4] * r1 + m[5] * r2;
        for (int c = 0; c < 3; ++c) x[3 * (size_t)v + c] = 0.0;
        r[3 * (size_t)v] = r0; r[3 * (size_t)v + 1] = r1; r[3 * (size_t)v + 2] = r2;
        p[3 * (size_t)v] = z0; p[3 * (size_t)v + 1] = z1; p[3 * (size_t)v + 2] = z2;
        rz = r0 * z0 + r1 * z1 + r2 * z2;
        bb = r0 * r0 + r1 * r1 + r2 * r2;
    }
    rz = warp_sum(rz);
    bb = warp_sum(bb);
    if ((threadIdx.x & 31) == 0) {
        if (rz != 0.0) atomicAdd(scal + 0, rz);
        if (bb != 0.0) atomicAdd(scal + 4, bb);
    }
}
// x += alpha p ; r -= alpha Ap ; z = Minv r (kept in Ap's storage) ; rz_new = r.z ; |r|^2       with alpha = rz / pAp
__global__ void __launch_bounds__(256) k_pcg_update(int nV, const double* __restrict__ Minv, const double* __restrict__ p, double* __restrict__ Ap_z, double* __restrict__ x,
    double* __restrict__ r, double* __restrict__ scal)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const double pAp = scal[1];
    const double alpha = pAp != 0.0 ? scal[0] / pAp : 0.0;
    double rz = 0.0, rr = 0.0;
    if (v < nV) {
        double rv[3];
        for (int c = 0; c < 3; ++c) {
            const size_t i = 3 * (size_t)v + c;
            x[i] += alpha * p[i];
            rv[c] = r[i] - alpha * Ap_z[i];
            r[i] = rv[c];
        }
        const double* m = Minv + 6 * (size_t)v;
        const double z0 = m[0] * rv[0] + m[1] * rv[1] + m[2] * rv[2], z1 = m[1] * rv[0] + m[3] * rv[1] + m[4] * rv[2], z2 = m[2] * rv[0] + m[4] * rv[1] + m[5] * rv[2];
        Ap_z[3 * (size_t)v] = z0; Ap_z[3 * (size_t)v + 1] = z1; Ap_z[3 * (size_t)v + 2] = z2;
        rz = rv[0] * z0 + rv[1] * z1 + rv[2] * z2;
        rr = rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2];
    }
    rz = warp_sum(rz);
    rr = warp_sum(rr);
    if ((threadIdx.x & 31) == 0) {
        if (rz != 0.0) atomicAdd(scal + 2, rz);
        if (rr != 0.0) atomicAdd(scal + 3, rr);
    }
}
// p = z + beta p with beta = rz_new / rz ; then roll the scalars for the next iteration (one thread)
__global__ void __launch_bounds__(256) k_pcg_direction(int n, const double* __restrict__ z, double* __restrict__ p, const double* __restrict__ scal)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double rz = scal[0];
    const double beta = rz != 0.0 ? scal[2] / rz : 0.0;
    if (i < n) p[i] = z[i] + beta * p[i];
}
__global__ void k_pcg_roll(double* scal, double* history, int it)
{
    if (threadIdx.x == 0) {
        history[it] = scal[3]; // |r|^2 after this iteration
        scal[0] = scal[2];
        scal[1] = scal[2] = scal[3] = 0.0;
    }
}
// mean |p| over the surface vertices (SpatialHash.hpp:603-612) for a direction that was produced on the device: fixed-order two-level sum
__global__ void __launch_bounds__(256) k_psize(int nSV, const int* __restrict__ SVI, int nVdof, const double* __restrict__ p, double* __restrict__ partials)
{
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nSV; i += gridDim.x * blockDim.x) {
        const int v = SVI[i];
        if (v >= nVdof) continue; // the obstacle's surface vertices do not count (SpatialHash::build sees the mesh alone)
        s += fabs(p[3 * (size_t)v]) + fabs(p[3 * (size_t)v + 1]) + fabs(p[3 * (size_t)v + 2]);
    }
    s = warp_sum(s);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[w];
        partials[blockIdx.x] = t;
    }
}

} // namespace ipcgpu

using namespace ipcgpu;

#define CKS(call)                                                      \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) {                                       \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); \
            return IPCGPU_ERR_CUDA;                                    \
        }                                                              \
    } while (0)

static inline int nblk(long long n, int b) { return (int)((n + b - 1) / b); }

// full row structure of the symmetric matrix from its upper-triangular CSR (host, once per pattern)
int solver_build_full_pattern(ipcgpu_ctx* ctx, const int* ia, const int* ja)
{
    const int n = ctx->n_rows, base = ctx->index_base;
    std::vector<int> cnt((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i)
        for (int k = ia[i] - base; k < ia[i + 1] - base; ++k) {
            const int j = ja[k] - base;
            ++cnt[i + 1];
            if (j != i) ++cnt[j + 1];
        }
    for (int i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
    const size_t nf = (size_t)cnt[n];
    std::vector<int> fja(nf), fpos(nf), cur(cnt.begin(), cnt.end() - 1);
    // lower part first per row (transposed entries arrive in ascending source row = ascending column), then the row's own upper entries
    for (int i = 0; i < n; ++i)
        for (int k = ia[i] - base; k < ia[i + 1] - base; ++k) {
            const int j = ja[k] - base;
            if (j != i) { fja[cur[j]] = i; fpos[cur[j]] = k; ++cur[j]; }
        }
    for (int i = 0; i < n; ++i)
        for (int k = ia[i] - base; k < ia[i + 1] - base; ++k) {
            fja[cur[i]] = ja[k] - base; fpos[cur[i]] = k; ++cur[i];
        }
    bool ok = ctx->fia.upload(cnt.data(), cnt.size(), ctx->stream) && ctx->fja.upload(fja.data(), nf, ctx->stream) && ctx->fpos.upload(fpos.data(), nf, ctx->stream);
    if (!ok) {
        ctx->err = "upload of the full-row pattern failed";
        return IPCGPU_ERR_CUDA;
    }
    CKS(cudaStreamSynchronize(ctx->stream));
    ctx->full_pattern_ready = true;
    return 0;
}

// PCG on the device-resident matrix.  rhs_dev: device vector (3 nV) scaled by `sign`.  The solution is left in ctx->sol.
int solver_pcg(ipcgpu_ctx* ctx, const double* rhs_dev, double sign, double rel_tol, int max_iter, int* iters_out, double* rel_res_out)
{
    cudaStream_t st = ctx->stream;
    const int n = ctx->n_rows, nV = ctx->nV;
    bool ok = ctx->sol.reserve(n) && ctx->pcg_r.reserve(n) && ctx->pcg_p.reserve(n) && ctx->pcg_q.reserve(n) && ctx->pcg_minv.reserve((size_t)6 * nV) && ctx->pcg_scal.reserve(8)
        && ctx->pcg_hist.reserve((size_t)std::max(max_iter, 1) + 1);
    if (!ok) {
        ctx->err = "PCG workspace allocation failed";
        return IPCGPU_ERR_CUDA;
    }
    CKS(cudaMemsetAsync(ctx->pcg_scal.p, 0, 8 * sizeof(double), st));
    k_block_jacobi<<<nblk(nV, 256), 256, 0, st>>>(nV, ctx->ia.p, ctx->index_base, ctx->a.p, ctx->pcg_minv.p);
    k_pcg_init<<<nblk(nV, 256), 256, 0, st>>>(nV, rhs_dev, sign, ctx->pcg_minv.p, ctx->sol.p, ctx->pcg_r.p, ctx->pcg_p.p, ctx->pcg_scal.p);
    ctx->launches += 2;
    double* h = ctx->h_scalar;
    CKS(cudaMemcpyAsync(h, ctx->pcg_scal.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CKS(cudaStreamSynchronize(st));
    const double bb = h[4];
    int it = 0;
    double rr = bb;
    const int check_every = 25; // the host looks at the residual every 25 iterations (one small read-back)
    if (bb > 0.0) {
        while (it < max_iter) {
            const int burst = std::min(check_every, max_iter - it);
            for (int b = 0; b < burst; ++b, ++it) {
                k_spmv_dot<<<kSMs * 8, 256, 0, st>>>(n, ctx->fia.p, ctx->fja.p, ctx->fpos.p, ctx->a.p, ctx->pcg_p.p, ctx->pcg_q.p, ctx->pcg_scal.p, 1);
                k_pcg_update<<<nblk(nV, 256), 256, 0, st>>>(nV, ctx->pcg_minv.p, ctx->pcg_p.p, ctx->pcg_q.p, ctx->sol.p, ctx->pcg_r.p, ctx->pcg_scal.p);
                k_pcg_direction<<<nblk(n, 256), 256, 0, st>>>(n, ctx->pcg_q.p, ctx->pcg_p.p, ctx->pcg_scal.p);
                k_pcg_roll<<<1, 32, 0, st>>>(ctx->pcg_scal.p, ctx->pcg_hist.p, it);
                ctx->launches += 4;
            }
            CKS(cudaMemcpyAsync(h, ctx->pcg_hist.p + (it - 1), sizeof(double), cudaMemcpyDeviceToHost, st));
            CKS(cudaStreamSynchronize(st));
            rr = h[0];
            if (!(rr == rr)) break; // NaN: the matrix was not positive definite
            if (std::sqrt(rr) <= rel_tol * std::sqrt(bb)) break;
        }
    }
    CKS(cudaGetLastError());
    if (iters_out) *iters_out = it;
    if (rel_res_out) *rel_res_out = bb > 0.0 ? std::sqrt(rr / bb) : 0.0;
    return 0;
}

// the device-resident solution becomes the search direction of the step-bound stages (pSize by a fixed-order device sum)
int solver_adopt_direction(ipcgpu_ctx* ctx)
{
    cudaStream_t st = ctx->stream;
    CKS(cudaMemcpyAsync(ctx->dir.p, ctx->sol.p, (size_t)ctx->n_rows * sizeof(double), cudaMemcpyDeviceToDevice, st));
    double pSize = 0.0;
    if (ctx->nSV > 0) {
        const int nb = 64;
        if (!ctx->pcg_scal.reserve(8) || !ctx->partials.reserve(nb + 8)) return IPCGPU_ERR_CUDA;
        k_psize<<<nb, 256, 0, st>>>(ctx->nSV, ctx->SVI.p, ctx->nVdof, ctx->dir.p, ctx->partials.p);
        ++ctx->launches;
        std::vector<double> hp(nb);
        CKS(cudaMemcpyAsync(hp.data(), ctx->partials.p, nb * sizeof(double), cudaMemcpyDeviceToHost, st));
        CKS(cudaStreamSynchronize(st));
        for (double v : hp) pSize += v;
        long long nMeshSV = 0;
        for (int v : ctx->h_SVI) nMeshSV += v < ctx->nVdof ? 1 : 0;
        pSize = nMeshSV > 0 ? pSize / (double)(nMeshSV * 3) : 0.0;
    }
    ctx->pSize = pSize;
    if (!ctx->pSize_dev.reserve(1)) return IPCGPU_ERR_CUDA;
    CKS(cudaMemcpyAsync(ctx->pSize_dev.p, &ctx->pSize, sizeof(double), cudaMemcpyHostToDevice, st));
    CKS(cudaStreamSynchronize(st));
    ctx->pSize_surface = ctx->surface_ready;
    ctx->dir_valid = true;
    return 0;
}
