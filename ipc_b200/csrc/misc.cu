// misc.cu -- small per-vertex kernels the Newton driver runs between the hot stages.
#include "common.cuh"
#include "kernels.h"

namespace ipcgpu {

// Optimizer::stepForward (src/TimeStepper/Optimizer.cpp:2919-2938): x = x0 + alpha * p
__global__ void __launch_bounds__(256) k_step_forward(int nV, const double* __restrict__ x0, const double* __restrict__ p, double alpha, double* __restrict__ x)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) x[(size_t)c * nV + v] = x0[(size_t)c * nV + v] + alpha * p[3 * (size_t)v + c];
}

void step_forward(int nV, const double* x0, const double* p, double alpha, double* x, cudaStream_t st)
{
    if (nV > 0) k_step_forward<<<(nV + 255) / 256, 256, 0, st>>>(nV, x0, p, alpha, x);
}

} // namespace ipcgpu
