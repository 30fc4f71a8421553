// common.cuh -- shared helpers for the ipc_b200 CUDA hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>


#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__

namespace ipcgpu {

constexpr int kSMs = 148; // B200: 2 dies x 74 SMs

// 3x3 row-major register matrix
struct M3 {
    double m[9];
    DEV double& operator()(int i, int j) { return m[3 * i + j]; }
    DEV double operator()(int i, int j) const { return m[3 * i + j]; }
};

DEV double det3(const M3& A)
{
    return A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1))
        - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0))
        + A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
}

// order-preserving map of a non-negative double onto uint64 (for atomicMin on step sizes)
DEV unsigned long long dbl_to_ord(double x) { return (unsigned long long)__double_as_longlong(x); }
DEV double ord_to_dbl(unsigned long long u) { return __longlong_as_double((long long)u); }

template <typename T>
DEV T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
DEV double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

} // namespace ipcgpu
