// pair_common.cuh -- MMCVID decoding, CSR lookup and Dirichlet helpers shared by the per-pair kernels (barrier.cu, friction.cu)
#pragma once
#include "contact.cuh"

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

} // namespace ipcgpu
