// kernels.h -- launcher declarations shared between the .cu translation units and the C-ABI (api.cu)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ipcgpu {

struct ElasticArgs {
    int nV, nT;
    int t_begin, t_end;      // this rank's tet range (multi-GPU partition)
    const double* V;         // SoA [x|y|z] current positions
    const int* T;            // SoA [v0|v1|v2|v3]
    const double* Ainv;      // SoA 9 x nT, q = 3*i+j row-major index of Dm^-1
    const double* vol;
    const double* mu;
    const double* lam;
    int energy;              // 0 NH, 1 FCR
};

// elastic.cu
void elastic_energy(const ElasticArgs& p, double* e_per_tet, double* partials, double coef, double* out, cudaStream_t st);
int elastic_energy_blocks(int nTets);
void elastic_grad_hess(const ElasticArgs& p, double coef, int projectSPD, bool need_g, bool need_h, double* gcont, double* hblk, cudaStream_t st);
void gather_gradient(int nV, const int* inc_ptr, const int* inc, const double* gcont, const uint8_t* dbc, int projectDBC, int accumulate, double* g, cudaStream_t st);
void assemble_csr(int nSlots, const int* slot_v, const int* slot_u, const int* slot_off, const int* con_ptr, const unsigned* con_src,
    const double* hblk, const uint8_t* dbc, int projectDBC, const double* mass, int accumulate, double* a, cudaStream_t st);
void diag_mass_dbc(int nV, const int* ia, int base, const uint8_t* dbc, int projectDBC, const double* mass, double* a, cudaStream_t st);
void slot_offsets(int nSlots, const int* slot_v, const int* slot_u, const int* ia, const int* ja, int base, int* slot_off, int* err, cudaStream_t st);
void inversion_step(const ElasticArgs& p, const double* dir, double slack, double* per_tet, unsigned long long* min_ord, cudaStream_t st);


// ---- contact ------------------------------------------------------------------------------------------
struct SurfArgs {
    int nV;
    const double* V;      // SoA current positions
    const double* Vrest;  // SoA rest positions
    const uint8_t* dbc;   // nullable
    const int* vCoDim;    // nullable (=> 3)
    int nSV; const int* SVI;
    int nSE; const int* SE;   // interleaved (first, second)   [Mesh::SFEdges]
    int nSF; const int* SF;   // SoA [v0|v1|v2]                [Mesh::SF column-major]
};

struct BarrierArgs {
    int nV;
    const double* V;
    const double* Vrest;
    const uint8_t* dbc;
    const int* SE;
    const int4* cs; int nC;          // active set (MMCVID encoding)
    const int4* para; const int2* para_e; int nP; // mollified (nearly parallel EE) set + (eI,eJ)
    double dHat, kappa;
    int projectDBC;
    const int* ia; const int* ja; int base;
};

// barrier.cu
void barrier_energy(const BarrierArgs& p, double* partials, int* bad, cudaStream_t st);
int barrier_energy_blocks(int n);
void barrier_gradient(const BarrierArgs& p, double* g, cudaStream_t st);
void barrier_hessian(const BarrierArgs& p, double* a, int* err, double* Hraw /* 144 per pair */, int* rows /* 5 per pair: 4 vertex ids, then n flags */, cudaStream_t st);
// elastic.cu (shared fixed-order reduction)
void reduce_sum(const double* partials, int n, double scale, double* out, cudaStream_t st);

// misc.cu
void step_forward(int nV, const double* x0_soa, const double* p_interleaved, double alpha, double* x_soa, cudaStream_t st);

} // namespace ipcgpu
