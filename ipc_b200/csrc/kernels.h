// kernels.h -- launcher declarations shared between the .cu translation units and the C-ABI (api.cu)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ipcgpu {

// Device-resident scalars of one Newton iteration.  Every stage reads its inputs from here and leaves its outputs here, so a whole
// iteration is one stream of launches (and in-stream NCCL reductions) with a single read-back at the end (ipcgpu_fetch_iteration).
struct IterState {
    unsigned long long step_ord;      // running step-size bound (order-preserving integer image of a non-negative double)
    unsigned long long inv_ord;       // min inversion step over this rank's tets (then over all ranks)
    unsigned long long ccd_ord;       // running minimum of the narrow phase in flight (every pair prunes against it)
    unsigned long long cand_range[2]; // [begin, end) of the candidate list this rank's narrow phase walks
    unsigned long long n_full_cand;   // candidates of the last full CCD (this rank)
    double max_t;                     // step on entry of the narrow phase in flight (max_t of every pair)
    double alpha_grid;                // sweep length of the last swept grid (after the span rescale of SpatialHash.hpp:603-618)
    double radius;                    // query inflation of the swept broad phase = one reference voxel
    double ref_lo[3], ref_inv_h;      // reference swept-grid geometry (SpatialHash.hpp:589-640)
    double alpha_stage[4];            // step after: inversion filter, partial CCD, swept-grid rescale, full CCD
    double energy[4];                 // elastic, barrier, friction, inertia (cross-rank sums once reduced)
    int ref_count[3];
    int n_set[3];                     // active / mollified / candidate counts of the last constraint set (this rank's lists)
    int flags[8];                     // IPCGPU_FLAG_* slots (nonzero = raised); cleared by ipcgpu_fetch_iteration
    int grid_axis_cells;              // cells per axis the broad-phase grids built since the last fetch would have liked (sort-width tuning)
    int pad_;
    int checks[2];                    // line-search safeguards: inverted tets, surface triangles crossed by an edge (this rank's share, then sums)
    unsigned long long ccd_stats[8];  // survivors, warnings, deferred, longest / total pair cycles, boxes (thread pass, warp pass), candidates
};
enum { FLAG_NONPOSITIVE_DISTANCE = 0, FLAG_SET_CAPACITY = 1, FLAG_CCD_CAPACITY = 2, FLAG_ZERO_CCD_DISTANCE = 3, FLAG_PATTERN = 4, FLAG_TI_WARNINGS = 5, FLAG_EXCHANGE_CAPACITY = 6 };

struct ElasticArgs {
    int nV, nT;
    int t_begin, t_end;      // this rank's OWNED tet range (energy, inversion filter: every tet exactly once)
    int n_list;              // gradient/Hessian kernel: number of tets this rank assembles (all tets that touch its rows)
    const int* tet_list;     // their ids, ascending; nullptr = the contiguous range [t_begin, t_begin + n_list)
    const double* V;         // SoA [x|y|z] current positions
    const int* T;            // SoA [v0|v1|v2|v3]
    const double* Ainv;      // SoA 9 x nT, q = 3*i+j row-major index of Dm^-1
    const double* vol;
    const double* mu;
    const double* lam;
    int energy;              // 0 NH, 1 FCR
    int e_row_lo, e_row_hi;  // fused energy of the gradient/Hessian kernel: a tet is counted by the rank that owns its smallest vertex
};

// elastic.cu
void elastic_energy(const ElasticArgs& p, double* e_per_tet, double* partials, double coef, double* out, cudaStream_t st);
int elastic_energy_blocks(int nTets);
// e_partials != nullptr: the kernel also leaves one partial sum of psi * vol per CTA there (elastic_grad_hess_blocks of them)
// hdst != nullptr: the Hessian blocks go to the SLOT-MAJOR intermediate hcon (destination offsets hdst, 10 per local tet) instead of the tile-major hblk
void elastic_grad_hess(const ElasticArgs& p, double coef, int projectSPD, bool need_g, bool need_h, double* gcont, double* hblk, cudaStream_t st, double* e_partials = nullptr,
    const unsigned* hdst = nullptr, double* hcon = nullptr);
void assemble_slot_major(int nSlots, const int* slot_v, const int* slot_u, const int* slot_off, const unsigned* cbase, const double* hcon, const uint8_t* dbc, int projectDBC,
    int accumulate, double* a, cudaStream_t st);
int elastic_grad_hess_blocks(int n_list);
void gather_gradient(int nV, const int* inc_ptr, const int* inc, const double* gcont, const uint8_t* dbc, int projectDBC, int accumulate, double* g, cudaStream_t st);
void assemble_csr(int nSlots, const int* slot_v, const int* slot_u, const int* slot_off, const int* con_ptr, const unsigned* con_src,
    const double* hblk, const uint8_t* dbc, int projectDBC, const double* mass, int accumulate, double* a, cudaStream_t st);
void diag_mass_dbc(int nV, const int* ia, int base, const uint8_t* dbc, int projectDBC, const double* mass, double* a, cudaStream_t st);
void slot_offsets(int nSlots, const int* slot_v, const int* slot_u, const int* ia, const int* ja, int base, int* slot_off, int* err, cudaStream_t st);
void diag_mass_dbc_range(int v0, int v1, const int* ia, int base, const uint8_t* dbc, int projectDBC, const double* mass, double* a, cudaStream_t st);
void inversion_step(const ElasticArgs& p, const double* dir, double slack, double* per_tet, IterState* st_dev, cudaStream_t st);
void inversion_apply(IterState* st_dev, int nT, cudaStream_t st);          // Energy.cpp:576-579 on the device-resident step
void step_set(IterState* st_dev, double alpha, cudaStream_t st);           // step_ord = alpha
void energy_store(IterState* st_dev, int slot, const double* src, cudaStream_t st);
void pack_scalars(const IterState* st_dev, unsigned local_mask, double* buf, cudaStream_t st);   // deferred cross-rank scalars -> 14 doubles
void unpack_scalars(IterState* st_dev, unsigned local_mask, const double* buf, cudaStream_t st);


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
    // kinematic obstacle (MeshCO): vertices >= nVdof belong to a triangle mesh without degrees of freedom that rides at the tail of the vertex
    // arrays (ipcgpu_set_obstacle_tail; INT_MAX = none).  Pairs between the mesh and the obstacle follow MeshCO.cpp, pairs inside the obstacle
    // do not exist.  ee_as_vf: Tight-Inclusion evaluates mesh-obstacle edge pairs through the vertex-face routine (MeshCO.cpp:1609)
    int nVdof, ee_as_vf;
};
// which body a surface primitive belongs to is decided by any one of its vertices (primitives do not straddle bodies)
__device__ __forceinline__ bool obstacle_vertex(const SurfArgs& s, int v) { return v >= s.nVdof; }

struct BarrierArgs {
    int nV;
    const double* V;
    const double* Vrest;
    const uint8_t* dbc;
    const int* SE;
    const int4* cs; const int* nC;          // active set (MMCVID encoding) and its DEVICE-resident count
    const int4* para; const int2* para_e; const int* nP; // mollified (nearly parallel EE) set + (eI,eJ), device count
    // multi-rank: E and g are taken over a contiguous share [n*rank/nranks, n*(rank+1)/nranks) of each list (share = 1) or over the
    // whole list (share = 0: the lists are already this rank's own); the Hessian is assembled by ROW OWNER: a rank processes every
    // pair that touches a vertex in [row_lo, row_hi) and scatters only the block rows it owns
    int rank, nranks, share;
    int row_lo, row_hi;
    double dHat, kappa;
    int projectDBC;
    const int* ia; const int* ja; int base;
    int nVdof; // first obstacle vertex (SurfArgs::nVdof)
};

// barrier.cu
void barrier_energy(const BarrierArgs& p, double* partials, int* bad, cudaStream_t st);
int barrier_energy_blocks();
void barrier_gradient(const BarrierArgs& p, double* g, cudaStream_t st);
void evaluate_constraints(const BarrierArgs& p, double* val, cudaStream_t st);
void constraint_jacobian_t(const BarrierArgs& p, const double* input, double coef, double* g, cudaStream_t st);
void para_gradient(const BarrierArgs& p, double* g, cudaStream_t st);
// Hraw: 144 doubles per owned pair; rows: 4 vertex ids per owned pair; psd: makePD "unchanged" flag per owned pair; n_owned: device counter
void barrier_hessian_build_project(const BarrierArgs& p, int* flags, double* Hraw, int* rows, int* psd, int* n_owned, int capacity, cudaStream_t st);
void barrier_hessian_scatter(const BarrierArgs& p, double* a, int* flags, const double* Hraw, const int* rows, const int* psd, const int* n_owned, int capacity, cudaStream_t st);
// friction.cu -- lagged friction of the self-contact pairs (SelfCollisionHandler.cpp:2481-2987)
struct FrictionArgs {
    int nV;
    const double* V;       // current positions (SoA)
    const double* Vt;      // positions at the start of the time step, result.V_prev (SoA)
    const uint8_t* dbc;
    const int4* cs; const int* n;   // LAGGED active set (MMActiveSet_lastH) and its device-resident size
    const double* lambda;           // MMLambda_lastH
    const double2* coord;           // MMDistCoord
    const double* basis;            // MMTanBasis: 6 per pair, column-major 3x2
    double eps2, coef;              // fricDHat, selfFric
    int projectDBC;
    const int* ia; const int* ja; int base;
    int rank, nranks;               // E and g: contiguous share of the list; H: by row owner [row_lo, row_hi)
    int row_lo, row_hi;
};
void friction_lag(const BarrierArgs& p, int4* cs_out, int* n_out, double* lambda, double2* coord, double* basis, int capacity, int* bad, cudaStream_t st);
void friction_energy(const FrictionArgs& p, double* partials, cudaStream_t st);
int friction_energy_blocks();
void friction_gradient(const FrictionArgs& p, double* g, cudaStream_t st);
void friction_hessian(const FrictionArgs& p, double* a, int* err, cudaStream_t st);
// elastic.cu (shared fixed-order reduction)
void reduce_sum(const double* partials, int n, double scale, double* out, cudaStream_t st);

// misc.cu
void step_forward(int nV, const double* x0_soa, const double* p_interleaved, double alpha, double* x_soa, cudaStream_t st);
// inertia term of Optimizer::computeEnergyVal / computeGradient (Optimizer.cpp:3227-3239, :3439-3450)
int inertia_energy_blocks(int nV);
void inertia_energy(int v0, int v1, int nV, const double* x_soa, const double* xtilde_soa, const double* mass, double* partials, cudaStream_t st);
void inertia_gradient(int nV, const double* x_soa, const double* xtilde_soa, const double* mass, const uint8_t* dbc, int projectDBC, double* g, cudaStream_t st);

} // namespace ipcgpu
