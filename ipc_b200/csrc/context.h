// context.h -- the ipcgpu_ctx object behind the C ABI: device buffers, scatter maps, streams, NCCL.
#pragma once
#include "kernels.h"
#include "broadphase_types.h"
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace ipcgpu {

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    // grow-only allocation; returns false on CUDA failure
    bool reserve(size_t count)
    {
        if (count <= n) return true;
        release();
        if (cudaMalloc(&p, count * sizeof(T)) != cudaSuccess) {
            p = nullptr;
            return false;
        }
        n = count;
        return true;
    }
    bool upload(const T* h, size_t count, cudaStream_t st)
    {
        if (!reserve(count)) return false;
        return count == 0 || cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st) == cudaSuccess;
    }
};

// device workspace of the contact stages (constraint.cu / ccd.cu)
struct ContactWork {
    DevBuf<Box> vbox, ebox, tbox, tsbox, esbox; // per-primitive boxes; t/e boxes again in grid-sorted order
    DevBuf<unsigned long long> bounds, tkeys, ekeys, key_tmp, skey, skey2;
    DevBuf<Grid> grid;
    DevBuf<int> tvals, evals, val_tmp, counters, sidx, sidx2;
    DevBuf<int4> act, dup, para, tmp4;
    DevBuf<int2> para_e, cand, tmp2;
    DevBuf<unsigned char> cub_tmp;
    DevBuf<unsigned> ttab_key, etab_key; // cell hash tables of the triangle / edge grids
    DevBuf<int2> ttab_start, etab_start; // [first, last+1) entry range per table slot
    unsigned tab_mask = 0;
    DevBuf<int2> bp_pairs; // broad-phase pair lists (PT then EE), bp_cap each
    size_t bp_cap = 0;
    int cap = 0;
    int nC = 0, nP = 0, nK = 0; // current active / mollified / candidate counts
};

// device workspace of the CCD stage (ccd.cu)
struct CcdWork {
    DevBuf<int> vmin, vmax, counters;
    DevBuf<int2> cand;
    DevBuf<unsigned> surv, surv2;
    bool wide_level_set = false;
    DevBuf<unsigned char> scratch;
    DevBuf<unsigned long long> ncand, bounds;
    // reference swept-grid geometry (SpatialHash.hpp:589-640) of the last ipcgpu_hash_build_swept
    double ref_lo[3] = { 0, 0, 0 }, ref_inv_h = 0.0, alpha_grid = 0.0;
    int ref_count[3] = { 0, 0, 0 };
    bool swept_ready = false;
    unsigned last_survivors = 0;
    unsigned long long last_deferred = 0, last_longest_cycles = 0, last_total_cycles = 0;
    int last_warnings = 0;
    unsigned long long last_candidates = 0, last_boxes_thread = 0, last_boxes_warp = 0;
};

} // namespace ipcgpu

struct ipcgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;

    // partition
    int rank = 0, nranks = 1;
    void* nccl_comm = nullptr;

    // mesh
    int nV = 0, nT = 0, energy = 0;
    int t_begin = 0, t_end = 0;
    ipcgpu::DevBuf<double> V, Vsaved, Vrest, Ainv, vol, mu, lam, mass;
    ipcgpu::DevBuf<int> T;
    ipcgpu::DevBuf<uint8_t> dbc;
    bool has_mass = false, has_dbc = false, state_saved = false;
    std::vector<int> h_T; // host copy of tets (maps are rebuilt when the partition changes)

    // surface (Mesh::SVI / SFEdges / SF) and contact workspace
    int nSV = 0, nSE = 0, nSF = 0;
    ipcgpu::DevBuf<int> SVI, SE, SF, vCoDim;
    bool has_codim = false, surface_ready = false;
    int pair_capacity = 1 << 20;
    bool canonical_order = true; // sort the contact lists lexicographically after the build
    bool partition_contact = false, lists_local = false; // multi-rank: build only this rank's share of the contact sets
    ipcgpu::ContactWork cw;
    ipcgpu::CcdWork ccd;
    size_t ccd_capacity = (size_t)1 << 23; // candidate pairs
    std::vector<int> h_SVI;                // host copy (pSize of the swept build is a serial host sum, SpatialHash.hpp:603-612)
    std::vector<double> h_dir;             // host shadow of the last uploaded search direction
    ipcgpu::DevBuf<double> bpartials, bHraw;
    ipcgpu::DevBuf<int> brows;

    // gradient gather map (local tets)
    ipcgpu::DevBuf<int> inc_ptr, inc;
    // Hessian slots (mesh-topology vertex pairs v<=u touched by local tets) and contributions
    int nSlots = 0;
    ipcgpu::DevBuf<int> slot_v, slot_u, slot_off, con_ptr;
    ipcgpu::DevBuf<unsigned> con_src;
    bool maps_ready = false, offsets_ready = false;

    // CSR
    int n_rows = 0, nnz = 0, index_base = 0;
    ipcgpu::DevBuf<int> ia, ja;
    ipcgpu::DevBuf<double> a;
    ipcgpu::DevBuf<int> flag; // device error flag

    // work / result buffers
    ipcgpu::DevBuf<double> gcont, hblk, g, e_per_tet, partials, scalar_out, inv_steps, dir;
    ipcgpu::DevBuf<unsigned long long> min_ord;
    double* h_scalar = nullptr; // pinned staging for scalars (4 doubles)

    // profiling: event pairs per stage (only when enabled)
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof[16];
    cudaEvent_t timer_a = nullptr, timer_b = nullptr;
    cudaEvent_t prof_begin(int stage)
    {
        if (!profiling) return nullptr;
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        prof[stage].emplace_back(a, b);
        cudaEventRecord(a, stream);
        return b;
    }
    void prof_end(cudaEvent_t b)
    {
        if (b) cudaEventRecord(b, stream);
    }

    ipcgpu::ElasticArgs eargs() const
    {
        ipcgpu::ElasticArgs p;
        p.nV = nV; p.nT = nT; p.t_begin = t_begin; p.t_end = t_end;
        p.V = V.p; p.T = T.p; p.Ainv = Ainv.p; p.vol = vol.p; p.mu = mu.p; p.lam = lam.p; p.energy = energy;
        return p;
    }
};
