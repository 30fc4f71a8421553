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
    DevBuf<Box> vbox, ebox, tbox;   // per-primitive boxes (primitive order)
    DevBuf<QEntry> centries;        // grid-sorted entries (triangles first, then edges): quantised box + id
    DevBuf<unsigned long long> bounds, skey, skey2;
    DevBuf<unsigned> ckeys, key_tmp; // ckeys: sorted 32-bit cell keys (cell | type bits) of the combined triangle + edge + vertex grid
    DevBuf<Grid> grid;
    DevBuf<int> cvals, val_tmp, counters, sidx, sidx2;
    DevBuf<int4> act, dup, para, tmp4;
    DevBuf<int2> para_e, cand, tmp2;
    DevBuf<unsigned char> cub_tmp;
    DevBuf<int> cell_cnt, cell_off; // dense per-(type, cell) counters and their exclusive prefix sum: entry range of a cell = two adjacent offsets
    DevBuf<int2> bp_pairs; // broad-phase pair lists (PT then EE), bp_cap each
    size_t bp_cap = 0;
    int cap = 0;
    // counters (device ints): [0] active, [1] PP/PE duplicates, [2] mollified, [3] candidates, [4] overflow, [8]/[9] broad-phase pairs PT/EE,
    // [10]/[11] active / mollified of the GLOBAL lists after the cross-rank exchange, [12] pair Hessians owned by this rank
    int nC = 0, nP = 0, nK = 0; // host mirrors of [0], [2], [3]; -1 = not read back (device-resident iteration)
    bool want_cand = false;
    unsigned dup_tab = 1024;    // slots of the PP/PE duplicate-merge table
    int axis_bits = 10;         // key bits per axis of the grid sorts (re-tuned from IterState::grid_axis_cells at every fetch)
    int built_axis_bits = 10;   // ... of the grids that are currently built (position of the type bit)
    int built_vertices = 0;     // surface-vertex entries in the combined sorted array (0: the build carried none)
    bool built_voxel_entries = false; // the entries of the current grid carry reference-voxel ranges (swept grid) instead of quantised boxes
    // barrier stage workspace, sized by the pair capacity
    DevBuf<double> bHraw, bpartials, bval; // (bval: per-constraint values of ipcgpu_evaluate_constraints / inputs of ..._jacobian_t)
    DevBuf<int> brows, bpsd;
    // multi-rank exchange of the pair lists (one fixed-size message per rank, see k_pack_lists)
    int xcap = 0;
    size_t xstride = 0;
    DevBuf<int4> xsend, xrecv, gact, gpara;
    DevBuf<int2> gpara_e;
    bool lists_global = false; // gact / gpara hold the global lists of the last partitioned build
    // lagged friction data (MMActiveSet_lastH, MMLambda_lastH, MMDistCoord, MMTanBasis): snapshot taken by ipcgpu_friction_lag
    DevBuf<int4> fr_cs;
    DevBuf<int> fr_n;
    DevBuf<double> fr_lambda, fr_basis, fr_partials;
    DevBuf<double2> fr_coord;
    int fr_host_n = -1;        // host mirror of the lagged count (-1 = not read back)
    bool fr_ready = false;
};

// device workspace of the CCD stage (ccd.cu)
struct CcdWork {
    DevBuf<int> vmin, vmax, counters;
    DevBuf<int2> cand;
    DevBuf<unsigned> surv, surv2;
    bool wide_level_set = false;
    DevBuf<unsigned char> scratch;
    DevBuf<unsigned long long> ncand, bounds;
    bool swept_ready = false; // (the reference swept-grid geometry of the last build lives in IterState)
    unsigned last_survivors = 0;
    unsigned long long last_deferred = 0, last_longest_cycles = 0, last_total_cycles = 0;
    int last_warnings = 0;
    unsigned long long last_candidates = 0, last_boxes_thread = 0, last_boxes_warp = 0;
};

} // namespace ipcgpu

struct ipcgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // side stream: the build + projection of the pair Hessians (latency-bound, a fraction of one wave) run next to the elastic assembly;
    // ev_inputs marks, on the main stream, the last point at which their inputs (positions, contact sets) changed
    cudaStream_t side = nullptr;
    // copy stream: ipcgpu_download_range_async forks it off the main stream, so that a result (gradient, CSR values) travels to the host
    // while the later stages of the iteration run; joined by ipcgpu_fetch_iteration / ipcgpu_sync / ipcgpu_capture_end
    cudaStream_t copy = nullptr;
    cudaEvent_t ev_copy_fork = nullptr, ev_copy_join = nullptr;
    bool copy_pending = false;
    cudaEvent_t ev_inputs = nullptr, ev_join = nullptr, ev_scatter = nullptr; // (ev_scatter: the previous scatter has read the pair Hessians)
    bool scatter_marked = false;
    bool inputs_marked = false;
    void mark_inputs()
    {
        if (ev_inputs) inputs_marked = (cudaEventRecord(ev_inputs, stream) == cudaSuccess);
    }
    std::string err;
    uint64_t launches = 0;

    // partition: tets [t_begin, t_end) are OWNED (energy, inversion filter); rows of vertices [v_begin, v_end) are owned, and the
    // gradient/Hessian kernel runs over every tet that touches them (tet_list, halo tets duplicated) so that the CSR needs no reduction
    int rank = 0, nranks = 1;
    void* nccl_comm = nullptr;
    int v_begin = 0, v_end = 0, n_list = 0;
    ipcgpu::DevBuf<int> tet_list;
    long long a_begin = 0, a_end = 0; // CSR value range of the owned rows

    // device-resident scalars of the iteration in flight + pinned host mirror
    ipcgpu::DevBuf<ipcgpu::IterState> iter;
    ipcgpu::IterState* h_iter = nullptr;
    double pSize = 0.0;     // mean |p| over the surface vertices (SpatialHash.hpp:603-612), computed when p is uploaded
    bool dir_valid = false, pSize_surface = false;
    bool energy_local[4] = { false, false, false, false }; // IterState::energy[s] still holds this rank's partial sum
    bool a_all_dirty = false;                // a cross-rank completion filled rows this rank does not own
    bool checks_local = false;               // IterState::checks still holds this rank's partial counts
    std::vector<int> h_ia;                   // host copy of the CSR row starts (value range of the owned rows)

    // mesh
    int nV = 0, nT = 0, energy = 0;
    int t_begin = 0, t_end = 0;
    ipcgpu::DevBuf<double> V, Vsaved, Vrest, Ainv, vol, mu, lam, mass, Vprev, xtilde;
    bool prev_set = false, xtilde_set = false; // result.V_prev (friction) / xTilta (inertia) uploaded
    ipcgpu::DevBuf<int> T;
    ipcgpu::DevBuf<uint8_t> dbc;
    bool has_mass = false, has_dbc = false, state_saved = false;
    std::vector<int> h_T; // host copy of tets (maps are rebuilt when the partition changes)

    // surface (Mesh::SVI / SFEdges / SF) and contact workspace
    int nSV = 0, nSE = 0, nSF = 0;
    int nVdof = 0x7fffffff;  // first obstacle vertex (ipcgpu_set_obstacle_tail): vertices from here on have no degrees of freedom; INT_MAX = no obstacle
    int ee_as_vf = 1;        // Tight-Inclusion of mesh-obstacle edge pairs through the vertex-face routine, as MeshCO.cpp:1609 calls it
    ipcgpu::DevBuf<int> SVI, SE, SF, vCoDim;
    bool has_codim = false, surface_ready = false;
    int pair_capacity = 1 << 20;
    int exchange_capacity = 1 << 16; // pairs per rank and list in the fixed-size message of the cross-rank pair-list exchange
    bool canonical_order = true; // sort the contact lists lexicographically after the build
    bool partition_contact = false, lists_local = false; // multi-rank: build only this rank's share of the contact sets
    ipcgpu::ContactWork cw;
    ipcgpu::CcdWork ccd;
    size_t ccd_capacity = (size_t)1 << 23; // candidate pairs
    std::vector<int> h_SVI;                // host copy (pSize of the swept build is a serial host sum, SpatialHash.hpp:603-612)
    double debug_prune_seed = -1.0;        // test hook, see ipcgpu_ccd_debug_seed_bound

    // gradient gather map (local tets)
    ipcgpu::DevBuf<int> inc_ptr, inc;
    // Hessian slots (mesh-topology vertex pairs v<=u touched by local tets) and contributions
    int nSlots = 0;
    ipcgpu::DevBuf<int> slot_v, slot_u, slot_off, con_ptr;
    ipcgpu::DevBuf<unsigned> con_src, hdst, cbase; // hdst / cbase: slot-major intermediate (destination of each tet block / start of each slot's run)
    ipcgpu::DevBuf<double> hcon;                   // the slot-major intermediate itself: contributions of a CSR block slot contiguous
    int hess_layout = 0;                           // 0 tile-major hblk + index-list assembly (default: faster), 1 slot-major hcon + streaming assembly
    bool hblk_valid = false;
    bool maps_ready = false, offsets_ready = false;

    // CSR
    int n_rows = 0, nnz = 0, index_base = 0;
    ipcgpu::DevBuf<int> ia, ja;
    ipcgpu::DevBuf<double> a;
    ipcgpu::DevBuf<int> flag; // device error flag
    // device-resident linear solve (solve.cu): full-row structure of the symmetric matrix + PCG workspace
    ipcgpu::DevBuf<int> fia, fja, fpos;
    bool full_pattern_ready = false;
    ipcgpu::DevBuf<double> sol, pcg_b, pcg_r, pcg_p, pcg_q, pcg_minv, pcg_scal, pcg_hist;

    // work / result buffers
    ipcgpu::DevBuf<double> gcont, hblk, g, e_per_tet, partials, scalar_out, inv_steps, dir, in_partials, e_partials2;
    ipcgpu::DevBuf<double> pSize_dev; // mean |p| of the uploaded search direction, read by the swept-grid kernel from device memory (graph replay)

    // CUDA graphs of device-resident call sequences (ipcgpu_capture_begin / _end / ipcgpu_graph_launch).  A captured sequence mutates a
    // few host-side state words (which scalars are still rank-local, which lists are global ...); they are snapshotted at the end of the
    // capture and re-applied at every replay.  `epoch` is bumped by every call that may reallocate or re-partition: older graphs are refused.
    struct HostState {
        bool energy_local[4], checks_local, lists_local, lists_global, want_cand, swept_ready, fr_ready, inputs_marked, scatter_marked;
        int nC, nP, nK, fr_host_n;
    };
    struct GraphRec {
        cudaGraphExec_t exec = nullptr;
        cudaGraph_t graph = nullptr;
        uint64_t launches = 0, epoch = 0;
        bool dirty_at_begin = false;
        HostState hs;
    };
    std::vector<GraphRec> graphs;
    bool capturing = false;
    uint64_t epoch = 0, launches_at_capture = 0;
    bool dirty_at_capture = false;
    double* h_scalar = nullptr; // pinned staging for scalars

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
        p.n_list = n_list; p.tet_list = (nranks > 1) ? tet_list.p : nullptr;
        p.V = V.p; p.T = T.p; p.Ainv = Ainv.p; p.vol = vol.p; p.mu = mu.p; p.lam = lam.p; p.energy = energy;
        p.e_row_lo = (nranks > 1) ? v_begin : 0; p.e_row_hi = (nranks > 1) ? v_end : nV;
        return p;
    }
};
