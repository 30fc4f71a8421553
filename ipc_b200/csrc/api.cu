// api.cu -- extern "C" entry points declared in include/ipcgpu.h
//
// Execution model.  Every stage enqueues its kernels (and, with several ranks, its NCCL reductions) on the context's stream and leaves
// its scalar results in the device-resident iteration state (kernels.h: IterState).  An entry point synchronises with the host only if
// the caller hands it a host output pointer; with NULL outputs a whole Newton iteration is one uninterrupted stream, read back once by
// ipcgpu_fetch_iteration.  The synchronous forms are the same code followed by that read-back.
#include "../../include/ipcgpu.h"
#include "context.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <dlfcn.h>
#include <numeric>

using namespace ipcgpu;

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                        \
            return IPCGPU_ERR_CUDA;                                                               \
        }                                                                                         \
    } while (0)
#define REQUIRE(cond, code, msg)                                                                  \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            ctx->err = (msg);                                                                     \
            return (code);                                                                        \
        }                                                                                         \
    } while (0)
#define ALLOC(buf, count) REQUIRE((buf).reserve(count), IPCGPU_ERR_CUDA, "cudaMalloc failed for " #buf)

// ---------------------------------------------------------------------------------------------------
// NCCL through dlopen: the library that torch already loaded (libnccl.so.2) is reused when present.
// ---------------------------------------------------------------------------------------------------
namespace {
struct Id128 {
    char b[128];
};
struct Nccl {
    void* h = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /* ncclUniqueId by value: 128 bytes */ Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
Nccl g_nccl;
bool nccl_load(std::string& err)
{
    if (g_nccl.h) return true;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* n : names) {
        g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) {
        err = "dlopen(libnccl.so.2) failed";
        return false;
    }
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(g_nccl.h, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.h, "ncclAllReduce");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(g_nccl.h, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.AllGather) {
        err = "NCCL symbols missing";
        return false;
    }
    return true;
}
constexpr int kNcclInt32 = 2;   // ncclInt32
constexpr int kNcclFloat64 = 8; // ncclDouble
constexpr int kNcclUint64 = 5;  // ncclUint64
constexpr int kNcclSum = 0, kNcclMax = 2, kNcclMin = 3;
} // namespace

// min over ranks of a device-resident uint64 (the order-preserving image of a step), in stream; no-op on one rank
int nccl_min_u64(ipcgpu_ctx* ctx, unsigned long long* word)
{
    if (ctx->nranks <= 1) return IPCGPU_OK;
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_ALLREDUCE);
    int r = g_nccl.AllReduce(word, word, 1, kNcclUint64, kNcclMin, ctx->nccl_comm, ctx->stream);
    ctx->prof_end(pe);
    REQUIRE(r == 0, IPCGPU_ERR_NCCL, "ncclAllReduce(min step) failed");
    return IPCGPU_OK;
}

// make the main stream wait for the copies ipcgpu_download_range_async forked off (a graph edge when capturing)
static int join_copy_stream(ipcgpu_ctx* ctx)
{
    if (!ctx->copy_pending) return IPCGPU_OK;
    CK(cudaEventRecord(ctx->ev_copy_join, ctx->copy));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_join, 0));
    ctx->copy_pending = false;
    return IPCGPU_OK;
}

// one D2H copy of the iteration state + stream synchronisation
int fetch_iter_state(ipcgpu_ctx* ctx)
{
    {
        int rcj = join_copy_stream(ctx);
        if (rcj) return rcj;
    }
    CK(cudaMemcpyAsync(ctx->h_iter, ctx->iter.p, sizeof(IterState), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return IPCGPU_OK;
}
static int clear_flag(ipcgpu_ctx* ctx, int idx)
{
    CK(cudaMemsetAsync(&ctx->iter.p->flags[idx], 0, sizeof(int), ctx->stream));
    return IPCGPU_OK;
}

// ---------------------------------------------------------------------------------------------------
// map building (host, once per mesh/partition): vertex->incident (tet,local) lists and Hessian slots.
// Partition (nranks > 1): rank r owns the rows of the vertex range [v_begin, v_end) -- chosen so that the incident-tet counts
// balance -- and assembles EVERY tet that touches one of them (tets on a range boundary are computed by both neighbours), so that
// each rank's part of the CSR is complete and the Hessian needs no cross-rank reduction.
// ---------------------------------------------------------------------------------------------------
static int build_maps(ipcgpu_ctx* ctx)
{
    const int nT = ctx->nT, nV = ctx->nV;
    const std::vector<int>& T = ctx->h_T;
    ctx->t_begin = (int)((int64_t)nT * ctx->rank / ctx->nranks);
    ctx->t_end = (int)((int64_t)nT * (ctx->rank + 1) / ctx->nranks);
    // vertex ranges balanced by incident-tet count
    std::vector<int64_t> cum(nV + 1, 0);
    for (size_t i = 0; i < (size_t)4 * nT; ++i) ++cum[T[i] + 1];
    for (int v = 0; v < nV; ++v) cum[v + 1] += cum[v];
    auto boundary = [&](int r) -> int {
        if (r <= 0) return 0;
        if (r >= ctx->nranks) return nV;
        const int64_t target = cum[nV] * r / ctx->nranks;
        return (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
    };
    const int vb = std::min(boundary(ctx->rank), nV), ve = std::max(vb, std::min(boundary(ctx->rank + 1), nV));
    ctx->v_begin = vb;
    ctx->v_end = ve;
    // tets touching the owned rows, ascending
    std::vector<int> list;
    if (ctx->nranks == 1) {
        list.resize(nT);
        std::iota(list.begin(), list.end(), 0);
    }
    else {
        list.reserve((size_t)(nT / ctx->nranks) + 1024);
        for (int t = 0; t < nT; ++t) {
            bool touch = false;
            for (int k = 0; k < 4; ++k) {
                const int v = T[(size_t)k * nT + t];
                touch = touch || (v >= vb && v < ve);
            }
            if (touch) list.push_back(t);
        }
    }
    const int nL = (int)list.size();
    ctx->n_list = nL;
    if (list.empty()) list.push_back(0); // keep the upload non-empty
    REQUIRE(ctx->tet_list.upload(list.data(), list.size(), ctx->stream), IPCGPU_ERR_CUDA, "upload of the tet list failed");
    // incidence of the OWNED vertices: counting sort by vertex; entries 4*localTet+loc ascending
    std::vector<int> ptr(nV + 1, 0);
    for (int l = 0; l < nL; ++l)
        for (int k = 0; k < 4; ++k) {
            const int v = T[(size_t)k * nT + list[l]];
            if (v >= vb && v < ve) ++ptr[v + 1];
        }
    for (int v = 0; v < nV; ++v) ptr[v + 1] += ptr[v];
    std::vector<int> inc((size_t)std::max(ptr[nV], 1)), cur(ptr.begin(), ptr.end() - 1);
    for (int l = 0; l < nL; ++l)
        for (int k = 0; k < 4; ++k) {
            const int v = T[(size_t)k * nT + list[l]];
            if (v >= vb && v < ve) inc[cur[v]++] = 4 * l + k;
        }
    if (!ctx->inc_ptr.upload(ptr.data(), ptr.size(), ctx->stream) || !ctx->inc.upload(inc.data(), inc.size(), ctx->stream)) {
        ctx->err = "upload of incidence map failed";
        return IPCGPU_ERR_CUDA;
    }
    // slots: (v<=u) pairs whose ROW vertex v is owned; contributions (key, src) sorted by key then tet
    REQUIRE(((uint64_t)nL + 64ull) * 78ull < 0xffffffffull, IPCGPU_ERR_CAPACITY, "local tet count too large for 32-bit block offsets");
    struct KS {
        uint64_t key;
        unsigned src;
        unsigned tet;
        unsigned slot10; // block slot of the tet: 0..3 diagonal blocks, 4..9 the vertex pairs (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
    };
    std::vector<KS> ks;
    ks.reserve((size_t)10 * nL);
    static const int pa[6] = { 0, 0, 0, 1, 1, 2 }, pb[6] = { 1, 2, 3, 2, 3, 3 };
    for (int l = 0; l < nL; ++l) {
        const int t = list[l];
        int v[4];
        for (int k = 0; k < 4; ++k) v[k] = T[(size_t)k * nT + t];
        // tile-major block addresses (elastic.cu): (l/64)*64*78 + o*64 + (l%64)*len
        const unsigned tl = (unsigned)l, tile_base = (tl / 64u) * (64u * 78u), tin = tl % 64u;
        for (int a = 0; a < 4; ++a)
            if (v[a] >= vb && v[a] < ve) ks.push_back({ ((uint64_t)v[a] << 32) | (uint32_t)v[a], tile_base + 6u * a * 64u + tin * 6u, tl, (unsigned)a });
        for (int q = 0; q < 6; ++q) {
            const int lo = std::min(v[pa[q]], v[pb[q]]), hi = std::max(v[pa[q]], v[pb[q]]);
            if (lo >= vb && lo < ve) ks.push_back({ ((uint64_t)lo << 32) | (uint32_t)hi, tile_base + (24u + 9u * q) * 64u + tin * 9u, tl, 4u + (unsigned)q });
        }
    }
    std::sort(ks.begin(), ks.end(), [](const KS& a, const KS& b) { return a.key < b.key || (a.key == b.key && (a.tet < b.tet || (a.tet == b.tet && a.src < b.src))); });
    std::vector<int> sv, su, cptr;
    std::vector<unsigned> csrc(std::max<size_t>(ks.size(), 1)), cwho(std::max<size_t>(ks.size(), 1)); // cwho: 10 * local tet + block slot of the contribution
    for (size_t i = 0; i < ks.size(); ++i) {
        if (i == 0 || ks[i].key != ks[i - 1].key) {
            sv.push_back((int)(ks[i].key >> 32));
            su.push_back((int)(ks[i].key & 0xffffffffu));
            cptr.push_back((int)i);
        }
        csrc[i] = ks[i].src;
        cwho[i] = ks[i].tet * 10u + ks[i].slot10;
    }
    cptr.push_back((int)ks.size());
    {
        // Slot order = work order of k_assemble_csr (9 threads per slot, a warp covers ~3.5 slots and runs as long as its longest
        // contribution list).  In key order every 7th slot is a diagonal block with ~23 contributions against ~5 for an off-diagonal
        // one, so half of the warps idled most lanes for 20 iterations.  Off-diagonal slots first, then the diagonal ones (key order
        // inside each group keeps the CSR writes local): warps see uniform list lengths.
        const size_t nS = sv.size();
        std::vector<int> order;
        order.reserve(nS);
        for (size_t i = 0; i < nS; ++i)
            if (sv[i] != su[i]) order.push_back((int)i);
        for (size_t i = 0; i < nS; ++i)
            if (sv[i] == su[i]) order.push_back((int)i);
        std::vector<int> sv2(nS), su2(nS), cptr2;
        std::vector<unsigned> csrc2(csrc.size()), cwho2(cwho.size());
        cptr2.reserve(nS + 1);
        size_t pos = 0;
        for (size_t k = 0; k < nS; ++k) {
            const int i = order[k];
            sv2[k] = sv[i];
            su2[k] = su[i];
            cptr2.push_back((int)pos);
            for (int c = cptr[i]; c < cptr[i + 1]; ++c) { csrc2[pos] = csrc[c]; cwho2[pos] = cwho[c]; ++pos; }
        }
        cptr2.push_back((int)pos);
        sv.swap(sv2); su.swap(su2); cptr.swap(cptr2); csrc.swap(csrc2); cwho.swap(cwho2);
    }
    // slot-major intermediate: the contributions of a slot are contiguous (ascending tet order inside the slot), 6 doubles per diagonal and
    // 9 per off-diagonal contribution; hdst tells the per-tet kernel where each of a tet's ten blocks goes (0xffffffff: a row this rank does
    // not own), cbase where a slot's run starts
    std::vector<unsigned> hdst((size_t)10 * std::max(nL, 1), 0xffffffffu), cbase(sv.size() + 1, 0u);
    {
        uint64_t run = 0;
        for (size_t k = 0; k < sv.size() && !ks.empty(); ++k) {
            cbase[k] = (unsigned)run;
            const unsigned len = (sv[k] == su[k]) ? 6u : 9u;
            for (int c = cptr[k]; c < cptr[k + 1]; ++c) {
                hdst[cwho[c]] = (unsigned)run;
                run += len;
            }
        }
        cbase[sv.size()] = (unsigned)run;
        REQUIRE(run < 0xffffffffull, IPCGPU_ERR_CAPACITY, "slot-major intermediate too large for 32-bit offsets");
    }
    ctx->nSlots = (int)sv.size();
    if (sv.empty()) { sv.push_back(0); su.push_back(0); } // keep the uploads non-empty
    bool ok = ctx->slot_v.upload(sv.data(), sv.size(), ctx->stream) && ctx->slot_u.upload(su.data(), su.size(), ctx->stream)
        && ctx->con_ptr.upload(cptr.data(), cptr.size(), ctx->stream) && ctx->con_src.upload(csrc.data(), csrc.size(), ctx->stream)
        && ctx->slot_off.reserve((size_t)3 * std::max(1, ctx->nSlots)) && ctx->hdst.upload(hdst.data(), hdst.size(), ctx->stream)
        && ctx->cbase.upload(cbase.data(), cbase.size(), ctx->stream);
    REQUIRE(ok, IPCGPU_ERR_CUDA, "upload of Hessian scatter map failed");
    ALLOC(ctx->gcont, (size_t)12 * std::max(1, nL));
    ALLOC(ctx->hblk, (size_t)78 * 64 * ((size_t)(std::max(1, nL) + 63) / 64));
    ALLOC(ctx->hcon, (size_t)78 * std::max(1, nL));
    ALLOC(ctx->partials, (size_t)std::max(1, elastic_energy_blocks(ctx->t_end - ctx->t_begin)) + 8);
    CK(cudaStreamSynchronize(ctx->stream)); // host vectors go out of scope
    ctx->maps_ready = true;
    ctx->offsets_ready = false;
    return IPCGPU_OK;
}

static void owned_value_range(ipcgpu_ctx* ctx)
{
    // CSR value range of the owned rows [3 v_begin, 3 v_end)
    if (ctx->h_ia.empty()) return;
    ctx->a_begin = (long long)ctx->h_ia[(size_t)3 * ctx->v_begin] - ctx->index_base;
    ctx->a_end = (long long)ctx->h_ia[(size_t)3 * ctx->v_end] - ctx->index_base;
}

static int ensure_offsets(ipcgpu_ctx* ctx)
{
    if (ctx->offsets_ready) return IPCGPU_OK;
    REQUIRE(ctx->maps_ready, IPCGPU_ERR_STATE, "ipcgpu_set_mesh must precede Hessian assembly");
    REQUIRE(ctx->n_rows == 3 * ctx->nV, IPCGPU_ERR_STATE, "ipcgpu_set_csr must be called with n_rows = 3*nV");
    CK(cudaMemsetAsync(ctx->flag.p, 0, sizeof(int), ctx->stream));
    slot_offsets(ctx->nSlots, ctx->slot_v.p, ctx->slot_u.p, ctx->ia.p, ctx->ja.p, ctx->index_base, ctx->slot_off.p, ctx->flag.p, ctx->stream);
    ++ctx->launches;
    int h = 0;
    CK(cudaMemcpyAsync(&h, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    REQUIRE(h == 0, IPCGPU_ERR_PATTERN, "CSR pattern misses a block of the mesh topology (row<=col entries of every tet vertex pair are required)");
    ctx->offsets_ready = true;
    return IPCGPU_OK;
}

// C++ linkage helpers implemented in constraint.cu / ccd.cu
int contact_alloc(ipcgpu_ctx* ctx);
int contact_constraint_set(ipcgpu_ctx* ctx, double dHat, int wantCand, int* nC, int* nPara, int* nCand);
int contact_sync_counts(ipcgpu_ctx* ctx);
void contact_pack_lists(ipcgpu_ctx* ctx);
void contact_unpack_lists(ipcgpu_ctx* ctx);
int ccd_alloc(ipcgpu_ctx* ctx);
int ccd_narrow(ipcgpu_ctx* ctx, const int2* cand, const int* n32, const unsigned long long* n64, unsigned long long cap, int share, double tol, const double* err_vf,
    const double* err_ee, int stage, const int* overflow);
int ccd_build_swept(ipcgpu_ctx* ctx, double h);
int ccd_full(ipcgpu_ctx* ctx, double tol, const double* err_vf, const double* err_ee);
int ccd_read_back(ipcgpu_ctx* ctx, double* alpha_out);
int solver_build_full_pattern(ipcgpu_ctx* ctx, const int* ia, const int* ja); // solve.cu
int solver_pcg(ipcgpu_ctx* ctx, const double* rhs_dev, double sign, double rel_tol, int max_iter, int* iters_out, double* rel_res_out);
int solver_adopt_direction(ipcgpu_ctx* ctx);
int safeguard_inversion(ipcgpu_ctx* ctx);     // safeguard.cu
int safeguard_intersections(ipcgpu_ctx* ctx); // safeguard.cu

// deferred error flags of the iteration state -> status code (first raised flag wins) and message
static int status_from_flags(ipcgpu_ctx* ctx, const int* f)
{
    if (f[FLAG_NONPOSITIVE_DISTANCE]) {
        ctx->err = "a constraint has d <= 0 (the reference exits here, Optimizer.cpp:3296-3306)";
        return IPCGPU_ERR_NONPOSITIVE_DISTANCE;
    }
    if (f[FLAG_SET_CAPACITY]) {
        ctx->err = "constraint-set capacity exceeded (raise it with ipcgpu_set_pair_capacity)";
        return IPCGPU_ERR_CAPACITY;
    }
    if (f[FLAG_EXCHANGE_CAPACITY]) {
        ctx->err = "pair-list exchange capacity exceeded (65536 pairs per rank and list)";
        return IPCGPU_ERR_CAPACITY;
    }
    if (f[FLAG_CCD_CAPACITY]) {
        ctx->err = "CCD candidate capacity exceeded (raise it with ipcgpu_set_ccd_capacity)";
        return IPCGPU_ERR_CAPACITY;
    }
    if (f[FLAG_PATTERN]) {
        ctx->err = "CSR pattern misses a contact block: call ipcgpu_set_csr with the augmented pattern (augmentConnectivity, SelfCollisionHandler.cpp:330-415)";
        return IPCGPU_ERR_PATTERN;
    }
    return IPCGPU_OK;
}

extern "C" {

int ipcgpu_create(int device, ipcgpu_ctx** out)
{
    if (!out) return IPCGPU_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return IPCGPU_ERR_CUDA; // no CPU fallback by design
    if (device < 0 || device >= ndev) return IPCGPU_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return IPCGPU_ERR_CUDA;
    ipcgpu_ctx* ctx = new ipcgpu_ctx();
    ctx->device = device;
    void* hi = nullptr;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess || cudaMallocHost(&ctx->h_scalar, 512) != cudaSuccess
        || cudaMallocHost(&hi, sizeof(IterState)) != cudaSuccess || !ctx->flag.reserve(4) || !ctx->scalar_out.reserve(32) || !ctx->iter.reserve(1)
        || cudaMemsetAsync(ctx->iter.p, 0, sizeof(IterState), ctx->stream) != cudaSuccess) {
        delete ctx;
        return IPCGPU_ERR_CUDA;
    }
    ctx->h_iter = static_cast<IterState*>(hi);
    std::memset(ctx->h_iter, 0, sizeof(IterState));
    {
        const char* e = std::getenv("IPCGPU_HESS_LAYOUT");
        if (e) ctx->hess_layout = std::atoi(e) == 0 ? 0 : 1;
    }
    {   // side stream for the pair-Hessian build + projection (IPCGPU_BARRIER_OVERLAP=0 keeps everything on one stream).  Replayed from a
        // graph the overlap wins 0.11 ms per iteration on C5 (4.10 -> 3.99 ms); enqueued eagerly the extra event calls cost about as much
        // host time as the overlap wins
        const char* e = std::getenv("IPCGPU_BARRIER_OVERLAP");
        if (!(e && std::atoi(e) == 0)) {
            if (cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_inputs, cudaEventDisableTiming) != cudaSuccess
                || cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess
                || cudaEventCreateWithFlags(&ctx->ev_scatter, cudaEventDisableTiming) != cudaSuccess) {
                if (ctx->side) cudaStreamDestroy(ctx->side);
                ctx->side = nullptr; // fall back to the single-stream order
            }
        }
    }
    step_set(ctx->iter.p, 1.0, ctx->stream);
    *out = ctx;
    return IPCGPU_OK;
}

void ipcgpu_destroy(ipcgpu_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->side) cudaStreamSynchronize(ctx->side);
    if (ctx->copy) {
        cudaStreamSynchronize(ctx->copy);
        cudaStreamDestroy(ctx->copy);
        cudaEventDestroy(ctx->ev_copy_fork);
        cudaEventDestroy(ctx->ev_copy_join);
    }
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& r : ctx->graphs) {
        if (r.exec) cudaGraphExecDestroy(r.exec);
        if (r.graph) cudaGraphDestroy(r.graph);
    }
    if (ctx->ev_inputs) cudaEventDestroy(ctx->ev_inputs);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_scatter) cudaEventDestroy(ctx->ev_scatter);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    for (auto& v : ctx->prof)
        for (auto& pr : v) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
    if (ctx->timer_a) cudaEventDestroy(ctx->timer_a);
    if (ctx->timer_b) cudaEventDestroy(ctx->timer_b);
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->h_scalar) cudaFreeHost(ctx->h_scalar);
    if (ctx->h_iter) cudaFreeHost(ctx->h_iter);
    delete ctx;
}

const char* ipcgpu_last_error(const ipcgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int ipcgpu_host_alloc(void** ptr, uint64_t bytes) { return cudaMallocHost(ptr, bytes) == cudaSuccess ? IPCGPU_OK : IPCGPU_ERR_CUDA; }
int ipcgpu_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? IPCGPU_OK : IPCGPU_ERR_CUDA; }
int ipcgpu_sync(ipcgpu_ctx* ctx)
{
    int rcj = join_copy_stream(ctx);
    if (rcj) return rcj;
    CK(cudaStreamSynchronize(ctx->stream));
    return IPCGPU_OK;
}
uint64_t ipcgpu_launch_count(const ipcgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ipcgpu_comm_unique_id(void* id128)
{
    std::string err;
    if (!id128 || !nccl_load(err)) return IPCGPU_ERR_NCCL;
    return g_nccl.GetUniqueId(id128) == 0 ? IPCGPU_OK : IPCGPU_ERR_NCCL;
}

int ipcgpu_comm_init(ipcgpu_ctx* ctx, int rank, int nranks, const void* id128)
{
    ++ctx->epoch; // graphs captured before this call are refused (buffers, partition or list order may change)
    REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, IPCGPU_ERR_ARG, "bad rank/nranks");
    CK(cudaSetDevice(ctx->device));
    ctx->rank = rank;
    ctx->nranks = nranks;
    if (nranks > 1) {
        REQUIRE(id128 != nullptr, IPCGPU_ERR_ARG, "nccl unique id required for nranks>1");
        REQUIRE(nccl_load(ctx->err), IPCGPU_ERR_NCCL, ctx->err);
        Id128 id;
        std::memcpy(id.b, id128, 128);
        int r = g_nccl.CommInitRank(&ctx->nccl_comm, nranks, id, rank);
        REQUIRE(r == 0, IPCGPU_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    }
    if (ctx->nT > 0) { // re-partition an already loaded mesh
        int rc = build_maps(ctx);
        if (rc) return rc;
        owned_value_range(ctx);
        if (ctx->surface_ready && (rc = contact_alloc(ctx))) return rc;
    }
    return IPCGPU_OK;
}

int ipcgpu_partition_info(ipcgpu_ctx* ctx, int* rank, int* nranks, int* tet_begin, int* tet_end, int* row_vertex_begin, int* row_vertex_end, int64_t* value_begin,
    int64_t* value_end, int* n_assembled_tets)
{
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    if (tet_begin) *tet_begin = ctx->t_begin;
    if (tet_end) *tet_end = ctx->t_end;
    if (row_vertex_begin) *row_vertex_begin = ctx->v_begin;
    if (row_vertex_end) *row_vertex_end = ctx->v_end;
    if (value_begin) *value_begin = ctx->a_begin;
    if (value_end) *value_end = ctx->a_end;
    if (n_assembled_tets) *n_assembled_tets = ctx->n_list;
    return IPCGPU_OK;
}

int ipcgpu_set_mesh(ipcgpu_ctx* ctx, int nV, int nT, const double* Vrest, const int* tets, const double* restTriInv, const double* vol,
    const double* mu, const double* lam, const double* mass, const uint8_t* dbc, int energy)
{
    ++ctx->epoch; // graphs captured before this call are refused (buffers, partition or list order may change)
    REQUIRE(nV > 0 && nT >= 0 && Vrest && tets && restTriInv && vol && mu && lam, IPCGPU_ERR_ARG, "ipcgpu_set_mesh: null or empty input");
    REQUIRE(energy == IPCGPU_NEOHOOKEAN || energy == IPCGPU_FIXED_COROT, IPCGPU_ERR_ARG, "unknown energy type");
    CK(cudaSetDevice(ctx->device));
    for (size_t i = 0; i < (size_t)4 * nT; ++i) REQUIRE(tets[i] >= 0 && tets[i] < nV, IPCGPU_ERR_ARG, "tet vertex index out of range");
    ctx->nV = nV;
    ctx->nT = nT;
    ctx->energy = energy;
    ctx->nVdof = 0x7fffffff; // a new mesh has no obstacle tail until ipcgpu_set_obstacle_tail names one
    ctx->h_T.assign(tets, tets + (size_t)4 * nT);
    ctx->h_ia.clear();
    ctx->nnz = 0;
    ctx->surface_ready = false;
    ctx->dir_valid = false;
    // Dm^-1: reference layout is per-tet column-major; device layout is SoA over the row-major index q=3i+j
    std::vector<double> A((size_t)9 * std::max(nT, 1));
    for (int t = 0; t < nT; ++t)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) A[(size_t)(3 * i + j) * nT + t] = restTriInv[(size_t)9 * t + i + 3 * j];
    bool ok = ctx->Vrest.upload(Vrest, (size_t)3 * nV, ctx->stream) && ctx->V.upload(Vrest, (size_t)3 * nV, ctx->stream)
        && ctx->Vsaved.reserve((size_t)3 * nV) && ctx->T.upload(tets, (size_t)4 * nT, ctx->stream)
        && ctx->Ainv.upload(A.data(), (size_t)9 * nT, ctx->stream) && ctx->vol.upload(vol, nT, ctx->stream)
        && ctx->mu.upload(mu, nT, ctx->stream) && ctx->lam.upload(lam, nT, ctx->stream);
    REQUIRE(ok, IPCGPU_ERR_CUDA, "mesh upload failed");
    ctx->has_mass = mass != nullptr;
    if (mass) REQUIRE(ctx->mass.upload(mass, nV, ctx->stream), IPCGPU_ERR_CUDA, "mass upload failed");
    ctx->has_dbc = dbc != nullptr;
    if (dbc) REQUIRE(ctx->dbc.upload(dbc, nV, ctx->stream), IPCGPU_ERR_CUDA, "dbc upload failed");
    ALLOC(ctx->g, (size_t)3 * nV);
    ALLOC(ctx->dir, (size_t)3 * nV);
    ALLOC(ctx->e_per_tet, (size_t)std::max(nT, 1));
    ALLOC(ctx->inv_steps, (size_t)std::max(nT, 1));
    CK(cudaStreamSynchronize(ctx->stream));
    return build_maps(ctx);
}

int ipcgpu_set_csr(ipcgpu_ctx* ctx, int n_rows, const int* ia, const int* ja, int index_base)
{
    ++ctx->epoch; // graphs captured before this call are refused (buffers, partition or list order may change)
    REQUIRE(n_rows > 0 && ia && ja && (index_base == 0 || index_base == 1), IPCGPU_ERR_ARG, "ipcgpu_set_csr: bad arguments");
    REQUIRE(ctx->nV > 0 && n_rows == 3 * ctx->nV, IPCGPU_ERR_ARG, "ipcgpu_set_csr: n_rows must be 3*nV of the mesh set before");
    CK(cudaSetDevice(ctx->device));
    const int nnz = ia[n_rows] - index_base;
    REQUIRE(nnz >= 0, IPCGPU_ERR_ARG, "ipcgpu_set_csr: negative nnz");
    ctx->n_rows = n_rows;
    ctx->nnz = nnz;
    ctx->index_base = index_base;
    ctx->h_ia.assign(ia, ia + (size_t)n_rows + 1);
    bool ok = ctx->ia.upload(ia, (size_t)n_rows + 1, ctx->stream) && ctx->ja.upload(ja, (size_t)nnz, ctx->stream) && ctx->a.reserve((size_t)std::max(nnz, 1));
    REQUIRE(ok, IPCGPU_ERR_CUDA, "CSR upload failed");
    CK(cudaMemsetAsync(ctx->a.p, 0, (size_t)nnz * sizeof(double), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->a_all_dirty = false;
    ctx->offsets_ready = false;
    ctx->full_pattern_ready = false;
    owned_value_range(ctx);
    return IPCGPU_OK;
}

int ipcgpu_set_state(ipcgpu_ctx* ctx, const double* V)
{
    REQUIRE(ctx->nV > 0, IPCGPU_ERR_STATE, "ipcgpu_set_mesh first");
    CK(cudaSetDevice(ctx->device));
    if (V) CK(cudaMemcpyAsync(ctx->V.p, V, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->mark_inputs();
    return IPCGPU_OK;
}

int ipcgpu_save_state(ipcgpu_ctx* ctx)
{
    REQUIRE(ctx->nV > 0, IPCGPU_ERR_STATE, "ipcgpu_set_mesh first");
    CK(cudaMemcpyAsync(ctx->Vsaved.p, ctx->V.p, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->state_saved = true;
    return IPCGPU_OK;
}

// upload the search direction; pSize = mean |p| over the surface vertices in the reference's serial order (SpatialHash.hpp:603-612),
// taken straight from the caller's array (it is only needed by the swept build and costs one pass over the surface)
static int upload_dir(ipcgpu_ctx* ctx, const double* p)
{
    if (p) {
        CK(cudaMemcpyAsync(ctx->dir.p, p, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        double pSize = 0;
        // (mesh.SVI: with an obstacle attached the surface vertices of the tail do not count -- SpatialHash::build sees the mesh alone)
        int nMeshSV = 0;
        for (int i = 0; i < ctx->nSV; ++i) {
            const int v = ctx->h_SVI[i];
            if (v >= ctx->nVdof) continue;
            ++nMeshSV;
            pSize += std::abs(p[3 * (size_t)v]);
            pSize += std::abs(p[3 * (size_t)v + 1]);
            pSize += std::abs(p[3 * (size_t)v + 2]);
        }
        ctx->pSize = nMeshSV > 0 ? pSize / (double)((long long)nMeshSV * 3) : 0.0;
        // the swept-grid kernel reads it from device memory, so that a captured graph stays valid when the direction changes
        ALLOC(ctx->pSize_dev, 1);
        double* hp = ctx->h_scalar + 32; // pinned staging slot of its own (the copy is asynchronous)
        *hp = ctx->pSize;
        CK(cudaMemcpyAsync(ctx->pSize_dev.p, hp, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->pSize_surface = ctx->surface_ready;
        ctx->dir_valid = true; // (no synchronisation: like every host input of the deferred mode, p must stay untouched until the next fetch)
    }
    REQUIRE(ctx->dir_valid, IPCGPU_ERR_STATE, "no search direction uploaded yet");
    return IPCGPU_OK;
}

int ipcgpu_set_search_dir(ipcgpu_ctx* ctx, const double* p)
{
    REQUIRE(ctx->nV > 0 && p, IPCGPU_ERR_ARG, "ipcgpu_set_search_dir: mesh and p required");
    CK(cudaSetDevice(ctx->device));
    return upload_dir(ctx, p);
}

int ipcgpu_step_forward(ipcgpu_ctx* ctx, const double* p, double alpha)
{
    REQUIRE(ctx->nV > 0, IPCGPU_ERR_STATE, "ipcgpu_set_mesh first");
    REQUIRE(ctx->state_saved, IPCGPU_ERR_STATE, "ipcgpu_save_state must precede ipcgpu_step_forward");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_dir(ctx, p);
    if (rc) return rc;
    step_forward(ctx->nV, ctx->Vsaved.p, ctx->dir.p, alpha, ctx->V.p, ctx->stream);
    ++ctx->launches;
    CK(cudaGetLastError());
    ctx->mark_inputs();
    return IPCGPU_OK;
}

int ipcgpu_elastic_energy(ipcgpu_ctx* ctx, double coef, int /*redoSVD*/, double* E)
{
    REQUIRE(ctx->maps_ready, IPCGPU_ERR_STATE, "ipcgpu_set_mesh first");
    CK(cudaSetDevice(ctx->device));
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_ELASTIC_ENERGY);
    elastic_energy(ctx->eargs(), ctx->e_per_tet.p, ctx->partials.p, coef, ctx->scalar_out.p, ctx->stream);
    ctx->prof_end(pe);
    ctx->launches += 2;
    if (ctx->nranks > 1 && E) { // host result requested: complete it now; NULL = local sum, reduced by ipcgpu_fetch_iteration
        int r = g_nccl.AllReduce(ctx->scalar_out.p, ctx->scalar_out.p, 1, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
        REQUIRE(r == 0, IPCGPU_ERR_NCCL, "ncclAllReduce(energy) failed");
    }
    energy_store(ctx->iter.p, 0, ctx->scalar_out.p, ctx->stream);
    ++ctx->launches;
    ctx->energy_local[0] = (ctx->nranks > 1 && !E);
    CK(cudaGetLastError());
    if (E) {
        CK(cudaMemcpyAsync(ctx->h_scalar, ctx->scalar_out.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        *E = ctx->h_scalar[0];
    }
    return IPCGPU_OK;
}

// zero the part of the value array this rank writes (everything after a cross-rank completion has filled the other rows)
static int zero_values(ipcgpu_ctx* ctx)
{
    if (ctx->nranks > 1 && !ctx->a_all_dirty) {
        if (ctx->a_end > ctx->a_begin) CK(cudaMemsetAsync(ctx->a.p + ctx->a_begin, 0, (size_t)(ctx->a_end - ctx->a_begin) * sizeof(double), ctx->stream));
    }
    else CK(cudaMemsetAsync(ctx->a.p, 0, (size_t)ctx->nnz * sizeof(double), ctx->stream));
    ctx->a_all_dirty = false;
    return IPCGPU_OK;
}

static int run_grad_hess(ipcgpu_ctx* ctx, double coef, int projectSPD, int projectDBC, bool need_g, bool need_h, int add_mass, bool with_energy = false)
{
    REQUIRE(ctx->maps_ready, IPCGPU_ERR_STATE, "ipcgpu_set_mesh first");
    if (need_h) {
        int rc = ensure_offsets(ctx);
        if (rc) return rc;
    }
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_ELASTIC_TET);
    double* e_part = nullptr;
    if (with_energy) { // psi * vol per CTA, summed in fixed order below (computeEnergyVal at the same state: the SVD is shared)
        ALLOC(ctx->e_partials2, (size_t)elastic_grad_hess_blocks(ctx->n_list) + 8);
        e_part = ctx->e_partials2.p;
    }
    const bool slot_major = need_h && ctx->hess_layout == 1;
    elastic_grad_hess(ctx->eargs(), coef, projectSPD, need_g, need_h, ctx->gcont.p, ctx->hblk.p, ctx->stream, e_part, slot_major ? ctx->hdst.p : nullptr, ctx->hcon.p);
    ctx->hblk_valid = need_h && !slot_major;
    ctx->prof_end(pe);
    ++ctx->launches;
    if (with_energy) {
        pe = ctx->prof_begin(IPCGPU_STAGE_ELASTIC_ENERGY);
        reduce_sum(e_part, elastic_grad_hess_blocks(ctx->n_list), coef, ctx->scalar_out.p, ctx->stream);
        energy_store(ctx->iter.p, 0, ctx->scalar_out.p, ctx->stream);
        ctx->prof_end(pe);
        ctx->launches += 2;
        ctx->energy_local[0] = ctx->nranks > 1;
    }
    if (need_g) {
        // owned vertices gather their complete sums (every incident tet is in this rank's list); the other rows are written as zeros
        pe = ctx->prof_begin(IPCGPU_STAGE_GATHER_GRADIENT);
        gather_gradient(ctx->nV, ctx->inc_ptr.p, ctx->inc.p, ctx->gcont.p, ctx->has_dbc ? ctx->dbc.p : nullptr, projectDBC, 0, ctx->g.p, ctx->stream);
        ctx->prof_end(pe);
        ++ctx->launches;
    }
    if (need_h) {
        pe = ctx->prof_begin(IPCGPU_STAGE_ASSEMBLE_CSR);
        if (slot_major)
            assemble_slot_major(ctx->nSlots, ctx->slot_v.p, ctx->slot_u.p, ctx->slot_off.p, ctx->cbase.p, ctx->hcon.p, ctx->has_dbc ? ctx->dbc.p : nullptr, projectDBC, 1, ctx->a.p,
                ctx->stream);
        else
            assemble_csr(ctx->nSlots, ctx->slot_v.p, ctx->slot_u.p, ctx->slot_off.p, ctx->con_ptr.p, ctx->con_src.p, ctx->hblk.p,
                ctx->has_dbc ? ctx->dbc.p : nullptr, projectDBC, nullptr, 1, ctx->a.p, ctx->stream);
        // per-vertex diagonal terms (mass, Dirichlet identity) of the owned rows
        const double* m = (add_mass && ctx->has_mass) ? ctx->mass.p : nullptr;
        diag_mass_dbc_range(ctx->v_begin, ctx->v_end, ctx->ia.p, ctx->index_base, ctx->has_dbc ? ctx->dbc.p : nullptr, projectDBC, m, ctx->a.p, ctx->stream);
        ctx->prof_end(pe);
        ctx->launches += 2;
    }
    CK(cudaGetLastError());
    return IPCGPU_OK;
}

int ipcgpu_elastic_gradient(ipcgpu_ctx* ctx, double coef, int /*redoSVD*/, int projectDBC, double* g)
{
    CK(cudaSetDevice(ctx->device));
    int rc = run_grad_hess(ctx, coef, 1, projectDBC, true, false, 0);
    if (rc) return rc;
    if (ctx->nranks > 1 && g) { // host result requested: complete it across ranks; NULL = deferred (ipcgpu_allreduce_grad_hess)
        rc = ipcgpu_allreduce_grad_hess(ctx, 1, 0);
        if (rc) return rc;
    }
    if (g) {
        CK(cudaMemcpyAsync(g, ctx->g.p, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return IPCGPU_OK;
}

// host value array in (addCoeff semantics): rank 0 contributes it, everybody else starts from zero
static int upload_values(ipcgpu_ctx* ctx, const double* a_host)
{
    if (ctx->rank == 0) CK(cudaMemcpyAsync(ctx->a.p, a_host, (size_t)ctx->nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    else CK(cudaMemsetAsync(ctx->a.p, 0, (size_t)ctx->nnz * sizeof(double), ctx->stream));
    return IPCGPU_OK;
}
static int download_values(ipcgpu_ctx* ctx, double* a_host)
{
    if (ctx->nranks > 1) {
        int rc = ipcgpu_allreduce_grad_hess(ctx, 0, 1);
        if (rc) return rc;
    }
    CK(cudaMemcpyAsync(a_host, ctx->a.p, (size_t)ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return IPCGPU_OK;
}

int ipcgpu_elastic_hessian(ipcgpu_ctx* ctx, double coef, int /*redoSVD*/, int projectSPD, int projectDBC, double* a_inout)
{
    CK(cudaSetDevice(ctx->device));
    REQUIRE(ctx->nnz > 0, IPCGPU_ERR_STATE, "ipcgpu_set_csr first");
    int rc;
    if (a_inout && (rc = upload_values(ctx, a_inout))) return rc;
    if ((rc = run_grad_hess(ctx, coef, projectSPD, projectDBC, false, true, 0))) return rc;
    if (a_inout) return download_values(ctx, a_inout);
    return IPCGPU_OK;
}

int ipcgpu_elastic_grad_hess(ipcgpu_ctx* ctx, double coef, int projectSPD, int projectDBC, int add_mass, double* g, double* a)
{
    CK(cudaSetDevice(ctx->device));
    REQUIRE(ctx->nnz > 0, IPCGPU_ERR_STATE, "ipcgpu_set_csr first");
    // the value array is rebuilt from scratch (LinSysSolver::setZero, then addCoeff of every term): slots that no local tet touches
    // -- contact-only blocks of the augmented pattern -- must not keep last iteration's values
    int rc = zero_values(ctx);
    if (rc) return rc;
    if ((rc = run_grad_hess(ctx, coef, projectSPD, projectDBC, true, true, add_mass))) return rc;
    if (ctx->nranks > 1 && (g || a)) {
        rc = ipcgpu_allreduce_grad_hess(ctx, g ? 1 : 0, a ? 1 : 0);
        if (rc) return rc;
    }
    if (g) CK(cudaMemcpyAsync(g, ctx->g.p, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (a) CK(cudaMemcpyAsync(a, ctx->a.p, (size_t)ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (g || a) CK(cudaStreamSynchronize(ctx->stream));
    return IPCGPU_OK;
}

int ipcgpu_elastic_energy_grad_hess(ipcgpu_ctx* ctx, double coef, int projectSPD, int projectDBC, int add_mass, double* E, double* g, double* a)
{
    CK(cudaSetDevice(ctx->device));
    REQUIRE(ctx->nnz > 0, IPCGPU_ERR_STATE, "ipcgpu_set_csr first");
    // the value array is rebuilt from scratch (LinSysSolver::setZero, then addCoeff of every term): slots that no local tet touches
    // -- contact-only blocks of the augmented pattern -- must not keep last iteration's values
    int rc = zero_values(ctx);
    if (rc) return rc;
    if ((rc = run_grad_hess(ctx, coef, projectSPD, projectDBC, true, true, add_mass, true))) return rc;
    if (ctx->nranks > 1 && (g || a)) {
        rc = ipcgpu_allreduce_grad_hess(ctx, g ? 1 : 0, a ? 1 : 0);
        if (rc) return rc;
    }
    if (g) CK(cudaMemcpyAsync(g, ctx->g.p, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (a) CK(cudaMemcpyAsync(a, ctx->a.p, (size_t)ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (E) { // host result requested: complete it across the ranks now
        if (ctx->nranks > 1) {
            int r = g_nccl.AllReduce(&ctx->iter.p->energy[0], &ctx->iter.p->energy[0], 1, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
            REQUIRE(r == 0, IPCGPU_ERR_NCCL, "ncclAllReduce(energy) failed");
            ctx->energy_local[0] = false;
        }
        CK(cudaMemcpyAsync(ctx->h_scalar, &ctx->iter.p->energy[0], sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        *E = ctx->h_scalar[0];
    }
    else if (g || a) CK(cudaStreamSynchronize(ctx->stream));
    return IPCGPU_OK;
}

// ---- step bound: device-resident chain ---------------------------------------------------------------------
int ipcgpu_step_bound_set(ipcgpu_ctx* ctx, double alpha)
{
    REQUIRE(alpha >= 0.0, IPCGPU_ERR_ARG, "the step must be non-negative");
    CK(cudaSetDevice(ctx->device));
    step_set(ctx->iter.p, alpha, ctx->stream);
    ++ctx->launches;
    CK(cudaGetLastError());
    return IPCGPU_OK;
}

int ipcgpu_inversion_step(ipcgpu_ctx* ctx, const double* p, double slack, double* alpha_inout)
{
    REQUIRE(ctx->maps_ready, IPCGPU_ERR_STATE, "ipcgpu_set_mesh first");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_dir(ctx, p);
    if (rc) return rc;
    if (alpha_inout && (rc = ipcgpu_step_bound_set(ctx, *alpha_inout))) return rc;
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_INVERSION);
    inversion_step(ctx->eargs(), ctx->dir.p, slack, ctx->inv_steps.p, ctx->iter.p, ctx->stream);
    ctx->prof_end(pe);
    ctx->launches += 2;
    if ((rc = nccl_min_u64(ctx, &ctx->iter.p->inv_ord))) return rc;
    inversion_apply(ctx->iter.p, ctx->nT, ctx->stream); // Energy.cpp:576-579
    ++ctx->launches;
    CK(cudaGetLastError());
    if (alpha_inout) return ccd_read_back(ctx, alpha_inout);
    return IPCGPU_OK;
}

// ---- contact ---------------------------------------------------------------------------------------------
int ipcgpu_set_surface(ipcgpu_ctx* ctx, int nSV, const int* SVI, int nSE, const int* SE, int nSF, const int* SF, const int* vCoDim)
{
    ++ctx->epoch; // graphs captured before this call are refused (buffers, partition or list order may change)
    REQUIRE(ctx->nV > 0, IPCGPU_ERR_STATE, "ipcgpu_set_mesh first");
    REQUIRE(nSV >= 0 && nSE >= 0 && nSF >= 0 && (nSV == 0 || SVI) && (nSE == 0 || SE) && (nSF == 0 || SF), IPCGPU_ERR_ARG, "ipcgpu_set_surface: bad arguments");
    CK(cudaSetDevice(ctx->device));
    for (int i = 0; i < nSV; ++i) REQUIRE(SVI[i] >= 0 && SVI[i] < ctx->nV, IPCGPU_ERR_ARG, "SVI out of range");
    for (int i = 0; i < 2 * nSE; ++i) REQUIRE(SE[i] >= 0 && SE[i] < ctx->nV, IPCGPU_ERR_ARG, "SFEdges out of range");
    for (size_t i = 0; i < (size_t)3 * nSF; ++i) REQUIRE(SF[i] >= 0 && SF[i] < ctx->nV, IPCGPU_ERR_ARG, "SF out of range");
    ctx->nSV = nSV; ctx->nSE = nSE; ctx->nSF = nSF;
    bool ok = ctx->SVI.upload(SVI, std::max(nSV, 0), ctx->stream) && ctx->SE.upload(SE, (size_t)2 * nSE, ctx->stream) && ctx->SF.upload(SF, (size_t)3 * nSF, ctx->stream);
    REQUIRE(ok, IPCGPU_ERR_CUDA, "surface upload failed");
    ctx->has_codim = vCoDim != nullptr;
    if (vCoDim) REQUIRE(ctx->vCoDim.upload(vCoDim, ctx->nV, ctx->stream), IPCGPU_ERR_CUDA, "codim upload failed");
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->h_SVI.assign(SVI, SVI + nSV);
    ctx->pSize_surface = false; // pSize belongs to the surface
    int rc = contact_alloc(ctx);
    if (rc) return rc;
    if ((rc = ccd_alloc(ctx))) return rc;
    ctx->surface_ready = true;
    return IPCGPU_OK;
}

int ipcgpu_set_obstacle_tail(ipcgpu_ctx* ctx, int first_obstacle_vertex, int ee_through_vf_routine)
{
    ++ctx->epoch; // graphs captured before this call are refused (the pair rules change)
    REQUIRE(ctx->nV > 0, IPCGPU_ERR_STATE, "ipcgpu_set_mesh first");
    if (first_obstacle_vertex < 0 || first_obstacle_vertex >= ctx->nV) { // no obstacle
        ctx->nVdof = 0x7fffffff;
        ctx->ee_as_vf = ee_through_vf_routine ? 1 : 0;
        ctx->pSize_surface = false;
        return IPCGPU_OK;
    }
    REQUIRE(first_obstacle_vertex > 0, IPCGPU_ERR_ARG, "the mesh needs at least one vertex of its own");
    for (size_t i = 0; i < ctx->h_T.size(); ++i) REQUIRE(ctx->h_T[i] < first_obstacle_vertex, IPCGPU_ERR_ARG, "a tetrahedron uses an obstacle vertex");
    REQUIRE(ctx->has_dbc, IPCGPU_ERR_STATE, "the obstacle's vertices must be flagged Dirichlet (1) in ipcgpu_set_mesh: their rows never reach the system");
    {
        std::vector<uint8_t> tail((size_t)(ctx->nV - first_obstacle_vertex));
        CK(cudaMemcpyAsync(tail.data(), ctx->dbc.p + first_obstacle_vertex, tail.size(), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (uint8_t f : tail) REQUIRE(f == 1, IPCGPU_ERR_ARG, "the obstacle's vertices must be flagged Dirichlet (1) in ipcgpu_set_mesh");
    }
    ctx->nVdof = first_obstacle_vertex;
    ctx->ee_as_vf = ee_through_vf_routine ? 1 : 0;
    ctx->pSize_surface = false; // the mean |p| of the swept build is taken over the MESH's surface vertices: upload the direction again
    return IPCGPU_OK;
}

int ipcgpu_set_obstacle_positions(ipcgpu_ctx* ctx, const double* Vo_soa)
{
    REQUIRE(ctx->nV > 0 && ctx->nVdof < ctx->nV, IPCGPU_ERR_STATE, "ipcgpu_set_obstacle_tail first");
    REQUIRE(Vo_soa, IPCGPU_ERR_ARG, "null argument");
    CK(cudaSetDevice(ctx->device));
    const size_t nVo = (size_t)(ctx->nV - ctx->nVdof);
    // current AND rest positions: the obstacle has no rest shape of its own, compute_eps_x takes its current edge lengths
    // (MeshCollisionUtils.hpp:2976-2981); SoA with the stride of the whole vertex array
    // (the saved line-search state keeps the old tail: move the obstacle BETWEEN line searches, or call ipcgpu_save_state again afterwards)
    for (double* dst : { ctx->V.p, ctx->Vrest.p })
        CK(cudaMemcpy2DAsync(dst + ctx->nVdof, (size_t)ctx->nV * sizeof(double), Vo_soa, nVo * sizeof(double), nVo * sizeof(double), 3, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); // (pageable host memory)
    ctx->mark_inputs();
    return IPCGPU_OK;
}

int ipcgpu_set_ccd_capacity(ipcgpu_ctx* ctx, uint64_t capacity)
{
    ++ctx->epoch; // graphs captured before this call are refused (buffers, partition or list order may change)
    REQUIRE(capacity > 0 && capacity < 0xffffffffull, IPCGPU_ERR_ARG, "capacity out of range");
    ctx->ccd_capacity = (size_t)capacity;
    if (ctx->surface_ready) return ccd_alloc(ctx);
    return IPCGPU_OK;
}

int ipcgpu_ti_error(const double* V, int nV, const double* p, double err_vf[3], double err_ee[3])
{
    if (!V || nV <= 0 || !err_vf || !err_ee) return IPCGPU_ERR_ARG;
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for (int v = 0; v < nV; ++v)
        for (int c = 0; c < 3; ++c) {
            const double x = V[(size_t)c * nV + v];
            lo[c] = std::min(lo[c], x);
            hi[c] = std::max(hi[c], x);
            if (p) {
                const double y = x + p[3 * (size_t)v + c];
                lo[c] = std::min(lo[c], y);
                hi[c] = std::max(hi[c], y);
            }
        }
    double diag2 = 0.0;
    for (int c = 0; c < 3; ++c) diag2 += (hi[c] - lo[c]) * (hi[c] - lo[c]);
    const double radius = 0.5 * std::sqrt(diag2);
    // Tight-Inclusion get_numerical_error with minimum separation: filter * max(1, |x|max)^3
    const double ee_filter = 7.105427357601002e-15, vf_filter = 7.549516567451064e-15;
    for (int c = 0; c < 3; ++c) {
        const double center = 0.5 * (lo[c] + hi[c]);
        const double a = center - 10.0 * radius / std::sqrt(3.0), b = center + 10.0 * radius / std::sqrt(3.0);
        double m = std::max(std::fabs(a), std::fabs(b));
        m = std::max(m, 1.0);
        err_ee[c] = m * m * m * ee_filter;
        err_vf[c] = m * m * m * vf_filter;
    }
    return IPCGPU_OK;
}

int ipcgpu_ccd_debug_seed_bound(ipcgpu_ctx* ctx, double toi)
{
    ctx->debug_prune_seed = toi;
    return IPCGPU_OK;
}

int ipcgpu_ccd_partial_ti(ipcgpu_ctx* ctx, const double* p, double tol, const double err_vf[3], const double err_ee[3], double* alpha_inout)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    REQUIRE(err_vf && err_ee, IPCGPU_ERR_ARG, "null argument");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_dir(ctx, p);
    if (rc) return rc;
    if (alpha_inout && (rc = ipcgpu_step_bound_set(ctx, *alpha_inout))) return rc;
    // CFL_FOR_CCD != 0: an empty candidate list leaves the step unchanged (:700).  Replicated lists: every rank walks a contiguous
    // slice; partitioned lists (ipcgpu_set_contact_partition) are already this rank's own.
    ContactWork& w = ctx->cw;
    const int share = (ctx->nranks > 1 && !ctx->lists_local) ? 1 : 0;
    if ((rc = ccd_narrow(ctx, w.cand.p, w.counters.p + 3, nullptr, (unsigned long long)4 * w.cap, share, tol, err_vf, err_ee, 1, nullptr))) return rc;
    if (alpha_inout) return ccd_read_back(ctx, alpha_inout);
    return IPCGPU_OK;
}

int ipcgpu_hash_build_swept(ipcgpu_ctx* ctx, const double* p, double* alpha_inout, double h)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    REQUIRE(h > 0.0, IPCGPU_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_dir(ctx, p);
    if (rc) return rc;
    REQUIRE(ctx->pSize_surface, IPCGPU_ERR_STATE, "the search direction was uploaded before ipcgpu_set_surface: upload it again");
    if (alpha_inout && (rc = ipcgpu_step_bound_set(ctx, *alpha_inout))) return rc;
    if ((rc = ccd_build_swept(ctx, h))) return rc;
    if (alpha_inout) return ccd_read_back(ctx, alpha_inout);
    return IPCGPU_OK;
}

int ipcgpu_ccd_full_ti(ipcgpu_ctx* ctx, double tol, const double err_vf[3], const double err_ee[3], double* alpha_inout, uint64_t* n_candidates)
{
    REQUIRE(ctx->surface_ready && ctx->ccd.swept_ready, IPCGPU_ERR_STATE, "ipcgpu_hash_build_swept first");
    REQUIRE(err_vf && err_ee, IPCGPU_ERR_ARG, "null argument");
    CK(cudaSetDevice(ctx->device));
    int rc;
    // (the swept grid was built for the step the chain held then; a host step that differs from it only lowers max_t)
    if (alpha_inout && (rc = ipcgpu_step_bound_set(ctx, *alpha_inout))) return rc;
    if ((rc = ccd_full(ctx, tol, err_vf, err_ee))) return rc;
    if (alpha_inout || n_candidates) {
        if ((rc = ccd_read_back(ctx, alpha_inout))) return rc;
        if (n_candidates) *n_candidates = ctx->h_iter->n_full_cand;
        if (ctx->h_iter->flags[FLAG_CCD_CAPACITY]) {
            clear_flag(ctx, FLAG_CCD_CAPACITY);
            int only[8] = { 0 };
            only[FLAG_CCD_CAPACITY] = 1;
            return status_from_flags(ctx, only);
        }
    }
    return IPCGPU_OK;
}

int ipcgpu_ccd_stats(ipcgpu_ctx* ctx, uint64_t* candidates, uint64_t* survivors, uint64_t* warnings)
{
    if (candidates) *candidates = ctx->ccd.last_candidates;
    if (survivors) *survivors = ctx->ccd.last_survivors;
    if (warnings) *warnings = (uint64_t)ctx->ccd.last_warnings;
    return IPCGPU_OK;
}

int ipcgpu_ccd_stats_ex(ipcgpu_ctx* ctx, uint64_t* deferred, uint64_t* boxes_thread_pass, uint64_t* boxes_warp_pass)
{
    if (deferred) *deferred = ctx->ccd.last_deferred;
    if (boxes_thread_pass) *boxes_thread_pass = ctx->ccd.last_boxes_thread;
    if (boxes_warp_pass) *boxes_warp_pass = ctx->ccd.last_boxes_warp;
    return IPCGPU_OK;
}

int ipcgpu_ccd_stats_timing(ipcgpu_ctx* ctx, uint64_t* longest_pair_cycles, uint64_t* total_pair_cycles)
{
    if (longest_pair_cycles) *longest_pair_cycles = ctx->ccd.last_longest_cycles;
    if (total_pair_cycles) *total_pair_cycles = ctx->ccd.last_total_cycles;
    return IPCGPU_OK;
}

int ipcgpu_set_hessian_layout(ipcgpu_ctx* ctx, int layout)
{
    ++ctx->epoch; // graphs captured before this call are refused (another kernel pair)
    REQUIRE(layout == 0 || layout == 1, IPCGPU_ERR_ARG, "layout: 0 tile-major (per-tet blocks downloadable), 1 slot-major");
    ctx->hess_layout = layout;
    return IPCGPU_OK;
}

int ipcgpu_set_exchange_capacity(ipcgpu_ctx* ctx, int pairs_per_rank)
{
    ++ctx->epoch; // graphs captured before this call are refused (the message buffers change)
    REQUIRE(pairs_per_rank > 0, IPCGPU_ERR_ARG, "capacity must be positive");
    ctx->exchange_capacity = pairs_per_rank;
    if (ctx->surface_ready) return contact_alloc(ctx);
    return IPCGPU_OK;
}

int ipcgpu_set_pair_capacity(ipcgpu_ctx* ctx, int capacity)
{
    ++ctx->epoch; // graphs captured before this call are refused (buffers, partition or list order may change)
    REQUIRE(capacity > 0, IPCGPU_ERR_ARG, "capacity must be positive");
    ctx->pair_capacity = capacity;
    if (ctx->surface_ready) return contact_alloc(ctx);
    return IPCGPU_OK;
}

int ipcgpu_constraint_set(ipcgpu_ctx* ctx, double dHat, int getPTEE, int* nC, int* nPara, int* nCand)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    REQUIRE(dHat > 0.0, IPCGPU_ERR_ARG, "dHat must be positive");
    CK(cudaSetDevice(ctx->device));
    int rc = contact_constraint_set(ctx, dHat, getPTEE, nC, nPara, nCand);
    ctx->lists_local = (rc == 0) && ctx->partition_contact && ctx->nranks > 1;
    ctx->cw.lists_global = false;
    if (rc == 0 && ctx->lists_local) {
        // every rank holds a disjoint part of the sets: exchange them (one fixed-size message per rank) so that each rank can assemble
        // the Hessian rows it owns from ALL pairs that touch them
        ContactWork& w = ctx->cw;
        contact_pack_lists(ctx);
        cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_ALLREDUCE);
        int r = g_nccl.AllGather(w.xsend.p, w.xrecv.p, w.xstride * 4, kNcclInt32, ctx->nccl_comm, ctx->stream);
        ctx->prof_end(pe);
        REQUIRE(r == 0, IPCGPU_ERR_NCCL, "ncclAllGather(pair lists) failed");
        contact_unpack_lists(ctx);
        w.lists_global = true;
        CK(cudaGetLastError());
    }
    ctx->mark_inputs();
    return rc;
}

int ipcgpu_set_canonical_order(ipcgpu_ctx* ctx, int enable)
{
    ++ctx->epoch; // graphs captured before this call are refused (buffers, partition or list order may change)
    ctx->canonical_order = enable != 0;
    return IPCGPU_OK;
}

int ipcgpu_set_contact_partition(ipcgpu_ctx* ctx, int enable)
{
    ++ctx->epoch; // graphs captured before this call are refused (buffers, partition or list order may change)
    ctx->partition_contact = enable != 0;
    return IPCGPU_OK;
}

int ipcgpu_get_constraint_set(ipcgpu_ctx* ctx, int* mm, int* para, int* para_e, int* cand)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    ContactWork& w = ctx->cw;
    if (w.nC < 0) { // built without a read-back: fetch the sizes now
        int rc = contact_sync_counts(ctx);
        if (rc) return rc;
    }
    if (mm && w.nC) CK(cudaMemcpyAsync(mm, w.act.p, (size_t)w.nC * sizeof(int4), cudaMemcpyDeviceToHost, ctx->stream));
    if (para && w.nP) CK(cudaMemcpyAsync(para, w.para.p, (size_t)w.nP * sizeof(int4), cudaMemcpyDeviceToHost, ctx->stream));
    if (para_e && w.nP) CK(cudaMemcpyAsync(para_e, w.para_e.p, (size_t)w.nP * sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream));
    if (cand && w.nK) CK(cudaMemcpyAsync(cand, w.cand.p, (size_t)w.nK * sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return IPCGPU_OK;
}

int ipcgpu_constraint_set_sizes(ipcgpu_ctx* ctx, int* nC, int* nPara, int* nCand)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    ContactWork& w = ctx->cw;
    if (w.nC < 0) {
        int rc = contact_sync_counts(ctx);
        if (rc) return rc;
    }
    if (nC) *nC = w.nC;
    if (nPara) *nPara = w.nP;
    if (nCand) *nCand = w.nK;
    return IPCGPU_OK;
}

int ipcgpu_set_constraint_set(ipcgpu_ctx* ctx, int nC, const int* mm, int nP, const int* para, const int* para_e, int nK, const int* cand)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    ContactWork& w = ctx->cw;
    REQUIRE(nC >= 0 && nP >= 0 && nK >= 0 && nC <= w.cap && nP <= w.cap && nK <= 4 * w.cap, IPCGPU_ERR_CAPACITY, "set exceeds the pair capacity");
    if (nC) CK(cudaMemcpyAsync(w.act.p, mm, (size_t)nC * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
    if (nP) CK(cudaMemcpyAsync(w.para.p, para, (size_t)nP * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
    if (nP) CK(cudaMemcpyAsync(w.para_e.p, para_e, (size_t)nP * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    if (nK) CK(cudaMemcpyAsync(w.cand.p, cand, (size_t)nK * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    int* h = reinterpret_cast<int*>(ctx->h_scalar);
    for (int i = 0; i < 16; ++i) h[i] = 0;
    h[0] = nC; h[2] = nP; h[3] = nK;
    CK(cudaMemcpyAsync(w.counters.p, h, 16 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)); // the consumers read the sizes on the device
    CK(cudaStreamSynchronize(ctx->stream));
    w.nC = nC; w.nP = nP; w.nK = nK;
    w.want_cand = nK > 0;
    ctx->lists_local = false; // uploaded sets are the global ones
    w.lists_global = false;
    ctx->mark_inputs();
    return IPCGPU_OK;
}

static BarrierArgs barrier_args(ipcgpu_ctx* ctx, double dHat, double kappa, int projectDBC)
{
    BarrierArgs p;
    ContactWork& w = ctx->cw;
    p.nV = ctx->nV; p.V = ctx->V.p; p.Vrest = ctx->Vrest.p; p.dbc = ctx->has_dbc ? ctx->dbc.p : nullptr; p.SE = ctx->SE.p;
    p.nVdof = ctx->nVdof;
    // one rank: the lists as built.  Several ranks: the GLOBAL lists (replicated build, or partitioned build + exchange); energy and
    // gradient take a contiguous share of them, the Hessian goes by row owner.
    if (ctx->nranks > 1 && w.lists_global) {
        p.cs = w.gact.p; p.nC = w.counters.p + 10; p.para = w.gpara.p; p.para_e = w.gpara_e.p; p.nP = w.counters.p + 11;
    }
    else {
        p.cs = w.act.p; p.nC = w.counters.p + 0; p.para = w.para.p; p.para_e = w.para_e.p; p.nP = w.counters.p + 2;
    }
    p.rank = ctx->rank; p.nranks = ctx->nranks; p.share = ctx->nranks > 1 ? 1 : 0;
    p.row_lo = ctx->nranks > 1 ? ctx->v_begin : 0;
    p.row_hi = ctx->nranks > 1 ? ctx->v_end : ctx->nV;
    p.dHat = dHat; p.kappa = kappa; p.projectDBC = projectDBC;
    p.ia = ctx->ia.p; p.ja = ctx->ja.p; p.base = ctx->index_base;
    return p;
}

int ipcgpu_barrier_energy(ipcgpu_ctx* ctx, double dHat, double kappa, double* E)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    CK(cudaSetDevice(ctx->device));
    BarrierArgs p = barrier_args(ctx, dHat, kappa, 0);
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_BARRIER);
    barrier_energy(p, ctx->cw.bpartials.p, &ctx->iter.p->flags[FLAG_NONPOSITIVE_DISTANCE], ctx->stream);
    reduce_sum(ctx->cw.bpartials.p, barrier_energy_blocks(), kappa, ctx->scalar_out.p + 1, ctx->stream);
    ctx->prof_end(pe);
    ctx->launches += 2;
    if (ctx->nranks > 1 && E) {
        int r = g_nccl.AllReduce(ctx->scalar_out.p + 1, ctx->scalar_out.p + 1, 1, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
        REQUIRE(r == 0, IPCGPU_ERR_NCCL, "ncclAllReduce(barrier energy) failed");
    }
    energy_store(ctx->iter.p, 1, ctx->scalar_out.p + 1, ctx->stream);
    ++ctx->launches;
    ctx->energy_local[1] = (ctx->nranks > 1 && !E);
    CK(cudaGetLastError());
    if (E) { // synchronous form: the d <= 0 flag is checked right here (every rank checks its own share of the pairs)
        int rc = fetch_iter_state(ctx);
        if (rc) return rc;
        *E = ctx->h_iter->energy[1];
        if (ctx->h_iter->flags[FLAG_NONPOSITIVE_DISTANCE]) {
            clear_flag(ctx, FLAG_NONPOSITIVE_DISTANCE);
            int only[8] = { 0 };
            only[FLAG_NONPOSITIVE_DISTANCE] = 1;
            return status_from_flags(ctx, only);
        }
    }
    return IPCGPU_OK;
}

int ipcgpu_barrier_gradient(ipcgpu_ctx* ctx, double dHat, double kappa, double* g_inout)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    CK(cudaSetDevice(ctx->device));
    if (g_inout) {
        if (ctx->rank == 0) CK(cudaMemcpyAsync(ctx->g.p, g_inout, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        else CK(cudaMemsetAsync(ctx->g.p, 0, (size_t)3 * ctx->nV * sizeof(double), ctx->stream));
    }
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_BARRIER);
    barrier_gradient(barrier_args(ctx, dHat, kappa, 0), ctx->g.p, ctx->stream);
    ctx->prof_end(pe);
    ++ctx->launches;
    CK(cudaGetLastError());
    if (g_inout && ctx->nranks > 1) {
        int rc = ipcgpu_allreduce_grad_hess(ctx, 1, 0);
        if (rc) return rc;
    }
    if (g_inout) {
        CK(cudaMemcpyAsync(g_inout, ctx->g.p, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return IPCGPU_OK;
}

// host gradient in/out around a kernel that accumulates into the device gradient (addCoeff-like semantics: rank 0 contributes the input)
static int gradient_roundtrip_begin(ipcgpu_ctx* ctx, const double* g_in)
{
    if (ctx->rank == 0) CK(cudaMemcpyAsync(ctx->g.p, g_in, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    else CK(cudaMemsetAsync(ctx->g.p, 0, (size_t)3 * ctx->nV * sizeof(double), ctx->stream));
    return IPCGPU_OK;
}
static int gradient_roundtrip_end(ipcgpu_ctx* ctx, double* g_out)
{
    if (ctx->nranks > 1) {
        int rc = ipcgpu_allreduce_grad_hess(ctx, 1, 0);
        if (rc) return rc;
    }
    CK(cudaMemcpyAsync(g_out, ctx->g.p, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return IPCGPU_OK;
}

int ipcgpu_evaluate_constraints(ipcgpu_ctx* ctx, double* val, int n)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    CK(cudaSetDevice(ctx->device));
    ContactWork& w = ctx->cw;
    if (w.nC < 0) {
        int rc = contact_sync_counts(ctx);
        if (rc) return rc;
    }
    REQUIRE(n == w.nC && (val || n == 0), IPCGPU_ERR_ARG, "ipcgpu_evaluate_constraints: n must be the size of the active set on this context");
    ALLOC(w.bval, (size_t)std::max(w.cap, 1));
    BarrierArgs p = barrier_args(ctx, 1.0, 1.0, 0);
    p.cs = w.act.p; p.nC = w.counters.p + 0; // this context's own active list (what ipcgpu_get_constraint_set returns), not the exchanged one
    evaluate_constraints(p, w.bval.p, ctx->stream);
    ++ctx->launches;
    CK(cudaGetLastError());
    if (n) CK(cudaMemcpyAsync(val, w.bval.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return IPCGPU_OK;
}

int ipcgpu_constraint_jacobian_t(ipcgpu_ctx* ctx, const double* input, int n, double coef, double* g_inout)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    REQUIRE(g_inout != nullptr, IPCGPU_ERR_ARG, "null gradient");
    CK(cudaSetDevice(ctx->device));
    ContactWork& w = ctx->cw;
    if (w.nC < 0) {
        int rc = contact_sync_counts(ctx);
        if (rc) return rc;
    }
    REQUIRE(n == w.nC && (input || n == 0), IPCGPU_ERR_ARG, "ipcgpu_constraint_jacobian_t: n must be the size of the active set on this context");
    REQUIRE(!(ctx->nranks > 1 && ctx->lists_local), IPCGPU_ERR_STATE, "per-constraint input needs the replicated sets (ipcgpu_set_contact_partition(0))");
    ALLOC(w.bval, (size_t)std::max(w.cap, 1));
    int rc = gradient_roundtrip_begin(ctx, g_inout);
    if (rc) return rc;
    if (n) CK(cudaMemcpyAsync(w.bval.p, input, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    BarrierArgs p = barrier_args(ctx, 1.0, 1.0, 0);
    p.cs = w.act.p; p.nC = w.counters.p + 0;
    constraint_jacobian_t(p, w.bval.p, coef, ctx->g.p, ctx->stream);
    ++ctx->launches;
    CK(cudaGetLastError());
    return gradient_roundtrip_end(ctx, g_inout);
}

int ipcgpu_para_ee_gradient(ipcgpu_ctx* ctx, double dHat, double kappa, double* g_inout)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    REQUIRE(g_inout != nullptr, IPCGPU_ERR_ARG, "null gradient");
    CK(cudaSetDevice(ctx->device));
    int rc = gradient_roundtrip_begin(ctx, g_inout);
    if (rc) return rc;
    para_gradient(barrier_args(ctx, dHat, kappa, 0), ctx->g.p, ctx->stream);
    ++ctx->launches;
    CK(cudaGetLastError());
    return gradient_roundtrip_end(ctx, g_inout);
}

int ipcgpu_barrier_hessian(ipcgpu_ctx* ctx, double dHat, double kappa, int projectDBC, double* a_inout)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    REQUIRE(ctx->nnz > 0, IPCGPU_ERR_STATE, "ipcgpu_set_csr first");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if (a_inout && (rc = upload_values(ctx, a_inout))) return rc;
    ContactWork& w = ctx->cw;
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_BARRIER);
    const BarrierArgs bp = barrier_args(ctx, dHat, kappa, projectDBC);
    if (ctx->side && ctx->inputs_marked) {
        // build + projection on the side stream, ordered after the last change of their inputs (positions, contact sets) -- i.e. next to
        // whatever the main stream has queued since (the elastic assembly); the scatter joins the main stream, after the elastic writes
        CK(cudaStreamWaitEvent(ctx->side, ctx->ev_inputs, 0));
        if (ctx->scatter_marked) CK(cudaStreamWaitEvent(ctx->side, ctx->ev_scatter, 0)); // back-to-back calls: the last scatter still reads the buffers
        barrier_hessian_build_project(bp, ctx->iter.p->flags, w.bHraw.p, w.brows.p, w.bpsd.p, w.counters.p + 12, w.cap, ctx->side);
        CK(cudaEventRecord(ctx->ev_join, ctx->side));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    }
    else barrier_hessian_build_project(bp, ctx->iter.p->flags, w.bHraw.p, w.brows.p, w.bpsd.p, w.counters.p + 12, w.cap, ctx->stream);
    barrier_hessian_scatter(bp, ctx->a.p, ctx->iter.p->flags, w.bHraw.p, w.brows.p, w.bpsd.p, w.counters.p + 12, w.cap, ctx->stream);
    if (ctx->side && ctx->ev_scatter) ctx->scatter_marked = (cudaEventRecord(ctx->ev_scatter, ctx->stream) == cudaSuccess);
    ctx->prof_end(pe);
    ctx->launches += 3;
    CK(cudaGetLastError());
    if (a_inout) { // synchronous form: complete across ranks, download, check the pattern flag now
        if ((rc = download_values(ctx, a_inout))) return rc;
        if ((rc = fetch_iter_state(ctx))) return rc;
        if (ctx->h_iter->flags[FLAG_PATTERN] || ctx->h_iter->flags[FLAG_SET_CAPACITY]) {
            int only[8] = { 0 };
            only[FLAG_PATTERN] = ctx->h_iter->flags[FLAG_PATTERN];
            only[FLAG_SET_CAPACITY] = ctx->h_iter->flags[FLAG_SET_CAPACITY];
            clear_flag(ctx, FLAG_PATTERN);
            clear_flag(ctx, FLAG_SET_CAPACITY);
            return status_from_flags(ctx, only);
        }
    }
    return IPCGPU_OK;
}

// ---- lagged friction of the self-contact pairs --------------------------------------------------------------
int ipcgpu_set_prev_state(ipcgpu_ctx* ctx, const double* V_prev_soa)
{
    REQUIRE(ctx->nV > 0, IPCGPU_ERR_STATE, "ipcgpu_set_mesh first");
    CK(cudaSetDevice(ctx->device));
    ALLOC(ctx->Vprev, (size_t)3 * ctx->nV);
    if (V_prev_soa) CK(cudaMemcpyAsync(ctx->Vprev.p, V_prev_soa, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    else CK(cudaMemcpyAsync(ctx->Vprev.p, ctx->V.p, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream)); // V_prev = V at the start of a step
    ctx->prev_set = true;
    return IPCGPU_OK;
}

static int friction_alloc(ipcgpu_ctx* ctx)
{
    ContactWork& w = ctx->cw;
    const size_t cap = (size_t)std::max(ctx->pair_capacity, 1);
    ALLOC(w.fr_cs, cap);
    ALLOC(w.fr_n, 1);
    ALLOC(w.fr_lambda, cap);
    ALLOC(w.fr_coord, cap);
    ALLOC(w.fr_basis, 6 * cap);
    ALLOC(w.fr_partials, (size_t)friction_energy_blocks() + 8);
    return IPCGPU_OK;
}

int ipcgpu_friction_lag(ipcgpu_ctx* ctx, double dHat, double kappa, int* n_pairs)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    CK(cudaSetDevice(ctx->device));
    int rc = friction_alloc(ctx);
    if (rc) return rc;
    ContactWork& w = ctx->cw;
    // the lagged set is the whole active set on every rank (the global list after a partitioned build)
    friction_lag(barrier_args(ctx, dHat, kappa, 0), w.fr_cs.p, w.fr_n.p, w.fr_lambda.p, w.fr_coord.p, w.fr_basis.p, ctx->pair_capacity,
        &ctx->iter.p->flags[FLAG_NONPOSITIVE_DISTANCE], ctx->stream);
    ++ctx->launches;
    CK(cudaGetLastError());
    w.fr_ready = true;
    w.fr_host_n = -1;
    if (n_pairs) {
        CK(cudaMemcpyAsync(ctx->h_scalar, w.fr_n.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        w.fr_host_n = *reinterpret_cast<int*>(ctx->h_scalar);
        *n_pairs = w.fr_host_n;
    }
    return IPCGPU_OK;
}

static int friction_host_count(ipcgpu_ctx* ctx)
{
    ContactWork& w = ctx->cw;
    if (w.fr_host_n < 0) {
        CK(cudaMemcpyAsync(ctx->h_scalar, w.fr_n.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        w.fr_host_n = *reinterpret_cast<int*>(ctx->h_scalar);
    }
    return IPCGPU_OK;
}

int ipcgpu_get_friction_data(ipcgpu_ctx* ctx, int* n_pairs, int* mmcvid4, double* lambda, double* coord2, double* basis6)
{
    REQUIRE(ctx->cw.fr_ready, IPCGPU_ERR_STATE, "ipcgpu_friction_lag / ipcgpu_set_friction_data first");
    CK(cudaSetDevice(ctx->device));
    int rc = friction_host_count(ctx);
    if (rc) return rc;
    ContactWork& w = ctx->cw;
    const size_t n = (size_t)w.fr_host_n;
    if (n_pairs) *n_pairs = w.fr_host_n;
    if (n && mmcvid4) CK(cudaMemcpyAsync(mmcvid4, w.fr_cs.p, n * sizeof(int4), cudaMemcpyDeviceToHost, ctx->stream));
    if (n && lambda) CK(cudaMemcpyAsync(lambda, w.fr_lambda.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (n && coord2) CK(cudaMemcpyAsync(coord2, w.fr_coord.p, n * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    if (n && basis6) CK(cudaMemcpyAsync(basis6, w.fr_basis.p, 6 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return IPCGPU_OK;
}

int ipcgpu_set_friction_data(ipcgpu_ctx* ctx, int n_pairs, const int* mmcvid4, const double* lambda, const double* coord2, const double* basis6)
{
    REQUIRE(ctx->nV > 0, IPCGPU_ERR_STATE, "ipcgpu_set_mesh first");
    REQUIRE(n_pairs >= 0 && n_pairs <= ctx->pair_capacity, IPCGPU_ERR_CAPACITY, "friction set larger than the pair capacity");
    REQUIRE(n_pairs == 0 || (mmcvid4 && lambda && coord2 && basis6), IPCGPU_ERR_ARG, "null friction arrays");
    CK(cudaSetDevice(ctx->device));
    int rc = friction_alloc(ctx);
    if (rc) return rc;
    ContactWork& w = ctx->cw;
    const size_t n = (size_t)n_pairs;
    if (n) {
        CK(cudaMemcpyAsync(w.fr_cs.p, mmcvid4, n * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(w.fr_lambda.p, lambda, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(w.fr_coord.p, coord2, n * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(w.fr_basis.p, basis6, 6 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaMemcpyAsync(w.fr_n.p, &n_pairs, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); // n_pairs lives on the caller's stack
    w.fr_host_n = n_pairs;
    w.fr_ready = true;
    return IPCGPU_OK;
}

static FrictionArgs friction_args(ipcgpu_ctx* ctx, double eps2, double coef, int projectDBC)
{
    FrictionArgs p;
    ContactWork& w = ctx->cw;
    p.nV = ctx->nV; p.V = ctx->V.p; p.Vt = ctx->Vprev.p; p.dbc = ctx->has_dbc ? ctx->dbc.p : nullptr;
    p.cs = w.fr_cs.p; p.n = w.fr_n.p; p.lambda = w.fr_lambda.p; p.coord = w.fr_coord.p; p.basis = w.fr_basis.p;
    p.eps2 = eps2; p.coef = coef; p.projectDBC = projectDBC;
    p.ia = ctx->ia.p; p.ja = ctx->ja.p; p.base = ctx->index_base;
    p.rank = ctx->rank; p.nranks = ctx->nranks;
    p.row_lo = ctx->nranks > 1 ? ctx->v_begin : 0;
    p.row_hi = ctx->nranks > 1 ? ctx->v_end : ctx->nV;
    return p;
}
#define REQUIRE_FRICTION() \
    REQUIRE(ctx->cw.fr_ready, IPCGPU_ERR_STATE, "ipcgpu_friction_lag / ipcgpu_set_friction_data first"); \
    REQUIRE(ctx->prev_set, IPCGPU_ERR_STATE, "ipcgpu_set_prev_state first")

int ipcgpu_friction_energy(ipcgpu_ctx* ctx, double eps2, double coef, double* E)
{
    REQUIRE_FRICTION();
    REQUIRE(eps2 > 0.0, IPCGPU_ERR_ARG, "fricDHat must be positive");
    CK(cudaSetDevice(ctx->device));
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_BARRIER);
    friction_energy(friction_args(ctx, eps2, coef, 0), ctx->cw.fr_partials.p, ctx->stream);
    reduce_sum(ctx->cw.fr_partials.p, friction_energy_blocks(), coef, ctx->scalar_out.p + 2, ctx->stream);
    ctx->prof_end(pe);
    ctx->launches += 2;
    if (ctx->nranks > 1 && E) {
        int r = g_nccl.AllReduce(ctx->scalar_out.p + 2, ctx->scalar_out.p + 2, 1, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
        REQUIRE(r == 0, IPCGPU_ERR_NCCL, "ncclAllReduce(friction energy) failed");
    }
    energy_store(ctx->iter.p, 2, ctx->scalar_out.p + 2, ctx->stream);
    ++ctx->launches;
    ctx->energy_local[2] = (ctx->nranks > 1 && !E);
    CK(cudaGetLastError());
    if (E) {
        CK(cudaMemcpyAsync(ctx->h_scalar, ctx->scalar_out.p + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        *E = ctx->h_scalar[0];
    }
    return IPCGPU_OK;
}

int ipcgpu_friction_gradient(ipcgpu_ctx* ctx, double eps2, double coef, double* g_inout)
{
    REQUIRE_FRICTION();
    REQUIRE(eps2 > 0.0, IPCGPU_ERR_ARG, "fricDHat must be positive");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if (g_inout && (rc = gradient_roundtrip_begin(ctx, g_inout))) return rc;
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_BARRIER);
    friction_gradient(friction_args(ctx, eps2, coef, 0), ctx->g.p, ctx->stream);
    ctx->prof_end(pe);
    ++ctx->launches;
    CK(cudaGetLastError());
    if (g_inout) return gradient_roundtrip_end(ctx, g_inout);
    return IPCGPU_OK;
}

int ipcgpu_friction_hessian(ipcgpu_ctx* ctx, double eps2, double coef, int projectDBC, double* a_inout)
{
    REQUIRE_FRICTION();
    REQUIRE(eps2 > 0.0, IPCGPU_ERR_ARG, "fricDHat must be positive");
    REQUIRE(ctx->nnz > 0, IPCGPU_ERR_STATE, "ipcgpu_set_csr first");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if (a_inout && (rc = upload_values(ctx, a_inout))) return rc;
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_BARRIER);
    friction_hessian(friction_args(ctx, eps2, coef, projectDBC), ctx->a.p, ctx->iter.p->flags + FLAG_PATTERN, ctx->stream);
    ctx->prof_end(pe);
    ++ctx->launches;
    CK(cudaGetLastError());
    if (a_inout) {
        if ((rc = download_values(ctx, a_inout))) return rc;
        if ((rc = fetch_iter_state(ctx))) return rc;
        if (ctx->h_iter->flags[FLAG_PATTERN]) {
            int only[8] = { 0 };
            only[FLAG_PATTERN] = 1;
            clear_flag(ctx, FLAG_PATTERN);
            return status_from_flags(ctx, only);
        }
    }
    return IPCGPU_OK;
}

// ---- inertia term (Optimizer.cpp:3227-3239, :3439-3450) ---------------------------------------------------------------------
int ipcgpu_set_xtilde(ipcgpu_ctx* ctx, const double* xtilde_soa)
{
    REQUIRE(ctx->nV > 0, IPCGPU_ERR_STATE, "ipcgpu_set_mesh first");
    REQUIRE(xtilde_soa != nullptr, IPCGPU_ERR_ARG, "null xTilta");
    CK(cudaSetDevice(ctx->device));
    ALLOC(ctx->xtilde, (size_t)3 * ctx->nV);
    CK(cudaMemcpyAsync(ctx->xtilde.p, xtilde_soa, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->xtilde_set = true;
    return IPCGPU_OK;
}

int ipcgpu_inertia_energy(ipcgpu_ctx* ctx, double* E)
{
    REQUIRE(ctx->xtilde_set && ctx->has_mass, IPCGPU_ERR_STATE, "ipcgpu_set_xtilde and a mass diagonal (ipcgpu_set_mesh) first");
    CK(cudaSetDevice(ctx->device));
    // vertex blocks [nV r / N, nV (r+1) / N): every vertex exactly once across the ranks
    const int v0 = (int)((long long)ctx->nV * ctx->rank / ctx->nranks), v1 = (int)((long long)ctx->nV * (ctx->rank + 1) / ctx->nranks);
    ALLOC(ctx->in_partials, (size_t)inertia_energy_blocks(ctx->nV) + 8);
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_ELASTIC_ENERGY);
    inertia_energy(v0, v1, ctx->nV, ctx->V.p, ctx->xtilde.p, ctx->mass.p, ctx->in_partials.p, ctx->stream);
    reduce_sum(ctx->in_partials.p, inertia_energy_blocks(v1 - v0), 1.0, ctx->scalar_out.p + 3, ctx->stream);
    ctx->prof_end(pe);
    ctx->launches += 2;
    if (ctx->nranks > 1 && E) {
        int r = g_nccl.AllReduce(ctx->scalar_out.p + 3, ctx->scalar_out.p + 3, 1, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
        REQUIRE(r == 0, IPCGPU_ERR_NCCL, "ncclAllReduce(inertia energy) failed");
    }
    energy_store(ctx->iter.p, 3, ctx->scalar_out.p + 3, ctx->stream);
    ++ctx->launches;
    ctx->energy_local[3] = (ctx->nranks > 1 && !E);
    CK(cudaGetLastError());
    if (E) {
        CK(cudaMemcpyAsync(ctx->h_scalar, ctx->scalar_out.p + 3, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        *E = ctx->h_scalar[0];
    }
    return IPCGPU_OK;
}

int ipcgpu_inertia_gradient(ipcgpu_ctx* ctx, int projectDBC, double* g_inout)
{
    REQUIRE(ctx->xtilde_set && ctx->has_mass, IPCGPU_ERR_STATE, "ipcgpu_set_xtilde and a mass diagonal (ipcgpu_set_mesh) first");
    CK(cudaSetDevice(ctx->device));
    if (g_inout) CK(cudaMemcpyAsync(ctx->g.p, g_inout, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // Device-resident form with several ranks: the gradient is summed over the ranks later (ipcgpu_allreduce_grad_hess), so only rank 0 adds
    // the per-vertex term.  Host form: every rank holds the caller's vector and adds the full term -- no reduction needed.
    if (g_inout || ctx->nranks == 1 || ctx->rank == 0) {
        inertia_gradient(ctx->nV, ctx->V.p, ctx->xtilde.p, ctx->mass.p, ctx->has_dbc ? ctx->dbc.p : nullptr, projectDBC, ctx->g.p, ctx->stream);
        ++ctx->launches;
    }
    CK(cudaGetLastError());
    if (g_inout) {
        CK(cudaMemcpyAsync(g_inout, ctx->g.p, (size_t)3 * ctx->nV * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return IPCGPU_OK;
}

// ---- CUDA graphs of device-resident call sequences ----------------------------------------------------------------------------
// An iteration in the NULL-output form is ~65 launches whose arguments do not change while the scene, the pattern, dHat / kappa and
// the tolerances stay the same: positions, search direction, list sizes and step bounds all live in device memory.  Enqueued one by
// one the front of the iteration (grid builds: ~25 kernels of 3-15 us) is bound by the host's launch rate; captured once and
// replayed, the whole sequence is one cudaGraphLaunch.
static ipcgpu_ctx::HostState snapshot_host_state(const ipcgpu_ctx* ctx)
{
    ipcgpu_ctx::HostState h;
    for (int s = 0; s < 4; ++s) h.energy_local[s] = ctx->energy_local[s];
    h.checks_local = ctx->checks_local;
    h.lists_local = ctx->lists_local;
    h.lists_global = ctx->cw.lists_global;
    h.want_cand = ctx->cw.want_cand;
    h.swept_ready = ctx->ccd.swept_ready;
    h.fr_ready = ctx->cw.fr_ready;
    h.inputs_marked = false; // events recorded inside a capture cannot be waited on outside of it
    h.scatter_marked = false;
    h.nC = ctx->cw.nC; h.nP = ctx->cw.nP; h.nK = ctx->cw.nK; h.fr_host_n = ctx->cw.fr_host_n;
    return h;
}
static void apply_host_state(ipcgpu_ctx* ctx, const ipcgpu_ctx::HostState& h)
{
    for (int s = 0; s < 4; ++s) ctx->energy_local[s] = h.energy_local[s];
    ctx->checks_local = h.checks_local;
    ctx->lists_local = h.lists_local;
    ctx->cw.lists_global = h.lists_global;
    ctx->cw.want_cand = h.want_cand;
    ctx->ccd.swept_ready = h.swept_ready;
    ctx->cw.fr_ready = h.fr_ready;
    ctx->inputs_marked = h.inputs_marked;
    ctx->scatter_marked = h.scatter_marked;
    ctx->cw.nC = h.nC; ctx->cw.nP = h.nP; ctx->cw.nK = h.nK; ctx->cw.fr_host_n = h.fr_host_n;
}

int ipcgpu_capture_begin(ipcgpu_ctx* ctx)
{
    REQUIRE(!ctx->capturing, IPCGPU_ERR_STATE, "a capture is already in progress");
    REQUIRE(!ctx->profiling, IPCGPU_ERR_STATE, "switch the stage timers off (ipcgpu_profile(ctx, 0)) before capturing");
    CK(cudaSetDevice(ctx->device));
    ctx->inputs_marked = false;  // the side-stream fork of the pair Hessians must hang on an event recorded INSIDE the capture
    ctx->scatter_marked = false;
    ctx->launches_at_capture = ctx->launches;
    ctx->dirty_at_capture = ctx->a_all_dirty;
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    return IPCGPU_OK;
}

int ipcgpu_capture_end(ipcgpu_ctx* ctx, int* graph_id)
{
    REQUIRE(ctx->capturing, IPCGPU_ERR_STATE, "no capture in progress");
    REQUIRE(graph_id != nullptr, IPCGPU_ERR_ARG, "null graph id");
    {
        int rcj = join_copy_stream(ctx); // a forked copy branch must rejoin the capturing stream
        if (rcj) return rcj;
    }
    ctx->capturing = false;
    ipcgpu_ctx::GraphRec rec;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &rec.graph);
    if (e != cudaSuccess || rec.graph == nullptr) {
        cudaGetLastError();
        ctx->err = std::string("stream capture failed (a call inside the capture synchronised or copied to the host? use NULL outputs): ") + cudaGetErrorString(e);
        return IPCGPU_ERR_CUDA;
    }
    e = cudaGraphInstantiate(&rec.exec, rec.graph, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(rec.graph);
        ctx->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e);
        return IPCGPU_ERR_CUDA;
    }
    rec.launches = ctx->launches - ctx->launches_at_capture;
    rec.epoch = ctx->epoch;
    rec.dirty_at_begin = ctx->dirty_at_capture;
    rec.hs = snapshot_host_state(ctx);
    ctx->launches = ctx->launches_at_capture; // nothing ran yet
    ctx->inputs_marked = false;
    ctx->scatter_marked = false;
    ctx->graphs.push_back(rec);
    *graph_id = (int)ctx->graphs.size() - 1;
    return IPCGPU_OK;
}

int ipcgpu_graph_launch(ipcgpu_ctx* ctx, int graph_id)
{
    REQUIRE(graph_id >= 0 && graph_id < (int)ctx->graphs.size() && ctx->graphs[graph_id].exec, IPCGPU_ERR_ARG, "unknown graph id");
    REQUIRE(!ctx->capturing, IPCGPU_ERR_STATE, "a capture is in progress");
    const ipcgpu_ctx::GraphRec& rec = ctx->graphs[graph_id];
    REQUIRE(rec.epoch == ctx->epoch, IPCGPU_ERR_STATE, "the scene, pattern, partition or capacities changed since this graph was captured: capture it again");
    CK(cudaSetDevice(ctx->device));
    if (ctx->a_all_dirty && !rec.dirty_at_begin) // a cross-rank completion filled rows the captured clear does not cover
        CK(cudaMemsetAsync(ctx->a.p, 0, (size_t)ctx->nnz * sizeof(double), ctx->stream));
    CK(cudaGraphLaunch(rec.exec, ctx->stream));
    ctx->a_all_dirty = false;
    apply_host_state(ctx, rec.hs);
    ctx->launches += rec.launches;
    return IPCGPU_OK;
}

int ipcgpu_graph_destroy(ipcgpu_ctx* ctx, int graph_id)
{
    REQUIRE(graph_id >= 0 && graph_id < (int)ctx->graphs.size(), IPCGPU_ERR_ARG, "unknown graph id");
    ipcgpu_ctx::GraphRec& rec = ctx->graphs[graph_id];
    if (rec.exec) cudaGraphExecDestroy(rec.exec);
    if (rec.graph) cudaGraphDestroy(rec.graph);
    rec.exec = nullptr;
    rec.graph = nullptr;
    return IPCGPU_OK;
}

// ---- line-search safeguards -----------------------------------------------------------------------
static int reduce_checks(ipcgpu_ctx* ctx)
{
    if (ctx->nranks > 1 && ctx->checks_local) {
        int r = g_nccl.AllReduce(ctx->iter.p->checks, ctx->iter.p->checks, 2, kNcclInt32, kNcclSum, ctx->nccl_comm, ctx->stream);
        REQUIRE(r == 0, IPCGPU_ERR_NCCL, "ncclAllReduce(safeguard counts) failed");
    }
    ctx->checks_local = false;
    return IPCGPU_OK;
}

int ipcgpu_check_inversion(ipcgpu_ctx* ctx, int* n_inverted)
{
    REQUIRE(ctx->maps_ready, IPCGPU_ERR_STATE, "ipcgpu_set_mesh first");
    CK(cudaSetDevice(ctx->device));
    int rc = safeguard_inversion(ctx);
    REQUIRE(rc == 0, rc, "inversion check launch failed");
    ctx->checks_local = true;
    if (n_inverted) {
        CK(cudaMemsetAsync(&ctx->iter.p->checks[1], 0, sizeof(int), ctx->stream)); // (the partner count is not pending: keep the sum clean)
        if ((rc = reduce_checks(ctx))) return rc;
        if ((rc = fetch_iter_state(ctx))) return rc;
        *n_inverted = ctx->h_iter->checks[0];
    }
    return IPCGPU_OK;
}

int ipcgpu_intersection_free(ipcgpu_ctx* ctx, int* ok)
{
    REQUIRE(ctx->surface_ready, IPCGPU_ERR_STATE, "ipcgpu_set_surface first");
    CK(cudaSetDevice(ctx->device));
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_HASH);
    int rc = safeguard_intersections(ctx);
    ctx->prof_end(pe);
    REQUIRE(rc == 0, rc, "intersection check launch failed");
    ctx->checks_local = true;
    if (ok) {
        CK(cudaMemsetAsync(&ctx->iter.p->checks[0], 0, sizeof(int), ctx->stream));
        if ((rc = reduce_checks(ctx))) return rc;
        if ((rc = fetch_iter_state(ctx))) return rc;
        *ok = ctx->h_iter->checks[1] == 0 ? 1 : 0;
    }
    return IPCGPU_OK;
}

// ---- device-resident linear solve hand-off ----------------------------------------------------------
int ipcgpu_solve_pcg(ipcgpu_ctx* ctx, const double* rhs, double rel_tol, int max_iter, double* x, int adopt_as_search_dir, int* iters, double* rel_residual)
{
    REQUIRE(ctx->nnz > 0 && ctx->n_rows == 3 * ctx->nV, IPCGPU_ERR_STATE, "ipcgpu_set_csr first");
    REQUIRE(ctx->nranks == 1, IPCGPU_ERR_STATE, "the built-in solver runs on one rank (a distributed solver takes each rank's rows: ipcgpu_partition_info)");
    REQUIRE(rel_tol > 0.0 && max_iter > 0, IPCGPU_ERR_ARG, "bad tolerance / iteration limit");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->full_pattern_ready) { // once per sparsity pattern: rows of both triangles, gathered through a position map
        std::vector<int> ia((size_t)ctx->n_rows + 1), ja((size_t)ctx->nnz);
        CK(cudaMemcpyAsync(ia.data(), ctx->ia.p, ia.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ja.data(), ctx->ja.p, ja.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        int rc = solver_build_full_pattern(ctx, ia.data(), ja.data());
        if (rc) return rc;
    }
    const double* rhs_dev = ctx->g.p;
    double sign = -1.0; // Newton: H p = -g (Optimizer.cpp:2350-2352)
    if (rhs) { // a host right-hand side is staged in a buffer of its own
        ALLOC(ctx->pcg_b, (size_t)ctx->n_rows);
        CK(cudaMemcpyAsync(ctx->pcg_b.p, rhs, (size_t)ctx->n_rows * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        rhs_dev = ctx->pcg_b.p;
        sign = 1.0;
    }
    int rc = solver_pcg(ctx, rhs_dev, sign, rel_tol, max_iter, iters, rel_residual);
    if (rc) return rc;
    if (adopt_as_search_dir && (rc = solver_adopt_direction(ctx))) return rc;
    if (x) {
        CK(cudaMemcpyAsync(x, ctx->sol.p, (size_t)ctx->n_rows * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return IPCGPU_OK;
}

int ipcgpu_csr_set_zero(ipcgpu_ctx* ctx)
{
    REQUIRE(ctx->nnz > 0, IPCGPU_ERR_STATE, "ipcgpu_set_csr first");
    return zero_values(ctx);
}

int ipcgpu_allreduce_grad_hess(ipcgpu_ctx* ctx, int with_gradient, int with_hessian)
{
    if (ctx->nranks <= 1) return IPCGPU_OK;
    REQUIRE(ctx->nccl_comm != nullptr, IPCGPU_ERR_STATE, "ipcgpu_comm_init first");
    cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_ALLREDUCE);
    if (with_gradient) { // elastic part: every rank holds the complete rows it owns (zeros elsewhere); barrier part: partial sums
        int r = g_nccl.AllReduce(ctx->g.p, ctx->g.p, (size_t)3 * ctx->nV, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
        REQUIRE(r == 0, IPCGPU_ERR_NCCL, "ncclAllReduce(gradient) failed");
    }
    if (with_hessian) {
        // The rows a rank owns are already complete (row-owner assembly): this only matters to a caller that wants the WHOLE matrix on
        // every rank.  Non-owned rows are zero, so a sum completes it.
        int r = g_nccl.AllReduce(ctx->a.p, ctx->a.p, (size_t)ctx->nnz, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
        REQUIRE(r == 0, IPCGPU_ERR_NCCL, "ncclAllReduce(csr values) failed");
        ctx->a_all_dirty = true; // the next rebuild must clear every row, not only the owned ones
    }
    ctx->prof_end(pe);
    return IPCGPU_OK;
}

int ipcgpu_fetch_iteration(ipcgpu_ctx* ctx, ipcgpu_iteration* out)
{
    REQUIRE(out != nullptr, IPCGPU_ERR_ARG, "null output");
    CK(cudaSetDevice(ctx->device));
    if (ctx->nranks > 1) {
        // complete the deferred scalars across ranks in ONE collective: locally summed energies, the error flags (so that every rank
        // returns the same status) and the safeguard counts (round 2, first half: up to six separate NCCL calls here)
        unsigned mask = 0;
        for (int s = 0; s < 4; ++s)
            if (ctx->energy_local[s]) mask |= 1u << s;
        if (ctx->checks_local) mask |= 1u << 4;
        cudaEvent_t pe = ctx->prof_begin(IPCGPU_STAGE_ALLREDUCE);
        pack_scalars(ctx->iter.p, mask, ctx->scalar_out.p + 16, ctx->stream);
        int r = g_nccl.AllReduce(ctx->scalar_out.p + 16, ctx->scalar_out.p + 16, 14, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
        unpack_scalars(ctx->iter.p, mask, ctx->scalar_out.p + 16, ctx->stream);
        ctx->prof_end(pe);
        ctx->launches += 2;
        REQUIRE(r == 0, IPCGPU_ERR_NCCL, "ncclAllReduce(iteration scalars) failed");
        for (int s = 0; s < 4; ++s) ctx->energy_local[s] = false;
    }
    ctx->checks_local = false;
    int rc = ccd_read_back(ctx, nullptr);
    if (rc) return rc;
    const IterState& h = *ctx->h_iter;
    out->energy_elastic = h.energy[0];
    out->energy_barrier = h.energy[1];
    out->energy_friction = h.energy[2];
    out->energy_inertia = h.energy[3];
    out->alpha_inversion = h.alpha_stage[0];
    out->alpha_partial_ccd = h.alpha_stage[1];
    out->alpha_swept_grid = h.alpha_stage[2];
    out->alpha_full_ccd = h.alpha_stage[3];
    std::memcpy(&out->alpha, &h.step_ord, sizeof(double));
    out->n_active = h.n_set[0];
    out->n_mollified = h.n_set[1];
    out->n_candidates = h.n_set[2];
    out->n_full_ccd_candidates = h.n_full_cand;
    out->ti_warnings = (uint64_t)h.flags[FLAG_TI_WARNINGS];
    out->n_inverted_tets = h.checks[0];
    out->n_intersected_triangles = h.checks[1];
    ContactWork& w = ctx->cw;
    w.nC = h.n_set[0]; w.nP = h.n_set[1]; w.nK = h.n_set[2];
    const int status = status_from_flags(ctx, h.flags);
    out->status = status;
    CK(cudaMemsetAsync(ctx->iter.p->flags, 0, 8 * sizeof(int), ctx->stream)); // flags are per fetch
    if (h.grid_axis_cells > 0) { // sort width of the next iteration's grid builds: enough bits for 1.5x the cells this one wanted
        int bits = 3;
        while (bits < 10 && (1 << bits) - 2 < (h.grid_axis_cells * 3) / 2 + 2) ++bits;
        w.axis_bits = bits;
        CK(cudaMemsetAsync(&ctx->iter.p->grid_axis_cells, 0, sizeof(int), ctx->stream));
    }
    return status;
}

int ipcgpu_profile(ipcgpu_ctx* ctx, int enable)
{
    REQUIRE(!ctx->capturing, IPCGPU_ERR_STATE, "stage timers cannot be switched inside a capture");
    CK(cudaStreamSynchronize(ctx->stream));
    for (int s = 0; s < IPCGPU_STAGE_COUNT; ++s) {
        for (auto& pr : ctx->prof[s]) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        ctx->prof[s].clear();
    }
    ctx->profiling = enable != 0;
    return IPCGPU_OK;
}

int ipcgpu_profile_read(ipcgpu_ctx* ctx, int stage, double* total_ms, int* count)
{
    REQUIRE(stage >= 0 && stage < IPCGPU_STAGE_COUNT && total_ms && count, IPCGPU_ERR_ARG, "bad stage");
    CK(cudaStreamSynchronize(ctx->stream));
    double tot = 0.0;
    for (auto& pr : ctx->prof[stage]) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, pr.first, pr.second));
        tot += ms;
    }
    *total_ms = tot;
    *count = (int)ctx->prof[stage].size();
    return IPCGPU_OK;
}

int ipcgpu_timer_start(ipcgpu_ctx* ctx)
{
    if (!ctx->timer_a) {
        CK(cudaEventCreate(&ctx->timer_a));
        CK(cudaEventCreate(&ctx->timer_b));
    }
    CK(cudaEventRecord(ctx->timer_a, ctx->stream));
    return IPCGPU_OK;
}

int ipcgpu_timer_stop(ipcgpu_ctx* ctx, double* ms)
{
    REQUIRE(ctx->timer_a && ms, IPCGPU_ERR_STATE, "ipcgpu_timer_start first");
    CK(cudaEventRecord(ctx->timer_b, ctx->stream));
    CK(cudaEventSynchronize(ctx->timer_b));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, ctx->timer_a, ctx->timer_b));
    *ms = f;
    return IPCGPU_OK;
}

static int buf_info(ipcgpu_ctx* ctx, int which, double** p, uint64_t* n)
{
    const uint64_t nL = (uint64_t)ctx->n_list;
    switch (which) {
    case IPCGPU_BUF_GRADIENT: *p = ctx->g.p; *n = (uint64_t)3 * ctx->nV; return 0;
    case IPCGPU_BUF_CSR_VALUES: *p = ctx->a.p; *n = (uint64_t)ctx->nnz; return 0;
    case IPCGPU_BUF_ENERGY_PER_TET: *p = ctx->e_per_tet.p; *n = (uint64_t)ctx->nT; return 0;
    case IPCGPU_BUF_TET_HESSIANS: *p = ctx->hblk.p; *n = 78 * 64 * ((nL + 63) / 64); return ctx->hblk_valid ? 0 : 2; /* tile-major, see elastic.cu */
    case IPCGPU_BUF_TET_GRADIENTS: *p = ctx->gcont.p; *n = 12 * nL; return 0;
    case IPCGPU_BUF_INVERSION_STEPS: *p = ctx->inv_steps.p; *n = (uint64_t)ctx->nT; return 0;
    default: return 1;
    }
}

int ipcgpu_download(ipcgpu_ctx* ctx, int which, double* dst, uint64_t count) { return ipcgpu_download_range(ctx, which, 0, count, dst); }

int ipcgpu_download_range_async(ipcgpu_ctx* ctx, int which, uint64_t offset, uint64_t count, double* dst_pinned)
{
    double* p;
    uint64_t n;
    REQUIRE(buf_info(ctx, which, &p, &n) == 0, IPCGPU_ERR_ARG, "unknown buffer id");
    REQUIRE((dst_pinned || count == 0) && offset + count <= n, IPCGPU_ERR_ARG, "download: bad destination or range");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->copy) {
        REQUIRE(!ctx->capturing, IPCGPU_ERR_STATE, "call ipcgpu_download_range_async once outside a capture first (it creates the copy stream)");
        CK(cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_copy_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_copy_join, cudaEventDisableTiming));
    }
    if (count == 0) return IPCGPU_OK;
    // everything enqueued so far produces the buffer: the copy starts after it and runs next to whatever follows on the main stream
    CK(cudaEventRecord(ctx->ev_copy_fork, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy, ctx->ev_copy_fork, 0));
    CK(cudaMemcpyAsync(dst_pinned, p + offset, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy));
    ctx->copy_pending = true;
    return IPCGPU_OK;
}

int ipcgpu_download_range(ipcgpu_ctx* ctx, int which, uint64_t offset, uint64_t count, double* dst)
{
    double* p;
    uint64_t n;
    {
        const int bi = buf_info(ctx, which, &p, &n);
        REQUIRE(bi != 2, IPCGPU_ERR_STATE, "the per-tet Hessian blocks are only kept by the tile-major layout: ipcgpu_set_hessian_layout(ctx, 0) before the Hessian call");
        REQUIRE(bi == 0, IPCGPU_ERR_ARG, "unknown buffer id");
    }
    REQUIRE((dst || count == 0) && offset + count <= n, IPCGPU_ERR_ARG, "download: bad destination or range");
    if (count) CK(cudaMemcpyAsync(dst, p + offset, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return IPCGPU_OK;
}

void* ipcgpu_device_ptr(ipcgpu_ctx* ctx, int which)
{
    double* p;
    uint64_t n;
    return buf_info(ctx, which, &p, &n) == 0 ? p : nullptr;
}

} // extern "C"
