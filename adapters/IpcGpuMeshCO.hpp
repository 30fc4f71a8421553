// IpcGpuMeshCO.hpp -- reference-side binding of the kinematic-obstacle hand-off (include/ipcgpu.h: ipcgpu_set_obstacle_tail).
//
// Compiled inside an ipc-sim/IPC checkout next to IpcGpuAdapters.hpp (which it extends); not built in this repository.
//
// The reference keeps one handler per collision object: MeshCO<3> (src/CollisionObject/MeshCO.hpp:39-233) next to the static
// SelfCollisionHandler<3>, each with its own active set (Optimizer.hpp: MMActiveSet[coI], .back() = self contact) and its own pass over
// the spatial hash.  On the device the obstacle rides at the tail of the mesh's arrays and ONE pass serves both handlers; this header
//   - builds the merged arrays (attach / setState / padded search direction),
//   - splits what the device reports into the two handlers' lists and translates the obstacle's share into MeshCO's own MMCVID
//     encoding (MeshCO.cpp:83-120: negative entries -v-1 are mesh vertices, non-negative entries are obstacle vertices),
//   - gives the obstacle's share of every quantity the reference asks MeshCO for, with MeshCO's signatures.
// Call map (Optimizer.cpp):                                   device work                              who triggers it
//   :2464 / :2466  computeConstraintSet (obstacle, then self)   ipcgpu_constraint_set, once per state    whichever is called first
//   :3270 / :3300  evaluateConstraints                          ipcgpu_evaluate_constraints (merged)     each handler reads its share
//   :3486 / :3500  leftMultiplyConstraintJacobianT              ipcgpu_constraint_jacobian_t             each handler: input scattered to its share
//   :3686 / :3693  augmentIPHessian (+ ParaEE)                  ipcgpu_barrier_hessian, both shares      the self-contact adapter; MeshCO's is a no-op
//   :1899 / :1925  largestFeasibleStepSize_TightInclusion       ipcgpu_ccd_partial_ti, both shares       the self-contact adapter; MeshCO's leaves stepSize
//   :1971 / :2000  largestFeasibleStepSize_CCD_TightInclusion   ipcgpu_ccd_full_ti, both shares          the self-contact adapter; MeshCO's leaves stepSize
//   :2645 / :2650  checkEdgeTriIntersectionIfAny                ipcgpu_intersection_free, both shares    the self-contact adapter; MeshCO's returns true
// (the step bounds are minima and the reference hands the running step from one handler to the next, so "all pairs in one call" and
// "obstacle first, then self" give the same number; the Hessian and the intersection flag are sums / conjunctions.)
// Vectors sized by the vertex count carry the tail: gradient, search direction and solution have 3 * (nV + nVo) entries on the device
// side; padDirection / the prefix copies below convert.  The solver's CSR is the mesh's own: the device pattern appends one identity
// diagonal per obstacle row, and because those rows come last the mesh's values are a PREFIX of the device's value array
// (GpuMeshCO::setPattern / fetchValues).
#pragma once
#include "IpcGpuAdapters.hpp"
#include "MeshCOEncoding.hpp"

namespace IPC {

struct GpuMeshCO {
    inline static IpcGpuScene* gpu = nullptr;
    inline static int nV = 0, nVo = 0, nSE = 0, nSF = 0, nSV = 0; // the MESH's counts (the obstacle's follow them in the merged arrays)
    inline static std::vector<double> Vall;                         // SoA scratch of the merged positions
    // the last merged device result, split: entry i of a handler's list is entry idx_*[i] of the device-resident list
    inline static std::vector<MMCVID> self_set, self_para, co_set, co_para;
    inline static std::vector<std::pair<int, int>> self_para_e, co_para_e, self_cand, co_cand;
    inline static std::vector<int> idx_self, idx_co;
    inline static unsigned long long state_epoch = 0, set_epoch = ~0ull;

    // ---- scene -----------------------------------------------------------------------------------------------------------------
    // MeshCO's Base::V (nVo x 3), Base::F (nFo x 3), edges (MeshCO.cpp:37-80) appended to the mesh's arrays; call INSTEAD of IpcGpuScene::setMesh
    static void attach(IpcGpuScene& scene, const Mesh<3>& mesh, const Eigen::MatrixXd& Vo, const Eigen::MatrixXi& Fo, const std::vector<std::pair<int, int>>& Eo,
        int energyType, bool eeThroughVfRoutine = true)
    {
        gpu = &scene;
        nV = (int)mesh.V.rows(); nVo = (int)Vo.rows(); nSE = (int)mesh.SFEdges.size(); nSF = (int)mesh.SF.rows(); nSV = (int)mesh.SVI.size();
        const int n = nV + nVo, nT = (int)mesh.F.rows(), nFo = (int)Fo.rows(), nEo = (int)Eo.size();
        std::vector<double> rest((size_t)3 * n), mass(n, 0.0), A((size_t)9 * nT);
        std::vector<uint8_t> dbc(n, 1); // the tail is Dirichlet: its rows never reach the system
        std::vector<int> codim(n, 3), svi(nSV + nVo), se((size_t)2 * (nSE + nEo)), sf((size_t)3 * (nSF + nFo));
        Eigen::VectorXd m = mesh.massMatrix.diagonal();
        for (int c = 0; c < 3; ++c) {
            std::copy(mesh.V_rest.data() + (size_t)c * nV, mesh.V_rest.data() + (size_t)(c + 1) * nV, rest.begin() + (size_t)c * n);
            std::copy(Vo.data() + (size_t)c * nVo, Vo.data() + (size_t)(c + 1) * nVo, rest.begin() + (size_t)c * n + nV); // rest = current (MeshCollisionUtils.hpp:2976-2981)
        }
        for (int v = 0; v < nV; ++v) { mass[v] = m.data()[v]; dbc[v] = static_cast<uint8_t>(mesh.vertexDBCType[v]); codim[v] = mesh.vICoDim(v); }
        for (int t = 0; t < nT; ++t) std::copy(mesh.restTriInv[t].data(), mesh.restTriInv[t].data() + 9, A.data() + 9 * t);
        IpcGpuScene::check(scene.ctx, ipcgpu_set_mesh(scene.ctx, n, nT, rest.data(), mesh.F.data(), A.data(), mesh.triArea.data(), mesh.u.data(), mesh.lambda.data(),
                                          mass.data(), dbc.data(), energyType), "ipcgpu_set_mesh");
        for (int i = 0; i < nSV; ++i) svi[i] = mesh.SVI.data()[i];
        for (int i = 0; i < nVo; ++i) svi[nSV + i] = nV + i; // every obstacle vertex is a surface point (MeshCO.cpp:1899)
        for (int e = 0; e < nSE; ++e) { se[2 * e] = mesh.SFEdges[e].first; se[2 * e + 1] = mesh.SFEdges[e].second; }
        for (int e = 0; e < nEo; ++e) { se[2 * (nSE + e)] = nV + Eo[e].first; se[2 * (nSE + e) + 1] = nV + Eo[e].second; }
        for (int c = 0; c < 3; ++c) {
            for (int f = 0; f < nSF; ++f) sf[(size_t)c * (nSF + nFo) + f] = mesh.SF.data()[(size_t)c * nSF + f];
            for (int f = 0; f < nFo; ++f) sf[(size_t)c * (nSF + nFo) + nSF + f] = nV + Fo.data()[(size_t)c * nFo + f];
        }
        IpcGpuScene::check(scene.ctx, ipcgpu_set_surface(scene.ctx, nSV + nVo, svi.data(), nSE + nEo, se.data(), nSF + nFo, sf.data(), codim.data()), "ipcgpu_set_surface");
        IpcGpuScene::check(scene.ctx, ipcgpu_set_obstacle_tail(scene.ctx, nV, eeThroughVfRoutine ? 1 : 0), "ipcgpu_set_obstacle_tail");
        Vall.assign(rest.begin(), rest.end());
        ++state_epoch;
    }
    // after every LinSysSolver::set_pattern: the solver's own pattern (3 nV rows) plus the identity rows of the tail -- upper-triangular 3x3
    // diagonal blocks like every other vertex (row 3v: 3v 3v+1 3v+2, row 3v+1: 3v+1 3v+2, row 3v+2: 3v+2).  No mesh row has a column in the tail.
    inline static int nnz_mesh = 0;
    static void setPattern(LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>* solver, int indexBase)
    {
        const int rows = solver->getNumRows();
        const int* ia = solver->get_ia().data();
        const int* ja = solver->get_ja().data();
        nnz_mesh = ia[rows] - indexBase;
        std::vector<int> ia2(ia, ia + rows + 1), ja2(ja, ja + nnz_mesh);
        for (int v = 0; v < nVo; ++v)
            for (int r = 0; r < 3; ++r) {
                for (int c = r; c < 3; ++c) ja2.push_back(rows + 3 * v + c + indexBase);
                ia2.push_back(ia2.back() + (3 - r));
            }
        IpcGpuScene::check(gpu->ctx, ipcgpu_set_csr(gpu->ctx, rows + 3 * nVo, ia2.data(), ja2.data(), indexBase), "ipcgpu_set_csr");
    }
    // the mesh's values = the first nnz_mesh entries of the device's value array (the tail's identity rows follow): call the assembly entry
    // points with a NULL host array (device-resident accumulation) and fetch the prefix once the matrix is complete
    static void fetchValues(LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>* solver)
    {
        IpcGpuScene::check(gpu->ctx, ipcgpu_download_range(gpu->ctx, IPCGPU_BUF_CSR_VALUES, 0, (uint64_t)nnz_mesh, solver->get_a().data()), "ipcgpu_download_range");
    }
    // ... and the mesh's gradient = the first 3 nV entries of the device's
    static void fetchGradient(Eigen::VectorXd& gradient)
    {
        IpcGpuScene::check(gpu->ctx, ipcgpu_download_range(gpu->ctx, IPCGPU_BUF_GRADIENT, 0, (uint64_t)3 * nV, gradient.data()), "ipcgpu_download_range");
    }
    // whenever mesh.V changes (the obstacle's part stays what it is)
    static void setState(const Mesh<3>& mesh)
    {
        const int n = nV + nVo;
        for (int c = 0; c < 3; ++c) std::copy(mesh.V.data() + (size_t)c * nV, mesh.V.data() + (size_t)(c + 1) * nV, Vall.begin() + (size_t)c * n);
        IpcGpuScene::check(gpu->ctx, ipcgpu_set_state(gpu->ctx, Vall.data()), "ipcgpu_set_state");
        ++state_epoch;
    }
    // MeshCO::move (MeshCO.cpp:3327-3400) after it has updated Base::V
    static void moved(const Eigen::MatrixXd& Vo)
    {
        const int n = nV + nVo;
        for (int c = 0; c < 3; ++c) std::copy(Vo.data() + (size_t)c * nVo, Vo.data() + (size_t)(c + 1) * nVo, Vall.begin() + (size_t)c * n + nV);
        IpcGpuScene::check(gpu->ctx, ipcgpu_set_obstacle_positions(gpu->ctx, Vo.data()), "ipcgpu_set_obstacle_positions");
        ++state_epoch;
    }
    // the search direction with a zero tail (the obstacle does not move during a line search: MeshCO.cpp:790-800 passes Base::V twice)
    static std::vector<double> padDirection(const Eigen::VectorXd& searchDir)
    {
        std::vector<double> p((size_t)3 * (nV + nVo), 0.0);
        std::copy(searchDir.data(), searchDir.data() + (size_t)3 * nV, p.begin());
        return p;
    }

    // ---- encodings (adapters/MeshCOEncoding.hpp: plain functions, run against the oracle's translation by tests/test_adapters_compile.py) ------
    static bool touchesObstacle(const int* q) { return meshco_encoding::touches_obstacle(q, nV); }
    // merged self-contact entry -> MeshCO entry (slot 3 keeps a multiplicity < 0 as it is)
    static MMCVID toMeshCO(const int* q)
    {
        int m[4];
        meshco_encoding::to_meshco(q, nV, m);
        return MMCVID(m[0], m[1], m[2], m[3]);
    }
    // MeshCO entry -> merged self-contact entry (the inverse; MeshCO.cpp:83-120 for which slot holds what)
    static MMCVID toMerged(const int* m)
    {
        int q[4];
        meshco_encoding::to_merged(m, nV, q);
        return MMCVID(q[0], q[1], q[2], q[3]);
    }

    // ---- the one device pass behind both handlers' computeConstraintSet (Optimizer.cpp:2464-2470) -------------------------------------
    static void refreshSets(double dHat, bool getPTEE)
    {
        if (set_epoch == state_epoch) return;
        int nC = 0, nP = 0, nK = 0;
        IpcGpuScene::check(gpu->ctx, ipcgpu_constraint_set(gpu->ctx, dHat, getPTEE, &nC, &nP, &nK), "ipcgpu_constraint_set");
        std::vector<int> mm((size_t)4 * nC + 4), pa((size_t)4 * nP + 4), pe((size_t)2 * nP + 2), cd((size_t)2 * nK + 2);
        IpcGpuScene::check(gpu->ctx, ipcgpu_get_constraint_set(gpu->ctx, mm.data(), pa.data(), pe.data(), getPTEE ? cd.data() : nullptr), "ipcgpu_get_constraint_set");
        self_set.clear(); co_set.clear(); self_para.clear(); co_para.clear(); self_para_e.clear(); co_para_e.clear(); self_cand.clear(); co_cand.clear();
        idx_self.clear(); idx_co.clear();
        for (int c = 0; c < nC; ++c) {
            const int* q = &mm[(size_t)4 * c];
            if (touchesObstacle(q)) { co_set.push_back(toMeshCO(q)); idx_co.push_back(c); }
            else { self_set.push_back(MMCVID(q[0], q[1], q[2], q[3])); idx_self.push_back(c); }
        }
        for (int c = 0; c < nP; ++c) {
            const int* q = &pa[(size_t)4 * c];
            const int eI = pe[(size_t)2 * c], eJ = pe[(size_t)2 * c + 1];
            if (touchesObstacle(q) || eJ >= nSE) { co_para.push_back(toMeshCO(q)); co_para_e.emplace_back(eI, eJ >= 0 ? eJ - nSE : -1); }
            else { self_para.push_back(MMCVID(q[0], q[1], q[2], q[3])); self_para_e.emplace_back(eI, eJ); }
        }
        for (int c = 0; getPTEE && c < nK; ++c) { // cs_PTEE (MeshCO.cpp:2144-2161): PT (-svI-1, sfI), TP (-sfI-1, -vI-1), EE (eI mesh, eJ obstacle)
            const int a = cd[(size_t)2 * c], b = cd[(size_t)2 * c + 1];
            if (a < 0) {
                const int sv = -a - 1;
                if (sv < nSV && b < nSF) self_cand.emplace_back(a, b);
                else if (sv < nSV) co_cand.emplace_back(a, b - nSF);
                else co_cand.emplace_back(-b - 1, -(sv - nSV) - 1);
            }
            else if (b < nSE) self_cand.emplace_back(a, b);
            else co_cand.emplace_back(a, b - nSE);
        }
        set_epoch = state_epoch;
    }
    // MeshCO<3>::computeConstraintSet (MeshCO.hpp:150-156)
    static void computeConstraintSet(const Mesh<3>&, const SpatialHash<3>&, double dHat, std::vector<MMCVID>& constraintSet, std::vector<MMCVID>& paraEEMMCVIDSet,
        std::vector<std::pair<int, int>>& paraEEeIeJSet, bool getPTEE, std::vector<std::pair<int, int>>& cs_PTEE)
    {
        refreshSets(dHat, getPTEE);
        constraintSet = co_set; paraEEMMCVIDSet = co_para; paraEEeIeJSet = co_para_e;
        if (getPTEE) cs_PTEE = co_cand;
    }
    // SelfCollisionHandler<3>::computeConstraintSet when an obstacle is attached (replaces GpuSelfCollisionHandler's)
    static void computeSelfConstraintSet(const Mesh<3>&, const SpatialHash<3>&, double dHat, std::vector<MMCVID>& constraintSet, std::vector<MMCVID>& paraEEMMCVIDSet,
        std::vector<std::pair<int, int>>& paraEEeIeJSet, bool getPTEE, std::vector<std::pair<int, int>>& cs_PTEE)
    {
        refreshSets(dHat, getPTEE);
        constraintSet = self_set; paraEEMMCVIDSet = self_para; paraEEeIeJSet = self_para_e;
        if (getPTEE) cs_PTEE = self_cand;
    }
    // CollisionObject::evaluateConstraints over MeshCO::evaluateConstraint (MeshCO.cpp:83-120): appends this handler's distances.
    // (`activeSet` must be the list computeConstraintSet returned for the current state: the device evaluates its resident list)
    static void evaluateConstraints(const Mesh<3>&, const std::vector<MMCVID>& activeSet, Eigen::VectorXd& val, double /*coef*/ = 1.0, bool self = false)
    {
        const std::vector<int>& idx = self ? idx_self : idx_co;
        if (activeSet.size() != idx.size()) { spdlog::error("GpuMeshCO::evaluateConstraints: the list is not the one the device holds"); exit(-1); }
        std::vector<double> all(idx_self.size() + idx_co.size() + 1);
        IpcGpuScene::check(gpu->ctx, ipcgpu_evaluate_constraints(gpu->ctx, all.data(), (int)(idx_self.size() + idx_co.size())), "ipcgpu_evaluate_constraints");
        const int start = (int)val.size();
        val.conservativeResize(start + (int)idx.size());
        for (size_t i = 0; i < idx.size(); ++i) val.data()[start + i] = all[idx[i]];
    }
    // MeshCO<3>::leftMultiplyConstraintJacobianT (MeshCO.cpp:122-200): out (3 nV of the mesh) += coef * mult * input[c] * grad d_c, mesh rows only
    static void leftMultiplyConstraintJacobianT(const Mesh<3>&, const std::vector<MMCVID>& activeSet, const Eigen::VectorXd& input, Eigen::VectorXd& output_incremental,
        double coef = 1.0, bool self = false)
    {
        const std::vector<int>& idx = self ? idx_self : idx_co;
        if (activeSet.size() != idx.size()) { spdlog::error("GpuMeshCO::leftMultiplyConstraintJacobianT: the list is not the one the device holds"); exit(-1); }
        const size_t nAll = idx_self.size() + idx_co.size();
        std::vector<double> in(nAll + 1, 0.0), out((size_t)3 * (nV + nVo), 0.0); // the other handler's entries contribute nothing
        for (size_t i = 0; i < idx.size(); ++i) in[idx[i]] = input.data()[i];
        IpcGpuScene::check(gpu->ctx, ipcgpu_constraint_jacobian_t(gpu->ctx, in.data(), (int)nAll, coef, out.data()), "ipcgpu_constraint_jacobian_t");
        for (size_t i = 0; i < (size_t)3 * nV; ++i) output_incremental.data()[i] += out[i]; // the tail's rows are Dirichlet rows: dropped
    }
    // MeshCO<3>::augmentIPHessian / augmentParaEEHessian (MeshCO.cpp:407-586, :2314-2520): added by the self-contact adapter's
    // ipcgpu_barrier_hessian, which walks the merged list (both handlers' pairs, mesh rows and columns only)
    static void augmentIPHessian(const Mesh<3>&, const std::vector<MMCVID>&, LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>*, double, double = 1.0, bool = true) {}
    static void augmentParaEEHessian(const Mesh<3>&, const std::vector<MMCVID>&, const std::vector<std::pair<int, int>>&, LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>*, double,
        double, bool)
    {
    }
    // MeshCO<3>::largestFeasibleStepSize_TightInclusion / _CCD_TightInclusion (MeshCO.cpp:742-980, :1388-1668): the self-contact adapter's
    // ipcgpu_ccd_partial_ti / ipcgpu_ccd_full_ti take the minimum over both handlers' pairs (hand them padDirection(searchDir))
    static void largestFeasibleStepSize_TightInclusion(const Mesh<3>&, const SpatialHash<3>&, const Eigen::VectorXd&, double, const std::vector<std::pair<int, int>>&, double&) {}
    static void largestFeasibleStepSize_CCD_TightInclusion(const Mesh<3>&, const SpatialHash<3>&, const Eigen::VectorXd&, double, double&) {}
    // MeshCO<3>::checkEdgeTriIntersectionIfAny (MeshCO.cpp:2611-2678): part of the self-contact adapter's ipcgpu_intersection_free
    static bool checkEdgeTriIntersectionIfAny(const Mesh<3>&, const SpatialHash<3>&) { return true; }
};

} // namespace IPC
