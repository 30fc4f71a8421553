// IpcGpuAdapters.hpp -- reference-side binding of libipcgpu.so (include/ipcgpu.h).
//
// This header is compiled INSIDE an ipc-sim/IPC checkout (it includes the reference's own headers: Eigen, Mesh.hpp,
// Energy.hpp, LinSysSolver.hpp ...), next to the files it replaces; it is NOT built in this repository, where Eigen /
// TBB / libigl are absent.  It shows exactly how the reference's plug-in interfaces map onto the
// C ABI, so that `Optimizer` drives the GPU path unchanged:
//
//   main.cpp:1383-1392   energyTerms.emplace_back(new IPC::GpuElasticEnergy<DIM>(gpu, /*NeoHookean*/0));
//   SelfCollisionHandler.cpp  -> the static functions below replace the bodies of the same names
//   Optimizer.cpp:1965        sh.build(result, searchDir, alpha, h)  -> IpcGpuScene::hashBuildSwept
//
// Everything is a thin marshalling layer: Eigen objects already have the memory layout the ABI expects
// (column-major MatrixXd/MatrixXi == SoA, VectorXd gradient == interleaved), so no copies are made on the host side.
#pragma once
#include "../include/ipcgpu.h"

#include "Energy.hpp"          // IPC::Energy<dim>                      (src/Energy/Energy.hpp:38-131)
#include "LinSysSolver.hpp"    // IPC::LinSysSolver                     (src/LinSysSolver/LinSysSolver.hpp:34-37)
#include "Mesh.hpp"            // IPC::Mesh<dim>                        (src/Mesh.hpp:61-144)
#include "MeshCollisionUtils.hpp" // IPC::MMCVID                        (src/CollisionObject/MeshCollisionUtils.hpp:24-110)
#include "SpatialHash.hpp"     // IPC::SpatialHash<dim>                 (src/Utils/SpatialHash.hpp:21-22)
#include "CCDUtils.hpp"        // tight_inclusion_{vf,ee}_err           (src/Utils/CCDUtils.hpp:24)

#include <spdlog/spdlog.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <vector>

namespace IPC {

// One per simulation: owns the device context and mirrors the scene arrays once (Mesh::computeFeatures output).
class IpcGpuScene {
public:
    ipcgpu_ctx* ctx = nullptr;

    explicit IpcGpuScene(int device = 0)
    {
        if (ipcgpu_create(device, &ctx) != IPCGPU_OK) throw std::runtime_error("ipcgpu_create failed: no CUDA device (there is no CPU fallback)");
    }
    ~IpcGpuScene() { ipcgpu_destroy(ctx); }

    static void check(ipcgpu_ctx* c, int rc, const char* what)
    {
        if (rc == IPCGPU_OK) return;
        // the reference's own convention on invariant violations is spdlog::error + exit (Optimizer.cpp:2031-2033, 3296-3306);
        // the ABI never exits, the adapter decides
        spdlog::error("{}: {}", what, ipcgpu_last_error(c));
        if (rc == IPCGPU_ERR_NONPOSITIVE_DISTANCE) exit(0);
        exit(-1);
    }

    // after Mesh::computeFeatures / setLameParam and AnimScripter::initAnimScript (Optimizer.cpp:170)
    void setMesh(const Mesh<3>& mesh, int energyType)
    {
        std::vector<double> A(9 * mesh.F.rows());
        for (int t = 0; t < mesh.F.rows(); ++t) std::copy(mesh.restTriInv[t].data(), mesh.restTriInv[t].data() + 9, A.data() + 9 * t);
        std::vector<uint8_t> dbc(mesh.V.rows());
        for (int v = 0; v < mesh.V.rows(); ++v) dbc[v] = static_cast<uint8_t>(mesh.vertexDBCType[v]); // NOT_DBC=0, ZERO=1, NONZERO=2 (Mesh.hpp:135-144)
        Eigen::VectorXd mass = mesh.massMatrix.diagonal();
        check(ctx, ipcgpu_set_mesh(ctx, (int)mesh.V.rows(), (int)mesh.F.rows(), mesh.V_rest.data(), mesh.F.data(), A.data(), mesh.triArea.data(),
                       mesh.u.data(), mesh.lambda.data(), mass.data(), dbc.data(), energyType), "ipcgpu_set_mesh");
        std::vector<int> se(2 * mesh.SFEdges.size()), codim(mesh.V.rows());
        for (size_t e = 0; e < mesh.SFEdges.size(); ++e) { se[2 * e] = mesh.SFEdges[e].first; se[2 * e + 1] = mesh.SFEdges[e].second; }
        for (int v = 0; v < mesh.V.rows(); ++v) codim[v] = mesh.vICoDim(v);
        check(ctx, ipcgpu_set_surface(ctx, (int)mesh.SVI.size(), mesh.SVI.data(), (int)mesh.SFEdges.size(), se.data(), (int)mesh.SF.rows(), mesh.SF.data(), codim.data()),
            "ipcgpu_set_surface");
    }
    // after every LinSysSolver::set_pattern (Optimizer.cpp:457-507, 3556-3595)
    void setPattern(LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>* solver, int indexBase /* 1, or 0 after CHOLMODSolver::set_pattern */)
    {
        check(ctx, ipcgpu_set_csr(ctx, solver->getNumRows(), solver->get_ia().data(), solver->get_ja().data(), indexBase), "ipcgpu_set_csr");
    }
    // whenever mesh.V changes (stepForward, Optimizer.cpp:2919-2938)
    void setState(const Mesh<3>& mesh) { check(ctx, ipcgpu_set_state(ctx, mesh.V.data()), "ipcgpu_set_state"); }
};

// Energy<3> plug-in (replaces NeoHookeanEnergy / FixedCoRotEnergy objects created at main.cpp:1383-1392)
template <int dim>
class GpuElasticEnergy : public Energy<dim> {
    IpcGpuScene& gpu;

public:
    GpuElasticEnergy(IpcGpuScene& scene, int energyType)
        : Energy<dim>(/*needElemInvSafeGuard=*/energyType == IPCGPU_NEOHOOKEAN)
        , gpu(scene)
    {
    }
    // Energy.hpp:42-47.  redoSVD is accepted for interface fidelity; F/SVD are recomputed on the device (DESIGN.md "SVD cache")
    void computeEnergyVal(const Mesh<dim>& data, int redoSVD, std::vector<AutoFlipSVD<Eigen::Matrix<double, dim, dim>>>&,
        std::vector<Eigen::Matrix<double, dim, dim>>&, double coef, double& energyVal) const override
    {
        IpcGpuScene::check(gpu.ctx, ipcgpu_elastic_energy(gpu.ctx, coef, redoSVD, &energyVal), "ipcgpu_elastic_energy");
    }
    // Energy.hpp:56-62
    void computeGradient(const Mesh<dim>& data, bool redoSVD, std::vector<AutoFlipSVD<Eigen::Matrix<double, dim, dim>>>&,
        std::vector<Eigen::Matrix<double, dim, dim>>&, double coef, Eigen::VectorXd& gradient, bool projectDBC = true) const override
    {
        gradient.conservativeResize(data.V.rows() * dim); // Energy.cpp:267
        IpcGpuScene::check(gpu.ctx, ipcgpu_elastic_gradient(gpu.ctx, coef, redoSVD, projectDBC, gradient.data()), "ipcgpu_elastic_gradient");
    }
    // Energy.hpp:64-71: the sink's value array is written wholesale (LinSysSolver::get_a() is mutable, LinSysSolver.hpp:463-466)
    void computeHessian(const Mesh<dim>& data, bool redoSVD, std::vector<AutoFlipSVD<Eigen::Matrix<double, dim, dim>>>&,
        std::vector<Eigen::Matrix<double, dim, dim>>&, double coef, LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>* sink, bool projectSPD = true,
        bool projectDBC = true) const override
    {
        IpcGpuScene::check(gpu.ctx, ipcgpu_elastic_hessian(gpu.ctx, coef, redoSVD, projectSPD, projectDBC, sink->get_a().data()), "ipcgpu_elastic_hessian");
    }
    // Energy.hpp:118-120
    void filterStepSize(const Mesh<dim>& data, const Eigen::VectorXd& searchDir, double& stepSize) const override
    {
        if (this->needElemInvSafeGuard) IpcGpuScene::check(gpu.ctx, ipcgpu_inversion_step(gpu.ctx, searchDir.data(), 0.2, &stepSize), "ipcgpu_inversion_step");
    }
};

// SelfCollisionHandler<3> statics (SelfCollisionHandler.hpp:23-232): same names, same SIGNATURES, same argument meaning -- the bodies of the
// reference's functions are replaced one for one, Optimizer.cpp is not edited.  The device keeps the sets it built last; every function that
// the reference hands an `activeSet` checks (size + 64-bit FNV hash of the ints) that it is the device-resident one and uploads it when it
// is not (ipcgpu_set_constraint_set), so a caller that edits or swaps the vectors between calls is still served correctly.
struct GpuSetTag { // size + hash of a constraint list
    size_t n;
    uint64_t h;
};
struct GpuSelfCollisionHandler {
    inline static IpcGpuScene* gpu = nullptr; // set once by main()
    using SetTag = GpuSetTag;
    inline static SetTag active_tag{ ~size_t(0), 0 }, para_tag{ ~size_t(0), 0 }; // what the device holds
    // host copies of the sets the device holds (the reference hands the active set and the mollified sets to DIFFERENT functions, the
    // upload takes all of them)
    inline static std::vector<MMCVID> last_active, last_para;
    inline static std::vector<std::pair<int, int>> last_para_e;

    static uint64_t fnv(const int* p, size_t n)
    {
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < n; ++i) { h ^= (uint32_t)p[i]; h *= 1099511628211ull; }
        return h;
    }
    static SetTag tag_of(const std::vector<MMCVID>& s) { return { s.size(), fnv(reinterpret_cast<const int*>(s.data()), 4 * s.size()) }; }
    // make the device-resident sets equal to what the caller holds
    static void ensureSets(const std::vector<MMCVID>& activeSet, const std::vector<MMCVID>& para, const std::vector<std::pair<int, int>>& paraE)
    {
        const SetTag ta = tag_of(activeSet), tp = tag_of(para);
        if (ta.n == active_tag.n && ta.h == active_tag.h && tp.n == para_tag.n && tp.h == para_tag.h) return;
        IpcGpuScene::check(gpu->ctx, ipcgpu_set_constraint_set(gpu->ctx, (int)activeSet.size(), reinterpret_cast<const int*>(activeSet.data()), (int)para.size(),
                                         reinterpret_cast<const int*>(para.data()), reinterpret_cast<const int*>(paraE.data()), 0, nullptr),
            "ipcgpu_set_constraint_set");
        active_tag = ta;
        para_tag = tp;
        last_active = activeSet;
        last_para = para;
        last_para_e = paraE;
    }

    // :2149-2478. `sh` is unused: the broad phase lives on the device.
    static void computeConstraintSet(const Mesh<3>& mesh, const SpatialHash<3>& /*sh*/, double dHat, std::vector<MMCVID>& constraintSet,
        std::vector<MMCVID>& paraEEMMCVIDSet, std::vector<std::pair<int, int>>& paraEEeIeJSet, bool getPTEE, std::vector<std::pair<int, int>>& cs_PTEE)
    {
        int nC = 0, nP = 0, nK = 0;
        IpcGpuScene::check(gpu->ctx, ipcgpu_constraint_set(gpu->ctx, dHat, getPTEE, &nC, &nP, &nK), "ipcgpu_constraint_set");
        constraintSet.assign(nC, MMCVID());
        paraEEMMCVIDSet.assign(nP, MMCVID());
        paraEEeIeJSet.assign(nP, {});
        if (getPTEE) cs_PTEE.assign(nK, {});
        static_assert(sizeof(MMCVID) == 4 * sizeof(int), "MMCVID is std::array<int,4>");
        static_assert(sizeof(std::pair<int, int>) == 2 * sizeof(int), "pair<int,int> is two ints");
        IpcGpuScene::check(gpu->ctx, ipcgpu_get_constraint_set(gpu->ctx, reinterpret_cast<int*>(constraintSet.data()), reinterpret_cast<int*>(paraEEMMCVIDSet.data()),
                                         reinterpret_cast<int*>(paraEEeIeJSet.data()), getPTEE ? reinterpret_cast<int*>(cs_PTEE.data()) : nullptr),
            "ipcgpu_get_constraint_set");
        active_tag = tag_of(constraintSet);
        para_tag = tag_of(paraEEMMCVIDSet);
        last_active = constraintSet;
        last_para = paraEEMMCVIDSet;
        last_para_e = paraEEeIeJSet;
    }
    // :64-81  (appends, like the reference: Optimizer.cpp:3493 relies on it)
    static void evaluateConstraints(const Mesh<3>&, const std::vector<MMCVID>& activeSet, Eigen::VectorXd& val, double /*coef*/ = 1.0)
    {
        ensureSets(activeSet, last_para, last_para_e);
        const int start = (int)val.size();
        val.conservativeResize(start + activeSet.size());
        IpcGpuScene::check(gpu->ctx, ipcgpu_evaluate_constraints(gpu->ctx, val.data() + start, (int)activeSet.size()), "ipcgpu_evaluate_constraints");
    }
    // :84-148  out += coef * mult * input[c] * grad d_c
    static void leftMultiplyConstraintJacobianT(const Mesh<3>&, const std::vector<MMCVID>& activeSet, const Eigen::VectorXd& input, Eigen::VectorXd& output_incremental,
        double coef = 1.0)
    {
        ensureSets(activeSet, last_para, last_para_e);
        IpcGpuScene::check(gpu->ctx, ipcgpu_constraint_jacobian_t(gpu->ctx, input.data(), (int)activeSet.size(), coef, output_incremental.data()),
            "ipcgpu_constraint_jacobian_t");
    }
    // :2990-3045
    static void augmentParaEEGradient(const Mesh<3>&, const std::vector<MMCVID>& paraEEMMCVIDSet, const std::vector<std::pair<int, int>>& paraEEeIeJSet,
        Eigen::VectorXd& grad_inc, double dHat, double coef)
    {
        ensureSets(std::vector<MMCVID>(last_active), paraEEMMCVIDSet, paraEEeIeJSet);
        IpcGpuScene::check(gpu->ctx, ipcgpu_para_ee_gradient(gpu->ctx, dHat, coef, grad_inc.data()), "ipcgpu_para_ee_gradient");
    }
    // :418-561 (reference signature).  The device call also adds the mollified pairs' Hessian (augmentParaEEHessian, :3049-3201), so that
    // function is a no-op below.
    static void augmentIPHessian(const Mesh<3>&, const std::vector<MMCVID>& activeSet, LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>* mtr, double dHat, double coef,
        bool projectDBC)
    {
        ensureSets(activeSet, last_para, last_para_e);
        IpcGpuScene::check(gpu->ctx, ipcgpu_barrier_hessian(gpu->ctx, dHat, coef, projectDBC, mtr->get_a().data()), "ipcgpu_barrier_hessian");
    }
    static void augmentParaEEHessian(const Mesh<3>&, const std::vector<MMCVID>&, const std::vector<std::pair<int, int>>&, LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>*, double,
        double, bool)
    {
        // already added by augmentIPHessian (one device pass over both lists)
    }
    // :690-866
    static void largestFeasibleStepSize_TightInclusion(const Mesh<3>&, const SpatialHash<3>&, const Eigen::VectorXd& searchDir, double tolerance,
        const std::vector<std::pair<int, int>>&, std::vector<std::pair<int, int>>&, double& stepSize)
    {
        IpcGpuScene::check(gpu->ctx, ipcgpu_ccd_partial_ti(gpu->ctx, searchDir.data(), tolerance, tight_inclusion_vf_err.data(), tight_inclusion_ee_err.data(), &stepSize),
            "ipcgpu_ccd_partial_ti");
    }
    // SpatialHash::build(mesh, searchDir, curMaxStepSize, voxelSize) at Optimizer.cpp:1965 (alpha is mutated, SpatialHash.hpp:603-618)
    static void hashBuildSwept(const Eigen::VectorXd& searchDir, double& alpha, double voxelSize)
    {
        IpcGpuScene::check(gpu->ctx, ipcgpu_hash_build_swept(gpu->ctx, searchDir.data(), &alpha, voxelSize), "ipcgpu_hash_build_swept");
    }
    // :1370-1630
    static void largestFeasibleStepSize_CCD_TightInclusion(const Mesh<3>&, const SpatialHash<3>&, const Eigen::VectorXd&, double tolerance,
        std::vector<std::pair<int, int>>&, double& stepSize)
    {
        IpcGpuScene::check(gpu->ctx, ipcgpu_ccd_full_ti(gpu->ctx, tolerance, tight_inclusion_vf_err.data(), tight_inclusion_ee_err.data(), &stepSize, nullptr),
            "ipcgpu_ccd_full_ti");
    }
    // ---- lagged friction (reference signatures, SelfCollisionHandler.hpp:182-208).  The reference hands V / Vt / the lagged containers to
    // every call; the adapter uploads what it is handed (positions every call, the lagged containers when their tag changes) ------------
    inline static SetTag fric_tag{ ~size_t(0), 0 };
    static void ensureFriction(const std::vector<MMCVID>& constraintSet, const Eigen::VectorXd& multipliers, const std::vector<Eigen::Vector2d>& MMDistCoord,
        const std::vector<Eigen::Matrix<double, 3, 2>>& MMTanBasis)
    {
        static_assert(sizeof(Eigen::Vector2d) == 2 * sizeof(double) && sizeof(Eigen::Matrix<double, 3, 2>) == 6 * sizeof(double), "fixed-size Eigen types are plain arrays");
        SetTag t = tag_of(constraintSet);
        t.h ^= fnv(reinterpret_cast<const int*>(multipliers.data()), 2 * constraintSet.size()) * 31u;
        t.h ^= fnv(reinterpret_cast<const int*>(MMDistCoord.data()), 4 * constraintSet.size()) * 131u;
        t.h ^= fnv(reinterpret_cast<const int*>(MMTanBasis.data()), 12 * constraintSet.size()) * 1031u;
        if (t.n == fric_tag.n && t.h == fric_tag.h) return;
        IpcGpuScene::check(gpu->ctx, ipcgpu_set_friction_data(gpu->ctx, (int)constraintSet.size(), reinterpret_cast<const int*>(constraintSet.data()), multipliers.data(),
                                          reinterpret_cast<const double*>(MMDistCoord.data()), reinterpret_cast<const double*>(MMTanBasis.data())),
            "ipcgpu_set_friction_data");
        fric_tag = t;
    }
    // :2481-2527 -- computed on the device at the current device state, copied into the caller's containers.  The multipliers of
    // Optimizer.cpp:1582-1591 come out of the same kernel (gpuFrictionMultipliers below), so the host loop there can go.
    static void computeDistCoordAndTanBasis(const Mesh<3>& mesh, const std::vector<MMCVID>& constraintSet, std::vector<Eigen::Vector2d>& MMDistCoord,
        std::vector<Eigen::Matrix<double, 3, 2>>& MMTanBasis, double dHat = 1.0, double kappa = 0.0, Eigen::VectorXd* multipliers = nullptr)
    {
        ensureSets(constraintSet, last_para, last_para_e);
        IpcGpuScene::check(gpu->ctx, ipcgpu_set_state(gpu->ctx, mesh.V.data()), "ipcgpu_set_state");
        int n = 0;
        IpcGpuScene::check(gpu->ctx, ipcgpu_friction_lag(gpu->ctx, dHat, kappa, &n), "ipcgpu_friction_lag");
        MMDistCoord.resize(constraintSet.size());
        MMTanBasis.resize(constraintSet.size());
        if (multipliers) multipliers->conservativeResize(constraintSet.size());
        IpcGpuScene::check(gpu->ctx, ipcgpu_get_friction_data(gpu->ctx, &n, nullptr, multipliers ? multipliers->data() : nullptr, reinterpret_cast<double*>(MMDistCoord.data()),
                                          reinterpret_cast<double*>(MMTanBasis.data())),
            "ipcgpu_get_friction_data");
        fric_tag = SetTag{ ~size_t(0), 0 }; // the caller may post-process its containers: re-validate at the next evaluator call
    }
    // :2529-2596
    static void computeFrictionEnergy(const Eigen::MatrixXd& V, const Eigen::MatrixXd& Vt, const std::vector<MMCVID>& constraintSet, const Eigen::VectorXd& multipliers,
        const std::vector<Eigen::Vector2d>& MMDistCoord, const std::vector<Eigen::Matrix<double, 3, 2>>& MMTanBasis, double& Ef, double eps2, double coef)
    {
        ensureFriction(constraintSet, multipliers, MMDistCoord, MMTanBasis);
        IpcGpuScene::check(gpu->ctx, ipcgpu_set_state(gpu->ctx, V.data()), "ipcgpu_set_state");
        IpcGpuScene::check(gpu->ctx, ipcgpu_set_prev_state(gpu->ctx, Vt.data()), "ipcgpu_set_prev_state");
        IpcGpuScene::check(gpu->ctx, ipcgpu_friction_energy(gpu->ctx, eps2, coef, &Ef), "ipcgpu_friction_energy");
    }
    // :2598-2735
    static void augmentFrictionGradient(const Eigen::MatrixXd& V, const Eigen::MatrixXd& Vt, const std::vector<MMCVID>& constraintSet, const Eigen::VectorXd& multipliers,
        const std::vector<Eigen::Vector2d>& MMDistCoord, const std::vector<Eigen::Matrix<double, 3, 2>>& MMTanBasis, Eigen::VectorXd& grad_inc, double eps2, double coef)
    {
        ensureFriction(constraintSet, multipliers, MMDistCoord, MMTanBasis);
        IpcGpuScene::check(gpu->ctx, ipcgpu_set_state(gpu->ctx, V.data()), "ipcgpu_set_state");
        IpcGpuScene::check(gpu->ctx, ipcgpu_set_prev_state(gpu->ctx, Vt.data()), "ipcgpu_set_prev_state");
        IpcGpuScene::check(gpu->ctx, ipcgpu_friction_gradient(gpu->ctx, eps2, coef, grad_inc.data()), "ipcgpu_friction_gradient");
    }
    // :2708-2743 (the LinSysSolver overload; the std::function overload of :2745 exists for the SQP path, which is out of scope)
    static void augmentFrictionHessian(const Mesh<3>& mesh, const Eigen::MatrixXd& Vt, const std::vector<MMCVID>& constraintSet, const Eigen::VectorXd& multipliers,
        const std::vector<Eigen::Vector2d>& MMDistCoord, const std::vector<Eigen::Matrix<double, 3, 2>>& MMTanBasis, LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>* H_inc,
        double eps2, double coef, bool projectDBC)
    {
        ensureFriction(constraintSet, multipliers, MMDistCoord, MMTanBasis);
        IpcGpuScene::check(gpu->ctx, ipcgpu_set_state(gpu->ctx, mesh.V.data()), "ipcgpu_set_state");
        IpcGpuScene::check(gpu->ctx, ipcgpu_set_prev_state(gpu->ctx, Vt.data()), "ipcgpu_set_prev_state");
        IpcGpuScene::check(gpu->ctx, ipcgpu_friction_hessian(gpu->ctx, eps2, coef, projectDBC, H_inc->get_a().data()), "ipcgpu_friction_hessian");
    }
    // :3254-3340 (edge-triangle part; see include/ipcgpu.h for the point-in-tetrahedron remark)
    static bool checkEdgeTriIntersectionIfAny(const Mesh<3>&, const SpatialHash<3>&)
    {
        int ok = 0;
        IpcGpuScene::check(gpu->ctx, ipcgpu_intersection_free(gpu->ctx, &ok), "ipcgpu_intersection_free");
        return ok != 0;
    }
};

// Mesh<3>::checkInversion(bool mute) (Mesh.cpp:745-763) on the device state
inline bool gpuCheckInversion(IpcGpuScene& gpu)
{
    int n = 0;
    IpcGpuScene::check(gpu.ctx, ipcgpu_check_inversion(gpu.ctx, &n), "ipcgpu_check_inversion");
    return n == 0;
}

} // namespace IPC
