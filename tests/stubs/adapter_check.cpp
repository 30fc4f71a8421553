// Instantiates every adapter against the restated reference interface (syntax/semantic check only: g++ -fsyntax-only).
#include "IpcGpuAdapters.hpp"
#include "IpcGpuMeshCO.hpp"
template class IPC::GpuElasticEnergy<3>;
void touch(IPC::IpcGpuScene& s, const IPC::Mesh<3>& m, IPC::LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>* sol, const IPC::SpatialHash<3>& sh)
{
    s.setMesh(m, IPCGPU_NEOHOOKEAN);
    s.setPattern(sol, 1);
    s.setState(m);
    std::vector<IPC::MMCVID> a, b;
    std::vector<std::pair<int, int>> c, d;
    Eigen::VectorXd v;
    double x = 1.0;
    using H = IPC::GpuSelfCollisionHandler;
    H::computeConstraintSet(m, sh, 1e-6, a, b, c, true, d);
    // the reference's own call sequence (Optimizer.cpp:3492-3502, 3693-3695), reference signatures (SelfCollisionHandler.hpp:23-33)
    H::evaluateConstraints(m, a, v);
    H::leftMultiplyConstraintJacobianT(m, a, v, v, 1e8);
    H::augmentParaEEGradient(m, b, c, v, 1e-6, 1e8);
    H::augmentIPHessian(m, a, sol, 1e-6, 1e8, true);
    H::augmentParaEEHessian(m, b, c, sol, 1e-6, 1e8, true);
    // friction (SelfCollisionHandler.hpp:182-208; Optimizer.cpp:1594, 3371, 3505, 3699)
    std::vector<Eigen::Vector2d> co;
    std::vector<Eigen::Matrix<double, 3, 2>> ba;
    Eigen::MatrixXd Vt;
    H::computeDistCoordAndTanBasis(m, a, co, ba);
    H::computeDistCoordAndTanBasis(m, a, co, ba, 1e-6, 1e8, &v);
    H::computeFrictionEnergy(m.V, Vt, a, v, co, ba, x, 1e-9, 0.3);
    H::augmentFrictionGradient(m.V, Vt, a, v, co, ba, v, 1e-9, 0.3);
    H::augmentFrictionHessian(m, Vt, a, v, co, ba, sol, 1e-9, 0.3, true);
    (void)H::checkEdgeTriIntersectionIfAny(m, sh);
    (void)IPC::gpuCheckInversion(s);
    H::largestFeasibleStepSize_TightInclusion(m, sh, v, 1e-6, d, d, x);
    H::hashBuildSwept(v, x, 0.1);
    H::largestFeasibleStepSize_CCD_TightInclusion(m, sh, v, 1e-6, d, x);
    // kinematic obstacle: MeshCO's signatures (MeshCO.hpp:56-181) over the merged device pass
    using O = IPC::GpuMeshCO;
    Eigen::MatrixXd Vo;
    Eigen::MatrixXi Fo;
    O::attach(s, m, Vo, Fo, c, IPCGPU_NEOHOOKEAN);
    O::setPattern(sol, 1);
    O::setState(m);
    O::moved(Vo);
    std::vector<double> pd = O::padDirection(v);
    O::computeConstraintSet(m, sh, 1e-6, a, b, c, true, d);
    O::computeSelfConstraintSet(m, sh, 1e-6, a, b, c, true, d);
    O::evaluateConstraints(m, a, v);
    O::leftMultiplyConstraintJacobianT(m, a, v, v, 1e8);
    O::augmentIPHessian(m, a, sol, 1e-6, 1e8, true);
    O::augmentParaEEHessian(m, b, c, sol, 1e-6, 1e8, true);
    O::largestFeasibleStepSize_TightInclusion(m, sh, v, 1e-6, d, x);
    O::largestFeasibleStepSize_CCD_TightInclusion(m, sh, v, 1e-6, x);
    (void)O::checkEdgeTriIntersectionIfAny(m, sh);
    O::fetchValues(sol);
    O::fetchGradient(v);
    int q[4] = { 0, 1, 2, 3 };
    (void)O::toMerged(O::toMeshCO(q).data.data());
    (void)pd;
}
