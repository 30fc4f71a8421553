// The virtual interface of IPC::Energy<dim> that GpuElasticEnergy overrides (src/Energy/Energy.hpp:27-62, 118-120): if a signature
// here drifted from the adapter's, `override` fails to compile.  Test scaffolding only.
#pragma once
#include "LinSysSolver.hpp"
#include "Mesh.hpp"
#include <vector>
namespace IPC {
template <typename MatType>
class AutoFlipSVD; // src/Utils/AutoFlipSVD.hpp
template <int dim>
class Energy {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
protected:
    const bool needElemInvSafeGuard;

public:
    Energy(bool p_needElemInvSafeGuard);
    virtual ~Energy(void);
    virtual void computeEnergyVal(const Mesh<dim>& data, int redoSVD, std::vector<AutoFlipSVD<Eigen::Matrix<double, dim, dim>>>& svd,
        std::vector<Eigen::Matrix<double, dim, dim>>& F, double coef, double& energyVal) const;
    virtual void computeGradient(const Mesh<dim>& data, bool redoSVD, std::vector<AutoFlipSVD<Eigen::Matrix<double, dim, dim>>>& svd,
        std::vector<Eigen::Matrix<double, dim, dim>>& F, double coef, Eigen::VectorXd& gradient, bool projectDBC = true) const;
    virtual void computeHessian(const Mesh<dim>& data, bool redoSVD, std::vector<AutoFlipSVD<Eigen::Matrix<double, dim, dim>>>& svd,
        std::vector<Eigen::Matrix<double, dim, dim>>& F, double coef, LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>* linSysSolver,
        bool projectSPD = true, bool projectDBC = true) const;
    virtual void filterStepSize(const Mesh<dim>& data, const Eigen::VectorXd& searchDir, double& stepSize) const;
};
} // namespace IPC
