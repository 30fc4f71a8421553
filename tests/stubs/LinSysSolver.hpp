// IPC::LinSysSolver accessors used by the adapters (src/LinSysSolver/LinSysSolver.hpp:31-37, 451-466).  Test scaffolding only.
#pragma once
#include <Eigen/Eigen>
namespace IPC {
template <typename vectorTypeI, typename vectorTypeS>
class LinSysSolver {
public:
    virtual ~LinSysSolver() = default;
    virtual int getNumRows(void) const;
    virtual Eigen::VectorXi& get_ia(void);
    virtual Eigen::VectorXi& get_ja(void);
    virtual Eigen::VectorXd& get_a(void);
    virtual const Eigen::VectorXd& get_a(void) const;
};
} // namespace IPC
