// Declarations of the IPC::Mesh<dim> members the adapters read, restated from the reference for a compile check
// (src/Mesh.hpp:61-149: V_rest, V, F, SF :63-67; SVI :68; massMatrix :83; u, lambda :85; triArea :86; vertexDBCType :89;
// restTriInv :92; SFEdges :98; vICoDim :149).  Test scaffolding only.
#pragma once
#include <Eigen/Eigen>
#include <utility>
#include <vector>
namespace IPC {
enum class DirichletBCType { NOT_DBC, ZERO, NONZERO };
template <int dim>
class Mesh {
public:
    Eigen::MatrixXd V_rest, V;
    Eigen::MatrixXi F, SF;
    Eigen::VectorXi SVI;
    Eigen::SparseMatrix<double> massMatrix;
    Eigen::VectorXd u, lambda;
    Eigen::VectorXd triArea;
    std::vector<DirichletBCType> vertexDBCType;
    std::vector<Eigen::Matrix<double, dim, dim>> restTriInv;
    std::vector<std::pair<int, int>> SFEdges;
    int vICoDim(int vI) const;
};
} // namespace IPC
