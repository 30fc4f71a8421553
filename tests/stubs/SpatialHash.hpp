// IPC::SpatialHash<dim> is only passed through by the adapters (src/Utils/SpatialHash.hpp:21-22).  Test scaffolding only.
#pragma once
namespace IPC {
template <int dim>
class SpatialHash {};
} // namespace IPC
