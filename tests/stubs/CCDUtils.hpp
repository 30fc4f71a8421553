// Tight-Inclusion numerical-error globals (src/Utils/CCDUtils.hpp:24, CCDUtils.cpp:16).  Test scaffolding only.
#pragma once
#include <array>
namespace IPC {
extern std::array<double, 3> tight_inclusion_vf_err, tight_inclusion_ee_err;
}
