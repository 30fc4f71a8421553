// IPC::MMCVID layout (src/CollisionObject/MeshCollisionUtils.hpp:24-40).  Test scaffolding only.
#pragma once
#include <array>
namespace IPC {
class MMCVID {
public:
    std::array<int, 4> data;
    MMCVID(int a, int b, int c, int d) : data{ a, b, c, d } {}
    MMCVID(int a = -1) : data{ a, -1, -1, -1 } {}
};
} // namespace IPC
