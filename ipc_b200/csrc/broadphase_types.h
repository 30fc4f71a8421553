// broadphase_types.h -- POD types shared by host (context.h) and device code (broadphase.cuh)
#pragma once
namespace ipcgpu {
struct Grid {
    double ox, oy, oz, inv_h;
    int nx, ny, nz;
};
struct Box {
    double lo[3], hi[3];
};
} // namespace ipcgpu
