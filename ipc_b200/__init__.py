"""ipc_b200 -- B200-native (sm_100a) IPC Newton hot path behind IPC's own plug-in interfaces.

csrc/   hand-written CUDA kernels + the extern "C" ABI (include/ipcgpu.h) -> libipcgpu.so
lib.py  ctypes binding of that ABI
mesh.py host-side scene precompute (what Mesh<3>::computeFeatures / LinSysSolver::set_pattern produce)
"""
__all__ = ["lib", "mesh"]
