"""The reference-side binding (adapters/IpcGpuAdapters.hpp) cannot be built against the real reference here (no Eigen/TBB/libigl), so it
is compile-checked against an interface-only restatement of the reference classes it plugs into (tests/stubs/, each file citing the
reference lines it restates): wrong `override` signatures, misspelt members or ABI misuse fail here."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_adapters_compile_against_the_restated_interface():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Wno-unused-parameter", "-I", os.path.join(ROOT, "tests", "stubs"),
           "-I", os.path.join(ROOT, "adapters"), "-I", ROOT, os.path.join(ROOT, "tests", "stubs", "adapter_check.cpp")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_meshco_encoding_header_runs_against_the_oracle(tmp_path):
    """adapters/MeshCOEncoding.hpp (what GpuMeshCO uses to split the device's merged lists) is plain C++: build it into a small shared object,
    run it on random entries of all six kinds and compare with the oracle's translation and with the Python mirror (ipc_b200/obstacle.py)."""
    import ctypes as C
    import sys

    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as orc
    from ipc_b200 import obstacle as OB
    src = tmp_path / "enc.cpp"
    src.write_text('#include "MeshCOEncoding.hpp"\n'
                   'extern "C" void to_merged(const int* m, int nV, int* q) { meshco_encoding::to_merged(m, nV, q); }\n'
                   'extern "C" void to_meshco(const int* q, int nV, int* m) { meshco_encoding::to_meshco(q, nV, m); }\n'
                   'extern "C" int touches(const int* q, int nV) { return meshco_encoding::touches_obstacle(q, nV) ? 1 : 0; }\n')
    so = tmp_path / "libenc.so"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-shared", "-fPIC", "-I", os.path.join(ROOT, "adapters"), str(src), "-o", str(so)])
    enc = C.CDLL(str(so))
    rng = np.random.default_rng(0)
    nV, nVo, nSE = 1000, 300, 400
    mv = lambda n=1: rng.integers(0, nV, n)
    ov = lambda n=1: rng.integers(0, nVo, n)
    rows = []
    for _ in range(200):
        mult = -int(rng.integers(1, 5))
        rows += [[*mv(2), *ov(2)], [-mv()[0] - 1, ov()[0], -1, mult], [-mv()[0] - 1, *ov(2), mult], [-mv()[0] - 1, *ov(3)],
                 [*(-mv(3) - 1), ov()[0]], [*(-mv(2) - 1), ov()[0], mult]]
    co = np.array(rows, dtype=np.int32)
    merged = np.empty_like(co)
    for k in range(len(co)):
        enc.to_merged(orc.i(co[k]), nV, orc.i(merged[k]))
    ref = np.empty_like(co)
    pe = np.zeros((0, 2), np.int32)
    orc.lib().orc_meshco_to_merged(nV, nSE, orc.i(co), len(co), orc.i(ref), orc.i(pe), 0, orc.i(pe))
    assert np.array_equal(merged, ref)
    back = np.empty_like(co)
    for k in range(len(co)):
        assert enc.touches(orc.i(merged[k]), nV) == 1 and OB.involves_obstacle(merged[k], nV)
        enc.to_meshco(orc.i(merged[k]), nV, orc.i(back[k]))
        assert list(back[k]) == OB.merged_to_meshco(merged[k], nV)
    assert np.array_equal(back, co)
    own = np.array([[-5, 6, 7, 8], [1, 2, 3, 4], [-5, 6, -1, -2]], dtype=np.int32)
    assert all(enc.touches(orc.i(q), nV) == 0 for q in own)
