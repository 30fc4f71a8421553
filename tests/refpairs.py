"""ctypes wrapper of oracle/_ref/libref_pairs.so: the REFERENCE's own scalar codegen bodies (MeshCollisionUtils.hpp,
BarrierFunctions.hpp) compiled from /root/reference by oracle/build_ref.sh.  Used to pin the oracle restatement and
to generate tests/golden/pairs_golden.json.  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_pairs.so")
_dp = C.POINTER(C.c_double)


def available():
    return os.path.exists(PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(PATH)
    return _lib


def _call(name, v, n):
    v = np.ascontiguousarray(v, dtype=np.float64).ravel()
    out = np.empty(n)
    getattr(lib(), name)(v.ctypes.data_as(_dp), out.ctypes.data_as(_dp))
    return out


def g_PE(v): return _call("ref_g_PE", v, 9)
def H_PE(v): return _call("ref_H_PE", v, 81).reshape(9, 9)
def g_PT(v): return _call("ref_g_PT", v, 12)
def H_PT(v): return _call("ref_H_PT", v, 144).reshape(12, 12)
def g_EE(v): return _call("ref_g_EE", v, 12)
def H_EE(v): return _call("ref_H_EE", v, 144).reshape(12, 12)
def EEcross_g(v): return _call("ref_EEcross_g", v, 12)
def EEcross_H(v): return _call("ref_EEcross_H", v, 144).reshape(12, 12)


def barrier(d, dHat):
    b, g, H = C.c_double(), C.c_double(), C.c_double()
    lib().ref_barrier(C.c_double(d), C.c_double(dHat), C.byref(b), C.byref(g), C.byref(H))
    return b.value, g.value, H.value


def q(x, eps):
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    lib().ref_q(C.c_double(x), C.c_double(eps), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value
