"""Generates tests/golden/pairs_golden.json by EXECUTING the reference's own codegen (oracle/_ref/libref_pairs.so, built by
oracle/build_ref.sh from /root/reference/src/CollisionObject/MeshCollisionUtils.hpp and src/Utils/BarrierFunctions.hpp)
on the reference-authored stencil (MeshCollisionUtils.hpp:180-182, 635-638, 1234-1238, 2019-2023) and seeded random stencils.
The JSON travels to the GPU box; /root/reference does not.   Run: python tests/golden/gen_pairs_golden.py"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import refpairs as R

rng = np.random.default_rng(20260922)
stencil = np.array([[0, 0, 0], [1, 0.1, 0], [0, 1.1, -0.1], [0, 0.1, -1.1]], dtype=float)
cases = [stencil] + [rng.standard_normal((4, 3)) for _ in range(12)] + [1e-3 * rng.standard_normal((4, 3)) + 5.0 for _ in range(3)]
out = {"cases": [], "barrier": [], "q": []}
for X in cases:
    out["cases"].append({"X": X.tolist(), "g_PE": R.g_PE(X[:3]).tolist(), "H_PE": R.H_PE(X[:3]).tolist(), "g_PT": R.g_PT(X).tolist(),
                         "H_PT": R.H_PT(X).tolist(), "g_EE": R.g_EE(X).tolist(), "H_EE": R.H_EE(X).tolist(),
                         "cross_g": R.EEcross_g(X).tolist(), "cross_H": R.EEcross_H(X).tolist()})
for dHat in (1e-6, 1e-2, 1.0):
    for f in (1e-6, 1e-3, 0.1, 0.5, 0.9, 0.999999):
        b, g, H = R.barrier(f * dHat, dHat)
        out["barrier"].append({"d": f * dHat, "dHat": dHat, "b": b, "g": g, "H": H})
for eps in (1e-3, 7.5):
    for f in (0.0, 0.1, 0.5, 0.99):
        q, qg, qH = R.q(f * eps, eps)
        out["q"].append({"x": f * eps, "eps": eps, "q": q, "qg": qg, "qH": qH})
p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pairs_golden.json")
json.dump(out, open(p, "w"))
print("wrote", p)
