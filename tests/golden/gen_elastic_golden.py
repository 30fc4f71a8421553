"""Generates tests/golden/elastic_golden.json with mpmath (50 digits) from the closed forms the reference
implements (NeoHookeanEnergy.cpp:55-153, FixedCoRotEnergy.cpp:62-153), independent of any SVD code:
  NH : psi = mu/2 (|F|^2 - 3) - mu ln J + lam/2 (ln J)^2 ;  P = mu (F - F^-T) + lam ln J F^-T
  FCR: psi = mu |F - R|^2 + lam/2 (J-1)^2 (R = polar rotation) ; P = 2 mu (F - R) + lam (J-1) cof(F)
dP/dF (unprojected, row-major vec(F) index 3i+j as in Energy.cpp:535-542) by high-precision central differences.
Material of the reference's unit tests: YM=100, PR=0.4 (Energy.cpp:588).
Run:  python tests/golden/gen_elastic_golden.py
"""
import json, os
import mpmath as mp

mp.mp.dps = 50
YM, PR = mp.mpf(100), mp.mpf("0.4")
MU = YM / 2 / (1 + PR)
LAM = YM * PR / (1 + PR) / (1 - 2 * PR)

def cof(F):
    return mp.matrix([[F[1,1]*F[2,2]-F[1,2]*F[2,1], F[1,2]*F[2,0]-F[1,0]*F[2,2], F[1,0]*F[2,1]-F[1,1]*F[2,0]],
                      [F[0,2]*F[2,1]-F[0,1]*F[2,2], F[0,0]*F[2,2]-F[0,2]*F[2,0], F[0,1]*F[2,0]-F[0,0]*F[2,1]],
                      [F[0,1]*F[1,2]-F[0,2]*F[1,1], F[0,2]*F[1,0]-F[0,0]*F[1,2], F[0,0]*F[1,1]-F[0,1]*F[1,0]]])

def polar_R(F):
    # Newton iteration R <- (R + R^-T)/2 converges to the orthogonal polar factor; for det F<0 it converges to the
    # improper factor, so flip the smallest singular direction the way the rotation-variant SVD does (sigma_3<0).
    U, S, Vt = mp.svd_r(F)
    if mp.det(U) < 0:
        U[:, 2] = -U[:, 2]; S[2] = -S[2]
    if mp.det(Vt) < 0:
        Vt[2, :] = -Vt[2, :]; S[2] = -S[2]
    return U * Vt, S

def psi(et, F):
    J = mp.det(F)
    if et == 0:
        return MU/2*(sum(F[i,j]**2 for i in range(3) for j in range(3)) - 3) - MU*mp.log(J) + LAM/2*mp.log(J)**2
    R, _ = polar_R(F)
    D = F - R
    return MU*sum(D[i,j]**2 for i in range(3) for j in range(3)) + LAM/2*(J-1)**2

def P(et, F):
    J = mp.det(F); C = cof(F)
    if et == 0:
        FinvT = C / J
        return MU*(F - FinvT) + LAM*mp.log(J)*FinvT
    R, _ = polar_R(F)
    return 2*MU*(F - R) + LAM*(J-1)*C

def dPdF(et, F, h=mp.mpf(10)**-18):
    out = [[None]*9 for _ in range(9)]
    for r in range(3):
        for s in range(3):
            Fp = F.copy(); Fm = F.copy(); Fp[r,s] += h; Fm[r,s] -= h
            D = (P(et, Fp) - P(et, Fm)) / (2*h)
            for i in range(3):
                for j in range(3):
                    out[3*i+j][3*r+s] = D[i,j]
    return out

CASES = {
    "identity": [[1,0,0],[0,1,0],[0,0,1]],
    "stretch": [["1.3","0.1","-0.2"],["0.05","0.9","0.15"],["-0.1","0.2","1.1"]],
    "compress": [["0.6","0.05","0.0"],["0.1","0.7","-0.05"],["0.02","0.03","0.5"]],
    "shear_rot": [["0.8","-0.6","0.3"],["0.55","0.85","-0.1"],["0.05","0.2","1.2"]],
    "near_degenerate": [["1.0","1e-9","0"],["0","1.0","1e-9"],["0","0","1.0"]],
}
INVERTED = {"inverted": [["0.9","0.1","0.0"],["0.05","1.1","0.1"],["0.0","0.2","-0.4"]]}  # FCR only

def f(x): return float(x)
res = {"mu": f(MU), "lam": f(LAM), "cases": []}
for name, Fl in list(CASES.items()) + list(INVERTED.items()):
    F = mp.matrix([[mp.mpf(str(x)) for x in row] for row in Fl])
    for et in (0, 1):
        if name == "inverted" and et == 0:
            continue
        Pm = P(et, F)
        H = dPdF(et, F)
        _, S = polar_R(F)
        res["cases"].append({"name": name, "energy": et, "F": [[f(F[i,j]) for j in range(3)] for i in range(3)],
                             "psi": f(psi(et, F)), "P": [[f(Pm[i,j]) for j in range(3)] for i in range(3)],
                             "dPdF": [[f(H[i][j]) for j in range(9)] for i in range(9)],
                             "sigma": sorted([f(abs(s)) for s in S], reverse=True), "detF": f(mp.det(F))})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "elastic_golden.json")
json.dump(res, open(out, "w"), indent=0)
print("wrote", out, len(res["cases"]), "cases")
