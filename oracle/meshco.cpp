/* CPU ORACLE (test infrastructure only -- the product never links or calls this file).
 *
 * Kinematic mesh obstacles: the reference's MeshCO<3> (src/CollisionObject/MeshCO.cpp), restated for the barrier / Tight-Inclusion
 * path of SURVEY.md section 8 row f3.  An obstacle is a triangle mesh without degrees of freedom (Base::V, edges, Base::F); it meets the
 * simulated mesh through the same point-triangle / edge-edge machinery as self contact, with three differences this file keeps:
 *   - the active set uses MeshCO's own MMCVID encoding (sign = body: negative entries -v-1 are MESH vertices, non-negative entries are
 *     OBSTACLE vertices; MeshCO.cpp:83-120):
 *         EE (m0, m1, o0, o1)   PP (-m-1, o, -1, -mult)   PE (-m-1, o0, o1, -mult)   PT (-m-1, o0, o1, o2)
 *         TP (-m0-1, -m1-1, -m2-1, o)   EP (-m0-1, -m1-1, o, -mult)
 *     and a mesh-vertex / obstacle-vertex PP pair found from the mesh point's side (:1827-1848) and from the obstacle point's side
 *     (:1916-1935) is ONE entry whose multiplicity counts both;
 *   - no filters: no shared-vertex test (the bodies share nothing), no Dirichlet / codimension test (:1795-2100);
 *   - gradient and Hessian are taken of the FULL pair stencil (the obstacle's vertices included: makePD projects the full 6x6 / 9x9 /
 *     12x12 block, :430-560) and only the mesh vertices' rows and columns reach the system;
 *   - Tight-Inclusion: edge-edge pairs go through vertexFaceCCD_double with the edge-edge error bound (:900-940, :1609-1655 -- the
 *     reference calls the vertex-face routine there; ee_as_vf = 1 reproduces it, 0 calls the edge-edge routine), and in the partial CCD
 *     the retry of an edge-edge pair assigns its result to a shadowing local (:929), so the pair counts as colliding with whatever time
 *     the retry left.
 * The pair-level arithmetic (distances, derivatives, barrier, mollifier, makePD, Tight-Inclusion) is the oracle's own, shared with the
 * self-contact restatement: MeshCO.cpp calls the same MeshCollisionUtils / IglUtils / inclusion_ccd functions as SelfCollisionHandler.cpp.
 * E / g / H are evaluated by translating MeshCO entries to the self-contact encoding over a merged vertex numbering (obstacle vertex k =
 * nV + k, Dirichlet, rest position = current position: compute_eps_x(mesh, Base::V, ...), MeshCollisionUtils.hpp:2976-2981) and calling
 * the self-contact oracle; argument orders of every pair function are preserved by the translation (see co_to_merged).
 * PARITY: no reference test covers MeshCO; pair math is pinned through oracle/contact.cpp, Tight-Inclusion stays "parity unpinned".
 */
#include "oracle.h"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

namespace {

typedef std::array<int, 4> Q;

inline void mvert(const orc_surf* s, int v, double* x) { x[0] = s->V[v]; x[1] = s->V[(size_t)s->nV + v]; x[2] = s->V[(size_t)2 * s->nV + v]; }
inline void overt(const orc_obstacle* o, int v, double* x) { x[0] = o->V[v]; x[1] = o->V[(size_t)o->nV + v]; x[2] = o->V[(size_t)2 * o->nV + v]; }
inline double sq(double a) { return a * a; }

struct Box { double lo[3], hi[3]; };
inline Box box_of(const double* x, int n, double pad)
{
    Box b;
    for (int c = 0; c < 3; ++c) {
        b.lo[c] = b.hi[c] = x[c];
        for (int k = 1; k < n; ++k) { b.lo[c] = std::min(b.lo[c], x[3 * k + c]); b.hi[c] = std::max(b.hi[c], x[3 * k + c]); }
        b.lo[c] -= pad; b.hi[c] += pad;
    }
    return b;
}
inline bool overlap(const Box& a, const Box& b)
{
    for (int c = 0; c < 3; ++c)
        if (a.hi[c] < b.lo[c] || b.hi[c] < a.lo[c]) return false;
    return true;
}

/* merged numbering: mesh vertices 0..nV-1, obstacle vertex k -> nV + k */
struct Merged {
    std::vector<double> V, Vrest;
    std::vector<uint8_t> dbc;
    std::vector<int> SE;
    orc_surf s;
};
void merge(const orc_surf* s, const orc_obstacle* o, Merged& m)
{
    const int n = s->nV + o->nV;
    m.V.resize((size_t)3 * n); m.Vrest.resize((size_t)3 * n); m.dbc.assign(n, 1);
    for (int c = 0; c < 3; ++c) {
        for (int v = 0; v < s->nV; ++v) {
            m.V[(size_t)c * n + v] = s->V[(size_t)c * s->nV + v];
            m.Vrest[(size_t)c * n + v] = s->Vrest[(size_t)c * s->nV + v];
        }
        for (int v = 0; v < o->nV; ++v) m.V[(size_t)c * n + s->nV + v] = m.Vrest[(size_t)c * n + s->nV + v] = o->V[(size_t)c * o->nV + v];
    }
    for (int v = 0; v < s->nV; ++v) m.dbc[v] = s->dbc ? s->dbc[v] : 0;
    m.SE.resize((size_t)2 * (s->nSE + o->nE));
    for (int e = 0; e < 2 * s->nSE; ++e) m.SE[e] = s->SE[e];
    for (int e = 0; e < 2 * o->nE; ++e) m.SE[2 * (size_t)s->nSE + e] = s->nV + o->E[e];
    m.s = *s;
    m.s.nV = n; m.s.V = m.V.data(); m.s.Vrest = m.Vrest.data(); m.s.dbc = m.dbc.data();
    m.s.nSE = s->nSE + o->nE; m.s.SE = m.SE.data();
    m.s.vCoDim = nullptr;
}
/* MeshCO entry -> self-contact entry over the merged numbering.  The pair functions see their arguments in the same order:
 *   TP: d_PT(obstacle point, mesh triangle) (:111) = PT entry with the obstacle vertex as the point;  EP likewise (:115) */
Q co_to_merged(const int mm[4], int nV)
{
    if (mm[0] >= 0) return { mm[0], mm[1], nV + mm[2], mm[3] >= 0 ? nV + mm[3] : mm[3] };
    if (mm[1] >= 0) return { mm[0], nV + mm[1], mm[2] >= 0 ? nV + mm[2] : mm[2], mm[3] >= 0 ? nV + mm[3] : mm[3] };
    if (mm[2] < 0) return { -(nV + mm[3]) - 1, -mm[0] - 1, -mm[1] - 1, -mm[2] - 1 };
    return { -(nV + mm[2]) - 1, -mm[0] - 1, -mm[1] - 1, mm[3] };
}
void translate(const orc_surf* s, const orc_obstacle* o, const int* mmcvid, int nC, const int* para, const int* para_e, int nP, std::vector<int>& act, std::vector<int>& par,
    std::vector<int>& parE)
{
    act.resize((size_t)4 * std::max(nC, 1)); par.resize((size_t)4 * std::max(nP, 1)); parE.resize((size_t)2 * std::max(nP, 1));
    for (int c = 0; c < nC; ++c) { Q q = co_to_merged(mmcvid + 4 * c, s->nV); std::memcpy(&act[4 * (size_t)c], q.data(), 16); }
    for (int c = 0; c < nP; ++c) {
        Q q = co_to_merged(para + 4 * c, s->nV);
        std::memcpy(&par[4 * (size_t)c], q.data(), 16);
        parE[2 * (size_t)c] = para_e[2 * c];                                              /* mesh edge, or -1 */
        parE[2 * (size_t)c + 1] = para_e[2 * c + 1] < 0 ? -1 : s->nSE + para_e[2 * c + 1]; /* obstacle edge */
    }
}

/* one candidate of the step bound: points in the order MeshCO passes them to inclusion_ccd; kind 0 = PT (mesh point, obstacle
 * triangle), 1 = TP (obstacle point, mesh triangle), 2 = EE (mesh edge, obstacle edge).  Returns -1 zero distance, 0 / 1 (toi) */
int co_pair_ccd(int kind, const double* x0, const double* x1, double tol, const double evf[3], const double eee[3], double max_t, int ee_as_vf, bool partial, double* toi_out)
{
    double d;
    if (kind == 2) orc_edge_edge_d(x0, &d);
    else orc_point_tri_d(x0, &d);
    d = std::sqrt(d);
    if (d == 0) return -1;
    const bool vf_routine = kind != 2 || ee_as_vf;
    const double* err = kind == 2 ? eee : evf;
    auto run = [&](double ms, int max_itr, double* toi) {
        double out_tol;
        return vf_routine ? orc_ti_vf(x0, x1, err, ms, tol, max_t, max_itr, 1, toi, &out_tol) : orc_ti_ee(x0, x1, err, ms, tol, max_t, max_itr, 1, toi, &out_tol);
    };
    double toi;
    int hit = run(std::min(0.2 * d, 1e-6), 1000000, &toi);
    if (hit && toi < 1e-6) {
        const int again = run(0.0, -1, &toi);
        if (again) toi *= 0.8;
        if (!(kind == 2 && partial)) hit = again; /* :929 -- `bool has_collision = ...` shadows the outer flag in the partial CCD's EE branch */
    }
    if (hit) { *toi_out = toi; return 1; }
    return 0;
}

} // namespace

extern "C" {

/* MeshCO<3>::computeConstraintSet (MeshCO.cpp:1795-2223), brute force with a conservative box reject (the set does not depend on the
 * spatial hash).  cand = cs_PTEE (:2144-2161): PT (-svI-1, sfI), TP (-sfI-1, -vI-1), EE (eI mesh, eJ obstacle).  Outputs sorted. */
int orc_meshco_constraint_set(const orc_surf* s, const orc_obstacle* o, double dHat, int cap, int* mmcvid, int* nC, int capP, int* para, int* para_eIeJ, int* nPara, int capK,
    int* cand, int* nCand, int nthreads)
{
    const double pad = std::sqrt(dHat);
    const int nth = nthreads > 0 ? nthreads : 1;
    std::vector<std::vector<Q>> csPT(o->nF), csTP(o->nV), csEE(o->nE);
    std::vector<std::vector<int>> cPT(o->nF), cTP(o->nV), cEE(o->nE);
    auto PPq = [](int m, int ov, int last) { return Q{ -m - 1, ov, -1, last }; };
    /* point-triangle: obstacle triangle x mesh surface vertex (:1803-1895) */
#pragma omp parallel for num_threads(nth) schedule(dynamic, 4)
    for (int sfI = 0; sfI < o->nF; ++sfI) {
        const int t[3] = { o->F[sfI], o->F[(size_t)o->nF + sfI], o->F[(size_t)2 * o->nF + sfI] };
        double x[12];
        for (int k = 0; k < 3; ++k) overt(o, t[k], x + 3 + 3 * k);
        const Box tb = box_of(x + 3, 3, pad);
        for (int svI = 0; svI < s->nSV; ++svI) {
            const int vI = s->SVI[svI];
            mvert(s, vI, x);
            if (!overlap(tb, box_of(x, 1, 0.0))) continue;
            const int dtype = orc_dType_PT(x);
            double d, y[9];
            Q q;
            if (dtype <= 2) {
                std::memcpy(y, x, 24); std::memcpy(y + 3, x + 3 * (dtype + 1), 24); orc_d_PP(y, &d);
                q = PPq(vI, t[dtype], -1);
            }
            else if (dtype <= 5) {
                const int a = dtype - 3, b = (dtype - 2) % 3;
                std::memcpy(y, x, 24); std::memcpy(y + 3, x + 3 * (a + 1), 24); std::memcpy(y + 6, x + 3 * (b + 1), 24); orc_d_PE(y, &d);
                q = { -vI - 1, t[a], t[b], -1 };
            }
            else {
                orc_d_PT(x, &d);
                q = { -vI - 1, t[0], t[1], t[2] };
            }
            if (d < dHat) { csPT[sfI].push_back(q); cPT[sfI].push_back(svI); }
        }
    }
    /* triangle-point: obstacle vertex x mesh surface triangle (:1897-1990) */
#pragma omp parallel for num_threads(nth) schedule(dynamic, 4)
    for (int vI = 0; vI < o->nV; ++vI) {
        double x[12];
        overt(o, vI, x);
        const Box pb = box_of(x, 1, pad);
        for (int sfI = 0; sfI < s->nSF; ++sfI) {
            const int t[3] = { s->SF[sfI], s->SF[(size_t)s->nSF + sfI], s->SF[(size_t)2 * s->nSF + sfI] };
            for (int k = 0; k < 3; ++k) mvert(s, t[k], x + 3 + 3 * k);
            if (!overlap(pb, box_of(x + 3, 3, 0.0))) continue;
            const int dtype = orc_dType_PT(x);
            double d, y[9];
            Q q;
            if (dtype <= 2) {
                std::memcpy(y, x, 24); std::memcpy(y + 3, x + 3 * (dtype + 1), 24); orc_d_PP(y, &d);
                q = PPq(t[dtype], vI, -1);
            }
            else if (dtype <= 5) {
                const int a = dtype - 3, b = (dtype - 2) % 3;
                std::memcpy(y, x, 24); std::memcpy(y + 3, x + 3 * (a + 1), 24); std::memcpy(y + 6, x + 3 * (b + 1), 24); orc_d_PE(y, &d);
                q = { -t[a] - 1, -t[b] - 1, vI, -1 };
            }
            else {
                orc_d_PT(x, &d);
                q = { -t[0] - 1, -t[1] - 1, -t[2] - 1, vI };
            }
            if (d < dHat) { csTP[vI].push_back(q); cTP[vI].push_back(sfI); }
        }
    }
    /* edge-edge: obstacle edge eJ x mesh surface edge eI (:1996-2140) */
#pragma omp parallel for num_threads(nth) schedule(dynamic, 4)
    for (int eJ = 0; eJ < o->nE; ++eJ) {
        const int b0 = o->E[2 * eJ], b1 = o->E[2 * eJ + 1];
        double x[12];
        overt(o, b0, x + 6); overt(o, b1, x + 9);
        const Box eb = box_of(x + 6, 2, pad);
        const double lenJ = sq(x[6] - x[9]) + sq(x[7] - x[10]) + sq(x[8] - x[11]); /* (V_MCO.row(eJ0) - V_MCO.row(eJ1)).squaredNorm() */
        for (int eI = 0; eI < s->nSE; ++eI) {
            const int a0 = s->SE[2 * eI], a1 = s->SE[2 * eI + 1];
            mvert(s, a0, x); mvert(s, a1, x + 3);
            if (!overlap(eb, box_of(x, 2, 0.0))) continue;
            const int dtype = orc_dType_EE(x);
            double cr;
            orc_ee_cross(x, &cr, nullptr, nullptr);
            double r0[3], r1[3];
            for (int c = 0; c < 3; ++c) { r0[c] = s->Vrest[(size_t)c * s->nV + a0]; r1[c] = s->Vrest[(size_t)c * s->nV + a1]; }
            const double eps_x = 1.0e-3 * (sq(r0[0] - r1[0]) + sq(r0[1] - r1[1]) + sq(r0[2] - r1[2])) * lenJ; /* MeshCollisionUtils.hpp:2976-2981 */
            const int add_e = (cr < eps_x) ? -eI - 2 : -1;
            double d, y[9];
            Q q;
            const double* P[4] = { x, x + 3, x + 6, x + 9 };
            auto pp = [&](int m, int mi, int ov, int oi) { std::memcpy(y, P[mi], 24); std::memcpy(y + 3, P[oi], 24); orc_d_PP(y, &d); q = { -m - 1, ov, -1, add_e }; };
            switch (dtype) {
            case 0: pp(a0, 0, b0, 2); break;
            case 1: pp(a0, 0, b1, 3); break;
            case 2: std::memcpy(y, P[0], 24); std::memcpy(y + 3, P[2], 24); std::memcpy(y + 6, P[3], 24); orc_d_PE(y, &d); q = { -a0 - 1, b0, b1, add_e }; break;
            case 3: pp(a1, 1, b0, 2); break;
            case 4: pp(a1, 1, b1, 3); break;
            case 5: std::memcpy(y, P[1], 24); std::memcpy(y + 3, P[2], 24); std::memcpy(y + 6, P[3], 24); orc_d_PE(y, &d); q = { -a1 - 1, b0, b1, add_e }; break;
            case 6: std::memcpy(y, P[2], 24); std::memcpy(y + 3, P[0], 24); std::memcpy(y + 6, P[1], 24); orc_d_PE(y, &d); q = { -a0 - 1, -a1 - 1, b0, add_e }; break;
            case 7: std::memcpy(y, P[3], 24); std::memcpy(y + 3, P[0], 24); std::memcpy(y + 6, P[1], 24); orc_d_PE(y, &d); q = { -a0 - 1, -a1 - 1, b1, add_e }; break;
            default:
                orc_d_EE(x, &d);
                q = (add_e <= -2) ? Q{ a0, a1, b0, -b1 - s->nSE - 2 } : Q{ a0, a1, b0, b1 };
            }
            if (d < dHat) { csEE[eJ].push_back(q); cEE[eJ].push_back(eI); }
        }
    }
    /* merge (:2144-2222) */
    std::vector<Q> act, par;
    std::vector<std::array<int, 2>> parE, cnd;
    std::map<Q, int> counter;
    for (int sfI = 0; sfI < o->nF; ++sfI) {
        for (int svI : cPT[sfI]) cnd.push_back({ -svI - 1, sfI });
        for (const Q& c : csPT[sfI]) {
            if (c[3] < 0) ++counter[c];
            else act.push_back(c);
        }
    }
    for (int vI = 0; vI < o->nV; ++vI) {
        for (int sfI : cTP[vI]) cnd.push_back({ -sfI - 1, -vI - 1 });
        for (const Q& c : csTP[vI]) {
            if (c[3] < 0) ++counter[c];
            else act.push_back(c);
        }
    }
    for (int eJ = 0; eJ < o->nE; ++eJ) {
        for (int eI : cEE[eJ]) cnd.push_back({ eI, eJ });
        for (const Q& c : csEE[eJ]) {
            if (c[3] >= 0) act.push_back(c);
            else if (c[3] == -1) ++counter[c];
            else if (c[3] >= -s->nSE - 1) {
                par.push_back({ c[0], c[1], c[2], -1 });
                parE.push_back({ -c[3] - 2, eJ });
            }
            else {
                par.push_back({ c[0], c[1], c[2], -c[3] - s->nSE - 2 });
                parE.push_back({ -1, -1 });
            }
        }
    }
    for (const auto& kv : counter) act.push_back({ kv.first[0], kv.first[1], kv.first[2], -kv.second });
    std::sort(act.begin(), act.end());
    std::vector<int> order(par.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return par[a] < par[b] || (par[a] == par[b] && parE[a] < parE[b]); });
    std::sort(cnd.begin(), cnd.end());
    *nC = (int)act.size();
    *nPara = (int)par.size();
    *nCand = (int)cnd.size();
    if ((int)act.size() > cap || (int)par.size() > capP || (int)cnd.size() > capK) return -1;
    for (size_t i = 0; i < act.size(); ++i) std::memcpy(mmcvid + 4 * i, act[i].data(), 16);
    for (size_t i = 0; i < par.size(); ++i) {
        std::memcpy(para + 4 * i, par[order[i]].data(), 16);
        std::memcpy(para_eIeJ + 2 * i, parE[order[i]].data(), 8);
    }
    for (size_t i = 0; i < cnd.size(); ++i) std::memcpy(cand + 2 * i, cnd[i].data(), 8);
    return 0;
}

/* MeshCO entries -> self-contact entries over the merged numbering (the form the GPU library reports them in) */
void orc_meshco_to_merged(int nV, int nSE, const int* mmcvid, int nC, int* out, const int* para_eIeJ, int nP, int* para_e_out)
{
    for (int c = 0; c < nC; ++c) { Q q = co_to_merged(mmcvid + 4 * c, nV); std::memcpy(out + 4 * c, q.data(), 16); }
    for (int c = 0; c < nP; ++c) {
        para_e_out[2 * c] = para_eIeJ[2 * c];
        para_e_out[2 * c + 1] = para_eIeJ[2 * c + 1] < 0 ? -1 : nSE + para_eIeJ[2 * c + 1];
    }
}

/* Optimizer.cpp:3268-3289 (+ MeshCO.cpp:83-120, :2226-2263) */
int orc_meshco_energy(const orc_surf* s, const orc_obstacle* o, const int* mmcvid, int nC, const int* para, const int* para_eIeJ, int nPara, double dHat, double kappa, double* E)
{
    Merged m;
    merge(s, o, m);
    std::vector<int> act, par, parE;
    translate(s, o, mmcvid, nC, para, para_eIeJ, nPara, act, par, parE);
    return orc_barrier_energy(&m.s, act.data(), nC, par.data(), parE.data(), nPara, dHat, kappa, E);
}

/* Optimizer.cpp:3480-3491 (+ MeshCO.cpp:122-200, :2266-2311); g (3 nV of the MESH, interleaved) += */
void orc_meshco_gradient(const orc_surf* s, const orc_obstacle* o, const int* mmcvid, int nC, const int* para, const int* para_eIeJ, int nPara, double dHat, double kappa, double* g)
{
    Merged m;
    merge(s, o, m);
    std::vector<int> act, par, parE;
    translate(s, o, mmcvid, nC, para, para_eIeJ, nPara, act, par, parE);
    std::vector<double> gm((size_t)3 * m.s.nV, 0.0);
    orc_barrier_gradient(&m.s, act.data(), nC, par.data(), parE.data(), nPara, dHat, kappa, 0, gm.data());
    for (size_t i = 0; i < (size_t)3 * s->nV; ++i) g[i] += gm[i]; /* the obstacle's entries have no rows (:143-195) */
}

/* Optimizer.cpp:3686-3689 (+ MeshCO.cpp:407-586, :2314-2520); the CSR is the MESH's (3 nV rows); a += */
void orc_meshco_hessian_csr(const orc_surf* s, const orc_obstacle* o, const int* mmcvid, int nC, const int* para, const int* para_eIeJ, int nPara, double dHat, double kappa,
    int projectDBC, const int* ia, const int* ja, int index_base, double* a, int nthreads)
{
    Merged m;
    merge(s, o, m);
    std::vector<int> act, par, parE;
    translate(s, o, mmcvid, nC, para, para_eIeJ, nPara, act, par, parE);
    /* obstacle vertices are flagged Dirichlet in the merged surface: their rows and columns are skipped (rowIStart = -1, :441-536) and
     * never looked up in the mesh's pattern */
    orc_barrier_hessian_csr(&m.s, act.data(), nC, par.data(), parE.data(), nPara, dHat, kappa, projectDBC, ia, ja, index_base, a, nthreads);
}

/* MeshCO<3>::largestFeasibleStepSize_TightInclusion (MeshCO.cpp:742-980) over cs_PTEE; every pair sees the step on entry as max_t
 * (the canonical semantics of oracle/ccd.cpp).  Returns 1 if some initial distance was zero (step 0). */
int orc_meshco_ccd_partial(const orc_surf* s, const orc_obstacle* o, const double* p, const int* cand, int nCand, double tol, const double evf[3], const double eee[3],
    int ee_as_vf, double* alpha, int nthreads)
{
    const double max_t = *alpha;
    double best = *alpha;
    int zero = 0;
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 8) reduction(min : best) reduction(max : zero)
    for (int c = 0; c < nCand; ++c) {
        const int a = cand[2 * c], b = cand[2 * c + 1];
        double x0[12], x1[12];
        int kind;
        if (a < 0 && b < 0) { /* TP (:760-840): obstacle vertex, mesh triangle */
            kind = 1;
            const int vI = -b - 1, sfI = -a - 1;
            overt(o, vI, x0);
            std::memcpy(x1, x0, 24);
            for (int k = 0; k < 3; ++k) {
                const int v = s->SF[(size_t)k * s->nSF + sfI];
                mvert(s, v, x0 + 3 + 3 * k);
                for (int q = 0; q < 3; ++q) x1[3 + 3 * k + q] = x0[3 + 3 * k + q] + p[3 * (size_t)v + q];
            }
        }
        else if (a < 0) { /* PT (:842-915): mesh vertex, obstacle triangle */
            kind = 0;
            const int v = s->SVI[-a - 1];
            mvert(s, v, x0);
            for (int q = 0; q < 3; ++q) x1[q] = x0[q] + p[3 * (size_t)v + q];
            for (int k = 0; k < 3; ++k) { overt(o, o->F[(size_t)k * o->nF + b], x0 + 3 + 3 * k); std::memcpy(x1 + 3 + 3 * k, x0 + 3 + 3 * k, 24); }
        }
        else { /* EE (:917-975): mesh edge, obstacle edge */
            kind = 2;
            for (int k = 0; k < 2; ++k) {
                const int v = s->SE[2 * a + k];
                mvert(s, v, x0 + 3 * k);
                for (int q = 0; q < 3; ++q) x1[3 * k + q] = x0[3 * k + q] + p[3 * (size_t)v + q];
                overt(o, o->E[2 * b + k], x0 + 6 + 3 * k);
                std::memcpy(x1 + 6 + 3 * k, x0 + 6 + 3 * k, 24);
            }
        }
        double toi;
        const int r = co_pair_ccd(kind, x0, x1, tol, evf, eee, max_t, ee_as_vf, true, &toi);
        if (r < 0) zero = 1;
        else if (r > 0 && toi < best) best = toi;
    }
    *alpha = zero ? 0.0 : best;
    return zero;
}

/* MeshCO<3>::largestFeasibleStepSize_CCD_TightInclusion (MeshCO.cpp:1388-1668).  Candidates: every (obstacle primitive, mesh primitive)
 * pair whose boxes -- the mesh primitive's swept over [0, alpha on entry], the obstacle's at rest -- overlap: a superset of the pairs
 * that can collide, like the reference's hash queries (:1403, :1490, :1585; the zero-distance rule is applied to these candidates). */
int orc_meshco_ccd_full(const orc_surf* s, const orc_obstacle* o, const double* p, double tol, const double evf[3], const double eee[3], int ee_as_vf, double* alpha,
    long long* nPairs, int nthreads)
{
    const double max_t = *alpha;
    double best = *alpha;
    int zero = 0;
    long long np = 0;
    const int nth = nthreads > 0 ? nthreads : 1;
    auto swept = [&](const int* v, int n, double* x0, double* x1) {
        for (int k = 0; k < n; ++k) {
            mvert(s, v[k], x0 + 3 * k);
            for (int q = 0; q < 3; ++q) x1[3 * k + q] = x0[3 * k + q] + p[3 * (size_t)v[k] + q];
        }
    };
    auto swept_box = [&](const double* x0, const double* x1, int n) {
        double y[24];
        for (int k = 0; k < 3 * n; ++k) { y[k] = x0[k]; y[3 * n + k] = x0[k] + max_t * (x1[k] - x0[k]); }
        return box_of(y, 2 * n, 0.0);
    };
#pragma omp parallel for num_threads(nth) schedule(dynamic, 4) reduction(min : best) reduction(max : zero) reduction(+ : np)
    for (int sfI = 0; sfI < o->nF; ++sfI) { /* :1396-1478 */
        double x0[12], x1[12];
        for (int k = 0; k < 3; ++k) { overt(o, o->F[(size_t)k * o->nF + sfI], x0 + 3 + 3 * k); std::memcpy(x1 + 3 + 3 * k, x0 + 3 + 3 * k, 24); }
        const Box tb = box_of(x0 + 3, 3, 0.0);
        for (int svI = 0; svI < s->nSV; ++svI) {
            const int v = s->SVI[svI];
            swept(&v, 1, x0, x1);
            if (!overlap(tb, swept_box(x0, x1, 1))) continue;
            ++np;
            double toi;
            const int r = co_pair_ccd(0, x0, x1, tol, evf, eee, max_t, ee_as_vf, false, &toi);
            if (r < 0) zero = 1;
            else if (r > 0 && toi < best) best = toi;
        }
    }
#pragma omp parallel for num_threads(nth) schedule(dynamic, 4) reduction(min : best) reduction(max : zero) reduction(+ : np)
    for (int vI = 0; vI < o->nV; ++vI) { /* :1484-1570 */
        double x0[12], x1[12];
        overt(o, vI, x0);
        std::memcpy(x1, x0, 24);
        const Box pb = box_of(x0, 1, 0.0);
        for (int sfI = 0; sfI < s->nSF; ++sfI) {
            const int t[3] = { s->SF[sfI], s->SF[(size_t)s->nSF + sfI], s->SF[(size_t)2 * s->nSF + sfI] };
            swept(t, 3, x0 + 3, x1 + 3);
            if (!overlap(pb, swept_box(x0 + 3, x1 + 3, 3))) continue;
            ++np;
            double toi;
            const int r = co_pair_ccd(1, x0, x1, tol, evf, eee, max_t, ee_as_vf, false, &toi);
            if (r < 0) zero = 1;
            else if (r > 0 && toi < best) best = toi;
        }
    }
#pragma omp parallel for num_threads(nth) schedule(dynamic, 4) reduction(min : best) reduction(max : zero) reduction(+ : np)
    for (int eJ = 0; eJ < o->nE; ++eJ) { /* :1576-1660 */
        double x0[12], x1[12];
        for (int k = 0; k < 2; ++k) { overt(o, o->E[2 * eJ + k], x0 + 6 + 3 * k); std::memcpy(x1 + 6 + 3 * k, x0 + 6 + 3 * k, 24); }
        const Box eb = box_of(x0 + 6, 2, 0.0);
        for (int eI = 0; eI < s->nSE; ++eI) {
            const int e[2] = { s->SE[2 * eI], s->SE[2 * eI + 1] };
            swept(e, 2, x0, x1);
            if (!overlap(eb, swept_box(x0, x1, 2))) continue;
            ++np;
            double toi;
            const int r = co_pair_ccd(2, x0, x1, tol, evf, eee, max_t, ee_as_vf, false, &toi);
            if (r < 0) zero = 1;
            else if (r > 0 && toi < best) best = toi;
        }
    }
    *alpha = zero ? 0.0 : best;
    if (nPairs) *nPairs = np;
    return zero;
}

} // extern "C"
