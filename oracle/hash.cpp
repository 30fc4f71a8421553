/*
 * hash.cpp -- CPU ORACLE (test infrastructure only; see oracle.h header).
 * The reference's spatial hash (src/Utils/SpatialHash.hpp) restated with the same data structure
 * (std::unordered_map<int, std::vector<int>> voxel, serial inserts, parallel per-primitive voxel lists), and the two
 * drivers that use it: computeConstraintSet (SelfCollisionHandler.cpp:2149-2478) and the full CCD sweep (:1370-1630).
 * These are the algorithmically faithful CPU baselines timed by bench.py; tests check that they return exactly what the
 * brute-force versions in contact.cpp / ccd.cpp return (the active set / step bound do not depend on the hash).
 */
#include "oracle.h"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

struct RefHash {
    double lo[3], inv_h;
    int count[3], count01;
    int edgeStart, triStart;
    std::unordered_map<int, std::vector<int>> voxel;
    std::vector<std::vector<int>> occupancy; /* swept build only: voxels of every point / edge */

    void axis_index(const double* pos, int* out) const /* :841-845 */
    {
        for (int c = 0; c < 3; ++c) out[c] = (int)std::floor((pos[c] - lo[c]) * inv_h);
    }
    int voxel_index(const int* a) const { return a[0] + a[1] * count[0] + a[2] * count01; } /* :850-854 */
};

inline void vert_of(const orc_surf* s, int v, double* out)
{
    out[0] = s->V[v];
    out[1] = s->V[(size_t)s->nV + v];
    out[2] = s->V[(size_t)2 * s->nV + v];
}
inline bool is_dbc(const orc_surf* s, int v) { return s->dbc && s->dbc[v] != 0; }
inline int codim(const orc_surf* s, int v) { return s->vCoDim ? s->vCoDim[v] : 3; }

void add_range(const RefHash& h, const int* mins, const int* maxs, std::vector<int>& out)
{
    for (int iz = mins[2]; iz <= maxs[2]; ++iz)
        for (int iy = mins[1]; iy <= maxs[1]; ++iy)
            for (int ix = mins[0]; ix <= maxs[0]; ++ix) out.push_back(ix + iy * h.count[0] + iz * h.count01);
}

/* :46-201 (static, alpha = 0, p ignored) and :589-750 (swept) */
void build(RefHash& h, const orc_surf* s, const double* p, double alpha, const orc_grid* g, bool swept, int nthreads)
{
    for (int c = 0; c < 3; ++c) { h.lo[c] = g->lo[c]; h.count[c] = g->count[c]; }
    h.inv_h = g->inv_h;
    h.count01 = h.count[0] * h.count[1];
    h.edgeStart = s->nSV;
    h.triStart = s->nSV + s->nSE;
    std::vector<std::array<int, 3>> vmin(s->nV), vmax(s->nV);
#pragma omp parallel for num_threads(nthreads)
    for (int i = 0; i < s->nSV; ++i) {
        int v = s->SVI[i];
        double x[3], xt[3];
        vert_of(s, v, x);
        int a[3], b[3];
        h.axis_index(x, a);
        if (swept) {
            for (int c = 0; c < 3; ++c) xt[c] = x[c] + alpha * p[3 * (size_t)v + c];
            h.axis_index(xt, b);
        }
        else std::memcpy(b, a, sizeof(a));
        for (int c = 0; c < 3; ++c) { vmin[v][c] = std::min(a[c], b[c]); vmax[v][c] = std::max(a[c], b[c]); }
    }
    h.voxel.clear();
    std::vector<std::vector<int>> loc_v(s->nSV), loc_e(s->nSE), loc_t(s->nSF);
#pragma omp parallel for num_threads(nthreads)
    for (int i = 0; i < s->nSV; ++i) add_range(h, vmin[s->SVI[i]].data(), vmax[s->SVI[i]].data(), loc_v[i]);
#pragma omp parallel for num_threads(nthreads)
    for (int e = 0; e < s->nSE; ++e) {
        int a = s->SE[2 * e], b = s->SE[2 * e + 1], mn[3], mx[3];
        for (int c = 0; c < 3; ++c) { mn[c] = std::min(vmin[a][c], vmin[b][c]); mx[c] = std::max(vmax[a][c], vmax[b][c]); }
        add_range(h, mn, mx, loc_e[e]);
    }
#pragma omp parallel for num_threads(nthreads)
    for (int f = 0; f < s->nSF; ++f) {
        int a = s->SF[f], b = s->SF[(size_t)s->nSF + f], d = s->SF[(size_t)2 * s->nSF + f], mn[3], mx[3];
        for (int c = 0; c < 3; ++c) {
            mn[c] = std::min(std::min(vmin[a][c], vmin[b][c]), vmin[d][c]);
            mx[c] = std::max(std::max(vmax[a][c], vmax[b][c]), vmax[d][c]);
        }
        add_range(h, mn, mx, loc_t[f]);
    }
    /* serial inserts, as in the reference (:91-93, :188-197, :740-749) */
    for (int i = 0; i < s->nSV; ++i)
        for (int vx : loc_v[i]) h.voxel[vx].push_back(i);
    for (int e = 0; e < s->nSE; ++e)
        for (int vx : loc_e[e]) h.voxel[vx].push_back(e + h.edgeStart);
    for (int f = 0; f < s->nSF; ++f)
        for (int vx : loc_t[f]) h.voxel[vx].push_back(f + h.triStart);
    if (swept) {
        h.occupancy.assign(h.triStart, {});
        for (int i = 0; i < s->nSV; ++i) h.occupancy[i].swap(loc_v[i]);
        for (int e = 0; e < s->nSE; ++e) h.occupancy[e + h.edgeStart].swap(loc_e[e]);
    }
}

/* queryPointForTriangles(pos, radius) :203-229 */
void query_point_tris(const RefHash& h, const double* pos, double radius, std::unordered_set<int>& out)
{
    double a[3], b[3];
    int mins[3], maxs[3];
    for (int c = 0; c < 3; ++c) { a[c] = pos[c] - radius; b[c] = pos[c] + radius; }
    h.axis_index(a, mins);
    h.axis_index(b, maxs);
    for (int c = 0; c < 3; ++c) { mins[c] = std::max(mins[c], 0); maxs[c] = std::min(maxs[c], h.count[c] - 1); }
    out.clear();
    for (int iz = mins[2]; iz <= maxs[2]; ++iz)
        for (int iy = mins[1]; iy <= maxs[1]; ++iy)
            for (int ix = mins[0]; ix <= maxs[0]; ++ix) {
                auto it = h.voxel.find(ix + iy * h.count[0] + iz * h.count01);
                if (it == h.voxel.end()) continue;
                for (int id : it->second)
                    if (id >= h.triStart) out.insert(id - h.triStart);
            }
}
/* queryEdgeForEdgesWithBBoxCheck(mesh, vBegin, vEnd, radius, edgeInds, eIq) :375-421 */
void query_edge_edges(const RefHash& h, const orc_surf* s, const double* x0, const double* x1, double radius, int eIq, std::vector<int>& out)
{
    double lb[3], rt[3];
    int mins[3], maxs[3];
    for (int c = 0; c < 3; ++c) { lb[c] = std::min(x0[c], x1[c]) - radius; rt[c] = std::max(x0[c], x1[c]) + radius; }
    h.axis_index(lb, mins);
    h.axis_index(rt, maxs);
    for (int c = 0; c < 3; ++c) { mins[c] = std::max(mins[c], 0); maxs[c] = std::min(maxs[c], h.count[c] - 1); }
    out.clear();
    for (int iz = mins[2]; iz <= maxs[2]; ++iz)
        for (int iy = mins[1]; iy <= maxs[1]; ++iy)
            for (int ix = mins[0]; ix <= maxs[0]; ++ix) {
                auto it = h.voxel.find(ix + iy * h.count[0] + iz * h.count01);
                if (it == h.voxel.end()) continue;
                for (int id : it->second) {
                    if (id < h.edgeStart || id >= h.triStart || id - h.edgeStart <= eIq) continue;
                    int eJ = id - h.edgeStart;
                    double y0[3], y1[3];
                    vert_of(s, s->SE[2 * eJ], y0);
                    vert_of(s, s->SE[2 * eJ + 1], y1);
                    bool sep = false;
                    for (int c = 0; c < 3; ++c) {
                        double jl = std::min(y0[c], y1[c]), jh = std::max(y0[c], y1[c]);
                        if (jl - rt[c] > 0.0 || lb[c] - jh > 0.0) sep = true;
                    }
                    if (!sep) out.push_back(eJ);
                }
            }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
}

int pair_ccd_ext(bool vf, const int v[4], const orc_surf* s, const double* p, double tol, const double err[3], double max_t, double* toi_out)
{
    double x0[12], x1[12];
    for (int k = 0; k < 4; ++k) {
        vert_of(s, v[k], x0 + 3 * k);
        for (int c = 0; c < 3; ++c) x1[3 * k + c] = x0[3 * k + c] + p[3 * (size_t)v[k] + c];
    }
    double d;
    if (vf) orc_point_tri_d(x0, &d);
    else orc_edge_edge_d(x0, &d);
    d = std::sqrt(d);
    if (d == 0) return -1;
    double toi, ot;
    int hit = vf ? orc_ti_vf(x0, x1, err, std::min(0.2 * d, 1e-6), tol, max_t, 1000000, 1, &toi, &ot) : orc_ti_ee(x0, x1, err, std::min(0.2 * d, 1e-6), tol, max_t, 1000000, 1, &toi, &ot);
    if (hit && toi < 1e-6) {
        hit = vf ? orc_ti_vf(x0, x1, err, 0.0, tol, max_t, 1000000, 1, &toi, &ot) : orc_ti_ee(x0, x1, err, 0.0, tol, max_t, 1000000, 1, &toi, &ot);
        if (hit) toi *= 0.8;
    }
    if (hit) { *toi_out = toi; return 1; }
    return 0;
}

} // namespace

extern "C" {

/* computeConstraintSet through the reference-style hash; same outputs (canonically sorted) as orc_constraint_set */
int orc_constraint_set_hashed(const orc_surf* s, double dHat, double voxel_size, int cap, int* mmcvid, int* nC, int capP, int* para, int* para_eIeJ, int* nPara,
    int capK, int* cand, int* nCand, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    orc_grid g;
    orc_grid_static(s, voxel_size, &g);
    RefHash h;
    build(h, s, nullptr, 0.0, &g, false, nthreads);
    const double sq = std::sqrt(dHat);
    typedef std::array<int, 4> Q;
    std::vector<std::vector<Q>> csPT(s->nSV), csEE(s->nSE);
    std::vector<std::vector<int>> candPT(s->nSV), candEE(s->nSE);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16)
    for (int svI = 0; svI < s->nSV; ++svI) {
        int vI = s->SVI[svI];
        double x[12];
        vert_of(s, vI, x);
        std::unordered_set<int> tris;
        query_point_tris(h, x, sq, tris);
        for (int sfI : tris) {
            int t[3] = { s->SF[sfI], s->SF[(size_t)s->nSF + sfI], s->SF[(size_t)2 * s->nSF + sfI] };
            if (vI == t[0] || vI == t[1] || vI == t[2]) continue;
            if ((codim(s, vI) < 3 && codim(s, t[0]) < 3) || (is_dbc(s, vI) && is_dbc(s, t[0]) && is_dbc(s, t[1]) && is_dbc(s, t[2]))) continue;
            for (int k = 0; k < 3; ++k) vert_of(s, t[k], x + 3 + 3 * k);
            int ty = orc_dType_PT(x);
            double d, y[9];
            Q q;
            if (ty < 3) { std::memcpy(y, x, 24); std::memcpy(y + 3, x + 3 * (ty + 1), 24); orc_d_PP(y, &d); q = { -vI - 1, t[ty], -1, -1 }; }
            else if (ty < 6) {
                int a = ty - 3, b = (ty - 2) % 3;
                std::memcpy(y, x, 24); std::memcpy(y + 3, x + 3 * (a + 1), 24); std::memcpy(y + 6, x + 3 * (b + 1), 24);
                orc_d_PE(y, &d);
                q = { -vI - 1, t[a], t[b], -1 };
            }
            else { orc_d_PT(x, &d); q = { -vI - 1, t[0], t[1], t[2] }; }
            if (d < dHat) { csPT[svI].push_back(q); candPT[svI].push_back(sfI); }
        }
    }
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16)
    for (int eI = 0; eI < s->nSE; ++eI) {
        int a0 = s->SE[2 * eI], a1 = s->SE[2 * eI + 1];
        double x[12];
        vert_of(s, a0, x);
        vert_of(s, a1, x + 3);
        std::vector<int> edges;
        query_edge_edges(h, s, x, x + 3, sq, eI, edges);
        for (int eJ : edges) {
            int b0 = s->SE[2 * eJ], b1 = s->SE[2 * eJ + 1];
            if (a0 == b0 || a0 == b1 || a1 == b0 || a1 == b1 || eI > eJ) continue;
            if ((codim(s, a0) < 3 && codim(s, b0) < 3) || (is_dbc(s, a0) && is_dbc(s, a1) && is_dbc(s, b0) && is_dbc(s, b1))) continue;
            vert_of(s, b0, x + 6);
            vert_of(s, b1, x + 9);
            int ty = orc_dType_EE(x);
            double cr;
            orc_ee_cross(x, &cr, nullptr, nullptr);
            double r0[3], r1[3], r2[3], r3[3];
            for (int c = 0; c < 3; ++c) {
                r0[c] = s->Vrest[(size_t)c * s->nV + a0]; r1[c] = s->Vrest[(size_t)c * s->nV + a1];
                r2[c] = s->Vrest[(size_t)c * s->nV + b0]; r3[c] = s->Vrest[(size_t)c * s->nV + b1];
            }
            double la = 0, lb = 0;
            for (int c = 0; c < 3; ++c) { la += (r0[c] - r1[c]) * (r0[c] - r1[c]); lb += (r2[c] - r3[c]) * (r2[c] - r3[c]); }
            int add_e = (cr < 1.0e-3 * la * lb) ? -eJ - 2 : -1;
            double d, y[9];
            Q q;
            const int vid[4] = { a0, a1, b0, b1 };
            auto PP = [&](int i, int j) { std::memcpy(y, x + 3 * i, 24); std::memcpy(y + 3, x + 3 * j, 24); orc_d_PP(y, &d); q = { -vid[i] - 1, vid[j], -1, add_e }; };
            auto PE = [&](int i, int j, int k) { std::memcpy(y, x + 3 * i, 24); std::memcpy(y + 3, x + 3 * j, 24); std::memcpy(y + 6, x + 3 * k, 24); orc_d_PE(y, &d); q = { -vid[i] - 1, vid[j], vid[k], add_e }; };
            switch (ty) {
            case 0: PP(0, 2); break;
            case 1: PP(0, 3); break;
            case 2: PE(0, 2, 3); break;
            case 3: PP(1, 2); break;
            case 4: PP(1, 3); break;
            case 5: PE(1, 2, 3); break;
            case 6: PE(2, 0, 1); break;
            case 7: PE(3, 0, 1); break;
            default: orc_d_EE(x, &d); q = (add_e <= -2) ? Q{ a0, a1, b0, -b1 - s->nSE - 2 } : Q{ a0, a1, b0, b1 };
            }
            if (d < dHat) { csEE[eI].push_back(q); candEE[eI].push_back(eJ); }
        }
    }
    /* serial merge (:2411-2476), canonical sort */
    std::vector<Q> act, par;
    std::vector<std::array<int, 2>> parE, cnd;
    std::map<Q, int> counter;
    for (int svI = 0; svI < s->nSV; ++svI) {
        for (int sfI : candPT[svI]) cnd.push_back({ -svI - 1, sfI });
        for (const Q& c : csPT[svI]) { if (c[3] < 0) ++counter[c]; else act.push_back(c); }
    }
    for (int eI = 0; eI < s->nSE; ++eI) {
        for (int eJ : candEE[eI]) cnd.push_back({ eI, eJ });
        for (const Q& c : csEE[eI]) {
            if (c[3] >= 0) act.push_back(c);
            else if (c[3] == -1) ++counter[c];
            else if (c[3] >= -s->nSE - 1) { par.push_back({ c[0], c[1], c[2], -1 }); parE.push_back({ eI, -c[3] - 2 }); }
            else { par.push_back({ c[0], c[1], c[2], -c[3] - s->nSE - 2 }); parE.push_back({ -1, -1 }); }
        }
    }
    for (const auto& kv : counter) act.push_back({ kv.first[0], kv.first[1], kv.first[2], -kv.second });
    std::sort(act.begin(), act.end());
    std::vector<int> order(par.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return par[a] < par[b] || (par[a] == par[b] && parE[a] < parE[b]); });
    std::sort(cnd.begin(), cnd.end());
    *nC = (int)act.size(); *nPara = (int)par.size(); *nCand = (int)cnd.size();
    if ((int)act.size() > cap || (int)par.size() > capP || (int)cnd.size() > capK) return -1;
    for (size_t i = 0; i < act.size(); ++i) std::memcpy(mmcvid + 4 * i, act[i].data(), 16);
    for (size_t i = 0; i < par.size(); ++i) { std::memcpy(para + 4 * i, par[order[i]].data(), 16); std::memcpy(para_eIeJ + 2 * i, parE[order[i]].data(), 8); }
    for (size_t i = 0; i < cnd.size(); ++i) std::memcpy(cand + 2 * i, cnd[i].data(), 8);
    return 0;
}

/* swept hash build + full CCD sweep through it (SpatialHash.hpp:589-832, SelfCollisionHandler.cpp:1370-1630) */
int orc_ccd_full_hashed(const orc_surf* s, const double* p, double* alpha_inout, double voxel_size, double tol, const double err_vf[3], const double err_ee[3],
    long long* nPairs, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    orc_grid g;
    orc_grid_swept(s, p, alpha_inout, voxel_size, &g);
    const double alpha = *alpha_inout;
    RefHash h;
    build(h, s, p, alpha, &g, true, nthreads);
    const double max_t = alpha;
    double best = alpha;
    int zero = 0;
    long long pairs = 0;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 8) reduction(min : best) reduction(max : zero) reduction(+ : pairs)
    for (int svI = 0; svI < s->nSV; ++svI) {
        int vI = s->SVI[svI];
        std::unordered_set<int> tris;
        for (int vx : h.occupancy[svI]) { /* queryPointForPrimitives (:752-773), triangle part */
            auto it = h.voxel.find(vx);
            if (it == h.voxel.end()) continue;
            for (int id : it->second)
                if (id >= h.triStart) tris.insert(id - h.triStart);
        }
        for (int f : tris) {
            int v[4] = { vI, s->SF[f], s->SF[(size_t)s->nSF + f], s->SF[(size_t)2 * s->nSF + f] };
            if (vI == v[1] || vI == v[2] || vI == v[3]) continue;
            if ((codim(s, vI) < 3 && codim(s, v[1]) < 3) || (is_dbc(s, vI) && is_dbc(s, v[1]) && is_dbc(s, v[2]) && is_dbc(s, v[3]))) continue;
            ++pairs;
            double toi;
            int r = pair_ccd_ext(true, v, s, p, tol, err_vf, max_t, &toi);
            if (r < 0) zero = 1;
            else if (r > 0 && toi < best) best = toi;
        }
    }
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 8) reduction(min : best) reduction(max : zero) reduction(+ : pairs)
    for (int eI = 0; eI < s->nSE; ++eI) {
        const int a0 = s->SE[2 * eI], a1 = s->SE[2 * eI + 1];
        double bl[3], bh[3];
        for (int c = 0; c < 3; ++c) {
            double x0 = s->V[(size_t)c * s->nV + a0], x1 = s->V[(size_t)c * s->nV + a1];
            double y0 = x0 + max_t * p[3 * (size_t)a0 + c], y1 = x1 + max_t * p[3 * (size_t)a1 + c];
            bh[c] = std::max(std::max(x0, y0), std::max(x1, y1));
            bl[c] = std::min(std::min(x0, y0), std::min(x1, y1));
        }
        std::unordered_set<int> edges; /* queryEdgeForEdgesWithBBoxCheck (:803-832) */
        for (int vx : h.occupancy[eI + h.edgeStart]) {
            auto it = h.voxel.find(vx);
            if (it == h.voxel.end()) continue;
            for (int id : it->second) {
                if (id < h.edgeStart || id >= h.triStart || id - h.edgeStart <= eI) continue;
                int eJ = id - h.edgeStart;
                int b0 = s->SE[2 * eJ], b1 = s->SE[2 * eJ + 1];
                bool sep = false;
                for (int c = 0; c < 3; ++c) {
                    double x0 = s->V[(size_t)c * s->nV + b0], x1 = s->V[(size_t)c * s->nV + b1];
                    double y0 = x0 + max_t * p[3 * (size_t)b0 + c], y1 = x1 + max_t * p[3 * (size_t)b1 + c];
                    double jh = std::max(std::max(x0, y0), std::max(x1, y1)), jl = std::min(std::min(x0, y0), std::min(x1, y1));
                    if (jl - bh[c] > 0.0 || bl[c] - jh > 0.0) sep = true;
                }
                if (!sep) edges.insert(eJ);
            }
        }
        for (int eJ : edges) {
            const int b0 = s->SE[2 * eJ], b1 = s->SE[2 * eJ + 1];
            if (a0 == b0 || a0 == b1 || a1 == b0 || a1 == b1) continue;
            if ((codim(s, a0) < 3 && codim(s, b0) < 3) || (is_dbc(s, a0) && is_dbc(s, a1) && is_dbc(s, b0) && is_dbc(s, b1))) continue;
            ++pairs;
            int v[4] = { a0, a1, b0, b1 };
            double toi;
            int r = pair_ccd_ext(false, v, s, p, tol, err_ee, max_t, &toi);
            if (r < 0) zero = 1;
            else if (r > 0 && toi < best) best = toi;
        }
    }
    if (nPairs) *nPairs = pairs;
    *alpha_inout = zero ? 0.0 : best;
    return zero;
}

} // extern "C"
