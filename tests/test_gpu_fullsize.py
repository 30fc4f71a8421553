"""GPU parity at BASELINE.json's FULL size: the 1M-tet ball pile of bench.py, every stage of the hot path through the C ABI against
the oracle's reference-style (hashed) drivers on all host cores.  Integer sets identical, E/g 1e-10, CSR values 1e-9, step bit-exact.
Size-independent properties ride along: the elastic and barrier forces each sum to zero, the step bound is intersection free."""
import os
import struct
import sys

import numpy as np
import pytest

import oracle as orc
from ipc_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def bits(x):
    return struct.pack("<d", float(x))


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.timeout(900)
def test_full_size_pile_every_stage(gpu_ctx):
    import bench

    class A:
        tets, res = 1_000_000, 10

    m, info = bench.build_scene(A)
    assert m.nT >= 1_000_000
    nth = os.cpu_count() or 8
    dHat, p, kappa, dt2, tol = info["dHat"], info["p"], bench.KAPPA, bench.DT2, bench.TI_TOL
    hvox = m.avgEdgeLen / 3.0
    ctx = gpu_ctx
    ctx.set_mesh(m.V_rest_soa, m.T_soa, m.restTriInv, m.vol, m.mu, m.lam, m.mass, m.dbc, m.energy)
    ctx.set_surface(m.SVI, m.SFEdges, m.SF_soa, m.vCoDim)
    ctx.set_state(m.V_soa)
    o, s = orc.Elastic(m), orc.Surf(m)

    # constraint set: identical integer sets
    mm, pa, pe, cand = ctx.constraint_set(dHat, 1)
    mm_r, pa_r, pe_r, cand_r = s.constraint_set_hashed(dHat, hvox, nth)
    assert len(mm_r) > 10_000
    assert np.array_equal(mm, mm_r) and np.array_equal(pa, pa_r) and np.array_equal(pe, pe_r) and np.array_equal(cand, cand_r)

    ia, ja = m.csr_pattern(1, extra_pairs=bench.contact_pattern_pairs(m, mm, pa, pe))
    ctx.set_csr(ia, ja, 1)

    # energies
    E, Er = ctx.elastic_energy(dt2), o.energy(dt2, nth)[0]
    assert abs(E - Er) <= RTOL * abs(Er)
    Eb = ctx.barrier_energy(dHat, kappa)
    Ebr, bad = s.barrier_energy(mm_r, pa_r, pe_r, dHat, kappa)
    assert bad == 0 and abs(Eb - Ebr) <= RTOL * abs(Ebr)

    # gradients (+ property: internal forces of each kind sum to zero)
    g = ctx.elastic_gradient(dt2, 1, 1)
    g_r = o.gradient(dt2, 1, nth)
    assert rel(g, g_r) <= RTOL
    assert np.abs(g.reshape(-1, 3).sum(0)).max() <= 1e-9 * np.abs(g).max()
    gb = np.zeros(3 * m.nV)
    ctx.barrier_gradient(dHat, kappa, gb)
    gb_r = np.zeros(3 * m.nV)
    s.barrier_gradient(mm_r, pa_r, pe_r, dHat, kappa, g=gb_r)
    assert rel(gb, gb_r) <= RTOL
    assert np.abs(gb.reshape(-1, 3).sum(0)).max() <= 1e-9 * np.abs(gb).max()

    # Hessian: elastic + mass + barrier in the solver's CSR
    a = np.zeros(ja.size)
    ctx.elastic_hessian(dt2, 1, 1, 1, a)
    ctx.barrier_hessian(dHat, kappa, 1, a)
    a_r = o.hessian_csr(dt2, ia, ja, 1, 1, 1, nthreads=nth)
    s.barrier_hessian_csr(mm_r, pa_r, pe_r, dHat, kappa, ia, ja, 1, 1, a=a_r, nthreads=nth)
    assert rel(a, a_r) <= 1e-9

    # step bounds: inversion filter, partial CCD, swept hash, full CCD -- the step is bit-exact
    al = ctx.inversion_step(p, 0.2, 1.0)
    al_r, _ = o.inversion_step(p, 0.2, 1.0)
    assert abs(al - al_r) <= 1e-9 * al_r
    evf, eee = L.Context.ti_error(m.V_soa, m.nV, p)
    al = ctx.ccd_partial(p, tol, evf, eee, al_r)
    al_r, _ = orc.ccd_partial(s, p, cand_r, tol, evf, eee, al_r, nth)
    assert bits(al) == bits(al_r)
    ag = ctx.hash_build_swept(p, al, hvox)
    al2, ncand = ctx.ccd_full(tol, evf, eee, ag)
    al2_r, _, npairs = orc.ccd_full_hashed(s, p, al_r, hvox, tol, evf, eee, nth)
    assert ncand == npairs and bits(al2) == bits(al2_r)
    assert ctx.ccd_stats()[2] == 0  # no conservative early-out was needed
    assert 0.0 < al2 <= 1.0

    # property: at 0.999 * step every active stencil still has positive distance
    V2 = m.V + 0.999 * al2 * p.reshape(-1, 3)
    rng = np.random.default_rng(1)
    for c in cand[rng.integers(0, len(cand), 300)]:
        if c[0] < 0:
            assert orc.point_tri_d(V2[[m.SVI[-c[0] - 1]] + list(m.SF[c[1]])]) > 0
        else:
            assert orc.edge_edge_d(V2[list(m.SFEdges[c[0]]) + list(m.SFEdges[c[1]])]) > 0
