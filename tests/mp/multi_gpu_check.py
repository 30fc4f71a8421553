"""Launched by torchrun (one rank per GPU): the N-rank hot path must reproduce the single-process oracle.
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mp/multi_gpu_check.py"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import oracle as orc
from bench import contact_pattern_pairs
from ipc_b200 import lib as L
from ipc_b200 import scenes


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scene = os.environ.get("IPCGPU_CHECK_SCENE", "pile4")
    m, info = scenes.ball_pile(4, res=8, seed=5, height=4) if scene == "pile4" else (scenes.ball_on_mat_c3(nx=60) if scene == "c3small" else scenes.sphere_pile_fcc(16))
    dHat, p, kappa, coef = info["dHat"], info["p"], 1e8, 0.025 ** 2
    ctx = L.Context(local)
    ids = [L.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(rank, world, ids[0])
    ctx.set_mesh(m.V_rest_soa, m.T_soa, m.restTriInv, m.vol, m.mu, m.lam, m.mass, m.dbc, m.energy)
    ctx.set_surface(m.SVI, m.SFEdges, m.SF_soa, m.vCoDim)
    ctx.set_state(m.V_soa)
    ctx.set_search_dir(p)
    mm, pa, pe, cand = ctx.constraint_set(dHat, 1)
    ia, ja = m.csr_pattern(1, extra_pairs=contact_pattern_pairs(m, mm, pa, pe))
    ctx.set_csr(ia, ja, 1)
    if os.environ.get("IPCGPU_PARTITION_CONTACT", "1") == "1":  # second build: partitioned sets (the pattern above came from the replicated one)
        ctx.set_contact_partition(1)
        ctx.constraint_set(dHat, 1, fetch=False)
    E = ctx.elastic_energy(coef) + ctx.barrier_energy(dHat, kappa)
    ctx.csr_set_zero()
    ctx.elastic_grad_hess(coef, 1, 1, 1, None, None)
    ctx.barrier_gradient(dHat, kappa, None)
    ctx.barrier_hessian(dHat, kappa, 1, None)
    ctx.allreduce_grad_hess(1, 1)
    g = ctx.download(L.BUF_GRADIENT, 3 * m.nV)
    a = ctx.download(L.BUF_CSR_VALUES, ja.size)
    evf, eee = L.Context.ti_error(m.V_soa, m.nV, None)
    al = ctx.inversion_step(None, 0.2, 1.0)
    al = ctx.ccd_partial(None, 1e-6, evf, eee, al)
    al = ctx.hash_build_swept(None, al, m.avgEdgeLen / 3)
    al, _ = ctx.ccd_full(1e-6, evf, eee, al)
    # the same iteration through the device-resident chain (what bench.py times): nothing read back until the fetch, and every rank
    # downloads only the CSR rows it owns
    ctx.constraint_set(dHat, 1, fetch=False, sizes=False)
    ctx.elastic_energy(coef, 1, want=False)
    ctx.barrier_energy(dHat, kappa, want=False)
    ctx.elastic_grad_hess(coef, 1, 1, 1, None, None)
    ctx.barrier_gradient(dHat, kappa, None)
    ctx.barrier_hessian(dHat, kappa, 1, None)
    ctx.allreduce_grad_hess(1, 0)
    ctx.step_bound_set(1.0)
    ctx.inversion_step(None, 0.2, None)
    ctx.ccd_partial(None, 1e-6, evf, eee, None)
    ctx.hash_build_swept(None, None, m.avgEdgeLen / 3)
    ctx.ccd_full(1e-6, evf, eee, None)
    it = ctx.fetch_iteration()
    part = ctx.partition_info()
    from ipc_b200 import partition as P  # the library's partition must be the documented rule
    vb_ = P.vertex_boundaries(m.T, m.nV, world)
    assert (part["row_vertex_begin"], part["row_vertex_end"]) == (vb_[rank], vb_[rank + 1]), (part, vb_)
    assert (part["tet_begin"], part["tet_end"]) == P.tet_range(m.nT, rank, world)
    assert part["n_assembled_tets"] == len(P.assembled_tets(m.T, vb_[rank], vb_[rank + 1]))
    assert (part["value_begin"], part["value_end"]) == P.owned_value_range(ia, 1, vb_[rank], vb_[rank + 1])
    g_def = ctx.download(L.BUF_GRADIENT, 3 * m.nV)
    a_own = np.zeros(ja.size)
    ctx.download_range_into(L.BUF_CSR_VALUES, part["value_begin"], a_own[part["value_begin"]:part["value_end"]])
    t = torch.from_numpy(a_own).cuda()
    dist.all_reduce(t)  # owned row ranges are disjoint and cover the matrix: their sum is the whole matrix
    a_def = t.cpu().numpy()
    E_def = it.energy_elastic + it.energy_barrier
    ok = True
    if rank == 0:
        o, s = orc.Elastic(m), orc.Surf(m)
        mm_r, pa_r, pe_r, cand_r = s.constraint_set(dHat, nthreads=8)
        assert np.array_equal(mm, mm_r) and np.array_equal(cand, cand_r)
        E_ref = o.energy(coef)[0] + s.barrier_energy(mm_r, pa_r, pe_r, dHat, kappa)[0]
        g_ref = s.barrier_gradient(mm_r, pa_r, pe_r, dHat, kappa, g=o.gradient(coef, 1))
        a_ref = o.hessian_csr(coef, ia, ja, 1, 1, 1)
        for v in range(m.nV):
            for r in range(3):
                a_ref[ia[3 * v + r] - 1] += m.mass[v]
        a_ref = s.barrier_hessian_csr(mm_r, pa_r, pe_r, dHat, kappa, ia, ja, 1, 1, a=a_ref)
        al_ref, _ = o.inversion_step(p, 0.2, 1.0)
        al_ref, _ = orc.ccd_partial(s, p, cand_r, 1e-6, evf, eee, al_ref, 8)
        gr, ag = orc.grid_swept(s, p, al_ref, m.avgEdgeLen / 3)
        al_ref, _, _ = orc.ccd_full(s, p, gr, ag, 1e-6, evf, eee, ag, 8)
        rel = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)
        res = dict(E=abs(E - E_ref) / abs(E_ref), g=rel(g, g_ref), a=rel(a, a_ref), alpha_bits=struct.pack("<d", al) == struct.pack("<d", al_ref),
                   E_deferred=abs(E_def - E_ref) / abs(E_ref), g_deferred=rel(g_def, g_ref), a_deferred_owned_rows=rel(a_def, a_ref),
                   alpha_deferred_bits=struct.pack("<d", it.alpha) == struct.pack("<d", al_ref), status=it.status, ti_warnings=int(it.ti_warnings))
        ok = (res["E"] <= 1e-10 and res["g"] <= 1e-10 and res["a"] <= 1e-9 and res["alpha_bits"] and res["E_deferred"] <= 1e-10 and res["g_deferred"] <= 1e-10
              and res["a_deferred_owned_rows"] <= 1e-9 and res["alpha_deferred_bits"] and it.status == 0 and it.ti_warnings == 0)
        print(f"MULTI_GPU_CHECK world={world} {'OK' if ok else 'FAIL'} {res} alpha={al}")
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
