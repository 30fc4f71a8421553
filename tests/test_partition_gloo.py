"""world_size-2 CPU (gloo) test of the multi-rank design: a block partition of the tets plus a SUM all-reduce of the
per-rank gradients / CSR values, and a MIN all-reduce of the per-rank step bounds, reproduce the single-process result."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as orc
    from ipc_b200 import mesh as M
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    V, T = M.grid_tets(4, 3, 3)
    m = M.Mesh(V, T, energy=0)
    M.deform(m, 3)
    ia, ja = m.csr_pattern(1)
    tb, te = m.nT * rank // world, m.nT * (rank + 1) // world  # the context's partition rule (api.cu)
    sub = M.Mesh.__new__(M.Mesh)
    sub.__dict__.update(m.__dict__)
    sub.T, sub.nT = m.T[tb:te], te - tb
    sub.restTriInv, sub.vol, sub.mu, sub.lam = m.restTriInv[tb:te], m.vol[tb:te], m.mu[tb:te], m.lam[tb:te]
    o = orc.Elastic(sub)
    g = torch.from_numpy(o.gradient(0.5, 0))
    a = torch.from_numpy(o.hessian_csr(0.5, ia, ja, 1, 1, 0))
    p = np.random.default_rng(1).standard_normal(3 * m.nV) * m.avgEdgeLen
    step = torch.tensor([o.inversion_step(p, 0.2, 1.0)[0]], dtype=torch.float64)
    dist.all_reduce(g, op=dist.ReduceOp.SUM); dist.all_reduce(a, op=dist.ReduceOp.SUM); dist.all_reduce(step, op=dist.ReduceOp.MIN)
    if rank == 0:
        full = orc.Elastic(m)
        ok = (np.allclose(g.numpy(), full.gradient(0.5, 0), rtol=0, atol=1e-12 * np.abs(g.numpy()).max())
              and np.allclose(a.numpy(), full.hessian_csr(0.5, ia, ja, 1, 1, 0), rtol=0, atol=1e-12 * np.abs(a.numpy()).max())
              and float(step) == full.inversion_step(p, 0.2, 1.0)[0])
        q.put(ok)
    dist.destroy_process_group()


def test_tet_partition_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29641, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert q.get(timeout=5) is True
