"""Synthetic scenes for the BASELINE.json configs (seeded; see SURVEY.md 8(d) / BASELINE.md table).

ball_pile(n_balls, ...) : jittered lattice of tetrahedral balls with surface gaps of U(0.3,1.2)*sqrt(dHat) and a
                          search direction toward the pile centre -- the "1M-tet ball pile, heavy CCD pair count".
twisted_mat(...)        : the "mat-twist ~100K tets, FixedCoRot, no contact" case.
"""
import numpy as np

from . import mesh as M


def ball_pile(n_balls, res=10, radius=0.5, seed=5, energy=0, dhat_rel=1e-3, gap_lo=0.3, gap_hi=1.2, deform_amp=0.01):
    """Balls on a jittered cubic lattice.  Returns (mesh, info) with info = dict(dHat, p, centers).

    Neighbouring balls are separated by gaps g ~ U(gap_lo, gap_hi) * sqrt(dHat) along the lattice axes, so a
    band of surface primitives of every ball is inside the barrier activation distance of its neighbours.
    """
    rng = np.random.default_rng(seed)
    side = int(np.ceil(n_balls ** (1.0 / 3.0)))
    # scene bbox diag decides dHat (Optimizer.cpp:274-281): dHat = (dhat_rel)^2 * bboxDiag^2
    pitch0 = 2.0 * radius
    diag2 = 3.0 * (side * pitch0) ** 2
    dHat = dhat_rel ** 2 * diag2
    sq = np.sqrt(dHat)
    Vb, Tb = M.ball_tets(res, radius)
    parts, centers = [], []
    k = 0
    # cumulative positions with random gaps per axis-layer keep every axis-neighbour gap inside [gap_lo, gap_hi]*sqrt(dHat)
    offs = [np.concatenate([[0.0], np.cumsum(pitch0 + rng.uniform(gap_lo, gap_hi, side - 1) * sq)]) for _ in range(3)]
    for i in range(side):
        for j in range(side):
            for l in range(side):
                if k >= n_balls:
                    break
                c = np.array([offs[0][i], offs[1][j], offs[2][l]])
                # random rotation so that contacts are not axis aligned vertex-vertex only
                q = rng.standard_normal(4)
                q /= np.linalg.norm(q)
                w, x, y, z = q
                R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                              [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                              [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
                parts.append((Vb @ R.T + c, Tb))
                centers.append(c)
                k += 1
    m = M.merge_meshes(parts, energy=energy)
    centers = np.array(centers)
    # mild smooth deformation of every ball (keeps all tets positive), so that F != I everywhere
    X = m.V_rest
    m.V = X + deform_amp * radius * np.sin(2.0 * np.pi * X[:, [1, 2, 0]] / (2.0 * radius)) * 0.2
    # search direction: toward the pile centre, scaled so that alpha_CFL < 1 (forces the full-CCD branch)
    pile_c = centers.mean(0)
    nVb = Vb.shape[0]
    ball_of = np.repeat(np.arange(len(centers)), nVb)
    dirn = pile_c - centers[ball_of]
    nrm = np.linalg.norm(dirn, axis=1, keepdims=True)
    dirn = np.where(nrm > 0, dirn / np.maximum(nrm, 1e-300), 0.0)
    p = dirn * (4.0 * sq) * rng.uniform(0.5, 1.0, (m.nV, 1)) + rng.normal(0, 0.2 * sq, (m.nV, 3))
    info = dict(dHat=dHat, p=np.ascontiguousarray(p).ravel(), centers=centers, radius=radius, n_balls=len(centers))
    return m, info


def twisted_mat(nx=26, ny=26, nz=25, seed=2, energy=1, invert_frac=0.001):
    V, T = M.grid_tets(nx, ny, nz)
    m = M.Mesh(V, T, energy=energy)
    M.deform(m, seed, twist=np.pi, amp=0.02, noise=0.02)
    if invert_frac > 0 and energy == 1:
        rng = np.random.default_rng(seed)
        k = max(1, int(invert_frac * m.nV))
        m.V[rng.integers(0, m.nV, k)] += 0.8 * m.avgEdgeLen * rng.standard_normal((k, 3))
    return m
