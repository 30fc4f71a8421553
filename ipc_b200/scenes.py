"""Synthetic scenes for the BASELINE.json configs (seeded; see SURVEY.md 8(d) / BASELINE.md table).

ball_pile(n_balls, ...) : jittered lattice of tetrahedral balls with surface gaps of U(0.3,1.2)*sqrt(dHat) and a
                          search direction toward the pile centre -- the "1M-tet ball pile, heavy CCD pair count".
twisted_mat(...)        : the "mat-twist ~100K tets, FixedCoRot, no contact" case.
"""
import numpy as np

from . import mesh as M


def ball_pile(n_balls, res=10, radius=0.5, seed=5, energy=0, dhat_rel=1e-3, gap_lo=0.3, gap_hi=1.2, q=6.0, height=5):
    """Columns of stacked L_q "balls" (rounded, nearly flat poles).  Returns (mesh, info).

    Balls of one column are separated along z by gaps g ~ U(gap_lo, gap_hi) * sqrt(dHat) (pole to pole), rotated about z
    by a random angle (so facing edges are generically not parallel) and shifted laterally by a small jitter; columns
    are far enough apart not to interact.  info = dict(dHat, p, ...): dHat = (dhat_rel)^2 * bboxDiag^2
    (Optimizer.cpp:274-281) and p is a search direction that squeezes every column toward its middle with
    |p| ~ 2-4 sqrt(dHat), i.e. alpha_CFL < 1: the full-CCD branch of Optimizer.cpp:1961 is taken.
    """
    rng = np.random.default_rng(seed)
    n_cols = int(np.ceil(n_balls / height))
    side = int(np.ceil(np.sqrt(n_cols)))
    col_pitch = 2.0 * radius + 0.5
    ext = np.array([side * col_pitch, side * col_pitch, height * 2.0 * radius])
    dHat = dhat_rel ** 2 * float((ext ** 2).sum())
    sq = np.sqrt(dHat)
    Vb, Tb = M.superball_tets(res, radius, q)
    parts, centers, col_mid = [], [], []
    k = 0
    for ci in range(n_cols):
        cx, cy = (ci % side) * col_pitch, (ci // side) * col_pitch
        z = 0.0
        zs = []
        for l in range(height):
            if k >= n_balls:
                break
            th = rng.uniform(0, 2 * np.pi)
            R = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]])
            c = np.array([cx + rng.normal(0, 0.02), cy + rng.normal(0, 0.02), z])
            parts.append((Vb @ R.T + c, Tb))
            centers.append(c)
            zs.append(z)
            z += 2.0 * radius + rng.uniform(gap_lo, gap_hi) * sq
            k += 1
        col_mid += [0.5 * (zs[0] + zs[-1])] * len(zs)
    m = M.merge_meshes(parts, energy=energy)
    centers = np.array(centers)
    col_mid = np.array(col_mid)
    # mild smooth volumetric deformation (keeps every tet positive and leaves the poles' gaps almost untouched)
    X = m.V_rest
    nVb = Vb.shape[0]
    ball_of = np.repeat(np.arange(len(centers)), nVb)
    local = X - centers[ball_of]
    m.V = X + 0.01 * radius * np.stack([np.sin(2 * np.pi * local[:, 1]), np.sin(2 * np.pi * local[:, 0]), 0 * local[:, 2]], axis=1)
    # search direction: every ball moves toward its column's middle
    dz = col_mid[ball_of] - centers[ball_of, 2]
    p = np.zeros((m.nV, 3))
    p[:, 2] = np.sign(dz) * 3.0 * sq * rng.uniform(0.6, 1.0, m.nV)
    p += rng.normal(0, 0.1 * sq, (m.nV, 3))
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


def ball_on_mat(nx=40, res=6, seed=3, energy=0, dhat_rel=1e-3, gap_lo=0.2, gap_hi=1.5):
    """BASELINE config C3 ("ball-on-mat, barrier contact + CCD line search"): a one-cell-thick mat of nx x nx x 1 cells (all surface, as
    the reference's thin mats are) and a ball hovering over its middle at a gap of U(gap_lo, gap_hi)*sqrt(dHat); the search direction
    pushes the ball down by up to 2 sqrt(dHat) (nx=200 -> 240,000 mat tets).  Returns (mesh, info) like ball_pile."""
    rng = np.random.default_rng(seed)
    h = 1.0 / nx
    Vm, Tm = M.grid_tets(nx, nx, 1, h=h)
    radius = 0.15
    Vb, Tb = M.superball_tets(res, radius, 2.0)
    ext = np.array([1.0, 1.0, h + 2 * radius])
    dHat = dhat_rel ** 2 * float((ext ** 2).sum())
    sq = np.sqrt(dHat)
    gap = rng.uniform(gap_lo, gap_hi) * sq
    nVm = Vm.shape[0]
    Vm_def = Vm + 0.02 * h * rng.standard_normal((nVm, 3)) * np.array([1.0, 1.0, 0.2])
    # the ball's lowest vertex sits `gap` above the highest mat vertex; off the lattice so that no feature pair is exactly degenerate
    c = np.array([0.5 + 0.31 * h, 0.5 - 0.17 * h, Vm_def[:, 2].max() + gap - Vb[:, 2].min()])
    m = M.merge_meshes([(Vm, Tm), (Vb + c, Tb)], energy=energy)
    m.V = m.V_rest.copy()
    m.V[:nVm] = Vm_def
    p = np.zeros((m.nV, 3))
    p[nVm:, 2] = -rng.uniform(0.0, 2.0, m.nV - nVm) * sq
    p += rng.normal(0, 0.05 * sq, (m.nV, 3))
    return m, dict(dHat=dHat, p=np.ascontiguousarray(p).ravel(), n_mat_verts=nVm, gap=gap)
