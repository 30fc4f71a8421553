"""Host-side scene data: synthetic tet-mesh generators and the per-scene precompute that the reference
does once in Mesh<3>::computeFeatures / setLameParam (src/Mesh.cpp:415-527, 661-671) and
LinSysSolver::set_pattern (src/LinSysSolver/LinSysSolver.hpp:46-150).

This is input preparation (numpy), not the hot path; everything here produces exactly the arrays the C ABI
(include/ipcgpu.h) takes, in the reference's own memory layouts.
"""
import numpy as np

# Kuhn split of the unit cube into 6 positively oriented tets around the main diagonal (0,0,0)-(1,1,1)
# (same family of splits as input/tetMeshes/cube.msh: 8 vertices / 6 tets).
_PERMS = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]


def _perm_sign(p):
    s = 1
    p = list(p)
    for i in range(3):
        for j in range(i + 1, 3):
            if p[i] > p[j]:
                s = -s
    return s


def grid_tets(nx, ny, nz, h=None, origin=(0.0, 0.0, 0.0)):
    """nx*ny*nz cells, 6 tets each. Returns V (nV,3) float64, T (nT,4) int32 with det[x1-x0,x2-x0,x3-x0] > 0."""
    if h is None:
        h = 1.0 / nx
    ii, jj, kk = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    V = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1).astype(np.float64) * h + np.asarray(origin)

    def vid(i, j, k):
        return (i * (ny + 1) + j) * (nz + 1) + k

    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ci, cj, ck = ci.ravel(), cj.ravel(), ck.ravel()
    tets = []
    for p in _PERMS:
        off = np.zeros((4, 3), dtype=np.int64)
        for s in range(3):
            off[s + 1] = off[s]
            off[s + 1, p[s]] += 1
        vs = [vid(ci + off[s, 0], cj + off[s, 1], ck + off[s, 2]) for s in range(4)]
        if _perm_sign(p) < 0:
            vs[2], vs[3] = vs[3], vs[2]
        tets.append(np.stack(vs, axis=1))
    # interleave so that the 6 tets of a cell are adjacent (memory locality like a mesher's output)
    T = np.stack(tets, axis=1).reshape(-1, 4).astype(np.int32)
    return V, T


def ball_tets(n, radius=1.0, center=(0.0, 0.0, 0.0)):
    """A tetrahedral ball: the n^3 cube grid on [-1,1]^3 warped radially (p -> p*|p|_inf/|p|_2)."""
    V, T = grid_tets(n, n, n, h=2.0 / n, origin=(-1.0, -1.0, -1.0))
    linf = np.abs(V).max(axis=1)
    l2 = np.linalg.norm(V, axis=1)
    scale = np.where(l2 > 0, linf / np.maximum(l2, 1e-300), 0.0)
    V = V * scale[:, None] * radius + np.asarray(center)
    return V, T


def superball_tets(n, radius=1.0, q=6.0, center=(0.0, 0.0, 0.0)):
    """A rounded-cube "ball" (unit ball of the L_q norm): p -> p*|p|_inf/|p|_q.  q=2 is the round ball; q=6 has nearly
    flat poles, so that stacked balls touch over a patch of surface primitives instead of a single point."""
    V, T = grid_tets(n, n, n, h=2.0 / n, origin=(-1.0, -1.0, -1.0))
    linf = np.abs(V).max(axis=1)
    lq = (np.abs(V) ** q).sum(axis=1) ** (1.0 / q)
    scale = np.where(lq > 0, linf / np.maximum(lq, 1e-300), 0.0)
    V = V * scale[:, None] * radius + np.asarray(center)
    return V, T


def boundary_faces(T):
    """Outward-oriented boundary triangles of a positively oriented tet mesh."""
    a, b, c, d = T[:, 0], T[:, 1], T[:, 2], T[:, 3]
    faces = np.concatenate([np.stack(f, axis=1) for f in ((a, c, b), (a, b, d), (a, d, c), (b, c, d))], axis=0)
    key = np.sort(faces, axis=1)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    keep = cnt[inv.ravel()] == 1
    SF = faces[keep]
    # deterministic order (the reference's $Surface block is sorted by vertex ids)
    order = np.lexsort((SF[:, 2], SF[:, 1], SF[:, 0]))
    return SF[order].astype(np.int32)


def surface_edges(SF):
    """Mesh.cpp:495-516: directed edge (a,b) is kept unless (b,a) was inserted earlier; result sorted."""
    a = np.concatenate([SF[:, 0], SF[:, 1], SF[:, 2]])
    b = np.concatenate([SF[:, 1], SF[:, 2], SF[:, 0]])
    n = SF.shape[0]
    seq = np.concatenate([3 * np.arange(n), 3 * np.arange(n) + 1, 3 * np.arange(n) + 2])
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    order = np.lexsort((seq, hi, lo))
    lo_s, hi_s = lo[order], hi[order]
    first = np.ones(order.size, dtype=bool)
    first[1:] = (lo_s[1:] != lo_s[:-1]) | (hi_s[1:] != hi_s[:-1])
    sel = order[first]
    E = np.stack([a[sel], b[sel]], axis=1)
    E = E[np.lexsort((E[:, 1], E[:, 0]))]
    return E.astype(np.int32)


class Mesh:
    """Arrays the reference's Mesh<3> owns, in the layouts its Eigen members expose through .data()."""

    def __init__(self, V_rest, T, YM=1e5, PR=0.4, density=1000.0, energy=0, SF=None):
        self.V_rest = np.ascontiguousarray(V_rest, dtype=np.float64)  # (nV,3)
        self.V = self.V_rest.copy()
        self.T = np.ascontiguousarray(T, dtype=np.int32)  # (nT,4)
        self.nV, self.nT = self.V_rest.shape[0], self.T.shape[0]
        self.energy = energy
        x = self.V_rest[self.T]  # (nT,4,3)
        X0 = np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0], x[:, 3] - x[:, 0]], axis=2)  # columns = edge vectors
        self.vol = np.linalg.det(X0) / 6.0  # triArea, Mesh.cpp:455
        Ainv = np.linalg.inv(X0)  # restTriInv, Mesh.cpp:449
        self.restTriInv = np.ascontiguousarray(Ainv.transpose(0, 2, 1)).reshape(self.nT, 9)  # per-tet column-major
        self.mu = np.full(self.nT, YM / 2.0 / (1.0 + PR))  # Mesh.cpp:663
        self.lam = np.full(self.nT, YM * PR / (1.0 + PR) / (1.0 - 2.0 * PR))  # Mesh.cpp:664
        self.mass = np.zeros(self.nV)
        np.add.at(self.mass, self.T.ravel(), np.repeat(self.vol * density / 4.0, 4))
        self.dbc = np.zeros(self.nV, dtype=np.uint8)
        # igl::avg_edge_length over the cyclic tet edges (0,1)(1,2)(2,3)(3,0), used for the hash cell size
        cyc = [(0, 1), (1, 2), (2, 3), (3, 0)]
        self.avgEdgeLen = float(np.mean([np.linalg.norm(x[:, i] - x[:, j], axis=1).mean() for i, j in cyc]))
        # surface triangles: the file's own $Surface block when the mesh came from a reference .msh (IglUtils.cpp:548-560),
        # otherwise the boundary faces of the tets (what the reference's findSurfaceTris does for files without one)
        self.SF = boundary_faces(self.T) if SF is None or len(SF) == 0 else np.ascontiguousarray(SF, dtype=np.int32)
        self.SVI = np.unique(self.SF).astype(np.int32)  # Mesh::computeBoundaryVert (sorted surface vertices)
        self.SFEdges = surface_edges(self.SF)
        self.bbox_diag2 = float(((self.V_rest.max(0) - self.V_rest.min(0)) ** 2).sum())
        self.vCoDim = np.full(self.nV, 3, dtype=np.int32)
        self._nbr = None

    # ---- layouts for the C ABI ---------------------------------------------------------------------
    @property
    def V_soa(self):
        return np.ascontiguousarray(self.V.T).ravel()

    @property
    def V_rest_soa(self):
        return np.ascontiguousarray(self.V_rest.T).ravel()

    @property
    def T_soa(self):
        return np.ascontiguousarray(self.T.T).ravel()

    @property
    def SF_soa(self):
        return np.ascontiguousarray(self.SF.T).ravel()

    # ---- adjacency + CSR pattern --------------------------------------------------------------------
    def neighbor_pairs(self, extra_pairs=None):
        """Unique undirected vertex pairs (lo<hi): tet edges + surface edges (+ contact pairs)."""
        T = self.T.astype(np.int64)
        pr = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
        a = np.concatenate([T[:, i] for i, _ in pr])
        b = np.concatenate([T[:, j] for _, j in pr])
        if extra_pairs is not None and len(extra_pairs):
            e = np.asarray(extra_pairs, dtype=np.int64).reshape(-1, 2)
            a = np.concatenate([a, e[:, 0]])
            b = np.concatenate([b, e[:, 1]])
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        keep = lo != hi
        key = np.unique(lo[keep] * self.nV + hi[keep])
        return (key // self.nV).astype(np.int64), (key % self.nV).astype(np.int64)

    def csr_pattern(self, index_base=1, extra_pairs=None):
        """LinSysSolver::set_pattern: upper-triangular 3x3-block CSR (row 3v: [3v,3v+1,3v+2, 3n.. n>v])."""
        lo, hi = self.neighbor_pairs(extra_pairs)
        nV = self.nV
        up = np.bincount(lo, minlength=nV)  # upper neighbours per vertex (lo sorted, hi ascending within lo)
        ptr = np.zeros(nV + 1, dtype=np.int64)
        np.cumsum(up, out=ptr[1:])
        rownnz = np.stack([3 + 3 * up, 2 + 3 * up, 1 + 3 * up], axis=1).ravel()
        ia = np.zeros(3 * nV + 1, dtype=np.int64)
        np.cumsum(rownnz, out=ia[1:])
        nnz = int(ia[-1])
        ja = np.empty(nnz, dtype=np.int64)
        for r in range(3):
            start = ia[r:3 * nV:3]
            for c in range(r, 3):
                ja[start + (c - r)] = 3 * np.arange(nV) + c
            # neighbour blocks
            pos = start[lo] + (3 - r) + 3 * (np.arange(lo.size) - ptr[lo])
            for c in range(3):
                ja[pos + c] = 3 * hi + c
        return (ia + index_base).astype(np.int32), (ja + index_base).astype(np.int32)


def merge_meshes(parts, **kw):
    """Concatenate (V,T[,SF]) parts into one Mesh (multiple bodies = one reference Mesh with several components).  When every part
    brings its surface triangles they are concatenated too (the reference appends each shape's $Surface block, Config/main load
    loop); otherwise the surface is recomputed from the tets."""
    Vs, Ts, Fs, off = [], [], [], 0
    all_sf = all(len(pt) > 2 and pt[2] is not None and len(pt[2]) for pt in parts)
    for pt in parts:
        V, T = pt[0], pt[1]
        Vs.append(V)
        Ts.append(T + off)
        if all_sf:
            Fs.append(np.asarray(pt[2]) + off)
        off += V.shape[0]
    SF = np.concatenate(Fs).astype(np.int32) if all_sf else None
    return Mesh(np.concatenate(Vs), np.concatenate(Ts).astype(np.int32), SF=SF, **kw)


def deform(mesh, seed, twist=0.5, amp=0.02, noise=0.02, require_positive=True):
    """SURVEY 8(d) deformed state: twist about z + sine field + Gaussian noise (scaled by the mean edge)."""
    rng = np.random.default_rng(seed)
    X = mesh.V_rest
    ext = X.max(0) - X.min(0)
    c = 0.5 * (X.max(0) + X.min(0))
    h = mesh.avgEdgeLen
    for attempt in range(8):
        z = (X[:, 2] - X[:, 2].min()) / max(ext[2], 1e-300)
        th = twist * z
        co, si = np.cos(th), np.sin(th)
        x = X - c
        Y = np.stack([co * x[:, 0] - si * x[:, 1], si * x[:, 0] + co * x[:, 1], x[:, 2]], axis=1) + c
        Y += amp * ext.max() * np.sin(3.0 * X[:, [1, 2, 0]] / max(ext.max(), 1e-300) * np.pi)
        Y += rng.normal(0.0, noise * h, size=Y.shape)
        mesh.V = Y
        if not require_positive:
            break
        xx = Y[mesh.T]
        det = np.linalg.det(np.stack([xx[:, 1] - xx[:, 0], xx[:, 2] - xx[:, 0], xx[:, 3] - xx[:, 0]], axis=2))
        if det.min() > 0:
            break
        noise *= 0.5
        amp *= 0.5
    return mesh
