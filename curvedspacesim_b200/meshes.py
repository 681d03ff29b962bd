"""Mesh input for the hot path: OFF loading in the reference's corner convention and the synthetic
meshes BASELINE.json names (icosphere, remeshed torus) plus small closed-form test surfaces.

The reference loads OFF files into CGAL::Surface_mesh (src/models/triangulatedMeshSpace.cpp:42-72)
and reads the corners of a face with halfedges_around_face + source()
(src/utility/meshUtilities.cpp:13-34).  Under CGAL's add_face the face's halfedge is the one that
closes the loop, so an OFF line ``3 a b c`` is seen as corners (c, a, b).
That rotation lives in exactly one place: :func:`reference_corners`.
"""
from __future__ import annotations

import numpy as np


def load_off(path: str):
    """Return (V float64 [nV,3], F int32 [nF,3]) exactly as written in the file."""
    with open(path) as fh:
        toks = fh.read().split()
    if toks[0] != "OFF":
        raise ValueError("Invalid input file.")  # same message as triangulatedMeshSpace.cpp:55
    nv, nf = int(toks[1]), int(toks[2])
    pos = 4
    V = np.array(toks[pos:pos + 3 * nv], dtype=np.float64).reshape(nv, 3)
    pos += 3 * nv
    F = np.empty((nf, 3), dtype=np.int32)
    for i in range(nf):
        if int(toks[pos]) != 3:
            raise ValueError("Non-triangular mesh")  # triangulatedMeshSpace.cpp:61
        F[i] = (int(toks[pos + 1]), int(toks[pos + 2]), int(toks[pos + 3]))
        pos += 4
    return V, F


def save_off(path: str, V, F):
    with open(path, "w") as fh:
        fh.write("OFF\n%d %d 0\n" % (len(V), len(F)))
        for p in V:
            fh.write("%.17g %.17g %.17g\n" % tuple(p))
        for f in F:
            fh.write("3 %d %d %d\n" % tuple(f))


def reference_corners(F):
    """OFF face (a,b,c) -> corner order (c,a,b) used by every barycentric coordinate in the reference."""
    F = np.asarray(F, dtype=np.int32)
    return np.ascontiguousarray(F[:, [2, 0, 1]])


def face_areas(V, F):
    a = V[F[:, 1]] - V[F[:, 0]]
    b = V[F[:, 2]] - V[F[:, 0]]
    return 0.5 * np.linalg.norm(np.cross(a, b), axis=1)


def icosphere(nu: int, radius: float = 1.0):
    """Class-I geodesic icosphere of frequency nu: V = 10 nu^2 + 2, F = 20 nu^2."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    base = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                     [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    base /= np.linalg.norm(base[0])
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
             (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
             (8, 6, 7), (9, 8, 1)]
    verts = {}
    pts = []

    def vid(key, p):
        i = verts.get(key)
        if i is None:
            i = len(pts)
            verts[key] = i
            pts.append(p)
        return i

    def key_for(a, b, c, i, j, k):
        # canonical key: sorted (corner id, weight) pairs with non-zero weight
        return tuple(sorted((v, w) for v, w in ((a, i), (b, j), (c, k)) if w > 0))

    F = []
    for (a, b, c) in faces:
        idx = {}
        for i in range(nu + 1):
            for j in range(nu + 1 - i):
                k = nu - i - j
                p = (i * base[a] + j * base[b] + k * base[c]) / nu
                idx[(i, j)] = vid(key_for(a, b, c, i, j, k), p)
        for i in range(nu):
            for j in range(nu - i):
                F.append((idx[(i, j)], idx[(i + 1, j)], idx[(i, j + 1)]))
                if i + j < nu - 1:
                    F.append((idx[(i + 1, j)], idx[(i + 1, j + 1)], idx[(i, j + 1)]))
    V = np.array(pts, dtype=np.float64)
    V *= radius / np.linalg.norm(V, axis=1, keepdims=True)
    F = np.array(F, dtype=np.int32)
    # make every face counter-clockwise seen from outside
    n = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]])
    flip = np.einsum("ij,ij->i", n, V[F[:, 0]]) < 0
    F[flip] = F[flip][:, [0, 2, 1]]
    return V, F


def torus(nu: int = 1250, nv: int = 400, R: float = 3.0, r: float = 1.0, jitter: float = 0.0, seed: int = 13377,
          flip_parity: bool = True):
    """Parametric torus grid, two triangles per quad (1250 x 400 -> 1 M faces).

    ``jitter`` moves every vertex along the parameter directions by at most jitter * local edge and
    ``flip_parity`` alternates the quad diagonal, a deterministic stand-in for isotropic remeshing
    (CGAL's PMP::isotropic_remeshing is not available here)."""
    rng = np.random.default_rng(seed)
    iu, iv = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    u = iu.astype(np.float64)
    v = iv.astype(np.float64)
    if jitter > 0:
        u = u + jitter * (rng.random(u.shape) * 2 - 1)
        v = v + jitter * (rng.random(v.shape) * 2 - 1)
    th = 2 * np.pi * u / nu
    ph = 2 * np.pi * v / nv
    X = (R + r * np.cos(ph)) * np.cos(th)
    Y = (R + r * np.cos(ph)) * np.sin(th)
    Z = r * np.sin(ph)
    V = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    a = (iu * nv + iv).ravel()
    b = (((iu + 1) % nu) * nv + iv).ravel()
    c = (((iu + 1) % nu) * nv + (iv + 1) % nv).ravel()
    d = (iu * nv + (iv + 1) % nv).ravel()
    par = ((iu + iv) % 2 == 0).ravel() if flip_parity else np.ones(a.shape, bool)
    # quad a(u,v) b(u+1,v) c(u+1,v+1) d(u,v+1); outward normal = d_theta x d_phi
    t1 = np.where(par[:, None], np.stack([a, b, c], 1), np.stack([a, b, d], 1))
    t2 = np.where(par[:, None], np.stack([a, c, d], 1), np.stack([b, c, d], 1))
    F = np.concatenate([t1, t2]).astype(np.int32)
    return V, F


def plane_grid(nx: int, ny: int, lx: float = 1.0, ly: float = 1.0, normal_up: bool = True, tilt=None):
    """Flat triangulated rectangle (open boundary): geodesics are straight chords."""
    xs = np.linspace(0, lx, nx + 1)
    ys = np.linspace(0, ly, ny + 1)
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    V = np.stack([gx, gy, np.zeros_like(gx)], -1).reshape(-1, 3)
    F = []
    for i in range(nx):
        for j in range(ny):
            a = i * (ny + 1) + j
            b = (i + 1) * (ny + 1) + j
            c = (i + 1) * (ny + 1) + j + 1
            d = i * (ny + 1) + j + 1
            if (i + j) % 2 == 0:
                F += [(a, b, c), (a, c, d)]
            else:
                F += [(a, b, d), (b, c, d)]
    F = np.array(F, dtype=np.int32)
    if tilt is not None:
        V = V @ np.asarray(tilt, dtype=np.float64).T
    return V, F


def bowl(nx: int = 24, ny: int = 24, depth: float = 0.35, jitter: float = 0.15, seed: int = 13377):
    """Open curved sheet z = depth (x^2 + y^2) over [-1,1]^2 with jittered interior vertices: a synthetic stand-in for the
    reference's open example meshes (silo_omega0.012_R2.0.off, sp_rb20_isotropic.off) for the open-mesh boundary rules."""
    V, F = plane_grid(nx, ny, 2.0, 2.0)
    V = V - np.array([1.0, 1.0, 0.0])
    rng = np.random.default_rng(seed)
    interior = (np.abs(V[:, 0]) < 1 - 1e-9) & (np.abs(V[:, 1]) < 1 - 1e-9)
    h = 2.0 / max(nx, ny)
    V[interior, :2] += (rng.random((int(interior.sum()), 2)) - 0.5) * 2 * jitter * h
    V[:, 2] = depth * (V[:, 0] ** 2 + V[:, 1] ** 2)
    return V, F


def cube(n: int = 1, side: float = 1.0):
    """Closed cube surface, each side split into n x n quads -> 2 triangles; outward orientation."""
    verts = {}
    pts = []

    def vid(p):
        key = tuple(np.round(p * 4 * n).astype(int))
        i = verts.get(key)
        if i is None:
            i = len(pts)
            verts[key] = i
            pts.append(p)
        return i

    F = []
    # (origin, du, dv) with du x dv = outward normal
    sides = [((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, 0), (0, 1, 0), (1, 0, 0)), ((1, 0, 0), (0, 1, 0), (0, 0, 1)),
             ((0, 0, 0), (0, 0, 1), (0, 1, 0)), ((0, 1, 0), (0, 0, 1), (1, 0, 0)), ((0, 0, 0), (1, 0, 0), (0, 0, 1))]
    for o, du, dv in sides:
        o, du, dv = (np.array(x, dtype=np.float64) for x in (o, du, dv))
        for i in range(n):
            for j in range(n):
                p = [o + (du * (i + a) + dv * (j + b)) / n for a, b in ((0, 0), (1, 0), (1, 1), (0, 1))]
                q = [vid(x * side) for x in p]
                F += [(q[0], q[1], q[2]), (q[0], q[2], q[3])]
    return np.array(pts, dtype=np.float64), np.array(F, dtype=np.int32)


def build_adjacency(corners):
    """adj[f,k] = face across the edge opposite corner k (-1 border), adjk[f,k] = that edge's index there."""
    corners = np.asarray(corners, dtype=np.int64)
    nF = len(corners)
    a = corners[:, [1, 2, 0]].ravel()  # edge k: corner k+1 -> corner k+2
    b = corners[:, [2, 0, 1]].ravel()
    big = int(corners.max()) + 1
    fwd = a * big + b
    rev = b * big + a
    order = np.argsort(fwd, kind="stable")
    sf = fwd[order]
    if np.any(sf[1:] == sf[:-1]):
        raise ValueError("mesh: duplicated directed edge (non-manifold or inconsistently oriented)")
    pos = np.searchsorted(sf, rev)
    pos = np.minimum(pos, len(sf) - 1)
    found = sf[pos] == rev
    he = np.where(found, order[pos], -1)
    adj = np.where(found, he // 3, -1).astype(np.int32).reshape(nF, 3)
    adjk = np.where(found, he % 3, -1).astype(np.int32).reshape(nF, 3)
    return adj, adjk
