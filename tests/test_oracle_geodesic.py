"""CPU oracle: exact geodesics pinned by closed-form cases, the independent exhaustive-unfolding checker
and metric properties (SURVEY.md §8(c) "KATs we must create": (1) planar patch, (2) developable exact
cases, (3) boundary pseudo-source, (4) symmetry/property tests, (5) full-mesh vs patch
self-consistency of meshTesting.cpp:201-203, (7) brute-force cross-check).

The reference ships no golden vectors for CGAL::Surface_mesh_shortest_path (parity unpinned); these
tests are what pins the oracle instead."""
from __future__ import annotations

import math

import numpy as np
import pytest

from bruteforce_geodesic import BruteGeodesic
from curvedspacesim_b200 import meshes
from helpers import random_positions
from oracle_binding import Oracle


def _unit(v):
    v = np.asarray(v, np.float64)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _locate(V, corners, p):
    """(face, bary) of a point lying on the mesh (brute force; tests only)."""
    for f, c in enumerate(corners):
        a, b, cc = V[c[0]], V[c[1]], V[c[2]]
        n = np.cross(b - a, cc - a)
        if abs(np.dot(p - a, n)) > 1e-12 * max(1.0, np.linalg.norm(n)):
            continue
        m = np.stack([b - a, cc - a], 1)
        uv, *_ = np.linalg.lstsq(m, p - a, rcond=None)
        w = np.array([1 - uv[0] - uv[1], uv[0], uv[1]])
        if np.all(w > 1e-9):
            return f, w
    raise ValueError("point not strictly inside any face: %r" % (p,))


def _query(orc, V, corners, src, tgts):
    sf, sb = _locate(V, corners, np.asarray(src, float))
    loc = [_locate(V, corners, np.asarray(t, float)) for t in tgts]
    d, ts, te, tie, st = orc.distance(sf, sb, [l[0] for l in loc], np.array([l[1] for l in loc]))
    return d, ts, te


# ---------------------------------------------------------------------------------------------- (1)
@pytest.mark.parametrize("tilt", [None, "rot"])
def test_planar_patch_is_euclidean(tilt):
    R = None
    if tilt:
        a, b = 0.7, -0.4
        Rx = np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]])
        Ry = np.array([[math.cos(b), 0, math.sin(b)], [0, 1, 0], [-math.sin(b), 0, math.cos(b)]])
        R = Rx @ Ry
    V, F = meshes.plane_grid(7, 5, 1.4, 1.0, tilt=R)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    rng = np.random.default_rng(1)
    face, bary = random_positions(len(F), 40, rng)
    P = orc.euclidean(face, bary)
    for s in range(0, 40, 8):
        tf, tb = np.delete(face, s), np.delete(bary, s, 0)
        d, ts, te, tie, _ = orc.distance(face[s], bary[s], tf, tb)
        chord = np.delete(P, s, 0) - P[s]
        L = np.linalg.norm(chord, axis=1)
        assert np.max(np.abs(d - L)) < 1e-13
        assert np.max(np.abs(ts - chord / L[:, None])) < 1e-12
        assert np.max(np.abs(te - chord / L[:, None])) < 1e-12


# ---------------------------------------------------------------------------------------------- (2)
def test_cube_closed_form():
    V, F = meshes.cube(3, 1.0)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    # top face z=1 -> +x face: straight unfolding across the shared edge x=1,z=1
    src = (0.81, 0.47, 1.0)
    tg = [(1.0, 0.52, 0.77), (1.0, 0.31, 0.58)]
    d, ts, te = _query(orc, V, corners, src, tg)
    for i, t in enumerate(tg):
        exp = math.hypot((1 - src[0]) + (1 - t[2]), src[1] - t[1])
        assert abs(d[i] - exp) < 1e-13
        # start tangent lies in the top face and points along the unfolded chord
        u = _unit([(1 - src[0]) + (1 - t[2]), t[1] - src[1]])
        assert np.allclose(ts[i], [u[0], u[1], 0.0], atol=1e-12)
        assert np.allclose(te[i], [0.0, u[1], -u[0]], atol=1e-12)
    # top -> bottom through the +x side: |..| = (1-x) + 1 + (1-x'), dy
    src = (0.83, 0.47, 1.0)
    t = (0.79, 0.55, 0.0)
    d, ts, te = _query(orc, V, corners, src, [t])
    assert abs(d[0] - math.hypot((1 - src[0]) + 1 + (1 - t[0]), src[1] - t[1])) < 1e-13
    # centre of a face to the centre of an adjacent face (many tied unfoldings do not matter: distance is unique)
    d, _, _ = _query(orc, V, corners, (0.5 + 1e-3, 0.5 + 2e-3, 1.0), [(1.0, 0.5 + 2e-3, 0.5)])
    assert abs(d[0] - (1.0 - 1e-3)) < 1e-13


def test_regular_tetrahedron_face_centroids():
    a = 1.0
    V = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], float) * a / (2 * math.sqrt(2))
    F = np.array([[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]], np.int32)
    n = np.cross(V[F[0, 1]] - V[F[0, 0]], V[F[0, 2]] - V[F[0, 0]])
    if np.dot(n, V[F[0]].mean(0)) < 0:
        F = F[:, ::-1].copy()
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    b = np.array([1 / 3, 1 / 3, 1 / 3])
    for f in (1, 2, 3):
        d, ts, te, tie, _ = orc.distance(0, b, [f], b[None])
        assert abs(d[0] - a / math.sqrt(3)) < 1e-13  # two centroid-to-edge-midpoint legs, collinear when unfolded
    # off-centre source: nothing is tied, and the closed form is the straight unfolded chord
    bs = np.array([0.5, 0.3, 0.2])
    d, ts, te, tie, _ = orc.distance(0, bs, [1], np.array([[0.25, 0.35, 0.4]]))
    bg = BruteGeodesic(V, corners)
    D, TS, TE = bg.solve(0, bs, [1], np.array([[0.25, 0.35, 0.4]]))
    assert abs(d[0] - D[0]) < 1e-13 and np.allclose(ts[0], TS[0], atol=1e-12) and np.allclose(te[0], TE[0], atol=1e-12)


# ---------------------------------------------------------------------------------------------- (3)
def _l_shape():
    V, F = meshes.plane_grid(4, 4, 1.0, 1.0)
    cen = V[F].mean(1)
    keep = ~((cen[:, 0] > 0.5) & (cen[:, 1] > 0.5))
    return V, F[keep].copy()


def test_boundary_vertex_is_a_pseudo_source():
    V, F = _l_shape()
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    src, tgt, bend = np.array([0.9, 0.3, 0]), np.array([0.3, 0.9, 0]), np.array([0.5, 0.5, 0])
    d, ts, te = _query(orc, V, corners, src, [tgt])
    assert abs(d[0] - (np.linalg.norm(bend - src) + np.linalg.norm(tgt - bend))) < 1e-13
    assert np.allclose(ts[0], _unit(bend - src), atol=1e-12)
    assert np.allclose(te[0], _unit(tgt - bend), atol=1e-12)
    # a target that IS visible keeps the straight chord
    t2 = np.array([0.21, 0.33, 0])
    d, ts, te = _query(orc, V, corners, src, [t2])
    assert abs(d[0] - np.linalg.norm(t2 - src)) < 1e-13


def test_disconnected_patch_reports_unreachable():
    # two separate planar sheets in one mesh: the global-mesh branch returns dist < 0 (CGAL returns (-1, end))
    V1, F1 = meshes.plane_grid(2, 2)
    V2, F2 = meshes.plane_grid(2, 2)
    V = np.concatenate([V1, V2 + [3.0, 0, 0]])
    F = np.concatenate([F1, F2 + len(V1)]).astype(np.int32)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    b = np.array([0.3, 0.3, 0.4])
    d, ts, te, tie, _ = orc.distance(0, b, [len(F1) + 1, 1], np.array([b, b]))
    assert d[0] < 0 and d[1] > 0
    # with submeshing on the reference substitutes the sentinel (2*maxDist, (0,0,1)) (triangulatedMeshSpace.cpp:198-203)
    orc.set_submeshing(True, 0.7)
    d, ts, te, tie, _ = orc.distance(0, b, [len(F1) + 1, 1], np.array([b, b]), threshold=10.0)
    assert d[0] == 1.4 and np.array_equal(ts[0], [0, 0, 1]) and np.array_equal(te[0], [0, 0, 1])


# ---------------------------------------------------------------------------------------------- (7)
def _golden_mesh(name):
    import sys

    from helpers import GOLDEN

    sys.path.insert(0, GOLDEN)
    from make_golden import golden_mesh

    return golden_mesh(name)


def test_against_stored_exhaustive_unfolding():
    """tests/golden/bruteforce_geodesics.npz holds the answers of the independent exhaustive-unfolding checker
    (minutes of CPU; produced by tests/golden/make_golden.py) on 7 small meshes x 3 sources x 13 targets."""
    import os

    from helpers import GOLDEN

    g = np.load(os.path.join(GOLDEN, "bruteforce_geodesics.npz"))
    assert len(g["names"]) == 7
    worst = 0.0
    for name in g["names"]:
        V, F = _golden_mesh(str(name))
        corners = meshes.reference_corners(F)
        orc = Oracle(V, corners)
        face, bary = g[name + "/face"], g[name + "/bary"]
        for s in range(g[name + "/D"].shape[0]):
            tf, tb = np.delete(face, s), np.delete(bary, s, 0)
            d, ts, te, tie, _ = orc.distance(face[s], bary[s], tf, tb)
            D, TS, TE = g[name + "/D"][s], g[name + "/TS"][s], g[name + "/TE"][s]
            assert np.all(np.isfinite(D))
            worst = max(worst, float(np.max(np.abs(d - D) / D)))
            ok = tie == 0
            assert np.max(np.abs(ts[ok] - TS[ok])) < 1e-9
            assert np.max(np.abs(te[ok] - TE[ok])) < 1e-9
    assert worst < 1e-12


@pytest.mark.parametrize("name", ["cube1", "lshape"])
def test_against_live_exhaustive_unfolding(name):
    """The checker itself is exercised live on the two cheapest cases (different seed from the fixture)."""
    V, F = _golden_mesh(name)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    bg = BruteGeodesic(V, corners)
    rng = np.random.default_rng(23)
    face, bary = random_positions(len(F), 8, rng)
    bary = np.clip(bary, 0.02, None)
    bary /= bary.sum(1, keepdims=True)
    d, ts, te, tie, _ = orc.distance(face[0], bary[0], face[1:], bary[1:])
    D, TS, TE = bg.solve(face[0], bary[0], face[1:], bary[1:], depth=14)
    assert np.max(np.abs(d - D) / D) < 1e-12
    ok = tie == 0
    assert np.max(np.abs(ts[ok] - TS[ok])) < 1e-9 and np.max(np.abs(te[ok] - TE[ok])) < 1e-9


# ---------------------------------------------------------------------------------------------- (4)
def test_metric_properties_on_icosphere():
    V, F = meshes.icosphere(6)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    rng = np.random.default_rng(7)
    n = 24
    face, bary = random_positions(len(F), n, rng)
    P = orc.euclidean(face, bary)
    D = np.zeros((n, n))
    TS = np.zeros((n, n, 3))
    TE = np.zeros((n, n, 3))
    for i in range(n):
        d, ts, te, tie, _ = orc.distance(face[i], bary[i], face, bary)
        D[i], TS[i], TE[i] = d, ts, te
    off = ~np.eye(n, dtype=bool)
    assert np.max(np.abs(D - D.T)[off] / D[off]) < 1e-12                      # d(a,b) = d(b,a)
    assert np.max(np.abs(TS + np.transpose(TE, (1, 0, 2)))[off]) < 1e-9       # start(a->b) = -end(b->a)
    E = np.linalg.norm(P[:, None] - P[None], axis=2)
    assert np.all(D[off] >= E[off] * (1 - 1e-14))                            # d >= Euclidean
    for k in range(n):                                                       # triangle inequality
        assert np.all(D <= D[:, [k]] + D[[k], :] + 1e-12)
    assert np.allclose(np.linalg.norm(TS[off], axis=1), 1, atol=1e-12)
    # tangents lie in the plane of their face
    nrm = _unit(np.cross(V[corners[:, 1]] - V[corners[:, 0]], V[corners[:, 2]] - V[corners[:, 0]]))
    assert np.max(np.abs(np.einsum("ijk,ik->ij", TS, nrm[face]))[off]) < 1e-12
    assert np.max(np.abs(np.einsum("ijk,jk->ij", TE, nrm[face]))[off]) < 1e-12


def test_icosphere_converges_to_great_circle():
    errs = []
    for nu in (8, 16, 32):
        V, F = meshes.icosphere(nu)
        corners = meshes.reference_corners(F)
        orc = Oracle(V, corners)
        rng = np.random.default_rng(5)
        face, bary = random_positions(len(F), 12, rng)
        P = orc.euclidean(face, bary)
        d, *_ = orc.distance(face[0], bary[0], face[1:], bary[1:])
        U = _unit(P)
        arc = np.arccos(np.clip(U[1:] @ U[0], -1, 1))
        errs.append(float(np.max(np.abs(d - arc) / arc)))
    assert errs[0] < 2e-2 and errs[1] < errs[0] / 2.5 and errs[2] < errs[1] / 2.5  # O(h^2)


# ---------------------------------------------------------------------------------------------- (5)
def test_full_mesh_vs_patch_self_consistency():
    """meshTesting.cpp:201-203: distances computed on the per-source patch equal the full-mesh ones for
    targets closer than the cut-off."""
    V, F = meshes.icosphere(10)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    rng = np.random.default_rng(3)
    face, bary = random_positions(len(F), 300, rng)
    P = orc.euclidean(face, bary)
    cutoff = 0.5
    checked = 0
    for s in range(0, 300, 30):
        e = np.linalg.norm(P - P[s], axis=1)
        sel = np.where((e < cutoff) & (np.arange(300) != s))[0]
        if len(sel) == 0:
            continue
        orc.set_submeshing(False)
        dfull, tsf, *_ = orc.distance(face[s], bary[s], face[sel], bary[sel])
        orc.set_submeshing(True, cutoff)
        dsub, tss, *_ = orc.distance(face[s], bary[s], face[sel], bary[sel], threshold=float(e[sel].max()))
        near = dfull < float(e[sel].max())
        assert np.max(np.abs(dfull[near] - dsub[near])) < 1e-12
        assert np.all(dsub >= dfull - 1e-12)  # a patch can only lengthen paths
        checked += int(near.sum())
    assert checked > 10
