"""Generates the committed fixtures under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

The reference ships no golden vectors and cannot be built or imported here (CGAL/Boost/HDF5/MPI are
absent; it is C++), so there are two kinds of fixture, both produced by code in this repository:

  bruteforce_geodesics.npz  distances / start / end tangents from tests/bruteforce_geodesic.py, the
                            exhaustive-unfolding + Dijkstra checker that shares no code with the oracle or
                            the kernels (minutes of CPU, which is why the results are stored).  These PIN the
                            oracle (tests/test_oracle_geodesic.py) and the CUDA path (tests/test_gpu_parity.py).
  closed_form.npz           analytically known distances (plane, cube unfoldings, tetrahedron, L-shaped
                            notch) with the query points they belong to.
  oracle_regression.npz     outputs of the CPU oracle on small seeded configurations (neighbour lists,
                            distances, tangents, forces, state after 50 NVE steps).  A regression guard for the
                            oracle itself and a second, travel-able comparison target for the GPU tests.
  locate_regression.npz     R^3 points and the mesh positions the oracle's R3PositionsToMeshPositions restatement
                            assigns to them (the oracle is checked against an independent numpy closest-point
                            formulation in tests/test_oracle_model.py); same two roles.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
for p in (ROOT, TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)

from bruteforce_geodesic import BruteGeodesic  # noqa: E402
from curvedspacesim_b200 import meshes  # noqa: E402
from helpers import interaction_range, make_state, random_positions  # noqa: E402


def golden_mesh(name):
    """Meshes are regenerated from their name (deterministic generators), never stored."""
    if name == "icosphere2":
        return meshes.icosphere(2)
    if name == "icosphere3":
        return meshes.icosphere(3)
    if name == "torus8x5":
        return meshes.torus(8, 5, R=3.0, r=1.0, jitter=0.15, seed=3)
    if name == "torus10x6":
        return meshes.torus(10, 6, R=3.0, r=1.0, jitter=0.2, seed=5)
    if name == "cube1":
        return meshes.cube(1, 1.0)
    if name == "cube2":
        return meshes.cube(2, 1.0)
    if name == "lshape":
        V, F = meshes.plane_grid(4, 4, 1.0, 1.0)
        cen = V[F].mean(1)
        return V, F[~((cen[:, 0] > 0.5) & (cen[:, 1] > 0.5))].copy()
    if name == "icosphere16":
        return meshes.icosphere(16)
    if name == "torus60x24":
        return meshes.torus(60, 24, jitter=0.2)
    raise KeyError(name)


def brute_cases():
    out = {}
    names = ["icosphere2", "icosphere3", "torus8x5", "torus10x6", "cube1", "cube2", "lshape"]
    for name in names:
        V, F = golden_mesh(name)
        corners = meshes.reference_corners(F)
        rng = np.random.default_rng(11)
        n = 14
        face, bary = random_positions(len(F), n, rng)
        bary = np.clip(bary, 0.02, None)
        bary /= bary.sum(1, keepdims=True)
        bg = BruteGeodesic(V, corners)
        nsrc = 3
        D = np.zeros((nsrc, n - 1))
        TS = np.zeros((nsrc, n - 1, 3))
        TE = np.zeros((nsrc, n - 1, 3))
        for s in range(nsrc):
            tf, tb = np.delete(face, s), np.delete(bary, s, 0)
            D[s], TS[s], TE[s] = bg.solve(face[s], bary[s], tf, tb, depth=30)
            print(name, "source", s, "done", flush=True)
        out[name + "/face"], out[name + "/bary"] = face, bary
        out[name + "/D"], out[name + "/TS"], out[name + "/TE"] = D, TS, TE
    np.savez_compressed(os.path.join(HERE, "bruteforce_geodesics.npz"), names=np.array(names), **out)


def oracle_cases():
    from oracle_binding import Oracle, force_params

    out = {}
    names = []
    for name, N, pot in (("icosphere16", 200, "harmonic"), ("torus60x24", 500, "gaussian")):
        V, F = golden_mesh(name)
        corners, face, bary, vel = make_state(V, F, N)
        orc = Oracle(V, corners)
        _, _, area = orc.mesh_info()
        rc = interaction_range(area, N)
        if pot == "harmonic":
            kind, params = force_params("harmonic", k=1.0, sigma=rc)
        else:
            kind, params = force_params("gaussian", alpha=1.0, sigma=0.5 * rc, range=rc)
        orc.set_submeshing(True, rc)
        orc.set_state(face, bary, vel)
        off, idx, d, ts, te = orc.find_neighbors(rc)
        frc = orc.compute_forces(kind, params)
        orc.run_nve(kind, params, 0.01, 50)
        f2, b2, v2, fr2 = orc.get_state()
        key = name + "_N%d_%s" % (N, pot)
        names.append(key)
        for k, v in (("N", N), ("rc", rc), ("kind", kind), ("params", params), ("off", off), ("idx", idx), ("dist", d), ("ts", ts),
                     ("te", te), ("frc", frc), ("face50", f2), ("bary50", b2), ("vel50", v2), ("frc50", fr2)):
            out[key + "/" + k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "oracle_regression.npz"), names=np.array(names), **out)


def locate_points(V, corners, seed=3):
    """On-surface (nudged by 1e-9), clearly off-surface, far away, on vertices, on edge midpoints, the origin."""
    rng = np.random.default_rng(seed)
    face0, bary0 = random_positions(len(corners), 300, rng)
    on = np.einsum("nk,nkd->nd", bary0, V[corners[face0]])
    ext = float(np.abs(V).max())
    return np.concatenate([on + 1e-9 * rng.standard_normal((300, 3)), on + 0.05 * ext * rng.standard_normal((300, 3)),
                           4.0 * ext * rng.standard_normal((50, 3)), V[:50], 0.5 * (V[corners[:50, 1]] + V[corners[:50, 2]]),
                           np.zeros((1, 3))])


def locate_cases():
    from oracle_binding import Oracle

    out, names = {}, []
    for name in ("icosphere16", "torus60x24"):
        V, F = golden_mesh(name)
        corners = meshes.reference_corners(F)
        P = locate_points(V, corners)
        f, b = Oracle(V, corners).locate(P)
        names.append(name)
        out[name + "/xyz"], out[name + "/face"], out[name + "/bary"] = P, f, b
    np.savez_compressed(os.path.join(HERE, "locate_regression.npz"), names=np.array(names), **out)


def closed_form():
    """(mesh name, source point, target point, distance) tuples with analytically known answers."""
    rows = []
    # cube side 1: top -> +x face across the shared edge; top -> bottom through the +x side
    for src, t in (((0.81, 0.47, 1.0), (1.0, 0.52, 0.77)), ((0.81, 0.47, 1.0), (1.0, 0.31, 0.58))):
        rows.append(("cube2", src, t, math.hypot((1 - src[0]) + (1 - t[2]), src[1] - t[1])))
    src, t = (0.83, 0.47, 1.0), (0.79, 0.55, 0.0)
    rows.append(("cube2", src, t, math.hypot((1 - src[0]) + 1 + (1 - t[0]), src[1] - t[1])))
    # L-shaped notch: the path bends at the reflex boundary vertex (0.5, 0.5)
    s, t, b = np.array([0.9, 0.3, 0]), np.array([0.3, 0.9, 0]), np.array([0.5, 0.5, 0])
    rows.append(("lshape", tuple(s), tuple(t), float(np.linalg.norm(b - s) + np.linalg.norm(t - b))))
    s, t = np.array([0.9, 0.3, 0]), np.array([0.21, 0.33, 0])
    rows.append(("lshape", tuple(s), tuple(t), float(np.linalg.norm(t - s))))
    np.savez(os.path.join(HERE, "closed_form.npz"), mesh=np.array([r[0] for r in rows]), src=np.array([r[1] for r in rows]),
             tgt=np.array([r[2] for r in rows]), dist=np.array([r[3] for r in rows]))


if __name__ == "__main__":
    what = sys.argv[1:] or ["closed", "oracle", "brute", "locate"]
    if "closed" in what:
        closed_form()
    if "oracle" in what:
        oracle_cases()
    if "brute" in what:
        brute_cases()
    if "locate" in what:
        locate_cases()
