"""CPU oracle: cell list, patch, walker, pair forces and updaters (the non-geodesic rows of SURVEY.md §8(a)).

Ordering / bit contracts are checked against literal pure-Python transcriptions of the reference
(tests/pyref.py); the walker and the updaters by the properties the reference's own debugging mains
print (meshTesting.cpp, flatSpaceSimulation.cpp) and by numpy restatements of the formulas."""
from __future__ import annotations

import math

import numpy as np
import pytest

import pyref
from curvedspacesim_b200 import meshes
from helpers import csr_rows, interaction_range, make_state, random_positions, random_velocities
from oracle_binding import Oracle, force_params

import os

from helpers import GOLDEN as GOLDEN_DIR


def _setup(V, F, N, seed=13377, area_fraction=0.9):
    corners, face, bary, vel = make_state(V, F, N, seed=seed)
    orc = Oracle(V, corners)
    mn, mx, area = orc.mesh_info()
    rc = interaction_range(area, N, area_fraction)
    orc.set_submeshing(True, rc)
    orc.set_state(face, bary, vel)
    return orc, corners, face, bary, vel, rc, (mn, mx, area)


# --------------------------------------------------------------------------------- mesh / types
def test_mesh_info_bbox_contains_origin_and_area():
    # triangulatedMeshSpace::updateMeshSpanAndTree seeds the bounding box with 0 (q1)
    V, F = meshes.icosphere(4)
    V = V + np.array([5.0, 0.0, 0.0])
    orc = Oracle(V, meshes.reference_corners(F))
    mn, mx, area = orc.mesh_info()
    assert mn[0] == 0.0 and abs(mx[0] - 6.0) < 1e-12
    assert abs(area - meshes.face_areas(V, F).sum()) < 1e-12


def test_adjacency_matches_numpy_builder():
    V, F = meshes.torus(12, 8, jitter=0.1)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    adj, adjk = orc.adjacency()
    a2, k2 = meshes.build_adjacency(corners)
    assert np.array_equal(adj, a2) and np.array_equal(adjk, k2)
    # edge k of f is edge adjk of adj: shared vertices agree
    for f in range(0, len(F), 7):
        for k in range(3):
            g, kk = adj[f, k], adjk[f, k]
            assert {corners[f, (k + 1) % 3], corners[f, (k + 2) % 3]} == {corners[g, (kk + 1) % 3], corners[g, (kk + 2) % 3]}


def test_euclidean_is_normalised_barycentric_sum():
    V, F = meshes.icosphere(5)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    face, bary = random_positions(len(F), 50, np.random.default_rng(0))
    bary = bary * 1.7  # SMSP::point normalises by the sum
    P = orc.euclidean(face, bary)
    exp = np.einsum("ik,ikd->id", bary, V[corners[face]]) / bary.sum(1, keepdims=True)
    assert np.max(np.abs(P - exp)) < 1e-15


# --------------------------------------------------------------------------------- cell list (row 12)
@pytest.mark.parametrize("mesh,N", [("ico", 300), ("torus", 400)])
def test_candidate_lists_match_literal_transcription(mesh, N):
    V, F = meshes.icosphere(8) if mesh == "ico" else meshes.torus(40, 16, jitter=0.2)
    orc, corners, face, bary, vel, rc, (mn, mx, area) = _setup(V, F, N)
    off, idx, maxd = orc.candidates(rc)
    P = orc.euclidean(face, bary)
    ref, refmax = pyref.candidate_lists(P, mn, mx, rc)
    rows = csr_rows(off, idx)
    for i in range(N):
        assert list(rows[i]) == ref[i]          # membership AND order are part of the contract
        assert maxd[i] == refmax[i]             # bit-equal R = sqrt(max d^2)
    # the stencil loses nobody: same sets as the brute-force Euclidean ball
    D2 = ((P[:, None] - P[None]) ** 2).sum(-1)
    for i in range(0, N, 17):
        ball = set(np.where((D2[i] < rc * rc) & (np.arange(N) != i))[0].tolist())
        assert ball == set(ref[i])


def test_grid_size_rule():
    V, F = meshes.torus(30, 12)
    orc, *_ , rc, (mn, mx, area) = _setup(V, F, 100)
    n, cs = orc.cell_grid(rc)
    for d in range(3):
        assert n[d] == max(1, int(math.floor((mx[d] - mn[d]) / rc)))
        assert cs[d] == (mx[d] - mn[d]) / n[d]


# --------------------------------------------------------------------------------- patch (row 15)
def test_patch_face_set_matches_literal_dfs():
    V, F = meshes.torus(40, 16, jitter=0.2)
    orc, corners, face, bary, vel, rc, _ = _setup(V, F, 400)
    adj, _ = orc.adjacency()
    off, idx, maxd = orc.candidates(rc)
    P = orc.euclidean(face, bary)
    rows = csr_rows(off, idx)
    sizes = []
    for i in range(0, 400, 9):
        if len(rows[i]) == 0:
            continue
        R = min(rc, maxd[i])
        got = orc.patch(face[i], bary[i], face[rows[i]], R)
        exp = pyref.patch_faces(V, corners, adj, face[i], P[i], face[rows[i]], R)
        assert len(set(got.tolist())) == len(got)
        assert set(got.tolist()) == exp
        sizes.append(len(got))
    assert max(sizes) > 4


def test_patch_early_exits():
    V, F = meshes.icosphere(6)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    adj, _ = orc.adjacency()
    b = np.array([0.3, 0.3, 0.4])
    assert orc.patch(5, b, [5, 5], 10.0).tolist() == [5]                               # submesher.cpp:79-80
    got = orc.patch(5, b, [5, int(adj[5, 1])], 10.0)                                    # :97-98
    assert set(got.tolist()) == {5, *adj[5].tolist()}
    # a far goal face that the flood fill cannot reach is appended (:143-144)
    got = orc.patch(5, b, [len(F) - 1], 1e-3)
    assert set(got.tolist()) == {5, *adj[5].tolist(), len(F) - 1}


# --------------------------------------------------------------------------------- walker (rows 4-8)
def test_walker_on_a_plane_is_a_straight_line_and_identity_transport():
    V, F = meshes.plane_grid(12, 12, 1.0, 1.0)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    rng = np.random.default_rng(2)
    face, bary = random_positions(len(F), 200, rng)
    P = orc.euclidean(face, bary)
    disp = rng.standard_normal((200, 3)) * 0.2
    disp[:, 2] = rng.standard_normal(200)  # the normal component is projected away (THRESHOLD, :462-465)
    tgt = P + disp * [1, 1, 0]
    inside = np.all((tgt[:, :2] > 0.01) & (tgt[:, :2] < 0.99), axis=1)
    vec = rng.standard_normal((200, 2, 3)) * [1, 1, 0]
    f2, b2, d2, v2, flags, cr = orc.transport(face[inside], bary[inside], disp[inside], vec[inside])
    P2 = orc.euclidean(f2, b2)
    assert np.all(flags == 0) and cr.max() >= 3
    assert np.max(np.abs(P2 - tgt[inside])) < 1e-9        # 1e-11 / 1e-13 clamps perturb ~1e-11 per crossing
    assert np.max(np.abs(v2 - vec[inside])) < 1e-12


def test_walker_conserves_norm_stays_in_plane_and_is_reversible():
    V, F = meshes.icosphere(10)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    rng = np.random.default_rng(4)
    n = 300
    face, bary = random_positions(len(F), n, rng)
    vel = random_velocities(V, corners, face, 1.0, rng)
    vel /= np.linalg.norm(vel, axis=1, keepdims=True)
    L = rng.uniform(0.05, 1.5, n)
    disp = vel * L[:, None]
    f2, b2, d2, v2, flags, cr = orc.transport(face, bary, disp, vel[:, None, :])
    ok = flags == 0
    assert ok.mean() > 0.99 and cr.max() > 10
    v2 = v2[:, 0]
    nrm = np.cross(V[corners[:, 1]] - V[corners[:, 0]], V[corners[:, 2]] - V[corners[:, 0]])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    assert np.max(np.abs(np.linalg.norm(v2[ok], axis=1) - 1)) < 1e-13          # parallel transport is an isometry
    assert np.max(np.abs(np.einsum("ij,ij->i", v2, nrm[f2]))[ok]) < 1e-12      # vector stays tangent
    # walk back along the transported direction: returns to the start
    f3, b3, d3, v3, fl3, _ = orc.transport(f2, b2, -v2 * L[:, None], v2[:, None, :])
    ok &= fl3 == 0
    P0, P3 = orc.euclidean(face, bary), orc.euclidean(f3, b3)
    assert np.max(np.abs(P0 - P3)[ok]) < 1e-8
    assert np.array_equal(f3[ok], face[ok]) or (f3[ok] == face[ok]).mean() > 0.98  # start points near an edge may land next door


def test_walking_the_start_tangent_reaches_the_target():
    """Ties the geodesic engine to the walker: x_target = exp_source(d * startTangent)."""
    V, F = meshes.icosphere(8)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    rng = np.random.default_rng(6)
    face, bary = random_positions(len(F), 40, rng)
    P = orc.euclidean(face, bary)
    d, ts, te, tie, _ = orc.distance(face[0], bary[0], face[1:], bary[1:])
    sel = np.where((d < 1.2) & (tie == 0))[0]
    assert len(sel) > 5
    sf = np.full(len(sel), face[0], np.int32)
    sb = np.repeat(bary[:1], len(sel), 0)
    f2, b2, d2, v2, flags, _ = orc.transport(sf, sb, ts[sel] * d[sel, None], ts[sel][:, None, :])
    ok = flags == 0
    P2 = orc.euclidean(f2, b2)
    assert np.max(np.abs(P2 - P[1:][sel])[ok]) < 1e-8
    assert np.max(np.abs(v2[:, 0] - te[sel])[ok]) < 1e-8  # transported start tangent = end tangent


def test_walker_flags_vertex_crossing():
    V, F = meshes.plane_grid(2, 2, 1.0, 1.0)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    # aim exactly through the centre vertex (0.5, 0.5)
    f0 = 0
    c = V[corners[f0]]
    b = np.array([0.5, 0.25, 0.25])
    p = b @ c
    target_vertex = np.array([0.5, 0.5, 0.0])
    disp = (target_vertex - p) * 2.0
    f2, b2, d2, v2, flags, cr = orc.transport([f0], [b], [disp])
    assert flags[0] & 1  # WALK_VERTEX: flagged and counted, excluded from trajectory parity


# --------------------------------------------------------------------------------- forces (row 9)
@pytest.mark.parametrize("pot", ["harmonic", "gaussian"])
def test_pair_forces_and_energy_match_numpy(pot):
    V, F = meshes.icosphere(12)
    orc, corners, face, bary, vel, rc, _ = _setup(V, F, 250)
    if pot == "harmonic":
        kind, params = force_params("harmonic", k=2.5, sigma=rc)
    else:
        kind, params = force_params("gaussian", alpha=1.3, sigma=0.5 * rc, range=rc)
    off, idx, d, ts, te = orc.find_neighbors(rc)
    frc = orc.compute_forces(kind, params)
    exp = np.zeros((250, 3))
    e = 0.0
    for i in range(250):
        for j in range(off[i], off[i + 1]):
            if pot == "harmonic":
                if d[j] <= rc:
                    exp[i] += -2.5 * (rc - d[j]) * ts[j]
                if d[j] < rc:
                    e += 0.5 * 2.5 * (rc - d[j]) ** 2
            else:
                s = 0.5 * rc
                exp[i] += -(d[j] * 1.3 * math.exp(-d[j] ** 2 / (2 * s * s)) / (math.sqrt(2 * math.pi) * s * math.sqrt(s))) * ts[j]
                e += 1.3 * math.exp(-d[j] ** 2 / (2 * s * s)) / (math.sqrt(2 * math.pi) * s)
    assert np.max(np.abs(frc - exp)) < 1e-13 * max(1.0, np.abs(exp).max())
    assert abs(orc.compute_energy(kind, params) - e) < 1e-12 * max(1.0, abs(e))
    # zero=False accumulates (simulation::computeForces passes zero only for the first force)
    frc2 = orc.compute_forces(kind, params, zero=False)
    assert np.max(np.abs(frc2 - 2 * exp)) < 1e-12 * max(1.0, np.abs(exp).max())


def test_all_to_all_candidates_without_cell_list():
    V, F = meshes.icosphere(3)
    corners, face, bary, vel = make_state(V, F, 12)
    orc = Oracle(V, corners)
    orc.set_options(use_cell_list=False)
    orc.set_state(face, bary, vel)
    off, idx, d, ts, te = orc.find_neighbors(0.5)
    for i, row in enumerate(csr_rows(off, idx)):
        assert row.tolist() == [j for j in range(12) if j != i]  # baseNeighborStructure.cpp:17-36
    assert np.all(d > 0)


# --------------------------------------------------------------------------------- updaters (rows 2-2d)
def test_nve_conserves_energy_and_threads_are_bitwise_equal():
    V, F = meshes.icosphere(12)
    orc, corners, face, bary, vel, rc, _ = _setup(V, F, 300)
    kind, params = force_params("harmonic", k=1.0, sigma=rc)

    def total_energy(o):
        _, _, v, _ = o.get_state()
        return o.compute_energy(kind, params) * 0.5 + 0.5 * float((v * v).sum())  # each pair is counted twice

    orc.compute_forces(kind, params)
    e0 = total_energy(orc)
    orc.run_nve(kind, params, 0.002, 100)
    e1 = total_energy(orc)
    assert abs(e1 - e0) / abs(e0) < 2e-4
    c = orc.counters()
    assert c["crossings"] > 0 and c["nohit"] == 0 and c["nan"] == 0
    s1 = orc.get_state()
    # the same run sharded over 4 worker threads like mpiModel::determineIndexBounds is bitwise identical
    orc4 = Oracle(V, corners)
    orc4.set_submeshing(True, rc)
    orc4.set_options(True, False, 4)
    orc4.set_state(face, bary, vel)
    orc4.compute_forces(kind, params)
    orc4.run_nve(kind, params, 0.002, 100)
    for a, b in zip(s1, orc4.get_state()):
        assert np.array_equal(a, b)


def test_gradient_descent_and_fire_reduce_the_energy():
    V, F = meshes.icosphere(10)
    orc, corners, face, bary, vel, rc, _ = _setup(V, F, 200)
    kind, params = force_params("harmonic", k=1.0, sigma=rc)
    e0 = orc.compute_energy(kind, params)
    orc.run_gd(kind, params, 0.05, 30)
    e1 = orc.compute_energy(kind, params)
    assert e1 < e0
    orc.set_state(face, bary, np.zeros_like(vel))
    p = np.array([60, 0.01, 0.99, 0.1, 1e-5, 1.1, 0.95, 0.9, 4, 1e-12, 0.0])
    orc.fire_init(p, dt0=0.01, alpha0=0.99)
    _, out = orc.run_fire(kind, params)
    assert out[0] == 60  # ran to the iteration cap (fireMinimization.cpp:3-21)
    assert orc.compute_energy(kind, params) < 0.5 * e0


def test_nose_hoover_reaches_the_target_temperature():
    """tau = 1 as in curvedSpaceNVTSim.cpp:101.  (The first bath mass 2 (Ndof-1) T tau^2 makes the chain's
    set point KE = (Ndof-1) T tau^2, so the target is T only for tau = 1; the oracle restates that as coded.)"""
    V, F = meshes.icosphere(8)
    N = 100
    orc, corners, face, bary, vel, rc, _ = _setup(V, F, N)
    kind, params = force_params("harmonic", k=1.0, sigma=rc)
    orc.compute_forces(kind, params)
    orc.nvt_init(0.01, 0.5, tau=1.0, M=2)
    bath, ke, scale = orc.nvt_state()
    assert bath.shape == (3, 4) and bath[0, 3] == 2.0 * (N - 1) * 0.5 and bath[1, 3] == 0.5  # noseHooverNVT.cpp:28-36
    temps = []
    for _ in range(20):
        orc.run_nvt(kind, params, 50)
        _, _, v, _ = orc.get_state()
        temps.append(float((v * v).sum()) / (2 * N))  # getTemperatureFromKE, noseHooverNVT.cpp:141-150
    assert temps[0] < 0.25 and abs(temps[-1] - 0.5) < 0.05
    assert abs(orc.nvt_state()[1] - temps[-1] * N) < 1e-9  # kineticEnergy bookkeeping of propagateChain


# --------------------------------------------------------------------------------- regression fixture
def test_oracle_matches_its_committed_regression_fixture():
    """tests/golden/oracle_regression.npz (made by tests/golden/make_golden.py) is what the GPU tests also
    compare against on the box; here it guards the oracle itself against silent changes."""
    import os
    import sys

    from helpers import GOLDEN

    sys.path.insert(0, GOLDEN)
    from make_golden import golden_mesh

    g = np.load(os.path.join(GOLDEN, "oracle_regression.npz"))
    for key in g["names"]:
        key = str(key)
        name, N = key.split("_N")[0], int(g[key + "/N"])
        V, F = golden_mesh(name)
        corners, face, bary, vel = make_state(V, F, N)
        rc, kind, params = float(g[key + "/rc"]), int(g[key + "/kind"]), g[key + "/params"]
        orc = Oracle(V, corners)
        orc.set_submeshing(True, rc)
        orc.set_state(face, bary, vel)
        off, idx, d, ts, te = orc.find_neighbors(rc)
        assert np.array_equal(off, g[key + "/off"]) and np.array_equal(idx, g[key + "/idx"])
        assert np.array_equal(d, g[key + "/dist"]) and np.array_equal(ts, g[key + "/ts"]) and np.array_equal(te, g[key + "/te"])
        assert np.array_equal(orc.compute_forces(kind, params), g[key + "/frc"])
        orc.run_nve(kind, params, 0.01, 50)
        f2, b2, v2, fr2 = orc.get_state()
        assert np.array_equal(f2, g[key + "/face50"]) and np.array_equal(b2, g[key + "/bary50"]) and np.array_equal(v2, g[key + "/vel50"])


# --------------------------------------------------------------------------------- open meshes
def _open_plane():
    V, F = meshes.plane_grid(8, 8, 1.0, 1.0)
    return V, F, meshes.reference_corners(F)


def test_absorbing_boundary_stops_on_the_edge_and_projects_vectors():
    """absorbingOpenMeshSpace::updateAtBoundaryEdge (absorbingOpenMeshSpace.cpp:2-24): the particle stops where its path
    meets the border; transported vectors lose their outward component (triangulatedMeshSpace.cpp:411-426)."""
    V, F, corners = _open_plane()
    orc = Oracle(V, corners)
    orc.set_boundary(1)
    from test_oracle_geodesic import _locate

    f, b = _locate(V, corners, np.array([0.52, 0.47, 0.0]))
    disp = np.array([0.8, 0.1, 0.0])
    vecs = np.array([[[1.0, 0.3, 0.0], [-1.0, 0.2, 0.0]]])
    f2, b2, d2, v2, flags, cr = orc.transport([f], [b], [disp], vecs)
    P = orc.euclidean(f2, b2)[0]
    y_hit = 0.47 + 0.1 * (1.0 - 0.52) / 0.8
    assert flags[0] == 16 and abs(P[0] - 1.0) < 1e-9 and abs(P[1] - y_hit) < 1e-9
    assert np.allclose(v2[0, 0], [0.0, 0.3, 0.0], atol=1e-12)       # pointed over the boundary: outward part removed
    assert np.allclose(v2[0, 1], [-1.0, 0.2, 0.0], atol=1e-12)      # pointed inward: untouched
    # the closed space only flags the event and stops (the reference throws there)
    orc.set_boundary(0)
    f3, b3, _, v3, fl3, _ = orc.transport([f], [b], [disp], vecs)
    assert fl3[0] == 16 and np.allclose(v3[0, 0], [1.0, 0.3, 0.0], atol=1e-12)
    # a displacement that stays inside is unaffected by the boundary rule
    orc.set_boundary(1)
    f4, b4, _, _, fl4, _ = orc.transport([f], [b], [np.array([0.2, 0.1, 0.0])])
    assert fl4[0] == 0 and np.allclose(orc.euclidean(f4, b4)[0], [0.72, 0.57, 0.0], atol=1e-9)


def test_tangential_boundary_slides_along_the_edge():
    """tangentialOpenMeshSpace::updateAtBoundaryEdge (tangentialOpenMeshSpace.cpp:3-42): the displacement is redirected
    along the border edge in the direction that overlaps it, with the length of the displacement the face iteration
    started with times that overlap."""
    V, F, corners = _open_plane()
    orc = Oracle(V, corners)
    orc.set_boundary(2)
    from test_oracle_geodesic import _locate

    f, b = _locate(V, corners, np.array([0.52, 0.47, 0.0]))
    for dy in (0.1, -0.1):
        disp = np.array([0.8, dy, 0.0])
        f2, b2, d2, v2, flags, cr = orc.transport([f], [b], [disp], np.array([[[1.0, 0.3, 0.0]]]))
        P = orc.euclidean(f2, b2)[0]
        y_hit = 0.47 + dy * (1.0 - 0.52) / 0.8
        assert flags[0] == 16 and abs(P[0] - 1.0) < 1e-9
        assert (P[1] - y_hit) * dy > 0 and abs(P[1] - y_hit) < abs(dy)   # slid in the direction of the tangential component
        assert np.allclose(v2[0, 0], [0.0, 0.3, 0.0], atol=1e-12)
    # an NVE run on an open sheet keeps every particle on the mesh (short enough that no two particles have yet slid into
    # the same corner of the sheet: coincident particles have no tangent, in the reference as here)
    N = 60
    corners, face, bary, vel = make_state(V, F, N)
    orc.set_submeshing(True, 0.2)
    orc.set_state(face, bary, vel * 2)
    kind, params = force_params("harmonic", k=1.0, sigma=0.2)
    orc.compute_forces(kind, params)
    orc.run_nve(kind, params, 0.01, 40)
    f3, b3, v3, _ = orc.get_state()
    P = orc.euclidean(f3, b3)
    c = orc.counters()
    assert c["nan"] == 0 and c["border"] > 100
    assert np.all(b3 >= 0) and np.all((P[:, :2] >= -1e-9) & (P[:, :2] <= 1 + 1e-9))


# --------------------------------------------------------------------------------- observables
def test_stress_and_temperature_match_a_literal_numpy_evaluation():
    """simulation::computeMonodisperseStress (simulation.cpp:104-173) and noseHooverNVT::getTemperatureFromKE
    (noseHooverNVT.cpp:141-150) restated with numpy loops over the oracle's own neighbour lists."""
    V, F = meshes.icosphere(8)
    N = 150
    corners, face, bary, vel = make_state(V, F, N)
    orc = Oracle(V, corners)
    _, _, area = orc.mesh_info()
    rc = interaction_range(area, N, 2.0)
    orc.set_submeshing(True, rc)
    orc.set_state(face, bary, vel)
    for name, kw in (("harmonic", dict(k=1.0, sigma=rc)), ("gaussian", dict(alpha=1.0, sigma=0.5 * rc, range=rc))):
        kind, params = force_params(name, **kw)
        S = orc.compute_stress(kind, params)
        off, idx, d, ts, _ = orc.find_neighbors(rc)
        fdr, vv = np.zeros((3, 3)), np.zeros((3, 3))
        for i in range(N):
            for q in range(off[i], off[i + 1]):
                if name == "harmonic":
                    f = -1.0 * (rc - d[q]) * ts[q] if d[q] <= rc else np.zeros(3)
                else:
                    sg = 0.5 * rc
                    f = -(d[q] * np.exp(-d[q] ** 2 / (2 * sg * sg)) / (np.sqrt(2 * np.pi) * sg * np.sqrt(sg))) * ts[q]
                fdr += np.outer(f, ts[q])
                vv += np.outer(vel[i], vel[i])
        ref = (N / area) * vv / (2 * N) + fdr / (4 * area * N)
        assert len(idx) > N and np.max(np.abs(S - ref)) < 1e-12 * np.abs(ref).max()
        assert abs(np.trace(S)) > 0
    assert abs(orc.temperature() - (vel ** 2).sum() / (2 * N)) < 1e-14


# --------------------------------------------------------------------------------- R^3 -> mesh positions
def _closest_point_distances_numpy(P, V, corners):
    """Independent restatement (no Voronoi regions): the closest point of a triangle is the projection onto its plane when
    that lies inside, otherwise the closest point of one of its three edges.  Returns [n_points, n_faces] distances."""
    a, b, c = V[corners[:, 0]], V[corners[:, 1]], V[corners[:, 2]]
    out = np.empty((len(P), len(corners)))
    nrm = np.cross(b - a, c - a)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)

    def seg(p, s0, s1):
        d = s1 - s0
        t = np.clip(np.einsum("ij,ij->i", p - s0, d) / np.einsum("ij,ij->i", d, d), 0, 1)
        return np.linalg.norm(p - (s0 + t[:, None] * d), axis=1)

    for i, p in enumerate(P):
        pp = np.broadcast_to(p, a.shape)
        h = np.einsum("ij,ij->i", pp - a, nrm)
        q = pp - h[:, None] * nrm
        inside = np.ones(len(a), bool)
        for s0, s1 in ((a, b), (b, c), (c, a)):
            inside &= np.einsum("ij,ij->i", np.cross(s1 - s0, q - s0), nrm) >= 0
        de = np.minimum(np.minimum(seg(pp, a, b), seg(pp, b, c)), seg(pp, c, a))
        out[i] = np.where(inside, np.abs(h), de)
    return out


def test_locate_finds_the_closest_face_and_clamps_like_the_reference():
    # simpleModel::R3PositionsToMeshPositions (simpleModel.cpp:136-154) + clampBarycentricCoordinatesToFace (:114-134)
    V, F = meshes.torus(24, 12, R=3.0, r=1.0, jitter=0.2, seed=5)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    rng = np.random.default_rng(11)
    face0, bary0 = random_positions(len(F), 150, rng)
    on = orc.euclidean(face0, bary0)
    nrm = np.cross(V[corners[face0, 1]] - V[corners[face0, 0]], V[corners[face0, 2]] - V[corners[face0, 0]])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    slightly_off = on + 1e-9 * rng.standard_normal((150, 1)) * nrm          # what the reference expects as input
    far = 6.0 * rng.standard_normal((40, 3))                                 # anywhere, incl. outside the bounding box
    P = np.concatenate([slightly_off, far, V[:20], 0.5 * (V[corners[:10, 0]] + V[corners[:10, 1]])])
    f, b = orc.locate(P)
    D = _closest_point_distances_numpy(P, V, corners)
    x = orc.euclidean(f, b)
    d_found = np.linalg.norm(P - x, axis=1)
    assert np.all(np.abs(d_found - D.min(axis=1)) < 1e-10)                   # nothing on the mesh is closer
    assert np.all(np.abs(D[np.arange(len(P)), f] - D.min(axis=1)) < 1e-10)   # and the face returned attains it
    assert np.array_equal(f[:150], face0)                                    # interior points come back to their face
    assert np.max(np.abs(b[:150] - bary0)) < 1e-7
    # weights: all >= the clamp tolerance, and the reference's sequential renormalisation (sum == 1 only to ~1e-14)
    assert b.min() >= 1e-14 * (1 - 1e-12) and np.max(np.abs(b.sum(axis=1) - 1)) < 1e-12
    # a mesh vertex lies on several faces at distance 0: the lowest face index wins (documented tie rule)
    for k in range(20):
        incident = np.where((corners == k).any(axis=1))[0]
        assert f[190 + k] == incident.min()
    # literal clamp arithmetic on a corner point: weights (1, 0, 0) -> (1, tol, tol) divided one after the other
    fv, bv = orc.locate(V[corners[7, 0]][None, :])
    w = [1.0, 1e-14, 1e-14] if fv[0] == 7 else None
    if w is not None:
        w[0] = w[0] / (w[0] + w[1] + w[2])
        w[1] = w[1] / (w[0] + w[1] + w[2])
        w[2] = w[2] / (w[0] + w[1] + w[2])
        assert np.array_equal(bv[0], np.array(w))


def test_locate_matches_its_committed_fixture():
    """tests/golden/locate_regression.npz (tests/golden/make_golden.py locate) is what the GPU test also compares with."""
    import sys

    sys.path.insert(0, GOLDEN_DIR)
    from make_golden import golden_mesh

    g = np.load(os.path.join(GOLDEN_DIR, "locate_regression.npz"))
    for name in g["names"]:
        name = str(name)
        V, F = golden_mesh(name)
        f, b = Oracle(V, meshes.reference_corners(F)).locate(g[name + "/xyz"])
        assert np.array_equal(f, g[name + "/face"]) and np.array_equal(b, g[name + "/bary"])
