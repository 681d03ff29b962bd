"""GPU parity tests: the CUDA path, called through the C ABI (include/css_api.h), against the CPU oracle on
the same seeded inputs, against the committed golden fixtures, and - at BASELINE.json's full sizes -
through size-independent properties.

Bars (BASELINE.json north_star): bit-exact neighbour lists and face index after every displacement;
1e-9 relative geodesic distances and tangents; 1e-8 relative forces; 1e-6 trajectories after 1000 steps
excluding flagged vertex/edge-degenerate crossings.  The tolerances are written where they are applied."""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np
import pytest

from curvedspacesim_b200 import binding, meshes
from helpers import GOLDEN, csr_rows, interaction_range, make_state, random_positions, random_velocities
from oracle_binding import Oracle, force_params

pytestmark = pytest.mark.gpu

TOL_DIST = 1e-9     # relative, geodesic distances
TOL_TAN = 1e-9      # absolute on unit tangents
TOL_FORCE = 1e-8    # relative to the largest force component
TOL_TRAJ = 1e-6     # positions / velocities after 1000 steps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mesh(name):
    sys.path.insert(0, GOLDEN)
    from make_golden import golden_mesh

    return golden_mesh(name)


def _pair(V, F, N, pot="harmonic", gpu=None, area_fraction=0.9, want_end=True, seed=13377):
    corners, face, bary, vel = make_state(V, F, N, seed=seed)
    orc = Oracle(V, corners)
    _, _, area = orc.mesh_info()
    rc = interaction_range(area, N, area_fraction)
    if pot == "harmonic":
        kind, params = force_params("harmonic", k=1.0, sigma=rc)
    else:
        kind, params = force_params("gaussian", alpha=1.0, sigma=0.5 * rc, range=rc)
    orc.set_submeshing(True, rc)
    orc.set_state(face, bary, vel)
    ctx = gpu()
    ctx.set_mesh(V, corners)
    ctx.set_submeshing(True, rc)
    ctx.set_options(True, want_end)
    ctx.set_state(face, bary, vel)
    return orc, ctx, corners, face, bary, vel, rc, kind, params


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if len(a) else 0.0


CASES = [("icosphere16", 200, "harmonic"), ("torus60x24", 500, "gaussian"), ("icosphere40", 3000, "harmonic"),
         ("torus200x60", 5000, "harmonic")]


def _case_mesh(name):
    if name == "icosphere40":
        return meshes.icosphere(40)
    if name == "torus200x60":
        return meshes.torus(200, 60, R=3.0, r=1.0, jitter=0.2, seed=13377)
    return _mesh(name)


# ------------------------------------------------------------------------------ oracle parity, per step
@pytest.mark.parametrize("name,N,pot", CASES)
def test_neighbours_distances_tangents_forces(name, N, pot, gpu_ctx_factory):
    V, F = _case_mesh(name)
    orc, ctx, corners, face, bary, vel, rc, kind, params = _pair(V, F, N, pot, gpu_ctx_factory)
    assert np.array_equal(orc.euclidean(face, bary), ctx.euclidean(face, bary))            # bit-exact
    ctx.counters(reset=True)
    o_off, o_idx, o_d, o_ts, o_te = orc.find_neighbors(rc)
    g_off, g_idx, g_d, g_ts, g_te = ctx.find_neighbors(rc, want_end=True)
    assert np.array_equal(o_off, g_off) and np.array_equal(o_idx, g_idx)                  # bit-exact, ordered
    c, oc = ctx.counters(), orc.counters()
    assert c["patch_faces"] == oc["patch_faces"] and c["patch_verts"] == oc["patch_verts"]  # identical patches
    assert c["queries"] == len(o_idx) and c["overflow"] == 0 and c["kernels"] > 0
    assert len(o_idx) > N
    assert _rel(g_d, o_d) < TOL_DIST
    assert np.max(np.abs(g_ts - o_ts)) < TOL_TAN and np.max(np.abs(g_te - o_te)) < TOL_TAN
    f0 = orc.compute_forces(kind, params)
    ctx.compute_forces(kind, params)
    f1 = ctx.get_state()[3]
    assert np.max(np.abs(f0 - f1)) < TOL_FORCE * np.abs(f0).max()
    assert abs(ctx.compute_energy(kind, params) - orc.compute_energy(kind, params)) < 1e-9 * abs(orc.compute_energy(kind, params))
    assert ctx.counters()["overflow"] == 0


@pytest.mark.parametrize("name,N,pot", CASES[:3])
def test_move_is_bit_exact(name, N, pot, gpu_ctx_factory):
    V, F = _case_mesh(name)
    orc, ctx, corners, face, bary, vel, rc, kind, params = _pair(V, F, N, pot, gpu_ctx_factory)
    rng = np.random.default_rng(5)
    for scale in (0.3, 3.0):  # 3 rc: dozens of edge crossings per particle
        disp = rng.standard_normal((N, 3)) * scale * rc
        frc = rng.standard_normal((N, 3))
        orc.set_state(face, bary, vel, frc)
        ctx.set_state(face, bary, vel, frc)
        orc.move(disp.copy(), transport_force=True, transport_velocity=True)
        ctx.move(disp.copy(), transport_force=True, transport_velocity=True)
        of, ob, ov, ofr = orc.get_state()
        gf, gb, gv, gfr = ctx.get_state()
        assert np.array_equal(orc.walk_flags(), ctx.walk_flags())
        ok = orc.walk_flags() == 0
        assert ok.mean() > 0.99
        assert np.array_equal(of[ok], gf[ok])                                              # face index: bit-exact
        assert np.array_equal(ob[ok], gb[ok]) and np.array_equal(ov[ok], gv[ok]) and np.array_equal(ofr[ok], gfr[ok])
    assert ctx.counters()["crossings"] == orc.counters()["crossings"]


def test_config1_like_trajectory_1000_steps(gpu_ctx_factory):
    """BASELINE.json configs[0] shape (closed sphere of ~5k faces, N=100, harmonic, NVE dt=0.01, 1000 steps) on the
    synthetic icosphere-16 (5120 faces; the reference's sphere_radius1.off has 4860)."""
    V, F = _mesh("icosphere16")
    orc, ctx, corners, face, bary, vel, rc, kind, params = _pair(V, F, 100, "harmonic", gpu_ctx_factory)
    orc.compute_forces(kind, params)
    ctx.compute_forces(kind, params)
    flagged = np.zeros(100, bool)
    for _ in range(10):
        orc.run_nve(kind, params, 0.01, 100)
        ctx.step_nve(kind, params, 0.01, 100)
        flagged |= (orc.walk_flags() != 0) | (ctx.walk_flags() != 0)
    of, ob, ov, ofr = orc.get_state()
    gf, gb, gv, gfr = ctx.get_state()
    ok = ~flagged
    assert ok.sum() >= 98
    assert np.array_equal(of[ok], gf[ok])
    eo, eg = orc.euclidean(of, ob), orc.euclidean(gf, gb)
    assert np.max(np.abs(eo - eg)[ok]) < TOL_TRAJ and np.max(np.abs(ov - gv)[ok]) < TOL_TRAJ
    c = ctx.counters()
    assert c["overflow"] == 0 and c["walk_nohit"] == 0 and c["walk_nan"] == 0 and c["walk_itercap"] == 0


@pytest.mark.parametrize("name,N,pot,steps", [("torus60x24", 500, "gaussian", 200), ("icosphere40", 3000, "harmonic", 50)])
def test_nve_trajectories(name, N, pot, steps, gpu_ctx_factory):
    V, F = _case_mesh(name)
    orc, ctx, corners, face, bary, vel, rc, kind, params = _pair(V, F, N, pot, gpu_ctx_factory, want_end=False)
    orc.compute_forces(kind, params)
    ctx.compute_forces(kind, params)
    orc.set_options(True, False, 8)
    orc.run_nve(kind, params, 0.01, steps)
    ctx.step_nve(kind, params, 0.01, steps)
    of, ob, ov, ofr = orc.get_state()
    gf, gb, gv, gfr = ctx.get_state()
    assert np.array_equal(of, gf)
    assert np.max(np.abs(orc.euclidean(of, ob) - orc.euclidean(gf, gb))) < TOL_TRAJ
    assert np.max(np.abs(ov - gv)) < TOL_TRAJ
    assert np.max(np.abs(ofr - gfr)) < TOL_FORCE * np.abs(ofr).max() + TOL_TRAJ


def test_config3_like_fire_then_nose_hoover(gpu_ctx_factory):
    """BASELINE.json configs[2] shape: FIRE minimisation followed by Nose-Hoover NVT (M=2, tau=1), harmonic."""
    V, F = meshes.torus(80, 30, R=3.0, r=1.0, jitter=0.25, seed=7)  # non-isotropic, ~4.8k faces
    orc, ctx, corners, face, bary, vel, rc, kind, params = _pair(V, F, 800, "harmonic", gpu_ctx_factory, want_end=False)
    p = np.array([40, 0.01, 0.99, 0.1, 1e-5, 1.1, 0.95, 0.9, 4, 1e-12, 0.0])
    zero = np.zeros_like(vel)
    orc.set_state(face, bary, zero)
    ctx.set_state(face, bary, zero)
    orc.fire_init(p, dt0=0.01, alpha0=0.99)
    ctx.fire_init(p, dt0=0.01, alpha0=0.99)
    _, o_out = orc.run_fire(kind, params)
    g_out = ctx.fire_minimize(kind, params)
    assert o_out[0] == g_out[0] == 40
    assert abs(o_out[1] - g_out[1]) < 1e-9 * o_out[1] and o_out[2] == g_out[2] and o_out[3] == g_out[3]
    of, ob, ov, ofr = orc.get_state()
    gf, gb, gv, gfr = ctx.get_state()
    assert np.array_equal(of, gf) and np.max(np.abs(ob - gb)) < 1e-9 and np.max(np.abs(ov - gv)) < 1e-9
    assert abs(ctx.max_force() - np.sqrt((gfr * gfr).sum(1).max())) < 1e-12
    assert abs(ctx.force_norm() - np.sqrt((gfr * gfr).sum())) < 1e-10
    # phase B: thermalise
    orc.set_state(of, ob, vel)
    ctx.set_state(of, ob, vel)
    orc.compute_forces(kind, params)
    ctx.compute_forces(kind, params)
    orc.nvt_init(0.01, 0.2, tau=1.0, M=2)
    ctx.nvt_init(0.01, 0.2, tau=1.0, M=2)
    orc.run_nvt(kind, params, 60)
    ctx.step_nvt(kind, params, 60)
    of, ob, ov, ofr = orc.get_state()
    gf, gb, gv, gfr = ctx.get_state()
    assert np.array_equal(of, gf) and np.max(np.abs(ob - gb)) < 1e-8 and np.max(np.abs(ov - gv)) < 1e-8
    ob_, oke, osc = orc.nvt_state()
    gb_, gke, gsc = ctx.nvt_state()
    assert np.max(np.abs(ob_ - gb_)) < 1e-9 * max(1.0, np.abs(ob_).max()) and abs(oke - gke) < 1e-9 * oke


def test_gradient_descent(gpu_ctx_factory):
    V, F = _mesh("icosphere16")
    orc, ctx, corners, face, bary, vel, rc, kind, params = _pair(V, F, 200, "harmonic", gpu_ctx_factory)
    orc.run_gd(kind, params, 0.05, 25)
    ctx.step_gd(kind, params, 0.05, 25)
    of, ob, ov, ofr = orc.get_state()
    gf, gb, gv, gfr = ctx.get_state()
    assert np.array_equal(of, gf) and np.max(np.abs(ob - gb)) < 1e-10


# ------------------------------------------------------------------------------ per-call API (baseSpace)
def test_distance_per_call_global_and_submeshed(gpu_ctx_factory):
    V, F = _mesh("icosphere16")
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    ctx = gpu_ctx_factory()
    ctx.set_mesh(V, corners)
    rng = np.random.default_rng(2)
    face, bary = random_positions(len(F), 60, rng)
    for sub, thr in ((False, 1e20), (True, 0.6)):
        orc.set_submeshing(sub, 0.8)
        ctx.set_submeshing(sub, 0.8)
        for s in range(0, 60, 12):
            tf, tb = np.delete(face, s), np.delete(bary, s, 0)
            if sub:
                P = orc.euclidean(face, bary)
                near = np.linalg.norm(np.delete(P, s, 0) - P[s], axis=1) < thr
                tf, tb = tf[near], tb[near]
                if len(tf) == 0:
                    continue
            od, ots, ote, tie, _ = orc.distance(face[s], bary[s], tf, tb, threshold=thr)
            gd, gts, gte = ctx.distance(face[s], bary[s], tf, tb, threshold=thr)
            assert _rel(gd, od) < TOL_DIST
            ok = tie == 0
            assert np.max(np.abs(gts - ots)[ok]) < TOL_TAN and np.max(np.abs(gte - ote)[ok]) < TOL_TAN
    assert ctx.counters()["overflow"] == 0


@pytest.mark.parametrize("name", ["torus60x24", "icosphere16"])
def test_locate_r3_positions_is_bit_exact(name, gpu_ctx_factory):
    """css_locate = simpleModel::R3PositionsToMeshPositions (simpleModel.cpp:136-154): the grid search on the device returns the
    face and the clamped weights of the oracle's brute-force scan bit for bit (exact ties: lowest face index)."""
    V, F = _mesh(name)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    ctx = gpu_ctx_factory()
    ctx.set_mesh(V, corners)
    rng = np.random.default_rng(3)
    face0, bary0 = random_positions(len(F), 2000, rng)
    on = orc.euclidean(face0, bary0)
    ext = float(np.abs(V).max())
    P = np.concatenate([on + 1e-9 * rng.standard_normal((2000, 3)),        # slightly off the surface (the reference's use)
                        on + 0.05 * ext * rng.standard_normal((2000, 3)),  # clearly off
                        4.0 * ext * rng.standard_normal((200, 3)),         # far away, outside the grid
                        V[:200],                                           # on vertices: distance-0 ties between faces
                        0.5 * (V[corners[:200, 1]] + V[corners[:200, 2]]), # on edges
                        np.zeros((1, 3))])                                 # deep inside
    of, ob = orc.locate(P)
    gf, gb = ctx.locate(P)
    assert np.array_equal(of, gf)
    assert np.array_equal(ob, gb)
    assert np.array_equal(gf[:2000], face0)
    gf2, gb2 = ctx.locate(P[:10], clamp_tol=1e-6)                           # the tolerance is the caller's
    of2, ob2 = orc.locate(P[:10], clamp_tol=1e-6)
    assert np.array_equal(of2, gf2) and np.array_equal(ob2, gb2)
    assert ctx.locate(np.zeros((0, 3)))[0].shape == (0,)
    with pytest.raises(binding.CssError):
        ctx.locate(np.array([[np.nan, 0.0, 0.0]]))


def test_disconnected_sentinel_and_unreachable(gpu_ctx_factory):
    V1, F1 = meshes.plane_grid(2, 2)
    V = np.concatenate([V1, V1 + [3.0, 0, 0]])
    F = np.concatenate([F1, F1 + len(V1)]).astype(np.int32)
    corners = meshes.reference_corners(F)
    ctx = gpu_ctx_factory()
    ctx.set_mesh(V, corners)
    b = np.array([0.3, 0.3, 0.4])
    d, ts, te = ctx.distance(0, b, [len(F1) + 1, 1], np.array([b, b]))
    assert d[0] == -1.0 and d[1] > 0                      # global branch: CGAL reports (-1, end) for another component
    ctx.set_submeshing(True, 0.7)
    d, ts, te = ctx.distance(0, b, [len(F1) + 1, 1], np.array([b, b]), threshold=10.0)
    assert d[0] == 1.4 and np.array_equal(ts[0], [0, 0, 1]) and np.array_equal(te[0], [0, 0, 1])  # tMS.cpp:198-203
    assert ctx.counters()["disconnected"] == 2


def test_transport_per_call_with_vectors(gpu_ctx_factory):
    V, F = _mesh("torus60x24")
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    ctx = gpu_ctx_factory()
    ctx.set_mesh(V, corners)
    rng = np.random.default_rng(8)
    n = 400
    face, bary = random_positions(len(F), n, rng)
    vecs = np.stack([random_velocities(V, corners, face, 1.0, rng) for _ in range(3)], 1)
    disp = random_velocities(V, corners, face, 1.0, rng) * rng.uniform(0.0, 2.0, n)[:, None]
    disp[::7] = 0.0                                                                       # zero displacement stays put
    of, ob, od, ov, ofl, _ = orc.transport(face, bary, disp, vecs)
    gf, gb, gd, gv, gfl = ctx.transport(face, bary, disp, vecs)
    assert np.array_equal(ofl, gfl)
    ok = ofl == 0
    assert np.array_equal(of[ok], gf[ok]) and np.array_equal(ob[ok], gb[ok]) and np.array_equal(ov[ok], gv[ok])
    # displaceParticle = transport without vectors
    gf2, gb2, _, _, _ = ctx.transport(face, bary, disp)
    assert np.array_equal(gf2[ok], gf[ok]) and np.array_equal(gb2[ok], gb[ok])


# ------------------------------------------------------------------------------ open meshes
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("sheet", ["plane", "bowl"])
def test_open_mesh_boundary_rules(mode, sheet, gpu_ctx_factory):
    """absorbingOpenMeshSpace / tangentialOpenMeshSpace (openMeshSpace.cpp:114-238): the walker stops on, or slides along,
    border edges and projects transported vectors that point over the boundary.  Bit-exact against the oracle except
    flagged vertex events; border events themselves are part of the compared set."""
    V, F = meshes.plane_grid(12, 12, 2.0, 2.0) if sheet == "plane" else meshes.bowl(16, 16)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    orc.set_boundary(mode)
    ctx = gpu_ctx_factory()
    ctx.set_mesh(V, corners)
    ctx.set_boundary(mode)
    rng = np.random.default_rng(21 + mode)
    n = 2000
    face, bary = random_positions(len(F), n, rng)
    vecs = np.stack([random_velocities(V, corners, face, 1.0, rng) for _ in range(2)], 1)
    disp = random_velocities(V, corners, face, 1.0, rng) * rng.uniform(0.0, 1.5, n)[:, None]
    of, ob, od, ov, ofl, _ = orc.transport(face, bary, disp, vecs)
    gf, gb, gd, gv, gfl = ctx.transport(face, bary, disp, vecs)
    assert np.array_equal(ofl, gfl)
    assert (ofl & 16).sum() > 200                                                          # many particles met the border
    ok = (ofl & ~16) == 0
    assert ok.mean() > 0.95
    assert np.array_equal(of[ok], gf[ok]) and np.array_equal(ob[ok], gb[ok])
    assert np.array_equal(ov[ok], gv[ok]) and np.array_equal(od[ok], gd[ok])
    assert np.all(gb[ok] >= 0)                                                             # nobody left the sheet
    # the closed rule on the same inputs only flags (and differs from the open rules for the flagged particles)
    ctx.set_boundary(0)
    cf, cb, _, cv, cfl = ctx.transport(face, bary, disp, vecs)
    assert np.array_equal(cfl & 16, gfl & 16) or mode == 2                                 # tangential slides may meet more borders
    inside = gfl == 0
    assert np.array_equal(cf[inside], gf[inside]) and np.array_equal(cb[inside], gb[inside])
    # NVE steps with the fused step on the open sheet.  A particle parked in a corner of the sheet is a vertex-degenerate
    # event every step (flagged) and reacts discontinuously to 1e-13 differences, so the GPU is re-seeded with the oracle's
    # state before every step and the one-step map is compared: bit-exact faces, 1e-9 on everything else.
    N = 300
    corners, face, bary, vel = make_state(V, F, N)
    _, _, area = orc.mesh_info()
    rc = interaction_range(area, N)
    kind, params = force_params("harmonic", k=1.0, sigma=rc)
    for sim in (orc, ctx):
        sim.set_boundary(mode)
        sim.set_submeshing(True, rc)
        sim.set_state(face, bary, vel * 2)
        sim.compute_forces(kind, params)
    compared = 0
    for _ in range(40):
        of, ob, ov, ofr = orc.get_state()
        ctx.set_state(of, ob, ov, ofr)
        orc.run_nve(kind, params, 0.01, 1)
        ctx.step_nve(kind, params, 0.01, 1)
        assert np.array_equal(orc.walk_flags(), ctx.walk_flags())
        ok = (orc.walk_flags() & ~16) == 0
        of, ob, ov, ofr = orc.get_state()
        gf, gb, gv, gfr = ctx.get_state()
        assert np.array_equal(of[ok], gf[ok])
        assert np.max(np.abs(ob - gb)[ok]) < 1e-9 and np.max(np.abs(ov - gv)[ok]) < 1e-9
        compared += int(ok.sum())
    assert compared > 39 * N and orc.counters()["border"] > 100 and ctx.counters()["walk_border"] == orc.counters()["border"]

# ------------------------------------------------------------------------------ observables
@pytest.mark.parametrize("pot", ["harmonic", "gaussian"])
def test_stress_tensor_and_temperature(pot, gpu_ctx_factory):
    """simulation::computeMonodisperseStress (simulation.cpp:104-173), noseHooverNVT::getTemperatureFromKE."""
    V, F = _case_mesh("icosphere40")
    orc, ctx, corners, face, bary, vel, rc, kind, params = _pair(V, F, 3000, pot, gpu_ctx_factory, area_fraction=1.5)
    So, Sg = orc.compute_stress(kind, params), ctx.compute_stress(kind, params)
    assert np.abs(So).max() > 0 and np.max(np.abs(So - Sg)) < TOL_FORCE * np.abs(So).max()
    assert abs(orc.temperature() - ctx.temperature()) < 1e-12 * orc.temperature()


@pytest.mark.parametrize("pinned", [False, True])
def test_host_buffer_step_equals_the_device_resident_step(pinned, gpu_ctx_factory):
    """css_step_nve_host (upload, step, download with the position download overlapped on a second stream) against
    css_set_state + css_step_nve + css_get_state: bitwise, through the plain launches and the CUDA-graph replays."""
    V, F = _case_mesh("icosphere40")
    orc, ctx, corners, face, bary, vel, rc, kind, params = _pair(V, F, 3000, "harmonic", gpu_ctx_factory, want_end=False)
    ctx.compute_forces(kind, params)
    f0, b0, v0, fr0 = ctx.get_state()
    ref = gpu_ctx_factory()
    ref.set_mesh(V, corners)
    ref.set_submeshing(True, rc)
    ref.set_options(True, False)
    ref.set_state(f0, b0, v0, fr0)
    bufs = [f0.copy(), b0.copy(), v0.copy(), fr0.copy()]
    if pinned:
        import torch

        bufs = [torch.from_numpy(a).pin_memory().numpy() for a in bufs]
    for step in range(6):
        ctx.step_nve_host(kind, params, 0.01, *bufs)
        ref.step_nve(kind, params, 0.01, 1)
        rf, rb, rv, rfr = ref.get_state()
        assert np.array_equal(bufs[0], rf) and np.array_equal(bufs[1], rb), step
        assert np.array_equal(bufs[2], rv) and np.array_equal(bufs[3], rfr), step
        if step == 2:  # the host owns the state: a change made on the host is what the next step sees
            bufs[2] *= 0.5
            ref.set_velocities(bufs[2])
    assert not np.array_equal(bufs[1], b0)
    bad = bufs[0].copy()
    bad[5] = len(F)
    with pytest.raises(binding.CssError):
        ctx.step_nve_host(kind, params, 0.01, bad, bufs[1], bufs[2], bufs[3])


# ------------------------------------------------------------------------------ golden fixtures
def test_golden_bruteforce_geodesics(gpu_ctx_factory):
    g = np.load(os.path.join(GOLDEN, "bruteforce_geodesics.npz"))
    for name in g["names"]:
        V, F = _mesh(str(name))
        corners = meshes.reference_corners(F)
        ctx = gpu_ctx_factory()
        ctx.set_mesh(V, corners)
        face, bary = g[name + "/face"], g[name + "/bary"]
        for s in range(g[name + "/D"].shape[0]):
            tf, tb = np.delete(face, s), np.delete(bary, s, 0)
            d, ts, te = ctx.distance(face[s], bary[s], tf, tb)
            D, TS, TE = g[name + "/D"][s], g[name + "/TS"][s], g[name + "/TE"][s]
            assert _rel(d, D) < TOL_DIST
            same = (np.abs(ts - TS).max(1) < TOL_TAN) & (np.abs(te - TE).max(1) < TOL_TAN)
            assert same.mean() > 0.9  # tangents may legitimately differ where two shortest paths tie


def test_golden_closed_form(gpu_ctx_factory):
    from test_oracle_geodesic import _locate

    g = np.load(os.path.join(GOLDEN, "closed_form.npz"))
    for name, src, tgt, dist in zip(g["mesh"], g["src"], g["tgt"], g["dist"]):
        V, F = _mesh(str(name))
        corners = meshes.reference_corners(F)
        ctx = gpu_ctx_factory()
        ctx.set_mesh(V, corners)
        sf, sb = _locate(V, corners, src)
        tf, tb = _locate(V, corners, tgt)
        d, _, _ = ctx.distance(sf, sb, [tf], tb[None])
        assert abs(d[0] - dist) < 1e-12


def test_golden_oracle_regression(gpu_ctx_factory):
    g = np.load(os.path.join(GOLDEN, "oracle_regression.npz"))
    for key in g["names"]:
        key = str(key)
        name, N = key.split("_N")[0], int(g[key + "/N"])
        V, F = _mesh(name)
        corners, face, bary, vel = make_state(V, F, N)
        rc, kind, params = float(g[key + "/rc"]), int(g[key + "/kind"]), g[key + "/params"]
        ctx = gpu_ctx_factory()
        ctx.set_mesh(V, corners)
        ctx.set_submeshing(True, rc)
        ctx.set_options(True, True)
        ctx.set_state(face, bary, vel)
        off, idx, d, ts, te = ctx.find_neighbors(rc, want_end=True)
        assert np.array_equal(off, g[key + "/off"]) and np.array_equal(idx, g[key + "/idx"])
        assert _rel(d, g[key + "/dist"]) < TOL_DIST
        assert np.max(np.abs(ts - g[key + "/ts"])) < TOL_TAN and np.max(np.abs(te - g[key + "/te"])) < TOL_TAN
        ctx.compute_forces(kind, params)
        frc = ctx.get_state()[3]
        assert np.max(np.abs(frc - g[key + "/frc"])) < TOL_FORCE * np.abs(g[key + "/frc"]).max()
        ctx.step_nve(kind, params, 0.01, 50)
        f2, b2, v2, fr2 = ctx.get_state()
        assert np.array_equal(f2, g[key + "/face50"])
        assert np.max(np.abs(b2 - g[key + "/bary50"])) < 1e-9 and np.max(np.abs(v2 - g[key + "/vel50"])) < 1e-9


# ------------------------------------------------------------------------------ edge cases / errors
def test_edge_cases(gpu_ctx_factory):
    V, F = _mesh("icosphere16")
    corners = meshes.reference_corners(F)
    ctx = gpu_ctx_factory()
    ctx.set_mesh(V, corners)
    ctx.set_submeshing(True, 0.02)
    # sparse: nobody has a neighbour (K = 0 everywhere); forces are exactly zero and the step still runs
    _, face, bary, vel = make_state(V, F, 20)
    ctx.set_state(face, bary, vel)
    off, idx, d, ts, _ = ctx.find_neighbors(0.02)
    assert off[-1] == 0 and len(idx) == 0
    kind, params = force_params("harmonic", k=1.0, sigma=0.02)
    ctx.compute_forces(kind, params)
    assert np.all(ctx.get_state()[3] == 0)
    ctx.step_nve(kind, params, 0.01, 3)
    # a single particle
    ctx.set_state(face[:1], bary[:1], vel[:1])
    ctx.step_nve(kind, params, 0.01, 3)
    assert ctx.find_neighbors(0.02)[0].tolist() == [0, 0]
    # empty per-call batches
    assert ctx.euclidean(np.zeros(0, np.int32), np.zeros((0, 3))).shape == (0, 3)
    d, ts, te = ctx.distance(0, bary[0], np.zeros(0, np.int32), np.zeros((0, 3)))
    assert len(d) == 0
    # all-to-all candidates without a cell list (baseNeighborStructure), global-mesh geodesics
    orc = Oracle(V, corners)
    orc.set_options(use_cell_list=False)
    orc.set_state(face[:12], bary[:12], vel[:12])
    ctx.set_options(False, True)
    ctx.set_submeshing(False, 0.0)
    ctx.set_state(face[:12], bary[:12], vel[:12])
    o = orc.find_neighbors(1.0)
    g = ctx.find_neighbors(1.0, want_end=True)
    assert np.array_equal(o[0], g[0]) and np.array_equal(o[1], g[1]) and g[0][-1] == 12 * 11
    assert _rel(g[2], o[2]) < TOL_DIST


@pytest.mark.parametrize("area_fraction,kmin", [(12.0, 33), (3.5, 17)])
def test_dense_neighbourhoods_grow_the_stride(area_fraction, kmin, gpu_ctx_factory):
    """config 2b-like stress: K ~ 55 neighbours per particle and large patches (last, whole-mesh tier), and a
    medium density whose K of 17..32 and ~150-face patches run on the large record tier."""
    V, F = _mesh("torus60x24")
    orc, ctx, corners, face, bary, vel, rc, kind, params = _pair(V, F, 600, "harmonic", gpu_ctx_factory, area_fraction=area_fraction)
    o = orc.find_neighbors(rc)
    g = ctx.find_neighbors(rc, want_end=True)
    assert np.array_equal(o[0], g[0]) and np.array_equal(o[1], g[1])
    assert np.diff(o[0]).max() >= kmin                      # beyond the small tier (and, for the dense case, the initial stride)
    assert _rel(g[2], o[2]) < TOL_DIST and np.max(np.abs(g[3] - o[3])) < TOL_TAN
    c = ctx.counters()
    assert c["overflow"] == 0 and c["tier_retry"] > 0


def test_stride_guard_regrows_inside_fused_steps(gpu_ctx_factory):
    """More candidates than the neighbour stride (32) met for the first time INSIDE fused steps: at the first step of a fresh
    context (css_step_nve, css_step_nve_host, NVT, FIRE, GD) and in the middle of CUDA-graph replays (a converging flow piles
    the particles up).  The guard freezes the pipeline, the host regrows the stride, finishes the step and runs the rest:
    the result equals the oracle's, which has no stride at all."""
    V, F = _mesh("torus60x24")
    for upd in ("nve", "nve_host", "nvt", "fire", "gd"):
        orc, ctx, corners, face, bary, vel, rc, kind, params = _pair(V, F, 600, "harmonic", gpu_ctx_factory, area_fraction=12.0, want_end=False)
        f0 = orc.compute_forces(kind, params)
        ctx.set_state(face, bary, vel, f0)                   # no neighbour phase has run on the GPU yet: stride still 32
        if upd == "nve":
            orc.run_nve(kind, params, 0.01, 6)
            ctx.step_nve(kind, params, 0.01, 6)
        elif upd == "nve_host":
            orc.run_nve(kind, params, 0.01, 2)
            hf, hb, hv, hfr = face.copy(), bary.copy(), vel.copy(), f0.copy()
            for _ in range(2):
                ctx.step_nve_host(kind, params, 0.01, hf, hb.reshape(-1), hv.reshape(-1), hfr.reshape(-1))
            of, ob, ov, ofr = orc.get_state()
            assert np.array_equal(of, hf) and np.max(np.abs(ob - hb)) < 1e-9 and np.max(np.abs(ov - hv)) < 1e-9
        elif upd == "nvt":
            orc.nvt_init(0.01, 0.2, tau=1.0, M=2)
            ctx.nvt_init(0.01, 0.2, tau=1.0, M=2)
            orc.run_nvt(kind, params, 4)
            ctx.step_nvt(kind, params, 4)
        elif upd == "fire":
            p = np.array([5, 0.01, 0.99, 0.1, 1e-5, 1.1, 0.95, 0.9, 4, 1e-12, 0.0])
            orc.fire_init(p, dt0=0.01, alpha0=0.99)
            ctx.fire_init(p, dt0=0.01, alpha0=0.99)
            orc.run_fire(kind, params)
            ctx.fire_minimize(kind, params)
        else:
            orc.run_gd(kind, params, 0.01, 4)
            ctx.step_gd(kind, params, 0.01, 4)
        of, ob, ov, ofr = orc.get_state()
        gf, gb, gv, gfr = ctx.get_state()
        ok = (orc.walk_flags() == 0) & (ctx.walk_flags() == 0)
        assert np.array_equal(of[ok], gf[ok]), upd
        assert np.max(np.abs(ob - gb)[ok]) < 1e-9 and np.max(np.abs(ov - gv)[ok]) < 1e-9, upd
        assert np.max(np.abs(ofr - gfr)[ok]) < TOL_FORCE * np.abs(ofr).max(), upd
        c = ctx.counters()
        assert c["overflow"] == 0 and c["kmax_overflow"] == 0   # raised, handled and cleared
    # mid-run: particles streaming towards one point of a sphere; the stride is exceeded after the graph has been captured
    V, F = _mesh("icosphere16")
    orc, ctx, corners, face, bary, vel, rc, kind, params = _pair(V, F, 400, "harmonic", gpu_ctx_factory, want_end=False)
    x = orc.euclidean(face, bary)
    pole = np.array([0.0, 0.0, 1.0])
    tang = pole[None, :] - (x @ pole)[:, None] * x / np.sum(x * x, 1, keepdims=True)
    c3 = corners[face]
    n = np.cross(V[c3[:, 1]] - V[c3[:, 0]], V[c3[:, 2]] - V[c3[:, 0]])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    tang -= np.sum(tang * n, 1, keepdims=True) * n            # in the plane of each particle's face
    vel = 1.2 * tang
    orc.set_state(face, bary, vel)
    ctx.set_state(face, bary, vel)
    orc.compute_forces(kind, params)
    ctx.compute_forces(kind, params)
    assert np.diff(ctx.find_neighbors(rc)[0]).max() < 16
    flagged = np.zeros(400, bool)
    for _ in range(6):
        orc.run_nve(kind, params, 0.01, 30)
        ctx.step_nve(kind, params, 0.01, 30)
        flagged |= (orc.walk_flags() != 0) | (ctx.walk_flags() != 0)
    assert np.diff(orc.find_neighbors(rc)[0]).max() > 40      # the pile-up really went beyond the initial stride
    of, ob, ov, ofr = orc.get_state()
    gf, gb, gv, gfr = ctx.get_state()
    ok = ~flagged
    assert ok.sum() > 380 and np.array_equal(of[ok], gf[ok])
    assert np.max(np.abs(orc.euclidean(of, ob) - orc.euclidean(gf, gb))[ok]) < TOL_TRAJ and np.max(np.abs(ov - gv)[ok]) < TOL_TRAJ
    assert ctx.counters()["overflow"] == 0


def test_error_convention(gpu_ctx_factory):
    ctx = gpu_ctx_factory()
    V, F = _mesh("cube1")
    corners = meshes.reference_corners(F)
    with pytest.raises(binding.CssError) as e:
        ctx.find_neighbors(0.1)
    assert e.value.code == 5                                 # CSS_ESTATE
    bad = corners.copy()
    bad[0] = bad[0][::-1]                                    # inconsistent orientation
    with pytest.raises(binding.CssError) as e:
        ctx.set_mesh(V, bad)
    assert e.value.code == 3                                 # CSS_EMESH
    ctx.set_mesh(V, corners)
    with pytest.raises(binding.CssError) as e:
        ctx.set_state(np.array([len(F)], np.int32), np.array([[0.3, 0.3, 0.4]]))
    assert e.value.code == 1                                 # CSS_EINVAL
    with pytest.raises(binding.CssError):
        ctx.distance(0, [0.3, 0.3, 0.4], [99], [[0.3, 0.3, 0.4]])


def test_long_range_kernel_agrees_with_the_large_record_tier():
    """sphere_radius1-like patches (~150 faces) are served by the 240-face record tier (k_patch<Large> + k_windows<Large>, one warp
    per source); CSS_MAX_LARGE=0 hands them straight on to the block-cooperative fused kernel (k_geodesic_cta).  Two independent
    implementations of the propagation on the same inputs: same lists, distances and tangents to round-off."""
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from curvedspacesim_b200 import binding, meshes
from helpers import make_state, interaction_range
V, F = meshes.icosphere(16)
corners, face, bary, vel = make_state(V, F, 100)
area = float(meshes.face_areas(V, F).sum()); rc = interaction_range(area, 100)
ctx = binding.Context(0); ctx.set_mesh(V, corners); ctx.set_submeshing(True, rc); ctx.set_options(True, True); ctx.set_state(face, bary, vel)
off, idx, d, ts, te = ctx.find_neighbors(rc, want_end=True)
c = ctx.counters()
np.savez(sys.argv[1], off=off, idx=idx, d=d, ts=ts, te=te, retry=c["tier_retry"], overflow=c["overflow"], faces=c["patch_faces"])
""" % (ROOT, os.path.join(ROOT, "tests"))
    import tempfile

    outs = []
    with tempfile.TemporaryDirectory() as td:
        for i, skip in enumerate((None, "0")):
            env = dict(os.environ)
            env.pop("CSS_MAX_LARGE", None)
            if skip is not None:
                env["CSS_MAX_LARGE"] = skip
            out = os.path.join(td, "o%d.npz" % i)
            subprocess.check_call([sys.executable, "-c", code, out], env=env)
            outs.append(dict(np.load(out)))
    a, b = outs
    assert int(a["faces"]) > 100 * 100 and int(a["retry"]) >= 80                  # the large tier is where these sources live ...
    assert int(b["retry"]) >= int(a["retry"]) + 50 and int(b["overflow"]) == 0    # ... and with the switch they move on once more
    assert np.array_equal(a["off"], b["off"]) and np.array_equal(a["idx"], b["idx"]) and len(a["idx"]) > 100
    assert _rel(a["d"], b["d"]) < 1e-12 and np.max(np.abs(a["ts"] - b["ts"])) < 1e-10 and np.max(np.abs(a["te"] - b["te"])) < 1e-10


def test_half_warp_and_one_warp_tier0_agree():
    """Tier 0 runs two sources per warp (window_half_kernel.cu, 8 targets per propagation, groups for K > 8);
    CSS_WIN_HALF=0 selects the one-warp-per-source kernel.  Same neighbour lists, distances and tangents to round-off,
    on a dense state (K up to ~30: several target groups per source) and through fused NVE steps."""
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from curvedspacesim_b200 import binding, meshes
from helpers import make_state, interaction_range
from oracle_binding import force_params
V, F = meshes.torus(120, 40, R=3.0, r=1.0, jitter=0.2, seed=7)
N = 3000
corners, face, bary, vel = make_state(V, F, N)
area = float(meshes.face_areas(V, F).sum()); rc = interaction_range(area, N, 4.0)
kind, params = force_params("harmonic", k=1.0, sigma=rc)
ctx = binding.Context(0); ctx.set_mesh(V, corners); ctx.set_submeshing(True, rc); ctx.set_options(True, True); ctx.set_state(face, bary, vel)
off, idx, d, ts, te = ctx.find_neighbors(rc, want_end=True)
c = ctx.counters()
ctx.step_nve(kind, params, 0.002, 10)
f2, b2, v2, fr2 = ctx.get_state()
np.savez(sys.argv[1], off=off, idx=idx, d=d, ts=ts, te=te, retry=c["tier_retry"], overflow=c["overflow"], face=f2, bary=b2, vel=v2, frc=fr2,
         kmean=len(idx) / N, kmax=int(np.diff(off).max()), spilled=c["spilled"])
""" % (ROOT, os.path.join(ROOT, "tests"))
    import tempfile

    outs = []
    with tempfile.TemporaryDirectory() as td:
        for i, half in enumerate(("1", "0")):
            env = dict(os.environ, CSS_WIN_HALF=half)
            out = os.path.join(td, "o%d.npz" % i)
            subprocess.check_call([sys.executable, "-c", code, out], env=env, timeout=300)
            outs.append(dict(np.load(out)))
    a, b = outs
    assert int(a["kmax"]) > 16 and float(a["kmean"]) > 8          # the grouped path is exercised
    assert int(a["spilled"]) > 0 and int(b["spilled"]) == 0         # and so are the global spill stacks of the half-warp kernel
    assert int(a["overflow"]) == 0 and int(b["overflow"]) == 0
    assert np.array_equal(a["off"], b["off"]) and np.array_equal(a["idx"], b["idx"])
    assert _rel(a["d"], b["d"]) < 1e-12 and np.max(np.abs(a["ts"] - b["ts"])) < 1e-10 and np.max(np.abs(a["te"] - b["te"])) < 1e-10
    assert np.array_equal(a["face"], b["face"])
    assert np.max(np.abs(a["bary"] - b["bary"])) < 1e-9 and np.max(np.abs(a["frc"] - b["frc"])) < 1e-8 * np.abs(a["frc"]).max()


# ------------------------------------------------------------------------------ full size: properties
@pytest.mark.parametrize("workload", ["cfg4_icosphere_250kfaces_N25k", "cfg5_torus_1Mfaces_N100k"])
def test_full_size_properties(workload, gpu_ctx_factory):
    """BASELINE.json configs[3] and [4] at full size.  The oracle would need minutes here, so the checks are
    properties that hold for any correct answer:
      * the neighbour relation is symmetric and ordered as the cell stencil dictates;
      * d >= Euclidean chord, tangents are unit and lie in the source / target face planes;
      * exp map: walking d * startTangent from the source (css_transport) lands on the target, and the transported
        start tangent arrives as the end tangent - ties the geodesic kernel to the independent walker kernel;
      * symmetry d(i,j) = d(j,i), start(i->j) = -end(j->i) (the two are computed on different patches);
      * idempotence: a second call returns the same bits."""
    import bench

    wl = bench.Workload(workload, 0.01)
    V, F, corners, face, bary, vel, N, rc = wl.V, wl.F, wl.corners, wl.face, wl.bary, wl.vel, wl.N, wl.rc
    ctx = gpu_ctx_factory()
    ctx.set_mesh(V, corners)
    ctx.set_submeshing(True, rc)
    ctx.set_options(True, True)
    ctx.set_state(face, bary, vel)
    off, idx, d, ts, te = ctx.find_neighbors(rc, want_end=True)
    off2, idx2, d2, ts2, te2 = ctx.find_neighbors(rc, want_end=True)
    assert np.array_equal(off, off2) and np.array_equal(idx, idx2) and np.array_equal(d, d2) and np.array_equal(ts, ts2)
    c = ctx.counters()
    assert c["overflow"] == 0 and c["disconnected"] == 0
    src = np.repeat(np.arange(N), np.diff(off))
    assert len(idx) > 3 * N
    P = ctx.euclidean(face, bary)
    chord = np.linalg.norm(P[idx] - P[src], axis=1)
    assert np.all(chord < rc) and np.all(d >= chord * (1 - 1e-14)) and np.all(d < 1.2 * rc)
    # symmetric relation: (i,j) present <=> (j,i) present
    key = src.astype(np.int64) * N + idx
    rkey = idx.astype(np.int64) * N + src
    order = np.argsort(key)
    pos = np.searchsorted(key[order], rkey)
    assert np.all(pos < len(key)) and np.array_equal(key[order][pos], rkey)
    rev = order[pos]                                        # index of (j,i) for every (i,j)
    sym = np.abs(d - d[rev]) / d
    assert np.quantile(sym, 0.999) < TOL_DIST
    # The two directions are computed on DIFFERENT patches (each cut at its own source's largest candidate distance,
    # triangulatedMeshSpace.cpp:167-173), so a few pairs are legitimately asymmetric.  Those, plus a random sample of
    # sources, are checked against the oracle's distanceWithSubmeshing one source at a time.
    orc = Oracle(V, corners)
    orc.set_submeshing(True, rc)
    orc.set_state(face, bary)
    _, _, maxd = orc.candidates(rc)
    worst = np.unique(src[np.argsort(sym)[-40:]])
    sample = np.random.default_rng(1).choice(N, size=300, replace=False)
    for i in np.concatenate([worst, sample]):
        a, b = off[i], off[i + 1]
        if a == b:
            continue
        od, ots, ote, tie, _ = orc.distance(face[i], bary[i], face[idx[a:b]], bary[idx[a:b]], threshold=float(maxd[i]))
        assert _rel(d[a:b], od) < TOL_DIST
        ok = tie == 0
        assert np.max(np.abs(ts[a:b] - ots)[ok], initial=0) < TOL_TAN and np.max(np.abs(te[a:b] - ote)[ok], initial=0) < TOL_TAN
    anti = np.abs(ts + te[rev]).max(1)
    assert np.quantile(anti, 0.99) < 1e-8
    nrm = np.cross(V[corners[:, 1]] - V[corners[:, 0]], V[corners[:, 2]] - V[corners[:, 0]])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    assert np.max(np.abs(np.linalg.norm(ts, axis=1) - 1)) < 1e-12 and np.max(np.abs(np.linalg.norm(te, axis=1) - 1)) < 1e-12
    assert np.max(np.abs((ts * nrm[face[src]]).sum(1))) < 1e-10 and np.max(np.abs((te * nrm[face[idx]]).sum(1))) < 1e-10
    # exp map through the walker kernel on a 200k-query sample
    sel = np.random.default_rng(0).choice(len(idx), size=min(200000, len(idx)), replace=False)
    f2, b2, _, v2, flags = ctx.transport(face[src[sel]], bary[src[sel]], ts[sel] * d[sel, None], ts[sel][:, None, :])
    ok = flags == 0
    assert ok.mean() > 0.999
    P2 = ctx.euclidean(f2, b2)
    err = np.linalg.norm(P2 - P[idx[sel]], axis=1)
    # paths that bend at a pseudo-source (saddle or patch-boundary vertex) are not straightest geodesics, so a
    # straight walk misses their target: rare at these densities (< 0.1 % of the queries)
    assert np.quantile(err[ok], 0.999) < 1e-8 and (err[ok] > 1e-8).mean() < 1e-3
    assert np.quantile(np.abs(v2[:, 0] - te[sel]).max(1)[ok], 0.999) < 1e-7


def test_full_size_step_is_deterministic_and_flag_free(gpu_ctx_factory):
    import bench

    wl = bench.Workload("cfg4_icosphere_250kfaces_N25k", 0.01)
    V, F, corners, face, bary, vel, N, rc = wl.V, wl.F, wl.corners, wl.face, wl.bary, wl.vel, wl.N, wl.rc
    kind, params = force_params("harmonic", k=1.0, sigma=rc)
    res = []
    for _ in range(2):
        ctx = gpu_ctx_factory()
        ctx.set_mesh(V, corners)
        ctx.set_submeshing(True, rc)
        ctx.set_state(face, bary, vel)
        ctx.compute_forces(kind, params)
        ctx.step_nve(kind, params, 0.01, 20)
        res.append(ctx.get_state())
        c = ctx.counters()
        assert c["overflow"] == 0 and c["walk_nohit"] == 0 and c["walk_nan"] == 0 and c["walk_itercap"] == 0
        ctx.close()
    for a, b in zip(*res):
        assert np.array_equal(a, b)  # run-to-run bitwise identical (no atomics feed the results)
    v = res[0][2]
    nrm = np.cross(V[corners[:, 1]] - V[corners[:, 0]], V[corners[:, 2]] - V[corners[:, 0]])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    assert np.max(np.abs((v * nrm[res[0][0]]).sum(1))) < 1e-10  # velocities stay tangent after 20 transports


# ------------------------------------------------------------------------------ C++ host layer
def _read_dump(path):
    raw = np.fromfile(path, dtype=np.uint8)
    n = int(raw[:4].view(np.int32)[0])
    o = 4

    def take(count, dt):
        nonlocal o
        a = raw[o:o + count * np.dtype(dt).itemsize].view(dt).copy()
        o += count * np.dtype(dt).itemsize
        return a

    ini = (take(n, np.int32), take(3 * n, np.float64).reshape(n, 3), take(3 * n, np.float64).reshape(n, 3))
    fin = (take(n, np.int32), take(3 * n, np.float64).reshape(n, 3), take(3 * n, np.float64).reshape(n, 3), take(3 * n, np.float64).reshape(n, 3))
    assert o == len(raw)
    return n, ini, fin


@pytest.mark.parametrize("branch,fused", [(2, 0), (2, 1), (3, 1), (3, 0), (1, 1), (0, 1), (0, 0)])
def test_cpp_host_layer_matches_the_oracle(branch, fused, tmp_path):
    """The reference-shaped C++ program (host/example_simulation.cpp on host/css_host.hpp: closedMeshSpace, model, cell
    list, harmonicRepulsion, simulation, NVE / NVT / GD / FIRE updaters) against the oracle started from the SAME initial
    state (the program dumps it), both with the host-driven updaters and with the fused device step."""
    from curvedspacesim_b200 import build

    exe = build.build_host_example()
    V, F = _mesh("icosphere16")
    off = str(tmp_path / "m.off")
    meshes.save_off(off, V, F)
    N, iters, dt, T = 150, (30 if branch else 25), 0.01, 0.2
    dump = str(tmp_path / "d.bin")
    dbdir = str(tmp_path / "traj.cssdb")
    out = subprocess.run([exe, off, str(N), str(iters), str(branch), str(fused), dump, str(dt), str(T)], capture_output=True, text=True,
                         env=dict(os.environ, CSS_EXAMPLE_DB=dbdir))
    assert out.returncode == 0, out.stderr
    n, (f0, b0, v0), (f1, b1, v1, fr1) = _read_dump(dump)
    corners = meshes.reference_corners(F)
    orc = Oracle(V, corners)
    _, _, area = orc.mesh_info()
    rc = interaction_range(area, N)
    kind, params = force_params("harmonic", k=1.0, sigma=rc)
    orc.set_submeshing(True, rc)
    orc.set_state(f0, b0, v0)          # forces start at zero, as in the reference's mains
    if branch == 2:
        orc.run_nve(kind, params, dt, iters)
    elif branch == 3:
        orc.nvt_init(dt, T, tau=1.0, M=2)
        orc.run_nvt(kind, params, iters)
    elif branch == 1:
        orc.run_gd(kind, params, dt, iters)
    else:
        p = np.array([iters, dt, 0.99, 0.1, 1e-5, 1.1, 0.95, 0.9, 4, 1e-12, 0.0])
        orc.fire_init(p, dt0=dt, alpha0=0.99)
        orc.run_fire(kind, params)
    of, ob, ov, ofr = orc.get_state()
    assert np.array_equal(of, f1)
    assert np.max(np.abs(ob - b1)) < 1e-9 and np.max(np.abs(ov - v1)) < 1e-9
    assert np.max(np.abs(ofr - fr1)) < TOL_FORCE * max(np.abs(ofr).max(), 1e-300) + 1e-12
    # the trajectory database written by the program (host/css_database.hpp) holds the same two states
    from curvedspacesim_b200 import trajectory

    db = trajectory.SimpleModelDatabase(N, dbdir, "r")
    assert db.current_number_of_records() == 2
    first, last = db.read_state(0), db.read_state(1)
    assert np.array_equal(first["faceIndex"], f0) and np.array_equal(first["barycentricPosition"], b0) and np.array_equal(first["velocity"], v0)
    assert np.array_equal(last["faceIndex"], f1) and np.array_equal(last["barycentricPosition"], b1) and np.array_equal(last["force"], fr1)
    assert np.array_equal(last["R3position"], orc.euclidean(f1, b1))
    # observables printed by the program: simulation::computeMonodisperseStress, noseHooverNVT::getTemperatureFromKE
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("stress trace")][0].split()
    orc.set_state(f1, b1, v1, fr1)
    So = orc.compute_stress(kind, params)
    assert abs(float(line[2]) - np.trace(So)) < 1e-8 * abs(np.trace(So)) and abs(float(line[4]) - orc.temperature()) < 1e-12


# ------------------------------------------------------------------------------ multi-GPU
@pytest.mark.parametrize("branch", [2, 0])
def test_cpp_host_layer_user_written_force(branch, tmp_path):
    """A user-written `force` subclass (only pairwiseForce / pairwiseEnergy, src/forces/baseForce.h:30-57) on a gpuModel: the
    base class downloads the neighbour lists and calls the virtuals on the host (host/css_host.hpp force::computeForces).  The
    subclass states the harmonic law, so the run must reproduce the stock device functor's: same faces, state to 1e-9."""
    from curvedspacesim_b200 import build

    exe = build.build_host_example()
    V, F = _mesh("icosphere16")
    off = str(tmp_path / "m.off")
    meshes.save_off(off, V, F)
    res = []
    for user in (False, True):
        dump = str(tmp_path / ("d%d.bin" % user))
        env = dict(os.environ)
        if user:
            env["CSS_EXAMPLE_USERFORCE"] = "1"
        out = subprocess.run([exe, off, "150", "25", str(branch), "0", dump, "0.01", "0.2"], capture_output=True, text=True, env=env)
        assert out.returncode == 0, out.stderr
        res.append(_read_dump(dump))
    (n0, ini0, fin0), (n1, ini1, fin1) = res
    assert n0 == n1 and all(np.array_equal(a, b) for a, b in zip(ini0, ini1))
    assert np.array_equal(fin0[0], fin1[0])
    for a, b in zip(fin0[1:], fin1[1:]):
        assert np.max(np.abs(a - b)) < 1e-9


def test_cpp_host_layer_imports_r3_positions(tmp_path):
    """simpleModel::setMeshPositionsFromR3File / R3PositionsToMeshPositions of the C++ host layer (css_locate underneath): the final
    R^3 coordinates of a short run, written as a text file, come back as the mesh positions they were computed from."""
    from curvedspacesim_b200 import build

    exe = build.build_host_example()
    V, F = _mesh("icosphere16")
    off = str(tmp_path / "m.off")
    meshes.save_off(off, V, F)
    out = subprocess.run([exe, off, "120", "5", "2", "1", str(tmp_path / "d.bin")], capture_output=True, text=True,
                         env=dict(os.environ, CSS_EXAMPLE_R3=str(tmp_path / "r3.csv")))
    assert out.returncode == 0, out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith("R3 import:")][0].split()
    n, same, diff = int(line[2]), int(line[4]), float(line[-1])
    assert n == 120 and same >= 118 and diff < 1e-9      # (a point on an edge may legitimately come back on the neighbour)


def test_two_gpus_bitwise_equal_to_one(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = os.path.join(ROOT, "tests", "multigpu_worker.py")
    out = str(tmp_path)
    world = min(torch.cuda.device_count(), 8)
    # the exchange after every move: peer-memory stores fused into the walker (default) and the NCCL all-gather
    for tag, p2p in (("_p2p", "1"), ("_nccl", "0")):
        env = dict(os.environ, CSS_P2P=p2p, CSS_TAG=tag)
        subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr",
                               "127.0.0.1", "--master-port", "29611", script, out], env=env, timeout=600)
    subprocess.check_call([sys.executable, script, out], timeout=600)
    one = np.load(os.path.join(out, "world1_rank0.npz"))
    for tag in ("_p2p", "_nccl"):
        for r in range(world):
            two = np.load(os.path.join(out, "world%d_rank%d%s.npz" % (world, r, tag)))
            assert bool(two["peer"]) == (tag == "_p2p"), "peer-memory exchange was not established"
            assert int(two["timeouts"]) == 0
            assert np.array_equal(two["face"], one["face"]) and np.array_equal(two["bary"], one["bary"])
            lo, hi = int(two["lo"]), int(two["hi"])
            assert np.array_equal(two["vel"], one["vel"][lo:hi]) and np.array_equal(two["frc"], one["frc"][lo:hi])
            # the dense run whose first neighbour phase overflows the stride inside the fused call (stride guard, both flavours)
            assert int(two["ovf3"]) == 0 and int(one["ovf3"]) == 0
            assert np.array_equal(two["face3"], one["face3"]) and np.array_equal(two["bary3"], one["bary3"])
            lo, hi = int(two["lo3"]), int(two["hi3"])
            assert np.array_equal(two["vel3"], one["vel3"][lo:hi]) and np.array_equal(two["frc3"], one["frc3"][lo:hi])
    for r in range(world):  # NVT + NVE continuation: the two transports agree bit for bit
        a = np.load(os.path.join(out, "world%d_rank%d_p2p.npz" % (world, r)))
        b = np.load(os.path.join(out, "world%d_rank%d_nccl.npz" % (world, r)))
        for k in ("face2", "bary2", "vel2", "frc2", "ke"):
            assert np.array_equal(a[k], b[k]), k


# ------------------------------------------------------------------------------ committed fixture of css_locate
def test_golden_locate(gpu_ctx_factory):
    """css_locate against the committed fixture tests/golden/locate_regression.npz (points on / off / far from the surface, on
    vertices and edges): same faces, same clamped weights."""
    g = np.load(os.path.join(GOLDEN, "locate_regression.npz"))
    for name in g["names"]:
        name = str(name)
        V, F = _mesh(name)
        ctx = gpu_ctx_factory()
        ctx.set_mesh(V, meshes.reference_corners(F))
        f, b = ctx.locate(g[name + "/xyz"])
        assert np.array_equal(f, g[name + "/face"])
        assert np.max(np.abs(b - g[name + "/bary"])) < 1e-12
