"""Development check run on the GPU box: CUDA path (through the C ABI) vs the CPU oracle on small cases."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from curvedspacesim_b200 import binding, meshes  # noqa: E402
from helpers import interaction_range, make_state  # noqa: E402
from oracle_binding import Oracle, force_params  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if len(a) else 0.0


def check(name, V, F, N, steps=20, dt=0.01, area_fraction=0.9):
    corners, face, bary, vel = make_state(V, F, N)
    orc = Oracle(V, corners)
    _, _, area = orc.mesh_info()
    rc = interaction_range(area, N, area_fraction)
    kind, params = force_params("harmonic", k=1.0, sigma=rc)
    orc.set_submeshing(True, rc)
    orc.set_state(face, bary, vel)
    ctx = binding.Context(0)
    ctx.set_mesh(V, corners)
    ctx.set_submeshing(True, rc)
    ctx.set_options(True, True)
    ctx.set_state(face, bary, vel)
    # euclid
    e0 = orc.euclidean(face, bary)
    e1 = ctx.euclidean(face, bary)
    print(name, "euclid bit-equal:", np.array_equal(e0, e1))
    # neighbours
    t0 = time.time()
    o_off, o_idx, o_d, o_ts, o_te = orc.find_neighbors(rc)
    t1 = time.time()
    g_off, g_idx, g_d, g_ts, g_te = ctx.find_neighbors(rc, want_end=True)
    t2 = time.time()
    print(name, "neighbour lists bit-equal:", np.array_equal(o_off, g_off) and np.array_equal(o_idx, g_idx), "queries", len(o_idx),
          "oracle s", round(t1 - t0, 3), "gpu s", round(t2 - t1, 3))
    if np.array_equal(o_off, g_off):
        print(name, "dist rel err", rel(g_d, o_d), "ts err", float(np.abs(g_ts - o_ts).max()), "te err", float(np.abs(g_te - o_te).max()))
    print(name, "counters", ctx.counters(), "oracle", orc.counters())
    # forces
    f0 = orc.compute_forces(kind, params)
    ctx.compute_forces(kind, params)
    f1 = ctx.get_state()[3]
    print(name, "force max abs err", float(np.abs(f0 - f1).max()), "scale", float(np.abs(f0).max()))
    # one move from identical state
    rng = np.random.default_rng(5)
    disp = rng.standard_normal((N, 3)) * 0.3 * rc
    orc.set_state(face, bary, vel)
    ctx.set_state(face, bary, vel)
    orc.move(disp.copy())
    ctx.move(disp.copy())
    of, ob, ov, _ = orc.get_state()
    gf, gb, gv, _ = ctx.get_state()
    print(name, "move: face bit-equal", np.array_equal(of, gf), "bary bit-equal", np.array_equal(ob, gb), "vel bit-equal",
          np.array_equal(ov, gv), "max bary diff", float(np.abs(ob - gb).max()), "flags", int(ctx.walk_flags().sum()))
    # NVE trajectory
    orc.set_state(face, bary, vel)
    ctx.set_state(face, bary, vel)
    orc.compute_forces(kind, params)
    ctx.compute_forces(kind, params)
    t0 = time.time()
    orc.run_nve(kind, params, dt, steps)
    t1 = time.time()
    ctx.step_nve(kind, params, dt, steps)
    ctx.synchronize()
    t2 = time.time()
    of, ob, ov, ofr = orc.get_state()
    gf, gb, gv, gfr = ctx.get_state()
    eo = orc.euclidean(of, ob)
    eg = orc.euclidean(gf, gb)
    print(name, "NVE %d steps: faces equal %d/%d, max pos err %.3e, max vel err %.3e; oracle %.3fs gpu %.3fs" %
          (steps, int((of == gf).sum()), N, float(np.abs(eo - eg).max()), float(np.abs(ov - gv).max()), t1 - t0, t2 - t1))
    print(name, "final counters", ctx.counters())
    ctx.close()


if __name__ == "__main__":
    V, F = meshes.icosphere(16)
    check("icosphere16/N=200", V, F, 200)
    V, F = meshes.torus(60, 24, jitter=0.2)
    check("torus60x24/N=500", V, F, 500)
    V, F = meshes.icosphere(40)
    check("icosphere40/N=3000", V, F, 3000, steps=10)
