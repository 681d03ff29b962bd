"""Dev diagnostic (GPU box): open-mesh NVE, GPU vs oracle one step at a time; prints the first divergence."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from curvedspacesim_b200 import binding, meshes
from helpers import interaction_range, make_state
from oracle_binding import Oracle, force_params
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
V, F = meshes.bowl(16, 16)
N = 300
corners, face, bary, vel = make_state(V, F, N)
orc = Oracle(V, corners); ctx = binding.Context(0); ctx.set_mesh(V, corners)
_, _, area = orc.mesh_info()
rc = interaction_range(area, N)
kind, params = force_params("harmonic", k=1.0, sigma=rc)
for sim in (orc, ctx):
    sim.set_boundary(mode); sim.set_submeshing(True, rc); sim.set_state(face, bary, vel * 2); sim.compute_forces(kind, params)
np.set_printoptions(precision=17, linewidth=200)
for s in range(40):
    of0, ob0, ov0, ofr0 = orc.get_state(); gf0, gb0, gv0, gfr0 = ctx.get_state()
    orc.run_nve(kind, params, 0.01, 1); ctx.step_nve(kind, params, 0.01, 1)
    of, ob, ov, ofr = orc.get_state(); gf, gb, gv, gfr = ctx.get_state()
    db = np.abs(ob - gb).max(1); dv = np.abs(ov - gv).max(1); dfr = np.abs(ofr - gfr).max(1)
    fl_o, fl_g = orc.walk_flags(), ctx.walk_flags()
    print("step", s, "max db %.3e dv %.3e dfr %.3e" % (db.max(), dv.max(), dfr.max()), "flags o", np.unique(fl_o, return_counts=True), "g", np.unique(fl_g, return_counts=True), "face mismatch", (of != gf).sum())
    bad = np.where((db > 1e-9) | (dv > 1e-9))[0]
    if len(bad):
        i = bad[0]
        print("first bad particle", i, "flags", fl_o[i], fl_g[i])
        print(" before: face", of0[i], gf0[i], "bary", ob0[i], gb0[i], "\n vel", ov0[i], gv0[i], "\n frc", ofr0[i], gfr0[i])
        print(" after: face", of[i], gf[i], "bary", ob[i], gb[i], "\n vel", ov[i], gv[i], "\n frc", ofr[i], gfr[i])
        d = 0.01 * gv0[i] + 0.5e-4 * gfr0[i]
        vh = gv0[i] + 0.005 * gfr0[i]
        print(" replay gpu-state on oracle:", orc.transport([gf0[i]], [gb0[i]], [d], np.array([[vh]])))
        print(" replay gpu-state on gpu:   ", ctx.transport([gf0[i]], [gb0[i]], [d], np.array([[vh]])))
        break
