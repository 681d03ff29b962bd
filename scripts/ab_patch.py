"""developer tool: A/B timing + equality of two library builds on a bench workload (run on the GPU box).
usage: python scripts/ab_patch.py <workload> <libA> <libB> ..."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, json, numpy as np
sys.path.insert(0, %r)
import bench
from curvedspacesim_b200 import binding
wl = bench.Workload(sys.argv[1], 0.01)
kind, params = wl.force(binding.force_params)
ctx = binding.Context(0)
ctx.set_mesh(wl.V, wl.corners); ctx.set_submeshing(True, wl.rc); ctx.set_options(True, False)
ctx.set_state(wl.face, wl.bary, wl.vel)
ctx.compute_forces(kind, params)
ctx.counters(reset=True)
off, idx, d, ts, _ = ctx.find_neighbors(wl.rc)
c = ctx.counters()
ctx.step_nve(kind, params, 0.01, 10)
ctx.set_timing(True)
ctx.step_nve(kind, params, 0.01, 3)
P, W, R, Wk, Ce, T = [], [], [], [], [], []
for _ in range(20):
    ctx.timer_record(0); ctx.step_nve(kind, params, 0.01, 1); ctx.timer_record(1)
    T.append(ctx.timer_elapsed_ms(0, 1))
    k = ctx.last_stage_ms(); P.append(k["patch_ms"]); W.append(k["window_ms"]); R.append(k["retry_ms"])
    k = ctx.last_kernel_ms(); Wk.append(k["walk_ms"]); Ce.append(k["celllist_ms"])
f, b, v, fr = ctx.get_state()
import hashlib
h = hashlib.sha256(f.tobytes() + b.tobytes() + v.tobytes() + fr.tobytes()).hexdigest()[:16]
hn = hashlib.sha256(off.tobytes() + idx.tobytes() + d.tobytes() + ts.tobytes()).hexdigest()[:16]
print(json.dumps({"lib": os.environ.get("CSS_LIB_PATH", "default").split("/")[-1] + " stencil=" + os.environ.get("CSS_STENCIL", "1"), "step_ms": float(np.median(T)), "patch_ms": float(np.median(P)), "window_ms": float(np.median(W)),
                  "retry_ms": float(np.median(R)), "walk_ms": float(np.median(Wk)), "cell_ms": float(np.median(Ce)), "state_hash": h, "nbr_hash": hn,
                  "patch_faces": c["patch_faces"], "patch_verts": c["patch_verts"], "queries": c["queries"], "retry": c["tier_retry"], "overflow": c["overflow"],
                  "windows": c["windows"], "pseudo": c.get("pseudo_sources", -1), "clk": [c["clk_patch"], c["clk_batch"], c["clk_fan"], c["clk_prop"], c["clk_total"]]}))
''' % ROOT
wl = sys.argv[1]
for spec in sys.argv[2:]:  # "<lib or default>[:ENV=VALUE[:ENV=VALUE...]]"
    lib, *sets = spec.split(":")
    env = dict(os.environ)
    for kv in sets:
        env[kv.split("=")[0]] = kv.split("=")[1]
    if lib != "default":
        env["CSS_LIB_PATH"] = os.path.join(ROOT, lib)
    env["CSS_VERBOSE"] = "1"
    r = subprocess.run([sys.executable, "-c", code, wl], env=env, capture_output=True, text=True)
    for ln in r.stderr.splitlines():
        if ln.startswith("[css]"):
            print(ln, flush=True)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + spec + ": " + r.stderr[-800:], flush=True)
