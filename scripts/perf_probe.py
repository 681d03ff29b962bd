"""Developer probe: time the NVE step of one workload and print phase times + counters (one JSON line)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import bench  # noqa: E402
from curvedspacesim_b200 import binding  # noqa: E402

_cache = {}


def run(workload, steps=int(os.environ.get("PROBE_STEPS", "5")), label=""):
    if workload not in _cache:
        _cache[workload] = bench.Workload(workload, 0.01)
    wl = _cache[workload]
    V, F, corners, face, bary, vel, N, rc = wl.V, wl.F, wl.corners, wl.face, wl.bary, wl.vel, wl.N, wl.rc
    kind, params = binding.force_params("harmonic", k=1.0, sigma=rc)
    ctx = binding.Context(0)
    ctx.set_mesh(V, corners)
    ctx.set_submeshing(True, rc)
    ctx.set_state(face, bary, vel)
    ctx.compute_forces(kind, params)
    ctx.set_timing(True)
    for _ in range(3):
        ctx.step_nve(kind, params, 0.01, 1)
    ctx.counters(reset=True)
    ms, geo, stg = [], [], []
    for _ in range(steps):
        ctx.timer_record(0)
        ctx.step_nve(kind, params, 0.01, 1)
        ctx.timer_record(1)
        ms.append(ctx.timer_elapsed_ms(0, 1))
        geo.append(ctx.last_kernel_ms()["geodesic_ms"])
        stg.append(list(ctx.last_stage_ms().values())[:3])
    c = ctx.counters()
    ns = max(c["sources"], 1)
    out = {"label": label, "tune": os.environ.get("CSS_TUNE", ""), "workload": workload, "ms_per_step": float(np.mean(ms)),
           "geo_ms": float(np.mean(geo)), "patch_win_retry_ms": [round(float(x), 4) for x in np.mean(np.array(stg), 0)], "retry_max_ms": round(float(np.max(np.array(stg)[:, 2])), 4), "Mpts_per_s": N / np.mean(ms) / 1e3, "win_per_src": c["windows"] / ns, "ps_per_src": c["pseudo_sources"] / ns,
           "retry_frac": c["tier_retry"] / ns, "ovf": [c["ovf_candidates"], c["ovf_faces"], c["ovf_verts"], c["ovf_ring"]],
           "overflow": c["overflow"],
           "dbg_per_src": {k: round(c[k] / ns, 3) for k in ("clk_batch", "clk_fan", "clk_prop", "clk_patch", "clk_total")}}
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg5_torus_1Mfaces_N100k"
    run(wl, label=" ".join(sys.argv[2:]))
