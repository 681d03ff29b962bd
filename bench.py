"""bench.py — particle-timesteps/s (and geodesic queries/s) of the geodesic MD step on B200.

Default workload (BASELINE.json configs[4], SURVEY.md §8(d) config 5, the configuration the metric is quoted on; it
fits one GPU so it is also the N=1 workload): synthetic torus R=3 r=1, 1250x400 grid -> 1 M faces, N = 100 000
particles, harmonic repulsion k=1 sigma=r_c=2 sqrt(0.9 A/(N pi)), submeshing at r_c, cell list, velocity-Verlet NVE
dt=0.01, T=0.2, seed 13377.  A "step" is one performTimestep: walker + position exchange + cell list + per-source
patch, exact geodesics, pair forces.  Particles are block-sharded over the ranks exactly like
mpiModel::determineIndexBounds with the mesh replicated (strong scaling: N is fixed).

`--workload` also offers BASELINE.json configs 1-4: the reference's own example meshes (sphere_radius1.off N=100 NVE,
torusrb20.off N=2000 gaussian NVE, triangulatedElephant.off N=5000 Nose-Hoover) and the 250 k-face icosphere N=25 000.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm: the oracle port on all host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from curvedspacesim_b200 import meshes, sharding  # noqa: E402
from curvedspacesim_b200.initial_state import interaction_range, make_state  # noqa: E402

METRIC = "particle_timesteps_per_s"
UNIT = "particle-timesteps/s"
MESHDIR = os.path.join(ROOT, "tests", "golden", "meshes")  # copies of the reference's exampleMeshes/*.off (data only)


def _off(name):
    return lambda: meshes.load_off(os.path.join(MESHDIR, name))


WORKLOADS = {
    # name: mesh builder, N, potential, integrator
    # the reference's default executable (curvedSpaceSimulation.cpp:25-29): range 2.6, ~830-face patches -> whole-mesh tier
    "default_exe_torus_isotropic_N20": dict(mesh=_off("torus_isotropic_remesh.off"), N=20, potential="harmonic", integrator="nve"),
    "cfg1_sphere_radius1_N100": dict(mesh=_off("sphere_radius1.off"), N=100, potential="harmonic", integrator="nve"),
    "cfg2_torusrb20_N2000_gaussian": dict(mesh=_off("torusrb20.off"), N=2000, potential="gaussian", integrator="nve"),
    "cfg3_elephant_N5000_nvt": dict(mesh=_off("triangulatedElephant.off"), N=5000, potential="harmonic", integrator="nvt"),
    "cfg4_icosphere_250kfaces_N25k": dict(mesh=lambda: meshes.icosphere(112), N=25000, potential="harmonic", integrator="nve"),
    "cfg5_torus_1Mfaces_N100k": dict(mesh=lambda: meshes.torus(1250, 400, R=3.0, r=1.0, jitter=0.2, seed=13377), N=100000,
                                     potential="harmonic", integrator="nve"),
    "small_torus_24kfaces_N5k": dict(mesh=lambda: meshes.torus(200, 60, R=3.0, r=1.0, jitter=0.2, seed=13377), N=5000,
                                     potential="harmonic", integrator="nve"),
}
POTENTIAL_TEXT = {"harmonic": "harmonic k=1 sigma=r_c", "gaussian": "gaussian alpha=1 sigma=r_c/2 range=r_c"}
INTEGRATOR_TEXT = {"nve": "velocity-Verlet NVE", "nvt": "Nose-Hoover NVT M=2 tau=1 T=0.2 (2 moves + 1 force evaluation per step)"}
TEMPERATURE = 0.2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


class Workload:
    def __init__(self, name, dt):
        w = WORKLOADS[name]
        self.name, self.N, self.potential, self.integrator, self.dt = name, w["N"], w["potential"], w["integrator"], dt
        self.V, self.F = w["mesh"]()
        self.corners, self.face, self.bary, self.vel = make_state(self.V, self.F, self.N, seed=13377, T=TEMPERATURE)
        self.area = float(meshes.face_areas(self.V, self.F).sum())
        self.rc = interaction_range(self.area, self.N, 0.9)

    def force(self, force_params):
        if self.potential == "harmonic":
            return force_params("harmonic", k=1.0, sigma=self.rc)
        return force_params("gaussian", alpha=1.0, sigma=0.5 * self.rc, range=self.rc)

    def config(self):
        """Workload description shared verbatim by both arms (the driver compares the two `config` objects)."""
        return {"workload": self.name, "N": self.N, "faces": int(len(self.F)), "vertices": int(len(self.V)), "r_c": self.rc, "dt": self.dt,
                "potential": POTENTIAL_TEXT[self.potential], "integrator": INTEGRATOR_TEXT[self.integrator],
                "sharding": "particle blocks (mpiModel::determineIndexBounds), mesh replicated",
                "l2": "L2 flushed (256 MiB write, untimed) between timed GPU steps"}


def shard(N, rank, nranks):
    """mpiModel::determineIndexBounds (src/models/mpiModel.cpp:20-31)."""
    return sharding.index_bounds(N, rank, nranks)


def oracle_arm(wl, threads, steps, warmup, budget_s=None):
    """Times `steps` steps of the CPU oracle port (oracle/liboracle.so; the CGAL reference cannot be built in this image) with
    `threads` host threads sharded like mpiModel::determineIndexBounds.  With `budget_s` the step count is chosen from one
    calibration step so that the sample takes about that long (2..500 steps).
    Returns (particle-timesteps/s, seconds, counters, steps)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))  # the CPU arm is the one place bench.py may use the oracle
    from oracle_binding import Oracle, force_params

    orc = Oracle(wl.V, wl.corners)
    orc.set_submeshing(True, wl.rc)
    orc.set_options(True, False, threads)
    kind, params = wl.force(force_params)
    orc.set_state(wl.face, wl.bary, wl.vel)
    orc.compute_forces(kind, params)
    if wl.integrator == "nvt":
        orc.nvt_init(wl.dt, TEMPERATURE, tau=1.0, M=2)
        run = lambda n: orc.run_nvt(kind, params, n)  # noqa: E731
    else:
        run = lambda n: orc.run_nve(kind, params, wl.dt, n)  # noqa: E731
    for _ in range(warmup):
        run(1)
    if budget_s is not None:
        t1 = run(1)
        steps = int(min(500, max(2, budget_s / max(t1, 1e-9))))
    orc.counters(reset=True)
    t = run(steps)
    return wl.N * steps / t, t, orc.counters(), steps


def run_reference(args, rank, world):
    """CPU arm: the oracle port on all host cores (`cpu_baseline.kind: "port"`)."""
    if rank != 0:
        return
    wl = Workload(args.workload, args.dt)
    cores = os.cpu_count() or 1
    value, t, c, _ = oracle_arm(wl, cores, args.steps, args.warmup)
    sample = "full step of %s (N=%d), %d steps, %d threads sharded like mpiModel::determineIndexBounds; oracle built -O3 -march=x86-64-v3 " \
             "-ffp-contract=off" % (args.workload, wl.N, args.steps, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": wl.config(),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "windows_per_step": c["windows_processed"] / args.steps,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5_torus_1Mfaces_N100k", choices=sorted(WORKLOADS))
    ap.add_argument("--dt", type=float, default=0.01)
    ap.add_argument("--cpu-steps", type=int, default=2, help="minimum steps of a CPU baseline sample (rank 0, N=1 only)")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="target duration of each CPU baseline sample (all cores, one thread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    # rank 0 must print exactly ONE line on stdout: native libraries (NCCL's version banner, ...) write to file descriptor 1
    # directly, so everything but the final JSON line is sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    from curvedspacesim_b200 import binding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    wl = Workload(args.workload, args.dt)
    N, rc = wl.N, wl.rc
    lo, hi = shard(N, rank, world)
    nloc = hi - lo
    kind, params = wl.force(binding.force_params)
    nvt = wl.integrator == "nvt"

    ctx = binding.Context(local_rank)
    ctx.set_mesh(wl.V, wl.corners)
    ctx.set_submeshing(True, rc)
    ctx.set_options(True, False)
    if world > 1:
        uid = [binding.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    else:
        ctx.comm_init(0, 1, None)
    ctx.set_state(wl.face, wl.bary, wl.vel[lo:hi], None, n_local=nloc, min_idx=lo)
    ctx.compute_forces(kind, params)
    if nvt:
        ctx.nvt_init(args.dt, TEMPERATURE, tau=1.0, M=2)

    def step(n=1):
        if nvt:
            ctx.step_nvt(kind, params, n)
        else:
            ctx.step_nve(kind, params, args.dt, n)

    flush_buf = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def flush():
        if flush_buf is not None:
            flush_buf.zero_()
            torch.cuda.synchronize()

    def align():
        """Multi-rank: the ranks leave the host barrier tens of microseconds apart, and a step's first exchange would wait for the
        latest one inside the timed region.  An (untimed, idempotent) position all-gather queued right before the start event makes
        every rank's stream wait for the others on the device, so the timed steps start together."""
        if world > 1:
            dist.barrier()
            ctx.gather_positions()

    def timed_steps(k):
        """k steps, each timed with CUDA events on the launching stream, L2 flushed (untimed) in between."""
        ms = []
        for _ in range(k):
            flush()
            align()
            ctx.timer_record(0)
            step(1)
            ctx.timer_record(1)
            ms.append(ctx.timer_elapsed_ms(0, 1))
        return ms

    for _ in range(max(args.warmup, 3)):
        step(1)
    ctx.synchronize()
    ctx.counters(reset=True)

    # ---- timed region: exactly K steps.  The clock sampler runs across the whole region.
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    step_ms = timed_steps(args.steps)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    cnt = ctx.counters()  # counters of the timed region (queries, launches, patch statistics, flags)

    # ---- the same K steps again with the per-phase event records switched on (they cost a few microseconds per step, so
    # the headline above runs without them): kernel durations for the roofline and the phase table
    ctx.set_timing(True)
    for _ in range(3):
        step(1)
    ctx.synchronize()
    inst_ms, geo_ms, walk_ms, cell_ms, patch_ms, win_ms, retry_ms, gather_ms = [], [], [], [], [], [], [], []
    for _ in range(args.steps):
        flush()
        align()
        ctx.timer_record(0)
        step(1)
        ctx.timer_record(1)
        inst_ms.append(ctx.timer_elapsed_ms(0, 1))
        k = ctx.last_kernel_ms()
        geo_ms.append(k["geodesic_ms"])
        walk_ms.append(k["walk_ms"])
        cell_ms.append(k["celllist_ms"])
        k = ctx.last_stage_ms()
        patch_ms.append(k["patch_ms"])
        win_ms.append(k["window_ms"])
        retry_ms.append(k["retry_ms"])
        gather_ms.append(k["gather_ms"])
    barrier()
    ctx.set_timing(False)
    for _ in range(3):
        step(1)
    ctx.synchronize()
    inst_total_ms = max_over_ranks(float(np.sum(inst_ms)))
    total_ms = max_over_ranks(float(np.sum(step_ms)))
    value = N * args.steps / (total_ms * 1e-3)
    queries = sum_over_ranks(float(cnt["queries"]))
    geo_total_ms = max(max_over_ranks(float(np.sum(geo_ms))), 1e-9)

    # ---- hot-L2 back-to-back bracket (same K steps, no flush): reported beside the headline
    barrier()
    ctx.timer_record(2)
    step(args.steps)
    ctx.timer_record(3)
    hot_ms = max_over_ranks(ctx.timer_elapsed_ms(2, 3))
    barrier()

    # ---- the same K steps with end tangents materialised (a query = distance + 2 tangents; the MD step itself
    # consumes only the start tangent, simpleModel.cpp:105-107, so the headline runs without them)
    ctx.set_options(True, True)
    for _ in range(3):
        step(1)
    ctx.synchronize()
    ctx.set_timing(True)
    for _ in range(3):
        step(1)
    c0 = ctx.counters()
    te_step, te_geo = [], []
    for _ in range(args.steps):
        flush()
        align()
        ctx.timer_record(0)
        step(1)
        ctx.timer_record(1)
        te_step.append(ctx.timer_elapsed_ms(0, 1))
        te_geo.append(ctx.last_kernel_ms()["geodesic_ms"])
    c1 = ctx.counters()
    barrier()
    ctx.set_timing(False)
    ctx.set_options(True, False)
    for _ in range(3):
        step(1)
    ctx.synchronize()
    te_queries = sum_over_ranks(float(c1["queries"] - c0["queries"]))
    te_total = max_over_ranks(float(np.sum(te_step)))
    te_geo_total = max(max_over_ranks(float(np.sum(te_geo))), 1e-9)

    # ---- end to end through the C ABI with HOST buffers: every step uploads the step's inputs (positions, velocities,
    # forces) from pinned host memory, runs one step and reads the step's result back into the same pinned buffers; timed by
    # the host clock around the whole loop, barrier + synchronize on both sides.  NVE: css_step_nve_host (the position
    # download overlaps the force phase).  NVT: css_set_state + css_step_nvt + css_get_state.
    gf, gb, gv, gfr = ctx.get_state()
    hf = torch.from_numpy(gf.copy()).pin_memory().numpy()
    hb = torch.from_numpy(gb.copy()).pin_memory().numpy()
    hv = torch.from_numpy(gv.copy()).pin_memory().numpy()
    hfr = torch.from_numpy(gfr.copy()).pin_memory().numpy()
    e2e_steps = max(3, min(args.steps, 10))

    def host_step():
        if nvt:
            ctx.set_state(hf, hb, hv, hfr, n_local=nloc, min_idx=lo)
            ctx.step_nvt(kind, params, 1)
            ctx.get_state_into(hf, hb.reshape(-1), hv.reshape(-1), hfr.reshape(-1))
        else:
            ctx.step_nve_host(kind, params, args.dt, hf, hb, hv, hfr)

    for _ in range(4):  # untimed warm-up of the host path (two plain steps size everything, then the graph is captured)
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step()
    barrier()
    e2e_t = max_over_ranks(time.perf_counter() - t0)
    e2e_value = N * e2e_steps / e2e_t
    h2d = N * (4 + 24) + nloc * 48
    d2h = N * (4 + 24) + nloc * 48

    # ---- roofline.  SURVEY.md §8(d): algorithmic bytes per particle-timestep
    #   B_step = B_state + B_cell + B_patch + B_cand + B_out = 152 + 60 + 24 P_f + 24 P_v + 28 K + 36 K
    # (+152 for the second walker pass of an NVT step), with P_f, P_v, K counted by the kernels in this run.  `achieved`
    # charges the whole step's bytes to the dominant kernel's duration (CUDA events on the launching stream); the same bytes
    # over the whole step are in `roofline_step`.  `builder_bytes_per_source` is the kernel's own itemised traffic (record
    # sections, edge frames, vertices, targets, outputs) kept as a second, named figure.
    ns = max(cnt["sources"], 1)
    pf, pv, kq = cnt["patch_faces"] / ns, cnt["patch_verts"] / ns, cnt["queries"] / ns
    b_step = 152 + 60 + 24 * pf + 24 * pv + 28 * kq + 36 * kq + (152 if nvt else 0)
    builder_bytes = (16 + 5 * kq + 5 * pv + 12 * pf) + 48 * pf + 24 * pv + 48 * kq + 52 + 36 * kq + 24 + 48
    stage = {"k_patch": float(np.mean(patch_ms)), "k_windows_half": float(np.mean(win_ms)), "retry_tiers": float(np.mean(retry_ms))}
    dom = max(stage, key=stage.get)  # retry_tiers = k_patch<Large> + k_windows<Large> + the long-range CTA tiers: dominant on config 1 and the default exe
    dom_ms = stage[dom]
    achieved = (b_step * nloc) / (dom_ms * 1e-3) / 1e9
    pk, pk_kind = peaks()
    kmet = None
    mpath = os.path.join(ROOT, "profiles", "kernel_metrics.json")
    if os.path.exists(mpath):  # figures of the committed `ncu --set full` captures (profiles/README.md), keyed by workload
        with open(mpath) as fh:
            kmet = json.load(fh).get(args.workload)
    km = (kmet or {}).get(dom, {}) if world == 1 else {}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                "traffic": km.get("dram_bytes_per_launch"),
                "peak_source": pk_kind + " (MEASURED_PEAKS.json hbm_gbs, burst copy)" if pk_kind == "measured" else "fallback 6650 GB/s",
                "kernel": dom + {"k_windows_half": " (stage 2, two sources per warp: window propagation + queries + pair forces + half kick)",
                                 "k_patch": " (stage 1: ordered candidates + patch = the static stencil of the source's face restricted to the cut-off; flood fill where a face has no stencil)",
                                 "retry_tiers": " (patches above 96 faces: k_patch<Large> + k_windows<Large>, one warp per source, up to 240 faces; beyond that "
                                                "k_geodesic_cta, one CTA per source)"}[dom],
                "formula": "SURVEY 8(d): B_step = 152 + 60 + 24 P_f + 24 P_v + 64 K" + (" + 152 (second NVT move)" if nvt else ""),
                "algorithmic_bytes_per_source": b_step, "builder_bytes_per_source": builder_bytes,
                "frac_builder_bytes": (builder_bytes * nloc) / (dom_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                "kernel_ms_per_launch": dom_ms,
                "kernel_share_of_step": max_over_ranks(float(np.sum({"k_windows_half": win_ms, "k_patch": patch_ms, "retry_tiers": retry_ms}[dom]))) / inst_total_ms,
                "ms_per_step_with_phase_events": inst_total_ms / args.steps,
                "other_kernels_ms_per_step": {"k_patch": stage["k_patch"], "k_windows_half": stage["k_windows_half"],
                                              "retry_tiers": float(np.mean(retry_ms)), "k_walk": float(np.mean(walk_ms)),
                                              "celllist": float(np.mean(cell_ms))},
                "note": "mesh SoA + edge frames stay L2-resident; the kernels are issue/latency-bound fp64 work, not HBM-bound: "
                        "see roofline_issue and fp64 below"}
    roofline_step = {"achieved": (b_step * N) / (total_ms / args.steps * 1e-3) / 1e9, "unit": "GB/s", "peak": pk["hbm_gbs"],
                     "frac": (b_step * N) / (total_ms / args.steps * 1e-3) / 1e9 / pk["hbm_gbs"]}
    # issue-slot / FP64-pipe / lane figures of the committed ncu capture of the dominant kernel (not re-measured here: ncu
    # cannot run inside a timed bench), and the measured double-precision FMA peak of THIS device in THIS run
    fp64_peak = ctx.microbench(0, 5)
    l2_peak = ctx.microbench(1, 5)
    roofline_issue = None
    fp64 = {"peak_tflops_measured": fp64_peak, "l2_read_gbs_measured": l2_peak,
            "how": "css_microbench: 8 independent DFMA chains per thread, 148 x 8 blocks x 256 threads, best of 5 (CUDA events)"}
    if km:
        roofline_issue = {k: km.get(k) for k in ("issue_slot_pct", "fp64_pipe_pct", "lanes_per_instruction", "warps_active_pct",
                                                  "warp_instructions_per_source", "capture")}
        if km.get("fp64_flops_per_window") and dom == "k_windows_half":
            fl = km["fp64_flops_per_window"] * cnt["windows"] / max(args.steps, 1)  # this rank's launches
            fp64["achieved_tflops"] = fl / (dom_ms * 1e-3) / 1e12
            fp64["frac"] = fp64["achieved_tflops"] / fp64_peak if fp64_peak > 0 else None
            fp64["flops_per_window"] = km["fp64_flops_per_window"]
            fp64["flops_source"] = "ncu smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on (dfma x 2) of the committed capture / windows of that launch"

    # ---- CPU baselines (oracle port) on rank 0 at N=1 only: all host cores (= multirankSimulation over all cores) and one
    # thread (= the reference's single rank), each on a bounded sample of the same workload
    cpu = cpu1 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, t, _, nsteps = oracle_arm(wl, cores, args.cpu_steps, 1, budget_s=args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "seconds": t,
               "sample": "%d full steps of the same workload (N=%d) with %d threads sharded like mpiModel; oracle built -O3 -march=x86-64-v3 "
                         "-ffp-contract=off" % (nsteps, N, cores)}
        v1, t1, _, n1 = oracle_arm(wl, 1, args.cpu_steps, 0, budget_s=args.cpu_seconds)
        cpu1 = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port", "seconds": t1,
                "sample": "%d full steps of the same workload (N=%d) on one thread (the reference's single rank)" % (n1, N)}

    if rank == 0:
        cfg = wl.config()
        if args.no_flush:
            cfg["l2"] = "hot L2 (--no-flush)"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic" if "cfg4" in args.workload or "cfg5" in args.workload or "small" in args.workload
            else "reference example mesh, synthetic particle state",
            "config": cfg,
            "impl_config": {"exchange": "none (1 rank)" if world == 1 else ("peer-memory stores fused into the walker + flag barrier (NVLink, CUDA IPC)"
                                                                             if ctx.comm_info()[2] else "NCCL all-gather"),
                            "cuda_graph": not nvt, "end_tangents": False},
            "geodesic_queries_per_s": queries / (geo_total_ms * 1e-3),
            "queries_per_step": queries / args.steps,
            "with_end_tangents": {"value": N * args.steps / (te_total * 1e-3), "ms_per_step": te_total / args.steps,
                                  "geodesic_queries_per_s": te_queries / (te_geo_total * 1e-3),
                                  "note": "query = distance + start + end tangent (SURVEY 8(d)); same K steps, css_set_options(1, 1)"},
            "value_hot_l2": N * args.steps / (hot_ms * 1e-3), "ms_per_step_hot_l2": hot_ms / args.steps,
            "wall_s_timed_region": t_wall,
            "phase_ms_per_step": {"walk": float(np.mean(walk_ms)), "celllist": float(np.mean(cell_ms)), "geodesic_force": float(np.mean(geo_ms)),
                                  "patch_records": float(np.mean(patch_ms)), "window_propagation": float(np.mean(win_ms)),
                                  "position_allgather": float(np.mean(gather_ms))},
            "patch_mean": {"faces": pf, "verts": pv, "K": kq, "windows_per_source": cnt["windows"] / ns, "tier_retry_frac": cnt["tier_retry"] / ns},
            "flags": {k: cnt[k] for k in ("walk_vertex", "walk_nohit", "walk_itercap", "walk_nan", "walk_border", "disconnected", "overflow")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": cnt["kernels"],
            "roofline": roofline,
            "roofline_step": roofline_step,
            "roofline_issue": roofline_issue,
            "fp64": fp64,
            "cpu_baseline": cpu,
            "cpu_baseline_single_rank": cpu1,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
