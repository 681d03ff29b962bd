"""bench.py — particle-timesteps/s (and geodesic queries/s) of the geodesic MD step on B200.

Workload (BASELINE.json configs[4], SURVEY.md §8(d) config 5, the configuration the metric is quoted
on; it fits one GPU so it is also the N=1 workload): synthetic torus R=3 r=1, 1250x400 grid -> 1 M
faces, N = 100 000 particles, harmonic repulsion k=1 sigma=r_c=2 sqrt(0.9 A/(N pi)), submeshing at r_c,
cell list, velocity-Verlet NVE dt=0.01, T=0.2, seed 13377.  A "step" is one performTimestep
(velocityVerletNVE::performUpdate): walker + position exchange + cell list + per-source patch,
exact geodesics, pair forces.  Particles are block-sharded over the ranks exactly like
mpiModel::determineIndexBounds with the mesh replicated (strong scaling: N is fixed).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm: the oracle port on all host cores
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from curvedspacesim_b200 import meshes, sharding  # noqa: E402
from helpers import interaction_range, make_state  # noqa: E402

METRIC = "particle_timesteps_per_s"
UNIT = "particle-timesteps/s"

WORKLOADS = {
    # name: (mesh builder, N)
    "cfg5_torus_1Mfaces_N100k": (lambda: meshes.torus(1250, 400, R=3.0, r=1.0, jitter=0.2, seed=13377), 100000),
    "cfg4_icosphere_250kfaces_N25k": (lambda: meshes.icosphere(112), 25000),
    "small_torus_24kfaces_N5k": (lambda: meshes.torus(200, 60, R=3.0, r=1.0, jitter=0.2, seed=13377), 5000),
}


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


def build_workload(name):
    builder, N = WORKLOADS[name]
    V, F = builder()
    corners, face, bary, vel = make_state(V, F, N, seed=13377, T=0.2)
    area = float(meshes.face_areas(V, F).sum())
    rc = interaction_range(area, N, 0.9)
    return V, F, corners, face, bary, vel, N, rc


def shard(N, rank, nranks):
    """mpiModel::determineIndexBounds (src/models/mpiModel.cpp:20-31)."""
    return sharding.index_bounds(N, rank, nranks)


def run_reference(args, rank, world):
    """CPU arm: the oracle port (the CGAL reference cannot be built in this image) on all host cores."""
    if rank != 0:
        return
    from oracle_binding import Oracle, force_params

    V, F, corners, face, bary, vel, N, rc = build_workload(args.workload)
    cores = os.cpu_count() or 1
    orc = Oracle(V, corners)
    orc.set_submeshing(True, rc)
    orc.set_options(True, False, cores)
    kind, params = force_params("harmonic", k=1.0, sigma=rc)
    orc.set_state(face, bary, vel)
    orc.compute_forces(kind, params)
    for _ in range(args.warmup):
        orc.run_nve(kind, params, args.dt, 1)
    orc.counters(reset=True)
    t = orc.run_nve(kind, params, args.dt, args.steps)
    c = orc.counters()
    value = N * args.steps / t
    sample = "full step of %s (N=%d), %d steps, %d threads sharded like mpiModel::determineIndexBounds" % (args.workload, N, args.steps, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "N": N, "faces": int(len(F)), "r_c": rc, "dt": args.dt, "potential": "harmonic k=1",
                       "integrator": "velocity-Verlet NVE"},
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
    ap.add_argument("--cpu-steps", type=int, default=2, help="steps of the CPU baseline sample (rank 0, N=1 only)")
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
        os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout; rank 0 must print exactly one JSON line

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

    V, F, corners, face, bary, vel, N, rc = build_workload(args.workload)
    lo, hi = shard(N, rank, world)
    nloc = hi - lo
    kind, params = binding.force_params("harmonic", k=1.0, sigma=rc)

    ctx = binding.Context(local_rank)
    ctx.set_mesh(V, corners)
    ctx.set_submeshing(True, rc)
    ctx.set_options(True, False)
    if world > 1:
        uid = [binding.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    else:
        ctx.comm_init(0, 1, None)
    ctx.set_state(face, bary, vel[lo:hi], None, n_local=nloc, min_idx=lo)
    ctx.compute_forces(kind, params)
    flush_buf = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def flush():
        if flush_buf is not None:
            flush_buf.zero_()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        ctx.step_nve(kind, params, args.dt, 1)
    ctx.synchronize()
    c0 = ctx.counters(reset=True)

    # ---- timed region: exactly K steps; each step timed with CUDA events on the launching stream,
    # L2 flushed (untimed) between steps.  The clock sampler runs across the whole region.
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    step_ms = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush()
        if world > 1:
            dist.barrier()
        ctx.timer_record(0)
        ctx.step_nve(kind, params, args.dt, 1)
        ctx.timer_record(1)
        step_ms.append(ctx.timer_elapsed_ms(0, 1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    cnt = ctx.counters()  # counters of the timed region (queries, launches, patch statistics, flags)

    # ---- the same K steps again with the per-phase event records switched on (they cost a few microseconds per step, so
    # the headline above runs without them): kernel durations for the roofline and the phase table
    ctx.set_timing(True)
    for _ in range(3):
        ctx.step_nve(kind, params, args.dt, 1)
    ctx.synchronize()
    inst_ms, geo_ms, walk_ms, cell_ms, patch_ms, win_ms, retry_ms, gather_ms = [], [], [], [], [], [], [], []
    for _ in range(args.steps):
        flush()
        if world > 1:
            dist.barrier()
        ctx.timer_record(0)
        ctx.step_nve(kind, params, args.dt, 1)
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
        ctx.step_nve(kind, params, args.dt, 1)
    ctx.synchronize()
    inst_total_ms = max_over_ranks(float(np.sum(inst_ms)))
    total_ms = max_over_ranks(float(np.sum(step_ms)))
    value = N * args.steps / (total_ms * 1e-3)
    queries = sum_over_ranks(float(cnt["queries"]))
    geo_total_ms = max_over_ranks(float(np.sum(geo_ms)))

    # ---- hot-L2 back-to-back bracket (same K steps, no flush): reported beside the headline
    barrier()
    ctx.timer_record(2)
    ctx.step_nve(kind, params, args.dt, args.steps)
    ctx.timer_record(3)
    hot_ms = max_over_ranks(ctx.timer_elapsed_ms(2, 3))
    barrier()

    # ---- end to end through the C ABI with HOST buffers (css_step_nve_host): every step uploads the step's inputs
    # (positions, velocities, forces) from pinned host memory, runs one velocity-Verlet step and reads the step's result
    # back into the same pinned buffers (the position download overlaps the force phase); timed by the host clock around
    # the whole loop, barrier + synchronize on both sides
    gf, gb, gv, gfr = ctx.get_state()
    hf = torch.from_numpy(gf.copy()).pin_memory().numpy()
    hb = torch.from_numpy(gb.copy()).pin_memory().numpy()
    hv = torch.from_numpy(gv.copy()).pin_memory().numpy()
    hfr = torch.from_numpy(gfr.copy()).pin_memory().numpy()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(4):  # untimed warm-up of the host path (two plain steps size everything, then the graph is captured)
        ctx.step_nve_host(kind, params, args.dt, hf, hb, hv, hfr)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.step_nve_host(kind, params, args.dt, hf, hb, hv, hfr)
    barrier()
    e2e_t = max_over_ranks(time.perf_counter() - t0)
    e2e_value = N * e2e_steps / e2e_t
    h2d = N * (4 + 24) + nloc * 48
    d2h = N * (4 + 24) + nloc * 48

    # ---- roofline of the dominant kernel (k_windows_half, stage 2): algorithmic bytes per launch / measured duration.
    # Per source it must read its patch record (header 16, tIdx 4K, tFace K, velig P_v, gface 4 P_f, gvert 4 P_v,
    # fvert 4 P_f, fadj 4 P_f), the edge frames of the patch faces (48 P_f), the patch vertices (24 P_v), the targets'
    # barycentric + Euclidean positions (48 K), its own (52), and write idx/dist/start tangent (36 K), the force (24)
    # and the kicked velocity (read + write 48).  DESIGN.md "Measurement" states the same formula.
    ns = max(cnt["sources"], 1)
    pf, pv, kq = cnt["patch_faces"] / ns, cnt["patch_verts"] / ns, cnt["queries"] / ns
    bytes_per_source = (16 + 5 * kq + 5 * pv + 12 * pf) + 48 * pf + 24 * pv + 48 * kq + 52 + 36 * kq + 24 + 48
    win_ms_per_launch = float(np.mean(win_ms))
    achieved = (bytes_per_source * nloc) / (win_ms_per_launch * 1e-3) / 1e9
    pk, pk_kind = peaks()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_k_windows_traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this workload
        with open(tpath) as fh:
            tj = json.load(fh)
        if tj.get("workload") == args.workload and world == 1:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                "traffic": traffic,
                "peak_source": pk_kind + " (MEASURED_PEAKS.json hbm_gbs, burst copy)" if pk_kind == "measured" else "fallback 6650 GB/s",
                "kernel": "k_windows_half (stage 2, two sources per warp: window propagation + queries + pair forces + half kick)",
                "algorithmic_bytes_per_source": bytes_per_source, "kernel_ms_per_launch": win_ms_per_launch,
                "kernel_share_of_step": max_over_ranks(float(np.sum(win_ms))) / inst_total_ms,
                "ms_per_step_with_phase_events": inst_total_ms / args.steps,
                "other_kernels_ms_per_step": {"k_patch": float(np.mean(patch_ms)), "retry_tiers": float(np.mean(retry_ms)),
                                              "k_walk": float(np.mean(walk_ms)), "celllist": float(np.mean(cell_ms))},
                "note": "mesh SoA + edge frames + patch records stay L2-resident; the kernel is issue/latency-bound fp64 work, "
                        "not HBM-bound (see profiles/ for FP64-pipe and issue-slot utilisation)"}

    # ---- CPU baseline (oracle port, all host cores) on rank 0 at N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle_binding import Oracle, force_params as ofp

        cores = os.cpu_count() or 1
        orc = Oracle(V, corners)
        orc.set_submeshing(True, rc)
        orc.set_options(True, False, cores)
        okind, oparams = ofp("harmonic", k=1.0, sigma=rc)
        orc.set_state(face, bary, vel)
        orc.compute_forces(okind, oparams)
        tcpu = orc.run_nve(okind, oparams, args.dt, args.cpu_steps)
        cpu = {"value": N * args.cpu_steps / tcpu, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d full NVE steps of the same workload (N=%d) with %d threads sharded like mpiModel" % (args.cpu_steps, N, cores)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "N": N, "faces": int(len(F)), "vertices": int(len(V)), "r_c": rc, "dt": args.dt,
                       "potential": "harmonic k=1", "integrator": "velocity-Verlet NVE", "sharding": "particle blocks (mpiModel), mesh replicated",
                       "exchange": "none (1 rank)" if world == 1 else ("peer-memory stores fused into the walker + flag barrier (NVLink, CUDA IPC)"
                                                                        if ctx.comm_info()[2] else "NCCL all-gather"),
                       "l2": "hot-L2 run reported separately" if args.no_flush else "L2 flushed (256 MiB write, untimed) between timed steps"},
            "geodesic_queries_per_s": queries / (geo_total_ms * 1e-3),
            "queries_per_step": queries / args.steps,
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
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
