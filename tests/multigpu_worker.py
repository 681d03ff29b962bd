"""Worker of tests/test_gpu_parity.py::test_two_gpus_bitwise_equal_to_one.  Launched either directly
(world 1) or under torch.distributed.run (world 2..8): every rank owns a block of particles
(mpiModel::determineIndexBounds), the mesh is replicated, positions are exchanged after every move: over peer
memory (walker stores + flag barrier, the default) or the NCCL all-gather (CSS_P2P=0).  Dumps the final state per rank."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from curvedspacesim_b200 import binding, meshes, sharding  # noqa: E402
from helpers import interaction_range, make_state  # noqa: E402


def main(out_dir):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    V, F = meshes.torus(200, 60, R=3.0, r=1.0, jitter=0.2, seed=13377)
    N = 5001
    corners, face, bary, vel = make_state(V, F, N)
    area = float(meshes.face_areas(V, F).sum())
    rc = interaction_range(area, N)
    kind, params = binding.force_params("harmonic", k=1.0, sigma=rc)
    lo, hi = sharding.index_bounds(N, rank, world)
    ctx = binding.Context(local)
    ctx.set_mesh(V, corners)
    ctx.set_submeshing(True, rc)
    if world > 1:
        uid = [binding.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    ctx.set_state(face, bary, vel[lo:hi], None, n_local=hi - lo, min_idx=lo)
    ctx.compute_forces(kind, params)
    ctx.step_nve(kind, params, 0.01, 25)      # 2 plain steps, then CUDA-graph replays
    ke = ctx.reduce(binding.SUM, [0.5 * float((ctx.get_state()[2] ** 2).sum())])
    f, b, v, fr = ctx.get_state()
    # a few Nose-Hoover steps (two moves + one force evaluation each; the kinetic-energy sum folds per-rank partials, so this
    # part is compared between the two exchange transports at the same world size, not against world 1)
    ctx.nvt_init(0.01, 0.2, 1.0, 2)
    ctx.step_nvt(kind, params, 3)
    ctx.step_nve(kind, params, 0.01, 5)
    f2, b2, v2, fr2 = ctx.get_state()
    # dense neighbourhoods (K ~ 55 > the initial stride of 32) met for the first time inside a fused call: on one rank the exact
    # candidate count raises the stride guard, on several ranks the replicated coarse-block bound does (every rank in the same
    # step, without talking); either way the library regrows, finishes the step and goes on
    Vd, Fd = meshes.torus(60, 24, R=3.0, r=1.0, jitter=0.2, seed=13377)
    Nd = 601
    cd, fd, bd, vd = make_state(Vd, Fd, Nd)
    rcd = interaction_range(float(meshes.face_areas(Vd, Fd).sum()), Nd, 12.0)
    kd, pd = binding.force_params("harmonic", k=1.0, sigma=rcd)
    lod, hid = sharding.index_bounds(Nd, rank, world)
    dense = binding.Context(local)
    dense.set_mesh(Vd, cd)
    dense.set_submeshing(True, rcd)
    if world > 1:
        uid = [binding.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        dense.comm_init(rank, world, uid[0])
    dense.set_state(fd, bd, vd[lod:hid], None, n_local=hid - lod, min_idx=lod)
    dense.step_nve(kd, pd, 0.01, 4)
    f3, b3, v3, fr3 = dense.get_state()
    dcnt = dense.counters()
    dense.close()
    tag = os.environ.get("CSS_TAG", "")
    np.savez(os.path.join(out_dir, "world%d_rank%d%s.npz" % (world, rank, tag)), face=f, bary=b, vel=v, frc=fr, lo=lo, hi=hi, ke=ke,
             face2=f2, bary2=b2, vel2=v2, frc2=fr2, peer=ctx.comm_info()[2], timeouts=ctx.counters()["peer_timeout"],
             face3=f3, bary3=b3, vel3=v3, frc3=fr3, lo3=lod, hi3=hid, ovf3=dcnt["overflow"] + dcnt["kmax_overflow"])
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
