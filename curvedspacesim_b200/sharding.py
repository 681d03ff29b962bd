"""Host-side sharding logic of the multi-GPU path (one process per GPU), mirroring the reference's MPI model:

  index_bounds          mpiModel::determineIndexBounds            src/models/mpiModel.cpp:20-31
  pack_block            mpiModel::processSendingBuffer            src/models/mpiModel.cpp:124-133
  unpack_blocks         mpiModel::processReceivingBuffer + readFromGlobalPositions   :135-157
  fold_in_rank_order    mpiSimulation::manipulateUpdaterData      src/simulation/mpiSimulation.cpp:69-89

The device path (css_gather_positions / css_reduce in csrc/css_api.cu) implements the same layout with
NCCL; these functions are the host statement of that layout, used by bench.py to shard the workload and
by the world_size-2 gloo tests.  Pure index arithmetic: there is no geometry and no fallback here."""
from __future__ import annotations

import math

import numpy as np


def per_rank(n_total: int, nranks: int) -> int:
    """largestNumberOfParticlesPerRank = ceil(NTotal / totalRanks)."""
    return int(math.ceil(n_total / nranks))


def index_bounds(n_total: int, rank: int, nranks: int):
    """[lo, hi) owned by `rank`.  The last rank takes the remainder; the reference's latent failure when
    (R-1)*per >= N (an empty or negative last block) is guarded by clamping."""
    if not (0 <= rank < nranks):
        raise ValueError("rank %d outside [0, %d)" % (rank, nranks))
    per = per_rank(n_total, nranks)
    lo = rank * per
    hi = (rank + 1) * per
    if rank == nranks - 1:
        hi = n_total
    return min(lo, n_total), min(hi, n_total)


def pack_block(face, bary, lo: int, hi: int, per: int):
    """One rank's send buffers, padded to `per` entries: int32 face [per] and float64 bary [3 per]."""
    fi = np.zeros(per, np.int32)
    fd = np.zeros(3 * per, np.float64)
    n = hi - lo
    fi[:n] = face[lo:hi]
    fd[:3 * n] = np.asarray(bary, np.float64)[lo:hi].reshape(-1)
    return fi, fd


def unpack_blocks(recv_i, recv_d, n_total: int, nranks: int):
    """Gathered buffers ([nranks * per] int32, [nranks * 3 per] float64) -> replicated (face, bary)."""
    per = per_rank(n_total, nranks)
    face = np.zeros(n_total, np.int32)
    bary = np.zeros((n_total, 3), np.float64)
    ri = np.asarray(recv_i).reshape(nranks, per)
    rd = np.asarray(recv_d).reshape(nranks, per, 3)
    for r in range(nranks):
        lo, hi = index_bounds(n_total, r, nranks)
        face[lo:hi] = ri[r, :hi - lo]
        bary[lo:hi] = rd[r, :hi - lo]
    return face, bary


def fold_in_rank_order(partials, op: str = "sum"):
    """partials [nranks][k] -> [k]; the fold starts from 0 for both sum and max, as the reference's does."""
    acc = np.zeros(np.asarray(partials).shape[1], np.float64)
    for row in np.asarray(partials, np.float64):
        acc = np.maximum(acc, row) if op == "max" else acc + row
    return acc
