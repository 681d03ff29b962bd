"""True-reference golden vectors, when somebody has produced them off-box with ref_harness/dump_reference.cpp
(tests/golden/reference_dumps/<case>.out.bin).  None ship with the repository: the reference cannot be built in its
image, so these tests SKIP and say so - parity at the CGAL boundary stays 'unpinned' until
dumps are added."""
from __future__ import annotations

import glob
import json
import os
import sys

import numpy as np
import pytest

from curvedspacesim_b200 import meshes
from helpers import GOLDEN, make_state

DUMPS = sorted(glob.glob(os.path.join(GOLDEN, "reference_dumps", "*.out.bin")))


def _load(path):
    raw = np.fromfile(path, dtype=np.uint8)
    o = 0

    def take(n, dt):
        nonlocal o
        a = raw[o:o + n * np.dtype(dt).itemsize].view(dt).copy()
        o += n * np.dtype(dt).itemsize
        return a

    N = int(take(1, np.int32)[0])
    off = take(N + 1, np.int32)
    tot = int(off[-1])
    d = {"N": N, "off": off, "idx": take(tot, np.int32), "dist": take(tot, np.float64), "ts": take(3 * tot, np.float64).reshape(tot, 3),
         "te": take(3 * tot, np.float64).reshape(tot, 3), "frc": take(3 * N, np.float64).reshape(N, 3), "face": take(N, np.int32),
         "bary": take(3 * N, np.float64).reshape(N, 3), "vel": take(3 * N, np.float64).reshape(N, 3), "frc_end": take(3 * N, np.float64).reshape(N, 3)}
    if o < len(raw):  # trailing R3PositionsToMeshPositions block (dumps made by the current ref_harness)
        M = int(take(1, np.int32)[0])
        d["loc_xyz"] = take(3 * M, np.float64).reshape(M, 3)
        d["loc_face"] = take(M, np.int32)
        d["loc_bary"] = take(3 * M, np.float64).reshape(M, 3)
    assert o == len(raw)
    return d


def _compare_locate(locate, ref):
    """PMP::locate_with_AABB_tree + the reference's clamp against css_locate / the oracle: same face except where the point is
    equidistant from two faces (the documented tie rule), weights to 1e-9."""
    if "loc_xyz" not in ref:
        pytest.skip("this dump has no R3PositionsToMeshPositions block")
    f, b = locate(ref["loc_xyz"])
    same = f == ref["loc_face"]
    assert same.mean() > 0.99
    assert np.max(np.abs(b[same] - ref["loc_bary"][same])) < 1e-9


def _case(path):
    sys.path.insert(0, GOLDEN)
    from make_golden import golden_mesh

    name = os.path.basename(path)[:-len(".out.bin")]
    meta = json.load(open(os.path.join(GOLDEN, "reference_inputs", name + ".json")))
    V, F = golden_mesh(meta["mesh"])
    corners, face, bary, vel = make_state(V, F, meta["N"])
    return meta, V, corners, face, bary, vel, _load(path)


def _compare(impl, meta, face, bary, vel, ref, set_state, find, forces, nve, get_state):
    set_state(face, bary, vel)
    off, idx, d, ts, te = find(meta["range"])
    assert np.array_equal(off, ref["off"]) and np.array_equal(idx, ref["idx"])              # bit-exact neighbour lists
    assert np.max(np.abs(d - ref["dist"]) / ref["dist"]) < 1e-9
    assert np.max(np.abs(ts - ref["ts"])) < 1e-9 and np.max(np.abs(te - ref["te"])) < 1e-9
    f = forces()
    assert np.max(np.abs(f - ref["frc"])) < 1e-8 * np.abs(ref["frc"]).max()
    nve(meta["dt"], meta["steps"])
    f2, b2, v2, _ = get_state()
    assert np.array_equal(f2, ref["face"])                                                   # bit-exact faces
    assert np.max(np.abs(b2 - ref["bary"])) < 1e-6 and np.max(np.abs(v2 - ref["vel"])) < 1e-6


@pytest.mark.skipif(not DUMPS, reason="no true-reference dumps under tests/golden/reference_dumps (see ref_harness/README.md)")
@pytest.mark.parametrize("path", DUMPS)
def test_oracle_against_reference_dump(path):
    from oracle_binding import Oracle, force_params

    meta, V, corners, face, bary, vel, ref = _case(path)
    orc = Oracle(V, corners)
    orc.set_submeshing(True, meta["range"])
    kind, params = force_params("harmonic", k=1.0, sigma=meta["range"])
    _compare("oracle", meta, face, bary, vel, ref, orc.set_state, orc.find_neighbors, lambda: orc.compute_forces(kind, params),
             lambda dt, n: orc.run_nve(kind, params, dt, n), orc.get_state)
    _compare_locate(orc.locate, ref)


@pytest.mark.gpu
@pytest.mark.skipif(not DUMPS, reason="no true-reference dumps under tests/golden/reference_dumps (see ref_harness/README.md)")
@pytest.mark.parametrize("path", DUMPS)
def test_cuda_against_reference_dump(path, gpu_ctx_factory):
    from curvedspacesim_b200 import binding

    meta, V, corners, face, bary, vel, ref = _case(path)
    ctx = gpu_ctx_factory()
    ctx.set_mesh(V, corners)
    ctx.set_submeshing(True, meta["range"])
    ctx.set_options(True, True)
    kind, params = binding.force_params("harmonic", k=1.0, sigma=meta["range"])

    def forces():
        ctx.compute_forces(kind, params)
        return ctx.get_state()[3]

    _compare("cuda", meta, face, bary, vel, ref, ctx.set_state, lambda r: ctx.find_neighbors(r, want_end=True), forces,
             lambda dt, n: ctx.step_nve(kind, params, dt, n), ctx.get_state)
    _compare_locate(ctx.locate, ref)
