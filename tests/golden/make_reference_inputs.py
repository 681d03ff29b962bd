"""Writes the inputs of ref_harness/dump_reference.cpp (run it where curvedSpaceSim + CGAL are built):
tests/golden/reference_inputs/<case>.off, <case>.state.bin (N, face, bary, vel) and <case>.json (range, steps, dt).

OFF faces are written so that the reference's corner order (c, a, b) for an OFF line "3 a b c" reproduces the corner
order used everywhere in this repository (meshes.reference_corners), i.e. the plain generator faces are written."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

from curvedspacesim_b200 import meshes  # noqa: E402
from helpers import interaction_range, make_state  # noqa: E402
from make_golden import golden_mesh  # noqa: E402

CASES = [("icosphere16", 200, 50, 0.01), ("torus60x24", 500, 50, 0.01)]


def main():
    out = os.path.join(HERE, "reference_inputs")
    os.makedirs(out, exist_ok=True)
    for name, N, steps, dt in CASES:
        V, F = golden_mesh(name)
        corners, face, bary, vel = make_state(V, F, N)
        area = float(meshes.face_areas(V, F).sum())
        rc = interaction_range(area, N)
        meshes.save_off(os.path.join(out, name + ".off"), V, F)
        with open(os.path.join(out, name + ".state.bin"), "wb") as fh:
            fh.write(np.int32(N).tobytes())
            fh.write(face.astype(np.int32).tobytes())
            fh.write(np.ascontiguousarray(bary, np.float64).tobytes())
            fh.write(np.ascontiguousarray(vel, np.float64).tobytes())
        with open(os.path.join(out, name + ".json"), "w") as fh:
            json.dump({"mesh": name, "N": N, "range": rc, "steps": steps, "dt": dt,
                       "command": "./dump_reference.out %s.off %s.state.bin %s.out.bin %.17g %d %g" % (name, name, name, rc, steps, dt)}, fh, indent=1)
        print("wrote", name)


if __name__ == "__main__":
    main()
