"""Shared test helpers (the initial-state sampling laws live in curvedspacesim_b200/initial_state.py)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
for _p in (os.path.dirname(os.path.abspath(__file__)), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from curvedspacesim_b200 import meshes  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


from curvedspacesim_b200.initial_state import interaction_range, make_state, random_positions, random_velocities  # noqa: E402,F401


def csr_rows(off, arr):
    return [arr[off[i]:off[i + 1]] for i in range(len(off) - 1)]
