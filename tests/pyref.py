"""Literal pure-Python transcriptions (small cases only) of the two topology-fixing steps of the path,
used to check the C++ oracle's ordering/bit contract independently of it (TEST INFRASTRUCTURE):

  candidate_lists : hyperRectangularCellList::setGridSize/positionToCellIndex/sort/getCellNeighbors
                    (src/utility/hyperRectangularCellList.cpp:9-46, 71-80, 82-125, 128-159), Index3D
                    (src/utility/indexer.h:47-50) and cellListNeighborStructure::
                    constructCandidateNeighborList (src/utility/cellListNeighborStructure.cpp:45-84)
  patch_faces     : submesher::constructSubmeshFromSourceAndTargets (src/utility/submesher.cpp:55-147)
"""
from __future__ import annotations

import math

import numpy as np


def candidate_lists(P, mn, mx, rng):
    """P [N,3] Euclidean positions -> (list of ordered candidate lists, list of max distance)."""
    P = np.asarray(P, np.float64)
    n = [max(1, int(math.floor((mx[d] - mn[d]) / rng))) for d in range(3)]
    cs = [(mx[d] - mn[d]) / n[d] for d in range(3)]

    def coord(x, d):
        return max(0, min(n[d] - 1, int(math.floor((x - mn[d]) / cs[d]))))

    cells = {}
    for i in range(len(P)):  # ascending particle index inside each cell
        c = (coord(P[i, 0], 0), coord(P[i, 1], 1), coord(P[i, 2], 2))
        cells.setdefault(c, []).append(i)
    out, maxd = [], []
    r2 = rng * rng
    for i in range(len(P)):
        ix, iy, iz = coord(P[i, 0], 0), coord(P[i, 1], 1), coord(P[i, 2], 2)
        lst, m2 = [], 0.0
        for xx in range(max(0, ix - 1), min(n[0] - 1, ix + 1) + 1):  # xx outer, yy, zz inner
            for yy in range(max(0, iy - 1), min(n[1] - 1, iy + 1) + 1):
                for zz in range(max(0, iz - 1), min(n[2] - 1, iz + 1) + 1):
                    for j in cells.get((xx, yy, zz), ()):
                        if j == i:
                            continue
                        dx, dy, dz = P[i, 0] - P[j, 0], P[i, 1] - P[j, 1], P[i, 2] - P[j, 2]
                        d2 = dx * dx + dy * dy + dz * dz
                        if d2 < r2:
                            lst.append(j)
                            m2 = max(m2, d2)
        out.append(lst)
        maxd.append(math.sqrt(m2))
    return out, maxd


def patch_faces(V, corners, adj, src_face, src_point, target_faces, max_dist):
    """Face SET of the per-source patch (order is irrelevant: pure set closure)."""
    thr2 = max_dist * max_dist
    visited = {int(src_face)}
    goal = {int(t) for t in target_faces if int(t) != int(src_face)}
    if not goal:
        return visited
    stack = []
    for k in range(3):
        g = int(adj[src_face, k])
        if g < 0:
            continue
        visited.add(g)
        stack.append(g)
        goal.discard(g)
    if not goal:
        return visited
    while stack:
        cur = stack.pop()
        for k in range(3):
            g = int(adj[cur, k])
            if g < 0 or g in visited:
                continue
            far = True
            for c in corners[g]:
                d = src_point - V[c]
                if d[0] * d[0] + d[1] * d[1] + d[2] * d[2] <= thr2:
                    far = False
            if far:
                continue
            visited.add(g)
            goal.discard(g)
            stack.append(g)
    return visited | goal
