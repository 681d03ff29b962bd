"""Independent exact-geodesic checker for SMALL patches (test infrastructure).

It shares no code and no pruning logic with the oracle (oracle/geodesic.hpp) or the CUDA kernel:
  1. for every "emitter" (the source point and every patch vertex) it enumerates, by plain recursion
     and with no dominance filter, every straight unfolded segment that leaves the emitter, crosses a
     strip of faces and reaches a vertex or a target point ("visibility legs");
  2. Dijkstra over {source} U vertices U targets with those legs as edges.
Every leg is a real path on the surface, so the result is an upper bound; it is exact because the true
polyhedral geodesic is a chain of such legs bending only at vertices.  Exponential in principle, fine
for <= ~60 faces.
"""
from __future__ import annotations

import heapq
import math

import numpy as np

TOL = 1e-12


def _sub(a, b):
    return (a[0] - b[0], a[1] - b[1])


def _cross(a, b):
    return a[0] * b[1] - a[1] * b[0]


def _norm(a):
    return math.hypot(a[0], a[1])


class BruteGeodesic:
    def __init__(self, V, corners, faces=None):
        self.V = np.asarray(V, dtype=np.float64)
        C = np.asarray(corners, dtype=np.int64)
        self.faces = list(range(len(C))) if faces is None else [int(f) for f in faces]
        self.C = {f: tuple(int(x) for x in C[f]) for f in self.faces}
        half = {}
        for f in self.faces:
            c = self.C[f]
            for k in range(3):
                half[(c[(k + 1) % 3], c[(k + 2) % 3])] = (f, k)
        self.adj = {}
        for f in self.faces:
            c = self.C[f]
            for k in range(3):
                self.adj[(f, k)] = half.get((c[(k + 2) % 3], c[(k + 1) % 3]))
        self.verts = sorted({v for f in self.faces for v in self.C[f]})
        self.inc = {v: [] for v in self.verts}
        for f in self.faces:
            for k in range(3):
                self.inc[self.C[f][k]].append((f, k))

    # place the third corner given the entry edge A->B (2-D) of face g entered through edge e
    def _third(self, g, e, A, B):
        c = self.C[g]
        PA, PB, PC = self.V[c[(e + 1) % 3]], self.V[c[(e + 2) % 3]], self.V[c[e]]
        ab = PB - PA
        L = float(np.linalg.norm(ab))
        x = float(np.dot(PC - PA, ab)) / L
        y = float(np.linalg.norm(np.cross(ab, PC - PA))) / L
        ux, uy = (B[0] - A[0]), (B[1] - A[1])
        l2 = math.hypot(ux, uy)
        ux, uy = ux / l2, uy / l2
        return (A[0] + x * ux - y * uy, A[1] + x * uy + y * ux)

    @staticmethod
    def _hit(S, P, X, Y):
        d = _sub(P, S)
        den = _cross(_sub(Y, X), d)
        if den == 0:
            return None
        return _cross(_sub(S, X), d) / den

    def _lift(self, g, Q, d):
        """2-D direction d in the unfolded frame where face g has corner images Q[0..2] -> unit 3-D vector."""
        c = self.C[g]
        e1, e2 = _sub(Q[1], Q[0]), _sub(Q[2], Q[0])
        det = _cross(e1, e2)
        al, be = _cross(d, e2) / det, _cross(e1, d) / det
        r = al * (self.V[c[1]] - self.V[c[0]]) + be * (self.V[c[2]] - self.V[c[0]])
        return r / np.linalg.norm(r)

    def _unfold(self, g, e, A, B, S, PL, PR, depth, out, first):
        """S sees the part [PL, PR] of edge A->B (PL nearer A); record what it sees inside/through face g.
        `first` = function mapping a 2-D direction at the emitter to the 3-D start direction."""
        if depth <= 0:
            return
        c = self.C[g]
        iA, iB, iC = (e + 1) % 3, (e + 2) % 3, e
        Cp = self._third(g, e, A, B)
        Q = [None, None, None]
        Q[iA], Q[iB], Q[iC] = A, B, Cp
        AB = _sub(B, A)
        # targets in this face
        for t in self.face_targets.get(g, ()):
            b = self.tbary[t]
            s = b[0] + b[1] + b[2]
            T = ((b[0] * Q[0][0] + b[1] * Q[1][0] + b[2] * Q[2][0]) / s, (b[0] * Q[0][1] + b[1] * Q[1][1] + b[2] * Q[2][1]) / s)
            d = _sub(T, S)
            den = _cross(AB, d)
            if den == 0:
                continue
            mu = _cross(_sub(S, A), d) / den
            X = (A[0] + mu * AB[0], A[1] + mu * AB[1])
            # inside the visible part?
            if _cross(_sub(PL, S), _sub(X, S)) <= TOL and _cross(_sub(PR, S), _sub(X, S)) >= -TOL:
                out.append((("t", t), _norm(d), first(d), self._lift(g, Q, d)))
        dL, dR, dC = _sub(PL, S), _sub(PR, S), _sub(Cp, S)
        sL, sR = _cross(dL, dC), _cross(dR, dC)
        sc = _norm(dC)
        eL, eR = TOL * _norm(dL) * sc, TOL * _norm(dR) * sc
        inside = sL <= eL and sR >= -eR
        if inside:
            out.append((("v", c[iC]), sc, first(dC), self._lift(g, Q, dC)))
        # edge C->A of g (opposite corner B): neighbour sees A->C
        if sL <= eL:
            nb = self.adj[(g, iB)]
            if nb is not None:
                m0 = self._hit(S, PL, A, Cp)
                m1 = 1.0 if inside else self._hit(S, PR, A, Cp)
                if m0 is not None and m1 is not None:
                    m0, m1 = max(0.0, min(1.0, m0)), max(0.0, min(1.0, m1))
                    if m1 - m0 > 1e-13:
                        AC = _sub(Cp, A)
                        self._unfold(nb[0], nb[1], A, Cp, S, (A[0] + m0 * AC[0], A[1] + m0 * AC[1]),
                                     (A[0] + m1 * AC[0], A[1] + m1 * AC[1]), depth - 1, out, first)
        # edge B->C of g (opposite corner A): neighbour sees C->B
        if sR >= -eR:
            nb = self.adj[(g, iA)]
            if nb is not None:
                m0 = 0.0 if inside else self._hit(S, PL, Cp, B)
                m1 = self._hit(S, PR, Cp, B)
                if m0 is not None and m1 is not None:
                    m0, m1 = max(0.0, min(1.0, m0)), max(0.0, min(1.0, m1))
                    if m1 - m0 > 1e-13:
                        CB = _sub(B, Cp)
                        self._unfold(nb[0], nb[1], Cp, B, S, (Cp[0] + m0 * CB[0], Cp[1] + m0 * CB[1]),
                                     (Cp[0] + m1 * CB[0], Cp[1] + m1 * CB[1]), depth - 1, out, first)

    def _frame(self, f):
        c = self.C[f]
        P0, P1, P2 = self.V[c[0]], self.V[c[1]], self.V[c[2]]
        L = float(np.linalg.norm(P1 - P0))
        return [(0.0, 0.0), (L, 0.0), (float(np.dot(P2 - P0, P1 - P0)) / L, float(np.linalg.norm(np.cross(P1 - P0, P2 - P0))) / L)]

    def _legs_from_point(self, f0, bary, depth):
        out = []
        Q = self._frame(f0)
        s = bary[0] + bary[1] + bary[2]
        S = ((bary[0] * Q[0][0] + bary[1] * Q[1][0] + bary[2] * Q[2][0]) / s, (bary[0] * Q[0][1] + bary[1] * Q[1][1] + bary[2] * Q[2][1]) / s)
        c = self.C[f0]
        S3 = (bary[0] * self.V[c[0]] + bary[1] * self.V[c[1]] + bary[2] * self.V[c[2]]) / s

        def first(d):
            return self._lift(f0, Q, d)

        for k in range(3):
            d = self.V[c[k]] - S3
            L = float(np.linalg.norm(d))
            out.append((("v", c[k]), L, d / L, d / L))
        for t in self.face_targets.get(f0, ()):
            d = self.tpos[t] - S3
            L = float(np.linalg.norm(d))
            out.append((("t", t), L, d / L, d / L))
        for k in range(3):
            nb = self.adj[(f0, k)]
            if nb is None:
                continue
            A, B = Q[(k + 2) % 3], Q[(k + 1) % 3]
            self._unfold(nb[0], nb[1], A, B, S, A, B, depth, out, first)
        return out

    def _legs_from_vertex(self, v, depth):
        out = []
        pv = self.V[v]
        for (g, i) in self.inc[v]:
            c = self.C[g]
            vp, vq = c[(i + 1) % 3], c[(i + 2) % 3]
            ep, eq = self.V[vp] - pv, self.V[vq] - pv
            lp, lq = float(np.linalg.norm(ep)), float(np.linalg.norm(eq))
            out.append((("v", vp), lp, ep / lp, ep / lp))
            out.append((("v", vq), lq, eq / lq, eq / lq))
            for t in self.face_targets.get(g, ()):
                d = self.tpos[t] - pv
                L = float(np.linalg.norm(d))
                out.append((("t", t), L, d / L, d / L))
            nb = self.adj[(g, i)]
            if nb is None:
                continue
            P2 = (lp, 0.0)
            Q2 = (float(np.dot(eq, ep)) / lp, float(np.linalg.norm(np.cross(ep, eq))) / lp)
            Qf = [None, None, None]
            Qf[i], Qf[(i + 1) % 3], Qf[(i + 2) % 3] = (0.0, 0.0), P2, Q2

            def first(d, g=g, Qf=Qf):
                return self._lift(g, Qf, d)

            self._unfold(nb[0], nb[1], Q2, P2, (0.0, 0.0), Q2, P2, depth, out, first)
        return out

    def solve(self, src_face, src_bary, tfaces, tbary, depth=40):
        """Returns (dist[K], start_tangent[K,3], end_tangent[K,3]); dist = inf when unreachable."""
        K = len(tfaces)
        self.tbary = [tuple(float(x) for x in b) for b in tbary]
        self.tpos = []
        self.face_targets = {}
        for t in range(K):
            f = int(tfaces[t])
            c = self.C[f]
            b = self.tbary[t]
            self.tpos.append((b[0] * self.V[c[0]] + b[1] * self.V[c[1]] + b[2] * self.V[c[2]]) / (b[0] + b[1] + b[2]))
            self.face_targets.setdefault(f, []).append(t)
        dist = {("s", 0): 0.0}
        start = {}
        end = {}
        heap = [(0.0, 0, ("s", 0))]
        done = set()
        cnt = 1
        while heap:
            d, _, node = heapq.heappop(heap)
            if node in done:
                continue
            done.add(node)
            if node[0] == "t":
                continue
            legs = self._legs_from_point(int(src_face), tuple(float(x) for x in src_bary), depth) if node[0] == "s" \
                else self._legs_from_vertex(node[1], depth)
            for (to, L, d0, d1) in legs:
                nd = d + L
                if nd < dist.get(to, math.inf):
                    dist[to] = nd
                    start[to] = d0 if node[0] == "s" else start[node]
                    end[to] = d1
                    heapq.heappush(heap, (nd, cnt, to))
                    cnt += 1
        D = np.full(K, np.inf)
        TS = np.zeros((K, 3))
        TE = np.zeros((K, 3))
        for t in range(K):
            if ("t", t) in dist:
                D[t] = dist[("t", t)]
                TS[t] = start[("t", t)]
                TE[t] = end[("t", t)]
        return D, TS, TE
