// ORACLE (test infrastructure only). Exact polyhedral geodesics from one source on a patch.
//
// PARITY UNPINNED: this replaces CGAL 5.6 Surface_mesh_shortest_path (not vendored, not installable
// here), as used by the reference at src/models/triangulatedMeshSpace.cpp:189-196 (add_source_point,
// build_sequence_tree on the per-source submesh), :229-237 (global mesh) and
// src/utility/meshUtilities.cpp:360-380 (shortest_path_points_to_source_points -> distance, start and
// end tangents).  The reference ships no golden vectors for it.  The algorithm restated here is the
// published one CGAL implements (Chen & Han 1990 window unfolding with the Xin & Wang 2009 vertex-
// distance filter, Citations.md:8-21): windows are unfolded face by face in the plane of their
// (pseudo-)source; saddle vertices and patch-boundary vertices act as pseudo-sources; a query is the
// minimum over the windows entering the target's face and over the three corners of that face.  This
// oracle processes events in Dijkstra order with a binary heap (the CUDA kernel uses a different,
// breadth-first order), so agreement between the two is a real check.  Its own correctness is pinned
// by tests/: closed-form cases (plane, cube, prisms, notch) and an independent exhaustive-unfolding
// checker.
//
// Output convention (meshUtilities.cpp:368-379): both tangents are unit vectors in the source->target
// sense; "start" is at the source, "end" is at the target.
#pragma once
#include "mesh.hpp"
#include <limits>
#include <queue>
#include <unordered_map>
#include <vector>

namespace orc {

struct GeoTarget {
    int face;
    double b[3];
};
struct GeoResult {
    double dist = -1.0; // < 0 : unreachable (other connected component of the patch)
    V3 ts{0, 0, 0}, te{0, 0, 0};
    int tie = 0;        // two distinct candidates within 1e-12 relative
};
struct GeoStats {
    long windowsCreated = 0, windowsProcessed = 0, pseudoSources = 0, faces = 0, verts = 0;
    void add(const GeoStats& o)
    {
        windowsCreated += o.windowsCreated;
        windowsProcessed += o.windowsProcessed;
        pseudoSources += o.pseudoSources;
        faces += o.faces;
        verts += o.verts;
    }
};

// Global per-vertex flag: interior angle sum >= 2 pi (saddle or flat) -> may be a pseudo-source.
inline std::vector<char> saddleFlags(const Mesh& m)
{
    std::vector<double> ang(m.nV, 0.0);
    for (int f = 0; f < m.nF; ++f)
        for (int k = 0; k < 3; ++k) {
            const V3& p = m.v[m.c[3 * f + k]];
            V3 a = m.v[m.c[3 * f + (k + 1) % 3]] - p, b = m.v[m.c[3 * f + (k + 2) % 3]] - p;
            ang[m.c[3 * f + k]] += std::atan2(norm(cross(a, b)), dot(a, b));
        }
    std::vector<char> s(m.nV);
    for (int i = 0; i < m.nV; ++i) s[i] = ang[i] >= 2.0 * M_PI - 1e-9;
    return s;
}

class PatchGeodesic {
public:
    // faces == nullptr : the whole mesh is the patch (global branch, triangulatedMeshSpace.cpp:222-237)
    PatchGeodesic(const Mesh& mesh, const std::vector<char>& saddle, const std::vector<int>* faces) : m(mesh)
    {
        if (!faces) {
            nF = m.nF;
            nV = m.nV;
            whole = true;
        } else {
            whole = false;
            gface = *faces;
            nF = (int)gface.size();
            std::unordered_map<int, int> vmap;
            for (int i = 0; i < nF; ++i) fmap[gface[i]] = i;
            fv.resize(3 * nF);
            fadj.resize(3 * nF);
            for (int i = 0; i < nF; ++i)
                for (int k = 0; k < 3; ++k) {
                    int gv = m.c[3 * gface[i] + k];
                    auto it = vmap.find(gv);
                    if (it == vmap.end()) {
                        it = vmap.emplace(gv, (int)gvert.size()).first;
                        gvert.push_back(gv);
                    }
                    fv[3 * i + k] = it->second;
                    int ga = m.adj[3 * gface[i] + k];
                    auto fit = ga < 0 ? fmap.end() : fmap.find(ga);
                    fadj[3 * i + k] = fit == fmap.end() ? -1 : fit->second;
                }
            nV = (int)gvert.size();
        }
        // incident faces CSR + pseudo-source eligibility
        incStart.assign(nV + 1, 0);
        for (int i = 0; i < 3 * nF; ++i) incStart[V(i / 3, i % 3) + 1]++;
        for (int i = 0; i < nV; ++i) incStart[i + 1] += incStart[i];
        inc.resize(3 * (size_t)nF);
        std::vector<int> fill(incStart.begin(), incStart.end() - 1);
        for (int f = 0; f < nF; ++f)
            for (int k = 0; k < 3; ++k) inc[fill[V(f, k)]++] = 3 * f + k;
        elig.assign(nV, 0);
        for (int v = 0; v < nV; ++v) elig[v] = saddle[whole ? v : gvert[v]];
        for (int f = 0; f < nF; ++f)
            for (int k = 0; k < 3; ++k)
                if (ADJ(f, k) < 0) { // border edge k: its endpoints are corners k+1, k+2
                    elig[V(f, (k + 1) % 3)] = 1;
                    elig[V(f, (k + 2) % 3)] = 1;
                }
    }

    int localFace(int g) const
    {
        if (whole) return g;
        auto it = fmap.find(g);
        return it == fmap.end() ? -1 : it->second;
    }

    void solve(int srcFaceGlobal, const double sb[3], const std::vector<GeoTarget>& tg, std::vector<GeoResult>& out,
               GeoStats* stats = nullptr)
    {
        const double INF = std::numeric_limits<double>::infinity();
        int nT = (int)tg.size();
        out.assign(nT, GeoResult());
        best.assign(nT, INF);
        second.assign(nT, INF);
        D.assign(nV, INF);
        dir.assign(nV, V3{0, 0, 0});
        faceTargets.clear();
        tpos.resize(nT);
        heap = decltype(heap)();
        GeoStats st;
        st.faces = nF;
        st.verts = nV;

        int f0 = localFace(srcFaceGlobal);
        S3 = m.point(srcFaceGlobal, sb);
        for (int t = 0; t < nT; ++t) {
            int lf = localFace(tg[t].face);
            tpos[t] = m.point(tg[t].face, tg[t].b);
            if (lf < 0) continue; // cannot happen for patches built by patchFaces
            if (lf == f0) {       // same planar face: the chord
                V3 d = tpos[t] - S3;
                double L = norm(d);
                candidate(t, L, d / L, d / L, out);
            } else
                faceTargets[lf].push_back(t);
        }
        res = &out;
        tgs = &tg;
        updateBound();

        // root frame: corner 0 at the origin, corner 1 on +x, corner 2 above
        V3 P[3] = {X(f0, 0), X(f0, 1), X(f0, 2)};
        rootQ[0] = V2{0, 0};
        double L01 = norm(P[1] - P[0]);
        rootQ[1] = V2{L01, 0};
        rootQ[2] = V2{dot(P[2] - P[0], P[1] - P[0]) / L01, norm(cross(P[1] - P[0], P[2] - P[0])) / L01};
        for (int k = 0; k < 3; ++k) rootP[k] = P[k];
        double bs = sb[0] + sb[1] + sb[2];
        V2 S2{(sb[0] * rootQ[0].x + sb[1] * rootQ[1].x + sb[2] * rootQ[2].x) / bs,
              (sb[0] * rootQ[0].y + sb[1] * rootQ[1].y + sb[2] * rootQ[2].y) / bs};
        for (int k = 0; k < 3; ++k) {
            V3 d = P[k] - S3;
            double L = norm(d);
            relax(V(f0, k), L, d / L);
        }
        for (int k = 0; k < 3; ++k) {
            int g = ADJ(f0, k);
            if (g < 0) continue;
            Win w;
            w.face = g;
            w.e = ADJK(f0, k);
            w.A = rootQ[(k + 2) % 3];
            w.B = rootQ[(k + 1) % 3];
            w.S = S2;
            w.t0 = 0;
            w.t1 = 1;
            w.sigma = 0;
            w.psv = -1;
            push(w, st);
        }

        while (!heap.empty()) {
            Ev ev = heap.top();
            if (ev.key > bound * (1 + 1e-12)) break;
            heap.pop();
            if (ev.kind == 1) {
                if (ev.key > D[ev.v]) continue; // stale
                st.pseudoSources++;
                spawnFan(ev.v, st);
            } else {
                st.windowsProcessed++;
                propagate(ev.w, st);
            }
        }
        for (int t = 0; t < nT; ++t)
            if (best[t] < INF) {
                out[t].dist = best[t];
                out[t].tie = second[t] < INF && (second[t] - best[t]) <= 1e-12 * best[t];
            }
        if (stats) stats->add(st);
    }

    int nF = 0, nV = 0;

private:
    struct Win {
        int face, e, psv;
        V2 A, B, S;
        double t0, t1, sigma;
    };
    struct Ev {
        double key;
        int kind, v;
        Win w;
        bool operator<(const Ev& o) const { return key > o.key; }
    };

    const Mesh& m;
    bool whole = false;
    std::vector<int> gface, gvert, fv, fadj, incStart, inc;
    std::unordered_map<int, int> fmap;
    std::vector<char> elig;
    std::vector<double> D, best, second;
    std::vector<V3> dir, tpos;
    std::unordered_map<int, std::vector<int>> faceTargets;
    std::priority_queue<Ev> heap;
    std::vector<GeoResult>* res = nullptr;
    const std::vector<GeoTarget>* tgs = nullptr;
    double bound = 0;
    V3 S3;
    V2 rootQ[3];
    V3 rootP[3];

    int G(int f) const { return whole ? f : gface[f]; }
    int V(int f, int k) const { return whole ? m.c[3 * f + k] : fv[3 * f + k]; }
    int ADJ(int f, int k) const { return whole ? m.adj[3 * f + k] : fadj[3 * f + k]; }
    int ADJK(int f, int k) const { return m.adjk[3 * G(f) + k]; }
    const V3& X(int f, int k) const { return m.v[m.c[3 * G(f) + k]]; }
    const V3& XV(int v) const { return m.v[whole ? v : gvert[v]]; }

    void updateBound()
    {
        bound = 0;
        for (double b : best) bound = std::max(bound, b);
    }
    void candidate(int t, double d, const V3& ts, const V3& te, std::vector<GeoResult>& out)
    {
        if (d < best[t]) {
            if (dot(ts, out[t].ts) < 1 - 1e-9 || best[t] - d > 1e-12 * d) second[t] = best[t];
            best[t] = d;
            out[t].ts = ts;
            out[t].te = te;
            updateBound();
        } else if (d < second[t] && dot(ts, out[t].ts) < 1 - 1e-9)
            second[t] = d;
    }
    // direction given in a face's unfolded 2-D frame -> unit 3-D vector in that face's plane
    static V3 lift(const V2 Q[3], const V3 P[3], const V2& d)
    {
        V2 e1 = Q[1] - Q[0], e2 = Q[2] - Q[0];
        double det = cross2(e1, e2);
        double al = cross2(d, e2) / det, be = cross2(e1, d) / det;
        V3 r = al * (P[1] - P[0]) + be * (P[2] - P[0]);
        return r / norm(r);
    }
    void relax(int v, double d, const V3& startDir)
    {
        if (!(d < D[v])) return;
        D[v] = d;
        dir[v] = startDir;
        const V3& pv = XV(v);
        for (int s = incStart[v]; s < incStart[v + 1]; ++s) {
            auto it = faceTargets.find(inc[s] / 3);
            if (it == faceTargets.end()) continue;
            for (int t : it->second) {
                V3 e = tpos[t] - pv;
                double L = norm(e);
                candidate(t, d + L, startDir, e / L, *res);
            }
        }
        if (elig[v]) {
            Ev ev;
            ev.key = d;
            ev.kind = 1;
            ev.v = v;
            heap.push(ev);
        }
    }
    static double segDist(const V2& S, const V2& X0, const V2& X1)
    {
        V2 e = X1 - X0;
        double L2 = dot2(e, e);
        double s = L2 > 0 ? dot2(S - X0, e) / L2 : 0;
        s = std::max(0.0, std::min(1.0, s));
        return norm2(S - (X0 + s * e));
    }
    void push(const Win& w, GeoStats& st)
    {
        V2 X0 = w.A + w.t0 * (w.B - w.A), X1 = w.A + w.t1 * (w.B - w.A);
        Ev ev;
        ev.key = w.sigma + segDist(w.S, X0, X1);
        if (ev.key > bound * (1 + 1e-12)) return;
        ev.kind = 0;
        ev.v = -1;
        ev.w = w;
        heap.push(ev);
        st.windowsCreated++;
    }
    void spawnFan(int v, GeoStats& st)
    {
        const V3& pv = XV(v);
        for (int s = incStart[v]; s < incStart[v + 1]; ++s) {
            int g = inc[s] / 3, i = inc[s] % 3;
            int vp = V(g, (i + 1) % 3), vq = V(g, (i + 2) % 3);
            V3 ep = XV(vp) - pv, eq = XV(vq) - pv;
            double lp = norm(ep), lq = norm(eq);
            relax(vp, D[v] + lp, dir[v]);
            relax(vq, D[v] + lq, dir[v]);
            int g2 = ADJ(g, i);
            if (g2 < 0) continue;
            Win w;
            w.face = g2;
            w.e = ADJK(g, i);
            w.B = V2{lp, 0};
            w.A = V2{dot(eq, ep) / lp, norm(cross(ep, eq)) / lp};
            w.S = V2{0, 0};
            w.t0 = 0;
            w.t1 = 1;
            w.sigma = D[v];
            w.psv = v;
            push(w, st);
        }
    }
    // ray S->P against segment X + mu (Y - X)
    static double hit(const V2& S, const V2& P, const V2& X, const V2& Y)
    {
        V2 d = P - S;
        double den = cross2(Y - X, d);
        double mu = cross2(S - X, d) / den;
        if (!(mu == mu)) mu = 0.5;
        return std::max(0.0, std::min(1.0, mu));
    }
    bool dominated(double sigma, const V2& S, int v, const V2& pv, const V2& Xend) const
    {
        // Xin-Wang filter: the window end Xend is reached strictly shorter through vertex v
        return D[v] + norm2(pv - Xend) < (sigma + norm2(S - Xend)) * (1 - 1e-12);
    }
    void propagate(const Win& w, GeoStats& st)
    {
        int g = w.face, e = w.e;
        int iA = (e + 1) % 3, iB = (e + 2) % 3, iC = e;
        int vA = V(g, iA), vB = V(g, iB), vC = V(g, iC);
        const V3 &PA = X(g, iA), &PB = X(g, iB), &PC = X(g, iC);
        V2 AB = w.B - w.A;
        double L2d = norm2(AB);
        V2 u{AB.x / L2d, AB.y / L2d};
        double L3 = norm(PB - PA);
        double cx = dot(PC - PA, PB - PA) / L3, cy = norm(cross(PC - PA, PB - PA)) / L3;
        V2 C{w.A.x + cx * u.x - cy * u.y, w.A.y + cx * u.y + cy * u.x};
        V2 P0 = w.A + w.t0 * AB, P1 = w.A + w.t1 * AB;

        // queries: targets inside the entered face
        auto it = faceTargets.find(g);
        if (it != faceTargets.end()) {
            V2 Q[3];
            V3 P3[3];
            Q[iA] = w.A, Q[iB] = w.B, Q[iC] = C;
            P3[iA] = PA, P3[iB] = PB, P3[iC] = PC;
            for (int t : it->second) {
                const double* b = (*tgs)[t].b;
                double bs = b[0] + b[1] + b[2];
                V2 T{(b[0] * Q[0].x + b[1] * Q[1].x + b[2] * Q[2].x) / bs, (b[0] * Q[0].y + b[1] * Q[1].y + b[2] * Q[2].y) / bs};
                V2 d = T - w.S;
                double den = cross2(AB, d);
                if (den == 0) continue;
                double mu = cross2(w.S - w.A, d) / den;
                if (mu < w.t0 - 1e-12 || mu > w.t1 + 1e-12) continue;
                double L = norm2(d);
                V3 te = lift(Q, P3, d);
                V3 ts = w.psv < 0 ? lift(rootQ, rootP, d) : dir[w.psv];
                candidate(t, w.sigma + L, ts, te, *res);
            }
        }

        V2 dL = P0 - w.S, dR = P1 - w.S, dC = C - w.S;
        double sideL = cross2(dL, dC), sideR = cross2(dR, dC);
        double lc = norm2(dC);
        double epsL = 1e-12 * norm2(dL) * lc, epsR = 1e-12 * norm2(dR) * lc;
        int kAC = iB, kCB = iA; // edge A-C is opposite corner B; edge C-B is opposite corner A
        bool inside = !(sideL > epsL) && !(sideR < -epsR);
        if (inside) {
            V3 sd = w.psv < 0 ? lift(rootQ, rootP, dC) : dir[w.psv];
            relax(vC, w.sigma + lc, sd);
        }
        // left child: edge C->A of this face, seen from the neighbour as A->C
        if (!(sideL > epsL)) {
            int g2 = ADJ(g, kAC);
            if (g2 >= 0) {
                double m0 = hit(w.S, P0, w.A, C);
                double m1 = inside ? 1.0 : hit(w.S, P1, w.A, C);
                if (m1 - m0 > 1e-13) {
                    V2 XA = w.A + m0 * (C - w.A), XC = w.A + m1 * (C - w.A);
                    if (!dominated(w.sigma, w.S, vA, w.A, XC) && !dominated(w.sigma, w.S, vC, C, XA)
                        && !dominated(w.sigma, w.S, vB, w.B, XA)) {
                        Win c;
                        c.face = g2;
                        c.e = ADJK(g, kAC);
                        c.A = w.A;
                        c.B = C;
                        c.S = w.S;
                        c.t0 = m0;
                        c.t1 = m1;
                        c.sigma = w.sigma;
                        c.psv = w.psv;
                        push(c, st);
                    }
                }
            }
        }
        // right child: edge B->C of this face, seen from the neighbour as C->B
        if (!(sideR < -epsR)) {
            int g2 = ADJ(g, kCB);
            if (g2 >= 0) {
                double m0 = inside ? 0.0 : hit(w.S, P0, C, w.B);
                double m1 = hit(w.S, P1, C, w.B);
                if (m1 - m0 > 1e-13) {
                    V2 XC = C + m0 * (w.B - C), XB = C + m1 * (w.B - C);
                    if (!dominated(w.sigma, w.S, vB, w.B, XC) && !dominated(w.sigma, w.S, vC, C, XB)
                        && !dominated(w.sigma, w.S, vA, w.A, XB)) {
                        Win c;
                        c.face = g2;
                        c.e = ADJK(g, kCB);
                        c.A = C;
                        c.B = w.B;
                        c.S = w.S;
                        c.t0 = m0;
                        c.t1 = m1;
                        c.sigma = w.sigma;
                        c.psv = w.psv;
                        push(c, st);
                    }
                }
            }
        }
    }
};

} // namespace orc
