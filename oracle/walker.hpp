// ORACLE (test infrastructure only). Straightest-geodesic displacement with parallel transport.
// Operation-by-operation restatement of triangulatedMeshSpace::transportParticleAndVectors
// (src/models/triangulatedMeshSpace.cpp:448-564), updateForEdgeIntersection (:614-658),
// rotateAboutAxis (src/utility/functionUtilities.cpp:8-32), the clamps
// (src/utility/meshUtilities.cpp:138-182) and the barycentric edge intersections (:229-331).
//
// Deliberate, documented deviations (SURVEY.md §7 "hard parts", §8(a) q4-q6):
//  * rotation angle: the reference takes theta = acos(n.n') and then sin/cos(theta); here (and in the
//    CUDA kernel) cos = n.n' and sin = |n x n'| so no libm call can differ between host and device.
//    strictTrig=true restores the reference's acos/sin/cos for comparison.
//  * open meshes (openMeshSpace.cpp:114-238): a border edge ends the walk in the closed space (flag), stops the particle on
//    the edge in the absorbing space (absorbingOpenMeshSpace.cpp:2-24) or redirects the displacement along the edge in the
//    tangential space (tangentialOpenMeshSpace.cpp:3-42); transported vectors pointing over the boundary are projected
//    (triangulatedMeshSpace.cpp:411-426).
//  * vertex crossings (two edges hit): updateForVertexIntersection / throughVertex (:247-407, :566-611) are restated with
//    their INTENDED semantics, flagged WALK_VERTEX and counted: the vertex is the corner shared by the two
//    hit edges (the reference takes involvedVertex[0], which is that corner in one of the three cases only); the fan is
//    circulated clockwise from the source face through the face adjacency (= CGAL's Vertex_around_target_circulator started
//    at the successor of the vertex in the source face, :256-300); the outgoing heading leaves half of the total angle on
//    either side (:319-348); the remaining length is |target - vertex| (the reference re-uses the pre-hop length, :578-581);
//    transported vectors keep their angle to the path (the reference rotates them about n x n' only when n.n' < 0, :589).
//    Angles come from detAngle / detSinCos (IEEE + - * / sqrt in a fixed order) instead of acos / sin / cos so that host and
//    device agree bit for bit.  Boundary vertices of open meshes follow openMeshSpace::getBoundaryVertexHeading (openMeshSpace.cpp:3-70,
//    the ring edge that overlaps the heading most) with absorbing / tangentialOpenMeshSpace::updateAtBoundaryVertex.
//  * no-hit and runaway loops do not throw: they set a flag bit and stop in a defined way.
#pragma once
#include "mesh.hpp"
#include <vector>

#ifndef ORC_WALK_TRACE
#define ORC_WALK_TRACE(...) ((void)0)
#endif

namespace orc {

enum WalkFlags { WALK_VERTEX = 1, WALK_NOHIT = 2, WALK_ITERCAP = 4, WALK_NAN = 8, WALK_BORDER = 16 };
static const int WALK_MAX_CROSSINGS = 100000;
enum BoundaryMode { BOUNDARY_CLOSED = 0, BOUNDARY_ABSORBING = 1, BOUNDARY_TANGENTIAL = 2 };

// triangulatedMeshSpace::projectVectorsIfOverBoundary (:411-426) with projectVectorOrthogonalToDirection (meshUtilities.cpp:125-129)
inline void projectVectorsIfOverBoundary(V3* T, int nT, const V3& orthogonal, const V3& inward)
{
    bool pointsOut = dot(orthogonal, inward) < 0;
    for (int i = 0; i < nT; ++i) {
        bool along = dot(T[i], orthogonal) > 0;
        if ((pointsOut && along) || (!pointsOut && !along)) {
            V3 dhat = orthogonal / norm(orthogonal);
            T[i] = T[i] - dot(T[i], dhat) * dhat;
        }
    }
}

inline void belowZeroClamp(double b[3], double tol = 1e-11) // meshUtilities.cpp:138-150
{
    double s = 0;
    for (int i = 0; i < 3; ++i) {
        b[i] = (b[i] < tol) ? tol : b[i]; // std::max(b, tol)
        s += b[i];
    }
    for (int i = 0; i < 3; ++i) b[i] = b[i] / s;
}
inline void nearZeroClamp(double b[3], double tol = 1e-13) // meshUtilities.cpp:157-168
{
    double s = 0;
    for (int i = 0; i < 3; ++i) {
        if (b[i] > -tol && b[i] < tol) b[i] = tol;
        s += b[i];
    }
    for (int i = 0; i < 3; ++i) b[i] = b[i] / s;
}

// functionUtilities.cpp:8-32 with axis = {base, base + a}
inline V3 rotateAboutAxis(const V3& p, const V3& base, const V3& a, double s, double c)
{
    V3 tip = base + a;
    V3 ax = tip - base;
    double an = sqlen(ax);
    ax = ax / std::sqrt(an);
    V3 sh = p - base;
    double dp = ax.x * sh.x + ax.y * sh.y + ax.z * sh.z;
    V3 r;
    r.x = ax.x * dp * (1. - c) + sh.x * c + (ax.y * sh.z - ax.z * sh.y) * s;
    r.y = ax.y * dp * (1. - c) + sh.y * c + (ax.z * sh.x - ax.x * sh.z) * s;
    r.z = ax.z * dp * (1. - c) + sh.z * c + (ax.x * sh.y - ax.y * sh.x) * s;
    return r + base;
}

// S = source bary, E = target bary; returns number of hit edges, k of the last hit, I = last hit point
inline int edgeHits(const double S[3], const double E[3], int lastEdge, int hitK[2], double I[3])
{
    int nh = 0;
    auto record = [&](int k, double t2) {
        I[0] = S[0] + t2 * (E[0] - S[0]);
        I[1] = S[1] + t2 * (E[1] - S[1]);
        I[2] = S[2] + t2 * (E[2] - S[2]);
        hitK[nh < 2 ? nh : 1] = k; // [0] first hit, [1] last hit
        nh++;
    };
    if (lastEdge != 2) { // V1V2 (edge opposite corner 2), meshUtilities.cpp:229-245
        double den = S[0] + S[1] - E[0] - E[1];
        if (den != 0) {
            double t1 = (-E[1] + E[1] * S[0] + S[1] - E[0] * S[1]) / den;
            double t2 = (-1 + S[0] + S[1]) / den;
            if (t1 >= 0 && t1 <= 1 && t2 >= 0 && t2 <= 1) record(2, t2);
        }
    }
    if (lastEdge != 0) { // V2V3 (edge opposite corner 0), :246-263
        double den = -E[0] + S[0];
        if (den != 0) {
            double t1 = -(E[0] - S[0] + E[1] * S[0] - E[0] * S[1]) / den;
            double t2 = (S[0]) / den;
            if (t1 >= 0 && t1 <= 1 && t2 >= 0 && t2 <= 1) record(0, t2);
        }
    }
    if (lastEdge != 1) { // V3V1 (edge opposite corner 1), :264-281
        double den = S[1] - E[1];
        if (den != 0) {
            double t1 = (E[0] * S[1] - E[1] * S[0]) / den;
            double t2 = S[1] / den;
            if (t1 >= 0 && t1 <= 1 && t2 >= 0 && t2 <= 1) record(1, t2);
        }
    }
    return nh;
}


// ---- deterministic trigonometry: IEEE + - * / sqrt only, evaluated in this order by the oracle and by the CUDA walker ----
// angle in [0, pi] of a rotation with sine s >= 0 and cosine c (not necessarily normalised)
inline double detAngle(double s, double c)
{
    const double PI = 3.14159265358979323846;
    double r = std::sqrt(s * s + c * c);
    if (!(r > 0)) return 0.0;
    bool obtuse = c < 0;
    double ca = obtuse ? -c : c;
    double t = s / (r + ca);                  // tan(phi / 2), phi in [0, pi/2]
    t = t / (1.0 + std::sqrt(1.0 + t * t));   // tan(phi / 4)
    t = t / (1.0 + std::sqrt(1.0 + t * t));   // tan(phi / 8) <= 0.0985
    double t2 = t * t;
    double a = 1.0 / 19.0;
    a = 1.0 / 17.0 - t2 * a;
    a = 1.0 / 15.0 - t2 * a;
    a = 1.0 / 13.0 - t2 * a;
    a = 1.0 / 11.0 - t2 * a;
    a = 1.0 / 9.0 - t2 * a;
    a = 1.0 / 7.0 - t2 * a;
    a = 1.0 / 5.0 - t2 * a;
    a = 1.0 / 3.0 - t2 * a;
    a = 1.0 - t2 * a;
    double phi = 8.0 * (t * a);
    return obtuse ? PI - phi : phi;
}
// sine and cosine of x in [0, pi]
inline void detSinCos(double x, double& s, double& c)
{
    double y = x * 0.125, y2 = y * y;
    double ps = 1.0 - y2 / 272.0;             // sin y = y (1 - y2/6 (1 - y2/20 (1 - y2/42 ( ... ))))
    ps = 1.0 - y2 / 210.0 * ps;
    ps = 1.0 - y2 / 156.0 * ps;
    ps = 1.0 - y2 / 110.0 * ps;
    ps = 1.0 - y2 / 72.0 * ps;
    ps = 1.0 - y2 / 42.0 * ps;
    ps = 1.0 - y2 / 20.0 * ps;
    ps = 1.0 - y2 / 6.0 * ps;
    double pc = 1.0 - y2 / 306.0;             // cos y = 1 - y2/2 (1 - y2/12 (1 - y2/30 ( ... )))
    pc = 1.0 - y2 / 240.0 * pc;
    pc = 1.0 - y2 / 182.0 * pc;
    pc = 1.0 - y2 / 132.0 * pc;
    pc = 1.0 - y2 / 90.0 * pc;
    pc = 1.0 - y2 / 56.0 * pc;
    pc = 1.0 - y2 / 30.0 * pc;
    pc = 1.0 - y2 / 12.0 * pc;
    pc = 1.0 - y2 / 2.0 * pc;
    s = y * ps, c = pc;
    for (int i = 0; i < 3; ++i) {             // three angle doublings
        double s2 = 2.0 * s * c, c2 = 1.0 - 2.0 * s * s;
        s = s2, c = c2;
    }
}
inline double angleBetweenUnit(const V3& u, const V3& v) { return detAngle(norm(cross(u, v)), dot(u, v)); } // functionUtilities.cpp:34-45
inline V3 unitTo(const V3& from, const V3& to)
{
    V3 d = to - from;
    return d / norm(d);
}
static const int WALK_MAX_VALENCE = 64;

// one face of the fan of vertex v: v sits at corner kc of face g
struct FanFace {
    int g, kc;
};
// next face clockwise (CGAL Vertex_around_target_circulator++ for counter-clockwise faces): across the edge (v, successor of v)
inline FanFace fanClockwise(const Mesh& m, FanFace a)
{
    int e = (a.kc + 2) % 3, g = m.adj[3 * a.g + e];
    return FanFace{g, g < 0 ? 0 : (m.adjk[3 * a.g + e] + 2) % 3};
}
inline FanFace fanCounterClockwise(const Mesh& m, FanFace a)
{
    int e = (a.kc + 1) % 3, g = m.adj[3 * a.g + e];
    return FanFace{g, g < 0 ? 0 : (m.adjk[3 * a.g + e] + 1) % 3};
}
// interior angle of the fan face at v, and the unit edge vectors to the successor / predecessor of v in that face
inline double fanAngle(const Mesh& m, FanFace a, const V3& Pv, V3& eNext, V3& ePrev)
{
    eNext = unitTo(Pv, m.v[m.c[3 * a.g + (a.kc + 1) % 3]]);
    ePrev = unitTo(Pv, m.v[m.c[3 * a.g + (a.kc + 2) % 3]]);
    return angleBetweenUnit(eNext, ePrev);
}

// Straightest geodesic through the interior vertex at corner kv of face f (throughVertex :247-407 + updateForVertexIntersection
// :566-611, intended semantics).  travel = unit direction of motion inside f.  Returns false when the fan meets a border
// (boundary vertex) or is malformed; otherwise the landing face and the unit heading inside it.
inline bool throughVertex(const Mesh& m, int f, int kv, const V3& travel, int& gOut, V3& heading)
{
    const V3 Pv = m.v[m.c[3 * f + kv]];
    const FanFace src{f, kv};
    V3 eN, eP;
    // pass 1: total angle, sectors in circulation order (first clockwise neighbour ... source face last)
    double total = 0;
    int n = 0;
    for (FanFace a = fanClockwise(m, src);; a = fanClockwise(m, a)) {
        if (a.g < 0 || ++n > WALK_MAX_VALENCE) return false;
        total += fanAngle(m, a, Pv, eN, eP);
        if (a.g == f) break;
    }
    const double half = total / 2.0;
    // pass 2: from the incoming direction (seen from the vertex) clockwise until half of the total angle is used up
    V3 back = V3{0, 0, 0} - travel;
    fanAngle(m, src, Pv, eN, eP);
    double traveled = angleBetweenUnit(back, eN); // :322 firstAngle, measured to the successor of v in the source face
    FanFace land = src;
    if (traveled < half) {
        for (FanFace a = fanClockwise(m, src);; a = fanClockwise(m, a)) {
            traveled += fanAngle(m, a, Pv, eN, eP);
            land = a;
            if (traveled >= half || a.g == f) break;
        }
    }
    // heading: the successor edge of the landing face turned back by (traveled - half) towards its predecessor edge (:334-345)
    double delta = traveled - half, sn, cs;
    if (delta < 0) delta = 0;
    detSinCos(delta, sn, cs);
    V3 u = cross(eN, eP);
    u = u / norm(u);
    V3 w = cross(u, eN);
    heading = cs * eN + sn * w;
    heading = heading / norm(heading);
    gOut = land.g;
    return true;
}

// openMeshSpace::getBoundaryVertexHeading (openMeshSpace.cpp:3-70): among the ring edges of the boundary vertex that belong to a
// face other than the source face, the one that overlaps the heading most.  Returns false when no other face touches v.
inline bool boundaryVertexHeading(const Mesh& m, int f, int kv, const V3& dhat, int& gOut, int& wOut, V3& heading)
{
    const V3 Pv = m.v[m.c[3 * f + kv]];
    const FanFace src{f, kv};
    double best = -1;
    bool found = false;
    for (int dir = 0; dir < 2; ++dir) {
        int n = 0;
        for (FanFace a = dir ? fanCounterClockwise(m, src) : fanClockwise(m, src); a.g >= 0 && a.g != f && ++n <= WALK_MAX_VALENCE;
             a = dir ? fanCounterClockwise(m, a) : fanClockwise(m, a)) {
            for (int j = 1; j <= 2; ++j) {
                int w = m.c[3 * a.g + (a.kc + j) % 3];
                V3 out = unitTo(Pv, m.v[w]);
                double overlap = dot(out, dhat);
                if (overlap > best) best = overlap, heading = out, gOut = a.g, wOut = w, found = true;
            }
        }
    }
    return found;
}

// Returns flag word; crossings (optional) counts edge hops.
inline int transport(const Mesh& m, int& face, double bary[3], V3& disp, V3* T, int nT, bool strictTrig = false, int* crossings = nullptr,
                     int boundaryMode = BOUNDARY_CLOSED)
{
    int flags = 0;
    int f = face;
    double S[3] = {bary[0], bary[1], bary[2]};
    V3 p = m.point(f, S);                                  // :451
    V3 n = m.normal(f);                                    // :461
    double nd = dot(n, disp);
    if (std::fabs(nd) > 1e-14) disp = disp - nd * n;       // :462-465
    V3 q = p + disp;                                       // :466
    int last = -1;
    double E[3];
    int nCross = 0;
    for (;;) {
        const V3 &c0 = m.v[m.c[3 * f]], &c1 = m.v[m.c[3 * f + 1]], &c2 = m.v[m.c[3 * f + 2]];
        ericsonBary(c0, c1, c2, q, E);                     // :478
        nearZeroClamp(E);                                  // :479
        q = m.point(f, E);
        belowZeroClamp(S);                                 // :483
        p = m.point(f, S);
        disp = q - p;                                      // :484
        if (E[0] != E[0]) {                                // checkBaryNan, meshUtilities.cpp:184-195
            flags |= WALK_NAN;
            break;
        }
        if (!(E[0] < 0 || E[1] < 0 || E[2] < 0)) break;    // :490-497
        if (nCross >= WALK_MAX_CROSSINGS) {
            flags |= WALK_ITERCAP;
            belowZeroClamp(E);
            break;
        }
        int hk[2];
        double I[3];
        int nh = edgeHits(S, E, last, hk, I);              // :504-507
        if (nh == 0) {                                     // :534-541 (reference throws)
            flags |= WALK_NOHIT;
            belowZeroClamp(E);
            break;
        }
        belowZeroClamp(I);                                 // :513
        V3 x = m.point(f, I);                              // :514
        S[0] = I[0], S[1] = I[1], S[2] = I[2];             // :517
        p = x;                                             // :518
        int k = nh >= 2 ? hk[1] : hk[0];
        ORC_WALK_TRACE("hop %d face %d nh %d k %d S=(%.3e %.3e %.3e) E=(%.3e %.3e %.3e) |disp| %.3e adj %d\n", nCross, f, nh, k, S[0], S[1], S[2], E[0],
                       E[1], E[2], norm(disp), m.adj[3 * f + k]);
        if (nh >= 2 && hk[0] != hk[1]) {                   // :520-532 two edges hit: the path goes through their common vertex
            flags |= WALK_VERTEX;
            const int kv = 3 - hk[0] - hk[1];
            const V3 travel = disp / norm(disp);           // direction of motion inside f (parallel to toIntersection, :516)
            int g2 = -1;
            V3 heading;
            if (throughVertex(m, f, kv, travel, g2, heading)) {
                V3 n2 = m.normal(g2);
                double rem = norm(q - p);                  // what is left of the displacement beyond the vertex
                V3 side = cross(n, travel), side2 = cross(n2, heading);
                for (int i = 0; i < nT; ++i) {             // parallel transport: same components in the (path, side, normal) frames
                    double ta = dot(T[i], travel), tb = dot(T[i], side), tc = dot(T[i], n);
                    T[i] = ta * heading + tb * side2 + tc * n2;
                }
                disp = rem * heading;
                q = p + disp;
                f = g2;
                ericsonBary(m.v[m.c[3 * f]], m.v[m.c[3 * f + 1]], m.v[m.c[3 * f + 2]], p, S); // :606
                n = n2;
                last = -1;                                 // :609-610
                nCross++;
                continue;
            }
            // boundary vertex (the fan is open): closed space stops (the reference throws, :522-523)
            flags |= WALK_BORDER;
            int w = -1;
            if (boundaryMode == BOUNDARY_CLOSED || !boundaryVertexHeading(m, f, kv, travel, g2, w, heading)) {
                E[0] = S[0], E[1] = S[1], E[2] = S[2];
                break;
            }
            const int vIdx = m.c[3 * f + kv];
            const V3 Pv = m.v[vIdx];
            V3 rest = q - p;                               // remaining displacement
            f = g2;                                        // openMeshSpace.cpp:63-67: the source moves to the vertex, seen from the new face
            n = m.normal(f);
            ericsonBary(m.v[m.c[3 * f]], m.v[m.c[3 * f + 1]], m.v[m.c[3 * f + 2]], Pv, S);
            if (nT > 0) {                                  // projectVectorsForBoundaryVertex :72-100
                V3 orth = cross(n, heading);
                orth = orth / norm(orth);
                V3 inside = Pv;
                for (int j = 0; j < 3; ++j) {
                    int cv = m.c[3 * f + j];
                    if (cv != vIdx && cv != w) inside = m.v[cv];
                }
                projectVectorsIfOverBoundary(T, nT, orth, inside - Pv);
            }
            last = -1;
            if (boundaryMode == BOUNDARY_ABSORBING) {      // absorbingOpenMeshSpace.cpp:26-50: stop at the vertex
                E[0] = S[0], E[1] = S[1], E[2] = S[2];
                break;
            }
            double slide = dot(rest, heading);             // tangentialOpenMeshSpace.cpp:42-64: projectVectorOntoDirection
            if (!(slide > 1e-9 * norm(m.v[w] - Pv))) {
                E[0] = S[0], E[1] = S[1], E[2] = S[2];
                break;
            }
            disp = slide * heading;
            p = m.point(f, S);
            q = p + disp;
            nCross++;
            continue;
        }
        int g = m.adj[3 * f + k];
        if (g < 0) {                                       // border edge
            flags |= WALK_BORDER;
            if (boundaryMode == BOUNDARY_CLOSED) {         // :547-548 (the closed space throws)
                E[0] = S[0], E[1] = S[1], E[2] = S[2];
                break;
            }
            const V3 &ev1 = m.v[m.c[3 * f + (k + 1) % 3]], &ev2 = m.v[m.c[3 * f + (k + 2) % 3]], &inner = m.v[m.c[3 * f + k]];
            V3 edge = ev2 - ev1;
            V3 orth = cross(n, edge);
            orth = orth / norm(orth);
            V3 inward = inner - p;                         // p = point(sourceFace, sourceBCs) = the intersection point
            if (boundaryMode == BOUNDARY_ABSORBING) {      // absorbingOpenMeshSpace.cpp:2-24: stop on the edge
                if (nT > 0) projectVectorsIfOverBoundary(T, nT, orth, inward);
                E[0] = S[0], E[1] = S[1], E[2] = S[2];
                break;
            }
            // tangentialOpenMeshSpace.cpp:3-42: slide along the edge in the direction that overlaps the displacement
            V3 fwd = edge / norm(edge);
            V3 back = ev1 - ev2;
            V3 bwd = back / norm(back);
            double len = norm(disp);                       // NB the reference uses the pre-hop displacement (source -> target)
            V3 dhat = disp / len;
            double fd = dot(dhat, fwd), bd = dot(dhat, bwd);
            projectVectorsIfOverBoundary(T, nT, orth, inward);
            double slide = (fd > bd ? fd : bd) * len;
            // exactly perpendicular: nothing left to slide (the reference then leaves the particle at a target outside the
            // face; we stop on the edge).  Likewise when the slide is below 1e-9 edge lengths, i.e. at the scale of the 1e-11 barycentric clamp of the source
            // point: in a corner of the sheet the two clamps otherwise feed a limit cycle (the reference never leaves it).
            if (!(slide > 1e-9 * norm(edge))) {
                E[0] = S[0], E[1] = S[1], E[2] = S[2];
                break;
            }
            if (fd > bd) disp = (fd * len) * fwd;
            else disp = (bd * len) * bwd;
            q = p + disp;
            last = -1;
            nCross++;
            continue;
        }
        V3 n2 = m.normal(g);                               // :626
        double c = dot(n, n2);                             // :627
        V3 ax = cross(n, n2);                              // :632
        double an = norm(ax);
        if (c < 1 && an > 0) {                             // :630 (an>0 guard: reference would divide by zero)
            double s = an, cc = c;
            if (strictTrig) {
                double th = std::acos(c);                  // :628
                s = std::sin(th);
                cc = std::cos(th);
            }
            ax = ax / an;                                  // :633
            q = rotateAboutAxis(q, p, ax, s, cc);          // :636
            disp = q - p;                                  // :637
            for (int i = 0; i < nT; ++i) {                 // :639-645
                V3 tt = p + T[i];
                tt = rotateAboutAxis(tt, p, ax, s, cc);
                T[i] = tt - p;
            }
        }
        last = m.adjk[3 * f + k];                          // :649
        f = g;                                             // :652
        ericsonBary(m.v[m.c[3 * f]], m.v[m.c[3 * f + 1]], m.v[m.c[3 * f + 2]], p, S); // :655
        n = n2;                                            // :657
        nCross++;
    }
    face = f;                                              // :561-562
    bary[0] = E[0], bary[1] = E[1], bary[2] = E[2];
    if (crossings) *crossings = nCross;
    return flags;
}

} // namespace orc
