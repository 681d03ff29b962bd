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
//    (triangulatedMeshSpace.cpp:411-426).  Boundary VERTEX events take the edge rule of the last edge hit and are flagged.
//  * vertex crossings (two edges hit), no-hit and runaway loops do not throw: they set a flag bit and
//    continue/stop in a defined way; the north star excludes them from parity and counts them.
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
        if (nh >= 2) flags |= WALK_VERTEX;                 // :520-532 (see header: treated as a crossing of the last hit edge)
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
