// ORACLE (test infrastructure only).  R^3 point -> mesh position, the import path of the reference:
// simpleModel::R3PositionsToMeshPositions (src/models/simpleModel.cpp:136-154) =
//   PMP::locate_with_AABB_tree(pos, tree, mesh)            closest point of the mesh, its face, barycentric weights
//   simpleModel::clampBarycentricCoordinatesToFace         (src/models/simpleModel.cpp:114-134, clampTolerance 1e-14,
//                                                           src/models/simpleModel.h:101)
//
// PARITY UNPINNED at the CGAL boundary: locate_with_AABB_tree lives in CGAL 5.6 (Polygon_mesh_processing/locate.h, not
// vendored).  Restated from its published algorithm: closest point of each triangle by Voronoi-region classification
// (Ericson, Real-Time Collision Detection 5.1.5 -- what Construct_projected_point_3 computes), the nearest face wins,
// barycentric weights of that point by the dot-product (Cramer) formula of CGAL::barycentric_coordinates, weights within
// machine epsilon of 0 or 1 snapped when a weight left [0,1] (internal::snap_coordinates_to_border).  Here the search is
// brute force over all faces; among faces at exactly the same squared distance the LOWEST FACE INDEX wins (CGAL's answer
// depends on its AABB traversal order; a point that close to an edge or vertex is equally well represented by either face).
// The clamp is the reference's own code, restated literally including its sequential renormalisation.
#pragma once
#include "mesh.hpp"
#include <limits>

namespace orc {

// closest point of triangle (a, b, c) to p
inline V3 closestPointOnTriangle(const V3& p, const V3& a, const V3& b, const V3& c)
{
    const V3 ab = b - a, ac = c - a, ap = p - a;
    const double d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0.0 && d2 <= 0.0) return a;
    const V3 bp = p - b;
    const double d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0.0 && d4 <= d3) return b;
    const double vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
        const double v = d1 / (d1 - d3);
        return a + v * ab;
    }
    const V3 cp = p - c;
    const double d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0.0 && d5 <= d6) return c;
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
        const double w = d2 / (d2 - d6);
        return a + w * ac;
    }
    const double va = d3 * d6 - d5 * d4;
    if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
        const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        return b + w * (c - b);
    }
    const double denom = 1.0 / (va + vb + vc);
    const double v = vb * denom, w = vc * denom;
    return (a + v * ab) + w * ac;
}

// barycentric weights of q in (p0, p1, p2), then the snap and the reference's clamp
inline void locateWeights(const V3& q, const V3& p0, const V3& p1, const V3& p2, double clampTol, double out[3])
{
    const V3 v0 = p1 - p0, v1 = p2 - p0, v2 = q - p0;
    const double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    const double den = d00 * d11 - d01 * d01;
    const double v = (d11 * d20 - d01 * d21) / den, w = (d00 * d21 - d01 * d20) / den;
    double co[3] = {(1.0 - v) - w, v, w};
    if (co[0] < 0.0 || co[0] > 1.0 || co[1] < 0.0 || co[1] > 1.0 || co[2] < 0.0 || co[2] > 1.0) {
        const double eps = std::numeric_limits<double>::epsilon();
        double residue = 0.0;
        for (int i = 0; i < 3; ++i) {
            if (std::fabs(co[i]) <= eps) residue = residue + co[i], co[i] = 0.0;
            else if (std::fabs(1.0 - co[i]) <= eps) residue = residue - (1.0 - co[i]), co[i] = 1.0;
        }
        for (int i = 0; i < 3; ++i)
            if (co[i] != 0.0 && co[i] != 1.0) {
                co[i] = co[i] + residue;
                break;
            }
    }
    // simpleModel::clampBarycentricCoordinatesToFace: weights smaller than the tolerance are set to it, then the three
    // divisions run one after the other, each seeing the weights already divided before it
    double w1 = co[0], w2 = co[1], w3 = co[2];
    if (std::fabs(w1) < clampTol) w1 = clampTol;
    if (std::fabs(w2) < clampTol) w2 = clampTol;
    if (std::fabs(w3) < clampTol) w3 = clampTol;
    w1 = w1 / ((w1 + w2) + w3);
    w2 = w2 / ((w1 + w2) + w3);
    w3 = w3 / ((w1 + w2) + w3);
    out[0] = w1, out[1] = w2, out[2] = w3;
}

inline void locatePoint(const Mesh& m, const V3& p, double clampTol, int& face, double bary[3])
{
    double best = std::numeric_limits<double>::infinity();
    int bf = -1;
    for (int f = 0; f < m.nF; ++f) {
        const V3 q = closestPointOnTriangle(p, m.v[m.c[3 * f]], m.v[m.c[3 * f + 1]], m.v[m.c[3 * f + 2]]);
        const double d2 = sqlen(p - q);
        if (d2 < best) best = d2, bf = f; // ascending f: the lowest index keeps an exact tie
    }
    face = bf;
    const V3 &p0 = m.v[m.c[3 * bf]], &p1 = m.v[m.c[3 * bf + 1]], &p2 = m.v[m.c[3 * bf + 2]];
    locateWeights(closestPointOnTriangle(p, p0, p1, p2), p0, p1, p2, clampTol, bary);
}

} // namespace orc
