// ORACLE (test infrastructure only). Triangle mesh in the reference's conventions.
//
// The reference stores the surface as CGAL::Surface_mesh<point3> (src/utility/meshUtilities.h:25-29)
// and reads face corners with halfedges_around_face + source() (src/utility/meshUtilities.cpp:13-34).
// Here the mesh is plain arrays: corner[3f+k] is the k-th corner in that reference order
// (an OFF line "3 a b c" yields corners (c,a,b); the rotation is applied
// by the loader, not here).  "Edge k" of a face is the edge opposite corner k.
#pragma once
#include "vec3.hpp"
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace orc {

struct Mesh {
    int nV = 0, nF = 0;
    std::vector<V3> v;     // vertex positions
    std::vector<int> c;    // 3*nF corners (reference order)
    std::vector<int> adj;  // 3*nF: face across edge k (-1 = border)
    std::vector<int> adjk; // 3*nF: index of that same edge inside the neighbouring face
    std::vector<int> ringStart, ringFaces; // CSR vertex -> incident faces
    V3 bbmin{0, 0, 0}, bbmax{0, 0, 0};

    void set(int nV_, const double* xyz, int nF_, const int* corners)
    {
        nV = nV_;
        nF = nF_;
        v.resize(nV);
        for (int i = 0; i < nV; ++i) v[i] = V3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        c.assign(corners, corners + 3 * nF);
        buildAdjacency();
        buildSpan();
    }

    // triangulatedMeshSpace::updateMeshSpanAndTree (src/models/triangulatedMeshSpace.cpp:6-30):
    // the box is seeded with (0,0,0), not with the first vertex, so it always contains the origin.
    void buildSpan()
    {
        bbmin = V3{0, 0, 0};
        bbmax = V3{0, 0, 0};
        for (int i = 0; i < nV; ++i)
            for (int d = 0; d < 3; ++d) {
                if (v[i][d] < bbmin[d]) bbmin[d] = v[i][d];
                if (v[i][d] > bbmax[d]) bbmax[d] = v[i][d];
            }
    }

    void buildAdjacency()
    {
        adj.assign(3 * (size_t)nF, -1);
        adjk.assign(3 * (size_t)nF, -1);
        std::unordered_map<uint64_t, int> half; // directed edge (a->b) -> 3f+k
        half.reserve(3 * (size_t)nF * 2);
        auto key = [](int a, int b) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; };
        for (int f = 0; f < nF; ++f)
            for (int k = 0; k < 3; ++k) {
                int a = c[3 * f + (k + 1) % 3], b = c[3 * f + (k + 2) % 3];
                if (!half.emplace(key(a, b), 3 * f + k).second)
                    throw std::runtime_error("mesh: duplicated directed edge (non-manifold or inconsistently oriented)");
            }
        for (int f = 0; f < nF; ++f)
            for (int k = 0; k < 3; ++k) {
                int a = c[3 * f + (k + 1) % 3], b = c[3 * f + (k + 2) % 3];
                auto it = half.find(key(b, a));
                if (it != half.end()) {
                    adj[3 * f + k] = it->second / 3;
                    adjk[3 * f + k] = it->second % 3;
                }
            }
        ringStart.assign(nV + 1, 0);
        for (int i = 0; i < 3 * nF; ++i) ringStart[c[i] + 1]++;
        for (int i = 0; i < nV; ++i) ringStart[i + 1] += ringStart[i];
        ringFaces.resize(3 * (size_t)nF);
        std::vector<int> fill(ringStart.begin(), ringStart.end() - 1);
        for (int f = 0; f < nF; ++f)
            for (int k = 0; k < 3; ++k) ringFaces[fill[c[3 * f + k]]++] = f;
    }

    // PMP::construct_point / Surface_mesh_shortest_path::point:
    // (b0 p0 + b1 p1 + b2 p2) / (b0 + b1 + b2).  Call sites: triangulatedMeshSpace.cpp:89,451,514.
    V3 point(int f, const double b[3]) const
    {
        const V3& p0 = v[c[3 * f]];
        const V3& p1 = v[c[3 * f + 1]];
        const V3& p2 = v[c[3 * f + 2]];
        double s = b[0] + b[1] + b[2];
        return V3{(b[0] * p0.x + b[1] * p1.x + b[2] * p2.x) / s, (b[0] * p0.y + b[1] * p1.y + b[2] * p2.y) / s,
                  (b[0] * p0.z + b[1] * p1.z + b[2] * p2.z) / s};
    }

    // PMP::compute_face_normal: unit((p1-p0) x (p2-p0)) in corner order.
    V3 normal(int f) const
    {
        const V3& p0 = v[c[3 * f]];
        const V3& p1 = v[c[3 * f + 1]];
        const V3& p2 = v[c[3 * f + 2]];
        V3 n = cross(p1 - p0, p2 - p0);
        return n / norm(n);
    }

    double area() const // totalArea, src/utility/meshUtilities.cpp:80-85,103-117
    {
        double a = 0;
        for (int f = 0; f < nF; ++f) {
            const V3& p0 = v[c[3 * f]];
            a += norm(cross(v[c[3 * f + 1]] - p0, v[c[3 * f + 2]] - p0)) / 2.0;
        }
        return a;
    }
};

// PMP::barycentric_coordinates(p,q,r,x) (Ericson form); call sites
// src/models/triangulatedMeshSpace.cpp:478,606,655.
inline void ericsonBary(const V3& p, const V3& q, const V3& r, const V3& x, double out[3])
{
    V3 v0 = q - p, v1 = r - p, v2 = x - p;
    double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    double den = d00 * d11 - d01 * d01;
    double vv = (d11 * d20 - d01 * d21) / den;
    double ww = (d00 * d21 - d01 * d20) / den;
    out[0] = 1.0 - vv - ww;
    out[1] = vv;
    out[2] = ww;
}

} // namespace orc
