// ORACLE (test infrastructure only). Per-source local patch ("submesh") face set.
// Literal restatement of submesher::constructSubmeshFromSourceAndTargets
// (src/utility/submesher.cpp:55-147): explicit stack, same visiting rules, same early exits.
// The result is returned as a face list; the reference additionally re-indexes it into a fresh
// Surface_mesh (submesher.cpp:4-53), which does not change the metric.
#pragma once
#include "mesh.hpp"
#include <unordered_set>
#include <vector>

namespace orc {

inline std::vector<int> patchFaces(const Mesh& m, int srcFace, const double srcBary[3], const std::vector<int>& targetFaces,
                                   double maxDistFromSource)
{
    double thr2 = maxDistFromSource * maxDistFromSource; // submesher.cpp:62
    V3 sp = m.point(srcFace, srcBary);                   // :64
    std::unordered_set<int> goal, visited;
    std::vector<int> order; // insertion order, for a deterministic return value
    auto visit = [&](int f) {
        if (visited.insert(f).second) order.push_back(f);
    };
    visit(srcFace);                                      // :72
    for (int tf : targetFaces)
        if (tf != srcFace) goal.insert(tf);              // :76-78
    if (goal.empty()) return order;                      // :79-80
    std::vector<int> stack;
    for (int k = 0; k < 3; ++k) {                        // :86-96 (all three neighbours, unconditionally)
        int g = m.adj[3 * srcFace + k];
        if (g < 0) continue;
        visit(g);
        stack.push_back(g);
        goal.erase(g);
    }
    if (goal.empty()) return order;                      // :97-98
    while (!stack.empty()) {                             // :107-137
        int cur = stack.back();
        stack.pop_back();
        for (int k = 0; k < 3; ++k) {
            int g = m.adj[3 * cur + k];
            if (g < 0) continue;
            if (visited.count(g)) continue;
            const V3& a = m.v[m.c[3 * g]];
            const V3& b = m.v[m.c[3 * g + 1]];
            const V3& c = m.v[m.c[3 * g + 2]];
            if (sqlen(sp - a) > thr2 && sqlen(sp - b) > thr2 && sqlen(sp - c) > thr2) continue; // :123-128
            visit(g);
            goal.erase(g);
            stack.push_back(g);
        }
    }
    std::vector<int> rest(goal.begin(), goal.end());     // :143-144 leftover goal faces
    std::sort(rest.begin(), rest.end());
    for (int g : rest) visit(g);
    return order;
}

} // namespace orc
