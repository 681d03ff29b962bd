// ORACLE (test infrastructure only). Euclidean cell list over the ambient 3-D embedding.
// Restates src/utility/hyperRectangularCellList.cpp:9-46 (grid), :71-80 (bin), :82-125 (sort),
// :128-159 (stencil), src/utility/indexer.h:47-50 (linearisation) and
// src/utility/cellListNeighborStructure.cpp:45-84 (ordered candidate list + largest distance).
#pragma once
#include "vec3.hpp"
#include <algorithm>
#include <cmath>
#include <vector>

namespace orc {

struct CellList {
    double mn[3], mx[3];
    double range = -1;
    int n[3] = {1, 1, 1};
    double cs[3] = {1, 1, 1};
    std::vector<int> cellStart, cellItems; // CSR; items ascending in particle index inside each cell
    const std::vector<V3>* pts = nullptr;

    void setDomain(const V3& lo, const V3& hi)
    {
        for (int d = 0; d < 3; ++d) {
            mn[d] = lo[d];
            mx[d] = hi[d];
        }
        range = -1;
    }
    // hyperRectangularCellList::setGridSize: n_d = max(1, floor(extent_d / range)), cell_d = extent_d / n_d
    void setRange(double r)
    {
        if (r == range) return;
        range = r;
        for (int d = 0; d < 3; ++d) {
            double ext = mx[d] - mn[d];
            n[d] = std::max(1, (int)std::floor(ext / r));
            cs[d] = ext / n[d];
        }
    }
    int cellCoord(double x, int d) const
    {
        return std::max(0, std::min(n[d] - 1, (int)std::floor((x - mn[d]) / cs[d])));
    }
    int cellOf(const V3& p) const
    {
        int ix = cellCoord(p.x, 0), iy = cellCoord(p.y, 1), iz = cellCoord(p.z, 2);
        return ix + iy * n[0] + iz * n[0] * n[1]; // Index3D, indexer.h:47-50
    }
    // hyperRectangularCellList::sort: particles are appended to their cell in ascending index order.
    void build(const std::vector<V3>& p)
    {
        pts = &p;
        int total = n[0] * n[1] * n[2];
        cellStart.assign(total + 1, 0);
        std::vector<int> cid(p.size());
        for (size_t i = 0; i < p.size(); ++i) {
            cid[i] = cellOf(p[i]);
            cellStart[cid[i] + 1]++;
        }
        for (int i = 0; i < total; ++i) cellStart[i + 1] += cellStart[i];
        cellItems.resize(p.size());
        std::vector<int> fill(cellStart.begin(), cellStart.end() - 1);
        for (size_t i = 0; i < p.size(); ++i) cellItems[fill[cid[i]]++] = (int)i;
    }
    // cellListNeighborStructure::constructCandidateNeighborList: stencil xx outer / yy / zz inner,
    // candidate kept iff idx != self and dist^2 < range^2 (strict); returns sqrt(max dist^2).
    double candidates(int self, std::vector<int>& out) const
    {
        out.clear();
        const V3& p = (*pts)[self];
        int ix = cellCoord(p.x, 0), iy = cellCoord(p.y, 1), iz = cellCoord(p.z, 2);
        double r2 = range * range, maxd2 = 0;
        for (int xx = std::max(0, ix - 1); xx <= std::min(n[0] - 1, ix + 1); ++xx)
            for (int yy = std::max(0, iy - 1); yy <= std::min(n[1] - 1, iy + 1); ++yy)
                for (int zz = std::max(0, iz - 1); zz <= std::min(n[2] - 1, iz + 1); ++zz) {
                    int cidx = xx + yy * n[0] + zz * n[0] * n[1];
                    for (int s = cellStart[cidx]; s < cellStart[cidx + 1]; ++s) {
                        int j = cellItems[s];
                        if (j == self) continue;
                        const V3& q = (*pts)[j];
                        double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
                        double d2 = dx * dx + dy * dy + dz * dz; // CGAL::squared_distance
                        if (d2 < r2) {
                            out.push_back(j);
                            if (d2 > maxd2) maxd2 = d2;
                        }
                    }
                }
        return std::sqrt(maxd2);
    }
};

} // namespace orc
