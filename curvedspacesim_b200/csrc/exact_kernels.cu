// Kernels whose branches / indices must be bit-identical to the CPU oracle.  THIS FILE IS COMPILED WITH
// -fmad=false (no FMA contraction); division and sqrt are IEEE-rounded in fp64 on the GPU.
//
//   k_euclid_cell   meshPositionToEuclideanLocation (triangulatedMeshSpace.cpp:82-106) fused with the
//                   cell binning of hyperRectangularCellList::positionToCellIndex/sort (:71-125)
//   k_cell_place / k_cell_fill / k_cell_rank   scan-free, deterministic (ascending particle index) cell contents
//   k_walk          triangulatedMeshSpace::transportParticleAndVectors (:448-658), one thread per particle,
//                   optionally fused with the velocity-Verlet first half step (velocityVerletNVE.cpp:14-21)
//   k_axpy-type     updater arithmetic (src/updaters/*.cpp) and reductions
//   k_locate        simpleModel::R3PositionsToMeshPositions (simpleModel.cpp:136-154): closest face + clamped weights
#include "common.cuh"
#include "kernels.h"

namespace css {

__device__ __forceinline__ d3 operator+(const d3& a, const d3& b) { return d3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ d3 operator-(const d3& a, const d3& b) { return d3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ d3 operator*(double s, const d3& a) { return d3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ d3 operator/(const d3& a, double s) { return d3{a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ double dot(const d3& a, const d3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ d3 cross(const d3& a, const d3& b)
{
    return d3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double sqlen(const d3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ double norm(const d3& a) { return sqrt(sqlen(a)); }

__device__ __forceinline__ d3 ld3(const double* p, int i) { return d3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
__device__ __forceinline__ void st3(double* p, int i, const d3& v)
{
    p[3 * i] = v.x;
    p[3 * i + 1] = v.y;
    p[3 * i + 2] = v.z;
}

struct Tri {
    d3 p0, p1, p2;
};
__device__ __forceinline__ Tri ldtri(const MeshDev& m, int f)
{
    int4 c = __ldg(m.corner + f);
    return Tri{ldvert(m, c.x), ldvert(m, c.y), ldvert(m, c.z)};
}
__device__ __forceinline__ d3 tpoint(const Tri& t, const double b[3])
{
    double s = b[0] + b[1] + b[2];
    return d3{(b[0] * t.p0.x + b[1] * t.p1.x + b[2] * t.p2.x) / s, (b[0] * t.p0.y + b[1] * t.p1.y + b[2] * t.p2.y) / s,
              (b[0] * t.p0.z + b[1] * t.p1.z + b[2] * t.p2.z) / s};
}
__device__ __forceinline__ d3 tnormal(const Tri& t)
{
    d3 n = cross(t.p1 - t.p0, t.p2 - t.p0);
    return n / norm(n);
}
__device__ __forceinline__ void ericson(const Tri& t, const d3& x, double out[3])
{
    d3 v0 = t.p1 - t.p0, v1 = t.p2 - t.p0, v2 = x - t.p0;
    double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    double den = d00 * d11 - d01 * d01;
    double vv = (d11 * d20 - d01 * d21) / den;
    double ww = (d00 * d21 - d01 * d20) / den;
    out[0] = 1.0 - vv - ww;
    out[1] = vv;
    out[2] = ww;
}
__device__ __forceinline__ void belowZeroClamp(double b[3])
{
    const double tol = 1e-11;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        b[i] = (b[i] < tol) ? tol : b[i];
        s += b[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) b[i] = b[i] / s;
}
__device__ __forceinline__ void nearZeroClamp(double b[3])
{
    const double tol = 1e-13;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (b[i] > -tol && b[i] < tol) b[i] = tol;
        s += b[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) b[i] = b[i] / s;
}
__device__ __forceinline__ d3 rotateAboutAxis(const d3& p, const d3& base, const d3& a, double s, double c)
{
    d3 tip = base + a;
    d3 ax = tip - base;
    double an = sqlen(ax);
    ax = ax / sqrt(an);
    d3 sh = p - base;
    double dp = ax.x * sh.x + ax.y * sh.y + ax.z * sh.z;
    d3 r;
    r.x = ax.x * dp * (1. - c) + sh.x * c + (ax.y * sh.z - ax.z * sh.y) * s;
    r.y = ax.y * dp * (1. - c) + sh.y * c + (ax.z * sh.x - ax.x * sh.z) * s;
    r.z = ax.z * dp * (1. - c) + sh.z * c + (ax.x * sh.y - ax.y * sh.x) * s;
    return r + base;
}

// ------------------------------------------------------------------------------------------------
// coarse (optional, sharded runs only): occupancy of the 2x2x2 blocks of cells, for the replicated stride bound of k_cell_rank
__global__ void k_euclid_cell(MeshDev m, CellGrid g, int n, const int* __restrict__ face, const double* __restrict__ bary,
                              double* __restrict__ eucl, int* __restrict__ cellOf, int* __restrict__ cellCount, int* __restrict__ cellSlot,
                              int* __restrict__ coarse)
{
    PDL_ENTRY();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Tri t = ldtri(m, face[i]);
    double b[3] = {bary[3 * i], bary[3 * i + 1], bary[3 * i + 2]};
    d3 p = tpoint(t, b);
    st3(eucl, i, p);
    if (cellOf) {
        int ix = cellCoord(g, p.x, 0), iy = cellCoord(g, p.y, 1), iz = cellCoord(g, p.z, 2);
        int c = ix + iy * g.n[0] + iz * g.n[0] * g.n[1];
        cellOf[i] = c;
        cellSlot[i] = atomicAdd(cellCount + c, 1); // arrival order inside the cell; the first arrival places the cell (k_cell_place)
        if (coarse) atomicAdd(coarse + (ix >> 1) + (iy >> 1) * ((g.n[0] + 1) >> 1) + (iz >> 1) * ((g.n[0] + 1) >> 1) * ((g.n[1] + 1) >> 1), 1);
    }
}

// Cell contents without a scan over the grid (the grid of config 5 has 2.5 M cells for 100 k particles; a scan costs more than
// everything else in the cell list).  The particle that arrived first in a cell reserves the cell's range with one atomic bump
// (cellStart[c]; ranges are contiguous and disjoint but in no particular order, empty cells keep stale starts and are never
// read because their count is 0), every particle drops itself at start + arrival slot, and the members are then put in
// ascending particle order, as hyperRectangularCellList::sort produces by inserting particles in index order (:96-112).
// Consumers read a cell as items[cellStart[c] .. cellStart[c] + cellCount[c]).
__global__ void k_cell_place(int n, const int* __restrict__ cellOf, const int* __restrict__ cellSlot, const int* __restrict__ cellCount,
                             int* __restrict__ cellStart, int* __restrict__ bump)
{
    PDL_ENTRY();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const bool first = i < n && cellSlot[i] == 0;
    const int c = first ? cellOf[i] : 0;
    const int cnt = first ? cellCount[c] : 0;
    int incl = cnt; // one bump per warp: the ranges of the warp's cells are carved out of it by a prefix sum
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(bump, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (first) cellStart[c] = base + incl - cnt;
}
__global__ void k_cell_fill(int n, const int* __restrict__ cellOf, const int* __restrict__ cellSlot, const int* __restrict__ cellStart,
                            int* __restrict__ tmpItems)
{
    PDL_ENTRY();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tmpItems[cellStart[cellOf[i]] + cellSlot[i]] = i;
}
// rank of particle i inside its cell = number of members with a smaller index
// ... and, in sharded runs (coarse != nullptr), the replicated half of the stride guard (common.cuh): the particle's 3x3x3 cell
// stencil lies inside 2x2x2 coarse blocks, whose occupancy therefore bounds its candidate count.  Every rank evaluates the
// bound for ALL particles from the replicated positions, so every rank raises the guard in the same step without talking to
// the others.  (On one rank the exact candidate count found by stage 1 raises it, at no cost.)
__global__ void k_cell_rank(int n, const int* __restrict__ cellOf, const int* __restrict__ cellStart, const int* __restrict__ cellCount,
                            const int* __restrict__ tmpItems, int* __restrict__ items, const int* __restrict__ coarse, int nx, int ny, int nz,
                            int kmax, unsigned long long* __restrict__ counters)
{
    PDL_ENTRY();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int occ = 0;
    if (i < n) {
        int c = cellOf[i];
        int s0 = cellStart[c], cnt = cellCount[c];
        int r = 0;
        for (int s = 0; s < cnt; ++s) r += (tmpItems[s0 + s] < i);
        items[s0 + r] = i;
        if (coarse) {
            const int ix = c % nx, iy = (c / nx) % ny, iz = c / (nx * ny);
            const int cx = (nx + 1) >> 1, cy = (ny + 1) >> 1;
            const int x0 = max(0, ix - 1) >> 1, x1 = min(nx - 1, ix + 1) >> 1, y0 = max(0, iy - 1) >> 1, y1 = min(ny - 1, iy + 1) >> 1;
            const int z0 = max(0, iz - 1) >> 1, z1 = min(nz - 1, iz + 1) >> 1;
            for (int zz = z0; zz <= z1; ++zz)
                for (int yy = y0; yy <= y1; ++yy)
                    for (int xx = x0; xx <= x1; ++xx) occ += coarse[xx + yy * cx + zz * cx * cy];
            occ -= 1; // the particle itself
        }
    }
    if (!coarse) return;
    occ = __reduce_max_sync(0xffffffffu, occ);
    if ((threadIdx.x & 31) == 0 && occ > kmax) {
        atomicMax(counters + C_KMAX_NEED, (unsigned long long)occ);
        atomicAdd(counters + C_KMAX_OVERFLOW, 1ull);
    }
}

// ------------------------------------------------------------------------------------------------
// The walker.  One thread per particle.  mode bit0: displacement computed in-kernel as the velocity-
// Verlet first half step (disp = dt v + dt^2/2 f ; v += dt/2 f); otherwise disp is read from `disp`.
__device__ __forceinline__ int edgeHits(const double S[3], const double E[3], int last, int& firstK, int& lastK, double I[3])
{
    int nh = 0;
    if (last != 2) {
        double den = S[0] + S[1] - E[0] - E[1];
        if (den != 0) {
            double t1 = (-E[1] + E[1] * S[0] + S[1] - E[0] * S[1]) / den;
            double t2 = (-1 + S[0] + S[1]) / den;
            if (t1 >= 0 && t1 <= 1 && t2 >= 0 && t2 <= 1) {
                I[0] = S[0] + t2 * (E[0] - S[0]);
                I[1] = S[1] + t2 * (E[1] - S[1]);
                I[2] = S[2] + t2 * (E[2] - S[2]);
                if (!nh) firstK = 2;
                lastK = 2;
                nh++;
            }
        }
    }
    if (last != 0) {
        double den = -E[0] + S[0];
        if (den != 0) {
            double t1 = -(E[0] - S[0] + E[1] * S[0] - E[0] * S[1]) / den;
            double t2 = (S[0]) / den;
            if (t1 >= 0 && t1 <= 1 && t2 >= 0 && t2 <= 1) {
                I[0] = S[0] + t2 * (E[0] - S[0]);
                I[1] = S[1] + t2 * (E[1] - S[1]);
                I[2] = S[2] + t2 * (E[2] - S[2]);
                if (!nh) firstK = 0;
                lastK = 0;
                nh++;
            }
        }
    }
    if (last != 1) {
        double den = S[1] - E[1];
        if (den != 0) {
            double t1 = (E[0] * S[1] - E[1] * S[0]) / den;
            double t2 = S[1] / den;
            if (t1 >= 0 && t1 <= 1 && t2 >= 0 && t2 <= 1) {
                I[0] = S[0] + t2 * (E[0] - S[0]);
                I[1] = S[1] + t2 * (E[1] - S[1]);
                I[2] = S[2] + t2 * (E[2] - S[2]);
                if (!nh) firstK = 1;
                lastK = 1;
                nh++;
            }
        }
    }
    return nh;
}

// triangulatedMeshSpace::projectVectorsIfOverBoundary (triangulatedMeshSpace.cpp:411-426)
template <int NT> __device__ __forceinline__ void projectVectorsIfOverBoundary(d3 (&T)[NT > 0 ? NT : 1], int nT, const d3& orthogonal, const d3& inward)
{
    bool pointsOut = dot(orthogonal, inward) < 0;
#pragma unroll
    for (int i = 0; i < NT; ++i)
        if (i < nT) {
            bool along = dot(T[i], orthogonal) > 0;
            if ((pointsOut && along) || (!pointsOut && !along)) {
                d3 dhat = orthogonal / norm(orthogonal);
                T[i] = T[i] - dot(T[i], dhat) * dhat;
            }
        }
}


// ---- vertex events (rare; kept out of line).  The same operation sequence as the CPU checker used by the tests: throughVertex /
// updateForVertexIntersection (triangulatedMeshSpace.cpp:247-407, :566-611) with their intended semantics, and the boundary
// vertex rule of the open spaces (openMeshSpace.cpp:3-70, absorbing/tangentialOpenMeshSpace::updateAtBoundaryVertex).
// Angles use deterministic IEEE-only trigonometry so that host oracle and device agree bit for bit.
__device__ __noinline__ double detAngle(double s, double c)
{
    const double PI = 3.14159265358979323846;
    double r = sqrt(s * s + c * c);
    if (!(r > 0)) return 0.0;
    bool obtuse = c < 0;
    double ca = obtuse ? -c : c;
    double t = s / (r + ca);
    t = t / (1.0 + sqrt(1.0 + t * t));
    t = t / (1.0 + sqrt(1.0 + t * t));
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
__device__ __noinline__ void detSinCos(double x, double& s, double& c)
{
    double y = x * 0.125, y2 = y * y;
    double ps = 1.0 - y2 / 272.0;
    ps = 1.0 - y2 / 210.0 * ps;
    ps = 1.0 - y2 / 156.0 * ps;
    ps = 1.0 - y2 / 110.0 * ps;
    ps = 1.0 - y2 / 72.0 * ps;
    ps = 1.0 - y2 / 42.0 * ps;
    ps = 1.0 - y2 / 20.0 * ps;
    ps = 1.0 - y2 / 6.0 * ps;
    double pc = 1.0 - y2 / 306.0;
    pc = 1.0 - y2 / 240.0 * pc;
    pc = 1.0 - y2 / 182.0 * pc;
    pc = 1.0 - y2 / 132.0 * pc;
    pc = 1.0 - y2 / 90.0 * pc;
    pc = 1.0 - y2 / 56.0 * pc;
    pc = 1.0 - y2 / 30.0 * pc;
    pc = 1.0 - y2 / 12.0 * pc;
    pc = 1.0 - y2 / 2.0 * pc;
    s = y * ps, c = pc;
    for (int i = 0; i < 3; ++i) {
        double s2 = 2.0 * s * c, c2 = 1.0 - 2.0 * s * s;
        s = s2, c = c2;
    }
}
__device__ __forceinline__ double angleBetweenUnit(const d3& u, const d3& v) { return detAngle(norm(cross(u, v)), dot(u, v)); }
__device__ __forceinline__ d3 unitTo(const d3& from, const d3& to)
{
    d3 d = to - from;
    return d / norm(d);
}
#define CSS_WALK_MAX_VALENCE 64
struct FanFace {
    int g, kc; // the vertex sits at corner kc of face g
};
__device__ __forceinline__ int pick3(const int4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
__device__ __forceinline__ FanFace fanClockwise(const MeshDev& m, FanFace a)
{
    const int e = (a.kc + 2) % 3;
    const int4 ad = __ldg(m.adj + a.g);
    const int g = pick3(ad, e);
    return FanFace{g, g < 0 ? 0 : (((ad.w >> (2 * e)) & 3) + 2) % 3};
}
__device__ __forceinline__ FanFace fanCounterClockwise(const MeshDev& m, FanFace a)
{
    const int e = (a.kc + 1) % 3;
    const int4 ad = __ldg(m.adj + a.g);
    const int g = pick3(ad, e);
    return FanFace{g, g < 0 ? 0 : (((ad.w >> (2 * e)) & 3) + 1) % 3};
}
__device__ __forceinline__ double fanAngle(const MeshDev& m, FanFace a, const d3& Pv, d3& eNext, d3& ePrev)
{
    const int4 c = __ldg(m.corner + a.g);
    eNext = unitTo(Pv, ldvert(m, pick3(c, (a.kc + 1) % 3)));
    ePrev = unitTo(Pv, ldvert(m, pick3(c, (a.kc + 2) % 3)));
    return angleBetweenUnit(eNext, ePrev);
}
__device__ __noinline__ bool throughVertex(const MeshDev& m, int f, int kv, const d3& travel, int& gOut, d3& heading)
{
    const d3 Pv = ldvert(m, pick3(__ldg(m.corner + f), kv));
    const FanFace src{f, kv};
    d3 eN, eP;
    double total = 0;
    int n = 0;
    for (FanFace a = fanClockwise(m, src);; a = fanClockwise(m, a)) {
        if (a.g < 0 || ++n > CSS_WALK_MAX_VALENCE) return false;
        total += fanAngle(m, a, Pv, eN, eP);
        if (a.g == f) break;
    }
    const double half = total / 2.0;
    d3 back = d3{0, 0, 0} - travel;
    fanAngle(m, src, Pv, eN, eP);
    double traveled = angleBetweenUnit(back, eN);
    FanFace land = src;
    if (traveled < half) {
        for (FanFace a = fanClockwise(m, src);; a = fanClockwise(m, a)) {
            traveled += fanAngle(m, a, Pv, eN, eP);
            land = a;
            if (traveled >= half || a.g == f) break;
        }
    }
    double delta = traveled - half, sn, cs;
    if (delta < 0) delta = 0;
    detSinCos(delta, sn, cs);
    d3 u = cross(eN, eP);
    u = u / norm(u);
    d3 w = cross(u, eN);
    heading = cs * eN + sn * w;
    heading = heading / norm(heading);
    gOut = land.g;
    return true;
}
__device__ __noinline__ bool boundaryVertexHeading(const MeshDev& m, int f, int kv, const d3& dhat, int& gOut, int& wOut, d3& heading)
{
    const d3 Pv = ldvert(m, pick3(__ldg(m.corner + f), kv));
    const FanFace src{f, kv};
    double best = -1;
    bool found = false;
    for (int dir = 0; dir < 2; ++dir) {
        int n = 0;
        for (FanFace a = dir ? fanCounterClockwise(m, src) : fanClockwise(m, src); a.g >= 0 && a.g != f && ++n <= CSS_WALK_MAX_VALENCE;
             a = dir ? fanCounterClockwise(m, a) : fanClockwise(m, a)) {
            const int4 c = __ldg(m.corner + a.g);
            for (int j = 1; j <= 2; ++j) {
                int w = pick3(c, (a.kc + j) % 3);
                d3 out = unitTo(Pv, ldvert(m, w));
                double overlap = dot(out, dhat);
                if (overlap > best) best = overlap, heading = out, gOut = a.g, wOut = w, found = true;
            }
        }
    }
    return found;
}

// the path goes through the vertex at corner kv of face f.  Returns true when the walk goes on (state updated), false when it
// stops at S (closed space at a boundary vertex, absorbing space, nothing left to slide along)
template <int NT>
__device__ __noinline__ bool vertexEvent(const MeshDev& m, int& f, Tri& tri, int kv, double S[3], d3& p, d3& q, d3& disp, d3& n,
                                         d3 (&T)[NT > 0 ? NT : 1], int nT, int& flags)
{
    flags |= WALK_VERTEX;
    const d3 travel = disp / norm(disp);
    int g2 = -1;
    d3 heading;
    if (throughVertex(m, f, kv, travel, g2, heading)) {
        Tri tri2 = ldtri(m, g2);
        d3 n2 = tnormal(tri2);
        double rem = norm(q - p);
        d3 side = cross(n, travel), side2 = cross(n2, heading);
#pragma unroll
        for (int i = 0; i < NT; ++i)
            if (i < nT) {
                double ta = dot(T[i], travel), tb = dot(T[i], side), tc = dot(T[i], n);
                T[i] = ta * heading + tb * side2 + tc * n2;
            }
        disp = rem * heading;
        q = p + disp;
        f = g2;
        tri = tri2;
        ericson(tri, p, S);
        n = n2;
        return true;
    }
    flags |= WALK_BORDER;
    int w = -1;
    if (m.boundary == 0 || !boundaryVertexHeading(m, f, kv, travel, g2, w, heading)) return false;
    const int vIdx = pick3(__ldg(m.corner + f), kv);
    const d3 Pv = ldvert(m, vIdx);
    d3 rest = q - p;
    f = g2;
    tri = ldtri(m, f);
    n = tnormal(tri);
    ericson(tri, Pv, S);
    if (nT > 0) {
        d3 orth = cross(n, heading);
        orth = orth / norm(orth);
        const int4 c = __ldg(m.corner + f);
        d3 inside = Pv;
        for (int j = 0; j < 3; ++j) {
            int cv = pick3(c, j);
            if (cv != vIdx && cv != w) inside = ldvert(m, cv);
        }
        projectVectorsIfOverBoundary<NT>(T, nT, orth, inside - Pv);
    }
    if (m.boundary == 1) return false;
    double slide = dot(rest, heading);
    if (!(slide > 1e-9 * norm(ldvert(m, w) - Pv))) return false;
    disp = slide * heading;
    p = tpoint(tri, S);
    q = p + disp;
    return true;
}

template <int NT>
__device__ __forceinline__ int walkOne(const MeshDev& m, int& face, double bary[3], d3& disp, d3 (&T)[NT > 0 ? NT : 1], int nT,
                                       int& nCross)
{
    int flags = 0;
    int f = face;
    double S[3] = {bary[0], bary[1], bary[2]};
    Tri tri = ldtri(m, f);
    d3 p = tpoint(tri, S);
    d3 n = tnormal(tri);
    double nd = dot(n, disp);
    if (fabs(nd) > 1e-14) disp = disp - nd * n;
    d3 q = p + disp;
    int last = -1;
    double E[3];
    nCross = 0;
    for (;;) {
        ericson(tri, q, E);
        nearZeroClamp(E);
        q = tpoint(tri, E);
        belowZeroClamp(S);
        p = tpoint(tri, S);
        disp = q - p;
        if (E[0] != E[0]) {
            flags |= WALK_NAN;
            break;
        }
        if (!(E[0] < 0 || E[1] < 0 || E[2] < 0)) break;
        if (nCross >= CSS_WALK_MAX_CROSSINGS) {
            flags |= WALK_ITERCAP;
            belowZeroClamp(E);
            break;
        }
        int k0 = 0, k1 = 0;
        double I[3];
        int nh = edgeHits(S, E, last, k0, k1, I);
        if (nh == 0) {
            flags |= WALK_NOHIT;
            belowZeroClamp(E);
            break;
        }
        belowZeroClamp(I);
        d3 x = tpoint(tri, I);
        S[0] = I[0], S[1] = I[1], S[2] = I[2];
        p = x;
        int k = nh >= 2 ? k1 : k0;
        if (nh >= 2 && k0 != k1) { // two edges hit: the path goes through their common vertex (corner 3 - k0 - k1)
            last = -1;
            if (vertexEvent<NT>(m, f, tri, 3 - k0 - k1, S, p, q, disp, n, T, nT, flags)) {
                nCross++;
                continue;
            }
            E[0] = S[0], E[1] = S[1], E[2] = S[2];
            break;
        }
        int4 a = __ldg(m.adj + f);
        int g = k == 0 ? a.x : (k == 1 ? a.y : a.z);
        if (g < 0) { // border edge: openMeshSpace.cpp:114-238 and its absorbing / tangential subclasses
            flags |= WALK_BORDER;
            if (m.boundary == 0) {
                E[0] = S[0], E[1] = S[1], E[2] = S[2];
                break;
            }
            const d3 ev1 = k == 0 ? tri.p1 : (k == 1 ? tri.p2 : tri.p0), ev2 = k == 0 ? tri.p2 : (k == 1 ? tri.p0 : tri.p1);
            const d3 inner = k == 0 ? tri.p0 : (k == 1 ? tri.p1 : tri.p2);
            d3 edge = ev2 - ev1;
            d3 orth = cross(n, edge);
            orth = orth / norm(orth);
            d3 inward = inner - p;
            if (m.boundary == 1) { // absorbing: stop on the edge
                if (nT > 0) projectVectorsIfOverBoundary<NT>(T, nT, orth, inward);
                E[0] = S[0], E[1] = S[1], E[2] = S[2];
                break;
            }
            d3 fwd = edge / norm(edge); // tangential: slide along the edge
            d3 back = ev1 - ev2;
            d3 bwd = back / norm(back);
            double len = norm(disp);
            d3 dhat = disp / len;
            double fd = dot(dhat, fwd), bd = dot(dhat, bwd);
            projectVectorsIfOverBoundary<NT>(T, nT, orth, inward);
            double slide = (fd > bd ? fd : bd) * len;
            if (!(slide > 1e-9 * norm(edge))) { // nothing left to slide, or a slide at the scale of the 1e-11 source clamp (limit cycle in sheet corners)
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
        Tri tri2 = ldtri(m, g);
        d3 n2 = tnormal(tri2);
        double c = dot(n, n2);
        d3 ax = cross(n, n2);
        double an = norm(ax);
        if (c < 1 && an > 0) {
            ax = ax / an;
            q = rotateAboutAxis(q, p, ax, an, c);
            disp = q - p;
#pragma unroll
            for (int i = 0; i < NT; ++i)
                if (i < nT) {
                    d3 tt = p + T[i];
                    tt = rotateAboutAxis(tt, p, ax, an, c);
                    T[i] = tt - p;
                }
        }
        last = (a.w >> (2 * k)) & 3;
        f = g;
        tri = tri2;
        ericson(tri, p, S);
        n = n2;
        nCross++;
    }
    face = f;
    bary[0] = E[0], bary[1] = E[1], bary[2] = E[2];
    return flags;
}

// particles [0,n) of this rank; global index = minIdx + i in face/bary
#ifndef CSS_WALK_MINB
#define CSS_WALK_MINB 12
#endif
__global__ void __launch_bounds__(64, CSS_WALK_MINB) k_walk(MeshDev m, int n, int minIdx, int* __restrict__ face, double* __restrict__ bary, double* __restrict__ disp,
                       double* __restrict__ vel, double* __restrict__ frc, int transportForce, int transportVelocity, int mode,
                       double dt, int* __restrict__ flagsOut, unsigned long long* __restrict__ counters, PeerWin pw)
{
    PDL_ENTRY();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int fl = 0, nc = 0;
    if (strideGuardUp(counters)) return; // a step upstream is waiting for a larger neighbour stride: nothing moves (common.cuh)
    if (i == 0) atomicAdd(counters + C_STEP_GUARD, 1ull);
    if (i < n) {
        d3 d;
        if (mode & 1) { // velocityVerletNVE::velocityVerletFirstHalfStep
            d3 v = ld3(vel, i), f = ld3(frc, i);
            d = dt * v + (0.5 * dt * dt) * f;
            v = v + (0.5 * dt) * f;
            st3(vel, i, v);
        } else
            d = ld3(disp, i);
        d3 T[2];
        int nT = 0;
        if (transportForce) T[nT++] = ld3(frc, i);
        if (transportVelocity) T[nT++] = ld3(vel, i);
        int gi = minIdx + i;
        int f = face[gi];
        double b[3] = {bary[3 * gi], bary[3 * gi + 1], bary[3 * gi + 2]};
        fl = walkOne<2>(m, f, b, d, T, nT, nc);
        face[gi] = f;
        bary[3 * gi] = b[0], bary[3 * gi + 1] = b[1], bary[3 * gi + 2] = b[2];
        if (pw.n > 1) { // the exchange step of the multi-rank model, fused: store the new position into every peer's window
            for (int r = 0; r < pw.n; ++r)
                if (r != pw.rank) {
                    pw.face[r][gi] = f;
                    double* pb = pw.bary[r] + 3 * (size_t)gi;
                    pb[0] = b[0], pb[1] = b[1], pb[2] = b[2];
                }
            __threadfence_system();
        }
        st3(disp, i, d);
        nT = 0;
        if (transportForce) st3(frc, i, T[nT++]);
        if (transportVelocity) st3(vel, i, T[nT++]);
        if (flagsOut) flagsOut[i] = fl;
    }
    // counters: warp-aggregated
    unsigned mask = 0xffffffffu;
    for (int b = 0; b < 5; ++b) {
        unsigned bal = __ballot_sync(mask, (fl >> b) & 1);
        if ((threadIdx.x & 31) == 0 && bal) atomicAdd(counters + b, (unsigned long long)__popc(bal));
    }
    int s = nc;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(mask, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(counters + C_CROSSINGS, (unsigned long long)s);
}

// generic per-call transport (css_transport): arbitrary number of vectors handled NV at a time
__global__ void k_transport_generic(MeshDev m, int n, int* __restrict__ face, double* __restrict__ bary, double* __restrict__ disp,
                                    int nVec, double* __restrict__ vecs, int* __restrict__ flagsOut)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int f0 = face[i];
    double b0[3] = {bary[3 * i], bary[3 * i + 1], bary[3 * i + 2]};
    d3 d0 = ld3(disp, i);
    int fl = 0, nc;
    int f = f0;
    double b[3];
    d3 d = d0;
    if (nVec == 0) {
        d3 T[1];
        b[0] = b0[0], b[1] = b0[1], b[2] = b0[2];
        fl = walkOne<0>(m, f, b, d, T, 0, nc);
    }
    for (int v0 = 0; v0 < nVec; v0 += 2) { // each pass re-walks the same (deterministic) path carrying two vectors
        d3 T[2];
        int nT = min(2, nVec - v0);
        for (int j = 0; j < nT; ++j) T[j] = ld3(vecs, i * nVec + v0 + j);
        f = f0;
        b[0] = b0[0], b[1] = b0[1], b[2] = b0[2];
        d = d0;
        fl = walkOne<2>(m, f, b, d, T, nT, nc);
        for (int j = 0; j < nT; ++j) st3(vecs, i * nVec + v0 + j, T[j]);
    }
    face[i] = f;
    bary[3 * i] = b[0], bary[3 * i + 1] = b[1], bary[3 * i + 2] = b[2];
    st3(disp, i, d);
    if (flagsOut) flagsOut[i] = fl;
}

// ---------------------------------------------------------------------------------- updater math
// op 0: v += a*f                      (velocityVerletNVE second half: a = dt/2 ; NVT: a = dt)
// op 1: disp = a*f                    (gradientDescent.cpp:12-14)
// op 2: v = a*v                       (noseHooverNVT velocity rescale)
// op 3: disp = a*v                    (noseHooverNVT half move)
// op 4: v = (1-b)*v + (b*a)*f         (FIRE mixing, a = scaling, b = alpha)
// op 5: v = 0
__global__ void k_axpy(int op, int n, double a, double b, double* __restrict__ vel, const double* __restrict__ frc,
                       double* __restrict__ disp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (op == 0) st3(vel, i, ld3(vel, i) + a * ld3(frc, i));
    else if (op == 1) st3(disp, i, a * ld3(frc, i));
    else if (op == 2) st3(vel, i, a * ld3(vel, i));
    else if (op == 3) st3(disp, i, a * ld3(vel, i));
    else if (op == 4) st3(vel, i, (1 - b) * ld3(vel, i) + (b * a) * ld3(frc, i));
    else if (op == 5) st3(vel, i, d3{0, 0, 0});
}

// Deterministic reductions: every block writes one partial; a single block folds them in order.
// out[0]=sum f.f  out[1]=sum v.v  out[2]=sum f.v  out[3]=max f.f  out[4]=sum 0.5 v.v
// NOTE the reference accumulates these serially (fireMinimization.cpp:23-34, baseUpdater.cpp:22-54);
// a tree order differs by rounding only (documented in DESIGN.md).
__global__ void k_reduce_partial(int n, const double* __restrict__ vel, const double* __restrict__ frc, double* __restrict__ partial)
{
    __shared__ double sm[5][32];
    double ff = 0, vv = 0, fv = 0, mx = 0, ke = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        d3 v = ld3(vel, i), f = ld3(frc, i);
        double a = dot(f, f), b = dot(v, v);
        ff += a;
        vv += b;
        fv += dot(f, v);
        mx = a > mx ? a : mx;
        ke += 0.5 * (1.0) * b;
    }
    double vals[5] = {ff, vv, fv, mx, ke};
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        double x = vals[k];
        for (int o = 16; o; o >>= 1) {
            double y = __shfl_xor_sync(0xffffffffu, x, o);
            x = (k == 3) ? (x > y ? x : y) : x + y;
        }
        if (lane == 0) sm[k][w] = x;
    }
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            double x = lane < nw ? sm[k][lane] : 0.0;
            for (int o = 16; o; o >>= 1) {
                double y = __shfl_xor_sync(0xffffffffu, x, o);
                x = (k == 3) ? (x > y ? x : y) : x + y;
            }
            if (lane == 0) partial[blockIdx.x * 5 + k] = x;
        }
    }
}
__global__ void k_reduce_final(int nb, const double* __restrict__ partial, double* __restrict__ out)
{
    int k = threadIdx.x;
    if (k >= 5) return;
    double x = 0;
    for (int b = 0; b < nb; ++b) {
        double y = partial[b * 5 + k];
        x = (k == 3) ? (x > y ? x : y) : x + y;
    }
    out[k] = x;
}

// force::computeEnergy (baseForce.cpp:33-44): sum over particles, neighbours in list order
__device__ __forceinline__ double pairEnergy(const ForceParams& fp, double d)
{
    if (fp.kind == 0) { // harmonicRepulsion.cpp:3-16
        double ans = 0;
        if (d < fp.sigma) ans = 0.5 * fp.a * (fp.sigma - d) * (fp.sigma - d);
        return ans;
    }
    const double sqrtTwoPi = 2.50662827463100050241576528481104525300698674061;
    double twoSigmaSquared = 2.0 * fp.sigma * fp.sigma;
    return fp.a * exp(-d * d / twoSigmaSquared) / (sqrtTwoPi * fp.sigma);
}
__global__ void k_energy_partial(int n, int kmax, const int* __restrict__ nbrCount, const double* __restrict__ nbrDist, ForceParams fp,
                                 double* __restrict__ partial)
{
    __shared__ double sm[32];
    double e = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int K = nbrCount[i];
        for (int j = 0; j < K; ++j) e += pairEnergy(fp, nbrDist[(size_t)i * kmax + j]);
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0) sm[w] = e;
    __syncthreads();
    if (w == 0) {
        double x = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) partial[blockIdx.x] = x;
    }
}
__global__ void k_energy_final(int nb, const double* __restrict__ partial, double* __restrict__ out)
{
    double x = 0;
    for (int b = 0; b < nb; ++b) x += partial[b];
    out[0] = x;
}

// simulation::computeMonodisperseStress (src/simulation/simulation.cpp:104-173): sums over particles and their neighbours,
// in list order, of force (x) separation and - counted once per NEIGHBOUR, as the reference's loop nest does - v (x) v.
__device__ __forceinline__ d3 pairForceExact(const ForceParams& fp, const d3& sep, double d)
{
    if (fp.kind == 0) { // harmonicRepulsion.cpp:19-33
        if (d <= fp.sigma) return (-fp.a * (fp.sigma - d)) * sep;
        return d3{0, 0, 0};
    }
    const double sqrtTwoPi = 2.50662827463100050241576528481104525300698674061; // gaussianRepulsion.h:16-24
    double twoSigmaSquared = 2.0 * fp.sigma * fp.sigma;
    double pre = d * fp.a * exp(-d * d / twoSigmaSquared) / ((sqrtTwoPi * fp.sigma) * sqrt(fp.sigma));
    return (-pre) * sep;
}
__global__ void k_stress_partial(int n, int kmax, const int* __restrict__ nbrCount, const double* __restrict__ nbrDist,
                                 const double* __restrict__ nbrTs, const double* __restrict__ vel, ForceParams fp, double* __restrict__ partial)
{
    __shared__ double sm[8][18];
    double acc[18];
#pragma unroll
    for (int q = 0; q < 18; ++q) acc[q] = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int K = nbrCount[i];
        d3 v = ld3(vel, i);
        double vv[3] = {v.x, v.y, v.z};
        for (int j = 0; j < K; ++j) {
            size_t o = (size_t)i * kmax + j;
            d3 sep = ld3(nbrTs, o);
            d3 f = pairForceExact(fp, sep, nbrDist[o]);
            double ff[3] = {f.x, f.y, f.z}, ss[3] = {sep.x, sep.y, sep.z};
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) acc[3 * a + b] += ff[a] * ss[b], acc[9 + 3 * a + b] += vv[a] * vv[b];
        }
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 18; ++q) {
        double x = acc[q];
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sm[w][q] = x;
    }
    __syncthreads();
    if (threadIdx.x < 18) {
        double x = 0;
        for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) x += sm[ww][threadIdx.x];
        partial[blockIdx.x * 18 + threadIdx.x] = x;
    }
}
__global__ void k_stress_final(int nb, const double* __restrict__ partial, double* __restrict__ out)
{
    int q = threadIdx.x;
    if (q >= 18) return;
    double x = 0;
    for (int b = 0; b < nb; ++b) x += partial[b * 18 + q];
    out[q] = x;
}

// ---------------------------------------------------------------------------------- host launchers
// ---------------------------------------------------------------------------------------------------------------------
// R^3 point -> mesh position: simpleModel::R3PositionsToMeshPositions (src/models/simpleModel.cpp:136-154), i.e.
// PMP::locate_with_AABB_tree + simpleModel::clampBarycentricCoordinatesToFace (:114-134).  One thread per point.  The AABB
// tree is replaced by a uniform grid over the faces' bounding boxes (built once per mesh, css_api.cu): the thread visits the
// cells around its point shell by shell and stops when the next shell cannot hold anything closer.  The per-face arithmetic
// (closest point by Voronoi region, squared distance, barycentric weights, snap, clamp) is the oracle's, operation for
// operation, so face index and weights are bit-identical to the CPU restatement the tests check against; exact ties go to the lowest face index.
struct FaceGrid {
    double mn[3], h;
    int n[3];
    const int* cellStart; // [ncells + 1]
    const int* cellFaces; // face ids, every face listed in each cell its bounding box overlaps
};

__device__ __forceinline__ d3 closestPointOnTriangle(const d3& p, const d3& a, const d3& b, const d3& c)
{
    const d3 ab = b - a, ac = c - a, ap = p - a;
    const double d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0.0 && d2 <= 0.0) return a;
    const d3 bp = p - b;
    const double d3_ = dot(ab, bp), d4 = dot(ac, bp);
    if (d3_ >= 0.0 && d4 <= d3_) return b;
    const double vc = d1 * d4 - d3_ * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3_ <= 0.0) {
        const double v = d1 / (d1 - d3_);
        return a + v * ab;
    }
    const d3 cp = p - c;
    const double d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0.0 && d5 <= d6) return c;
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
        const double w = d2 / (d2 - d6);
        return a + w * ac;
    }
    const double va = d3_ * d6 - d5 * d4;
    if (va <= 0.0 && (d4 - d3_) >= 0.0 && (d5 - d6) >= 0.0) {
        const double w = (d4 - d3_) / ((d4 - d3_) + (d5 - d6));
        return b + w * (c - b);
    }
    const double denom = 1.0 / (va + vb + vc);
    const double v = vb * denom, w = vc * denom;
    return (a + v * ab) + w * ac;
}

__device__ __forceinline__ void locateWeights(const d3& q, const d3& p0, const d3& p1, const d3& p2, double clampTol, double* out)
{
    const d3 v0 = p1 - p0, v1 = p2 - p0, v2 = q - p0;
    const double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    const double den = d00 * d11 - d01 * d01;
    const double v = (d11 * d20 - d01 * d21) / den, w = (d00 * d21 - d01 * d20) / den;
    double co[3] = {(1.0 - v) - w, v, w};
    if (co[0] < 0.0 || co[0] > 1.0 || co[1] < 0.0 || co[1] > 1.0 || co[2] < 0.0 || co[2] > 1.0) {
        const double eps = 2.220446049250313e-16;
        double residue = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (fabs(co[i]) <= eps) residue = residue + co[i], co[i] = 0.0;
            else if (fabs(1.0 - co[i]) <= eps) residue = residue - (1.0 - co[i]), co[i] = 1.0;
        }
        bool dumped = false;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (!dumped && co[i] != 0.0 && co[i] != 1.0) co[i] = co[i] + residue, dumped = true;
    }
    double w1 = co[0], w2 = co[1], w3 = co[2];
    if (fabs(w1) < clampTol) w1 = clampTol;
    if (fabs(w2) < clampTol) w2 = clampTol;
    if (fabs(w3) < clampTol) w3 = clampTol;
    w1 = w1 / ((w1 + w2) + w3);
    w2 = w2 / ((w1 + w2) + w3);
    w3 = w3 / ((w1 + w2) + w3);
    out[0] = w1, out[1] = w2, out[2] = w3;
}

__global__ void __launch_bounds__(128) k_locate(MeshDev m, FaceGrid g, int n, const double* __restrict__ xyz, double clampTol,
                                                int* __restrict__ face, double* __restrict__ bary)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const d3 p = ld3(xyz, i);
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) { // nothing is closest to such a point: reported by the host
        face[i] = -1;
        bary[3 * i] = bary[3 * i + 1] = bary[3 * i + 2] = 0.0;
        return;
    }
    int c0[3];
    {
        const double q[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double t = floor((q[d] - g.mn[d]) / g.h);
            t = t < 0.0 ? 0.0 : t;
            c0[d] = min(g.n[d] - 1, (int)fmin(t, 2.0e9));
        }
    }
    const int rmax = max(g.n[0], max(g.n[1], g.n[2]));
    double best = __longlong_as_double(0x7ff0000000000000LL);
    int bf = -1;
    for (int r = 0; r <= rmax; ++r) {
        const int x0 = max(0, c0[0] - r), x1 = min(g.n[0] - 1, c0[0] + r);
        const int y0 = max(0, c0[1] - r), y1 = min(g.n[1] - 1, c0[1] + r);
        const int z0 = max(0, c0[2] - r), z1 = min(g.n[2] - 1, c0[2] + r);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
                const bool edge = abs(z - c0[2]) == r || abs(y - c0[1]) == r;
                for (int x = x0; x <= x1; x += (edge || r == 0) ? 1 : max(1, x1 - x0)) { // interior rows: only the two end cells
                    if (!edge && abs(x - c0[0]) != r) continue;
                    const size_t cell = ((size_t)z * g.n[1] + y) * g.n[0] + x;
                    for (int k = g.cellStart[cell], ke = g.cellStart[cell + 1]; k < ke; ++k) {
                        const int f = g.cellFaces[k];
                        const int4 c = __ldg(m.corner + f);
                        const d3 q = closestPointOnTriangle(p, ldvert(m, c.x), ldvert(m, c.y), ldvert(m, c.z));
                        const double d2 = sqlen(p - q);
                        if (d2 < best || (d2 == best && f < bf)) best = d2, bf = f;
                    }
                }
            }
        // every cell of shell r + 1 is at least r h away from a point inside (or clamped into) the centre cell
        const double lb = (double)r * g.h;
        if (bf >= 0 && best < lb * lb * (1.0 - 1e-9)) break;
    }
    face[i] = bf;
    if (bf >= 0) {
        const int4 c = __ldg(m.corner + bf);
        const d3 p0 = ldvert(m, c.x), p1 = ldvert(m, c.y), p2 = ldvert(m, c.z);
        locateWeights(closestPointOnTriangle(p, p0, p1, p2), p0, p1, p2, clampTol, bary + 3 * i);
    } else
        bary[3 * i] = bary[3 * i + 1] = bary[3 * i + 2] = 0.0;
}

void launchLocate(cudaStream_t st, const MeshDev& m, const double gmn[3], double h, const int gn[3], const int* cellStart, const int* cellFaces,
                  int n, const double* xyz, double clampTol, int* face, double* bary)
{
    if (n <= 0) return;
    FaceGrid g;
    for (int d = 0; d < 3; ++d) g.mn[d] = gmn[d], g.n[d] = gn[d];
    g.h = h, g.cellStart = cellStart, g.cellFaces = cellFaces;
    k_locate<<<(n + 127) / 128, 128, 0, st>>>(m, g, n, xyz, clampTol, face, bary);
}

static inline int gridFor(int n, int b) { return (n + b - 1) / b; }

void launchEuclidCell(cudaStream_t st, const MeshDev& m, const CellGrid& g, int n, const int* face, const double* bary, double* eucl,
                      int* cellOf, int* cellCount, int* cellSlot, int* coarse)
{
    if (n > 0) launchStep(k_euclid_cell, gridFor(n, 256), 256, 0, st, m, g, n, face, bary, eucl, cellOf, cellCount, cellSlot, coarse);
}
// cellCount[nCells] is the bump counter (cleared with the counts)
void launchCellBuild(cudaStream_t st, int n, int nCells, const int* cellOf, const int* cellSlot, int* cellCount, int* cellStart, int* tmpItems,
                     int* items, const int* coarse, const CellGrid& g, int kmax, unsigned long long* counters)
{
    if (n <= 0) return;
    launchStep(k_cell_place, gridFor(n, 256), 256, 0, st, n, cellOf, cellSlot, cellCount, cellStart, cellCount + nCells);
    launchStep(k_cell_fill, gridFor(n, 256), 256, 0, st, n, cellOf, cellSlot, cellStart, tmpItems);
    launchStep(k_cell_rank, gridFor(n, 256), 256, 0, st, n, cellOf, cellStart, cellCount, tmpItems, items, coarse, g.n[0], g.n[1], g.n[2], kmax, counters);
}
void launchWalk(cudaStream_t st, const MeshDev& m, int n, int minIdx, int* face, double* bary, double* disp, double* vel, double* frc,
                int transportForce, int transportVelocity, int mode, double dt, int* flags, unsigned long long* counters, const PeerWin& pw)
{
    if (n > 0)
        launchStep(k_walk, gridFor(n, 64), 64, 0, st, m, n, minIdx, face, bary, disp, vel, frc, transportForce, transportVelocity, mode, dt, flags,
                   counters, pw);
}

// ---------------------------------------------------------------------------------- peer-memory exchange
namespace {
__device__ __forceinline__ unsigned long long ldAcquireSys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stReleaseSys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globalTimerNs()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *p >= e; gives up after 10 s (a rank died or left the collective sequence) and counts the event
__device__ __forceinline__ void waitFlag(const unsigned long long* p, unsigned long long e, unsigned long long* counters)
{
    if (ldAcquireSys(p) >= e) return;
    const unsigned long long t0 = globalTimerNs();
    while (ldAcquireSys(p) < e) {
        __nanosleep(64);
        if (globalTimerNs() - t0 > 10000000000ull) {
            atomicAdd(counters + C_PEER_TIMEOUT, 1ull);
            return;
        }
    }
}
} // namespace

// before the walker: every peer must have drained the previous epoch from its staging copy
__global__ void k_peer_wait_consumed(PeerWin pw, const unsigned long long* __restrict__ epoch, unsigned long long* counters)
{
    const int r = threadIdx.x;
    if (r < pw.n && r != pw.rank) waitFlag(pw.flags[pw.rank] + (CSS_MAX_PEERS + r) * CSS_PEER_FLAG_STRIDE, *epoch, counters);
}
// after the walker: publish "this rank finished epoch e" in every window, wait until every rank did, advance the epoch
__global__ void k_peer_barrier(PeerWin pw, unsigned long long* epoch, unsigned long long* counters)
{
    const int r = threadIdx.x;
    const unsigned long long e = *epoch + 1;
    __syncwarp();
    __threadfence_system();
    if (r < pw.n && r != pw.rank) stReleaseSys(pw.flags[r] + pw.rank * CSS_PEER_FLAG_STRIDE, e);
    if (r < pw.n && r != pw.rank) waitFlag(pw.flags[pw.rank] + r * CSS_PEER_FLAG_STRIDE, e, counters);
    __syncwarp();
    if (r == 0) *epoch = e;
}
// staging -> live arrays for the particles the other ranks own; the last block publishes "this rank drained epoch e"
__global__ void k_peer_copy(PeerWin pw, int nTotal, int lo, int hi, int* __restrict__ face, double* __restrict__ bary,
                            const unsigned long long* __restrict__ epoch, unsigned* ticket)
{
    const int* sf = pw.face[pw.rank];
    const double* sb = pw.bary[pw.rank];
    const int nOther = nTotal - (hi - lo);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nOther; q += gridDim.x * blockDim.x) {
        const int i = q < lo ? q : q + (hi - lo);
        face[i] = __ldcg(sf + i);
        const double b0 = __ldcg(sb + 3 * (size_t)i), b1 = __ldcg(sb + 3 * (size_t)i + 1), b2 = __ldcg(sb + 3 * (size_t)i + 2);
        bary[3 * (size_t)i] = b0, bary[3 * (size_t)i + 1] = b1, bary[3 * (size_t)i + 2] = b2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1) {
            const unsigned long long e = *epoch;
            for (int r = 0; r < pw.n; ++r)
                if (r != pw.rank) stReleaseSys(pw.flags[r] + (CSS_MAX_PEERS + pw.rank) * CSS_PEER_FLAG_STRIDE, e);
        }
    }
}
void launchPeerWaitConsumed(cudaStream_t st, const PeerWin& pw, const unsigned long long* epoch, unsigned long long* counters)
{
    k_peer_wait_consumed<<<1, 32, 0, st>>>(pw, epoch, counters);
}
void launchPeerBarrier(cudaStream_t st, const PeerWin& pw, unsigned long long* epoch, unsigned long long* counters)
{
    k_peer_barrier<<<1, 32, 0, st>>>(pw, epoch, counters);
}
void launchPeerCopy(cudaStream_t st, const PeerWin& pw, int nTotal, int lo, int hi, int* face, double* bary, const unsigned long long* epoch,
                    unsigned* ticket)
{
    int nOther = nTotal - (hi - lo);
    int blocks = max(1, min(296, gridFor(max(nOther, 1), 256)));
    k_peer_copy<<<blocks, 256, 0, st>>>(pw, nTotal, lo, hi, face, bary, epoch, ticket);
}
void launchTransportGeneric(cudaStream_t st, const MeshDev& m, int n, int* face, double* bary, double* disp, int nVec, double* vecs,
                            int* flags)
{
    if (n > 0) k_transport_generic<<<gridFor(n, 128), 128, 0, st>>>(m, n, face, bary, disp, nVec, vecs, flags);
}
// Nose-Hoover chain on the device (noseHooverNVT.cpp:65-110, the same statements in the same order as the host version in css_api.cu;
// this file is compiled without FMA contraction, so only exp() may differ from the host's libm in the last bit).  One thread.
// nh = bx[M+1] | by[M+1] | bz[M+1] | bw[M+1] | KE, scale, T, dt2, dt4, dt8.  Frozen behind the stride guard like every step kernel.
__global__ void k_nh_chain(double* __restrict__ nh, int M, const double* __restrict__ red, int takeKE, const unsigned long long* counters)
{
    if (threadIdx.x || blockIdx.x) return;
    if (strideGuardUp(counters)) return;
    double *bx = nh, *by = nh + (M + 1), *bz = nh + 2 * (M + 1), *bw = nh + 3 * (M + 1), *sc = nh + 4 * (M + 1);
    if (takeKE) sc[0] = red[4];
    double KE = sc[0];
    const double T = sc[2], dt2 = sc[3], dt4 = sc[4], dt8 = sc[5];
    double ef = 0;
    for (int ii = M - 1; ii > 0; --ii) {
        bz[ii] = (bw[ii - 1] * by[ii - 1] * by[ii - 1] - T) / bw[ii];
        ef = exp(-dt8 * by[ii + 1]);
        by[ii] *= ef;
        by[ii] += bz[ii] * dt4;
        by[ii] *= ef;
    }
    bz[0] = (2.0 * KE / bw[0] - 1.0);
    ef = exp(-dt8 * by[1]);
    by[0] *= ef;
    by[0] += bz[0] * dt4;
    by[0] *= ef;
    for (int ii = 0; ii < M; ++ii) bx[ii] += dt2 * by[ii];
    const double scale = exp(-dt2 * by[0]);
    KE = scale * scale * KE;
    bz[0] = (2.0 * KE / bw[0] - 1.0);
    ef = exp(-dt8 * by[1]);
    by[0] *= ef;
    by[0] += bz[0] * dt4;
    by[0] *= ef;
    for (int ii = 1; ii < M; ++ii) {
        bz[ii] = (bw[ii - 1] * by[ii - 1] * by[ii - 1] - T) / bw[ii];
        ef = exp(-dt8 * by[ii + 1]);
        by[ii] *= ef;
        by[ii] += bz[ii] * dt4;
        by[ii] *= ef;
    }
    sc[0] = KE, sc[1] = scale;
}
// v *= *s with the factor in device memory (the chain's velocity scale); frozen behind the stride guard
__global__ void k_scale_dev(int n, const double* __restrict__ s, double* __restrict__ vel, const unsigned long long* counters)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || strideGuardUp(counters)) return;
    const double a = *s;
    st3(vel, i, a * ld3(vel, i));
}
void launchNhChain(cudaStream_t st, double* nh, int M, const double* red, int takeKE, const unsigned long long* counters)
{
    k_nh_chain<<<1, 32, 0, st>>>(nh, M, red, takeKE, counters);
}
void launchScaleDev(cudaStream_t st, int n, const double* s, double* vel, const unsigned long long* counters)
{
    if (n > 0) k_scale_dev<<<gridFor(n, 256), 256, 0, st>>>(n, s, vel, counters);
}
void launchAxpy(cudaStream_t st, int op, int n, double a, double b, double* vel, const double* frc, double* disp)
{
    if (n > 0) k_axpy<<<gridFor(n, 256), 256, 0, st>>>(op, n, a, b, vel, frc, disp);
}
void launchReduce(cudaStream_t st, int n, const double* vel, const double* frc, double* partial, double* out)
{
    int nb = n > 0 ? min(REDUCE_MAX_BLOCKS, gridFor(n, 256)) : 1;
    k_reduce_partial<<<nb, 256, 0, st>>>(n, vel, frc, partial);
    k_reduce_final<<<1, 32, 0, st>>>(nb, partial, out);
}

void launchEnergy(cudaStream_t st, int nLocal, int kmax, const int* nbrCount, const double* nbrDist, ForceParams fp, double* partial,
                  double* out)
{
    int nb = nLocal > 0 ? min(REDUCE_MAX_BLOCKS, gridFor(nLocal, 256)) : 1;
    k_energy_partial<<<nb, 256, 0, st>>>(nLocal, kmax, nbrCount, nbrDist, fp, partial);
    k_energy_final<<<1, 1, 0, st>>>(nb, partial, out);
}

void launchStress(cudaStream_t st, int nLocal, int kmax, const int* nbrCount, const double* nbrDist, const double* nbrTs, const double* vel,
                  ForceParams fp, double* partial, double* out)
{
    int nb = nLocal > 0 ? min(REDUCE_MAX_BLOCKS, gridFor(nLocal, 256)) : 1;
    k_stress_partial<<<nb, 256, 0, st>>>(nLocal, kmax, nbrCount, nbrDist, nbrTs, vel, fp, partial);
    k_stress_final<<<1, 32, 0, st>>>(nb, partial, out);
}

} // namespace css
