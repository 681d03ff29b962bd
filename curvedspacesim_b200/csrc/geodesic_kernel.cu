// Fused exact geodesic kernel for long ranges: ONE CTA PER SOURCE PARTICLE (k_geodesic_cta).
//
// For its source the CTA (1) gathers the ordered Euclidean candidate list from the cell list
// (cellListNeighborStructure::constructCandidateNeighborList, src/utility/cellListNeighborStructure.cpp:45-84) -- or takes the
// explicit targets of a css_distance query, or everybody else in the all-to-all mode --, (2) flood-fills the local patch with the
// rule of submesher::constructSubmeshFromSourceAndTargets (src/utility/submesher.cpp:55-147) into its workspace (shared memory
// for the long-range tier, global memory for the whole-mesh tier), (3) runs exact window propagation (Chen-Han unfolding with the
// Xin-Wang vertex-distance filter; saddle and patch-boundary vertices are pseudo-sources) over a FIFO ring of windows, one window
// per thread and pass, (4) answers the K target queries (distance, start tangent, end tangent), which replaces
// CGAL::Surface_mesh_shortest_path as used at src/models/triangulatedMeshSpace.cpp:189-203, :212-238 and
// src/utility/meshUtilities.cpp:360-380, and (5) optionally accumulates the pair force of force::computeForces
// (src/forces/baseForce.cpp:12-28) in neighbour order and applies the velocity half-kick (src/updaters/velocityVerletNVE.cpp:27-28).
// The short-range work of a step never comes here: it runs on the two-stage record kernels (stencil_kernel.cu / patch_kernel.cu,
// window_half_kernel.cu / window_kernel.cu); this kernel takes what outgrows their records, explicit queries and the all-to-all mode.
//
// Decisions that fix topology (candidate membership/order, patch membership) use exactly-rounded
// arithmetic (common.cuh x* helpers) and are bit-identical to the CPU oracle; the window geometry is
// ordinary fp64 with FMA contraction.  No tensor cores: nothing here is a dense contraction.
#include "common.cuh"
#include "kernels.h"

namespace css {

#define FULL 0xffffffffu
#define NONE16 0xFFFFu
static __device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000LL); }

enum { ST_OK = 0, ST_OVERFLOW = 1, ST_OVF_K = 1, ST_OVF_F = 2, ST_OVF_V = 3, ST_OVF_RING = 4 };

struct WS { // per-source workspace (one per CTA), carved from one contiguous block; arrays addressed as base + k*capacity
    double *dv, *dt, *dr, *root;     // per-vertex [7*maxV], per-target [13*kt], ring [9*ring], root frame [16]
    unsigned long long* wcnt;        // [16] per-CTA counters
    int *gface, *gvert, *fhKey, *vhKey, *tIdx, *tGFace, *rmeta, *misc;
    unsigned short *fvert /*4*maxF*/, *fadj /*4*maxF*/, *fhVal, *vhVal, *tFace, *tCorner /*4*kt*/, *rpsv;
    unsigned char *velig, *vdirty;
    int maxV, kt, ring;
#define WS_ACC(name, base, cap, k) __device__ __forceinline__ double* name() const { return base + (size_t)(k) * cap; }
    WS_ACC(vx, dv, maxV, 0) WS_ACC(vy, dv, maxV, 1) WS_ACC(vz, dv, maxV, 2) WS_ACC(D, dv, maxV, 3)
    WS_ACC(dirx, dv, maxV, 4) WS_ACC(diry, dv, maxV, 5) WS_ACC(dirz, dv, maxV, 6)
    WS_ACC(tbest, dt, kt, 0) WS_ACC(tsx, dt, kt, 1) WS_ACC(tsy, dt, kt, 2) WS_ACC(tsz, dt, kt, 3)
    WS_ACC(tex, dt, kt, 4) WS_ACC(tey, dt, kt, 5) WS_ACC(tez, dt, kt, 6) WS_ACC(tpx, dt, kt, 7) WS_ACC(tpy, dt, kt, 8)
    WS_ACC(tpz, dt, kt, 9) WS_ACC(tb0, dt, kt, 10) WS_ACC(tb1, dt, kt, 11) WS_ACC(tb2, dt, kt, 12)
    WS_ACC(rax, dr, ring, 0) WS_ACC(ray, dr, ring, 1) WS_ACC(rbx, dr, ring, 2) WS_ACC(rby, dr, ring, 3)
    WS_ACC(rsx, dr, ring, 4) WS_ACC(rsy, dr, ring, 5) WS_ACC(rt0, dr, ring, 6) WS_ACC(rt1, dr, ring, 7) WS_ACC(rsg, dr, ring, 8)
#undef WS_ACC
};

__host__ __device__ inline size_t al8(size_t x) { return (x + 7) & ~(size_t)7; }

__host__ __device__ inline size_t wsLayout(const GeoCaps& c, char* base, WS* w)
{
    size_t o = 0;
    WS dummy;
    WS& r = w ? *w : dummy;
    r.maxV = c.maxV, r.kt = c.kt, r.ring = c.ring;
    r.dv = reinterpret_cast<double*>(base + o), o += 8 * 7 * (size_t)c.maxV;
    r.dt = reinterpret_cast<double*>(base + o), o += 8 * 13 * (size_t)c.kt;
    r.dr = reinterpret_cast<double*>(base + o), o += 8 * 9 * (size_t)c.ring;
    r.root = reinterpret_cast<double*>(base + o), o += 8 * 16;
    r.wcnt = reinterpret_cast<unsigned long long*>(base + o), o += 8 * 16;
    auto I = [&](int*& p, size_t n) {
        p = reinterpret_cast<int*>(base + o);
        o += 4 * n;
    };
    auto S = [&](unsigned short*& p, size_t n) {
        p = reinterpret_cast<unsigned short*>(base + o);
        o += 2 * n;
    };
    auto B = [&](unsigned char*& p, size_t n) {
        p = reinterpret_cast<unsigned char*>(base + o);
        o += n;
    };
    I(r.gface, c.maxF), I(r.gvert, c.maxV), I(r.fhKey, c.hashF), I(r.vhKey, c.hashV), I(r.tIdx, c.kt), I(r.tGFace, c.kt);
    I(r.rmeta, c.ring), I(r.misc, 8);
    S(r.fvert, 4 * (size_t)c.maxF), S(r.fadj, 4 * (size_t)c.maxF), S(r.fhVal, c.hashF), S(r.vhVal, c.hashV), S(r.tFace, c.kt);
    S(r.tCorner, 4 * (size_t)c.kt), S(r.rpsv, c.ring);
    o = al8(o);
    B(r.velig, c.maxV), B(r.vdirty, c.maxV);
    return al8(o + 8);
}
size_t geoWorkspaceBytes(const GeoCaps& c) { return wsLayout(c, nullptr, nullptr); }

// ---- small helpers --------------------------------------------------------------------------------
struct v2 {
    double x, y;
};
__device__ __forceinline__ v2 operator-(const v2& a, const v2& b) { return v2{a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ v2 lerp2(const v2& a, const v2& b, double t) { return v2{a.x + t * (b.x - a.x), a.y + t * (b.y - a.y)}; }
__device__ __forceinline__ double cross2(const v2& a, const v2& b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ double len2(const v2& a) { return sqrt(a.x * a.x + a.y * a.y); }
__device__ __forceinline__ double dist2(const v2& a, const v2& b) { return len2(a - b); }

__device__ __forceinline__ unsigned hashInt(int k) { return (unsigned)k * 2654435761u; }

// open-addressing insert; returns slot; *isNew set when this call created the key
__device__ __forceinline__ int hashInsert(int* keys, int mask, int key, bool& isNew)
{
    unsigned h = (hashInt(key) >> 7) & mask;
    for (;;) {
        int old = atomicCAS(keys + h, -1, key);
        if (old == -1) {
            isNew = true;
            return (int)h;
        }
        if (old == key) {
            isNew = false;
            return (int)h;
        }
        h = (h + 1) & mask;
    }
}
__device__ __forceinline__ int hashFind(const int* keys, int mask, int key)
{
    unsigned h = (hashInt(key) >> 7) & mask;
    for (;;) {
        int k = keys[h];
        if (k == key) return (int)h;
        if (k == -1) return -1;
        h = (h + 1) & mask;
    }
}

__device__ __forceinline__ int warpInclusiveScan(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += y;
    }
    return v;
}
__device__ __forceinline__ double warpMax(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

__device__ __forceinline__ bool atomicMinD(double* addr, double v)
{ // non-negative doubles order like their bit patterns
    unsigned long long nv = (unsigned long long)__double_as_longlong(v);
    unsigned long long old = atomicMin(reinterpret_cast<unsigned long long*>(addr), nv);
    return nv < old;
}

__device__ __forceinline__ double frcp(double x);
__device__ __forceinline__ double hitParam(const v2& S, const v2& P, const v2& X, const v2& Y)
{ // ray S->P against X + mu (Y - X), clamped
    v2 d = P - S;
    double den = cross2(Y - X, d);
    double mu = cross2(S - X, d) * frcp(den);
    if (!(mu == mu)) mu = 0.5;
    return fmin(1.0, fmax(0.0, mu));
}

// fast (not correctly rounded, ~1 ulp) reciprocal / square root: hardware seed + Newton steps.  Used for the
// window geometry only; every topology-fixing decision uses the exactly rounded x* helpers of common.cuh.
__device__ __forceinline__ double frcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}
__device__ __forceinline__ double fsqrt(double x)
{
    if (!(x > 0)) return 0.0;
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    double s = x * y;
    return fma(fma(-s, s, x), 0.5 * y, s); // one correction step on the square root itself
}
// float-precision lengths for pruning decisions (always used with a conservative margin)
__device__ __forceinline__ float alen2(const v2& a) { return sqrtf((float)(a.x * a.x + a.y * a.y)); }
__device__ __forceinline__ float adist2(const v2& a, const v2& b) { return alen2(a - b); }
__device__ __forceinline__ float asegDist(const v2& S, const v2& X0, const v2& X1)
{
    float ex = (float)(X1.x - X0.x), ey = (float)(X1.y - X0.y), sx = (float)(S.x - X0.x), sy = (float)(S.y - X0.y);
    float L2 = ex * ex + ey * ey;
    float s = L2 > 0.f ? __fdividef(sx * ex + sy * ey, L2) : 0.f;
    s = fminf(1.f, fmaxf(0.f, s));
    float dx = sx - s * ex, dy = sy - s * ey;
    return sqrtf(dx * dx + dy * dy);
}

struct Child {
    int valid;
    int meta;
    v2 A, B;
    double t0, t1;
};

// pair force on the device: functor id + parameter block (a virtual pairwiseForce cannot be called here)
__device__ __forceinline__ d3 pairForce(const ForceParams& fp, const d3& sep, double d)
{
    if (fp.kind == 0) { // harmonicRepulsion.cpp:19-33
        if (d <= fp.sigma) {
            double s = -fp.a * (fp.sigma - d);
            return d3{s * sep.x, s * sep.y, s * sep.z};
        }
        return d3{0, 0, 0};
    }
    const double sqrtTwoPi = 2.50662827463100050241576528481104525300698674061; // gaussianRepulsion.h:16-24
    double twoSigmaSquared = 2.0 * fp.sigma * fp.sigma;
    double s32 = (sqrtTwoPi * fp.sigma) * sqrt(fp.sigma);
    double pre = d * fp.a * exp(-d * d / twoSigmaSquared) / s32;
    return d3{-pre * sep.x, -pre * sep.y, -pre * sep.z};
}

// =====================================================================================================================
// ONE CTA (CTA_NT threads) PER SOURCE.  The reference's default executable (N = 20 on torus_isotropic_remesh.off, range 2.6:
// patches of ~830 faces, ~7 000 windows per source, triangulatedMeshSpace.cpp:212-238) would leave one warp alone on an SM with
// nothing to hide its latencies behind (measured: 6.4 ms per step); here CTA_NT (384) windows are popped per pass (one per thread,
// both children), children are compacted into the ring by a two-level block scan (deterministic order), target and vertex
// improvements go through 64-bit atomic minima whose unique last writer stores the start / end directions after a barrier, and
// all pseudo-source fans of a round are spawned in ONE sweep over the (face, corner) pairs of the patch (0.74 ms per step).
// =====================================================================================================================
#ifndef CSS_CTA_NT
#define CSS_CTA_NT 384 // threads per source (build parameter; default-executable shape: 128 -> 0.87, 256 -> 0.67, 384 -> 0.63, 512 -> 0.65 ms)
#endif
constexpr int CTA_NT = CSS_CTA_NT, CTA_NW = CTA_NT / 32;
struct CtaScratch {
    int scan[2][CTA_NW];
    int src;
    int flag;
};
// exclusive prefix of v over the block in thread order; `total` = block sum.  One barrier: the scratch rows alternate.
__device__ __forceinline__ int blockExclusiveScan(int v, int tid, CtaScratch& sc, int& parity, int& total)
{
    const int lane = tid & 31, wid = tid >> 5;
    const int incl = warpInclusiveScan(v, lane);
    int* buf = sc.scan[parity];
    parity ^= 1;
    if (lane == 31) buf[wid] = incl;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < CTA_NW; ++k) {
        const int x = buf[k];
        base += k < wid ? x : 0;
        tot += x;
    }
    total = tot;
    return base + incl - v;
}

template <bool GLOBAL_WS>
__device__ int processSourceCta(const GeoArgs& a, const WS& w, int li, int tid, CtaScratch& sc, int& parity)
{
    const int lane = tid & 31, wid = tid >> 5;
    const GeoCaps& cp = a.caps;
    unsigned long long* cnt = w.wcnt;
    const long long tp0 = clock64();
    const bool explicitQ = a.xK >= 0;
    const int gi = a.minIdx + li;
    const int maskF = cp.hashF - 1, maskV = cp.hashV - 1, maskR = cp.ring - 1;

    // ---------------- source ----------------
    int sf;
    double sb0, sb1, sb2;
    d3 sp;
    if (explicitQ) {
        sf = a.xSrcFace;
        sb0 = a.xSrcBary[0], sb1 = a.xSrcBary[1], sb2 = a.xSrcBary[2];
        int4 c = __ldg(a.m.corner + sf);
        sp = xpoint(ldvert(a.m, c.x), ldvert(a.m, c.y), ldvert(a.m, c.z), sb0, sb1, sb2);
    } else {
        sf = a.face[gi];
        sb0 = a.bary[3 * gi], sb1 = a.bary[3 * gi + 1], sb2 = a.bary[3 * gi + 2];
        sp = d3{a.eucl[3 * gi], a.eucl[3 * gi + 1], a.eucl[3 * gi + 2]};
    }

    // ---------------- 1. ordered candidates (every warp repeats the gather; warp 0 stores) ----------------
    int K = 0;
    double R;
    if (explicitQ) {
        K = a.xK;
        R = a.xThreshold;
        if (K > cp.kt) return ST_OVF_K;
        for (int t = tid; t < K; t += CTA_NT) w.tIdx[t] = t;
    } else if (a.cellStart) {
        const CellGrid& g = a.grid;
        int ix = cellCoord(g, sp.x, 0), iy = cellCoord(g, sp.y, 1), iz = cellCoord(g, sp.z, 2);
        int x0 = max(0, ix - 1), x1 = min(g.n[0] - 1, ix + 1);
        int y0 = max(0, iy - 1), y1 = min(g.n[1] - 1, iy + 1);
        int z0 = max(0, iz - 1), z1 = min(g.n[2] - 1, iz + 1);
        int ny = y1 - y0 + 1, nz = z1 - z0 + 1, ncell = (x1 - x0 + 1) * ny * nz;
        int s0 = 0, s1 = 0;
        if (lane < ncell) { // stencil order: xx outer, yy, zz inner (hyperRectangularCellList.cpp:141-144)
            int xx = x0 + lane / (ny * nz), rem = lane % (ny * nz);
            int yy = y0 + rem / nz, zz = z0 + rem % nz;
            int c = xx + yy * g.n[0] + zz * g.n[0] * g.n[1];
            s0 = a.cellStart[c];
            s1 = s0 + a.cellCount[c];
        }
        int mine = 0;
        double maxd2 = 0;
        for (int s = s0; s < s1; ++s) {
            int j = a.cellItems[s];
            if (j == gi) continue;
            d3 q{a.eucl[3 * j], a.eucl[3 * j + 1], a.eucl[3 * j + 2]};
            double d2 = xsqlen(xsub3(sp, q));
            if (d2 < g.range2) {
                mine++;
                maxd2 = d2 > maxd2 ? d2 : maxd2;
            }
        }
        int incl = warpInclusiveScan(mine, lane);
        K = __shfl_sync(FULL, incl, 31);
        maxd2 = warpMax(maxd2);
        R = xsqrt(maxd2);
        if (K > cp.kt) return ST_OVF_K;
        if (wid == 0) {
            int pos = incl - mine;
            for (int s = s0; s < s1; ++s) {
                int j = a.cellItems[s];
                if (j == gi) continue;
                d3 q{a.eucl[3 * j], a.eucl[3 * j + 1], a.eucl[3 * j + 2]};
                double d2 = xsqlen(xsub3(sp, q));
                if (d2 < g.range2) w.tIdx[pos++] = j;
            }
        }
    } else { // baseNeighborStructure::constructCandidateNeighborList: everybody else, VERYLARGEDOUBLE
        K = a.nTotal - 1;
        R = 1e20;
        if (K > cp.kt) return ST_OVF_K;
        for (int t = tid; t < K; t += CTA_NT) w.tIdx[t] = t < gi ? t : t + 1;
    }
    if (!explicitQ && K > a.kmax) {
        if (tid == 0) atomicMax(a.counters + C_KMAX_NEED, (unsigned long long)K), atomicAdd(a.counters + C_KMAX_OVERFLOW, 1ull);
        K = a.kmax; // truncated; the host grows kmax and reruns
    }
    __syncthreads();
    if (K == 0) {
        if (tid == 0) {
            cnt[C_SOURCES]++;
            if (!explicitQ) {
                a.nbrCount[li] = 0;
                if (a.forceMode) {
                    d3 f = a.zero ? d3{0, 0, 0} : d3{a.frc[3 * li], a.frc[3 * li + 1], a.frc[3 * li + 2]};
                    a.frc[3 * li] = f.x, a.frc[3 * li + 1] = f.y, a.frc[3 * li + 2] = f.z;
                }
            }
        }
        return ST_OK;
    }
    // triangulatedMeshSpace::distanceWithSubmeshing :167-169
    double thr2 = dinf();
    if (a.submeshing) {
        double thr = a.maxDist;
        if (R < a.maxDist) thr = R;
        thr2 = xmul(thr, thr);
    }

    // ---------------- 2. targets ----------------
    for (int t = tid; t < K; t += CTA_NT) {
        int tf;
        double b0, b1, b2;
        d3 tp;
        if (explicitQ) {
            tf = a.xTgtFace[t];
            b0 = a.xTgtBary[3 * t], b1 = a.xTgtBary[3 * t + 1], b2 = a.xTgtBary[3 * t + 2];
            int4 c = __ldg(a.m.corner + tf);
            tp = xpoint(ldvert(a.m, c.x), ldvert(a.m, c.y), ldvert(a.m, c.z), b0, b1, b2);
        } else {
            int j = w.tIdx[t];
            tf = a.face[j];
            b0 = a.bary[3 * j], b1 = a.bary[3 * j + 1], b2 = a.bary[3 * j + 2];
            tp = d3{a.eucl[3 * j], a.eucl[3 * j + 1], a.eucl[3 * j + 2]};
        }
        w.tGFace[t] = tf;
        w.tb0()[t] = b0, w.tb1()[t] = b1, w.tb2()[t] = b2;
        w.tpx()[t] = tp.x, w.tpy()[t] = tp.y, w.tpz()[t] = tp.z;
        w.tbest()[t] = dinf();
        w.tsx()[t] = 0, w.tsy()[t] = 0, w.tsz()[t] = 0;
        w.tex()[t] = 0, w.tey()[t] = 0, w.tez()[t] = 0;
    }
    for (int h = tid; h < cp.hashF; h += CTA_NT) w.fhKey[h] = -1;
    for (int h = tid; h < cp.hashV; h += CTA_NT) w.vhKey[h] = -1;
    if (tid == 0) {
        w.misc[0] = 0; // nF
        w.misc[1] = 0; // nV
        w.misc[2] = 0; // overflow
    }
    __syncthreads();

    // ---------------- 3. patch flood fill (submesher.cpp:55-147), one frontier per round ----------------
    auto addFace = [&](int g) {
        bool isNew;
        int slot = hashInsert(w.fhKey, maskF, g, isNew);
        if (isNew) {
            int id = atomicAdd(&w.misc[0], 1);
            if (id < cp.maxF) {
                w.gface[id] = g;
                w.fhVal[slot] = (unsigned short)id;
            } else
                w.misc[2] = 1;
        }
    };
    if (tid == 0) addFace(sf);
    __syncthreads();
    {
        bool goal = false;
        for (int t = tid; t < K; t += CTA_NT) goal |= (w.tGFace[t] != sf);
        if (__syncthreads_or(goal)) {
            if (tid < 3) {
                int4 sadj = __ldg(a.m.adj + sf);
                int g = tid == 0 ? sadj.x : (tid == 1 ? sadj.y : sadj.z);
                if (g >= 0) addFace(g);
            }
            __syncthreads();
            bool missing = false;
            for (int t = tid; t < K; t += CTA_NT) missing |= (hashFind(w.fhKey, maskF, w.tGFace[t]) < 0);
            if (__syncthreads_or(missing)) {
                int head = 1;
                for (;;) {
                    __syncthreads();
                    const int tail = min(w.misc[0], cp.maxF);
                    const int ovf = w.misc[2];
                    __syncthreads();
                    if (ovf) return ST_OVF_F;
                    if (head >= tail) break;
                    // one thread per (frontier face, edge): the three dependent L2 round trips of a candidate (adjacency, corners,
                    // vertices) run side by side for the whole ring instead of edge after edge
                    for (int item = 3 * head + tid; item < 3 * tail; item += CTA_NT) {
                        const int idx = item / 3, k = item - 3 * idx;
                        const int4 ad = __ldg(a.m.adj + w.gface[idx]);
                        const int g = k == 0 ? ad.x : (k == 1 ? ad.y : ad.z);
                        if (g < 0) continue;
                        if (hashFind(w.fhKey, maskF, g) >= 0) continue;
                        const int4 c = __ldg(a.m.corner + g);
                        const d3 q0 = ldvert(a.m, c.x), q1 = ldvert(a.m, c.y), q2 = ldvert(a.m, c.z);
                        const bool far = xsqlen(xsub3(sp, q0)) > thr2 && xsqlen(xsub3(sp, q1)) > thr2 && xsqlen(xsub3(sp, q2)) > thr2;
                        if (far) continue;
                        if (w.misc[0] >= cp.maxF) { // full: nobody inserts any more, so a second look tells whether g is really missing
                            if (hashFind(w.fhKey, maskF, g) < 0) w.misc[2] = 1; // (another thread may have added g since the first look:
                            continue;                                           //  the whole-mesh tier fills up to exactly maxF faces)
                        }
                        addFace(g);
                    }
                    head = tail;
                }
                for (int t = tid; t < K; t += CTA_NT) // leftover goal faces (:143-144)
                    if (hashFind(w.fhKey, maskF, w.tGFace[t]) < 0) {
                        if (w.misc[0] >= cp.maxF) {
                            if (hashFind(w.fhKey, maskF, w.tGFace[t]) < 0) w.misc[2] = 1;
                        } else
                            addFace(w.tGFace[t]);
                    }
            }
        }
    }
    __syncthreads();
    const int nF = w.misc[0];
    const int ovfF = w.misc[2];
    __syncthreads(); // (pass 1 below may raise misc[2] again: everybody has read it by now)
    if (ovfF || nF > cp.maxF) return ST_OVF_F;

    // ---------------- 4. local indexing ----------------
    for (int f = tid; f < nF; f += CTA_NT) { // pass 1: create vertex keys
        int4 c = __ldg(a.m.corner + w.gface[f]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int gv = k == 0 ? c.x : (k == 1 ? c.y : c.z);
            bool isNew;
            int slot = hashInsert(w.vhKey, maskV, gv, isNew);
            if (isNew) {
                int id = atomicAdd(&w.misc[1], 1);
                if (id < cp.maxV) {
                    w.gvert[id] = gv;
                    w.vhVal[slot] = (unsigned short)id;
                } else
                    w.misc[2] = 1;
            }
        }
        if (w.misc[2]) break; // hash table might be full next: stop inserting
    }
    __syncthreads();
    if (w.misc[2] || w.misc[1] > cp.maxV) return ST_OVF_V;
    const int nV = w.misc[1];
    for (int v = tid; v < nV; v += CTA_NT) {
        int gv = w.gvert[v];
        d3 p = ldvert(a.m, gv);
        w.vx()[v] = p.x, w.vy()[v] = p.y, w.vz()[v] = p.z;
        w.D()[v] = dinf();
        w.velig[v] = a.m.saddle[gv]; // bit 0: pseudo-source eligible; bit 1: fan being spawned in the current round
        w.vdirty[v] = 0;
    }
    __syncthreads();
    for (int f = tid; f < nF; f += CTA_NT) { // pass 2: local corners + local adjacency + border eligibility
        int gf = w.gface[f];
        int4 c = __ldg(a.m.corner + gf);
        int4 ad = __ldg(a.m.adj + gf);
        unsigned short lv[3], la[3];
        lv[0] = w.vhVal[hashFind(w.vhKey, maskV, c.x)];
        lv[1] = w.vhVal[hashFind(w.vhKey, maskV, c.y)];
        lv[2] = w.vhVal[hashFind(w.vhKey, maskV, c.z)];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int g = k == 0 ? ad.x : (k == 1 ? ad.y : ad.z);
            int s = g < 0 ? -1 : hashFind(w.fhKey, maskF, g);
            la[k] = s < 0 ? (unsigned short)NONE16 : w.fhVal[s];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (la[k] == NONE16) {
                w.velig[lv[(k + 1) % 3]] = 1;
                w.velig[lv[(k + 2) % 3]] = 1;
            }
        w.fvert[4 * f] = lv[0], w.fvert[4 * f + 1] = lv[1], w.fvert[4 * f + 2] = lv[2], w.fvert[4 * f + 3] = (unsigned short)(ad.w & 63);
        w.fadj[4 * f] = la[0], w.fadj[4 * f + 1] = la[1], w.fadj[4 * f + 2] = la[2], w.fadj[4 * f + 3] = 0;
    }
    for (int t = tid; t < K; t += CTA_NT) {
        int s = hashFind(w.fhKey, maskF, w.tGFace[t]);
        w.tFace[t] = w.fhVal[s];
    }
    __syncthreads();
    for (int t = tid; t < K; t += CTA_NT) {
        unsigned short lf = w.tFace[t];
        w.tCorner[4 * t] = w.fvert[4 * lf], w.tCorner[4 * t + 1] = w.fvert[4 * lf + 1], w.tCorner[4 * t + 2] = w.fvert[4 * lf + 2];
        w.fadj[4 * lf + 3] = 1; // the face holds a target: windows entering it answer queries
    }

    // ---------------- 5. window propagation ----------------
    // root frame (source face = local face 0): corner 0 at the origin, corner 1 on +x, corner 2 above
    v2 rq0{0, 0}, rq1, rq2, S2;
    {
        unsigned short c0 = w.fvert[0], c1 = w.fvert[1], c2 = w.fvert[2];
        d3 P0{w.vx()[c0], w.vy()[c0], w.vz()[c0]}, P1{w.vx()[c1], w.vy()[c1], w.vz()[c1]}, P2{w.vx()[c2], w.vy()[c2], w.vz()[c2]};
        d3 e01{P1.x - P0.x, P1.y - P0.y, P1.z - P0.z}, e02{P2.x - P0.x, P2.y - P0.y, P2.z - P0.z};
        double L01 = sqrt(e01.x * e01.x + e01.y * e01.y + e01.z * e01.z);
        d3 ex{e01.x / L01, e01.y / L01, e01.z / L01};
        double x2 = e02.x * ex.x + e02.y * ex.y + e02.z * ex.z;
        d3 ey{e02.x - x2 * ex.x, e02.y - x2 * ex.y, e02.z - x2 * ex.z};
        double y2 = sqrt(ey.x * ey.x + ey.y * ey.y + ey.z * ey.z);
        ey = d3{ey.x / y2, ey.y / y2, ey.z / y2};
        rq1 = v2{L01, 0};
        rq2 = v2{x2, y2};
        double bs = sb0 + sb1 + sb2;
        S2 = v2{(sb1 * rq1.x + sb2 * rq2.x) / bs, (sb2 * rq2.y) / bs};
        if (tid == 0) {
            w.root[0] = ex.x, w.root[1] = ex.y, w.root[2] = ex.z, w.root[3] = ey.x, w.root[4] = ey.y, w.root[5] = ey.z;
        }
        // direct distances to the three corners of the source face
        if (tid < 3) {
            unsigned short cv = tid == 0 ? c0 : (tid == 1 ? c1 : c2);
            d3 P = tid == 0 ? P0 : (tid == 1 ? P1 : P2);
            d3 d{P.x - sp.x, P.y - sp.y, P.z - sp.z};
            double L = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
            w.D()[cv] = L;
            w.dirx()[cv] = d.x / L, w.diry()[cv] = d.y / L, w.dirz()[cv] = d.z / L;
            w.vdirty[cv] = 1;
        }
    }
    __syncthreads();
    auto liftRoot = [&](const v2& d, double& ox, double& oy, double& oz) {
        double rx = d.x * w.root[0] + d.y * w.root[3], ry = d.x * w.root[1] + d.y * w.root[4], rz = d.x * w.root[2] + d.y * w.root[5];
        double L = sqrt(rx * rx + ry * ry + rz * rz);
        ox = rx / L, oy = ry / L, oz = rz / L;
    };
    // targets lying in the source face: the chord
    for (int t = tid; t < K; t += CTA_NT)
        if (w.tFace[t] == 0) {
            d3 d{w.tpx()[t] - sp.x, w.tpy()[t] - sp.y, w.tpz()[t] - sp.z};
            double L = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
            w.tbest()[t] = L;
            w.tsx()[t] = d.x / L, w.tsy()[t] = d.y / L, w.tsz()[t] = d.z / L;
            w.tex()[t] = d.x / L, w.tey()[t] = d.y / L, w.tez()[t] = d.z / L;
        }

    int head = 0, tail = 0; // ring positions (monotone; index = pos & maskR), block-uniform
    { // root windows, in edge order
        const bool v0 = w.fadj[0] != NONE16, v1 = w.fadj[1] != NONE16, v2_ = w.fadj[2] != NONE16;
        if (tid < 3) {
            const int k = tid;
            const unsigned short g = w.fadj[k];
            if (g != NONE16) {
                const int kk = (w.fvert[3] >> (2 * k)) & 3;
                v2 q[3] = {rq0, rq1, rq2};
                const int p = (k == 0 ? 0 : (k == 1 ? (int)v0 : (int)v0 + (int)v1)) & maskR;
                const v2 A = q[(k + 2) % 3], B = q[(k + 1) % 3];
                w.rax()[p] = A.x, w.ray()[p] = A.y, w.rbx()[p] = B.x, w.rby()[p] = B.y, w.rsx()[p] = S2.x, w.rsy()[p] = S2.y;
                w.rt0()[p] = 0, w.rt1()[p] = 1, w.rsg()[p] = 0, w.rmeta[p] = (int)g | (kk << 16), w.rpsv[p] = NONE16;
            }
        }
        tail = (int)v0 + (int)v1 + (int)v2_;
    }
    __syncthreads();

    // one target query of a window: candidate length and the unfolded direction source image -> target
    struct Win {
        v2 A, B, C, S;
        double t0, t1, sg;
        int iA, iB, iC;
    };
    auto queryCand = [&](const Win& q, int t, double& cand, v2& dT) -> bool {
        double bA = q.iA == 0 ? w.tb0()[t] : (q.iA == 1 ? w.tb1()[t] : w.tb2()[t]);
        double bB = q.iB == 0 ? w.tb0()[t] : (q.iB == 1 ? w.tb1()[t] : w.tb2()[t]);
        double bC = q.iC == 0 ? w.tb0()[t] : (q.iC == 1 ? w.tb1()[t] : w.tb2()[t]);
        double bs = bA + bB + bC;
        v2 T{(bA * q.A.x + bB * q.B.x + bC * q.C.x) / bs, (bA * q.A.y + bB * q.B.y + bC * q.C.y) / bs};
        dT = T - q.S;
        double den = cross2(q.B - q.A, dT);
        if (den == 0) return false;
        double mu = cross2(q.S - q.A, dT) / den;
        if (!(mu >= q.t0 - 1e-12 && mu <= q.t1 + 1e-12)) return false;
        cand = q.sg + len2(dT);
        return true;
    };

    double U = dinf();
    unsigned long long nWin = 0, nPs = 0;
    long long tk0 = clock64(), clkFan = 0, clkBatch = 0;
    if (tid == 0) atomicAdd(a.counters + C_CLK_PATCH, (unsigned long long)(tk0 - tp0));
    for (;;) {
        // ================= windows first: drain the ring, CTA_NT windows per pass =================
        while (head != tail) {
            long long tb0 = clock64();
            { // bound U = max_t best[t] (a thread that reads a value lowered in this very pass only prunes more)
                double myMax = 0;
                for (int t = lane; t < K; t += 32) myMax = fmax(myMax, w.tbest()[t]);
                U = warpMax(myMax);
            }
            const double Ub = U * (1 + 1e-12);
            const int nb = min(CTA_NT, tail - head);
            bool active = tid < nb;
            const int p = (head + tid) & maskR;
            head += nb;
            Win q;
            q.A = v2{0, 0}, q.B = v2{1, 0}, q.S = v2{0, -1}, q.C = v2{0, 1};
            q.t0 = 0, q.t1 = 1, q.sg = 0;
            int meta = 0;
            unsigned short psv = NONE16;
            if (active) {
                q.A = v2{w.rax()[p], w.ray()[p]}, q.B = v2{w.rbx()[p], w.rby()[p]}, q.S = v2{w.rsx()[p], w.rsy()[p]};
                q.t0 = w.rt0()[p], q.t1 = w.rt1()[p], q.sg = w.rsg()[p], meta = w.rmeta[p], psv = w.rpsv[p];
            }
            // (the slots popped here are overwritten by this pass's pushes at the earliest: those come after the scan barrier)
            const v2 &A = q.A, &B = q.B, &S = q.S;
            const double t0 = q.t0, t1 = q.t1, sg = q.sg;
            const int g = meta & 0xFFFF, e = (meta >> 16) & 3;
            const v2 AB = B - A;
            const v2 P0 = lerp2(A, B, t0), P1 = lerp2(A, B, t1);
            if (active && sg + asegDist(S, P0, P1) * (1 - 1e-5) > Ub) active = false; // bound may have tightened since the push
            if (active) nWin++;
            const int iA = e == 2 ? 0 : e + 1, iB = e == 0 ? 2 : e - 1, iC = e;
            q.iA = iA, q.iB = iB, q.iC = iC;
            unsigned short vA = 0, vB = 0, vC = 0;
            int kkbits = 0;
            bool holdsTarget = false;
            if (active) {
                vA = w.fvert[4 * g + iA], vB = w.fvert[4 * g + iB], vC = w.fvert[4 * g + iC];
                kkbits = w.fvert[4 * g + 3];
                holdsTarget = w.fadj[4 * g + 3] != 0;
                double PAx = w.vx()[vA], PAy = w.vy()[vA], PAz = w.vz()[vA];
                double abx = w.vx()[vB] - PAx, aby = w.vy()[vB] - PAy, abz = w.vz()[vB] - PAz;
                double acx = w.vx()[vC] - PAx, acy = w.vy()[vC] - PAy, acz = w.vz()[vC] - PAz;
                double r = frcp(abx * abx + aby * aby + abz * abz);
                double crx = aby * acz - abz * acy, cry = abz * acx - abx * acz, crz = abx * acy - aby * acx;
                double cxn = (acx * abx + acy * aby + acz * abz) * r;
                double cyn = fsqrt(crx * crx + cry * cry + crz * crz) * r;
                q.C = v2{A.x + cxn * AB.x - cyn * AB.y, A.y + cxn * AB.y + cyn * AB.x};
            }
            const v2& C = q.C;
            // ---- queries: targets inside the entered face.  Every candidate goes through an atomic minimum; the thread whose
            // minimum is still standing after the barrier (exactly one per improved target) records how its path starts and ends.
            unsigned impMask = 0; // by ordinal of the target among those of this face (ordinals >= 31 share bit 31)
            if (holdsTarget) {
                int ord = 0;
                for (int t = 0; t < K; ++t) {
                    if (w.tFace[t] != g) continue;
                    double cand;
                    v2 dT;
                    if (queryCand(q, t, cand, dT) && cand < w.tbest()[t] && atomicMinD(&w.tbest()[t], cand)) impMask |= 1u << min(ord, 31);
                    ord++;
                }
            }
            if (__syncthreads_or(impMask != 0)) {
                if (impMask) {
                    int ord = 0;
                    for (int t = 0; t < K; ++t) {
                        if (w.tFace[t] != g) continue;
                        const bool mineT = (impMask >> min(ord, 31)) & 1u;
                        ord++;
                        if (!mineT) continue;
                        double cand;
                        v2 dT;
                        if (!queryCand(q, t, cand, dT) || cand != w.tbest()[t]) continue;
                        if (psv == NONE16) liftRoot(dT, w.tsx()[t], w.tsy()[t], w.tsz()[t]);
                        else w.tsx()[t] = w.dirx()[psv], w.tsy()[t] = w.diry()[psv], w.tsz()[t] = w.dirz()[psv];
                        if (a.nbrTe) { // end tangent: dT in the face's (u, u_perp) frame, lifted with the face's 3-D frame
                            double PAx = w.vx()[vA], PAy = w.vy()[vA], PAz = w.vz()[vA];
                            double abx = w.vx()[vB] - PAx, aby = w.vy()[vB] - PAy, abz = w.vz()[vB] - PAz;
                            double acx = w.vx()[vC] - PAx, acy = w.vy()[vC] - PAy, acz = w.vz()[vC] - PAz;
                            double L3 = sqrt(abx * abx + aby * aby + abz * abz);
                            double U3x = abx / L3, U3y = aby / L3, U3z = abz / L3;
                            double cx = acx * U3x + acy * U3y + acz * U3z;
                            double wx = acx - cx * U3x, wy = acy - cx * U3y, wz = acz - cx * U3z;
                            double cy = sqrt(wx * wx + wy * wy + wz * wz);
                            double L2d = len2(AB);
                            double ux = AB.x / L2d, uy = AB.y / L2d;
                            double du = dT.x * ux + dT.y * uy, dw = (-dT.x * uy + dT.y * ux) / cy;
                            double rx = du * U3x + dw * wx, ry = du * U3y + dw * wy, rz = du * U3z + dw * wz;
                            double L = sqrt(rx * rx + ry * ry + rz * rz);
                            w.tex()[t] = rx / L, w.tey()[t] = ry / L, w.tez()[t] = rz / L;
                        }
                    }
                }
            }
            // ---- children
            Child c0, c1;
            c0.valid = c1.valid = 0;
            bool improved = false;
            double dC = 0;
            if (active) {
                v2 dL = P0 - S, dR = P1 - S, dCv = C - S;
                double sideL = cross2(dL, dCv), sideR = cross2(dR, dCv);
                double lc2 = dCv.x * dCv.x + dCv.y * dCv.y;
                float lcf = sqrtf((float)lc2);
                double epsL = 1e-12 * (double)(alen2(dL) * lcf), epsR = 1e-12 * (double)(alen2(dR) * lcf);
                bool inside = !(sideL > epsL) && !(sideR < -epsR);
                double DA = w.D()[vA], DB = w.D()[vB], DC = w.D()[vC];
                if (inside) {
                    dC = sg + fsqrt(lc2);
                    if (dC < DC) {
                        improved = atomicMinD(&w.D()[vC], dC);
                        DC = fmin(DC, dC);
                    }
                }
                // Xin-Wang filter and bound test in float with a conservative margin: a window is dropped only
                // when it is dominated by clearly more than the rounding of the approximation
                const float keep = 1.f - 2e-5f;
                float fsg = (float)sg, fDA = (float)DA, fDB = (float)DB, fDC = (float)DC;
                float fUb = Ub < 1e30 ? (float)Ub * (1.f + 2e-5f) : 3e38f;
                if (!(sideL > epsL)) { // edge C->A of this face (opposite corner B), seen from the neighbour as A->C
                    unsigned short g2 = w.fadj[4 * g + iB];
                    if (g2 != NONE16) {
                        double m0 = hitParam(S, P0, A, C);
                        double m1 = inside ? 1.0 : hitParam(S, P1, A, C);
                        if (m1 - m0 > 1e-13) {
                            v2 XA = lerp2(A, C, m0), XC = lerp2(A, C, m1);
                            float sXA = fsg + adist2(S, XA), sXC = fsg + adist2(S, XC);
                            bool dom = (fDA + adist2(A, XC) < sXC * keep) || (fDC + adist2(C, XA) < sXA * keep) || (fDB + adist2(B, XA) < sXA * keep);
                            if (!dom && fsg + asegDist(S, XA, XC) <= fUb) {
                                c0.valid = 1;
                                c0.meta = (int)g2 | (((kkbits >> (2 * iB)) & 3) << 16);
                                c0.A = A, c0.B = C, c0.t0 = m0, c0.t1 = m1;
                            }
                        }
                    }
                }
                if (!(sideR < -epsR)) { // edge B->C of this face (opposite corner A), seen from the neighbour as C->B
                    unsigned short g2 = w.fadj[4 * g + iA];
                    if (g2 != NONE16) {
                        double m0 = inside ? 0.0 : hitParam(S, P0, C, B);
                        double m1 = hitParam(S, P1, C, B);
                        if (m1 - m0 > 1e-13) {
                            v2 XC = lerp2(C, B, m0), XB = lerp2(C, B, m1);
                            float sXC = fsg + adist2(S, XC), sXB = fsg + adist2(S, XB);
                            bool dom = (fDB + adist2(B, XC) < sXC * keep) || (fDC + adist2(C, XB) < sXB * keep) || (fDA + adist2(A, XB) < sXB * keep);
                            if (!dom && fsg + asegDist(S, XC, XB) <= fUb) {
                                c1.valid = 1;
                                c1.meta = (int)g2 | (((kkbits >> (2 * iA)) & 3) << 16);
                                c1.A = C, c1.B = B, c1.t0 = m0, c1.t1 = m1;
                            }
                        }
                    }
                }
            }
            const int mine = c0.valid + c1.valid;
            double sdx = 0, sdy = 0, sdz = 0; // start direction of the window's family, read before anybody may rewrite it
            if (improved) {
                if (psv == NONE16) liftRoot(C - S, sdx, sdy, sdz);
                else sdx = w.dirx()[psv], sdy = w.diry()[psv], sdz = w.dirz()[psv];
            }
            int tot;
            int pos = tail + blockExclusiveScan(mine, tid, sc, parity, tot); // (barrier: every atomic minimum of the pass is done)
            if (improved && dC == w.D()[vC]) { // the standing minimum writes the start direction carried to this vertex
                w.dirx()[vC] = sdx, w.diry()[vC] = sdy, w.dirz()[vC] = sdz;
                w.vdirty[vC] = 1;
            }
            if (tail + tot - head > cp.ring) return ST_OVF_RING;
            if (c0.valid) {
                int s = pos & maskR;
                w.rax()[s] = c0.A.x, w.ray()[s] = c0.A.y, w.rbx()[s] = c0.B.x, w.rby()[s] = c0.B.y, w.rsx()[s] = S.x, w.rsy()[s] = S.y;
                w.rt0()[s] = c0.t0, w.rt1()[s] = c0.t1, w.rsg()[s] = sg, w.rmeta[s] = c0.meta, w.rpsv[s] = psv;
                pos++;
            }
            if (c1.valid) {
                int s = pos & maskR;
                w.rax()[s] = c1.A.x, w.ray()[s] = c1.A.y, w.rbx()[s] = c1.B.x, w.rby()[s] = c1.B.y, w.rsx()[s] = S.x, w.rsy()[s] = S.y;
                w.rt0()[s] = c1.t0, w.rt1()[s] = c1.t1, w.rsg()[s] = sg, w.rmeta[s] = c1.meta, w.rpsv[s] = psv;
            }
            tail += tot;
            __syncthreads();
            clkBatch += clock64() - tb0;
        }

        // ================= ring empty: vertex candidates, then pseudo-source fans =================
        long long tf0 = clock64();
        // ---- vertex -> target candidates (a path may end with a straight leg from a corner of the target's face)
        for (int t = tid; t < K; t += CTA_NT) {
            double best = w.tbest()[t];
            if (w.tFace[t] != 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    unsigned short cv = w.tCorner[4 * t + k];
                    double dv = w.D()[cv];
                    if (dv < best) {
                        d3 ee{w.tpx()[t] - w.vx()[cv], w.tpy()[t] - w.vy()[cv], w.tpz()[t] - w.vz()[cv]};
                        double L = sqrt(ee.x * ee.x + ee.y * ee.y + ee.z * ee.z);
                        if (dv + L < best) {
                            best = dv + L;
                            w.tbest()[t] = best;
                            w.tsx()[t] = w.dirx()[cv], w.tsy()[t] = w.diry()[cv], w.tsz()[t] = w.dirz()[cv];
                            w.tex()[t] = ee.x / L, w.tey()[t] = ee.y / L, w.tez()[t] = ee.z / L;
                        }
                    }
                }
            }
        }
        __syncthreads();
        {
            double myMax = 0;
            for (int t = lane; t < K; t += 32) myMax = fmax(myMax, w.tbest()[t]);
            U = warpMax(myMax);
        }
        const double Ub = U * (1 + 1e-12);
        // ---- pseudo-source fans.  A vertex v can lie on a shortest path to target t only if
        //      D[v] + |x_v - x_t| (Euclidean lower bound of the remaining leg) beats the best path known to t.
        bool mark = false;
        for (int v = tid; v < nV; v += CTA_NT) {
            bool fl = w.vdirty[v] && (w.velig[v] & 1) && w.D()[v] <= Ub;
            w.vdirty[v] = 0;
            if (fl) {
                bool useful = false;
                double Dv = w.D()[v], px = w.vx()[v], py = w.vy()[v], pz = w.vz()[v];
                for (int t = 0; t < K && !useful; ++t) {
                    double ex = w.tpx()[t] - px, ey = w.tpy()[t] - py, ez = w.tpz()[t] - pz;
                    float lb = sqrtf((float)(ex * ex + ey * ey + ez * ez)) * (1.f - 2e-6f);
                    useful = Dv + (double)lb < w.tbest()[t];
                }
                if (useful) w.velig[v] |= 2, mark = true, nPs++;
            }
        }
        const bool spawned = __syncthreads_or(mark);
        bool overflow = false;
        if (spawned) {
            // one sweep over the (face, corner) pairs: the pairs whose corner is a marked vertex are the fan edges
            const int nItems = 3 * nF;
            for (int i0 = 0; i0 < nItems; i0 += CTA_NT) {
                const int item = i0 + tid;
                const int f = item / 3, i = item - 3 * f;
                bool has = false;
                unsigned short pv = 0, vp = 0, vq = 0;
                if (item < nItems) {
                    pv = w.fvert[4 * f + i];
                    has = (w.velig[pv] & 2) != 0;
                }
                if (!__syncthreads_or(has)) continue;
                Child ch;
                ch.valid = 0;
                double Dv = 0, lp = 0, lq = 0, dvx = 0, dvy = 0, dvz = 0;
                if (has) {
                    unsigned short c0 = w.fvert[4 * f], c1 = w.fvert[4 * f + 1], c2 = w.fvert[4 * f + 2];
                    vp = i == 0 ? c1 : (i == 1 ? c2 : c0), vq = i == 0 ? c2 : (i == 1 ? c0 : c1);
                    Dv = w.D()[pv];
                    dvx = w.dirx()[pv], dvy = w.diry()[pv], dvz = w.dirz()[pv];
                    d3 Pv{w.vx()[pv], w.vy()[pv], w.vz()[pv]};
                    d3 ep{w.vx()[vp] - Pv.x, w.vy()[vp] - Pv.y, w.vz()[vp] - Pv.z}, eq{w.vx()[vq] - Pv.x, w.vy()[vq] - Pv.y, w.vz()[vq] - Pv.z};
                    lp = fsqrt(ep.x * ep.x + ep.y * ep.y + ep.z * ep.z), lq = fsqrt(eq.x * eq.x + eq.y * eq.y + eq.z * eq.z);
                    unsigned short g2 = w.fadj[4 * f + i];
                    if (g2 != NONE16) {
                        double rlp = frcp(lp);
                        double qx = (eq.x * ep.x + eq.y * ep.y + eq.z * ep.z) * rlp;
                        d3 cr{ep.y * eq.z - ep.z * eq.y, ep.z * eq.x - ep.x * eq.z, ep.x * eq.y - ep.y * eq.x};
                        double qy = fsqrt(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z) * rlp;
                        int kk = (w.fvert[4 * f + 3] >> (2 * i)) & 3;
                        ch.valid = 1;
                        ch.meta = (int)g2 | (kk << 16);
                        ch.A = v2{qx, qy};
                        ch.B = v2{lp, 0};
                        ch.t0 = 0, ch.t1 = 1;
                        if (Dv + asegDist(v2{0, 0}, ch.A, ch.B) * (1 - 1e-5) > Ub) ch.valid = 0;
                    }
                }
                __syncthreads(); // every (D, direction) pair of this sweep step is read before any of them moves
                bool ip = false, iq = false;
                if (has) { // edge paths
                    ip = atomicMinD(&w.D()[vp], Dv + lp);
                    iq = atomicMinD(&w.D()[vq], Dv + lq);
                }
                int tot;
                int pos = tail + blockExclusiveScan(ch.valid, tid, sc, parity, tot); // (barrier)
                // vertices improved along an edge inherit the pseudo-source's start direction
                if (ip && w.D()[vp] == Dv + lp) w.dirx()[vp] = dvx, w.diry()[vp] = dvy, w.dirz()[vp] = dvz, w.vdirty[vp] = 1;
                if (iq && w.D()[vq] == Dv + lq) w.dirx()[vq] = dvx, w.diry()[vq] = dvy, w.dirz()[vq] = dvz, w.vdirty[vq] = 1;
                if (tail + tot - head > cp.ring) overflow = true;
                else if (ch.valid) {
                    int s = pos & maskR;
                    w.rax()[s] = ch.A.x, w.ray()[s] = ch.A.y, w.rbx()[s] = ch.B.x, w.rby()[s] = ch.B.y, w.rsx()[s] = 0, w.rsy()[s] = 0;
                    w.rt0()[s] = 0, w.rt1()[s] = 1, w.rsg()[s] = Dv, w.rmeta[s] = ch.meta, w.rpsv[s] = pv;
                }
                if (!overflow) tail += tot;
                __syncthreads();
            }
            for (int v = tid; v < nV; v += CTA_NT) w.velig[v] &= 1;
            __syncthreads();
        }
        clkFan += clock64() - tf0;
        if (overflow) return ST_OVF_RING;
        if (!spawned) break;
    }
    if (nPs) atomicAdd(&cnt[C_PSEUDO], nPs);
    if (tid == 0) {
        atomicAdd(a.counters + C_CLK_BATCH, (unsigned long long)clkBatch);
        atomicAdd(a.counters + C_CLK_FAN, (unsigned long long)clkFan);
        atomicAdd(a.counters + C_CLK_PROP, (unsigned long long)(clock64() - tk0));
    }
    {
        unsigned long long w0 = nWin;
        for (int o = 16; o; o >>= 1) w0 += __shfl_xor_sync(FULL, w0, o);
        if (lane == 0 && w0) atomicAdd(&cnt[C_WINDOWS], w0);
    }

    // ---------------- 6. results ----------------
    unsigned long long nDis = 0;
    for (int t = tid; t < K; t += CTA_NT) {
        double d = w.tbest()[t];
        if (!(d < dinf())) { // unreachable inside the patch
            nDis++;
            if (a.submeshing) { // triangulatedMeshSpace.cpp:198-203
                d = 2.0 * a.maxDist;
                w.tsx()[t] = 0, w.tsy()[t] = 0, w.tsz()[t] = 1;
                w.tex()[t] = 0, w.tey()[t] = 0, w.tez()[t] = 1;
            } else
                d = -1.0;
            w.tbest()[t] = d;
        }
        size_t o = explicitQ ? (size_t)t : (size_t)li * a.kmax + t;
        if (a.nbrIdx && !explicitQ) a.nbrIdx[o] = w.tIdx[t];
        a.nbrDist[o] = d;
        if (a.nbrTs) a.nbrTs[3 * o] = w.tsx()[t], a.nbrTs[3 * o + 1] = w.tsy()[t], a.nbrTs[3 * o + 2] = w.tsz()[t];
        if (a.nbrTe) a.nbrTe[3 * o] = w.tex()[t], a.nbrTe[3 * o + 1] = w.tey()[t], a.nbrTe[3 * o + 2] = w.tez()[t];
    }
    if (nDis) atomicAdd(&cnt[C_DISCONNECTED], nDis);
    __syncthreads();
    if (tid == 0) {
        cnt[C_SOURCES]++;
        cnt[C_QUERIES] += K;
        cnt[C_PATCH_FACES] += nF;
        cnt[C_PATCH_VERTS] += nV;
        if (!explicitQ) {
            a.nbrCount[li] = K;
            if (a.forceMode) { // force::computeForces: accumulate in neighbour order
                d3 f = a.zero ? d3{0, 0, 0} : d3{a.frc[3 * li], a.frc[3 * li + 1], a.frc[3 * li + 2]};
                for (int t = 0; t < K; ++t) {
                    d3 pf = pairForce(a.fp, d3{w.tsx()[t], w.tsy()[t], w.tsz()[t]}, w.tbest()[t]);
                    f.x += pf.x, f.y += pf.y, f.z += pf.z;
                }
                a.frc[3 * li] = f.x, a.frc[3 * li + 1] = f.y, a.frc[3 * li + 2] = f.z;
                if (a.kick != 0.0) { // velocityVerletNVE.cpp:27-28
                    a.vel[3 * li] += a.kick * f.x, a.vel[3 * li + 1] += a.kick * f.y, a.vel[3 * li + 2] += a.kick * f.z;
                }
            }
        }
    }
    return ST_OK;
}

template <bool GLOBAL_WS>
__global__ void __launch_bounds__(CTA_NT, 1) k_geodesic_cta(GeoArgs a, size_t wsBytes)
{
    extern __shared__ __align__(16) char smem[];
    __shared__ CtaScratch sc;
    const int tid = threadIdx.x;
    char* base = GLOBAL_WS ? a.gws + (size_t)blockIdx.x * wsBytes : smem;
    WS w;
    wsLayout(a.caps, base, &w);
    unsigned long long* cnt = w.wcnt;
    PDL_ENTRY();
    if (tid == 0) sc.flag = a.xK < 0 && strideGuardUp(a.counters); // one reader: the flag may go up while the blocks start
    __syncthreads();
    if (sc.flag) return; // neighbour phase waiting for a larger stride (common.cuh); explicit queries are not affected
    if (tid < 16) cnt[tid] = 0;
    __syncthreads();

    const int nSrc = a.srcList ? *a.srcCount : (a.xK >= 0 ? 1 : a.nLocal);
    int parity = 0;
    for (;;) {
        if (tid == 0) sc.src = atomicAdd(a.workCounter, 1);
        __syncthreads();
        const int s = sc.src;
        __syncthreads();
        if (s >= nSrc) break;
        int li = a.srcList ? a.srcList[s] : s;
        long long tc0 = clock64();
        int st = processSourceCta<GLOBAL_WS>(a, w, li, tid, sc, parity);
        __syncthreads();
        if (tid == 0) {
            atomicAdd(a.counters + C_CLK_TOTAL, (unsigned long long)(clock64() - tc0));
            if (st != ST_OK) {
                atomicAdd(a.counters + C_OVF_REASON + st - 1, 1ull);
                if (a.lastTier) {
                    cnt[C_OVERFLOW]++;
                    if (a.xK < 0) a.nbrCount[li] = 0;
                } else {
                    int r = atomicAdd(a.retryCount, 1);
                    a.retryList[r] = li;
                    cnt[C_TIER_RETRY]++;
                }
            }
        }
    }
    __syncthreads();
    if (tid < 16 && cnt[tid]) atomicAdd(a.counters + tid, cnt[tid]);
}

int geodesicMaxSmemPerBlock() { return 227 * 1024; }

// block-cooperative tier: one CTA per source; workspace in shared memory (a.gws == nullptr, one block per SM) or in global memory
// (block b uses a.gws + b * geoWorkspaceBytes(a.caps))
cudaError_t launchGeodesicCta(cudaStream_t st, const GeoArgs& a, int blocks)
{
    size_t wsBytes = geoWorkspaceBytes(a.caps);
    if (a.gws) return launchStep(k_geodesic_cta<true>, blocks, CTA_NT, 0, st, a, wsBytes);
    if (wsBytes + sizeof(CtaScratch) > (size_t)geodesicMaxSmemPerBlock()) return cudaErrorInvalidConfiguration;
    cudaFuncSetAttribute(k_geodesic_cta<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsBytes);
    return launchStep(k_geodesic_cta<false>, blocks, CTA_NT, wsBytes, st, a, wsBytes);
}

} // namespace css
