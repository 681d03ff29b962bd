// Device helpers shared by the window-propagation kernels (window_kernel.cu: one warp per source; window_half_kernel.cu:
// two sources per warp).  Everything is templated on the per-source shared-memory workspace W.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace css {

#define FULL 0xffffffffu

namespace {

constexpr unsigned char NOPSV = 255;

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000LL); }

struct v2 {
    double x, y;
};
__device__ __forceinline__ v2 operator-(const v2& a, const v2& b) { return v2{a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ v2 lerp2(const v2& a, const v2& b, double t) { return v2{fma(t, b.x - a.x, a.x), fma(t, b.y - a.y, a.y)}; }
__device__ __forceinline__ double cross2(const v2& a, const v2& b) { return a.x * b.y - a.y * b.x; }

__device__ __forceinline__ bool atomicMinD(double* addr, double v)
{ // non-negative doubles order like their bit patterns
    unsigned long long nv = (unsigned long long)__double_as_longlong(v);
    unsigned long long old = atomicMin(reinterpret_cast<unsigned long long*>(addr), nv);
    return nv < old;
}

// ~1 ulp reciprocal / square root (hardware seed + Newton); window geometry only
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
    return fma(fma(-s, s, x), 0.5 * y, s);
}
__device__ __forceinline__ double hitParam(const v2& S, const v2& P, const v2& X, const v2& Y)
{ // ray S->P against X + mu (Y - X), clamped to the segment
    v2 d = P - S;
    double den = cross2(Y - X, d);
    double mu = cross2(S - X, d) * frcp(den);
    if (!(mu == mu)) mu = 0.5;
    return fmin(1.0, fmax(0.0, mu));
}
// fp32 geometry for pruning decisions (always used with a conservative margin; approximate sqrt is plenty)
struct f2 {
    float x, y;
};
__device__ __forceinline__ f2 tof2(const v2& a) { return f2{(float)a.x, (float)a.y}; }
__device__ __forceinline__ float asqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float flen(float x, float y) { return asqrt(fmaf(x, x, y * y)); }
__device__ __forceinline__ float fdist(const f2& a, const f2& b) { return flen(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ f2 flerp(const f2& a, const f2& b, float t) { return f2{fmaf(t, b.x - a.x, a.x), fmaf(t, b.y - a.y, a.y)}; }
__device__ __forceinline__ float fsegDist(const f2& S, const f2& X0, const f2& X1)
{
    float ex = X1.x - X0.x, ey = X1.y - X0.y, sx = S.x - X0.x, sy = S.y - X0.y;
    float L2 = fmaf(ex, ex, ey * ey);
    float s = L2 > 0.f ? __fdividef(fmaf(sx, ex, sy * ey), L2) : 0.f;
    s = fminf(1.f, fmaxf(0.f, s));
    return flen(fmaf(-s, ex, sx), fmaf(-s, ey, sy));
}
// upper bound U = max_t best[t] in fp32, rounded up (non-negative floats order like their bit patterns; +inf stays +inf)
template <class W> __device__ __forceinline__ float warpBound(const W& w, int lane, int K)
{
    unsigned u = lane < K ? __float_as_uint(__double2float_ru(w.tbest[lane])) : 0u;
    return __uint_as_float(__reduce_max_sync(FULL, u)) * (1.f + 2e-5f);
}

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

template <class W> __device__ __forceinline__ d3 vpos(const MeshDev& m, const W& w, int v)
{
    if constexpr (W::lean) return ldvert(m, w.gvert[v]);
    else return d3{w.vx[v], w.vy[v], w.vz[v]};
}
template <class W> __device__ __forceinline__ double2 edgeFrame(const MeshDev& m, const W& w, int g, int e)
{
    if constexpr (W::lean) return __ldg(m.geo + 3 * (size_t)w.gface[g] + e);
    else return w.geo[3 * g + e];
}

// push up to one window per lane; returns false when the ring would overflow
template <class W> __device__ __forceinline__ bool pushWindows(W& w, int lane, int head, int& tail, bool valid, const v2& A, const v2& B,
                                            double t0, double t1, int meta, unsigned char psv, const double2& cg, float lb = 0.f)
{
    unsigned bal = __ballot_sync(FULL, valid);
    int tot = __popc(bal);
    if (tail + tot - head > W::R) return false;
    if (valid) {
        int q = (tail + __popc(bal & ((1u << lane) - 1))) & (W::R - 1);
        w.rax[q] = A.x, w.ray[q] = A.y, w.rbx[q] = B.x, w.rby[q] = B.y;
        w.rt0[q] = t0, w.rt1[q] = t1, w.rmeta[q] = meta, w.rpsv[q] = psv;
        if constexpr (W::lean) w.rcg[q] = cg;
        if constexpr (W::ringLb) w.rlb[q] = lb; // lower bound of every path through the window (checked again at pop time)
    }
    tail += tot;
    return true;
}

// pseudo-source fan of vertex pv (rare: kept out of line to keep the propagation loop compact)
template <class W> __device__ __noinline__ bool spawnFan(const MeshDev& m, W& w, int lane, int nF, int pv, float fUb, int head, int& tail)
{
    const double Dv = w.D[pv];
    const d3 Pv = vpos(m, w, pv);
    const double dvx = w.dirx[pv], dvy = w.diry[pv];
    bool ok = true;
    for (int f0 = 0; f0 < nF; f0 += 32) {
        int f = f0 + lane;
        bool valid = false;
        v2 A{0, 0}, B{0, 0};
        int meta = 0;
        float lb = 0.f;
        if (f < nF) {
            uchar4 fv = w.fvert[f];
            int i = fv.x == pv ? 0 : (fv.y == pv ? 1 : (fv.z == pv ? 2 : -1));
            if (i >= 0) {
                int vp = i == 0 ? fv.y : (i == 1 ? fv.z : fv.x), vq = i == 0 ? fv.z : (i == 1 ? fv.x : fv.y);
                const d3 Pp = vpos(m, w, vp), Pq = vpos(m, w, vq);
                d3 ep{Pp.x - Pv.x, Pp.y - Pv.y, Pp.z - Pv.z}, eq{Pq.x - Pv.x, Pq.y - Pv.y, Pq.z - Pv.z};
                double lp = fsqrt(ep.x * ep.x + ep.y * ep.y + ep.z * ep.z), lq = fsqrt(eq.x * eq.x + eq.y * eq.y + eq.z * eq.z);
                if (atomicMinD(&w.D[vp], Dv + lp)) w.vdirty[vp] = 2; // edge paths
                if (atomicMinD(&w.D[vq], Dv + lq)) w.vdirty[vq] = 2;
                uchar4 fa = w.fadj[f];
                int g2 = i == 0 ? fa.x : (i == 1 ? fa.y : fa.z);
                if (g2 != REC_NONE) {
                    double rlp = frcp(lp);
                    double qx = (eq.x * ep.x + eq.y * ep.y + eq.z * ep.z) * rlp;
                    d3 cr{ep.y * eq.z - ep.z * eq.y, ep.z * eq.x - ep.x * eq.z, ep.x * eq.y - ep.y * eq.x};
                    double qy = fsqrt(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z) * rlp;
                    int kk = (fv.w >> (2 * i)) & 3;
                    meta = g2 | (kk << 16);
                    A = v2{qx, qy};
                    B = v2{lp, 0};
                    lb = (float)Dv + fsegDist(f2{0.f, 0.f}, tof2(A), tof2(B)) * (1.f - 1e-5f);
                    valid = !(lb > fUb);
                }
            }
        }
        __syncwarp();
        if (f < nF) { // vertices improved along an edge inherit the pseudo-source's start direction
            uchar4 fv = w.fvert[f];
            if (w.vdirty[fv.x] == 2) w.dirx[fv.x] = dvx, w.diry[fv.x] = dvy;
            if (w.vdirty[fv.y] == 2) w.dirx[fv.y] = dvx, w.diry[fv.y] = dvy;
            if (w.vdirty[fv.z] == 2) w.dirx[fv.z] = dvx, w.diry[fv.z] = dvy;
        }
        __syncwarp();
        if (f < nF) {
            uchar4 fv = w.fvert[f];
            if (w.vdirty[fv.x] == 2) w.vdirty[fv.x] = 1;
            if (w.vdirty[fv.y] == 2) w.vdirty[fv.y] = 1;
            if (w.vdirty[fv.z] == 2) w.vdirty[fv.z] = 1;
        }
        double2 cg{0, 0};
        if constexpr (W::lean)
            if (valid) cg = edgeFrame(m, w, meta & 0xFF, (meta >> 16) & 3);
        if (ok) ok = pushWindows(w, lane, head, tail, valid, A, B, 0.0, 1.0, meta, (unsigned char)pv, cg, lb);
        __syncwarp();
    }
    return ok;
}

// 3-D unit end tangent of a path that enters face g through edge e with direction (du, dw) in that edge's frame
template <class W> __device__ __noinline__ d3 liftEnd(const MeshDev& m, const W& w, int g, int e, double du, double dw)
{
    uchar4 fv = w.fvert[g];
    int c0 = fv.x, c1 = fv.y, c2 = fv.z;
    int vA = e == 0 ? c1 : (e == 1 ? c2 : c0), vB = e == 0 ? c2 : (e == 1 ? c0 : c1), vC = e == 0 ? c0 : (e == 1 ? c1 : c2);
    const d3 PA = vpos(m, w, vA), PB = vpos(m, w, vB), PC = vpos(m, w, vC);
    double abx = PB.x - PA.x, aby = PB.y - PA.y, abz = PB.z - PA.z;
    double acx = PC.x - PA.x, acy = PC.y - PA.y, acz = PC.z - PA.z;
    double L3 = sqrt(abx * abx + aby * aby + abz * abz);
    double Ux = abx / L3, Uy = aby / L3, Uz = abz / L3;
    double cx = acx * Ux + acy * Uy + acz * Uz;
    double wx = acx - cx * Ux, wy = acy - cx * Uy, wz = acz - cx * Uz;
    double cy = sqrt(wx * wx + wy * wy + wz * wz);
    double dwn = dw / cy;
    double rx = du * Ux + dwn * wx, ry = du * Uy + dwn * wy, rz = du * Uz + dwn * wz;
    double L = sqrt(rx * rx + ry * ry + rz * rz);
    return d3{rx / L, ry / L, rz / L};
}

enum { WS_OK = 0, WS_RING = 1, WS_GROUPS = 2 };

} // namespace

} // namespace css
