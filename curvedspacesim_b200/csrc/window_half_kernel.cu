// Stage 2 of the many-source geodesic path, TWO SOURCES PER WARP: each half-warp (16 lanes) owns one source particle
// and runs the same exact window propagation as window_kernel.cu (Chen-Han unfolding + Xin-Wang filter, pseudo-source
// fans, target queries, tangent lifts, pair forces in neighbour order + the second velocity-Verlet half kick).
//
// Replaces CGAL::Surface_mesh_shortest_path as the reference uses it per source
// (src/models/triangulatedMeshSpace.cpp:189-203, src/utility/meshUtilities.cpp:360-380) and force::computeForces
// (src/forces/baseForce.cpp:12-28) with velocityVerletNVE's second half kick (src/updaters/velocityVerletNVE.cpp:27-28).
//
// Why half-warps: the one-warp-per-source kernel is latency-bound (5 warps per scheduler, one instruction issued per warp
// every ~9 cycles, 15 of 32 lanes active) because a BFS level of these patches holds only ~5 windows.  The two halves of a
// warp execute ONE instruction stream in lockstep -- every loop runs while either half has work, every branch that contains
// a warp collective is warp-uniform -- so an instruction advances two sources, and the registers (the occupancy limiter next
// to shared memory) are shared by two sources.  A pass propagates 8 windows per source (a lane pair per window, one child
// edge each).  The per-source workspace is slimmed (8 targets at a time, 32-window ring) so that twice as many sources are
// resident per SM.  A source with more than MAXK candidates is propagated once per group of MAXK targets in consecutive
// rounds of the same half; ring overflows go to the retry tier like everywhere else.  Pseudo-source fans (rare) are spawned
// by the whole warp for one source at a time.
//
// Results are order-independent minima with the lowest lane winning ties, so runs are bitwise repeatable and the values
// agree with the one-warp kernel and the oracle to round-off (~1e-15).
#include "window_common.cuh"
#include <cuda_pipeline.h>

namespace css {

namespace {

template <int K> struct MaskOf {
    using type = unsigned short;
};
template <> struct MaskOf<8> {
    using type = unsigned char;
};

template <class T> struct HalfSmem { // per-SOURCE workspace, two per warp (always the lean layout: frames / vertices from L2)
    static constexpr int F = T::MAXF, V = T::MAXV, K = T::MAXK, R = T::RING;
    static constexpr bool lean = true;
    static constexpr bool ringLb = true;
    using mask_t = typename MaskOf<(K <= 8 ? 8 : 16)>::type;
    static_assert(T::OFF_VELIG % 4 == 0 && V % 4 == 0, "velig is copied in 4-byte pieces");
    static_assert(K <= 16 && R >= 32, "one target per lane of the half; a pass pushes up to 16 children");
    int gface[F], gvert[V];
    double D[V + 1], dirx[V], diry[V]; // D[V] = 0: the sigma of the real source
    double rax[R], ray[R], rbx[R], rby[R], rt0[R], rt1[R];
    double2 rcg[R];
    double tbest[K], tb0[K], tb1[K], tb2[K], tsx[K], tsy[K], tdu[K], tdw[K];
    double tpx[K], tpy[K], tpz[K], tcd0[K], tcd1[K], tcd2[K];
    double root[6];
    double fpart[3]; // pair forces of the earlier target groups of this source
    int rmeta[R];
    float rlb[R]; // ring: lower bound (fp32, rounded down by the margin) of every path through the window, computed at push time
    int tIdx[K];
    int tcode[K]; // how the best path ends: 0 none, 1 chord in the source face, 2 + 4*(g | e << 8) window, 3 + 4*k corner k
    int towner[K];
    unsigned wcnt[16];
    float ub; // pruning bound: max over the targets of the best distance known, fp32 rounded up
    alignas(16) uchar4 fvert[F];
    uchar4 fadj[F];
    alignas(4) mask_t tmask[F]; // targets lying in each face (bit t)
    alignas(4) unsigned char tFace[K];
    alignas(4) unsigned char velig[V];
    unsigned char vdirty[V];
    unsigned char rpsv[R];
};

// refresh the pruning bound of both sources of the warp after target distances changed: max over the half's targets of the
// best distance known, fp32 rounded up (non-negative floats order like their bit patterns; +inf stays +inf).  Called by the
// whole warp from warp-uniform code; the xor offsets stay inside a half.
template <class W> __device__ __forceinline__ void updateBound(W& w, int hl, int K)
{
    unsigned u = hl < K ? __float_as_uint(__double2float_ru(w.tbest[hl])) : 0u;
#pragma unroll
    for (int o = 8; o; o >>= 1) u = max(u, __shfl_xor_sync(FULL, u, o));
    if (hl == 0) w.ub = __uint_as_float(u) * (1.f + 2e-5f);
}
// parameter on X + mu (Y - X) hit by the ray from the origin through P, clamped to the segment
__device__ __forceinline__ double hitParam0(const v2& P, const v2& X, const v2& Y)
{
    double den = cross2(Y - X, P);
    double mu = -cross2(X, P) * frcp(den);
    mu = mu == mu ? mu : 0.5; // plain compares and selects: fmin / fmax carry IEEE NaN handling that costs ~8 instructions each
    mu = mu < 0.0 ? 0.0 : mu;
    return mu > 1.0 ? 1.0 : mu;
}

// push up to one window per lane into the ring of the lane's own half; false when that ring would overflow
template <class W> __device__ __forceinline__ bool pushHalf(W& w, int hl, int hbase, int head, int& tail, bool valid, const v2& A, const v2& B,
                                                           double t0, double t1, int meta, unsigned char psv, const double2& cg, float lb)
{
    const unsigned bal = (__ballot_sync(FULL, valid) >> hbase) & 0xFFFFu;
    const int tot = __popc(bal);
    if (tail + tot - head > W::R) return false;
    if (valid) {
        int q = (tail + __popc(bal & ((1u << hl) - 1))) & (W::R - 1);
        w.rax[q] = A.x, w.ray[q] = A.y, w.rbx[q] = B.x, w.rby[q] = B.y;
        w.rt0[q] = t0, w.rt1[q] = t1, w.rmeta[q] = meta, w.rpsv[q] = psv;
        w.rcg[q] = cg;
        w.rlb[q] = lb;
    }
    tail += tot;
    return true;
}

} // namespace

constexpr int SPILL_CAP = 128, SPILL_DOUBLES = 9; // windows per source in the global spill stack; doubles per window

#ifndef CSS_HALF_WPB
#define CSS_HALF_WPB 2
#endif
#ifndef CSS_HALF_MINBLOCKS
#define CSS_HALF_MINBLOCKS 8
#endif

template <class T> __global__ void __launch_bounds__(32 * CSS_HALF_WPB, CSS_HALF_MINBLOCKS) k_windows_half(const __grid_constant__ WinArgs a)
{
    using W = HalfSmem<T>;
    constexpr int MASKR = T::RING - 1;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int hl = lane & 15, hbase = lane & 16, half = lane >> 4;
    const unsigned hmask = 0xFFFFu << hbase;
    W* const wpair = reinterpret_cast<W*>(smemRaw) + 2 * wib;
    W& w = wpair[half];
    PDL_ENTRY();
    if (strideGuardUp(a.counters)) return; // stage 1 did not run: the host regrows the neighbour stride and repeats the phase
    // spill stacks of this warp's two sources (global memory, L2-resident; see the push in the pass loop)
    double* const spillPair = a.spill + (size_t)(blockIdx.x * (blockDim.x >> 5) + wib) * 2 * SPILL_CAP * SPILL_DOUBLES;
    double* const spill = spillPair + (size_t)half * SPILL_CAP * SPILL_DOUBLES;
    w.wcnt[hl] = 0;
    __syncwarp();
    const int nWork = a.srcList ? min(*a.srcCount, a.maxRecords) : a.nLocal;

    // per-half round state (uniform inside a half)
    int s = 0, li = 0, base = 0, nFr = 0, nVr = 0, Ktot = 0;
    bool more = false, done = false;
    for (;;) {
        // ---------------- next piece of work for each half: the next target group of its source, or a new source ----------------
        if (more) base += T::MAXK;
        else if (!done) {
            base = 0;
            for (;;) {
                if (hl == 0) s = atomicAdd(a.workCounter, 1);
                s = __shfl_sync(hmask, s, hbase);
                if (s >= nWork) {
                    done = true;
                    break;
                }
                const int4 hdr = *reinterpret_cast<const int4*>(a.records + (size_t)s * T::BYTES);
                li = a.srcList ? a.srcList[s] : s;
                if (hdr.w) continue; // overflowed in stage 1: the retry tiers own this source
                if (hdr.z == 0) {    // no candidates: nothing to propagate
                    if (hl == 0) {
                        w.wcnt[C_SOURCES]++;
                        a.nbrCount[li] = 0;
                        if (a.forceMode) {
                            d3 f = a.zero ? d3{0, 0, 0} : d3{a.frc[3 * li], a.frc[3 * li + 1], a.frc[3 * li + 2]};
                            a.frc[3 * li] = f.x, a.frc[3 * li + 1] = f.y, a.frc[3 * li + 2] = f.z;
                            if (a.kick != 0.0) a.vel[3 * li] += a.kick * f.x, a.vel[3 * li + 1] += a.kick * f.y, a.vel[3 * li + 2] += a.kick * f.z;
                        }
                    }
                    continue;
                }
                nFr = hdr.x, nVr = hdr.y, Ktot = hdr.z;
                break;
            }
        }
        __syncwarp();
        if (__all_sync(FULL, done)) break;
        bool live = !done; // this half propagates a source in this round
        int nF = live ? nFr : 0, nV = live ? nVr : 0, K = live ? min(T::MAXK, Ktot - base) : 0;
        const unsigned char* rec = a.records + (size_t)(live ? s : 0) * T::BYTES;
        const int gi = a.minIdx + li;

        // ---------------- stage the patch ----------------
        d3 sp{0, 0, 0};
        double sb0 = 1, sb1 = 0, sb2 = 0;
        if (live) {
            // the record sections go global -> shared as asynchronous copies (cp.async): all of them are in flight at once and
            // overlap the dependent target fetches below, instead of one L2 round trip per section
            const int4* src = reinterpret_cast<const int4*>(rec + T::OFF_FVERT); // fvert | fadj are contiguous in the record and here
            int4* dst = reinterpret_cast<int4*>(w.fvert);
            for (int q = hl; q * 4 < nF; q += 16) {
                __pipeline_memcpy_async(dst + q, src + q, 16);
                __pipeline_memcpy_async(dst + q + T::MAXF / 4, src + q + T::MAXF / 4, 16);
            }
            const int* gface = reinterpret_cast<const int*>(rec + T::OFF_GFACE);
            for (int f = hl; f < nF; f += 16) __pipeline_memcpy_async(w.gface + f, gface + f, 4);
            const int* gvert = reinterpret_cast<const int*>(rec + T::OFF_GVERT);
            for (int v = hl; v < nV; v += 16) __pipeline_memcpy_async(w.gvert + v, gvert + v, 4);
            for (int q = hl; q * 4 < nV; q += 16) __pipeline_memcpy_async(w.velig + 4 * q, rec + T::OFF_VELIG + 4 * q, 4);
            __pipeline_commit();
            unsigned* tmw = reinterpret_cast<unsigned*>(w.tmask);
            for (int q = hl; q * 4 < nF * (int)sizeof(typename W::mask_t); q += 16) tmw[q] = 0;
            for (int v = hl; v < nV; v += 16) w.D[v] = dinf(), w.vdirty[v] = 0;
            if (hl < K) {
                int j = reinterpret_cast<const int*>(rec + T::OFF_TIDX)[base + hl];
                w.tIdx[hl] = j;
                w.tFace[hl] = rec[T::OFF_TFACE + base + hl];
                w.tb0[hl] = a.bary[3 * j], w.tb1[hl] = a.bary[3 * j + 1], w.tb2[hl] = a.bary[3 * j + 2];
                w.tpx[hl] = a.eucl[3 * j], w.tpy[hl] = a.eucl[3 * j + 1], w.tpz[hl] = a.eucl[3 * j + 2];
                w.tbest[hl] = dinf();
                w.tcode[hl] = 0;
                w.tsx[hl] = 0, w.tsy[hl] = 0;
                w.towner[hl] = 32;
            }
            sp = d3{a.eucl[3 * gi], a.eucl[3 * gi + 1], a.eucl[3 * gi + 2]};
            sb0 = a.bary[3 * gi], sb1 = a.bary[3 * gi + 1], sb2 = a.bary[3 * gi + 2];
            __pipeline_wait_prior(0);
        }
        __syncwarp();

        // ---------------- root frame, direct legs ----------------
        // source face = local face 0: corner 0 at the origin, corner 1 on +x, corner 2 above
        v2 rq0{0, 0}, rq1{1, 0}, rq2{0, 1};
        if (live) {
            uchar4 fv = w.fvert[0];
            const d3 P0 = vpos(a.m, w, fv.x), P1 = vpos(a.m, w, fv.y), P2 = vpos(a.m, w, fv.z);
            d3 e01{P1.x - P0.x, P1.y - P0.y, P1.z - P0.z}, e02{P2.x - P0.x, P2.y - P0.y, P2.z - P0.z};
            double L01 = sqrt(e01.x * e01.x + e01.y * e01.y + e01.z * e01.z);
            double rL = 1.0 / L01;
            d3 ex{e01.x * rL, e01.y * rL, e01.z * rL};
            double x2 = e02.x * ex.x + e02.y * ex.y + e02.z * ex.z;
            d3 ey{e02.x - x2 * ex.x, e02.y - x2 * ex.y, e02.z - x2 * ex.z};
            double y2 = sqrt(ey.x * ey.x + ey.y * ey.y + ey.z * ey.z);
            double ry = 1.0 / y2;
            ey = d3{ey.x * ry, ey.y * ry, ey.z * ry};
            rq1 = v2{L01, 0};
            rq2 = v2{x2, y2};
            double rbs = 1.0 / (sb0 + sb1 + sb2);
            const v2 S2{(sb1 * rq1.x + sb2 * rq2.x) * rbs, (sb2 * rq2.y) * rbs};
            // translate: the source becomes the origin of the root frame
            rq0 = v2{-S2.x, -S2.y};
            rq1 = rq1 - S2, rq2 = rq2 - S2;
            if (hl == 0) w.D[W::V] = 0.0;
            if (hl == 0) w.root[0] = ex.x, w.root[1] = ex.y, w.root[2] = ex.z, w.root[3] = ey.x, w.root[4] = ey.y, w.root[5] = ey.z;
            if (hl < 3) { // straight legs to the three corners of the source face
                int cv = hl == 0 ? fv.x : (hl == 1 ? fv.y : fv.z);
                v2 q = hl == 0 ? rq0 : (hl == 1 ? rq1 : rq2);
                d3 P = hl == 0 ? P0 : (hl == 1 ? P1 : P2);
                d3 d{P.x - sp.x, P.y - sp.y, P.z - sp.z};
                w.D[cv] = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
                w.dirx[cv] = q.x, w.diry[cv] = q.y;
                w.vdirty[cv] = 1;
            }
            if (hl < K) {
                int t = hl, lf = w.tFace[t];
                if (lf == 0) { // target in the source face: the chord
                    d3 d{w.tpx[t] - sp.x, w.tpy[t] - sp.y, w.tpz[t] - sp.z};
                    w.tbest[t] = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
                    w.tcode[t] = 1;
                } else {
                    constexpr int MB = (int)sizeof(typename W::mask_t);
                    atomicOr(reinterpret_cast<unsigned*>(w.tmask) + (lf * MB) / 4, (1u << t) << (8 * ((lf * MB) & 3)));
                    uchar4 tv = w.fvert[lf];
                    double px = w.tpx[t], py = w.tpy[t], pz = w.tpz[t];
                    const d3 Q0 = vpos(a.m, w, tv.x), Q1 = vpos(a.m, w, tv.y), Q2 = vpos(a.m, w, tv.z);
                    double ax = px - Q0.x, ay = py - Q0.y, az = pz - Q0.z;
                    double bx = px - Q1.x, by = py - Q1.y, bz = pz - Q1.z;
                    double cx = px - Q2.x, cy = py - Q2.y, cz = pz - Q2.z;
                    w.tcd0[t] = sqrt(ax * ax + ay * ay + az * az);
                    w.tcd1[t] = sqrt(bx * bx + by * by + bz * bz);
                    w.tcd2[t] = sqrt(cx * cx + cy * cy + cz * cz);
                }
            }
        }
        int head = 0, tail = 0, spillN = 0;
        bool failed = false;
        {
            bool valid = false;
            v2 A{0, 0}, B{0, 0};
            int meta = 0;
            double2 cg{0, 0};
            float lb0 = 0.f;
            if (live && hl < 3) {
                uchar4 fa = w.fadj[0];
                int g = hl == 0 ? fa.x : (hl == 1 ? fa.y : fa.z);
                if (g != REC_NONE) {
                    int kk = (w.fvert[0].w >> (2 * hl)) & 3;
                    valid = true;
                    meta = g | (kk << 16);
                    // edge k runs corner k+1 -> corner k+2; the neighbour sees it reversed
                    A = hl == 0 ? rq2 : (hl == 1 ? rq0 : rq1);
                    B = hl == 0 ? rq1 : (hl == 1 ? rq2 : rq0);
                    cg = edgeFrame(a.m, w, g, kk);
                    lb0 = fsegDist(f2{0.f, 0.f}, tof2(A), tof2(B)) * (1.f - 1e-5f);
                }
            }
            pushHalf(w, hl, hbase, head, tail, valid, A, B, 0.0, 1.0, meta, NOPSV, cg, lb0); // 3 <= ring
        }
        __syncwarp();
        updateBound(w, hl, K);
        __syncwarp();

        unsigned nWin = 0, nPs = 0;
#ifdef CSS_PASS_STATS
        unsigned nPass = 0, nPassIdle = 0, nPass4 = 0, nPop = 0, nOuter = 0; // per WARP pass (lane 0 reports)
#endif
        for (;;) {
#ifdef CSS_PASS_STATS
            nOuter++;
#endif
            // ========== drain both rings: 16 windows per pass, shared between the two sources on demand; a PAIR of lanes per
            // window, one child edge each.  Each source is guaranteed 8 slots; the slots one source leaves free go to the other,
            // so a pass is full whenever the two rings together hold 16 windows (a BFS level of one patch rarely does).
            for (;;) {
                // a source whose ring ran dry takes spilled windows back (see the push below); each half moves its own
                if (__builtin_expect(spillN > 0 && head == tail, 0)) { // (rare paths carry branch hints: the compiler moves them out of the loop body, which is instruction-fetch sensitive)
                    const int n = min(spillN, 16);
                    if (hl < n) {
                        const double* e = spill + (size_t)(spillN - 1 - hl) * SPILL_DOUBLES;
                        const int q = (tail + hl) & MASKR;
                        w.rax[q] = e[0], w.ray[q] = e[1], w.rbx[q] = e[2], w.rby[q] = e[3], w.rt0[q] = e[4], w.rt1[q] = e[5];
                        w.rcg[q] = double2{e[6], e[7]};
                        const long long mp = __double_as_longlong(e[8]);
                        const int psvr = (int)(mp >> 32) & 0xFF;
                        w.rmeta[q] = (int)(mp & 0xFFFFFFFFll), w.rpsv[q] = (unsigned char)psvr;
                        const v2 Ar{e[0], e[1]}, Br{e[2], e[3]};
                        w.rlb[q] = (float)w.D[min(psvr, W::V)] + fsegDist(f2{0.f, 0.f}, tof2(lerp2(Ar, Br, e[4])), tof2(lerp2(Ar, Br, e[5]))) * (1.f - 1e-5f);
                    }
                    tail += n, spillN -= n;
                }
                __syncwarp();
                const int head0 = __shfl_sync(FULL, head, 0), tail0 = __shfl_sync(FULL, tail, 0);
                const int head1 = __shfl_sync(FULL, head, 16), tail1 = __shfl_sync(FULL, tail, 16);
                const int avail0 = tail0 - head0, avail1 = tail1 - head1;
                if (avail0 + avail1 == 0) break; // (a spill stack is empty whenever its ring is, after the refill above)
                const int nb0 = min(avail0, max(8, 16 - avail1)), nb1 = min(avail1, 16 - nb0);
                const int slot = lane >> 1;
                const int src = slot >= nb0; // the source this lane pair works for in this pass
                const int k = src ? slot - nb0 : slot;
                bool active = src ? k < nb1 : true;
                W& wp = wpair[src];
                const float fUb = wp.ub; // max over the source's targets of the best distance so far (kept by updateBound)
                const int j = lane & 1; // which child edge this lane propagates into
#ifdef CSS_PASS_STATS
                nPass++, nPassIdle += (nb0 == 0 || nb1 == 0), nPass4 += nb0 + nb1 <= 8, nPop += nb0 + nb1;
#endif
                const int p = ((src ? head1 : head0) + k) & MASKR;
                head += half ? nb1 : nb0;
                // ---- pop.  Every lane reads its slot (the index is always inside the ring; lanes without a window compute on
                // stale data and are masked at every side effect): straight-line code instead of divergent regions.
                // Window families live in the frame of their (pseudo-)source, which sits at the ORIGIN of that frame (the root
                // frame is translated so that the real source is at the origin too); sigma = D[pseudo-source], D[V] = 0 stands
                // for the real source.
                const v2 A{wp.rax[p], wp.ray[p]}, B{wp.rbx[p], wp.rby[p]};
                const double t0 = wp.rt0[p], t1 = wp.rt1[p];
                const double2 cg = wp.rcg[p];
                const int meta = active ? wp.rmeta[p] : 0;
                const unsigned char psv = wp.rpsv[p];
                const double sg = wp.D[min((int)psv, W::V)];
                const int g = meta & 0xFF, e = (meta >> 16) & 3;
                const v2 AB = B - A;
                const v2 P0 = lerp2(A, B, t0), P1 = lerp2(A, B, t1);
                const float fsg = (float)sg;
                const f2 fO{0.f, 0.f};
                // The bound may have tightened since the push: the lower bound computed then is checked again.  (Should sigma have
                // dropped since, the stored bound is too high; the vertex is dirty in that case and its fan is spawned afresh.)
                if (wp.rlb[p] > fUb) active = false;
                // ---- unfold the entered face: apex C from the edge frame; corners / neighbours / edge indices rotated by e
                const unsigned fvw = *reinterpret_cast<const unsigned*>(&wp.fvert[g]), faw = *reinterpret_cast<const unsigned*>(&wp.fadj[g]);
                const unsigned fv3 = fvw & 0xFFFFFFu, fa3 = faw & 0xFFFFFFu, kb = (fvw >> 24) & 63u;
                const unsigned rv = __funnelshift_r(fv3 | (fv3 << 24), fv3 >> 8, 8 * e);  // bytes: corner e, e+1, e+2
                const unsigned ra = __funnelshift_r(fa3 | (fa3 << 24), fa3 >> 8, 8 * e);  // faces opposite corner e, e+1, e+2
                const unsigned rk = (kb | (kb << 6)) >> (2 * e);                          // 2-bit edge indices, same rotation
                const int vC = rv & 0xFF, vA = (rv >> 8) & 0xFF, vB = (rv >> 16) & 0xFF;
                const v2 C{fma(cg.x, AB.x, fma(-cg.y, AB.y, A.x)), fma(cg.x, AB.y, fma(cg.y, AB.x, A.y))};
                nWin += active && j == 0;
                unsigned tm = (active && j == 0) ? wp.tmask[g] : 0u; // the even lane of the pair answers the queries
                // (every slot of this pass is read before anybody pushes: the barrier after the vertex update below separates them)
                // ---- queries: targets inside the entered face (rare: ~K/nF of the windows enter a face that holds a target)
                while (__any_sync(FULL, tm != 0)) {
                    bool improvedT = false;
                    int myT = 0;
                    double cand = 0;
                    v2 dT{0, 0};
                    if (tm) {
                        int t = __ffs(tm) - 1;
                        tm &= tm - 1;
                        double b0 = wp.tb0[t], b1 = wp.tb1[t], b2 = wp.tb2[t];
                        double bA = e == 0 ? b1 : (e == 1 ? b2 : b0), bB = e == 0 ? b2 : (e == 1 ? b0 : b1), bC = e == 0 ? b0 : (e == 1 ? b1 : b2);
                        double rbs = frcp(bA + bB + bC);
                        v2 d{(bA * A.x + bB * B.x + bC * C.x) * rbs, (bA * A.y + bB * B.y + bC * C.y) * rbs}; // the target, seen from the source
                        const double den = cross2(AB, d);
                        const double mu = -cross2(A, d) * frcp(den); // (den = 0: mu is inf or NaN and fails the range test)
                        const double c = sg + fsqrt(d.x * d.x + d.y * d.y);
                        if (den != 0 && mu >= t0 - 1e-12 && mu <= t1 + 1e-12 && atomicMinD(&wp.tbest[t], c)) improvedT = true, myT = t, cand = c, dT = d;
                    }
                    if (__builtin_expect(__any_sync(FULL, improvedT), 0)) { // the winner (lowest lane among equal candidates) records how its path starts and ends
                        __syncwarp();
                        bool win = improvedT && wp.tbest[myT] == cand;
                        if (win) atomicMin(&wp.towner[myT], lane);
                        updateBound(w, hl, K);
                        __syncwarp();
                        if (win && wp.towner[myT] == lane) {
                            if (psv == NOPSV) wp.tsx[myT] = dT.x, wp.tsy[myT] = dT.y;
                            else wp.tsx[myT] = wp.dirx[psv], wp.tsy[myT] = wp.diry[psv];
                            wp.tcode[myT] = 2 + 4 * (g | (e << 8));
                            double rl = frcp(fsqrt(AB.x * AB.x + AB.y * AB.y));
                            wp.tdu[myT] = (dT.x * AB.x + dT.y * AB.y) * rl;
                            wp.tdw[myT] = (-dT.x * AB.y + dT.y * AB.x) * rl;
                        }
                        __syncwarp();
                        if (hl < K) w.towner[hl] = 32;
                        __syncwarp();
                    }
                }
                // ---- children
                const double sideL = cross2(P0, C), sideR = cross2(P1, C);
                const double lc2 = C.x * C.x + C.y * C.y;
                // |side| <= 1e-12 |d| |dC| counts as "on the ray" (squared form: no square roots)
                const bool leftOpen = !(sideL > 0 && sideL * sideL > 1e-24 * (P0.x * P0.x + P0.y * P0.y) * lc2);
                const bool rightOpen = !(sideR < 0 && sideR * sideR > 1e-24 * (P1.x * P1.x + P1.y * P1.y) * lc2);
                const bool inside = leftOpen && rightOpen;
                double DC = wp.D[vC];
                const double dC = sg + fsqrt(lc2);
                const bool upd = active && inside && dC < DC;
                bool improved = false;
                if (upd && j == 0) improved = atomicMinD(&wp.D[vC], dC);
                DC = upd ? dC : DC;
                const float fDA = (float)wp.D[vA], fDB = (float)wp.D[vB], fDC = (float)DC;
                __syncwarp();
                { // the winner writes the start direction carried to this vertex (its own for the real source, the pseudo-source's else)
                    const bool winner = improved && dC == wp.D[vC];
                    const int pq = min((int)psv, W::V - 1);
                    const double ndx = psv == NOPSV ? C.x : wp.dirx[pq], ndy = psv == NOPSV ? C.y : wp.diry[pq];
                    if (winner) wp.dirx[vC] = ndx, wp.diry[vC] = ndy, wp.vdirty[vC] = 1;
                }
                // child j = 0: edge C->A of this face (opposite corner B), entered by the neighbour as A->C
                // child j = 1: edge B->C of this face (opposite corner A), entered by the neighbour as C->B
                // Xin-Wang filter and bound test in fp32 with a conservative margin: a window is dropped only when it is
                // dominated by clearly more than the rounding of the approximation.
                {
                    // Straight-line, evaluated by every lane and masked at the end: some lane of the warp needs every one of these
                    // steps in almost every pass, so nested early-outs only add branch and re-convergence instructions.
                    const v2 X = j ? C : A, Y = j ? B : C;
                    // corner opposite the child edge: B = corner e+2 for j = 0, A = corner e+1 for j = 1
                    const int g2 = (ra >> (j ? 8 : 16)) & 0xFF, kk = (rk >> (j ? 2 : 4)) & 3;
                    const bool open = active && (j ? rightOpen : leftOpen) && g2 != REC_NONE;
                    double2 ccg{0, 0};
                    if (open) ccg = edgeFrame(a.m, wp, g2, kk); // needed by the child at the next pass: issued here, stored with the push
                    const double h0 = hitParam0(P0, X, Y), h1 = hitParam0(P1, X, Y);
                    const double m0 = (j == 1 && inside) ? 0.0 : h0, m1 = (j == 0 && inside) ? 1.0 : h1;
                    const f2 fA = tof2(A), fB = tof2(B), fC = tof2(C);
                    const f2 fX = j ? fC : fA, fY = j ? fB : fC, fO2 = j ? fA : fB;
                    const float dX = j ? fDC : fDA, dY = j ? fDB : fDC, dO = j ? fDA : fDB;
                    const f2 X0 = flerp(fX, fY, (float)m0), X1 = flerp(fX, fY, (float)m1);
                    const float lbc = fsg + fsegDist(fO, X0, X1) * (1.f - 1e-5f);
                    const bool reach = lbc <= fUb;
                    const float keep = 1.f - 2e-5f;
                    const float s0 = (fsg + flen(X0.x, X0.y)) * keep, s1 = (fsg + flen(X1.x, X1.y)) * keep;
                    const f2 Xn = j ? X1 : X0; // the end of the child interval next to the parent edge
                    const float sn = j ? s1 : s0;
                    const bool dom = (dX + fdist(fX, X1) < s1) | (dY + fdist(fY, X0) < s0) | (dO + fdist(fO2, Xn) < sn);
                    const bool valid = open & (m1 - m0 > 1e-13) & reach & !dom;
                    const int cmeta = g2 | (kk << 16);
                    // push the children into the ring of the source they belong to
                    const unsigned bal = __ballot_sync(FULL, valid);
                    const unsigned lanes0 = nb0 >= 16 ? FULL : (1u << (2 * nb0)) - 1u; // the lanes that worked for source 0
                    const int tot0 = __popc(bal & lanes0), tot1 = __popc(bal & ~lanes0);
                    // Children that do not fit into the ring of their source go to that source's spill stack in global memory
                    // (all children of the pass, so a ring never holds a partial pass); the windows come back when the ring runs
                    // dry.  Order does not matter for the result, every update is a minimum.  ~1e-3 of the sources of config 5.
                    const bool ovf0 = tail0 + tot0 - (head0 + nb0) > T::RING, ovf1 = tail1 + tot1 - (head1 + nb1) > T::RING;
                    const int rank = __popc(bal & (src ? ~lanes0 : lanes0) & ((1u << lane) - 1u));
                    const int toth = half ? tot1 : tot0;
                    if (__builtin_expect(!(ovf0 | ovf1), 1)) { // the common case: everything fits
                        if (valid) {
                            const int q = ((src ? tail1 : tail0) + rank) & MASKR;
                            wp.rax[q] = X.x, wp.ray[q] = X.y, wp.rbx[q] = Y.x, wp.rby[q] = Y.y;
                            wp.rt0[q] = m0, wp.rt1[q] = m1, wp.rmeta[q] = cmeta, wp.rpsv[q] = psv;
                            wp.rcg[q] = ccg, wp.rlb[q] = lbc;
                        }
                        tail += toth;
                    } else {
                        const int sp0 = __shfl_sync(FULL, spillN, 0), sp1 = __shfl_sync(FULL, spillN, 16);
                        if (valid) {
                            if (!(src ? ovf1 : ovf0)) {
                                const int q = ((src ? tail1 : tail0) + rank) & MASKR;
                                wp.rax[q] = X.x, wp.ray[q] = X.y, wp.rbx[q] = Y.x, wp.rby[q] = Y.y;
                                wp.rt0[q] = m0, wp.rt1[q] = m1, wp.rmeta[q] = cmeta, wp.rpsv[q] = psv;
                                wp.rcg[q] = ccg, wp.rlb[q] = lbc;
                            } else if ((src ? sp1 : sp0) + rank < SPILL_CAP) {
                                double* e = spillPair + ((size_t)src * SPILL_CAP + (src ? sp1 : sp0) + rank) * SPILL_DOUBLES;
                                e[0] = X.x, e[1] = X.y, e[2] = Y.x, e[3] = Y.y, e[4] = m0, e[5] = m1, e[6] = ccg.x, e[7] = ccg.y;
                                e[8] = __longlong_as_double((long long)(unsigned)cmeta | ((long long)psv << 32));
                            }
                        }
                        if (half ? ovf1 : ovf0) {
                            if (spillN + toth > SPILL_CAP) failed = true, live = false, K = nF = nV = 0, head = tail = 0, spillN = 0; // next tier
                            else {
                                spillN += toth;
                                if (hl == 0) atomicAdd(a.counters + C_SPILLED, (unsigned long long)toth);
                            }
                        } else
                            tail += toth;
                    }
                }
                __syncwarp();
            }

            // ================= rings empty: straight legs from face corners, then pseudo-source fans =================
            if (hl < K) {
                int t = hl;
                double best = w.tbest[t];
                if (w.tFace[t] != 0) {
                    uchar4 tv = w.fvert[w.tFace[t]];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        int cv = k == 0 ? tv.x : (k == 1 ? tv.y : tv.z);
                        double c = w.D[cv] + (k == 0 ? w.tcd0[t] : (k == 1 ? w.tcd1[t] : w.tcd2[t]));
                        if (c < best) {
                            best = c;
                            w.tbest[t] = c;
                            w.tsx[t] = w.dirx[cv], w.tsy[t] = w.diry[cv];
                            w.tcode[t] = 3 + 4 * k;
                        }
                    }
                }
            }
            __syncwarp();
            updateBound(w, hl, K);
            __syncwarp();
            const float fUb = w.ub;
            // A vertex v can lie on a shortest path to target t only if D[v] + |x_v - x_t| (Euclidean lower bound of the
            // remaining leg) beats the best path known to t.
            bool spawned = false;
            const int nVmax = __reduce_max_sync(FULL, nV);
            for (int v0i = 0; v0i < nVmax; v0i += 16) {
                int v = v0i + hl;
                bool fl = v < nV && w.vdirty[v] && w.velig[v] && (float)w.D[v] * (1.f - 1e-6f) <= fUb;
                if (v < nV) w.vdirty[v] = 0;
                if (fl) {
                    bool useful = false;
                    const d3 Pq = vpos(a.m, w, v);
                    double Dv = w.D[v], px = Pq.x, py = Pq.y, pz = Pq.z;
                    for (int t = 0; t < K && !useful; ++t) {
                        double ex = w.tpx[t] - px, ey = w.tpy[t] - py, ez = w.tpz[t] - pz;
                        float lb = sqrtf((float)(ex * ex + ey * ey + ez * ez)) * (1.f - 2e-6f);
                        useful = Dv + (double)lb < w.tbest[t];
                    }
                    fl = useful;
                }
                __syncwarp();
                unsigned bal = __ballot_sync(FULL, fl);
                while (bal) { // the whole warp spawns the fan of ONE vertex of ONE of the two sources at a time
                    const int b = __ffs(bal) - 1;
                    bal &= bal - 1;
                    const int h = b >> 4, pv = v0i + (b & 15);
                    const bool mine = half == h;
                    const int nFh = __shfl_sync(FULL, nF, 16 * h), headh = __shfl_sync(FULL, head, 16 * h);
                    int tailh = __shfl_sync(FULL, tail, 16 * h);
                    const float fUbh = __shfl_sync(FULL, fUb, 16 * h);
                    // a fan pushes one window per face around the vertex; with little room left in the ring the vertex stays dirty
                    // and is taken up again after the ring has been drained
                    bool okh = true;
                    if (tailh - headh > T::RING - 16) {
                        if (lane == 0) wpair[h].vdirty[pv] = 1;
                    } else
                        okh = spawnFan(a.m, wpair[h], lane, nFh, pv, fUbh, headh, tailh);
                    if (mine) {
                        spawned = true, nPs += (hl == 0 && tailh != tail);
                        tail = tailh;
                        if (!okh) failed = true, live = false, K = nF = nV = 0, head = tail = 0;
                    }
                }
            }
            if (!__any_sync(FULL, spawned)) break;
        }

        // ---------------- results, pair forces ----------------
        const bool firstGroup = base == 0, lastGroup = base + T::MAXK >= Ktot;
        unsigned nDis = 0;
        double dres = 0;
        d3 ts{0, 0, 1}, te{0, 0, 1};
        if (hl < K) {
            int t = hl;
            dres = w.tbest[t];
            int code = w.tcode[t];
            if (code == 0 || !(dres < dinf())) { // unreachable inside the patch (triangulatedMeshSpace.cpp:198-203)
                nDis = 1;
                dres = a.submeshing ? 2.0 * a.maxDist : -1.0;
                if (!a.submeshing) ts = te = d3{0, 0, 0};
            } else if (code == 1) {
                double rl = 1.0 / dres;
                ts = d3{(w.tpx[t] - sp.x) * rl, (w.tpy[t] - sp.y) * rl, (w.tpz[t] - sp.z) * rl};
                te = ts;
            } else {
                double dx = w.tsx[t], dy = w.tsy[t];
                double rx = dx * w.root[0] + dy * w.root[3], ry = dx * w.root[1] + dy * w.root[4], rz = dx * w.root[2] + dy * w.root[5];
                double rl = 1.0 / sqrt(rx * rx + ry * ry + rz * rz);
                ts = d3{rx * rl, ry * rl, rz * rl};
                if (a.nbrTe) {
                    if ((code & 3) == 2) {
                        int ge = code >> 2;
                        te = liftEnd(a.m, w, ge & 0xFF, ge >> 8, w.tdu[t], w.tdw[t]);
                    } else {
                        int k = code >> 2;
                        uchar4 tv = w.fvert[w.tFace[t]];
                        int cv = k == 0 ? tv.x : (k == 1 ? tv.y : tv.z);
                        const d3 Pc = vpos(a.m, w, cv);
                        double ex = w.tpx[t] - Pc.x, ey = w.tpy[t] - Pc.y, ez = w.tpz[t] - Pc.z;
                        double rl2 = 1.0 / sqrt(ex * ex + ey * ey + ez * ez);
                        te = d3{ex * rl2, ey * rl2, ez * rl2};
                    }
                }
            }
            size_t o = (size_t)li * a.kmax + base + t;
            a.nbrIdx[o] = w.tIdx[t];
            a.nbrDist[o] = dres;
            if (a.nbrTs) a.nbrTs[3 * o] = ts.x, a.nbrTs[3 * o + 1] = ts.y, a.nbrTs[3 * o + 2] = ts.z;
            if (a.nbrTe) a.nbrTe[3 * o] = te.x, a.nbrTe[3 * o + 1] = te.y, a.nbrTe[3 * o + 2] = te.z;
        }
        // force::computeForces accumulates in neighbour order jj = 0..K-1 (baseForce.cpp:22-26): the first lane of the half adds
        // the pair forces in that order, fetching them from the lanes that computed them
        d3 pf{0, 0, 0};
        if (a.forceMode && hl < K) pf = pairForce(a.fp, ts, dres);
        d3 f{0, 0, 0};
        if (a.forceMode) {
            if (live && hl == 0) {
                if (!firstGroup) f = d3{w.fpart[0], w.fpart[1], w.fpart[2]};
                else if (!a.zero) f = d3{a.frc[3 * li], a.frc[3 * li + 1], a.frc[3 * li + 2]};
            }
            const int Kmax = __reduce_max_sync(FULL, K);
            for (int t = 0; t < Kmax; ++t) {
                double x = __shfl_sync(FULL, pf.x, hbase + t), y = __shfl_sync(FULL, pf.y, hbase + t), z = __shfl_sync(FULL, pf.z, hbase + t);
                if (t < K) f.x += x, f.y += y, f.z += z;
            }
        }
        for (int o = 8; o; o >>= 1) { // sums over the half
            nDis += __shfl_xor_sync(FULL, nDis, o);
            nWin += __shfl_xor_sync(FULL, nWin, o);
        }
#ifdef CSS_PASS_STATS
        if (lane == 0) {
            atomicAdd(a.counters + C_CLK_BATCH, (unsigned long long)nPass), atomicAdd(a.counters + C_CLK_FAN, (unsigned long long)nPassIdle);
            atomicAdd(a.counters + C_CLK_PROP, (unsigned long long)nPass4), atomicAdd(a.counters + C_CLK_PATCH, (unsigned long long)nPop);
            atomicAdd(a.counters + C_CLK_TOTAL, (unsigned long long)nOuter);
        }
#endif
        more = false;
        if (hl == 0) {
            if (live) {
                unsigned* cnt = w.wcnt;
                cnt[C_DISCONNECTED] += nDis;
                cnt[C_WINDOWS] += nWin;
                cnt[C_PSEUDO] += nPs;
                cnt[C_QUERIES] += K;
                if (firstGroup) cnt[C_PATCH_FACES] += nF, cnt[C_PATCH_VERTS] += nV;
                if (lastGroup) {
                    cnt[C_SOURCES]++;
                    a.nbrCount[li] = Ktot;
                    if (a.forceMode) {
                        a.frc[3 * li] = f.x, a.frc[3 * li + 1] = f.y, a.frc[3 * li + 2] = f.z;
                        if (a.kick != 0.0) a.vel[3 * li] += a.kick * f.x, a.vel[3 * li + 1] += a.kick * f.y, a.vel[3 * li + 2] += a.kick * f.z;
                    }
                } else
                    w.fpart[0] = f.x, w.fpart[1] = f.y, w.fpart[2] = f.z;
            } else if (failed) { // ring overflow: the particle is untouched, the next tier reruns it
                int r = atomicAdd(a.retryCount, 1);
                a.retryList[r] = li;
                w.wcnt[C_TIER_RETRY]++;
                atomicAdd(a.counters + C_OVF_REASON + 3, 1ull);
            }
        }
        more = live && !lastGroup;
        __syncwarp();
    }
    __syncwarp();
    if (w.wcnt[hl]) atomicAdd(a.counters + hl, (unsigned long long)w.wcnt[hl]);
}

template <class T> cudaError_t launchWindowsHalf(cudaStream_t st, const WinArgs& a, int numSMs)
{
    constexpr int wpb = CSS_HALF_WPB;
    const size_t smem = sizeof(HalfSmem<T>) * 2 * wpb;
    static int perSMdev[64] = {0}; // the attribute and the occupancy are per device
    int dev = 0;
    cudaGetDevice(&dev);
    int& perSM = perSMdev[dev & 63];
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    if (!perSM) {
        cudaFuncSetAttribute(k_windows_half<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_windows_half<T>, wpb * 32, smem) != cudaSuccess || n < 1) n = 1;
        perSM = n;
    }
    int blocks = numSMs * min(perSM, CSS_HALF_MINBLOCKS); // the spill scratch is sized for this many blocks
    if (!a.srcList) blocks = min(blocks, max(1, (a.nLocal + 2 * wpb - 1) / (2 * wpb)));
    return launchStep(k_windows_half<T>, blocks, wpb * 32, smem, st, a);
}
template cudaError_t launchWindowsHalf<TierHalf>(cudaStream_t, const WinArgs&, int);
size_t windowsHalfSpillBytes(int numSMs)
{ // one pair of spill stacks per resident warp of the persistent grid
    return (size_t)numSMs * CSS_HALF_MINBLOCKS * CSS_HALF_WPB * 2 * SPILL_CAP * SPILL_DOUBLES * sizeof(double);
}

} // namespace css
