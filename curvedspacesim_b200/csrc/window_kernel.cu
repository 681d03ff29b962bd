// Stage 2 of the many-source geodesic path: exact window propagation on the patch records written by
// patch_kernel.cu, ONE WARP PER SOURCE, the patch staged in shared memory.  This is the kernel of tier 1 (TierLarge records:
// big patches, many candidates, coarse meshes) and of CSS_WIN_HALF=0; tier 0 runs window_half_kernel.cu (two sources per
// warp), which shares the helpers of window_common.cuh with this file.
//
// Replaces CGAL::Surface_mesh_shortest_path as the reference uses it per source
// (src/models/triangulatedMeshSpace.cpp:189-203: add_source_point, build_sequence_tree, then
// shortest_path_points_to_source_points per target, src/utility/meshUtilities.cpp:360-380) and fuses the
// consumer of its results, force::computeForces (src/forces/baseForce.cpp:12-28) with the second
// velocity-Verlet half kick (src/updaters/velocityVerletNVE.cpp:27-28).
//
// Algorithm: Chen-Han window unfolding with the Xin-Wang vertex-distance filter; saddle and patch-border
// vertices are pseudo-sources.  Windows live in a FIFO ring in shared memory and are propagated 16 at a
// time, a pair of lanes per window (one child edge each); children are compacted into the ring with a ballot.  A window is
// (A, B, S, t0, t1, sigma): the entered edge A->B and the image S of its (pseudo-)source in one common
// unfolded 2-D frame, the visible interval [t0, t1] of the edge, and the distance sigma from the true source
// to the pseudo-source.  Unfolding across a face is four FMAs with the precomputed edge frames
// (MeshDev::geo), so the propagation loop touches no 3-D geometry.  Start directions are carried as 2-D
// vectors in the source face's frame and lifted to 3-D once per target at the end; end tangents likewise.
//
// Everything here is ordinary fp64 with FMA contraction: results agree with the oracle to ~1e-15, far
// inside the 1e-9 bar.  Pruning decisions are taken in fp32 with conservative margins.

#include "window_common.cuh"

namespace css {

namespace {

// LEAN = true: edge frames and vertex positions are not staged; the kernel reads them from global memory (L2-resident)
// through the local->global maps.  10.4 KB instead of 15.9 KB per warp for TierSmall -> 20 instead of 14 warps per SM.
template <class T, bool LEAN> struct WinSmem { // per-warp workspace
    static constexpr int F = T::MAXF, V = T::MAXV, K = T::MAXK, R = T::RING;
    static constexpr bool lean = LEAN;
    static constexpr bool ringLb = false; // (the two-sources-per-warp kernel carries a lower bound per ring entry)
    double2 geo[LEAN ? 1 : 3 * F];
    double vx[LEAN ? 1 : V], vy[LEAN ? 1 : V], vz[LEAN ? 1 : V];
    int gface[LEAN ? F : 1], gvert[LEAN ? V : 1];
    double D[V], dirx[V], diry[V];
    double rax[R], ray[R], rbx[R], rby[R], rt0[R], rt1[R]; // ring: the (pseudo-)source image and sigma follow from rpsv
    double2 rcg[LEAN ? R : 1];                             // ring: edge frame of the face the window enters (fetched at push time)
    double tbest[K], tb0[K], tb1[K], tb2[K], tsx[K], tsy[K], tdu[K], tdw[K];
    double tpx[K], tpy[K], tpz[K], tcd0[K], tcd1[K], tcd2[K];
    double root[6];
    double fpart[3]; // pair forces of the earlier target groups of this source
    unsigned long long wcnt[16];
    int grp[2];      // [0] first target of the group being propagated, [1] number of targets of the source
    int rmeta[R];
    int tIdx[K];
    int tcode[K];  // how the best path ends: 0 none, 1 chord in the source face, 2 + 4*(g | e << 8) window, 3 + 4*k corner k
    int towner[K];
    unsigned tmask[F]; // targets lying in each face (bit t)
    alignas(16) uchar4 fvert[F];
    uchar4 fadj[F];
    alignas(16) unsigned char tFace[K];
    unsigned char velig[V];
    unsigned char vdirty[V];
    unsigned char rpsv[R];
};


// One propagation for the targets [base, base + T::MAXK) of the record, base = w.grp[0].  The pair forces are summed in
// neighbour order across the groups (the partial sum waits in w.fpart); force and kicked velocity are written by the last
// group only, so a ring overflow in any group leaves the particle untouched for the retry tiers.  The group state lives in
// shared memory and is re-read where it is needed: the common single-group path keeps no extra registers alive.
template <class T, bool LEAN, bool GROUPED> __device__ int processRecord(const WinArgs& a, WinSmem<T, LEAN>& w, int li, int rslot, int lane)
{
    constexpr int MASKR = T::RING - 1;
    unsigned long long* cnt = w.wcnt;
    const int gi = a.minIdx + li;
    const unsigned char* rec = a.records + (size_t)rslot * T::BYTES;
    const int4 hdr = *reinterpret_cast<const int4*>(rec);
    const int nF = hdr.x, nV = hdr.y;
    if (hdr.w) return WS_OK; // overflowed in stage 1: the retry tiers own this source
    int Kg = hdr.z;
    if constexpr (GROUPED) {
        Kg = min(T::MAXK, hdr.z - w.grp[0]);
        if (lane == 0) w.grp[1] = hdr.z; // read back by the caller: are there more target groups?
    } else if (hdr.z > T::MAXK)
        return WS_GROUPS; // more candidates than one propagation takes: the caller switches to the grouped instantiation
    const int K = Kg;
    if (hdr.z == 0) {
        if (lane == 0) {
            cnt[C_SOURCES]++;
            a.nbrCount[li] = 0;
            if (a.forceMode) {
                d3 f = a.zero ? d3{0, 0, 0} : d3{a.frc[3 * li], a.frc[3 * li + 1], a.frc[3 * li + 2]};
                a.frc[3 * li] = f.x, a.frc[3 * li + 1] = f.y, a.frc[3 * li + 2] = f.z;
                if (a.kick != 0.0) a.vel[3 * li] += a.kick * f.x, a.vel[3 * li + 1] += a.kick * f.y, a.vel[3 * li + 2] += a.kick * f.z;
            }
        }
        return WS_OK;
    }

    // ---------------- stage the patch ----------------
    {
        const int4* src = reinterpret_cast<const int4*>(rec + T::OFF_FVERT); // fvert | fadj are contiguous in both layouts
        int4* dst = reinterpret_cast<int4*>(w.fvert);
        for (int q = lane; q < (2 * 4 * T::MAXF) / 16; q += 32)
            if (q * 4 < nF || (q >= T::MAXF / 4 && (q - T::MAXF / 4) * 4 < nF)) dst[q] = src[q];
        const int* gface = reinterpret_cast<const int*>(rec + T::OFF_GFACE);
        if constexpr (LEAN) {
            for (int f = lane; f < nF; f += 32) w.gface[f] = gface[f];
        } else {
            for (int q = lane; q < 3 * nF; q += 32) {
                int f = q / 3, e = q - 3 * f;
                w.geo[q] = __ldg(a.m.geo + 3 * (size_t)gface[f] + e);
            }
        }
        for (int f = lane; f < nF; f += 32) w.tmask[f] = 0;
        const int* gvert = reinterpret_cast<const int*>(rec + T::OFF_GVERT);
        for (int v = lane; v < nV; v += 32) {
            if constexpr (LEAN) w.gvert[v] = gvert[v];
            else {
                d3 p = ldvert(a.m, gvert[v]);
                w.vx[v] = p.x, w.vy[v] = p.y, w.vz[v] = p.z;
            }
            w.D[v] = dinf();
            w.velig[v] = rec[T::OFF_VELIG + v];
            w.vdirty[v] = 0;
        }
        if (lane < K) {
            const int base = GROUPED ? w.grp[0] : 0;
            int j = reinterpret_cast<const int*>(rec + T::OFF_TIDX)[base + lane];
            w.tIdx[lane] = j;
            w.tFace[lane] = rec[T::OFF_TFACE + base + lane];
            w.tb0[lane] = a.bary[3 * j], w.tb1[lane] = a.bary[3 * j + 1], w.tb2[lane] = a.bary[3 * j + 2];
            w.tpx[lane] = a.eucl[3 * j], w.tpy[lane] = a.eucl[3 * j + 1], w.tpz[lane] = a.eucl[3 * j + 2];
            w.tbest[lane] = dinf();
            w.tcode[lane] = 0;
            w.tsx[lane] = 0, w.tsy[lane] = 0;
            w.towner[lane] = 32;
        }
    }
    const d3 sp{a.eucl[3 * gi], a.eucl[3 * gi + 1], a.eucl[3 * gi + 2]};
    const double sb0 = a.bary[3 * gi], sb1 = a.bary[3 * gi + 1], sb2 = a.bary[3 * gi + 2];
    __syncwarp();

    // ---------------- root frame, direct legs ----------------
    // source face = local face 0: corner 0 at the origin, corner 1 on +x, corner 2 above
    v2 rq1, rq2, S2;
    {
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
        S2 = v2{(sb1 * rq1.x + sb2 * rq2.x) * rbs, (sb2 * rq2.y) * rbs};
        if (lane == 0) w.root[0] = ex.x, w.root[1] = ex.y, w.root[2] = ex.z, w.root[3] = ey.x, w.root[4] = ey.y, w.root[5] = ey.z;
        if (lane < 3) { // straight legs to the three corners of the source face
            int cv = lane == 0 ? fv.x : (lane == 1 ? fv.y : fv.z);
            v2 q = lane == 0 ? v2{0, 0} : (lane == 1 ? rq1 : rq2);
            d3 P = lane == 0 ? P0 : (lane == 1 ? P1 : P2);
            d3 d{P.x - sp.x, P.y - sp.y, P.z - sp.z};
            w.D[cv] = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
            w.dirx[cv] = q.x - S2.x, w.diry[cv] = q.y - S2.y;
            w.vdirty[cv] = 1;
        }
        if (lane < K) {
            int t = lane, lf = w.tFace[t];
            if (lf == 0) { // target in the source face: the chord
                d3 d{w.tpx[t] - sp.x, w.tpy[t] - sp.y, w.tpz[t] - sp.z};
                w.tbest[t] = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
                w.tcode[t] = 1;
            } else {
                atomicOr(&w.tmask[lf], 1u << t);
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
    int head = 0, tail = 0;
    {
        bool valid = false;
        v2 A{0, 0}, B{0, 0};
        int meta = 0;
        if (lane < 3) {
            uchar4 fa = w.fadj[0];
            int g = lane == 0 ? fa.x : (lane == 1 ? fa.y : fa.z);
            if (g != REC_NONE) {
                int kk = (w.fvert[0].w >> (2 * lane)) & 3;
                valid = true;
                meta = g | (kk << 16);
                // edge k runs corner k+1 -> corner k+2; the neighbour sees it reversed
                A = lane == 0 ? rq2 : (lane == 1 ? v2{0, 0} : rq1);
                B = lane == 0 ? rq1 : (lane == 1 ? rq2 : v2{0, 0});
            }
        }
        double2 cg{0, 0};
        if constexpr (LEAN)
            if (valid) cg = edgeFrame(a.m, w, meta & 0xFF, (meta >> 16) & 3);
        pushWindows(w, lane, head, tail, valid, A, B, 0.0, 1.0, meta, NOPSV, cg);
    }
    __syncwarp();

    unsigned long long nWin = 0, nPs = 0;
#ifdef CSS_PASS_STATS
    unsigned nPass = 0, nPass3 = 0, nPass8 = 0, nPop = 0, nOuter = 0;
#endif
    for (;;) {
#ifdef CSS_PASS_STATS
        nOuter++;
#endif
        // ================= drain the ring, 16 windows per pass: a PAIR of lanes per window, one child edge each ==========
        // (the two lanes of a pair run the same instructions up to the children, so the pass costs one child instead of two;
        //  BFS levels of these patches rarely exceed 16 windows)
        while (head != tail) {
            const float fUb = warpBound(w, lane, K);
            const int nb = min(16, tail - head);
            const int j = lane & 1; // which child edge this lane propagates into
            bool active = (lane >> 1) < nb;
            const int p = (head + (lane >> 1)) & MASKR;
            head += nb;
            v2 A{0, 0}, B{1, 0}, S{0, -1};
            double t0 = 0, t1 = 1, sg = 0;
            int meta = 0;
            unsigned char psv = NOPSV;
            if (active) {
                A = v2{w.rax[p], w.ray[p]}, B = v2{w.rbx[p], w.rby[p]};
                t0 = w.rt0[p], t1 = w.rt1[p], meta = w.rmeta[p], psv = w.rpsv[p];
                // a window family lives in the frame of its (pseudo-)source: the real source sits at S2 in the root frame,
                // a pseudo-source at the origin of its fan frame with sigma = its current vertex distance
                if (psv == NOPSV) S = S2;
                else S = v2{0, 0}, sg = w.D[psv];
            }
            __syncwarp(); // every slot of this pass is read before anybody pushes
            const int g = meta & 0xFF, e = (meta >> 16) & 3;
            const v2 AB = B - A;
            const v2 P0 = lerp2(A, B, t0), P1 = lerp2(A, B, t1);
            const f2 fS = tof2(S);
            const float fsg = (float)sg;
            if (active && fsg + fsegDist(fS, tof2(P0), tof2(P1)) * (1.f - 1e-5f) > fUb) active = false; // bound tightened since the push
#ifdef CSS_PASS_STATS
            nPass++, nPass3 += nb <= 8, nPass8 += nb <= 4, nPop += nb; // passes, passes of <= 8 / <= 4 windows, windows popped
#endif
            // ---- unfold the entered face: apex C from the edge frame
            int vA = 0, vB = 0, vC = 0, kkbits = 0;
            uchar4 fa = make_uchar4(REC_NONE, REC_NONE, REC_NONE, 0);
            v2 C{0, 1};
            unsigned tm = 0;
            if (active) {
                nWin += j == 0;
                uchar4 fv = w.fvert[g];
                fa = w.fadj[g];
                kkbits = fv.w;
                vA = e == 0 ? fv.y : (e == 1 ? fv.z : fv.x);
                vB = e == 0 ? fv.z : (e == 1 ? fv.x : fv.y);
                vC = e == 0 ? fv.x : (e == 1 ? fv.y : fv.z);
                double2 cg;
                if constexpr (LEAN) cg = w.rcg[p];
                else cg = edgeFrame(a.m, w, g, e);
                C = v2{fma(cg.x, AB.x, fma(-cg.y, AB.y, A.x)), fma(cg.x, AB.y, fma(cg.y, AB.x, A.y))};
                if (j == 0) tm = w.tmask[g]; // the even lane of the pair answers the queries
            }
            // ---- queries: targets inside the entered face (rare: ~K/nF of the windows enter a face that holds a target)
            while (__any_sync(FULL, tm != 0)) {
                bool improvedT = false;
                int myT = 0;
                double cand = 0;
                v2 dT{0, 0};
                if (tm) {
                    int t = __ffs(tm) - 1;
                    tm &= tm - 1;
                    double b0 = w.tb0[t], b1 = w.tb1[t], b2 = w.tb2[t];
                    double bA = e == 0 ? b1 : (e == 1 ? b2 : b0), bB = e == 0 ? b2 : (e == 1 ? b0 : b1), bC = e == 0 ? b0 : (e == 1 ? b1 : b2);
                    double rbs = frcp(bA + bB + bC);
                    v2 Tq{(bA * A.x + bB * B.x + bC * C.x) * rbs, (bA * A.y + bB * B.y + bC * C.y) * rbs};
                    v2 d = Tq - S;
                    double den = cross2(AB, d);
                    if (den != 0) {
                        double mu = cross2(S - A, d) * frcp(den);
                        if (mu >= t0 - 1e-12 && mu <= t1 + 1e-12) {
                            double c = sg + fsqrt(d.x * d.x + d.y * d.y);
                            if (atomicMinD(&w.tbest[t], c)) improvedT = true, myT = t, cand = c, dT = d;
                        }
                    }
                }
                if (__any_sync(FULL, improvedT)) { // the winner (lowest lane among equal candidates) records how its path starts and ends
                    __syncwarp();
                    bool win = improvedT && w.tbest[myT] == cand;
                    if (win) atomicMin(&w.towner[myT], lane);
                    __syncwarp();
                    if (win && w.towner[myT] == lane) {
                        if (psv == NOPSV) w.tsx[myT] = dT.x, w.tsy[myT] = dT.y;
                        else w.tsx[myT] = w.dirx[psv], w.tsy[myT] = w.diry[psv];
                        w.tcode[myT] = 2 + 4 * (g | (e << 8));
                        double rl = frcp(fsqrt(AB.x * AB.x + AB.y * AB.y));
                        w.tdu[myT] = (dT.x * AB.x + dT.y * AB.y) * rl;
                        w.tdw[myT] = (-dT.x * AB.y + dT.y * AB.x) * rl;
                    }
                    __syncwarp();
                    if (lane < K) w.towner[lane] = 32;
                    __syncwarp();
                }
            }
            // ---- children
            bool improved = false, leftOpen = false, rightOpen = false, inside = false;
            double dC = 0;
            float fDA = 0.f, fDB = 0.f, fDC = 0.f;
            if (active) {
                const v2 dL = P0 - S, dR = P1 - S, dCv = C - S;
                const double sideL = cross2(dL, dCv), sideR = cross2(dR, dCv);
                const double lc2 = dCv.x * dCv.x + dCv.y * dCv.y;
                // |side| <= 1e-12 |d| |dC| counts as "on the ray" (squared form: no square roots)
                leftOpen = !(sideL > 0 && sideL * sideL > 1e-24 * (dL.x * dL.x + dL.y * dL.y) * lc2);
                rightOpen = !(sideR < 0 && sideR * sideR > 1e-24 * (dR.x * dR.x + dR.y * dR.y) * lc2);
                inside = leftOpen && rightOpen;
                double DC = w.D[vC];
                if (inside) {
                    dC = sg + fsqrt(lc2);
                    if (dC < DC) {
                        if (j == 0) improved = atomicMinD(&w.D[vC], dC);
                        DC = fmin(DC, dC);
                    }
                }
                fDA = (float)w.D[vA], fDB = (float)w.D[vB], fDC = (float)DC;
            }
            __syncwarp();
            if (improved && dC == w.D[vC]) { // the winner writes the start direction carried to this vertex
                if (psv == NOPSV) w.dirx[vC] = C.x - S.x, w.diry[vC] = C.y - S.y;
                else w.dirx[vC] = w.dirx[psv], w.diry[vC] = w.diry[psv];
                w.vdirty[vC] = 1;
            }
            // child j = 0: edge C->A of this face (opposite corner B), entered by the neighbour as A->C
            // child j = 1: edge B->C of this face (opposite corner A), entered by the neighbour as C->B
            // Xin-Wang filter and bound test in fp32 with a conservative margin: a window is dropped only when it is
            // dominated by clearly more than the rounding of the approximation.
            const f2 fA = tof2(A), fB = tof2(B), fC = tof2(C);
            {
                const v2 X = j ? C : A, Y = j ? B : C;
                bool valid = false;
                double m0 = 0, m1 = 1;
                int cmeta = 0;
                double2 ccg{0, 0};
                if (active && (j ? rightOpen : leftOpen)) {
                    const int io = j ? (e == 2 ? 0 : e + 1) : (e == 0 ? 2 : e - 1); // corner opposite the child edge: iA / iB
                    const int g2 = io == 0 ? fa.x : (io == 1 ? fa.y : fa.z);
                    if (g2 != REC_NONE) {
                        const int kk = (kkbits >> (2 * io)) & 3;
                        if constexpr (LEAN) ccg = edgeFrame(a.m, w, g2, kk); // needed by the child at the next pass: issued here, stored with the push
                        if (!(j == 1 && inside)) m0 = hitParam(S, P0, X, Y);
                        if (!(j == 0 && inside)) m1 = hitParam(S, P1, X, Y);
                        if (m1 - m0 > 1e-13) {
                            const f2 fX = j ? fC : fA, fY = j ? fB : fC, fO = j ? fA : fB;
                            const float dX = j ? fDC : fDA, dY = j ? fDB : fDC, dO = j ? fDA : fDB;
                            const f2 X0 = flerp(fX, fY, (float)m0), X1 = flerp(fX, fY, (float)m1);
                            if (fsg + fsegDist(fS, X0, X1) <= fUb) {
                                const float keep = 1.f - 2e-5f;
                                const float s0 = (fsg + fdist(fS, X0)) * keep, s1 = (fsg + fdist(fS, X1)) * keep;
                                const f2 Xn = j ? X1 : X0; // the end of the child interval next to the parent edge
                                const float sn = j ? s1 : s0;
                                const bool dom = (dX + fdist(fX, X1) < s1) || (dY + fdist(fY, X0) < s0) || (dO + fdist(fO, Xn) < sn);
                                valid = !dom;
                                cmeta = g2 | (kk << 16);
                            }
                        }
                    }
                }
                if (!pushWindows(w, lane, head, tail, valid, X, Y, m0, m1, cmeta, psv, ccg)) return WS_RING;
            }
            __syncwarp();
        }

        // ================= ring empty: straight legs from face corners, then pseudo-source fans =================
        if (lane < K) {
            int t = lane;
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
        const float fUb = warpBound(w, lane, K);
        // A vertex v can lie on a shortest path to target t only if D[v] + |x_v - x_t| (Euclidean lower bound of the
        // remaining leg) beats the best path known to t.
        bool spawned = false, ok = true;
        for (int v0i = 0; v0i < nV; v0i += 32) {
            int v = v0i + lane;
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
            unsigned bal = __ballot_sync(FULL, fl);
            while (bal) {
                int b = __ffs(bal) - 1;
                bal &= bal - 1;
                spawned = true;
                nPs++;
                ok = spawnFan(a.m, w, lane, nF, v0i + b, fUb, head, tail) && ok;
            }
        }
        if (!ok) return WS_RING;
        if (!spawned) break;
    }

    // ---------------- results, pair forces ----------------
    int base = 0, Ktot = K;
    if constexpr (GROUPED) base = *reinterpret_cast<volatile int*>(&w.grp[0]), Ktot = *reinterpret_cast<volatile int*>(&w.grp[1]);
    const bool firstGroup = base == 0, lastGroup = base + T::MAXK >= Ktot;
    unsigned long long nDis = 0;
    double dres = 0;
    d3 ts{0, 0, 1}, te{0, 0, 1};
    if (lane < K) {
        int t = lane;
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
    // force::computeForces accumulates in neighbour order jj = 0..K-1 (baseForce.cpp:22-26): lane 0 adds the pair
    // forces in that order, fetching them from the lanes that computed them
    d3 pf{0, 0, 0};
    if (a.forceMode && lane < K) pf = pairForce(a.fp, ts, dres);
    d3 f{0, 0, 0};
    if (a.forceMode) {
        if (!firstGroup) f = d3{w.fpart[0], w.fpart[1], w.fpart[2]};
        else if (lane == 0 && !a.zero) f = d3{a.frc[3 * li], a.frc[3 * li + 1], a.frc[3 * li + 2]};
        for (int t = 0; t < K; ++t) {
            double x = __shfl_sync(FULL, pf.x, t), y = __shfl_sync(FULL, pf.y, t), z = __shfl_sync(FULL, pf.z, t);
            f.x += x, f.y += y, f.z += z;
        }
    }
    for (int o = 16; o; o >>= 1) {
        nDis += __shfl_xor_sync(FULL, nDis, o);
        nWin += __shfl_xor_sync(FULL, nWin, o);
    }
    if (lane == 0) {
        cnt[C_DISCONNECTED] += nDis;
        cnt[C_WINDOWS] += nWin;
        cnt[C_PSEUDO] += nPs;
#ifdef CSS_PASS_STATS
        atomicAdd(a.counters + C_CLK_BATCH, (unsigned long long)nPass), atomicAdd(a.counters + C_CLK_FAN, (unsigned long long)nPass3);
        atomicAdd(a.counters + C_CLK_PROP, (unsigned long long)nPass8), atomicAdd(a.counters + C_CLK_PATCH, (unsigned long long)nPop);
        atomicAdd(a.counters + C_CLK_TOTAL, (unsigned long long)nOuter);
#endif
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
    }
    return WS_OK;
}

} // namespace

// rare path (~5e-6 of the sources of config 5): one propagation per group of T::MAXK targets, out of line so that the
// common path compiles exactly as if it did not exist
template <class T, bool LEAN> __device__ __noinline__ int processGroups(const WinArgs& a, WinSmem<T, LEAN>& w, int li, int rslot, int lane)
{
    int st = WS_OK;
    for (int base = 0;; base += T::MAXK) {
        if (lane == 0) w.grp[0] = base;
        __syncwarp();
        st = processRecord<T, LEAN, true>(a, w, li, rslot, lane);
        st = __shfl_sync(FULL, st, 0);
        __syncwarp();
        if (st != WS_OK || base + T::MAXK >= w.grp[1]) break;
    }
    return st;
}

#ifndef CSS_LEAN_MINBLOCKS
#define CSS_LEAN_MINBLOCKS 5
#endif
template <class T, bool LEAN> __global__ void __launch_bounds__(128, LEAN ? CSS_LEAN_MINBLOCKS : 4) k_windows(const __grid_constant__ WinArgs a)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WinSmem<T, LEAN>& w = reinterpret_cast<WinSmem<T, LEAN>*>(smemRaw)[wib];
    unsigned long long* cnt = w.wcnt;
    PDL_ENTRY();
    if (strideGuardUp(a.counters)) return; // stage 1 did not run (common.cuh)
    if (lane < 16) cnt[lane] = 0;
    __syncwarp();
    const int nWork = a.srcList ? min(*a.srcCount, a.maxRecords) : a.nLocal;
    for (;;) {
        int s = 0;
        if (lane == 0) s = atomicAdd(a.workCounter, 1);
        s = __shfl_sync(FULL, s, 0);
        if (s >= nWork) break;
        const int li = a.srcList ? a.srcList[s] : s;
        int st = processRecord<T, LEAN, false>(a, w, li, s, lane);
        st = __shfl_sync(FULL, st, 0);
        __syncwarp();
        if constexpr (T::RECK > T::MAXK)
            if (st == WS_GROUPS) st = processGroups<T, LEAN>(a, w, li, s, lane);
        if (st != WS_OK && lane == 0) { // ring overflow: rerun on the next tier
            int r = atomicAdd(a.retryCount, 1);
            a.retryList[r] = li;
            cnt[C_TIER_RETRY]++;
            atomicAdd(a.counters + C_OVF_REASON + 3, 1ull);
        }
    }
    __syncwarp();
    if (lane < 16 && cnt[lane]) atomicAdd(a.counters + lane, cnt[lane]);
}

template <class T, bool LEAN> static cudaError_t launchWindowsImpl(cudaStream_t st, const WinArgs& a, int warpsPerBlock, int numSMs)
{
    size_t smem = sizeof(WinSmem<T, LEAN>) * warpsPerBlock;
    static int perSMdev[64][5] = {{0}}; // the attribute and the occupancy are per device
    int dev = 0;
    cudaGetDevice(&dev);
    int* perSM = perSMdev[dev & 63];
    if (warpsPerBlock < 1 || warpsPerBlock > 4 || smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    if (!perSM[warpsPerBlock]) {
        cudaFuncSetAttribute(k_windows<T, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_windows<T, LEAN>, warpsPerBlock * 32, smem) != cudaSuccess || n < 1) n = 1;
        perSM[warpsPerBlock] = n;
    }
    int blocks = numSMs * perSM[warpsPerBlock];
    if (!a.srcList) blocks = min(blocks, max(1, (a.nLocal + warpsPerBlock - 1) / warpsPerBlock));
    return launchStep(k_windows<T, LEAN>, blocks, warpsPerBlock * 32, smem, st, a);
}
template <class T> cudaError_t launchWindows(cudaStream_t st, const WinArgs& a, int warpsPerBlock, int numSMs, bool lean)
{
    return lean ? launchWindowsImpl<T, true>(st, a, warpsPerBlock, numSMs) : launchWindowsImpl<T, false>(st, a, warpsPerBlock, numSMs);
}
template cudaError_t launchWindows<TierSmall>(cudaStream_t, const WinArgs&, int, int, bool);
template cudaError_t launchWindows<TierLarge>(cudaStream_t, const WinArgs&, int, int, bool);

} // namespace css
