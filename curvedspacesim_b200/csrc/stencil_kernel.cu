// Stage 1 of the many-source geodesic path, tier 0, through STATIC FACE STENCILS.
//
// What the reference does per source and per step (submesher::constructSubmeshFromSourceAndTargets,
// src/utility/submesher.cpp:55-147, called from triangulatedMeshSpace::distanceWithSubmeshing, src/models/
// triangulatedMeshSpace.cpp:155-205) is a flood fill over the faces that have a vertex within the cut-off
// R' = min(maximumDistance, largest candidate distance) of the source.  Its result depends on the source only through the
// source's position inside its face and R' <= maximumDistance.  So every patch that can ever be cut for a source lying in face
// f is a subset of one fixed face set, the STENCIL of f: the flood fill run with the cut-off maximumDistance + rho_f around
// the centroid of f (rho_f = largest corner distance from the centroid; |v - x| <= R' implies |v - c_f| <= R' + |x - c_f|).
// The mesh and maximumDistance are fixed for a run, the GPU has 180 GB: the stencils of all faces are built ONCE
// (k_stencil_build, 1.9 kB per face) and a step only RESTRICTS the stencil of the source's face:
//   * which stencil vertices lie within the cut-off            -> one ballot per 32 vertices
//   * which stencil faces have such a vertex (eligible)         -> one ballot per 32 faces
//   * which eligible faces the flood fill reaches: every stencil face carries the face it was discovered from when the stencil
//     was built (its parent); when the parent of every eligible face is itself eligible (or one of the unconditionally taken
//     neighbours of the source face) every eligible face is reached -- one bit test per face, true for all but a handful of
//     sources; otherwise an explicit label propagation over the eligible faces decides (same set as the reference, always)
//   * local numbering = rank inside the bit masks (popc prefix sums); local adjacency = stencil adjacency restricted to the
//     set; leftover goal faces (submesher.cpp:143-144) are just more bits.
// No hash tables, no breadth-first passes, no dependent chains of global loads: every phase is a few rounds of 32 lanes.
// The record that comes out is the one patch_kernel.cu writes (same sections, another -- equally arbitrary -- local numbering);
// stage 2 does not know the difference.  Sources whose face has no stencil (more than 128 faces / 96 vertices), whose targets
// fall outside the stencil or whose patch exceeds the record capacity go to the retry list and are flood-filled.
//
// Every decision that fixes topology (candidate membership and order, vertex inside / outside the cut-off) uses the exactly
// rounded x* helpers of common.cuh with the operands of patch_kernel.cu: the face sets are identical, not just close.
#include "common.cuh"
#include "kernels.h"

namespace css {

#define FULL 0xffffffffu

namespace {

__device__ __forceinline__ unsigned hashInt(int k) { return (unsigned)k * 2654435761u; }
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
__device__ __forceinline__ int pick3(const int4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
__device__ __forceinline__ int warpInclusiveScan(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += y;
    }
    return v;
}

// ============================================================================================ building the stencils (once)
struct BuildSmem {
    static constexpr int HF = 512, HV = 512, FR = 64;
    int fhKey[HF], vhKey[HV];
    unsigned char fhVal[HF], vhVal[HV];
    int4 frAdj[FR], frOpp[FR];
    int gface[STENCIL_F], gvert[STENCIL_V];
    unsigned fvert[STENCIL_F]; // v0 | v1 << 8 | v2 << 16 | kk bits << 24
    unsigned fadj[STENCIL_F];  // n0 | n1 << 8 | n2 << 16 | parent << 24
    unsigned char fin[STENCIL_F];
    unsigned char newid[STENCIL_F];
};

} // namespace

__global__ void __launch_bounds__(128) k_stencil_build(MeshDev m, double maxDist, unsigned char* __restrict__ out, unsigned* __restrict__ lenOut,
                                                       unsigned long long* __restrict__ stats)
{
    __shared__ BuildSmem sm[4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    BuildSmem& s = sm[wib];
    const unsigned ltMask = (1u << lane) - 1u;
    constexpr int HF = BuildSmem::HF, HV = BuildSmem::HV, FRM = BuildSmem::FR - 1;
    for (int sf = blockIdx.x * 4 + wib; sf < m.nF; sf += gridDim.x * 4) {
        unsigned char* rec = out + (size_t)sf * STENCIL_BYTES;
        // cut-off around the centroid that covers the cut-off ball of every source point of the face (with a generous margin
        // over the rounding of the per-step tests; a superset costs nothing but a few idle bits)
        const int4 sc = __ldg(m.corner + sf);
        const d3 p0 = ldvert(m, sc.x), p1 = ldvert(m, sc.y), p2 = ldvert(m, sc.z);
        const d3 c{(p0.x + p1.x + p2.x) / 3.0, (p0.y + p1.y + p2.y) / 3.0, (p0.z + p1.z + p2.z) / 3.0};
        auto dist = [&](const d3& p) { return sqrt((p.x - c.x) * (p.x - c.x) + (p.y - c.y) * (p.y - c.y) + (p.z - c.z) * (p.z - c.z)); };
        const double rho = fmax(dist(p0), fmax(dist(p1), dist(p2)));
        const double thr = (maxDist + rho) * (1.0 + 1e-9) + 1e-300, thr2 = thr * thr;
        auto inside = [&](const double2& xy, const double2& zw) {
            const double dx = xy.x - c.x, dy = xy.y - c.y, dz = zw.x - c.z;
            return dx * dx + dy * dy + dz * dz <= thr2;
        };
        for (int h = lane; h < HF; h += 32) s.fhKey[h] = -1;
        for (int h = lane; h < HV; h += 32) s.vhKey[h] = -1;
        __syncwarp();
        bool in = false;
        if (lane < 3) {
            const int gv = pick3(sc, lane);
            const double2* pv = reinterpret_cast<const double2*>(m.vert + gv);
            in = inside(__ldg(pv), __ldg(pv + 1));
            bool isNew;
            s.vhVal[hashInsert(s.vhKey, HV - 1, gv, isNew)] = (unsigned char)lane;
            s.gvert[lane] = gv;
        }
        const unsigned bits = __ballot_sync(FULL, in) & 7u;
        if (lane == 0) {
            const int4 sA = __ldg(m.adjopp + 2 * (size_t)sf), sO = __ldg(m.adjopp + 2 * (size_t)sf + 1);
            bool isNew;
            s.fhVal[hashInsert(s.fhKey, HF - 1, sf, isNew)] = 0;
            s.gface[0] = sf;
            s.fvert[0] = 0u | (1u << 8) | (2u << 16) | ((unsigned)(sA.w & 63) << 24);
            s.fadj[0] = 0x00FFFFFFu;
            s.fin[0] = (unsigned char)bits;
            s.frAdj[0] = sA, s.frOpp[0] = sO;
        }
        __syncwarp();
        int nF = 1, nV = 3, head = 0;
        bool ovf = false;
        while (head < nF) { // the flood fill of patch_kernel.cu with the stencil cut-off; it also records who discovered whom
            const int cnt = min(10, nF - head);
            const int slotF = lane / 3, k = lane - 3 * slotF;
            const bool active = slotF < cnt;
            const int i = head + slotF;
            int g = -1, d = -1, kk = 0;
            unsigned fvb = 0, fbits = 0;
            if (active) {
                const int4 A = s.frAdj[i & FRM], O = s.frOpp[i & FRM];
                g = pick3(A, k), d = pick3(O, k), kk = (A.w >> (2 * k)) & 3;
                fvb = s.fvert[i], fbits = s.fin[i];
            }
            const bool valid = active && g >= 0;
            int slot = valid ? hashFind(s.fhKey, HF - 1, g) : -1;
            const bool cand = valid && slot < 0;
            const unsigned ina = (fbits >> ((k + 1) % 3)) & 1u, inb = (fbits >> ((k + 2) % 3)) & 1u;
            bool ind = false, elig = false;
            int4 gA = make_int4(0, 0, 0, 0), gO = gA;
            if (cand) {
                const double2* pv = reinterpret_cast<const double2*>(m.vert + d);
                ind = inside(__ldg(pv), __ldg(pv + 1));
                gA = __ldg(m.adjopp + 2 * (size_t)g), gO = __ldg(m.adjopp + 2 * (size_t)g + 1);
                elig = i == 0 || ina || inb || ind;
            }
            bool isNew = false;
            if (elig) slot = hashInsert(s.fhKey, HF - 1, g, isNew);
            const bool win = elig && isNew;
            const unsigned bal = __ballot_sync(FULL, win);
            const int nAdd = __popc(bal);
            if (nF + nAdd > STENCIL_F || nF + nAdd - (head + cnt) > BuildSmem::FR) {
                ovf = true;
                break;
            }
            const int id = nF + __popc(bal & ltMask);
            int vs = -1;
            bool vnew = false;
            if (win) {
                s.fhVal[slot] = (unsigned char)id;
                s.gface[id] = g;
                s.frAdj[id & FRM] = gA, s.frOpp[id & FRM] = gO;
                vs = hashInsert(s.vhKey, HV - 1, d, vnew);
            }
            const unsigned vbal = __ballot_sync(FULL, win && vnew);
            if (nV + __popc(vbal) > STENCIL_V) {
                ovf = true;
                break;
            }
            if (win && vnew) {
                const int vid = nV + __popc(vbal & ltMask);
                s.vhVal[vs] = (unsigned char)vid;
                s.gvert[vid] = d;
            }
            __syncwarp();
            if (active) reinterpret_cast<unsigned char*>(s.fadj + i)[k] = (valid && slot >= 0) ? s.fhVal[slot] : (unsigned char)REC_NONE;
            if (win) {
                const unsigned ld = s.vhVal[vs], la = (fvb >> (8 * ((k + 1) % 3))) & 0xFFu, lb = (fvb >> (8 * ((k + 2) % 3))) & 0xFFu;
                s.fvert[id] = (ld << (8 * kk)) | (lb << (8 * ((kk + 1) % 3))) | (la << (8 * ((kk + 2) % 3))) | ((unsigned)(gA.w & 63) << 24);
                s.fin[id] = (unsigned char)(((unsigned)ind << kk) | (inb << ((kk + 1) % 3)) | (ina << ((kk + 2) % 3)));
                s.fadj[id] = 0x00FFFFFFu | ((unsigned)i << 24); // parent = the face that discovered it; neighbours filled when it is expanded
            }
            __syncwarp();
            nF += nAdd, nV += __popc(vbal), head += cnt;
        }
        __syncwarp();
        if (ovf) {
            if (lane == 0) {
                *reinterpret_cast<int4*>(rec) = make_int4(0, 0, 1, 0);
                lenOut[sf] = 0u;
                atomicAdd(stats, 1ull);
            }
            continue;
        }
        // faces 1.. are stored in ascending global id (targets are located by binary search); face 0 stays the face itself
        for (int f = lane; f < nF; f += 32) {
            int r = 0;
            if (f > 0) {
                const int g = s.gface[f];
                r = 1;
                for (int q = 1; q < nF; ++q) r += s.gface[q] < g;
            }
            s.newid[f] = (unsigned char)r;
        }
        __syncwarp();
        int* ogface = reinterpret_cast<int*>(rec + 16);
        unsigned* ofvert = reinterpret_cast<unsigned*>(rec + 16 + 4 * nF);
        unsigned* ofadj = reinterpret_cast<unsigned*>(rec + 16 + 8 * nF);
        int* ogvert = reinterpret_cast<int*>(rec + 16 + 12 * nF);
        for (int f = lane; f < nF; f += 32) {
            const int r = s.newid[f];
            const unsigned fa = s.fadj[f];
            unsigned o = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) { // three neighbours and the parent
                const unsigned n = (fa >> (8 * k)) & 0xFFu;
                o |= (n == REC_NONE ? (unsigned)REC_NONE : (unsigned)s.newid[n]) << (8 * k);
            }
            if (f == 0) o |= 0xFF000000u; // the face itself has no parent
            ogface[r] = s.gface[f];
            ofvert[r] = s.fvert[f];
            ofadj[r] = o;
        }
        for (int v = lane; v < nV; v += 32) ogvert[v] = s.gvert[v];
        if (lane == 0) {
            *reinterpret_cast<int4*>(rec) = make_int4(nF, nV, 0, 0);
            lenOut[sf] = (unsigned)nF | ((unsigned)nV << 16);
            atomicAdd(stats + 1, (unsigned long long)nF);
        }
        __syncwarp();
    }
}

// ============================================================================================ per step: restrict a stencil
namespace {
template <class T> struct RestrictSmem { // per warp
    int gface[STENCIL_F];              // stencil faces (ascending global id from entry 1): binary search of the target faces
    unsigned char frank[STENCIL_F];    // patch-local id of a stencil face (valid where the face is in the patch)
    unsigned char visb[STENCIL_F];     // stencil face is in the (tentative) patch
    unsigned char vrank[STENCIL_V];    // patch-local id of a stencil vertex
    unsigned char inVb[STENCIL_V];     // stencil vertex lies within the cut-off
    alignas(4) unsigned char usedVb[STENCIL_V];   // stencil vertex belongs to a patch face
    alignas(4) unsigned char borderVb[STENCIL_V]; // ... and to a border edge of the patch
    unsigned char owner[32];           // candidate phase: the stencil cell (lane) that holds the particle a lane tests
    int tIdx[T::RECK];
};
} // namespace

#ifndef CSS_STENCIL_MINB
#define CSS_STENCIL_MINB 8
#endif
template <class T> __global__ void __launch_bounds__(128, CSS_STENCIL_MINB) k_patch_stencil(const __grid_constant__ PatchArgs a)
{
    __shared__ RestrictSmem<T> smAll[4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    RestrictSmem<T>& sm = smAll[wib];
    PDL_ENTRY();
    if (strideGuardUp(a.counters)) return; // the cell-list build found a stencil fuller than the neighbour stride (common.cuh)
    const unsigned ltMask = (1u << lane) - 1u;
    unsigned long long nRetry = 0;
    const int nWork = a.nLocal; // tier 0: sources 0 .. nLocal - 1, record w belongs to local particle w
    // Sources are dealt round-robin to the resident warps (the work per source is even; no work counter, no atomics), so every
    // warp knows its NEXT source while it works on the current one and asks for that source's stencil record one source
    // ahead: the lines travel from HBM during ~10 microseconds of work instead of stalling the warp.
    const int wstride = gridDim.x * (blockDim.x >> 5);
    int w = blockIdx.x * (blockDim.x >> 5) + wib;
    int sfNext = 0;
    unsigned slenNext = 0;
    if (w < nWork) {
        sfNext = a.face[a.minIdx + w];
        slenNext = __ldg(a.stencilLen + sfNext);
        if (lane * 128 < 16 + 12 * (int)(slenNext & 0xFFFFu) + 4 * (int)(slenNext >> 16))
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.stencil + (size_t)sfNext * STENCIL_BYTES + lane * 128));
    }
    for (; w < nWork; w += wstride) {
        const int gi = a.minIdx + w;
        unsigned char* rec = a.records + (size_t)w * T::BYTES;
        int retry = 0; // 0 none, else 1 + overflow reason (1 stride, 2 candidates, 3 faces, 4 vertices)
        const int sf = sfNext;
        const unsigned char* const S = a.stencil + (size_t)sf * STENCIL_BYTES;
        const int nSF = slenNext & 0xFFFFu, nSV = slenNext >> 16;
        const int wNext = w + wstride;
        if (wNext < nWork) { // issued now, consumed at the end of this iteration
            sfNext = a.face[a.minIdx + wNext];
            slenNext = __ldg(a.stencilLen + sfNext);
        }
        const d3 sp{a.eucl[3 * gi], a.eucl[3 * gi + 1], a.eucl[3 * gi + 2]};
        // ---------------- 1. ordered candidates (as in patch_kernel.cu) ----------------
        int K = 0;
        double R = 0;
        {
            const CellGrid& g = a.grid;
            const int c0 = a.cellOf[gi], nxy = g.n[0] * g.n[1];
            const int iz = (int)(((double)c0 + 0.5) * a.invNxy), r0 = c0 - iz * nxy;
            const int iy = (int)(((double)r0 + 0.5) * a.invNx), ix = r0 - iy * g.n[0];
            const int xx = ix + lane / 9 - 1, yy = iy + (lane / 3) % 3 - 1, zz = iz + lane % 3 - 1;
            int s0 = 0, s1 = 0;
            if (lane < 27 && xx >= 0 && xx < g.n[0] && yy >= 0 && yy < g.n[1] && zz >= 0 && zz < g.n[2]) {
                const int c = xx + yy * g.n[0] + zz * nxy;
                s0 = a.cellStart[c];
                s1 = s0 + a.cellCount[c];
            }
            // One lane per stored particle of the 27 cells when they hold at most 32 (practically always): the cells' contents are
            // laid out lane by lane in stencil order (prefix sum of the cell counts), every lane tests its particle, the hits keep
            // that order -- the order of cellListNeighborStructure::constructCandidateNeighborList.
            const int cnt = s1 - s0;
            const int inclC = warpInclusiveScan(cnt, lane), tot = __shfl_sync(FULL, inclC, 31), exclC = inclC - cnt;
            double maxd2 = 0;
            if (tot <= 32) {
                for (int t = 0; t < cnt; ++t) sm.owner[exclC + t] = (unsigned char)lane;
                __syncwarp();
                const int c = lane < tot ? sm.owner[lane] : 0;
                const int q = __shfl_sync(FULL, s0, c) + lane - __shfl_sync(FULL, exclC, c);
                int jj = -1;
                bool hit = false;
                if (lane < tot) {
                    jj = a.cellItems[q];
                    if (jj != gi) {
                        d3 p{a.eucl[3 * jj], a.eucl[3 * jj + 1], a.eucl[3 * jj + 2]};
                        maxd2 = xsqlen(xsub3(sp, p));
                        hit = maxd2 < g.range2;
                    }
                }
                if (!hit) maxd2 = 0;
                const unsigned bal = __ballot_sync(FULL, hit);
                K = __popc(bal);
                if (K > a.kmax) { // neighbour stride too small: raises the stride guard (common.cuh)
                    if (lane == 0) atomicMax(a.counters + C_KMAX_NEED, (unsigned long long)K), atomicAdd(a.counters + C_KMAX_OVERFLOW, 1ull);
                    retry = 1;
                } else if (K > T::RECK)
                    retry = 2;
                else if (hit)
                    sm.tIdx[__popc(bal & ltMask)] = jj;
            } else { // crowded cells: every lane walks its own cell
                int mine = 0;
                for (int q = s0; q < s1; ++q) {
                    int j = a.cellItems[q];
                    if (j == gi) continue;
                    d3 p{a.eucl[3 * j], a.eucl[3 * j + 1], a.eucl[3 * j + 2]};
                    double d2 = xsqlen(xsub3(sp, p));
                    if (d2 < g.range2) mine++, maxd2 = d2 > maxd2 ? d2 : maxd2;
                }
                int incl = warpInclusiveScan(mine, lane);
                K = __shfl_sync(FULL, incl, 31);
                if (K > a.kmax) {
                    if (lane == 0) atomicMax(a.counters + C_KMAX_NEED, (unsigned long long)K), atomicAdd(a.counters + C_KMAX_OVERFLOW, 1ull);
                    retry = 1;
                } else if (K > T::RECK)
                    retry = 2;
                else {
                    int pos = incl - mine;
                    for (int q = s0; q < s1; ++q) {
                        int j = a.cellItems[q];
                        if (j == gi) continue;
                        d3 p{a.eucl[3 * j], a.eucl[3 * j + 1], a.eucl[3 * j + 2]};
                        if (xsqlen(xsub3(sp, p)) < g.range2) sm.tIdx[pos++] = j;
                    }
                }
            }
            if (!retry) { // maximum of non-negative doubles = maximum of their bit patterns: two 32-bit warp reductions
                const unsigned hi = (unsigned)__double2hiint(maxd2), lo = (unsigned)__double2loint(maxd2);
                const unsigned mh = __reduce_max_sync(FULL, hi), ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
                R = xsqrt(__hiloint2double((int)mh, (int)ml));
            }
        }
        __syncwarp();
        int nF = 0, nV = 0;
        if (!retry && K > 0) {
            double thr2 = __longlong_as_double(0x7ff0000000000000LL);
            { // triangulatedMeshSpace::distanceWithSubmeshing :167-169 (this kernel only runs with submeshing on)
                double thr = a.maxDist;
                if (R < a.maxDist) thr = R;
                thr2 = xmul(thr, thr);
            }
            const int myT = lane < K ? sm.tIdx[lane] : -1;
            const int myTF = lane < K ? a.face[myT] : sf;
            const int* sgface = reinterpret_cast<const int*>(S + 16);
            const unsigned* sfvert = reinterpret_cast<const unsigned*>(S + 16 + 4 * nSF);
            const unsigned* sfadj = reinterpret_cast<const unsigned*>(S + 16 + 8 * nSF);
            const int* sgvert = reinterpret_cast<const int*>(S + 16 + 12 * nSF);
            int* ogface = reinterpret_cast<int*>(rec + T::OFF_GFACE);
            int* ogvert = reinterpret_cast<int*>(rec + T::OFF_GVERT);
            unsigned* ofvert = reinterpret_cast<unsigned*>(rec + T::OFF_FVERT);
            unsigned* ofadj = reinterpret_cast<unsigned*>(rec + T::OFF_FADJ);
            if (nSF == 0) retry = 3; // the face has no stencil (too many faces / vertices around it)
            else if (!__any_sync(FULL, myTF != sf)) {
                // every target lies in the source face: the patch is that face (submesher.cpp:79-80)
                const unsigned fv0 = __ldg(sfvert);
                if (lane < 3) {
                    ogvert[lane] = __ldg(sgvert + ((fv0 >> (8 * lane)) & 0xFFu));
                    rec[T::OFF_VELIG + lane] = 1; // all three corners are on the border of the patch
                }
                if (lane == 0) ogface[0] = sf, ofvert[0] = 0u | (1u << 8) | (2u << 16) | (fv0 & 0xFF000000u), ofadj[0] = 0x00FFFFFFu;
                if (lane < K) rec[T::OFF_TFACE + lane] = 0;
                nF = 1, nV = 3;
            } else {
                constexpr int FRND = STENCIL_F / 32, VRND = STENCIL_V / 32; // rounds of 32 lanes; a round runs only if the stencil reaches it
                // ---- 2. stencil vertices within the cut-off (flag bytes in shared memory: one load answers "is vertex v inside")
                int gvr[VRND];
                unsigned sadV[VRND];
                if (lane < STENCIL_V / 4) reinterpret_cast<unsigned*>(sm.usedVb)[lane] = 0u, reinterpret_cast<unsigned*>(sm.borderVb)[lane] = 0u;
#pragma unroll
                for (int r = 0; r < VRND; ++r) {
                    gvr[r] = 0, sadV[r] = 0;
                    if (32 * r < nSV) {
                        const int v = 32 * r + lane;
                        bool in = false, sad = false;
                        if (v < nSV) {
                            gvr[r] = __ldg(sgvert + v);
                            const double2* pv = reinterpret_cast<const double2*>(a.m.vert + gvr[r]);
                            const double2 xy = __ldg(pv), zw = __ldg(pv + 1);
                            in = !(xsqlen(xsub3(sp, d3{xy.x, xy.y, zw.x})) > thr2);
                            sad = zw.y != 0.0;
                        }
                        sm.inVb[v] = in;
                        sadV[r] = __ballot_sync(FULL, sad);
                    }
                }
                // stencil faces: global ids (staged for the binary search of the targets' faces), corner and neighbour ids
                int gfr[FRND];
                unsigned fv[FRND], fa[FRND];
                const unsigned fa0 = __ldg(sfadj); // neighbours of the source face: taken unconditionally (submesher.cpp:83-96)
                const unsigned n0 = fa0 & 0xFFu, n1 = (fa0 >> 8) & 0xFFu, n2 = (fa0 >> 16) & 0xFFu;
#pragma unroll
                for (int r = 0; r < FRND; ++r) {
                    gfr[r] = 0x7fffffff, fv[r] = 0, fa[r] = 0xFFFFFFFFu;
                    if (32 * r < nSF) {
                        const int f = 32 * r + lane;
                        if (f < nSF) gfr[r] = __ldg(sgface + f), fv[r] = __ldg(sfvert + f), fa[r] = __ldg(sfadj + f);
                        sm.gface[f] = gfr[r];
                    }
                }
                __syncwarp();
                // ---- 3. eligible faces, the unconditional neighbours of the source face, the targets' faces
                unsigned eligF[FRND], nMask[FRND], tMask[FRND];
#pragma unroll
                for (int r = 0; r < FRND; ++r) {
                    eligF[r] = 0, nMask[r] = 0;
                    if (32 * r < nSF) {
                        const unsigned f = 32 * r + lane;
                        const bool e = (int)f < nSF && (sm.inVb[fv[r] & 0xFFu] | sm.inVb[(fv[r] >> 8) & 0xFFu] | sm.inVb[(fv[r] >> 16) & 0xFFu]);
                        eligF[r] = __ballot_sync(FULL, e);
                        nMask[r] = __ballot_sync(FULL, f == 0u || f == n0 || f == n1 || f == n2);
                    }
                }
                int tl = -1; // stencil-local face of this lane's target
                if (lane < K) {
                    if (myTF == sf) tl = 0;
                    else {
                        int lo = 1, hi = nSF;
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (sm.gface[mid] < myTF) lo = mid + 1;
                            else hi = mid;
                        }
                        if (lo < nSF && sm.gface[lo] == myTF) tl = lo;
                    }
                }
                if (__any_sync(FULL, lane < K && tl < 0)) retry = 3; // a target outside the stencil (possible only for a leftover goal face)
                else {
                    bool allNear = true;
#pragma unroll
                    for (int r = 0; r < FRND; ++r) {
                        tMask[r] = 0;
                        if (32 * r < nSF) tMask[r] = __reduce_or_sync(FULL, (tl >= 0 && (tl >> 5) == r) ? 1u << (tl & 31) : 0u);
                        allNear &= (tMask[r] & ~nMask[r]) == 0u;
                    }
                    unsigned vis[FRND];
                    if (allNear) { // all goal faces among the source face and its neighbours: that is the patch (submesher.cpp:97-98)
#pragma unroll
                        for (int r = 0; r < FRND; ++r) vis[r] = nMask[r];
                    } else {
#pragma unroll
                        for (int r = 0; r < FRND; ++r) {
                            vis[r] = eligF[r] | nMask[r];
                            if (32 * r < nSF) sm.visb[32 * r + lane] = (vis[r] >> lane) & 1u;
                        }
                        __syncwarp();
                        bool bad = false; // an eligible face whose discoverer is not in the set: reachability has to be worked out
#pragma unroll
                        for (int r = 0; r < FRND; ++r)
                            if ((vis[r] >> lane) & ~(nMask[r] >> lane) & 1u) bad |= !sm.visb[fa[r] >> 24];
                        if (__any_sync(FULL, bad)) { // rare: label propagation from the source face and its neighbours over the eligible faces
                            unsigned cur[FRND];
#pragma unroll
                            for (int r = 0; r < FRND; ++r) cur[r] = nMask[r];
                            for (;;) {
                                __syncwarp();
#pragma unroll
                                for (int r = 0; r < FRND; ++r)
                                    if (32 * r < nSF) sm.visb[32 * r + lane] = (cur[r] >> lane) & 1u;
                                __syncwarp();
                                bool changed = false;
#pragma unroll
                                for (int r = 0; r < FRND; ++r) {
                                    if (32 * r < nSF) {
                                        bool c = false;
                                        if (((vis[r] & ~cur[r]) >> lane) & 1u) {
#pragma unroll
                                            for (int k = 0; k < 3; ++k) {
                                                const unsigned n = (fa[r] >> (8 * k)) & 0xFFu;
                                                c |= n != REC_NONE && sm.visb[n];
                                            }
                                        }
                                        const unsigned nb = __ballot_sync(FULL, c);
                                        cur[r] |= nb;
                                        changed |= nb != 0u;
                                    }
                                }
                                if (!changed) break;
                            }
#pragma unroll
                            for (int r = 0; r < FRND; ++r) vis[r] = cur[r];
                        }
#pragma unroll
                        for (int r = 0; r < FRND; ++r) vis[r] |= tMask[r]; // leftover goal faces (submesher.cpp:143-144)
                    }
                    // ---- 4. local numbering: ranks inside the masks
                    int fb[FRND + 1];
                    fb[0] = 0;
#pragma unroll
                    for (int r = 0; r < FRND; ++r) fb[r + 1] = fb[r] + __popc(vis[r]);
                    nF = fb[FRND];
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < FRND; ++r)
                        if (32 * r < nSF) sm.visb[32 * r + lane] = (vis[r] >> lane) & 1u;
                    __syncwarp();
                    if (nF > T::MAXF) retry = 4;
                    else {
                        unsigned nbv[FRND]; // per face: bit k = the face across edge k is in the patch
#pragma unroll
                        for (int r = 0; r < FRND; ++r) {
                            nbv[r] = 0;
                            if ((vis[r] >> lane) & 1u) {
                                const int f = 32 * r + lane;
                                sm.frank[f] = (unsigned char)(fb[r] + __popc(vis[r] & ltMask));
                                const unsigned v0 = fv[r] & 0xFFu, v1 = (fv[r] >> 8) & 0xFFu, v2 = (fv[r] >> 16) & 0xFFu;
                                sm.usedVb[v0] = 1, sm.usedVb[v1] = 1, sm.usedVb[v2] = 1;
                                const unsigned m0 = (fa[r] & 0xFFu), m1 = (fa[r] >> 8) & 0xFFu, m2 = (fa[r] >> 16) & 0xFFu;
                                const bool in0 = m0 != REC_NONE && sm.visb[m0], in1 = m1 != REC_NONE && sm.visb[m1], in2 = m2 != REC_NONE && sm.visb[m2];
                                nbv[r] = (unsigned)in0 | ((unsigned)in1 << 1) | ((unsigned)in2 << 2);
                                // a border edge k of the patch makes its endpoints (corners k+1, k+2) eligible pseudo-sources
                                if (!in0) sm.borderVb[v1] = 1, sm.borderVb[v2] = 1;
                                if (!in1) sm.borderVb[v2] = 1, sm.borderVb[v0] = 1;
                                if (!in2) sm.borderVb[v0] = 1, sm.borderVb[v1] = 1;
                            }
                        }
                        __syncwarp();
                        unsigned uV[VRND];
                        int vb[VRND + 1];
                        vb[0] = 0;
#pragma unroll
                        for (int r = 0; r < VRND; ++r) {
                            uV[r] = 0;
                            if (32 * r < nSV) uV[r] = __ballot_sync(FULL, sm.usedVb[32 * r + lane] != 0);
                            vb[r + 1] = vb[r] + __popc(uV[r]);
                        }
                        nV = vb[VRND];
                        if (nV > T::MAXV) retry = 5;
                        else {
                            // ---- 5. the record, straight to global memory
#pragma unroll
                            for (int r = 0; r < VRND; ++r)
                                if ((uV[r] >> lane) & 1u) {
                                    const int v = 32 * r + lane, vr = vb[r] + __popc(uV[r] & ltMask);
                                    sm.vrank[v] = (unsigned char)vr;
                                    ogvert[vr] = gvr[r];
                                    rec[T::OFF_VELIG + vr] = (unsigned char)(((sadV[r] >> lane) & 1u) | sm.borderVb[v]);
                                }
                            __syncwarp();
#pragma unroll
                            for (int r = 0; r < FRND; ++r)
                                if ((vis[r] >> lane) & 1u) {
                                    const int id = sm.frank[32 * r + lane];
                                    ogface[id] = gfr[r];
                                    ofvert[id] = (unsigned)sm.vrank[fv[r] & 0xFFu] | ((unsigned)sm.vrank[(fv[r] >> 8) & 0xFFu] << 8) |
                                                 ((unsigned)sm.vrank[(fv[r] >> 16) & 0xFFu] << 16) | (fv[r] & 0xFF000000u);
                                    const unsigned a0 = nbv[r] & 1u ? sm.frank[fa[r] & 0xFFu] : REC_NONE, a1 = nbv[r] & 2u ? sm.frank[(fa[r] >> 8) & 0xFFu] : REC_NONE,
                                                   a2 = nbv[r] & 4u ? sm.frank[(fa[r] >> 16) & 0xFFu] : REC_NONE;
                                    ofadj[id] = a0 | (a1 << 8) | (a2 << 16);
                                }
                            if (lane < K) rec[T::OFF_TFACE + lane] = sm.frank[tl];
                        }
                    }
                }
            }
            if (!retry) { // stage 2 stages the face sections in 16-byte and the eligibility flags in 4-byte pieces: define the tail of the last piece
                const int padF = (4 - (nF & 3)) & 3, padV = (4 - (nV & 3)) & 3;
                if (lane < padF) ofvert[nF + lane] = 0u, ofadj[nF + lane] = 0u;
                if (lane < padV) rec[T::OFF_VELIG + nV + lane] = 0;
            }
        }
        // ---------------- header, candidate ids, retry list ----------------
        if (retry == 3 && a.fallbackList) { // no stencil here: the flood fill (patch_kernel.cu) writes this record
            if (lane == 0) a.fallbackList[atomicAdd(a.fallbackCount, 1)] = w;
        } else if (retry) {
            if (lane == 0) {
                *reinterpret_cast<int4*>(rec) = make_int4(0, 0, 0, 1); // stage 2 skips this source: the large-capacity tier owns it
                int r = atomicAdd(a.retryCount, 1);
                a.retryList[r] = w;
                nRetry++;
                if (retry >= 2) atomicAdd(a.counters + C_OVF_REASON + (retry == 2 ? 0 : (retry == 5 ? 2 : 1)), 1ull); // candidates, faces, vertices
            }
        } else {
            if (lane == 0) *reinterpret_cast<int4*>(rec) = make_int4(nF, nV, K, 0);
            if (lane < K) reinterpret_cast<int*>(rec + T::OFF_TIDX)[lane] = sm.tIdx[lane];
        }
        if (wNext < nWork && lane * 128 < 16 + 12 * (int)(slenNext & 0xFFFFu) + 4 * (int)(slenNext >> 16))
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.stencil + (size_t)sfNext * STENCIL_BYTES + lane * 128));
        __syncwarp();
    }
    if (lane == 0 && nRetry) atomicAdd(a.counters + C_TIER_RETRY, nRetry);
}

size_t stencilBytes(int nF) { return (size_t)nF * STENCIL_BYTES; }

// builds the stencils of all faces; *nMissing = faces without a stencil (too many faces / vertices), *meanFaces = mean stencil size
cudaError_t buildStencils(cudaStream_t st, const MeshDev& m, double maxDist, unsigned char* out, unsigned* lenOut, unsigned long long* statsDev, int numSMs)
{
    cudaError_t e = cudaMemsetAsync(statsDev, 0, 2 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    const int blocks = std::max(1, std::min((m.nF + 3) / 4, numSMs * 8));
    k_stencil_build<<<blocks, 128, 0, st>>>(m, maxDist, out, lenOut, statsDev);
    return cudaGetLastError();
}

template <class T> cudaError_t launchPatchStencil(cudaStream_t st, const PatchArgs& a, int numSMs)
{
    if (a.srcList || a.maxRecords < a.nLocal || !a.stencil || !a.submeshing) return cudaErrorInvalidValue; // tier 0 with submeshing only
    static int perSMdev[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& perSM = perSMdev[dev & 63];
    if (!perSM)
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_patch_stencil<T>, 128, 0) != cudaSuccess || perSM < 1) perSM = 1;
    const int blocks = std::min(numSMs * perSM, std::max(1, (a.nLocal + 3) / 4));
    return launchStep(k_patch_stencil<T>, blocks, 128, 0, st, a);
}
template cudaError_t launchPatchStencil<TierSmall>(cudaStream_t, const PatchArgs&, int);

} // namespace css
