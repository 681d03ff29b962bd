// Stage 1 of the many-source geodesic path: ordered Euclidean candidates + local patch, ONE WARP PER SOURCE.
//
// For its source particle the warp
//   (1) gathers the ordered candidate list from the cell list, exactly as
//       cellListNeighborStructure::constructCandidateNeighborList does
//       (src/utility/cellListNeighborStructure.cpp:45-84; stencil order of
//       hyperRectangularCellList::getCellNeighbors, src/utility/hyperRectangularCellList.cpp:128-159),
//   (2) flood-fills the patch face set with the rule of submesher::constructSubmeshFromSourceAndTargets
//       (src/utility/submesher.cpp:55-147; cut-off R' = min(maximumDistance, largest candidate distance),
//       src/models/triangulatedMeshSpace.cpp:167-169),
//   (3) re-indexes faces and vertices locally (8-bit ids) and marks pseudo-source-eligible vertices
//       (saddle vertices of the mesh and vertices on the patch border), and
//   (4) writes one fixed-stride PATCH RECORD to global memory with coalesced 16-byte stores.
// Stage 2 (window_kernel.cu) streams the records back into shared memory and runs the exact window
// propagation.  Splitting the path keeps each kernel's instruction footprint inside the SM's instruction
// cache (the fused kernel spent most of its issue slots waiting for instruction fetches) and lets this
// integer / latency-bound stage run at 40+ resident warps per SM while the FP64 stage keeps its registers.
//
// Every decision here fixes topology (candidate membership and order, patch membership), so all
// arithmetic uses the exactly rounded x* helpers of common.cuh and is bit-identical to the CPU oracle.
// Sources whose candidates / patch exceed the record capacity are appended to the retry list and handled
// by the large-capacity tiers of geodesic_kernel.cu.
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
__device__ __forceinline__ int warpInclusiveScan(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += y;
    }
    return v;
}
// ------------------------------------------------------------------------------------------------------------------------
// Version 2 of the patch builder (default).  Same candidates, same face set, same record as buildPatch above; what changes is
// how the flood fill walks the mesh:
//   * ONE L2 round trip per breadth-first pass instead of three.  The table `adjopp` holds, per face, its three neighbours AND
//     the vertex of each neighbour that lies opposite the shared edge.  A frontier face therefore knows, from shared memory,
//     which face g lies across each edge and which single vertex d of g it has not seen yet; the pass fetches the position of
//     d (with the vertex's saddle flag in its fourth component) and, speculatively, g's own adjopp entry in the same trip.
//   * every vertex is tested against the cut-off once per face that discovers it (the other two corners of g are corners of the
//     frontier face: their inside bits are kept per face), instead of three position fetches per visited (face, edge) pair.
//   * local indexing is fused into the flood fill: the local corner ids of g are those of the frontier face plus (at most) one
//     new vertex, the local adjacency entry of an edge is known as soon as the face across it is looked up.  The two
//     re-indexing passes over the finished face set (6 hash probes per face) are gone.
//   * local ids come from ballot prefix sums, not atomics: the record is bitwise reproducible.
template <class T> struct PatchSmem2 {
    static constexpr int FR = T::MAXF <= 96 ? 32 : 64; // frontier ring: faces accepted but not yet expanded (power of two)
    int fhKey[T::HASHF];
    int vhKey[T::HASHV];
    unsigned char fhVal[T::HASHF];
    unsigned char vhVal[T::HASHV];
    int4 frAdj[FR], frOpp[FR];       // adjopp entry of the frontier faces, by local face id & (FR - 1)
    unsigned char fin[T::MAXF];      // per local face: bit k = corner k lies within the cut-off
    alignas(16) unsigned char rec[T::BYTES];
};

__device__ __forceinline__ int pick3(const int4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

template <class T> __device__ int buildPatch2(const PatchArgs& a, PatchSmem2<T>& s, int li, int lane)
{
    constexpr int HF = T::HASHF, HV = T::HASHV, FRM = PatchSmem2<T>::FR - 1;
    const int gi = a.minIdx + li;
    int* tIdx = reinterpret_cast<int*>(s.rec + T::OFF_TIDX);
    int* gface = reinterpret_cast<int*>(s.rec + T::OFF_GFACE);
    int* gvert = reinterpret_cast<int*>(s.rec + T::OFF_GVERT);
    unsigned char* tFace = s.rec + T::OFF_TFACE;
    unsigned char* velig = s.rec + T::OFF_VELIG;
    unsigned* fvert = reinterpret_cast<unsigned*>(s.rec + T::OFF_FVERT);
    unsigned* fadj = reinterpret_cast<unsigned*>(s.rec + T::OFF_FADJ);
    int* hdr = reinterpret_cast<int*>(s.rec);
    const unsigned ltMask = (1u << lane) - 1u;

    const int sf = a.face[gi];
    const d3 sp{a.eucl[3 * gi], a.eucl[3 * gi + 1], a.eucl[3 * gi + 2]};

    // ---------------- 1. ordered candidates (identical to version 1) ----------------
    int K = 0;
    double R;
    {
        const CellGrid& g = a.grid;
        // the source's cell is the one k_euclid_cell binned it into; decoded with exact reciprocal multiplications
        // ((c + 0.5) / d is at least 0.5 / d away from an integer, far more than the rounding of the product)
        const int c0 = a.cellOf[gi], nxy = g.n[0] * g.n[1];
        const int iz = (int)(((double)c0 + 0.5) * a.invNxy), r0 = c0 - iz * nxy;
        const int iy = (int)(((double)r0 + 0.5) * a.invNx), ix = r0 - iy * g.n[0];
        // lane -> cell of the 3 x 3 x 3 stencil, xx outer, yy, zz inner; cells outside the grid drop out, which leaves the
        // remaining ones in the order of hyperRectangularCellList::getCellNeighbors (clipped loops)
        const int xx = ix + lane / 9 - 1, yy = iy + (lane / 3) % 3 - 1, zz = iz + lane % 3 - 1;
        int s0 = 0, s1 = 0;
        if (lane < 27 && xx >= 0 && xx < g.n[0] && yy >= 0 && yy < g.n[1] && zz >= 0 && zz < g.n[2]) {
            const int c = xx + yy * g.n[0] + zz * nxy;
            s0 = a.cellStart[c];
            s1 = s0 + a.cellCount[c];
        }
        int mine = 0, h0 = -1, h1 = -1, h2 = -1, h3 = -1;
        double maxd2 = 0;
        for (int q = s0; q < s1; ++q) {
            int j = a.cellItems[q];
            if (j == gi) continue;
            d3 p{a.eucl[3 * j], a.eucl[3 * j + 1], a.eucl[3 * j + 2]};
            double d2 = xsqlen(xsub3(sp, p));
            if (d2 < g.range2) {
                if (mine == 0) h0 = j;
                else if (mine == 1) h1 = j;
                else if (mine == 2) h2 = j;
                else if (mine == 3) h3 = j;
                mine++;
                maxd2 = d2 > maxd2 ? d2 : maxd2;
            }
        }
        int incl = warpInclusiveScan(mine, lane);
        K = __shfl_sync(FULL, incl, 31);
        if (K > a.kmax) { // neighbour stride too small: raises the stride guard (common.cuh), the host regrows and repeats the phase
            if (lane == 0) atomicMax(a.counters + C_KMAX_NEED, (unsigned long long)K), atomicAdd(a.counters + C_KMAX_OVERFLOW, 1ull);
            return 1;
        }
        if (K > T::RECK) return 2; // 1 + reason (0 candidates, 1 faces, 2 vertices)
        { // maximum of non-negative doubles = maximum of their bit patterns: two 32-bit warp reductions
            const unsigned hi = (unsigned)__double2hiint(maxd2), lo = (unsigned)__double2loint(maxd2);
            const unsigned mh = __reduce_max_sync(FULL, hi), ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
            maxd2 = __hiloint2double((int)mh, (int)ml);
        }
        R = xsqrt(maxd2);
        int pos = incl - mine;
        if (mine <= 4) {
            if (mine > 0) tIdx[pos] = h0;
            if (mine > 1) tIdx[pos + 1] = h1;
            if (mine > 2) tIdx[pos + 2] = h2;
            if (mine > 3) tIdx[pos + 3] = h3;
        } else {
            for (int q = s0; q < s1; ++q) {
                int j = a.cellItems[q];
                if (j == gi) continue;
                d3 p{a.eucl[3 * j], a.eucl[3 * j + 1], a.eucl[3 * j + 2]};
                if (xsqlen(xsub3(sp, p)) < g.range2) tIdx[pos++] = j;
            }
        }
    }
    if (lane == 0) hdr[2] = K;
    if (K == 0) {
        if (lane == 0) hdr[0] = 0, hdr[1] = 0, hdr[3] = 0;
        return 0;
    }
    double thr2 = __longlong_as_double(0x7ff0000000000000LL);
    if (a.submeshing) { // triangulatedMeshSpace::distanceWithSubmeshing :167-169
        double thr = a.maxDist;
        if (R < a.maxDist) thr = R;
        thr2 = xmul(thr, thr);
    }

    // ---------------- 2. the source face ----------------
    for (int h = lane; h < HF; h += 32) s.fhKey[h] = -1;
    for (int h = lane; h < HV; h += 32) s.vhKey[h] = -1;
    __syncwarp();
    const int myTF = lane < K ? a.face[tIdx[lane]] : sf; // K <= T::RECK <= 32: one target per lane
    {
        const int4 sc = __ldg(a.m.corner + sf);
        bool in = false;
        if (lane < 3) {
            const int gv = pick3(sc, lane);
            const double2* pv = reinterpret_cast<const double2*>(a.m.vert + gv);
            const double2 xy = __ldg(pv), zw = __ldg(pv + 1);
            in = !(xsqlen(xsub3(sp, d3{xy.x, xy.y, zw.x})) > thr2);
            bool isNew;
            s.vhVal[hashInsert(s.vhKey, HV - 1, gv, isNew)] = (unsigned char)lane;
            gvert[lane] = gv;
            velig[lane] = zw.y != 0.0;
        }
        const unsigned bits = __ballot_sync(FULL, in) & 7u;
        if (lane == 0) {
            const int4 sA = __ldg(a.m.adjopp + 2 * (size_t)sf), sO = __ldg(a.m.adjopp + 2 * (size_t)sf + 1);
            bool isNew;
            s.fhVal[hashInsert(s.fhKey, HF - 1, sf, isNew)] = 0;
            gface[0] = sf;
            fvert[0] = 0u | (1u << 8) | (2u << 16) | ((unsigned)(sA.w & 63) << 24);
            fadj[0] = 0x00FFFFFFu; // REC_NONE x 3
            s.fin[0] = (unsigned char)bits;
            s.frAdj[0] = sA, s.frOpp[0] = sO;
        }
    }
    __syncwarp();
    int nF = 1, nV = 3, head = 0;
    // ---------------- 3. flood fill with fused local indexing ----------------
    if (__any_sync(FULL, myTF != sf)) {
        bool noAdd = false; // all goal faces are among the source face and its neighbours: only link what is there (submesher.cpp:97-98)
        while (head < nF) {
            const int cnt = min(10, nF - head);
            const int slotF = lane / 3, k = lane - 3 * slotF;
            const bool active = slotF < cnt;
            const int i = head + slotF;
            int g = -1, d = -1, kk = 0;
            unsigned fvb = 0, fbits = 0;
            if (active) {
                const int4 A = s.frAdj[i & FRM], O = s.frOpp[i & FRM];
                g = pick3(A, k), d = pick3(O, k), kk = (A.w >> (2 * k)) & 3;
                fvb = fvert[i], fbits = s.fin[i];
            }
            const bool valid = active && g >= 0;
            int slot = valid ? hashFind(s.fhKey, HF - 1, g) : -1;
            const bool cand = valid && slot < 0 && !noAdd;
            const unsigned ina = (fbits >> ((k + 1) % 3)) & 1u, inb = (fbits >> ((k + 2) % 3)) & 1u;
            bool ind = false, elig = false, sad = false;
            int4 gA = make_int4(0, 0, 0, 0), gO = gA;
            if (cand) { // the one global round trip of the pass: the unseen vertex of g and, speculatively, g's own adjopp entry
                const double2* pv = reinterpret_cast<const double2*>(a.m.vert + d);
                const double2 xy = __ldg(pv), zw = __ldg(pv + 1);
                gA = __ldg(a.m.adjopp + 2 * (size_t)g), gO = __ldg(a.m.adjopp + 2 * (size_t)g + 1);
                ind = !(xsqlen(xsub3(sp, d3{xy.x, xy.y, zw.x})) > thr2);
                sad = zw.y != 0.0;
                elig = i == 0 || ina || inb || ind; // the neighbours of the source face are taken unconditionally (submesher.cpp:83-96)
            }
            bool isNew = false;
            if (elig) slot = hashInsert(s.fhKey, HF - 1, g, isNew);
            const bool win = elig && isNew;
            const unsigned bal = __ballot_sync(FULL, win);
            const int nAdd = __popc(bal);
            if (nF + nAdd > T::MAXF || nF + nAdd - (head + cnt) > PatchSmem2<T>::FR) return 3;
            const int id = nF + __popc(bal & ltMask);
            int vs = -1;
            bool vnew = false;
            __syncwarp(); // the frontier slots read at the top of the pass may be the ones reused below (ring of FR entries)
            if (win) {
                s.fhVal[slot] = (unsigned char)id;
                gface[id] = g;
                s.frAdj[id & FRM] = gA, s.frOpp[id & FRM] = gO;
                vs = hashInsert(s.vhKey, HV - 1, d, vnew);
            }
            const unsigned vbal = __ballot_sync(FULL, win && vnew);
            if (nV + __popc(vbal) > T::MAXV) return 4;
            if (win && vnew) {
                const int vid = nV + __popc(vbal & ltMask);
                s.vhVal[vs] = (unsigned char)vid;
                gvert[vid] = d;
                velig[vid] = sad;
            }
            __syncwarp();
            if (active) reinterpret_cast<unsigned char*>(fadj + i)[k] = (valid && slot >= 0) ? s.fhVal[slot] : (unsigned char)REC_NONE;
            if (win) { // g's corner kk is the new vertex; its edge kk is the shared edge, seen from the other side
                const unsigned ld = s.vhVal[vs], la = (fvb >> (8 * ((k + 1) % 3))) & 0xFFu, lb = (fvb >> (8 * ((k + 2) % 3))) & 0xFFu;
                fvert[id] = (ld << (8 * kk)) | (lb << (8 * ((kk + 1) % 3))) | (la << (8 * ((kk + 2) % 3))) | ((unsigned)(gA.w & 63) << 24);
                s.fin[id] = (unsigned char)(((unsigned)ind << kk) | (inb << ((kk + 1) % 3)) | (ina << ((kk + 2) % 3)));
            }
            __syncwarp();
            nF += nAdd, nV += __popc(vbal);
            if (head == 0 && !__any_sync(FULL, hashFind(s.fhKey, HF - 1, myTF) < 0)) noAdd = true;
            head += cnt;
        }
        // leftover goal faces: inside the cut-off sphere but with no vertex inside it (submesher.cpp:143-144); rare
        unsigned miss = __ballot_sync(FULL, hashFind(s.fhKey, HF - 1, myTF) < 0);
        while (miss) {
            const int src = __ffs(miss) - 1;
            miss &= miss - 1;
            const int g = __shfl_sync(FULL, myTF, src);
            if (hashFind(s.fhKey, HF - 1, g) >= 0) continue; // two targets in the same leftover face
            if (nF >= T::MAXF) return 3;
            const int4 c = __ldg(a.m.corner + g), A = __ldg(a.m.adjopp + 2 * (size_t)g);
            int vs = -1;
            bool vnew = false, sad = false;
            if (lane < 3) {
                const int gv = pick3(c, lane);
                sad = __ldg(reinterpret_cast<const double2*>(a.m.vert + gv) + 1).y != 0.0;
                vs = hashInsert(s.vhKey, HV - 1, gv, vnew);
            }
            const unsigned vbal = __ballot_sync(FULL, vnew);
            if (nV + __popc(vbal) > T::MAXV) return 4;
            if (vnew) {
                const int vid = nV + __popc(vbal & ltMask);
                s.vhVal[vs] = (unsigned char)vid;
                gvert[vid] = pick3(c, lane);
                velig[vid] = sad;
            }
            __syncwarp();
            unsigned myv = lane < 3 ? s.vhVal[vs] : 0u, mya = REC_NONE;
            if (lane < 3) { // adjacency = the mesh adjacency restricted to the face set, in both directions
                const int nb = pick3(A, lane), sl = nb < 0 ? -1 : hashFind(s.fhKey, HF - 1, nb);
                if (sl >= 0) {
                    mya = s.fhVal[sl];
                    reinterpret_cast<unsigned char*>(fadj + mya)[(A.w >> (2 * lane)) & 3] = (unsigned char)nF;
                }
            }
            const unsigned v0 = __shfl_sync(FULL, myv, 0), v1 = __shfl_sync(FULL, myv, 1), v2 = __shfl_sync(FULL, myv, 2);
            const unsigned a0 = __shfl_sync(FULL, mya, 0), a1 = __shfl_sync(FULL, mya, 1), a2 = __shfl_sync(FULL, mya, 2);
            if (lane == 0) {
                bool isNew;
                s.fhVal[hashInsert(s.fhKey, HF - 1, g, isNew)] = (unsigned char)nF;
                gface[nF] = g;
                fvert[nF] = v0 | (v1 << 8) | (v2 << 16) | ((unsigned)(A.w & 63) << 24);
                fadj[nF] = a0 | (a1 << 8) | (a2 << 16);
            }
            __syncwarp();
            nF += 1, nV += __popc(vbal);
        }
    }
    __syncwarp();
    // ---------------- 4. patch-border vertices may act as pseudo-sources; targets -> local faces ----------------
    for (int f = lane; f < nF; f += 32) {
        const unsigned fa = fadj[f], fv = fvert[f];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (((fa >> (8 * k)) & 0xFFu) == REC_NONE) { // border edge k of the patch: its endpoints are corners k+1, k+2
                velig[(fv >> (8 * ((k + 1) % 3))) & 0xFFu] = 1;
                velig[(fv >> (8 * ((k + 2) % 3))) & 0xFFu] = 1;
            }
    }
    if (lane < K) tFace[lane] = s.fhVal[hashFind(s.fhKey, HF - 1, myTF)];
    if (lane == 0) hdr[0] = nF, hdr[1] = nV, hdr[3] = 0;
    return 0;
}

} // namespace

template <class T> using PatchWs = PatchSmem2<T>;
#define BUILD_PATCH buildPatch2

template <class T> __global__ void __launch_bounds__(PATCH_THREADS) k_patch(PatchArgs a)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    PatchWs<T>& s = reinterpret_cast<PatchWs<T>*>(smemRaw)[wib];
    unsigned long long nRetry = 0;
    PDL_ENTRY();
    if (strideGuardUp(a.counters)) return; // the cell-list build found a stencil fuller than the neighbour stride (common.cuh)
    const int nWork = a.srcList ? *a.srcCount : a.nLocal;
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(a.workCounter, 1);
        w = __shfl_sync(FULL, w, 0);
        if (w >= nWork) break;
        const int li = a.srcList ? a.srcList[w] : w;
        int* hdr = reinterpret_cast<int*>(s.rec);
        if (!a.recordByParticle && w >= a.maxRecords) { // no record slot left in this tier: hand the source on
            if (lane == 0) {
                int r = atomicAdd(a.retryCount, 1);
                a.retryList[r] = li;
                nRetry++;
            }
            continue;
        }
        int st = BUILD_PATCH<T>(a, s, li, lane);
        st = __shfl_sync(FULL, st, 0);
        __syncwarp();
        if (st != 0 && lane == 0) {
            hdr[0] = 0, hdr[1] = 0, hdr[2] = 0, hdr[3] = 1; // stage 2 skips this source
            int r = atomicAdd(a.retryCount, 1);
            a.retryList[r] = li;
            nRetry++;
            if (st >= 2) atomicAdd(a.counters + C_OVF_REASON + st - 2, 1ull);
        }
        __syncwarp();
        // coalesced record store, section by section and only as far as each section is used (a record is 1.6 kB apart, ~0.8 kB
        // of it carries data at the design point); what lies beyond is never read
        const int4* src = reinterpret_cast<const int4*>(s.rec);
        int4* dst = reinterpret_cast<int4*>(a.records + (size_t)(a.recordByParticle ? li : w) * T::BYTES);
        const int nF = hdr[0], nV = hdr[1];
        if (hdr[3] || nF == 0) {
            for (int q = lane; q < T::OFF_TFACE / 16; q += 32) dst[q] = src[q]; // header and candidate ids
        } else {
            static_assert(T::OFF_VELIG % 16 == 0 && T::OFF_GFACE % 16 == 0 && T::OFF_GVERT % 16 == 0 && T::OFF_FADJ % 16 == 0, "sections start on 16 bytes");
            for (int q = lane; q < T::OFF_VELIG / 16; q += 32) dst[q] = src[q]; // header, candidate ids, target faces
            for (int q = lane; q < (nV + 15) / 16; q += 32) dst[T::OFF_VELIG / 16 + q] = src[T::OFF_VELIG / 16 + q];
            for (int q = lane; q < (nF + 3) / 4; q += 32) {
                dst[T::OFF_GFACE / 16 + q] = src[T::OFF_GFACE / 16 + q];
                dst[T::OFF_FVERT / 16 + q] = src[T::OFF_FVERT / 16 + q];
                dst[T::OFF_FADJ / 16 + q] = src[T::OFF_FADJ / 16 + q];
            }
            for (int q = lane; q < (nV + 3) / 4; q += 32) dst[T::OFF_GVERT / 16 + q] = src[T::OFF_GVERT / 16 + q];
        }
        __syncwarp();
    }
    if (lane == 0 && nRetry) atomicAdd(a.counters + C_TIER_RETRY, nRetry);
}

template <class T> cudaError_t launchPatch(cudaStream_t st, const PatchArgs& a, int numSMs)
{
    size_t smem = sizeof(PatchWs<T>) * (PATCH_THREADS / 32);
    static int perSMdev[64] = {0}; // the attribute and the occupancy are per device
    int dev = 0;
    cudaGetDevice(&dev);
    int& perSM = perSMdev[dev & 63];
    if (!perSM) {
        cudaFuncSetAttribute(k_patch<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_patch<T>, PATCH_THREADS, smem) != cudaSuccess || perSM < 1) perSM = 1;
    }
    // persistent warps pulling sources from a work counter: one wave of resident blocks (a retry tier whose list
    // length is only known on the device gets one block per SM)
    int blocks = a.srcList ? numSMs : min(numSMs * perSM, max(1, (a.nLocal + PATCH_THREADS / 32 - 1) / (PATCH_THREADS / 32)));
    return launchStep(k_patch<T>, blocks, PATCH_THREADS, smem, st, a);
}
template cudaError_t launchPatch<TierSmall>(cudaStream_t, const PatchArgs&, int);
template cudaError_t launchPatch<TierLarge>(cudaStream_t, const PatchArgs&, int);

} // namespace css
