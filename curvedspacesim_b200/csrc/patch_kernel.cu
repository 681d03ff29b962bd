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

template <class T> struct PatchSmem { // per-warp scratch
    int fhKey[T::HASHF];             // global face id -> slot
    int vhKey[T::HASHV];             // global vertex id -> slot
    unsigned char fhVal[T::HASHF];   // slot -> local face id
    unsigned char vhVal[T::HASHV];   // slot -> local vertex id
    int misc[4];                     // nF, nV, overflow
    alignas(16) unsigned char rec[T::BYTES];
};

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
__device__ __forceinline__ double warpMaxD(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// returns 0 ok, 1 overflow (source goes to the retry tiers)
template <class T> __device__ int buildPatch(const PatchArgs& a, PatchSmem<T>& s, int li, int lane)
{
    constexpr int HF = T::HASHF, HV = T::HASHV;
    const int gi = a.minIdx + li;
    int* tIdx = reinterpret_cast<int*>(s.rec + T::OFF_TIDX);
    int* gface = reinterpret_cast<int*>(s.rec + T::OFF_GFACE);
    int* gvert = reinterpret_cast<int*>(s.rec + T::OFF_GVERT);
    unsigned char* tFace = s.rec + T::OFF_TFACE;
    unsigned char* velig = s.rec + T::OFF_VELIG;
    uchar4* fvert = reinterpret_cast<uchar4*>(s.rec + T::OFF_FVERT);
    uchar4* fadj = reinterpret_cast<uchar4*>(s.rec + T::OFF_FADJ);
    int* hdr = reinterpret_cast<int*>(s.rec);

    const int sf = a.face[gi];
    const d3 sp{a.eucl[3 * gi], a.eucl[3 * gi + 1], a.eucl[3 * gi + 2]};

    // ---------------- 1. ordered candidates ----------------
    int K = 0;
    double R;
    {
        const CellGrid& g = a.grid;
        int ix = cellCoord(g, sp.x, 0), iy = cellCoord(g, sp.y, 1), iz = cellCoord(g, sp.z, 2);
        int x0 = max(0, ix - 1), x1 = min(g.n[0] - 1, ix + 1);
        int y0 = max(0, iy - 1), y1 = min(g.n[1] - 1, iy + 1);
        int z0 = max(0, iz - 1), z1 = min(g.n[2] - 1, iz + 1);
        int ny = y1 - y0 + 1, nz = z1 - z0 + 1, ncell = (x1 - x0 + 1) * ny * nz;
        int s0 = 0, s1 = 0;
        if (lane < ncell) { // stencil order: xx outer, yy, zz inner
            int xx = x0 + lane / (ny * nz), rem = lane % (ny * nz);
            int yy = y0 + rem / nz, zz = z0 + rem % nz;
            int c = xx + yy * g.n[0] + zz * g.n[0] * g.n[1];
            s0 = a.cellStart[c];
            s1 = s0 + a.cellCount[c];
        }
        // one pass: each lane keeps the first few hits of its cell in registers (cells hold ~0.3 particles
        // on average at the target densities); a second pass over the cell is taken only when it has more
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
        if (K > a.kmax) { // neighbour stride too small: the host doubles it and reruns the step
            if (lane == 0) atomicMax(a.counters + C_KMAX_NEED, (unsigned long long)K), atomicAdd(a.counters + C_KMAX_OVERFLOW, 1ull);
            return 1;
        }
        if (K > T::RECK) return 2; // 1 + reason (0 candidates, 1 faces, 2 vertices)
        maxd2 = warpMaxD(maxd2);
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

    // ---------------- 2. flood fill ----------------
    for (int h = lane; h < HF; h += 32) s.fhKey[h] = -1;
    for (int h = lane; h < HV; h += 32) s.vhKey[h] = -1;
    if (lane == 0) s.misc[0] = 0, s.misc[1] = 0, s.misc[2] = 0;
    __syncwarp();
    auto addFace = [&](int g) {
        bool isNew;
        int slot = hashInsert(s.fhKey, HF - 1, g, isNew);
        if (isNew) {
            int id = atomicAdd(&s.misc[0], 1);
            if (id < T::MAXF) {
                gface[id] = g;
                s.fhVal[slot] = (unsigned char)id;
            } else
                s.misc[2] = 1;
        }
    };
    int myTF = lane < K ? a.face[tIdx[lane]] : sf; // K <= T::RECK <= 32: one target per lane
    if (lane == 0) addFace(sf);
    __syncwarp();
    if (__any_sync(FULL, myTF != sf)) {
        int4 sadj = __ldg(a.m.adj + sf);
        if (lane < 3) {
            int g = lane == 0 ? sadj.x : (lane == 1 ? sadj.y : sadj.z);
            if (g >= 0) addFace(g);
        }
        __syncwarp();
        if (__any_sync(FULL, hashFind(s.fhKey, HF - 1, myTF) < 0)) {
            // frontier faces [head, tail) x 3 edges, one (face, edge) pair per lane: 10 faces per pass
            int head = 1;
            for (;;) {
                __syncwarp();
                // one lane's view of the queue for the whole warp: a lane that read it later could already see pushes of this
                // pass, and lanes disagreeing on `head` would leave the loop at different times
                const int tail = min(__shfl_sync(FULL, s.misc[0], 0), T::MAXF);
                if (__shfl_sync(FULL, s.misc[2], 0)) return 3;
                if (head >= tail) break;
                int slotF = lane / 3, k = lane - 3 * slotF;
                int idx = head + slotF;
                if (slotF < 10 && idx < tail) {
                    int4 ad = __ldg(a.m.adj + gface[idx]);
                    int g = k == 0 ? ad.x : (k == 1 ? ad.y : ad.z);
                    if (g >= 0 && hashFind(s.fhKey, HF - 1, g) < 0) {
                        int4 c = __ldg(a.m.corner + g);
                        d3 p0 = ldvert(a.m, c.x), p1 = ldvert(a.m, c.y), p2 = ldvert(a.m, c.z);
                        bool far = xsqlen(xsub3(sp, p0)) > thr2 && xsqlen(xsub3(sp, p1)) > thr2 && xsqlen(xsub3(sp, p2)) > thr2;
                        if (!far) {
                            if (s.misc[0] >= T::MAXF) s.misc[2] = 1;
                            else addFace(g);
                        }
                    }
                }
                head = min(head + 10, tail);
            }
            if (hashFind(s.fhKey, HF - 1, myTF) < 0) { // leftover goal faces (submesher.cpp:143-144)
                if (s.misc[0] >= T::MAXF) s.misc[2] = 1;
                else addFace(myTF);
            }
        }
    }
    __syncwarp();
    if (s.misc[2] || s.misc[0] > T::MAXF) return 3;
    const int nF = s.misc[0];

    // ---------------- 3. local indexing ----------------
    for (int f = lane; f < nF; f += 32) {
        int4 c = __ldg(a.m.corner + gface[f]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int gv = k == 0 ? c.x : (k == 1 ? c.y : c.z);
            bool isNew;
            int slot = hashInsert(s.vhKey, HV - 1, gv, isNew);
            if (isNew) {
                int id = atomicAdd(&s.misc[1], 1);
                if (id < T::MAXV) {
                    gvert[id] = gv;
                    s.vhVal[slot] = (unsigned char)id;
                } else
                    s.misc[2] = 1;
            }
        }
        if (s.misc[2]) break; // the table is sized 2 x T::MAXV + 3 x 32 in-flight inserts: it cannot fill up before this trips
    }
    __syncwarp();
    if (s.misc[2] || s.misc[1] > T::MAXV) return 4;
    const int nV = s.misc[1];
    for (int v = lane; v < nV; v += 32) velig[v] = a.m.saddle[gvert[v]];
    __syncwarp();
    for (int f = lane; f < nF; f += 32) {
        int gf = gface[f];
        int4 c = __ldg(a.m.corner + gf);
        int4 ad = __ldg(a.m.adj + gf);
        unsigned char lv[3], la[3];
        lv[0] = s.vhVal[hashFind(s.vhKey, HV - 1, c.x)];
        lv[1] = s.vhVal[hashFind(s.vhKey, HV - 1, c.y)];
        lv[2] = s.vhVal[hashFind(s.vhKey, HV - 1, c.z)];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int g = k == 0 ? ad.x : (k == 1 ? ad.y : ad.z);
            int sl = g < 0 ? -1 : hashFind(s.fhKey, HF - 1, g);
            la[k] = sl < 0 ? (unsigned char)REC_NONE : s.fhVal[sl];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (la[k] == REC_NONE) { // patch border edge k: its endpoints are corners k+1, k+2
                velig[lv[(k + 1) % 3]] = 1;
                velig[lv[(k + 2) % 3]] = 1;
            }
        fvert[f] = make_uchar4(lv[0], lv[1], lv[2], (unsigned char)(ad.w & 63));
        fadj[f] = make_uchar4(la[0], la[1], la[2], 0);
    }
    if (lane < K) tFace[lane] = s.fhVal[hashFind(s.fhKey, HF - 1, myTF)];
    if (lane == 0) hdr[0] = nF, hdr[1] = nV, hdr[3] = 0;
    return 0;
}

} // namespace

template <class T> __global__ void __launch_bounds__(PATCH_THREADS) k_patch(PatchArgs a)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    PatchSmem<T>& s = reinterpret_cast<PatchSmem<T>*>(smemRaw)[wib];
    unsigned long long nRetry = 0;
    if (strideGuardUp(a.counters)) return; // the cell-list build found a stencil fuller than the neighbour stride (common.cuh)
    const int nWork = a.srcList ? *a.srcCount : a.nLocal;
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(a.workCounter, 1);
        w = __shfl_sync(FULL, w, 0);
        if (w >= nWork) break;
        const int li = a.srcList ? a.srcList[w] : w;
        int* hdr = reinterpret_cast<int*>(s.rec);
        if (w >= a.maxRecords) { // no record slot left in this tier: hand the source on
            if (lane == 0) {
                int r = atomicAdd(a.retryCount, 1);
                a.retryList[r] = li;
                nRetry++;
            }
            continue;
        }
        int st = buildPatch<T>(a, s, li, lane);
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
        // coalesced record store; the unused tail of a record is never read
        const int4* src = reinterpret_cast<const int4*>(s.rec);
        int4* dst = reinterpret_cast<int4*>(a.records + (size_t)w * T::BYTES);
        int nF = hdr[0];
        int used = hdr[3] ? 1 : (nF == 0 ? T::OFF_TFACE / 16 : T::BYTES / 16);
        for (int q = lane; q < used; q += 32) dst[q] = src[q];
        __syncwarp();
    }
    if (lane == 0 && nRetry) atomicAdd(a.counters + C_TIER_RETRY, nRetry);
}

template <class T> cudaError_t launchPatch(cudaStream_t st, const PatchArgs& a, int numSMs)
{
    size_t smem = sizeof(PatchSmem<T>) * (PATCH_THREADS / 32);
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
    k_patch<T><<<blocks, PATCH_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}
template cudaError_t launchPatch<TierSmall>(cudaStream_t, const PatchArgs&, int);
template cudaError_t launchPatch<TierLarge>(cudaStream_t, const PatchArgs&, int);

} // namespace css
