// Host-side launch interface between css_api.cu and the kernel translation units.
#pragma once
#include "common.cuh"
#include <algorithm>

namespace css {

#define REDUCE_MAX_BLOCKS 512

// per-source capacities of one geodesic tier (all counts per warp)
struct GeoCaps {
    int maxF, maxV, ring, kt, hashF, hashV; // ring, hashF, hashV are powers of two
};
size_t geoWorkspaceBytes(const GeoCaps& c);

struct GeoArgs {
    MeshDev m;
    CellGrid grid;
    int nTotal, nLocal, minIdx;
    const int* face;       // [nTotal] replicated positions
    const double* bary;    // [3 nTotal]
    const double* eucl;    // [3 nTotal]
    const int* cellStart;  // nullptr -> all-to-all candidates (baseNeighborStructure)
    const int* cellCount;  // cell c holds cellItems[cellStart[c] .. cellStart[c] + cellCount[c])
    const int* cellItems;
    int submeshing;
    double maxDist;
    // work: sources are srcList[0..*srcCount) (local indices) or 0..nLocal-1 when srcList == nullptr
    const int* srcList;
    const int* srcCount;
    int* workCounter;
    // explicit single-source query (css_distance); xK < 0 when unused
    int xK, xSrcFace;
    double xSrcBary[3], xThreshold;
    const int* xTgtFace;
    const double* xTgtBary;
    // outputs, fixed stride kmax per local particle
    int kmax;
    int* nbrCount;
    int* nbrIdx;
    double* nbrDist;
    double* nbrTs;
    double* nbrTe; // nullable
    // fused force (+ optional velocity kick v += kick * f)
    int forceMode; // 0 none, 1 accumulate into frc
    ForceParams fp;
    int zero;
    double* frc;
    double kick;
    double* vel;
    // retry list for the next tier
    int* retryList;
    int* retryCount;
    unsigned long long* counters;
    GeoCaps caps;
    char* gws; // global workspace (nullptr -> shared memory)
    int lastTier;
};

cudaError_t launchGeodesicCta(cudaStream_t st, const GeoArgs& a, int blocks); // one CTA per source (long-range tiers)

// ---- two-stage tier-0 path: patch records (patch_kernel.cu) -> window propagation (window_kernel.cu) ----
// Fixed-capacity patch record, one per local source particle, REC_BYTES apart (all offsets 16-byte aligned):
//   int   hdr[4]            nF, nV, K, status (1 = overflow: handled by the retry tiers, stage 2 skips it)
//   int   tIdx[RECK]        ordered candidate particle indices (the neighbour list)
//   u8    tFace[RECK]       local face of each target
//   u8    velig[REC_MAXV]   vertex may act as a pseudo-source (saddle of the mesh or on the patch border)
//   int   gface[REC_MAXF]   local face -> global face
//   int   gvert[REC_MAXV]   local vertex -> global vertex
//   uchar4 fvert[REC_MAXF]  local corner ids + kk bits (index of each edge inside the neighbour, 2 bits per edge)
//   uchar4 fadj[REC_MAXF]   local face across the edge opposite corner k (REC_NONE = patch border)
// Two capacity classes are instantiated: SMALL (tier 0, covers the design-point configs) and LARGE (tier 1: big patches
// or many candidates; also the fast path of coarse meshes such as config 1).
// K_ targets are propagated at a time (the per-target workspace of stage 2); a record holds up to RECK_ >= K_ candidates and
// stage 2 runs one propagation per group of K_ (sources with more than K_ candidates are ~5e-6 of config 5, and a second
// propagation inside the main launch is far cheaper than a lone warp in a retry launch).
template <int F_, int V_, int K_, int RING_, int RECK_ = K_> struct GeoTier {
    static constexpr int MAXF = F_, MAXV = V_, MAXK = K_, RING = RING_, RECK = RECK_; // faces < 255, RECK <= 32 (one target per lane in stage 1)
    static constexpr int OFF_TIDX = 16;
    static constexpr int OFF_TFACE = OFF_TIDX + 4 * RECK_;
    static constexpr int OFF_VELIG = OFF_TFACE + RECK_;
    static constexpr int OFF_GFACE = OFF_VELIG + V_;
    static constexpr int OFF_GVERT = OFF_GFACE + 4 * F_;
    static constexpr int OFF_FVERT = OFF_GVERT + 4 * V_;
    static constexpr int OFF_FADJ = OFF_FVERT + 4 * F_;
    static constexpr int BYTES = OFF_FADJ + 4 * F_;
    static constexpr int HASHF = F_ <= 96 ? 256 : 512, HASHV = V_ <= 64 ? 256 : 512; // >= 2 x capacity + in-flight inserts
    static_assert(BYTES % 16 == 0 && OFF_FVERT % 16 == 0 && OFF_GFACE % 4 == 0, "record sections must stay aligned");
    static_assert(F_ < 255 && V_ < 256 && K_ <= RECK_ && RECK_ <= 32 && (RING_ & (RING_ - 1)) == 0, "8-bit local ids, one target per lane");
};
#ifndef CSS_T0_K
#define CSS_T0_K 16
#endif
#ifndef CSS_T0_RING
#define CSS_T0_RING 64
#endif
#ifndef CSS_T0_RECK
#define CSS_T0_RECK 32
#endif
using TierSmall = GeoTier<96, 64, CSS_T0_K, CSS_T0_RING, CSS_T0_RECK>;
using TierLarge = GeoTier<240, 160, 32, 256>;
// Tier 0 as the two-sources-per-warp kernel sees it (window_half_kernel.cu): the SAME record layout as TierSmall, a slimmer
// per-source workspace (8 targets per propagation, 32-window ring) so that twice as many sources are resident per SM.
#ifndef CSS_TH_K
#define CSS_TH_K 8
#endif
#ifndef CSS_TH_RING
#define CSS_TH_RING 32
#endif
using TierHalf = GeoTier<96, 64, CSS_TH_K, CSS_TH_RING, CSS_T0_RECK>;
static_assert(TierHalf::BYTES == TierSmall::BYTES && TierHalf::OFF_FVERT == TierSmall::OFF_FVERT, "TierHalf reads TierSmall records");
#define REC_NONE 255
#define PATCH_THREADS 256

// Static face stencils (stencil_kernel.cu): per mesh face, the superset of every patch a source lying in that face can get.
// Records are STENCIL_BYTES apart; inside a record the sections are packed back to back, so a source reads (and prefetches)
// one contiguous run of 16 + 12 nSF + 4 nSV bytes:
//   int4  hdr              nSF, nSV, status, 0
//   int   gface[nSF]       stencil face -> global face; entry 0 is the face itself, entries 1.. ascend
//   u32   fvert[nSF]       stencil corner ids v0 | v1 << 8 | v2 << 16 | kk bits << 24
//   u32   fadj[nSF]        stencil face across edge k (REC_NONE = not in the stencil) n0 | n1 << 8 | n2 << 16 | parent << 24
//   int   gvert[nSV]       stencil vertex -> global vertex
// A separate table holds nSF | nSV << 16 per face (0 = the face has no stencil: more than STENCIL_F faces / STENCIL_V vertices).
#define STENCIL_F 160 /* multiples of 32; ids below REC_NONE */
#define STENCIL_V 128
#define STENCIL_BYTES (16 + 12 * STENCIL_F + 4 * STENCIL_V)

struct PatchArgs {
    MeshDev m;
    CellGrid grid;
    int nLocal, minIdx;
    const int* face;      // [nTotal]
    const double* eucl;   // [3 nTotal]
    const int* cellStart;
    const int* cellCount;
    const int* cellItems;
    const int* cellOf;    // [nTotal] linear cell index of every particle (k_euclid_cell)
    double invNx, invNxy; // 1 / grid.n[0], 1 / (grid.n[0] grid.n[1])
    int submeshing;
    double maxDist;
    int kmax;
    // work: local particle indices srcList[0..*srcCount) (records indexed by list position), or 0..nLocal-1 when null
    const int* srcList;
    const int* srcCount;
    int maxRecords;       // capacity of `records`; list entries beyond it go straight to the retry list
    int* workCounter;
    int* retryList;
    int* retryCount;
    unsigned long long* counters;
    unsigned char* records; // [maxRecords][Tier::BYTES]
    const unsigned char* stencil; // [nF][STENCIL_BYTES] (launchPatchStencil only)
    const unsigned* stencilLen;   // [nF] nSF | nSV << 16, 0 = no stencil
    // launchPatchStencil: sources it cannot serve (face without a stencil, target outside the stencil) are listed here and
    // flood-filled by launchPatch into the SAME tier-0 records (recordByParticle: record index = particle, not list position)
    int* fallbackList;
    int* fallbackCount;
    int recordByParticle;
};
template <class Tier> cudaError_t launchPatch(cudaStream_t st, const PatchArgs& a, int numSMs);
// tier 0 through static face stencils (stencil_kernel.cu); statsDev[0] = faces without a stencil, statsDev[1] = sum of stencil sizes
size_t stencilBytes(int nF);
cudaError_t buildStencils(cudaStream_t st, const MeshDev& m, double maxDist, unsigned char* out, unsigned* lenOut, unsigned long long* statsDev, int numSMs);
template <class Tier> cudaError_t launchPatchStencil(cudaStream_t st, const PatchArgs& a, int numSMs);

struct WinArgs {
    MeshDev m;
    int nLocal, minIdx;
    const int* face;
    const double* bary;
    const double* eucl;
    const unsigned char* records;
    const int* srcList;
    const int* srcCount;
    int maxRecords;
    int submeshing;
    double maxDist;
    int kmax;
    int* nbrCount;
    int* nbrIdx;
    double* nbrDist;
    double* nbrTs;
    double* nbrTe; // nullable
    int forceMode;
    ForceParams fp;
    int zero;
    double* frc;
    double kick;
    double* vel;
    int* workCounter;
    int* retryList;
    int* retryCount;
    unsigned long long* counters;
    double* spill; // launchWindowsHalf only: windowsHalfSpillBytes(numSMs) bytes of scratch
};
template <class Tier> cudaError_t launchWindows(cudaStream_t st, const WinArgs& a, int warpsPerBlock, int numSMs, bool lean);
template <class Tier> cudaError_t launchWindowsHalf(cudaStream_t st, const WinArgs& a, int numSMs); // two sources per warp
size_t windowsHalfSpillBytes(int numSMs);
int geodesicMaxSmemPerBlock();

void launchEuclidCell(cudaStream_t st, const MeshDev& m, const CellGrid& g, int n, const int* face, const double* bary, double* eucl,
                      int* cellOf, int* cellCount, int* cellSlot, int* coarse = nullptr);
void launchCellBuild(cudaStream_t st, int n, int nCells, const int* cellOf, const int* cellSlot, int* cellCount, int* cellStart, int* tmpItems,
                     int* items, const int* coarse, const CellGrid& g, int kmax, unsigned long long* counters);
// ---- peer-memory position exchange, fused with the walker (replaces the all-gather after every move:
// mpiSimulation::synchronizeAndTransferBuffers, src/simulation/mpiSimulation.cpp:11-42).  Every rank owns an exchange
// window (cudaMalloc + CUDA IPC, mapped by all peers over NVLink/NVSwitch): a staging copy of the replicated position
// arrays indexed by GLOBAL particle index, and two rows of epoch flags.  The walker stores each new position into its
// own live arrays and into every peer's staging copy; k_peer_barrier publishes "rank r finished epoch e" in every
// window and waits for all ranks; k_peer_copy moves the other ranks' blocks from staging to the live arrays and
// publishes "rank r drained epoch e", which the next walker waits for before it overwrites the staging copies.
#define CSS_MAX_PEERS 8
#define CSS_PEER_FLAG_STRIDE 16            /* one 128-byte line per writer */
#define CSS_PEER_FLAG_BYTES 4096           /* arrived[8 x 16] | consumed[8 x 16] (u64), padded */
struct PeerWin {
    int n, rank; // n <= 1: no peer exchange
    int* face[CSS_MAX_PEERS];
    double* bary[CSS_MAX_PEERS];
    unsigned long long* flags[CSS_MAX_PEERS];
};
void launchPeerWaitConsumed(cudaStream_t st, const PeerWin& pw, const unsigned long long* epoch, unsigned long long* counters);
void launchPeerBarrier(cudaStream_t st, const PeerWin& pw, unsigned long long* epoch, unsigned long long* counters);
void launchPeerCopy(cudaStream_t st, const PeerWin& pw, int nTotal, int lo, int hi, int* face, double* bary, const unsigned long long* epoch,
                    unsigned* ticket);
void launchWalk(cudaStream_t st, const MeshDev& m, int n, int minIdx, int* face, double* bary, double* disp, double* vel, double* frc,
                int transportForce, int transportVelocity, int mode, double dt, int* flags, unsigned long long* counters, const PeerWin& pw);
void launchTransportGeneric(cudaStream_t st, const MeshDev& m, int n, int* face, double* bary, double* disp, int nVec, double* vecs,
                            int* flags);
void launchLocate(cudaStream_t st, const MeshDev& m, const double gmn[3], double h, const int gn[3], const int* cellStart, const int* cellFaces,
                  int n, const double* xyz, double clampTol, int* face, double* bary);
void launchAxpy(cudaStream_t st, int op, int n, double a, double b, double* vel, const double* frc, double* disp);
void launchNhChain(cudaStream_t st, double* nh, int M, const double* red, int takeKE, const unsigned long long* counters); // Nose-Hoover chain, device-resident
void launchScaleDev(cudaStream_t st, int n, const double* s, double* vel, const unsigned long long* counters);
void launchReduce(cudaStream_t st, int n, const double* vel, const double* frc, double* partial, double* out);
void launchEnergy(cudaStream_t st, int nLocal, int kmax, const int* nbrCount, const double* nbrDist, ForceParams fp, double* partial,
                  double* out);
void launchStress(cudaStream_t st, int nLocal, int kmax, const int* nbrCount, const double* nbrDist, const double* nbrTs, const double* vel,
                  ForceParams fp, double* partial, double* out); // out[18]: sum f (x) dr | sum v (x) v

double runMicrobench(cudaStream_t st, int numSMs, int what, int reps); // microbench.cu: 0 FP64 FMA TFLOP/s, 1 L2 read GB/s

} // namespace css
