// Host-side launch interface between css_api.cu and the kernel translation units.
#pragma once
#include "common.cuh"

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

cudaError_t launchGeodesic(cudaStream_t st, const GeoArgs& a, int warpsPerBlock, int blocks);

// ---- two-stage tier-0 path: patch records (patch_kernel.cu) -> window propagation (window_kernel.cu) ----
// Fixed-capacity patch record, one per local source particle, REC_BYTES apart (all offsets 16-byte aligned):
//   int   hdr[4]            nF, nV, K, status (1 = overflow: handled by the retry tiers, stage 2 skips it)
//   int   tIdx[REC_MAXK]    ordered candidate particle indices (the neighbour list)
//   u8    tFace[REC_MAXK]   local face of each target
//   u8    velig[REC_MAXV]   vertex may act as a pseudo-source (saddle of the mesh or on the patch border)
//   int   gface[REC_MAXF]   local face -> global face
//   int   gvert[REC_MAXV]   local vertex -> global vertex
//   uchar4 fvert[REC_MAXF]  local corner ids + kk bits (index of each edge inside the neighbour, 2 bits per edge)
//   uchar4 fadj[REC_MAXF]   local face across the edge opposite corner k (REC_NONE = patch border)
#define REC_MAXF 96
#define REC_MAXV 64
#define REC_MAXK 16
#define REC_NONE 255
#define REC_OFF_TIDX 16
#define REC_OFF_TFACE (REC_OFF_TIDX + 4 * REC_MAXK)
#define REC_OFF_VELIG (REC_OFF_TFACE + REC_MAXK)
#define REC_OFF_GFACE (REC_OFF_VELIG + REC_MAXV)
#define REC_OFF_GVERT (REC_OFF_GFACE + 4 * REC_MAXF)
#define REC_OFF_FVERT (REC_OFF_GVERT + 4 * REC_MAXV)
#define REC_OFF_FADJ (REC_OFF_FVERT + 4 * REC_MAXF)
#define REC_BYTES (REC_OFF_FADJ + 4 * REC_MAXF)
#define PATCH_THREADS 256
#define WIN_RING 64

struct PatchArgs {
    MeshDev m;
    CellGrid grid;
    int nLocal, minIdx;
    const int* face;      // [nTotal]
    const double* eucl;   // [3 nTotal]
    const int* cellStart;
    const int* cellItems;
    int submeshing;
    double maxDist;
    int kmax;
    int* workCounter;
    int* retryList;
    int* retryCount;
    unsigned long long* counters;
    unsigned char* records; // [nLocal][REC_BYTES]
};
cudaError_t launchPatch(cudaStream_t st, const PatchArgs& a, int numSMs);
size_t patchSmemPerWarp();

struct WinArgs {
    MeshDev m;
    int nLocal, minIdx;
    const int* face;
    const double* bary;
    const double* eucl;
    const unsigned char* records;
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
};
cudaError_t launchWindows(cudaStream_t st, const WinArgs& a, int warpsPerBlock, int numSMs);
size_t windowSmemPerWarp();
int geodesicMaxSmemPerBlock();

void launchEuclidCell(cudaStream_t st, const MeshDev& m, const CellGrid& g, int n, const int* face, const double* bary, double* eucl,
                      int* cellOf, int* cellCount);
int scanBlocks(int nCells);
void launchCellBuild(cudaStream_t st, int n, int nCells, const int* cellOf, const int* cellCount, int* cellStart, int* blockSums,
                     int* fill, int* tmpItems, int* items);
void launchWalk(cudaStream_t st, const MeshDev& m, int n, int minIdx, int* face, double* bary, double* disp, double* vel, double* frc,
                int transportForce, int transportVelocity, int mode, double dt, int* flags, unsigned long long* counters);
void launchTransportGeneric(cudaStream_t st, const MeshDev& m, int n, int* face, double* bary, double* disp, int nVec, double* vecs,
                            int* flags);
void launchAxpy(cudaStream_t st, int op, int n, double a, double b, double* vel, const double* frc, double* disp);
void launchReduce(cudaStream_t st, int n, const double* vel, const double* frc, double* partial, double* out);
void launchEnergy(cudaStream_t st, int nLocal, int kmax, const int* nbrCount, const double* nbrDist, ForceParams fp, double* partial,
                  double* out);

} // namespace css
