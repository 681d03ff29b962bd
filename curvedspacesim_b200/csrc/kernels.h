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
