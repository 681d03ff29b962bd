/* css_api.h — C ABI of the B200-native geodesic MD hot path (libcurvedspacesim_b200.so).
 *
 * Drop-in boundary for curvedSpaceSim's per-timestep path.  Each entry point names the reference
 * interface it replaces (paths relative to the reference repository root):
 *
 *   css_set_mesh            triangulatedMeshSpace::loadMeshFromFile + updateMeshSpanAndTree
 *                           (src/models/triangulatedMeshSpace.cpp:6-72)
 *   css_set_submeshing      triangulatedMeshSpace::useSubmeshingRoutines (src/models/triangulatedMeshSpace.h:60-66)
 *   css_set_boundary        openMeshSpace / absorbingOpenMeshSpace / tangentialOpenMeshSpace (src/models/openMeshSpace.cpp:114-238)
 *   css_set_cell_domain     cellListNeighborStructure ctor (src/utility/cellListNeighborStructure.cpp:4-20)
 *   css_euclidean           triangulatedMeshSpace::meshPositionToEuclideanLocation (.cpp:82-106)
 *   css_locate              simpleModel::R3PositionsToMeshPositions + clampBarycentricCoordinatesToFace
 *                           (src/models/simpleModel.cpp:114-154: PMP::locate_with_AABB_tree, then the clamp)
 *   css_distance            baseSpace::distance / triangulatedMeshSpace::distance (src/models/baseSpace.h:36-39,
 *                           triangulatedMeshSpace.cpp:155-238)
 *   css_transport           baseSpace::transportParticleAndVectors / displaceParticle (baseSpace.h:29-32,
 *                           triangulatedMeshSpace.cpp:428-564)
 *   css_set_state/get_state simpleModel public arrays positions/velocities/forces (src/models/simpleModel.h:66-79);
 *                           sharding arguments = mpiModel::determineIndexBounds (src/models/mpiModel.cpp:20-31)
 *   css_find_neighbors      simpleModel::findNeighbors / mpiModel::findNeighbors (simpleModel.cpp:68-112, mpiModel.cpp:75-122)
 *   css_get_neighbors       simpleModel::neighbors / neighborVectors / neighborDistances (simpleModel.h:72-79)
 *   css_compute_forces      force::computeForces with harmonicRepulsion / gaussianRepulsion
 *                           (src/forces/baseForce.cpp:12-28, harmonicRepulsion.cpp:19-33, gaussianRepulsion.cpp:8-12)
 *   css_compute_energy      force::computeEnergy (baseForce.cpp:33-44)
 *   css_compute_stress      simulation::computeMonodisperseStress (src/simulation/simulation.cpp:104-173)
 *   css_temperature         noseHooverNVT::getTemperatureFromKE (src/updaters/noseHooverNVT.cpp:141-150)
 *   css_move                simpleModel::moveParticles (simpleModel.cpp:44-66)
 *   css_gather_positions    mpiSimulation::synchronizeAndTransferBuffers (src/simulation/mpiSimulation.cpp:11-42)
 *   css_reduce              mpiSimulation::manipulateUpdaterData (mpiSimulation.cpp:69-89)
 *   css_step_nve_host       the same step with the model's host vectors as input and output (simpleModel.h:60-75 public arrays)
 *   css_step_nve            velocityVerletNVE::performUpdate (src/updaters/velocityVerletNVE.cpp:3-29)
 *   css_step_gd             gradientDescent::performUpdate (src/updaters/gradientDescent.cpp:6-18)
 *   css_nvt_init/step_nvt   noseHooverNVT ctor/setBathVariables/performUpdate (src/updaters/noseHooverNVT.cpp:3-139)
 *   css_fire_init/minimize  fireMinimization::setFIREParameters/minimizeByFire (src/updaters/fireMinimization.cpp:3-90)
 *   css_max_force/force_norm updater::getMaxForce/getForceNorm (src/updaters/baseUpdater.cpp:22-54)
 *
 * Conventions: every function returns 0 on success and a non-zero CSS_E* code otherwise (the text is
 * available through css_last_error); nothing throws across the ABI.  All pointers are HOST pointers
 * unless the name ends in _dev.  Barycentric coordinates and vectors are packed [n][3] doubles, face
 * indices int32.  Face corners must be given in the reference's corner order.
 * A context is bound to one CUDA device and one host thread at a time (same contract as the
 * reference, which is not re-entrant).  There is no CPU fallback: every call runs CUDA kernels.
 */
#ifndef CSS_API_H
#define CSS_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct css_ctx css_ctx;

enum css_status {
    CSS_OK = 0,
    CSS_EINVAL = 1,     /* bad argument */
    CSS_ECUDA = 2,      /* CUDA runtime error */
    CSS_EMESH = 3,      /* mesh is not a consistently oriented manifold triangle mesh */
    CSS_ECAPACITY = 4,  /* a per-source patch/window/candidate capacity was exceeded on every tier */
    CSS_ESTATE = 5,     /* call made before the required state was set */
    CSS_ENCCL = 6       /* NCCL error / communicator not initialised */
};

enum css_force_kind { CSS_FORCE_HARMONIC = 0, CSS_FORCE_GAUSSIAN = 1 };
/* params: HARMONIC {k, sigma, range}; GAUSSIAN {alpha, sigma, range} (range = force::maximumInteractionRange) */

enum css_reduce_op { CSS_SUM = 0, CSS_MAX = 1 };

/* counters returned by css_counters (uint64 each) */
enum css_counter {
    CSS_C_WALK_VERTEX = 0,   /* displacement passed through a vertex (two edges hit) */
    CSS_C_WALK_NOHIT = 1,    /* target outside the face but no edge intersection found */
    CSS_C_WALK_ITERCAP = 2,  /* more than CSS_WALK_MAX_CROSSINGS edge crossings */
    CSS_C_WALK_NAN = 3,
    CSS_C_WALK_BORDER = 4,   /* border edge met in a closed-mesh space */
    CSS_C_DISCONNECTED = 5,  /* target unreachable inside the patch -> sentinel (2*maxDist, (0,0,1)) */
    CSS_C_TIES = 6,
    CSS_C_CROSSINGS = 7,     /* edge crossings walked */
    CSS_C_WINDOWS = 8,       /* windows propagated */
    CSS_C_PSEUDO = 9,        /* pseudo-source fans emitted */
    CSS_C_PATCH_FACES = 10,  /* sum of patch face counts */
    CSS_C_PATCH_VERTS = 11,
    CSS_C_QUERIES = 12,      /* (source,target) geodesic queries answered */
    CSS_C_SOURCES = 13,      /* sources processed */
    CSS_C_TIER_RETRY = 14,   /* sources re-run on a larger-capacity tier */
    CSS_C_OVERFLOW = 15,     /* sources that overflowed every tier (results invalid) */
    CSS_C_KERNELS = 16,      /* kernels launched by this context (host-side count) */
    CSS_C_KMAX_OVERFLOW = 17, /* neighbour stride too small (handled by regrowing and rerunning) */
    CSS_C_OVF_CANDIDATES = 18, CSS_C_OVF_FACES = 19, CSS_C_OVF_VERTS = 20, CSS_C_OVF_RING = 21, /* tier overflow reasons */
    CSS_C_CLK_BATCH = 22, CSS_C_CLK_FAN = 23, CSS_C_CLK_PROP = 24, CSS_C_CLK_PATCH = 25, CSS_C_CLK_TOTAL = 26, /* developer statistics */
    CSS_C_PEER_TIMEOUT = 27, /* a peer-exchange flag wait gave up */
    CSS_C_SPILLED = 28,      /* windows that left a full shared-memory ring for the global spill stack (and came back) */
    CSS_NUM_COUNTERS = 32
};

int css_create(css_ctx** ctx, int device);
int css_destroy(css_ctx* ctx);
const char* css_last_error(css_ctx* ctx);

int css_set_mesh(css_ctx* ctx, int nV, const double* xyz, int nF, const int32_t* corners);
int css_mesh_info(css_ctx* ctx, double bbmin[3], double bbmax[3], double* area);
int css_set_submeshing(css_ctx* ctx, int enabled, double maxDist);
int css_set_cell_domain(css_ctx* ctx, const double mn[3], const double mx[3]);
/* open-mesh boundary rule of the walker: 0 closed space (a border edge is flagged, triangulatedMeshSpace.cpp:547-548),
 * 1 absorbingOpenMeshSpace (src/models/absorbingOpenMeshSpace.cpp:2-50), 2 tangentialOpenMeshSpace (tangentialOpenMeshSpace.cpp:3-64) */
int css_set_boundary(css_ctx* ctx, int mode);
/* useCellList=0 reproduces baseNeighborStructure (all-to-all candidates) */
int css_set_options(css_ctx* ctx, int useCellList, int wantEndTangents);

/* ---- per-call API parity with baseSpace ---- */
int css_euclidean(css_ctx* ctx, int n, const int32_t* face, const double* bary, double* xyz);
/* closest mesh position of n points of R^3: face index and clamped barycentric weights (clampTol = simpleModel::clampTolerance,
 * 1e-14 in the reference); among faces at exactly the same distance the lowest index wins */
int css_locate(css_ctx* ctx, int n, const double* xyz, double clampTol, int32_t* face, double* bary);
int css_distance(css_ctx* ctx, int srcFace, const double srcBary[3], int K, const int32_t* tgtFace, const double* tgtBary,
                 double threshold, double* dist, double* startTan, double* endTan);
/* vecs is [n][nVec][3] (may be NULL when nVec==0); flags may be NULL */
int css_transport(css_ctx* ctx, int n, int32_t* face, double* bary, double* disp, int nVec, double* vecs, int32_t* flags);

/* ---- batched hot path (model state lives on the device) ---- */
/* face/bary: nTotal replicated positions; vel/frc: this rank's nLocal particles [minIdx, minIdx+nLocal) (may be NULL -> zero) */
int css_set_state(css_ctx* ctx, int nLocal, int nTotal, int minIdx, const int32_t* face, const double* bary, const double* vel,
                  const double* frc);
int css_get_state(css_ctx* ctx, int32_t* face /*nTotal*/, double* bary /*nTotal*/, double* vel /*nLocal*/, double* frc /*nLocal*/);
int css_set_velocities(css_ctx* ctx, const double* vel);
int css_set_forces(css_ctx* ctx, const double* frc);
int css_find_neighbors(css_ctx* ctx, double range, int64_t* totalNeighbors);
/* offsets has nLocal+1 entries; any output pointer may be NULL */
int css_get_neighbors(css_ctx* ctx, int32_t* offsets, int32_t* idx, double* dist, double* startTan, double* endTan);
int css_compute_forces(css_ctx* ctx, int kind, const double* params, int zero);
int css_compute_energy(css_ctx* ctx, int kind, const double* params, double* energy);
/* stress[3 a + b]: density kB <v_a v_b> / 2 + <f_a dr_b> / (2 d A), d = 2, density = N / A; v (x) v is accumulated once per
 * neighbour, as the reference's loop nest does.  Runs find_neighbors at the force's range first (like the reference). */
int css_compute_stress(css_ctx* ctx, int kind, const double* params, double stress[9]);
int css_temperature(css_ctx* ctx, double* temperature); /* sum v.v / (2 N) over all ranks */
int css_move(css_ctx* ctx, const double* disp /*nLocal, host; NULL = use the device displacement buffer*/, int transportForce,
             int transportVelocity);
int css_get_walk_flags(css_ctx* ctx, int32_t* flags /*nLocal*/);

/* ---- fused updaters ---- */
int css_step_nve(css_ctx* ctx, int kind, const double* params, double dt, int nsteps);
/* The same step for a HOST-resident state (simpleModel's public std::vectors): uploads face/bary [nTotal] and vel/frc
 * [nLocal] (page-locked buffers recommended), steps once, downloads the new state into the same buffers; the positions
 * come back on a second stream while the neighbour / force phase is still running.  Sharding as set by css_set_state. */
int css_step_nve_host(css_ctx* ctx, int kind, const double* params, double dt, int32_t* face, double* bary, double* vel, double* frc);
int css_step_gd(css_ctx* ctx, int kind, const double* params, double dt, int nsteps);
int css_nvt_init(css_ctx* ctx, double dt, double T, double tau, int M);
int css_step_nvt(css_ctx* ctx, int kind, const double* params, int nsteps);
int css_nvt_state(css_ctx* ctx, double* bath /*4*(M+1)*/, double* kineticEnergy, double* scale);
/* p = {maximumIterations, deltaT(ignored, as in the reference), alphaStart, deltaTMax, deltaTMin, deltaTInc, deltaTDec,
 *      alphaDec, nMin, forceCutoff, alphaMin}; NULL keeps the reference defaults */
int css_fire_init(css_ctx* ctx, const double* p, double dt0, double alpha0);
int css_fire_minimize(css_ctx* ctx, int kind, const double* params, double* out /*iterations, forceMax, dt, alpha*/);
int css_max_force(css_ctx* ctx, double* maxForce);
int css_force_norm(css_ctx* ctx, double* forceNorm);

/* ---- multi-GPU (one process per GPU) ---- */
int css_comm_unique_id(void* id128 /*128 bytes out*/);
int css_comm_init(css_ctx* ctx, int rank, int nranks, const void* id128);
int css_gather_positions(css_ctx* ctx); /* all-gather (face, bary) of every rank's block into the replicated arrays */
/* peerExchange = 1 when the exchange after every move runs over peer memory (the walker stores new positions straight into
 * every rank's CUDA-IPC-mapped window over NVLink, followed by a flag barrier) instead of the NCCL all-gather; decided
 * collectively at the first move after css_comm_init (CSS_P2P=0 in the environment keeps NCCL) */
int css_comm_info(css_ctx* ctx, int* rank, int* nranks, int* peerExchange);
int css_reduce(css_ctx* ctx, int op, int k, double* data); /* gather per-rank partials, fold in rank order */

/* ---- measurement / diagnostics ---- */
int css_counters(css_ctx* ctx, uint64_t* out /*CSS_NUM_COUNTERS*/, int reset);
int css_synchronize(css_ctx* ctx);
/* raw device pointers of the replicated position arrays (for callers that run their own collectives) */
int css_device_positions(css_ctx* ctx, void** face_dev, void** bary_dev);
/* timing of the last css_find_neighbors/compute_forces geodesic kernel, measured with CUDA events on the ctx stream */
int css_last_kernel_ms(css_ctx* ctx, float* geodesic_ms, float* walk_ms, float* celllist_ms);
int css_set_timing(css_ctx* ctx, int enabled);
/* split of the last geodesic phase: stage 1 (patch records), stage 2 (window propagation + forces), retry tiers;
 * gather_ms = the position all-gather between the walker and the neighbour phase of the last fused step */
int css_last_stage_ms(css_ctx* ctx, float* patch_ms, float* window_ms, float* retry_ms, float* gather_ms);
/* CUDA-event stopwatch on the context's own stream (slots 0..7) */
int css_timer_record(css_ctx* ctx, int slot);
int css_timer_elapsed_ms(css_ctx* ctx, int slotA, int slotB, float* ms); /* synchronises on slotB */
/* measured ceilings of this device for the roofline statements (replaces nothing in the reference; its profiler.h:16-58 is a
 * wall clock).  what = 0: double-precision FMA rate in TFLOP/s (8 independent FMA chains per thread, whole chip);
 * what = 1: L2 -> SM read bandwidth in GB/s over a 32 MiB buffer.  Best of `reps` CUDA-event-timed launches. */
int css_microbench(css_ctx* ctx, int what, int reps, double* value);

#ifdef __cplusplus
}
#endif
#endif /* CSS_API_H */
