// Shared device-side definitions for the geodesic MD hot path (sm_100a, fp64).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace css {

// ---- mesh in HBM (L2-resident at 1 M faces: 16 MB vertices + 16 MB corners + 16 MB adjacency) ----
// vertices : double4 {x,y,z,s}   one 32-byte sector per gather; s = 1.0 when the vertex is a saddle (angle sum >= 2 pi), else 0
// corners  : int4 {c0,c1,c2,0}   reference corner order
// adjacency: int4 {a0,a1,a2,kk}  a_k = face across the edge opposite corner k (-1 border);
//                                kk packs the index of that edge inside the neighbour, 2 bits per k
struct MeshDev {
    int nV, nF;
    const double4* vert;
    const int4* corner;
    const int4* adj;
    const unsigned char* saddle; // per-vertex: interior angle sum >= 2 pi  (pseudo-source candidate)
    // edge frames: for face f and edge e (opposite corner e, running A = corner e+1 -> B = corner e+2) the apex
    // C = corner e in the frame of that edge, normalised by |AB|:  C = A + cxn (B-A) + cyn perp(B-A).
    // geo[3 f + e] = {cxn, cyn}; lets a window be unfolded across a face with four FMAs and no 3-D geometry.
    const double2* geo;
    int boundary; // walker at a border edge: 0 closed space (flag + stop), 1 absorbing, 2 tangential (openMeshSpace variants)
    // flood-fill table of stage 1, two int4 per face (one 32-byte sector): {a0,a1,a2,kk} as in `adj`, then {o0,o1,o2,0} with
    // o_k = the vertex of neighbour a_k that lies opposite the shared edge (-1 at a border)
    const int4* adjopp;
};

struct CellGrid {
    double mn[3], cs[3];
    int n[3];
    double range2;
};

struct ForceParams {
    int kind;
    double a, sigma; // harmonic: k, sigma ; gaussian: alpha, sigma
};

struct d3 {
    double x, y, z;
};

// ---- exactly-rounded arithmetic (never contracted into FMA) --------------------------------------
// Used wherever a branch or an index decision must be bit-identical to the CPU oracle: cell binning,
// candidate selection, patch membership, the whole walker.
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double xsqrt(double a) { return __dsqrt_rn(a); }

__device__ __forceinline__ d3 xsub3(const d3& a, const d3& b) { return d3{xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)}; }
__device__ __forceinline__ d3 xadd3(const d3& a, const d3& b) { return d3{xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z)}; }
__device__ __forceinline__ d3 xscale(double s, const d3& a) { return d3{xmul(s, a.x), xmul(s, a.y), xmul(s, a.z)}; }
__device__ __forceinline__ d3 xdivs(const d3& a, double s) { return d3{xdiv(a.x, s), xdiv(a.y, s), xdiv(a.z, s)}; }
__device__ __forceinline__ double xdot(const d3& a, const d3& b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
__device__ __forceinline__ d3 xcross(const d3& a, const d3& b)
{
    return d3{xsub(xmul(a.y, b.z), xmul(a.z, b.y)), xsub(xmul(a.z, b.x), xmul(a.x, b.z)), xsub(xmul(a.x, b.y), xmul(a.y, b.x))};
}
__device__ __forceinline__ double xsqlen(const d3& a) { return xdot(a, a); }
__device__ __forceinline__ double xnorm(const d3& a) { return xsqrt(xsqlen(a)); }

__device__ __forceinline__ d3 ldvert(const MeshDev& m, int v)
{
    const double2* p = reinterpret_cast<const double2*>(m.vert + v);
    double2 a = __ldg(p);
    double2 b = __ldg(p + 1);
    return d3{a.x, a.y, b.x};
}

// (b0 p0 + b1 p1 + b2 p2) / (b0 + b1 + b2)   [PMP::construct_point, SURVEY.md §8(c)-C3]
__device__ __forceinline__ d3 xpoint(const d3& p0, const d3& p1, const d3& p2, double b0, double b1, double b2)
{
    double s = xadd(xadd(b0, b1), b2);
    return d3{xdiv(xadd(xadd(xmul(b0, p0.x), xmul(b1, p1.x)), xmul(b2, p2.x)), s),
              xdiv(xadd(xadd(xmul(b0, p0.y), xmul(b1, p1.y)), xmul(b2, p2.y)), s),
              xdiv(xadd(xadd(xmul(b0, p0.z), xmul(b1, p1.z)), xmul(b2, p2.z)), s)};
}

__device__ __forceinline__ int cellCoord(const CellGrid& g, double x, int d)
{
    int c = (int)floor(xdiv(xsub(x, g.mn[d]), g.cs[d]));
    return max(0, min(g.n[d] - 1, c));
}

// ---- programmatic dependent launch (PDL).  The kernels of a step are tiny next to their launch latency at the tail of the
// sequence (cell list, empty retry tiers) and the big ones end with a ragged tail.  Every step kernel is launched with the
// programmatic-stream-serialization attribute and starts with PDL_ENTRY(): it tells the scheduler that its successor may be
// made resident as soon as SM resources free up (launch_dependents) and then waits until its predecessor has completed and
// flushed (wait) -- memory ordering is exactly that of plain stream order, only launch latency and block scheduling overlap
// the predecessor's tail.  Without the attribute both instructions are no-ops.
#define PDL_ENTRY()                                             \
    do {                                                        \
        asm volatile("griddepcontrol.launch_dependents;");      \
        asm volatile("griddepcontrol.wait;" ::: "memory");      \
    } while (0)

#ifdef __CUDACC__
bool pdlEnabled(); // CSS_PDL=0 switches the attribute off (css_api.cu)
template <class... KArgs, class... Args>
inline cudaError_t launchStep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = pdlEnabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// walker flag bits / counters (mirrors include/css_api.h)
enum { WALK_VERTEX = 1, WALK_NOHIT = 2, WALK_ITERCAP = 4, WALK_NAN = 8, WALK_BORDER = 16 };
enum {
    C_WALK_VERTEX = 0, C_WALK_NOHIT, C_WALK_ITERCAP, C_WALK_NAN, C_WALK_BORDER, C_DISCONNECTED, C_TIES, C_CROSSINGS, C_WINDOWS,
    C_PSEUDO, C_PATCH_FACES, C_PATCH_VERTS, C_QUERIES, C_SOURCES, C_TIER_RETRY, C_OVERFLOW, C_KERNELS, C_KMAX_OVERFLOW,
    C_OVF_REASON /* 18..21: candidates, faces, vertices, ring */,
    C_CLK_BATCH = 22, C_CLK_FAN = 23, C_CLK_PROP = 24, C_CLK_PATCH = 25, C_CLK_TOTAL = 26, /* summed per-warp clock64 cycles */
    C_PEER_TIMEOUT = 27, /* a peer-exchange flag wait gave up (a rank died or left the collective sequence) */
    C_SPILLED = 28,      /* windows pushed to a global spill stack by k_windows_half */
    C_KMAX_NEED = 29,    /* neighbour stride that would have been enough (max 27-cell occupancy - 1), set with C_KMAX_OVERFLOW */
    C_STEP_GUARD = 30,   /* walker launches that really moved particles since the host last cleared it (see strideGuard) */
    NUM_COUNTERS = 32
};
#define CSS_WALK_MAX_CROSSINGS 100000

// Neighbour-stride guard.  Before stage 1 the cell-list build bounds every particle's candidate count by the occupancy of its
// 27-cell stencil; when that exceeds the neighbour stride it raises counters[C_KMAX_OVERFLOW].  While the flag is up every
// kernel of the step pipeline (patch, window, retry tiers, walker) returns at once, so a fused multi-step call (CUDA-graph
// replays included) freezes in a recoverable state: positions and velocities after the move of the failed step, forces not yet
// recomputed.  The host regrows the stride, finishes that step and runs the remaining ones.  The bound depends only on the
// replicated positions, so every rank of a sharded run raises the flag in the same step.
__device__ __forceinline__ bool strideGuardUp(const unsigned long long* counters) { return *(volatile const unsigned long long*)(counters + C_KMAX_OVERFLOW) != 0ull; }

} // namespace css
