// C ABI of the B200-native geodesic MD hot path (see include/css_api.h for the reference interfaces
// each entry point replaces).  Host side: owns device memory, one stream, the tiered launch of the
// geodesic kernel and (optionally) one NCCL communicator.  No CPU fallback exists: every entry point
// that computes launches CUDA kernels and fails with CSS_ECUDA when no device is usable.
#include "../../include/css_api.h"
#include "kernels.h"
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <nccl.h>
#include <string>
#include <unordered_map>
#include <vector>

using namespace css;

#define CSS_GATHER_PAD 64 /* spare position entries behind nTotal (>= nranks - 1, see css_gather_positions) */

struct css_ctx {
    int device = 0;
    cudaStream_t st = nullptr;
    std::string err;
    // mesh
    int nV = 0, nF = 0;
    double4* d_vert = nullptr;
    int4* d_corner = nullptr;
    int4* d_adj = nullptr;
    int4* d_adjopp = nullptr; // [2 nF] flood-fill table of stage 1 (common.cuh MeshDev::adjopp)
    // static face stencils of stage 1 (stencil_kernel.cu): built at the first neighbour phase after the mesh / the submeshing
    // cut-off changed; used while most faces have one
    unsigned char* d_stencil = nullptr;
    unsigned* d_stencilLen = nullptr;
    size_t stencilCap = 0;
    bool useStencil = true, stencilValid = false, stencilUsable = false;
    double stencilMeanFaces = 0;
    long long stencilMissing = 0;
    unsigned char* d_saddle = nullptr;
    double2* d_geo = nullptr; // edge frames, [3 nF]
    double bbmin[3] = {0, 0, 0}, bbmax[3] = {0, 0, 0}, area = 0;
    double cellMin[3] = {0, 0, 0}, cellMax[3] = {0, 0, 0};
    bool submeshing = false;
    double maxDist = 0;
    bool useCellList = true, wantEnd = false;
    int boundaryMode = 0; // 0 closed, 1 absorbing, 2 tangential
    // particles
    int nLocal = 0, nTotal = 0, minIdx = 0, capLocal = 0, capTotal = 0;
    int* d_face = nullptr;
    double *d_bary = nullptr, *d_eucl = nullptr, *d_vel = nullptr, *d_frc = nullptr, *d_disp = nullptr;
    int* d_walkFlags = nullptr;
    // cell list
    CellGrid grid{};
    double gridRange = -1;
    int nCells = 0, capCells = 0;
    int *d_cellOf = nullptr, *d_cellCount = nullptr, *d_cellStart = nullptr, *d_cellSlot = nullptr, *d_tmpItems = nullptr,
        *d_items = nullptr;
    // neighbours (fixed stride kmax)
    int kmax = 32, capNbr = 0;
    bool nbrHasEnd = false;
    int* d_nbrCount = nullptr;
    int* d_nbrIdx = nullptr;
    double *d_nbrDist = nullptr, *d_nbrTs = nullptr, *d_nbrTe = nullptr;
    bool nbrValid = false;
    // geodesic tiers
    int *d_work = nullptr, *d_retry[4] = {nullptr, nullptr, nullptr, nullptr}; // [3]: sources the stencil stage hands to the flood fill
    int capRetry = 0;
    char* d_gws = nullptr;
    size_t gwsBytes = 0;
    GeoCaps capsT2{0, 0, 0, 0, 0, 0}; // whole-mesh capacities of the last tier (set with the mesh)
    // one warp per SM on a shared-memory workspace (~195 kB): long-range patches of up to 1408 faces / 768 vertices (the reference's
    // default executable: N = 20, range 2.6 on torus_isotropic_remesh.off) stay out of the global-memory tier
    GeoCaps capsHuge{1408, 768, 1024, 128, 4096, 2048}; // long-range tier: one CTA per source, workspace in shared memory
    int t2Warps = 32;
    // two-stage tier 0: patch records (stage 1) -> window propagation (stage 2)
    bool winLean = true, winHalf = true;
    bool nvtOnDevice = true; // single-rank Nose-Hoover steps keep the chain on the device (CSS_NVT_HOST=1: host chain, one sync per step)
    int winWpb = 2;
    unsigned char *d_records = nullptr, *d_recordsL = nullptr;
    double* d_spill = nullptr; // window spill stacks of the two-sources-per-warp kernel (allocated once)
    // face grid of css_locate (built at the first call after css_set_mesh)
    int *d_fgStart = nullptr, *d_fgFaces = nullptr;
    double fgMin[3] = {0, 0, 0}, fgH = 0;
    int fgN[3] = {0, 0, 0};
    bool fgValid = false;
    size_t capRecords = 0, capRecordsL = 0;
    int numSMs = 148;
    // reductions / scratch
    double *d_partial = nullptr, *d_red = nullptr;
    unsigned long long* d_counters = nullptr;
    unsigned long long hostKernels = 0;
    // updaters
    struct {
        double dt, dt2, dt4, dt8, T, tau;
        int M = 0;
        std::vector<double> bx, by, bz, bw;
        double KE = 0, scale = 1;
        // single rank: the chain lives on the device (k_nh_chain), so css_step_nvt queues all its steps without a host round trip
        double* d = nullptr;      // bx | by | bz | bw | KE, scale, T, dt2, dt4, dt8
        int dM = 0;               // chain length the device block was allocated for
        bool deviceNewer = false; // the device copy is ahead of the host mirror (css_nvt_state downloads it)
        bool hostNewer = true;    // the host mirror has to be uploaded before the next device step
    } nh;
    struct FireState {
        double dt = 0.001, alpha = 0.99;
        int maximumIterations = 1000, nMin = 4, nSinceNegativePower = 0, iterations = 0;
        double alphaStart = 0.99, deltaTMax = 0.1, deltaTInc = 1.1, deltaTMin = 1e-5, deltaTDec = 0.95, alphaDec = 0.9, forceCutoff = 1e-12,
               alphaMin = 0.0;
        double forceMax = 0;
    } fire;
    // comm
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    int* d_sendI = nullptr;
    double* d_sendD = nullptr;
    int* d_recvI = nullptr;
    double* d_recvD = nullptr;
    int capComm = 0;
    double* d_redBuf = nullptr;
    // peer-memory exchange window (kernels.h PeerWin): cudaMalloc + CUDA IPC, mapped by every rank of the node
    bool p2pEnabled = true; // CSS_P2P=0 keeps the NCCL all-gather
    bool p2pFailed = false; // IPC mapping is not possible here: NCCL from now on (decided collectively)
    void* winLocal = nullptr;
    void* winPeer[CSS_MAX_PEERS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int winCap = 0; // particles
    PeerWin pw{};
    unsigned long long* d_epoch = nullptr;
    unsigned* d_ticket = nullptr;
    unsigned char* d_ipcBuf = nullptr;
    // CUDA graph of one fused NVE step (walker, gather, cell list, stages 1-2, retry tiers): one launch per step
    bool useGraph = true, capturing = false;
    cudaGraphExec_t nveExec = nullptr;
    uint64_t nveKey = 0;
    unsigned long long nveKernels = 0;
    int nveCalls = 0;
    // host-buffer step (css_step_nve_host): positions are downloaded on a second stream while the neighbour / force phase runs
    cudaStream_t stCopy = nullptr;
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    int32_t* ioFace = nullptr;
    double* ioBary = nullptr;
    bool ioNoGraph = false;
    // timing
    bool timing = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float msGeo = 0, msWalk = 0, msCell = 0;
    cudaEvent_t evS[4] = {nullptr, nullptr, nullptr, nullptr}; // stage boundaries inside the geodesic phase
    cudaEvent_t tev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// event record that stays a real, queryable event when the stream is being captured into a CUDA graph
static void recordEvent(css_ctx* c, cudaEvent_t e);
static void releasePeerWindow(css_ctx* c);

namespace css {
bool pdlEnabled()
{
    static const bool on = [] {
        const char* v = getenv("CSS_PDL");
        return !v || atoi(v) != 0;
    }();
    return on;
}
} // namespace css

static int fail(css_ctx* c, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}
#define CU(call)                                                                                                         \
    do {                                                                                                                 \
        cudaError_t e_ = (call);                                                                                         \
        if (e_ != cudaSuccess) return fail(ctx, CSS_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define NC(call)                                                                                                         \
    do {                                                                                                                 \
        ncclResult_t e_ = (call);                                                                                        \
        if (e_ != ncclSuccess) return fail(ctx, CSS_ENCCL, "%s failed: %s", #call, ncclGetErrorString(e_));              \
    } while (0)
#define BIND() CU(cudaSetDevice(ctx->device))

template <class T> static cudaError_t regrow(T*& p, size_t n)
{
    if (p) cudaFree(p);
    p = nullptr;
    return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
}

static void recordEvent(css_ctx* c, cudaEvent_t e)
{
    if (c->capturing) cudaEventRecordWithFlags(e, c->st, cudaEventRecordExternal);
    else cudaEventRecord(e, c->st);
}

// ------------------------------------------------------------------------------- peer exchange window
static void releasePeerWindow(css_ctx* c)
{
    for (int r = 0; r < CSS_MAX_PEERS; ++r) {
        if (c->winPeer[r] && c->winPeer[r] != c->winLocal) cudaIpcCloseMemHandle(c->winPeer[r]);
        c->winPeer[r] = nullptr;
    }
    if (c->winLocal) cudaFree(c->winLocal);
    c->winLocal = nullptr;
    c->winCap = 0;
    c->pw = PeerWin{};
}
static size_t peerFaceBytes(int cap) { return ((size_t)cap * sizeof(int) + 255) / 256 * 256; }

// Collective over the communicator (every rank calls it with the same nTotal, from the same place of the same call
// sequence): (re)creates the exchange windows for at least nTotal particles and maps every peer's window.  On any
// failure all ranks agree to stay on the NCCL all-gather.
static int ensurePeerWindow(css_ctx* ctx)
{
    if (ctx->nranks <= 1 || !ctx->p2pEnabled || ctx->p2pFailed || ctx->nranks > CSS_MAX_PEERS) return CSS_OK;
    // window capacity derived from nTotal only, so that every rank computes the same layout whatever its allocation history
    const int cap = (ctx->nTotal + 1023) / 1024 * 1024;
    if (ctx->pw.n > 1 && ctx->winCap == cap) return CSS_OK;
    if (ctx->capturing) return fail(ctx, CSS_ESTATE, "peer window must exist before a step is captured");
    const int R = ctx->nranks;
    if (!ctx->d_ipcBuf) {
        CU(regrow(ctx->d_ipcBuf, 128 * (size_t)(CSS_MAX_PEERS + 1)));
        CU(regrow(ctx->d_epoch, 1));
        CU(regrow(ctx->d_ticket, 1));
        CU(cudaMemsetAsync(ctx->d_ticket, 0, sizeof(unsigned), ctx->st));
    }
    struct Msg {
        cudaIpcMemHandle_t h;
        int ok, pad[15];
    };
    static_assert(sizeof(Msg) == 128, "IPC message size");
    auto exchange = [&](const Msg& mine, std::vector<Msg>& all) -> int {
        CU(cudaMemcpyAsync(ctx->d_ipcBuf, &mine, sizeof(Msg), cudaMemcpyHostToDevice, ctx->st));
        NC(ncclAllGather(ctx->d_ipcBuf, ctx->d_ipcBuf + 128, sizeof(Msg), ncclUint8, ctx->comm, ctx->st));
        all.resize(R);
        CU(cudaMemcpyAsync(all.data(), ctx->d_ipcBuf + 128, sizeof(Msg) * R, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        return CSS_OK;
    };
    std::vector<Msg> all;
    Msg mine{};
    // 1. nobody is still using the old windows (all earlier exchanges are stream-ordered before this all-gather)
    int rc = exchange(mine, all);
    if (rc) return rc;
    releasePeerWindow(ctx);
    // 2. new local window, flags and epoch zeroed before anybody can signal
    const size_t bytes = CSS_PEER_FLAG_BYTES + peerFaceBytes(cap) + sizeof(double) * 3 * (size_t)cap;
    mine.ok = cudaMalloc(&ctx->winLocal, bytes) == cudaSuccess;
    if (mine.ok) {
        mine.ok = cudaMemsetAsync(ctx->winLocal, 0, bytes, ctx->st) == cudaSuccess && cudaMemsetAsync(ctx->d_epoch, 0, 8, ctx->st) == cudaSuccess &&
                  cudaStreamSynchronize(ctx->st) == cudaSuccess && cudaIpcGetMemHandle(&mine.h, ctx->winLocal) == cudaSuccess;
    }
    (void)cudaGetLastError();
    rc = exchange(mine, all);
    if (rc) return rc;
    bool okAll = true;
    for (int r = 0; r < R; ++r) okAll = okAll && all[r].ok;
    // 3. map the peers
    Msg st2{};
    st2.ok = okAll;
    if (okAll) {
        for (int r = 0; r < R && st2.ok; ++r) {
            if (r == ctx->rank) ctx->winPeer[r] = ctx->winLocal;
            else if (cudaIpcOpenMemHandle(&ctx->winPeer[r], all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                ctx->winPeer[r] = nullptr;
                st2.ok = 0;
                (void)cudaGetLastError();
            }
        }
    }
    rc = exchange(st2, all);
    if (rc) return rc;
    for (int r = 0; r < R; ++r) okAll = okAll && all[r].ok;
    if (!okAll) { // every rank takes the same decision
        releasePeerWindow(ctx);
        ctx->p2pFailed = true;
        return CSS_OK;
    }
    ctx->winCap = cap;
    ctx->pw.n = R, ctx->pw.rank = ctx->rank;
    for (int r = 0; r < R; ++r) {
        unsigned char* base = (unsigned char*)ctx->winPeer[r];
        ctx->pw.flags[r] = (unsigned long long*)base;
        ctx->pw.face[r] = (int*)(base + CSS_PEER_FLAG_BYTES);
        ctx->pw.bary[r] = (double*)(base + CSS_PEER_FLAG_BYTES + peerFaceBytes(cap));
    }
    ctx->nveCalls = 0;
    return CSS_OK;
}

#pragma GCC visibility push(default)
extern "C" {

int css_create(css_ctx** out, int device)
{
    if (!out) return CSS_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return CSS_ECUDA;
    css_ctx* ctx = new css_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return CSS_ECUDA;
    }
    cudaDeviceGetAttribute(&ctx->numSMs, cudaDevAttrMultiProcessorCount, device);
    cudaMalloc(&ctx->d_counters, NUM_COUNTERS * sizeof(unsigned long long));
    cudaMemset(ctx->d_counters, 0, NUM_COUNTERS * sizeof(unsigned long long));
    cudaMalloc(&ctx->d_work, 16 * sizeof(int));
    cudaMalloc(&ctx->d_partial, REDUCE_MAX_BLOCKS * 18 * sizeof(double));
    cudaMalloc(&ctx->d_red, 24 * sizeof(double));
    for (auto& e : ctx->ev) cudaEventCreate(&e);
    for (auto& e : ctx->evS) cudaEventCreate(&e);
    for (auto& e : ctx->tev) cudaEventCreate(&e);
    if (const char* v = getenv("CSS_NO_GRAPH")) ctx->useGraph = atoi(v) == 0;
    if (const char* v = getenv("CSS_WIN_LEAN")) ctx->winLean = atoi(v) != 0;
    if (const char* v = getenv("CSS_WIN_HALF")) ctx->winHalf = atoi(v) != 0; // 0: one warp per source in tier 0 (window_kernel.cu)
    if (const char* v = getenv("CSS_NVT_HOST")) ctx->nvtOnDevice = atoi(v) == 0;
    if (const char* v = getenv("CSS_STENCIL")) ctx->useStencil = atoi(v) != 0; // 0: stage 1 of tier 0 flood-fills every patch (patch_kernel.cu)
    if (const char* v = getenv("CSS_WIN_WPB")) ctx->winWpb = std::max(1, std::min(4, atoi(v)));
    if (const char* v = getenv("CSS_P2P")) ctx->p2pEnabled = atoi(v) != 0;
    *out = ctx;
    return CSS_OK;
}

int css_destroy(css_ctx* ctx)
{
    if (!ctx) return CSS_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    if (ctx->nveExec) cudaGraphExecDestroy(ctx->nveExec);
    releasePeerWindow(ctx);
    if (ctx->comm) ncclCommDestroy(ctx->comm);
    void* ptrs[] = {ctx->d_vert,   ctx->d_corner,    ctx->d_adj,     ctx->d_saddle,  ctx->d_face,     ctx->d_bary,    ctx->d_eucl,
                    ctx->d_vel,    ctx->d_frc,       ctx->d_disp,    ctx->d_walkFlags, ctx->d_cellOf, ctx->d_cellCount, ctx->d_cellStart,
                    ctx->d_cellSlot, ctx->d_tmpItems, ctx->d_items,  ctx->d_nbrCount, ctx->d_nbrIdx,  ctx->d_nbrDist,
                    ctx->d_nbrTs,  ctx->d_nbrTe,     ctx->d_work,    ctx->d_retry[0], ctx->d_retry[1], ctx->d_retry[2], ctx->d_retry[3], ctx->d_gws,
                    ctx->d_partial, ctx->d_red,      ctx->d_counters, ctx->d_sendI,  ctx->d_sendD,    ctx->d_recvI,   ctx->d_recvD,
                    ctx->d_redBuf,  ctx->d_geo,       ctx->d_records, ctx->d_recordsL, ctx->d_epoch, ctx->d_ticket, ctx->d_ipcBuf, ctx->d_spill, ctx->d_fgStart, ctx->d_fgFaces, ctx->d_adjopp, ctx->d_stencil, ctx->d_stencilLen};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (auto& e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->tev)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->evS)
        if (e) cudaEventDestroy(e);
    if (ctx->evFork) cudaEventDestroy(ctx->evFork);
    if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
    if (ctx->stCopy) cudaStreamDestroy(ctx->stCopy);
    if (ctx->nh.d) cudaFree(ctx->nh.d);
    cudaStreamDestroy(ctx->st);
    delete ctx;
    return CSS_OK;
}

const char* css_last_error(css_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// ------------------------------------------------------------------------------------------ mesh
int css_set_mesh(css_ctx* ctx, int nV, const double* xyz, int nF, const int32_t* corners)
{
    if (!ctx || nV <= 0 || nF <= 0 || !xyz || !corners) return fail(ctx, CSS_EINVAL, "css_set_mesh: bad arguments");
    BIND();
    for (int i = 0; i < 3 * nF; ++i)
        if (corners[i] < 0 || corners[i] >= nV) return fail(ctx, CSS_EMESH, "Invalid input file."); // triangulatedMeshSpace.cpp:55
    // directed-edge map -> adjacency (what CGAL::Surface_mesh connectivity provides to the reference)
    std::unordered_map<uint64_t, int> half;
    half.reserve((size_t)nF * 6);
    auto key = [](int a, int b) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; };
    for (int f = 0; f < nF; ++f)
        for (int k = 0; k < 3; ++k) {
            int a = corners[3 * f + (k + 1) % 3], b = corners[3 * f + (k + 2) % 3];
            if (a == b || !half.emplace(key(a, b), 3 * f + k).second)
                return fail(ctx, CSS_EMESH, "mesh is not a consistently oriented manifold triangle mesh (face %d)", f);
        }
    std::vector<int4> hc(nF), ha(nF);
    for (int f = 0; f < nF; ++f) {
        int a3[3], kk = 0;
        for (int k = 0; k < 3; ++k) {
            int a = corners[3 * f + (k + 1) % 3], b = corners[3 * f + (k + 2) % 3];
            auto it = half.find(key(b, a));
            if (it == half.end()) a3[k] = -1;
            else {
                a3[k] = it->second / 3;
                kk |= (it->second % 3) << (2 * k);
            }
        }
        hc[f] = make_int4(corners[3 * f], corners[3 * f + 1], corners[3 * f + 2], 0);
        ha[f] = make_int4(a3[0], a3[1], a3[2], kk);
    }
    std::vector<int4> hao(2 * (size_t)nF);
    for (int f = 0; f < nF; ++f) {
        int o3[3];
        for (int k = 0; k < 3; ++k) {
            const int g = k == 0 ? ha[f].x : (k == 1 ? ha[f].y : ha[f].z);
            o3[k] = g < 0 ? -1 : corners[3 * g + ((ha[f].w >> (2 * k)) & 3)]; // corner kk of g is opposite its edge kk
        }
        hao[2 * (size_t)f] = ha[f];
        hao[2 * (size_t)f + 1] = make_int4(o3[0], o3[1], o3[2], 0);
    }
    // bounding box seeded with the origin (triangulatedMeshSpace::updateMeshSpanAndTree :9-30), area, angle sums
    std::vector<double4> hv(nV);
    for (int d = 0; d < 3; ++d) ctx->bbmin[d] = ctx->bbmax[d] = 0;
    for (int i = 0; i < nV; ++i) {
        hv[i] = make_double4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0);
        for (int d = 0; d < 3; ++d) {
            ctx->bbmin[d] = std::min(ctx->bbmin[d], xyz[3 * i + d]);
            ctx->bbmax[d] = std::max(ctx->bbmax[d], xyz[3 * i + d]);
        }
    }
    std::vector<double> ang(nV, 0.0);
    double area = 0;
    for (int f = 0; f < nF; ++f) {
        for (int k = 0; k < 3; ++k) {
            const double* p = xyz + 3 * corners[3 * f + k];
            const double* q = xyz + 3 * corners[3 * f + (k + 1) % 3];
            const double* r = xyz + 3 * corners[3 * f + (k + 2) % 3];
            double a[3] = {q[0] - p[0], q[1] - p[1], q[2] - p[2]}, b[3] = {r[0] - p[0], r[1] - p[1], r[2] - p[2]};
            double cr[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
            double cn = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
            ang[corners[3 * f + k]] += std::atan2(cn, a[0] * b[0] + a[1] * b[1] + a[2] * b[2]);
            if (k == 0) area += cn / 2.0;
        }
    }
    ctx->area = area;
    // edge frames (common.cuh MeshDev::geo): apex of each face in the frame of each of its edges
    std::vector<double2> hg(3 * (size_t)nF);
    for (int f = 0; f < nF; ++f)
        for (int e = 0; e < 3; ++e) {
            const double* A = xyz + 3 * corners[3 * f + (e + 1) % 3];
            const double* B = xyz + 3 * corners[3 * f + (e + 2) % 3];
            const double* C = xyz + 3 * corners[3 * f + e];
            double ab[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, ac[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
            double cr[3] = {ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]};
            double l2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
            hg[3 * (size_t)f + e] = make_double2((ac[0] * ab[0] + ac[1] * ab[1] + ac[2] * ab[2]) / l2,
                                                 std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]) / l2);
        }
    std::vector<unsigned char> sad(nV);
    for (int i = 0; i < nV; ++i) sad[i] = ang[i] >= 2.0 * M_PI - 1e-9, hv[i].w = sad[i] ? 1.0 : 0.0;
    for (int d = 0; d < 3; ++d) ctx->cellMin[d] = ctx->bbmin[d], ctx->cellMax[d] = ctx->bbmax[d];
    ctx->gridRange = -1;
    CU(regrow(ctx->d_vert, nV));
    CU(regrow(ctx->d_corner, nF));
    CU(regrow(ctx->d_adj, nF));
    CU(regrow(ctx->d_adjopp, 2 * (size_t)nF));
    CU(cudaMemcpy(ctx->d_adjopp, hao.data(), sizeof(int4) * 2 * (size_t)nF, cudaMemcpyHostToDevice));
    CU(regrow(ctx->d_saddle, nV));
    CU(regrow(ctx->d_geo, 3 * (size_t)nF));
    CU(cudaMemcpy(ctx->d_geo, hg.data(), sizeof(double2) * 3 * (size_t)nF, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_vert, hv.data(), sizeof(double4) * nV, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_corner, hc.data(), sizeof(int4) * nF, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_adj, ha.data(), sizeof(int4) * nF, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_saddle, sad.data(), nV, cudaMemcpyHostToDevice));
    ctx->nV = nV;
    ctx->nF = nF;
    ctx->stencilValid = false;
    ctx->fgValid = false;
    ctx->nbrValid = false;
    ctx->nveCalls = 0;
    // last-resort tier: the whole mesh fits (local ids are 16 bit)
    int mf = std::min(nF, 65534), mv = std::min(nV, 65534);
    auto p2 = [](int x) {
        int p = 1;
        while (p < x) p <<= 1;
        return p;
    };
    ctx->capsT2 = GeoCaps{mf, mv, p2(std::min(4 * mf + 64, 1 << 20)), 256, p2(2 * mf + 128), p2(2 * mv + 256)};
    return CSS_OK;
}

int css_mesh_info(css_ctx* ctx, double bbmin[3], double bbmax[3], double* area)
{
    if (!ctx || !ctx->nF) return fail(ctx, CSS_ESTATE, "mesh not set");
    for (int d = 0; d < 3; ++d) {
        if (bbmin) bbmin[d] = ctx->bbmin[d];
        if (bbmax) bbmax[d] = ctx->bbmax[d];
    }
    if (area) *area = ctx->area;
    return CSS_OK;
}

int css_set_submeshing(css_ctx* ctx, int enabled, double maxDist)
{
    if (!ctx) return CSS_EINVAL;
    ctx->nveCalls = 0;
    if ((enabled != 0) != ctx->submeshing || maxDist != ctx->maxDist) ctx->stencilValid = false;
    ctx->submeshing = enabled != 0;
    ctx->maxDist = maxDist;
    ctx->nbrValid = false;
    return CSS_OK;
}
int css_set_cell_domain(css_ctx* ctx, const double mn[3], const double mx[3])
{
    if (!ctx || !mn || !mx) return CSS_EINVAL;
    for (int d = 0; d < 3; ++d) ctx->cellMin[d] = mn[d], ctx->cellMax[d] = mx[d];
    ctx->gridRange = -1;
    ctx->nveCalls = 0;
    return CSS_OK;
}
int css_set_boundary(css_ctx* ctx, int mode)
{
    if (!ctx || mode < 0 || mode > 2) return fail(ctx, CSS_EINVAL, "css_set_boundary: mode must be 0 (closed), 1 (absorbing) or 2 (tangential)");
    ctx->boundaryMode = mode;
    ctx->nveCalls = 0;
    return CSS_OK;
}
int css_set_options(css_ctx* ctx, int useCellList, int wantEndTangents)
{
    if (!ctx) return CSS_EINVAL;
    ctx->nveCalls = 0;
    ctx->useCellList = useCellList != 0;
    ctx->wantEnd = wantEndTangents != 0;
    ctx->nbrValid = false;
    return CSS_OK;
}

static MeshDev meshDev(css_ctx* c) { return MeshDev{c->nV, c->nF, c->d_vert, c->d_corner, c->d_adj, c->d_saddle, c->d_geo, c->boundaryMode, c->d_adjopp}; }

// ------------------------------------------------------------------------------- per-call parity
int css_euclidean(css_ctx* ctx, int n, const int32_t* face, const double* bary, double* xyz)
{
    if (!ctx || !ctx->nF) return fail(ctx, CSS_ESTATE, "mesh not set");
    if (n <= 0) return CSS_OK;
    BIND();
    for (int i = 0; i < n; ++i)
        if (face[i] < 0 || face[i] >= ctx->nF) return fail(ctx, CSS_EINVAL, "css_euclidean: face index %d out of range", face[i]);
    int* df = nullptr;
    double *db = nullptr, *dx = nullptr;
    CU(cudaMalloc(&df, sizeof(int) * n));
    CU(cudaMalloc(&db, sizeof(double) * 3 * n));
    CU(cudaMalloc(&dx, sizeof(double) * 3 * n));
    CU(cudaMemcpyAsync(df, face, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(db, bary, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->st));
    launchEuclidCell(ctx->st, meshDev(ctx), ctx->grid, n, df, db, dx, nullptr, nullptr, nullptr);
    ctx->hostKernels++;
    CU(cudaMemcpyAsync(xyz, dx, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    cudaFree(df), cudaFree(db), cudaFree(dx);
    return CSS_OK;
}

// simpleModel::R3PositionsToMeshPositions (src/models/simpleModel.cpp:136-154).  The reference builds a CGAL AABB tree per call;
// here a uniform grid over the faces' bounding boxes is built once per mesh on the host (cell edge = twice the mean mesh edge,
// at most 2^24 cells) and searched on the device shell by shell (exact_kernels.cu k_locate).
static int buildFaceGrid(css_ctx* ctx)
{
    const int nV = ctx->nV, nF = ctx->nF;
    std::vector<double4> hv(nV);
    std::vector<int4> hc(nF);
    CU(cudaMemcpy(hv.data(), ctx->d_vert, sizeof(double4) * nV, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hc.data(), ctx->d_corner, sizeof(int4) * nF, cudaMemcpyDeviceToHost));
    double mn[3] = {hv[0].x, hv[0].y, hv[0].z}, mx[3] = {hv[0].x, hv[0].y, hv[0].z};
    for (int i = 1; i < nV; ++i) {
        const double q[3] = {hv[i].x, hv[i].y, hv[i].z};
        for (int d = 0; d < 3; ++d) mn[d] = std::min(mn[d], q[d]), mx[d] = std::max(mx[d], q[d]);
    }
    double esum = 0;
    for (int f = 0; f < nF; ++f) {
        const int c[3] = {hc[f].x, hc[f].y, hc[f].z};
        for (int k = 0; k < 3; ++k) {
            const double4 &a = hv[c[k]], &b = hv[c[(k + 1) % 3]];
            esum += std::sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z));
        }
    }
    double ext = std::max(mx[0] - mn[0], std::max(mx[1] - mn[1], mx[2] - mn[2]));
    double h = 2.0 * esum / (3.0 * nF);
    if (!(h > 0)) h = ext > 0 ? ext : 1.0;
    int n[3];
    for (;;) {
        double cells = 1;
        for (int d = 0; d < 3; ++d) n[d] = (int)std::floor((mx[d] - mn[d]) / h) + 1, cells *= n[d];
        if (cells <= (double)(1 << 24)) break;
        h *= 1.26; // halves the cell count
    }
    const size_t ncell = (size_t)n[0] * n[1] * n[2];
    auto cellRange = [&](int f, int lo[3], int hi[3]) {
        const int c[3] = {hc[f].x, hc[f].y, hc[f].z};
        for (int d = 0; d < 3; ++d) {
            double a = 1e300, b = -1e300;
            for (int k = 0; k < 3; ++k) {
                const double q = d == 0 ? hv[c[k]].x : (d == 1 ? hv[c[k]].y : hv[c[k]].z);
                a = std::min(a, q), b = std::max(b, q);
            }
            // one cell of slack on both sides: the device bins the query with its own rounding of the same division
            lo[d] = std::max(0, (int)std::floor((a - mn[d]) / h) - 1), hi[d] = std::min(n[d] - 1, (int)std::floor((b - mn[d]) / h) + 1);
        }
    };
    std::vector<int> start(ncell + 1, 0);
    int lo[3], hi[3];
    for (int f = 0; f < nF; ++f) {
        cellRange(f, lo, hi);
        for (int z = lo[2]; z <= hi[2]; ++z)
            for (int y = lo[1]; y <= hi[1]; ++y)
                for (int x = lo[0]; x <= hi[0]; ++x) start[((size_t)z * n[1] + y) * n[0] + x + 1]++;
    }
    for (size_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
    std::vector<int> faces((size_t)start[ncell]), fill(start.begin(), start.end() - 1);
    for (int f = 0; f < nF; ++f) {
        cellRange(f, lo, hi);
        for (int z = lo[2]; z <= hi[2]; ++z)
            for (int y = lo[1]; y <= hi[1]; ++y)
                for (int x = lo[0]; x <= hi[0]; ++x) faces[fill[((size_t)z * n[1] + y) * n[0] + x]++] = f;
    }
    CU(regrow(ctx->d_fgStart, ncell + 1));
    CU(regrow(ctx->d_fgFaces, faces.size()));
    CU(cudaMemcpy(ctx->d_fgStart, start.data(), sizeof(int) * (ncell + 1), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_fgFaces, faces.data(), sizeof(int) * faces.size(), cudaMemcpyHostToDevice));
    for (int d = 0; d < 3; ++d) ctx->fgMin[d] = mn[d], ctx->fgN[d] = n[d];
    ctx->fgH = h;
    ctx->fgValid = true;
    return CSS_OK;
}

int css_locate(css_ctx* ctx, int n, const double* xyz, double clampTol, int32_t* face, double* bary)
{
    if (!ctx || !ctx->nF) return fail(ctx, CSS_ESTATE, "mesh not set");
    if (n < 0 || (n > 0 && (!xyz || !face || !bary))) return fail(ctx, CSS_EINVAL, "css_locate: bad arguments");
    if (n == 0) return CSS_OK;
    BIND();
    if (!ctx->fgValid)
        if (int rc = buildFaceGrid(ctx)) return rc;
    int* df = nullptr;
    double *dx = nullptr, *db = nullptr;
    CU(cudaMalloc(&df, sizeof(int) * n));
    CU(cudaMalloc(&dx, sizeof(double) * 3 * n));
    CU(cudaMalloc(&db, sizeof(double) * 3 * n));
    CU(cudaMemcpyAsync(dx, xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->st));
    launchLocate(ctx->st, meshDev(ctx), ctx->fgMin, ctx->fgH, ctx->fgN, ctx->d_fgStart, ctx->d_fgFaces, n, dx, clampTol, df, db);
    ctx->hostKernels++;
    CU(cudaMemcpyAsync(face, df, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(bary, db, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    cudaFree(df), cudaFree(dx), cudaFree(db);
    for (int i = 0; i < n; ++i)
        if (face[i] < 0) return fail(ctx, CSS_EINVAL, "css_locate: point %d has no closest face (coordinates are not finite)", i);
    return CSS_OK;
}

int css_transport(css_ctx* ctx, int n, int32_t* face, double* bary, double* disp, int nVec, double* vecs, int32_t* flags)
{
    if (!ctx || !ctx->nF) return fail(ctx, CSS_ESTATE, "mesh not set");
    if (n <= 0) return CSS_OK;
    BIND();
    for (int i = 0; i < n; ++i)
        if (face[i] < 0 || face[i] >= ctx->nF) return fail(ctx, CSS_EINVAL, "css_transport: face index %d out of range", face[i]);
    int *df = nullptr, *dfl = nullptr;
    double *db = nullptr, *dd = nullptr, *dv = nullptr;
    CU(cudaMalloc(&df, sizeof(int) * n));
    CU(cudaMalloc(&dfl, sizeof(int) * n));
    CU(cudaMalloc(&db, sizeof(double) * 3 * n));
    CU(cudaMalloc(&dd, sizeof(double) * 3 * n));
    CU(cudaMalloc(&dv, sizeof(double) * 3 * (size_t)n * std::max(nVec, 1)));
    CU(cudaMemcpyAsync(df, face, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(db, bary, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(dd, disp, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->st));
    if (nVec > 0) CU(cudaMemcpyAsync(dv, vecs, sizeof(double) * 3 * (size_t)n * nVec, cudaMemcpyHostToDevice, ctx->st));
    launchTransportGeneric(ctx->st, meshDev(ctx), n, df, db, dd, nVec, dv, dfl);
    ctx->hostKernels++;
    CU(cudaMemcpyAsync(face, df, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(bary, db, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(disp, dd, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->st));
    if (nVec > 0) CU(cudaMemcpyAsync(vecs, dv, sizeof(double) * 3 * (size_t)n * nVec, cudaMemcpyDeviceToHost, ctx->st));
    if (flags) CU(cudaMemcpyAsync(flags, dfl, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    cudaFree(df), cudaFree(dfl), cudaFree(db), cudaFree(dd), cudaFree(dv);
    return CSS_OK;
}

// ------------------------------------------------------------------------------------- geodesics
// (re)builds the static face stencils for the current mesh and submeshing cut-off; never called while a step is being captured
static int ensureStencils(css_ctx* ctx)
{
    if (ctx->stencilValid || !ctx->useStencil || !ctx->submeshing) return CSS_OK;
    if (ctx->capturing) return fail(ctx, CSS_ESTATE, "face stencils must exist before a step is captured");
    ctx->stencilUsable = false;
    const size_t bytes = stencilBytes(ctx->nF);
    if (bytes > (32ull << 30)) { // meshes beyond ~17 M faces keep the flood fill
        ctx->stencilValid = true;
        return CSS_OK;
    }
    if (bytes > ctx->stencilCap) {
        CU(cudaStreamSynchronize(ctx->st));
        if (ctx->d_stencil) cudaFree(ctx->d_stencil);
        ctx->d_stencil = nullptr, ctx->stencilCap = 0;
        if (ctx->d_stencilLen) cudaFree(ctx->d_stencilLen);
        ctx->d_stencilLen = nullptr;
        if (cudaMalloc(&ctx->d_stencil, bytes) != cudaSuccess || cudaMalloc(&ctx->d_stencilLen, sizeof(unsigned) * (size_t)ctx->nF) != cudaSuccess) { // no room: the flood fill serves every source
            if (ctx->d_stencil) cudaFree(ctx->d_stencil);
            ctx->d_stencil = nullptr;
            (void)cudaGetLastError();
            ctx->stencilValid = true;
            return CSS_OK;
        }
        ctx->stencilCap = bytes;
    }
    unsigned long long* dstats = reinterpret_cast<unsigned long long*>(ctx->d_red); // 24 doubles of scratch
    CU(buildStencils(ctx->st, meshDev(ctx), ctx->maxDist, ctx->d_stencil, ctx->d_stencilLen, dstats, ctx->numSMs));
    ctx->hostKernels++;
    unsigned long long h[2] = {0, 0};
    CU(cudaMemcpyAsync(h, dstats, sizeof h, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    ctx->stencilMissing = (long long)h[0];
    ctx->stencilMeanFaces = ctx->nF > (int)h[0] ? (double)h[1] / (ctx->nF - (double)h[0]) : 0.0;
    ctx->stencilUsable = 2 * h[0] <= (unsigned long long)ctx->nF; // with most faces uncovered (coarse meshes) the flood fill is the better first stage
    ctx->stencilValid = true;
    if (getenv("CSS_VERBOSE"))
        fprintf(stderr, "[css] face stencils: %d faces, %.1f MB, mean %.1f stencil faces, %lld faces without one -> %s\n", ctx->nF, bytes / 1e6,
                ctx->stencilMeanFaces, ctx->stencilMissing, ctx->stencilUsable ? "stencil stage 1" : "flood-fill stage 1");
    return CSS_OK;
}

static int ensureTierBuffers(css_ctx* ctx, int nSrc)
{
    if (nSrc > ctx->capRetry) {
        for (int t = 0; t < 4; ++t) CU(regrow(ctx->d_retry[t], nSrc));
        ctx->capRetry = nSrc;
    }
    return CSS_OK;
}

// runs the tiered geodesic kernel for a prepared GeoArgs (srcList/work fields are filled here)
static int runGeodesicTiers(css_ctx* ctx, GeoArgs a, int nSrc)
{
    int rc = ensureTierBuffers(ctx, std::max(nSrc, 1));
    if (rc) return rc;
    CU(cudaMemsetAsync(ctx->d_work, 0, 16 * sizeof(int), ctx->st)); // [0..2] work counters, [4..6] retry counts
    auto ensureWorkspace = [&](size_t bytes) -> int { // global-memory workspace of the fused retry kernel
        if (bytes > ctx->gwsBytes) {
            CU(cudaStreamSynchronize(ctx->st));
            if (ctx->d_gws) cudaFree(ctx->d_gws);
            ctx->d_gws = nullptr;
            ctx->gwsBytes = 0;
            CU(cudaMalloc(&ctx->d_gws, bytes));
            ctx->gwsBytes = bytes;
        }
        return CSS_OK;
    };
    const bool staged = a.xK < 0 && a.cellStart != nullptr; // neighbour phase with a cell list: record tiers first
    if (staged && (rc = ensureStencils(ctx))) return rc;
    if (staged) {
        // Two-stage tiers: patch records (integer / latency-bound, high occupancy), then window propagation (fp64).
        // Tier 0 = TierSmall over every local source; tier 1 = TierLarge over the sources tier 0 handed on.
        static const int maxLargeEnv = getenv("CSS_MAX_LARGE") ? atoi(getenv("CSS_MAX_LARGE")) : -1; // developer switch: 0 hands tier 1's sources straight on
        const int maxLarge = maxLargeEnv >= 0 ? maxLargeEnv : std::min(std::max(nSrc, 1), 65536);
        size_t needS = (size_t)std::max(nSrc, 1) * TierSmall::BYTES, needL = (size_t)maxLarge * TierLarge::BYTES;
        if (needS > ctx->capRecords || needL > ctx->capRecordsL || (ctx->winHalf && !ctx->d_spill)) {
            CU(cudaStreamSynchronize(ctx->st));
            if (ctx->winHalf && !ctx->d_spill) CU(regrow(ctx->d_spill, windowsHalfSpillBytes(ctx->numSMs) / sizeof(double)));
            if (needS > ctx->capRecords) {
                CU(regrow(ctx->d_records, needS));
                ctx->capRecords = needS;
            }
            if (needL > ctx->capRecordsL) {
                CU(regrow(ctx->d_recordsL, needL));
                ctx->capRecordsL = needL;
            }
        }
        PatchArgs p{};
        p.m = a.m, p.grid = a.grid, p.nLocal = a.nLocal, p.minIdx = a.minIdx;
        p.face = a.face, p.eucl = a.eucl, p.cellStart = a.cellStart, p.cellCount = a.cellCount, p.cellItems = a.cellItems;
        p.cellOf = ctx->d_cellOf, p.invNx = 1.0 / a.grid.n[0], p.invNxy = 1.0 / ((double)a.grid.n[0] * a.grid.n[1]);
        p.submeshing = a.submeshing, p.maxDist = a.maxDist, p.kmax = a.kmax;
        p.counters = ctx->d_counters;
        WinArgs w{};
        w.m = a.m, w.nLocal = a.nLocal, w.minIdx = a.minIdx;
        w.face = a.face, w.bary = a.bary, w.eucl = a.eucl;
        w.submeshing = a.submeshing, w.maxDist = a.maxDist, w.kmax = a.kmax;
        w.nbrCount = a.nbrCount, w.nbrIdx = a.nbrIdx, w.nbrDist = a.nbrDist, w.nbrTs = a.nbrTs, w.nbrTe = a.nbrTe;
        w.forceMode = a.forceMode, w.fp = a.fp, w.zero = a.zero, w.frc = a.frc, w.kick = a.kick, w.vel = a.vel;
        w.counters = ctx->d_counters, w.spill = ctx->d_spill;
        // tier 0
        p.srcList = nullptr, p.srcCount = nullptr, p.maxRecords = std::max(nSrc, 1);
        p.workCounter = ctx->d_work + 0, p.retryList = ctx->d_retry[0], p.retryCount = ctx->d_work + 4, p.records = ctx->d_records;
        w.srcList = nullptr, w.srcCount = nullptr, w.maxRecords = p.maxRecords;
        w.workCounter = ctx->d_work + 3, w.retryList = ctx->d_retry[0], w.retryCount = ctx->d_work + 4, w.records = ctx->d_records;
        if (ctx->timing) recordEvent(ctx, ctx->evS[0]);
        p.stencil = ctx->d_stencil, p.stencilLen = ctx->d_stencilLen;
        if (ctx->useStencil && ctx->stencilUsable && a.submeshing) {
            // sources whose face has no stencil are flood-filled into the same records by a second launch; when every face has
            // one, the launch is skipped and the (very rare) source with a target outside its stencil goes to the large tier
            const bool fallback = ctx->stencilMissing > 0;
            p.fallbackList = fallback ? ctx->d_retry[3] : nullptr, p.fallbackCount = ctx->d_work + 9;
            CU(launchPatchStencil<TierSmall>(ctx->st, p, ctx->numSMs));
            if (fallback) {
                PatchArgs q = p;
                q.srcList = ctx->d_retry[3], q.srcCount = ctx->d_work + 9, q.workCounter = ctx->d_work + 10, q.recordByParticle = 1;
                CU(launchPatch<TierSmall>(ctx->st, q, ctx->numSMs));
                ctx->hostKernels++;
            }
        } else
            CU(launchPatch<TierSmall>(ctx->st, p, ctx->numSMs));
        if (ctx->timing) recordEvent(ctx, ctx->evS[1]);
        if (ctx->winHalf) CU(launchWindowsHalf<TierHalf>(ctx->st, w, ctx->numSMs));
        else CU(launchWindows<TierSmall>(ctx->st, w, ctx->winWpb, ctx->numSMs, ctx->winLean));
        if (ctx->timing) recordEvent(ctx, ctx->evS[2]);
        // tier 1 (work list = tier 0's retry list; its length is only known on the device)
        p.srcList = ctx->d_retry[0], p.srcCount = ctx->d_work + 4, p.maxRecords = maxLarge;
        p.workCounter = ctx->d_work + 1, p.retryList = ctx->d_retry[1], p.retryCount = ctx->d_work + 5, p.records = ctx->d_recordsL;
        w.srcList = p.srcList, w.srcCount = p.srcCount, w.maxRecords = maxLarge;
        w.workCounter = ctx->d_work + 7, w.retryList = ctx->d_retry[1], w.retryCount = ctx->d_work + 5, w.records = ctx->d_recordsL;
        CU(launchPatch<TierLarge>(ctx->st, p, ctx->numSMs));
        CU(launchWindows<TierLarge>(ctx->st, w, 1, ctx->numSMs, false));
        ctx->hostKernels += 4;
    }
    // Long-range tier: the fused kernel, ONE CTA PER SOURCE AND SM with its whole workspace in shared memory (384 windows per pass;
    // the global-memory tier below pays an L2 round trip for every ring / vertex access).  It serves what the record tiers handed
    // on and, directly, explicit css_distance queries and the all-to-all (no cell list) mode.
    {
        a.caps = ctx->capsHuge;
        a.srcList = staged ? ctx->d_retry[1] : nullptr, a.srcCount = staged ? ctx->d_work + 5 : nullptr;
        a.workCounter = ctx->d_work + 11;
        a.retryList = ctx->d_retry[2], a.retryCount = ctx->d_work + 6;
        a.gws = nullptr, a.lastTier = 0;
        CU(launchGeodesicCta(ctx->st, a, std::min(ctx->numSMs, staged ? ctx->numSMs : std::max(nSrc, 1))));
        ctx->hostKernels++;
    }
    // last tier: the same kernel with capacities sized for the whole mesh, global-memory workspace (one per block)
    {
        a.caps = ctx->capsT2;
        if (a.xK >= 0) a.caps.kt = std::max(a.caps.kt, a.xK);
        else if (!ctx->useCellList) a.caps.kt = std::max(a.caps.kt, a.nTotal);
        else a.caps.kt = std::max(a.caps.kt, a.kmax);
        size_t ws = geoWorkspaceBytes(a.caps);
        int blocks = ctx->t2Warps;
        if (int rc2 = ensureWorkspace(ws * blocks)) return rc2;
        a.srcList = ctx->d_retry[2], a.srcCount = ctx->d_work + 6;
        a.workCounter = ctx->d_work + 2;
        a.retryList = ctx->d_retry[1], a.retryCount = ctx->d_work + 12; // (the last tier retries nothing)
        a.gws = ctx->d_gws, a.lastTier = 1;
        CU(launchGeodesicCta(ctx->st, a, blocks));
        ctx->hostKernels++;
    }
    return CSS_OK;
}

int css_distance(css_ctx* ctx, int srcFace, const double srcBary[3], int K, const int32_t* tgtFace, const double* tgtBary,
                 double threshold, double* dist, double* startTan, double* endTan)
{
    if (!ctx || !ctx->nF) return fail(ctx, CSS_ESTATE, "mesh not set");
    if (K < 0 || srcFace < 0 || srcFace >= ctx->nF) return fail(ctx, CSS_EINVAL, "css_distance: bad source");
    if (K == 0) return CSS_OK;
    BIND();
    for (int i = 0; i < K; ++i)
        if (tgtFace[i] < 0 || tgtFace[i] >= ctx->nF) return fail(ctx, CSS_EINVAL, "css_distance: target face out of range");
    int* df = nullptr;
    double *db = nullptr, *dd = nullptr, *dts = nullptr, *dte = nullptr;
    CU(cudaMalloc(&df, sizeof(int) * K));
    CU(cudaMalloc(&db, sizeof(double) * 3 * K));
    CU(cudaMalloc(&dd, sizeof(double) * K));
    CU(cudaMalloc(&dts, sizeof(double) * 3 * K));
    CU(cudaMalloc(&dte, sizeof(double) * 3 * K));
    CU(cudaMemcpyAsync(df, tgtFace, sizeof(int) * K, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(db, tgtBary, sizeof(double) * 3 * K, cudaMemcpyHostToDevice, ctx->st));
    GeoArgs a{};
    a.m = meshDev(ctx);
    a.nTotal = 0, a.nLocal = 1, a.minIdx = 0;
    a.submeshing = ctx->submeshing, a.maxDist = ctx->maxDist;
    a.xK = K, a.xSrcFace = srcFace;
    a.xSrcBary[0] = srcBary[0], a.xSrcBary[1] = srcBary[1], a.xSrcBary[2] = srcBary[2];
    a.xThreshold = threshold;
    a.xTgtFace = df, a.xTgtBary = db;
    a.kmax = K;
    a.nbrDist = dd, a.nbrTs = dts, a.nbrTe = dte;
    a.counters = ctx->d_counters;
    unsigned long long before = 0, after = 0;
    CU(cudaMemcpyAsync(&before, ctx->d_counters + C_OVERFLOW, sizeof before, cudaMemcpyDeviceToHost, ctx->st));
    int rc = runGeodesicTiers(ctx, a, 1);
    if (rc) return rc;
    CU(cudaMemcpyAsync(&after, ctx->d_counters + C_OVERFLOW, sizeof after, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(dist, dd, sizeof(double) * K, cudaMemcpyDeviceToHost, ctx->st));
    if (startTan) CU(cudaMemcpyAsync(startTan, dts, sizeof(double) * 3 * K, cudaMemcpyDeviceToHost, ctx->st));
    if (endTan) CU(cudaMemcpyAsync(endTan, dte, sizeof(double) * 3 * K, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    cudaFree(df), cudaFree(db), cudaFree(dd), cudaFree(dts), cudaFree(dte);
    if (after != before) return fail(ctx, CSS_ECAPACITY, "css_distance: patch/window capacity exceeded on every tier");
    return CSS_OK;
}

// ------------------------------------------------------------------------------------- model state
static int ensureParticles(css_ctx* ctx, int nLocal, int nTotal)
{
    if (nTotal > ctx->capTotal) {
        // CSS_GATHER_PAD spare entries: in the padded all-gather of css_gather_positions the last rank's block is shorter than
        // ceil(N/R), and every rank sends a full block starting at its own offset
        CU(regrow(ctx->d_face, (size_t)nTotal + CSS_GATHER_PAD));
        CU(regrow(ctx->d_bary, 3 * ((size_t)nTotal + CSS_GATHER_PAD)));
        CU(regrow(ctx->d_eucl, 3 * (size_t)nTotal));
        CU(regrow(ctx->d_cellOf, nTotal));
        CU(regrow(ctx->d_cellSlot, nTotal));
        CU(regrow(ctx->d_tmpItems, nTotal));
        CU(regrow(ctx->d_items, nTotal));
        ctx->capTotal = nTotal;
    }
    if (nLocal > ctx->capLocal) {
        CU(regrow(ctx->d_vel, 3 * (size_t)nLocal));
        CU(regrow(ctx->d_frc, 3 * (size_t)nLocal));
        CU(regrow(ctx->d_disp, 3 * (size_t)nLocal));
        CU(regrow(ctx->d_walkFlags, nLocal));
        ctx->capLocal = nLocal;
        ctx->capNbr = 0;
    }
    return CSS_OK;
}
static int ensureNeighbors(css_ctx* ctx)
{
    size_t need = (size_t)std::max(ctx->nLocal, 1) * ctx->kmax;
    if (need > (size_t)ctx->capNbr || (ctx->wantEnd && !ctx->nbrHasEnd)) {
        CU(regrow(ctx->d_nbrCount, std::max(ctx->nLocal, 1)));
        CU(regrow(ctx->d_nbrIdx, need));
        CU(regrow(ctx->d_nbrDist, need));
        CU(regrow(ctx->d_nbrTs, 3 * need));
        if (ctx->wantEnd) CU(regrow(ctx->d_nbrTe, 3 * need));
        // the rows have a fixed stride and are downloaded whole (css_get_neighbors compacts them by the counts): define the unused tails once
        CU(cudaMemsetAsync(ctx->d_nbrIdx, 0, sizeof(int) * need, ctx->st));
        CU(cudaMemsetAsync(ctx->d_nbrDist, 0, sizeof(double) * need, ctx->st));
        CU(cudaMemsetAsync(ctx->d_nbrTs, 0, sizeof(double) * 3 * need, ctx->st));
        if (ctx->wantEnd) CU(cudaMemsetAsync(ctx->d_nbrTe, 0, sizeof(double) * 3 * need, ctx->st));
        ctx->nbrHasEnd = ctx->wantEnd;
        ctx->capNbr = (int)need;
    }
    return CSS_OK;
}

int css_set_state(css_ctx* ctx, int nLocal, int nTotal, int minIdx, const int32_t* face, const double* bary, const double* vel,
                  const double* frc)
{
    if (!ctx || !ctx->nF) return fail(ctx, CSS_ESTATE, "mesh not set");
    if (nLocal < 0 || nTotal < nLocal || minIdx < 0 || minIdx + nLocal > nTotal || !face || !bary)
        return fail(ctx, CSS_EINVAL, "css_set_state: bad sharding arguments");
    BIND();
    for (int i = 0; i < nTotal; ++i)
        if (face[i] < 0 || face[i] >= ctx->nF) return fail(ctx, CSS_EINVAL, "css_set_state: face index %d out of range", face[i]);
    int rc = ensureParticles(ctx, nLocal, nTotal);
    if (rc) return rc;
    if (nLocal != ctx->nLocal || nTotal != ctx->nTotal || minIdx != ctx->minIdx) ctx->nveCalls = 0; // buffers get re-sized by plain steps first
    ctx->nLocal = nLocal, ctx->nTotal = nTotal, ctx->minIdx = minIdx;
    CU(cudaMemcpyAsync(ctx->d_face, face, sizeof(int) * nTotal, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(ctx->d_bary, bary, sizeof(double) * 3 * nTotal, cudaMemcpyHostToDevice, ctx->st));
    if (vel) CU(cudaMemcpyAsync(ctx->d_vel, vel, sizeof(double) * 3 * nLocal, cudaMemcpyHostToDevice, ctx->st));
    else CU(cudaMemsetAsync(ctx->d_vel, 0, sizeof(double) * 3 * std::max(nLocal, 1), ctx->st));
    if (frc) CU(cudaMemcpyAsync(ctx->d_frc, frc, sizeof(double) * 3 * nLocal, cudaMemcpyHostToDevice, ctx->st));
    else CU(cudaMemsetAsync(ctx->d_frc, 0, sizeof(double) * 3 * std::max(nLocal, 1), ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    ctx->nbrValid = false;
    return CSS_OK;
}
int css_get_state(css_ctx* ctx, int32_t* face, double* bary, double* vel, double* frc)
{
    if (!ctx || !ctx->nTotal) return fail(ctx, CSS_ESTATE, "state not set");
    BIND();
    if (face) CU(cudaMemcpyAsync(face, ctx->d_face, sizeof(int) * ctx->nTotal, cudaMemcpyDeviceToHost, ctx->st));
    if (bary) CU(cudaMemcpyAsync(bary, ctx->d_bary, sizeof(double) * 3 * ctx->nTotal, cudaMemcpyDeviceToHost, ctx->st));
    if (vel) CU(cudaMemcpyAsync(vel, ctx->d_vel, sizeof(double) * 3 * ctx->nLocal, cudaMemcpyDeviceToHost, ctx->st));
    if (frc) CU(cudaMemcpyAsync(frc, ctx->d_frc, sizeof(double) * 3 * ctx->nLocal, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return CSS_OK;
}
int css_set_velocities(css_ctx* ctx, const double* vel)
{
    if (!ctx || !ctx->nTotal || !vel) return fail(ctx, CSS_ESTATE, "state not set");
    BIND();
    CU(cudaMemcpyAsync(ctx->d_vel, vel, sizeof(double) * 3 * ctx->nLocal, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return CSS_OK;
}
int css_set_forces(css_ctx* ctx, const double* frc)
{
    if (!ctx || !ctx->nTotal || !frc) return fail(ctx, CSS_ESTATE, "state not set");
    BIND();
    CU(cudaMemcpyAsync(ctx->d_frc, frc, sizeof(double) * 3 * ctx->nLocal, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return CSS_OK;
}

// hyperRectangularCellList::setGridSize (hyperRectangularCellList.cpp:9-46)
static int setupGrid(css_ctx* ctx, double range)
{
    if (range != ctx->gridRange) {
        long total = 1;
        for (int d = 0; d < 3; ++d) {
            double ext = ctx->cellMax[d] - ctx->cellMin[d];
            int n = std::max(1, (int)std::floor(ext / range));
            ctx->grid.n[d] = n;
            ctx->grid.cs[d] = ext / n;
            ctx->grid.mn[d] = ctx->cellMin[d];
            total *= n;
        }
        if (total > (1L << 28)) return fail(ctx, CSS_ECAPACITY, "cell grid of %ld cells is too large", total);
        ctx->nCells = (int)total;
        ctx->gridRange = range;
    }
    ctx->grid.range2 = range * range;
    if (ctx->nCells > ctx->capCells) {
        // counts | bump counter | occupancy of the 2x2x2 coarse blocks (sharded runs: replicated stride bound), cleared together
        CU(regrow(ctx->d_cellCount, 2 * (size_t)ctx->nCells + 2)); // (a coarse grid never has more blocks than the grid has cells)
        CU(regrow(ctx->d_cellStart, (size_t)ctx->nCells + 1));
        // only occupied cells get a start written each step; readers fetch start and count together and ignore the start of an
        // empty cell, so the array is defined once here (compute-sanitizer initcheck)
        CU(cudaMemsetAsync(ctx->d_cellStart, 0, sizeof(int) * ((size_t)ctx->nCells + 1), ctx->st));
        ctx->capCells = ctx->nCells;
    }
    return CSS_OK;
}

// findNeighbors (+ optional fused force / kick). forceMode 0: lists only.
static int findNeighborsImpl(css_ctx* ctx, double range, int forceMode, ForceParams fp, int zero, double kick)
{
    if (!ctx->nTotal) return fail(ctx, CSS_ESTATE, "state not set");
    int rc;
    MeshDev m = meshDev(ctx);
    if (ctx->timing) recordEvent(ctx, ctx->ev[0]);
    if (ctx->useCellList) {
        if ((rc = setupGrid(ctx, range))) return rc;
        const size_t nCoarse = (size_t)((ctx->grid.n[0] + 1) / 2) * ((ctx->grid.n[1] + 1) / 2) * ((ctx->grid.n[2] + 1) / 2);
        int* coarse = ctx->nranks > 1 ? ctx->d_cellCount + ctx->nCells + 1 : nullptr;
        CU(cudaMemsetAsync(ctx->d_cellCount, 0, sizeof(int) * ((size_t)ctx->nCells + 1 + (coarse ? nCoarse : 0)), ctx->st));
        launchEuclidCell(ctx->st, m, ctx->grid, ctx->nTotal, ctx->d_face, ctx->d_bary, ctx->d_eucl, ctx->d_cellOf, ctx->d_cellCount, ctx->d_cellSlot,
                         coarse);
        launchCellBuild(ctx->st, ctx->nTotal, ctx->nCells, ctx->d_cellOf, ctx->d_cellSlot, ctx->d_cellCount, ctx->d_cellStart, ctx->d_tmpItems,
                        ctx->d_items, coarse, ctx->grid, ctx->kmax, ctx->d_counters);
        ctx->hostKernels += 4;
    } else {
        launchEuclidCell(ctx->st, m, ctx->grid, ctx->nTotal, ctx->d_face, ctx->d_bary, ctx->d_eucl, nullptr, nullptr, nullptr);
        ctx->hostKernels += 1;
        if (ctx->nTotal - 1 > ctx->kmax) {
            ctx->kmax = ctx->nTotal - 1;
            ctx->capNbr = 0;
        }
    }
    if (ctx->timing) recordEvent(ctx, ctx->ev[1]);
    if ((rc = ensureNeighbors(ctx))) return rc;
    GeoArgs a{};
    a.m = m;
    a.grid = ctx->grid;
    a.nTotal = ctx->nTotal, a.nLocal = ctx->nLocal, a.minIdx = ctx->minIdx;
    a.face = ctx->d_face, a.bary = ctx->d_bary, a.eucl = ctx->d_eucl;
    a.cellStart = ctx->useCellList ? ctx->d_cellStart : nullptr;
    a.cellCount = ctx->d_cellCount;
    a.cellItems = ctx->d_items;
    a.submeshing = ctx->submeshing, a.maxDist = ctx->maxDist;
    a.xK = -1;
    a.kmax = ctx->kmax;
    a.nbrCount = ctx->d_nbrCount, a.nbrIdx = ctx->d_nbrIdx, a.nbrDist = ctx->d_nbrDist, a.nbrTs = ctx->d_nbrTs;
    a.nbrTe = ctx->wantEnd ? ctx->d_nbrTe : nullptr;
    a.forceMode = forceMode, a.fp = fp, a.zero = zero, a.frc = ctx->d_frc, a.kick = kick, a.vel = ctx->d_vel;
    a.counters = ctx->d_counters;
    if ((rc = runGeodesicTiers(ctx, a, ctx->nLocal))) return rc;
    if (ctx->timing) recordEvent(ctx, ctx->ev[2]);
    ctx->nbrValid = true;
    return CSS_OK;
}

// checks the capacity counters after a synchronisation point; grows kmax when the stride was too small
static int checkCapacity(css_ctx* ctx, bool* rerun, int* stepsDone = nullptr)
{
    unsigned long long h[NUM_COUNTERS];
    CU(cudaMemcpyAsync(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    if (rerun) *rerun = false;
    if (h[C_KMAX_OVERFLOW]) { // the stride guard is up (common.cuh): nothing downstream of the cell list has run
        CU(cudaMemsetAsync(ctx->d_counters + C_KMAX_OVERFLOW, 0, sizeof(unsigned long long), ctx->st));
        CU(cudaMemsetAsync(ctx->d_counters + C_KMAX_NEED, 0, sizeof(unsigned long long), ctx->st));
        int need = (int)std::min<unsigned long long>(h[C_KMAX_NEED], 1ull << 24), k = ctx->kmax * 2;
        while (k < need) k *= 2;
        ctx->kmax = k;
        ctx->capNbr = 0;
        ctx->nveCalls = 0;
        if (stepsDone) *stepsDone = (int)h[C_STEP_GUARD];
        if (rerun) *rerun = true;
        else return fail(ctx, CSS_ECAPACITY, "neighbour stride exceeded; stride grown to %d, rerun", k);
    }
    if (h[C_OVERFLOW]) {
        CU(cudaMemsetAsync(ctx->d_counters + C_OVERFLOW, 0, sizeof(unsigned long long), ctx->st));
        return fail(ctx, CSS_ECAPACITY, "%llu sources exceeded the patch/window capacity of every tier", h[C_OVERFLOW]);
    }
    if (h[C_PEER_TIMEOUT]) {
        CU(cudaMemsetAsync(ctx->d_counters + C_PEER_TIMEOUT, 0, sizeof(unsigned long long), ctx->st));
        return fail(ctx, CSS_ENCCL, "peer position exchange timed out (%llu flag waits gave up): a rank left the collective sequence", h[C_PEER_TIMEOUT]);
    }
    return CSS_OK;
}

// phase durations of the last launches (the stream has been synchronised by the caller)
static void readPhaseTimes(css_ctx* ctx, bool walked)
{
    if (!ctx->timing) return;
    if (cudaEventElapsedTime(&ctx->msCell, ctx->ev[0], ctx->ev[1]) != cudaSuccess) ctx->msCell = 0;
    if (cudaEventElapsedTime(&ctx->msGeo, ctx->ev[1], ctx->ev[2]) != cudaSuccess) ctx->msGeo = 0;
    if (walked && cudaEventElapsedTime(&ctx->msWalk, ctx->ev[3], ctx->ev[4]) != cudaSuccess) ctx->msWalk = 0;
    (void)cudaGetLastError();
}

static ForceParams mkForce(int kind, const double* p, double* range)
{
    ForceParams f;
    f.kind = kind;
    f.a = p[0];
    f.sigma = p[1];
    *range = p[2];
    return f;
}

int css_find_neighbors(css_ctx* ctx, double range, int64_t* totalNeighbors)
{
    if (!ctx) return CSS_EINVAL;
    BIND();
    for (int attempt = 0; attempt < 8; ++attempt) {
        int rc = findNeighborsImpl(ctx, range, 0, ForceParams{0, 0, 0}, 0, 0.0);
        if (rc) return rc;
        bool rerun = false;
        if ((rc = checkCapacity(ctx, &rerun))) return rc;
        if (!rerun) break;
    }
    if (ctx->timing) {
        cudaEventElapsedTime(&ctx->msCell, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&ctx->msGeo, ctx->ev[1], ctx->ev[2]);
    }
    if (totalNeighbors) {
        std::vector<int> cnt(std::max(ctx->nLocal, 1));
        CU(cudaMemcpy(cnt.data(), ctx->d_nbrCount, sizeof(int) * ctx->nLocal, cudaMemcpyDeviceToHost));
        int64_t t = 0;
        for (int i = 0; i < ctx->nLocal; ++i) t += cnt[i];
        *totalNeighbors = t;
    }
    return CSS_OK;
}

int css_get_neighbors(css_ctx* ctx, int32_t* offsets, int32_t* idx, double* dist, double* startTan, double* endTan)
{
    if (!ctx || !ctx->nbrValid) return fail(ctx, CSS_ESTATE, "css_get_neighbors: no neighbour lists (call css_find_neighbors)");
    if (endTan && !ctx->nbrHasEnd) return fail(ctx, CSS_ESTATE, "end tangents were not requested (css_set_options)");
    BIND();
    int n = ctx->nLocal, km = ctx->kmax;
    std::vector<int> cnt(std::max(n, 1)), hidx;
    std::vector<double> hd, ht;
    CU(cudaMemcpy(cnt.data(), ctx->d_nbrCount, sizeof(int) * n, cudaMemcpyDeviceToHost));
    size_t tot = (size_t)n * km;
    if (idx) {
        hidx.resize(tot);
        CU(cudaMemcpy(hidx.data(), ctx->d_nbrIdx, sizeof(int) * tot, cudaMemcpyDeviceToHost));
    }
    if (dist) {
        hd.resize(tot);
        CU(cudaMemcpy(hd.data(), ctx->d_nbrDist, sizeof(double) * tot, cudaMemcpyDeviceToHost));
    }
    size_t o = 0;
    for (int i = 0; i < n; ++i) {
        if (offsets) offsets[i] = (int)o;
        for (int j = 0; j < cnt[i]; ++j, ++o) {
            if (idx) idx[o] = hidx[(size_t)i * km + j];
            if (dist) dist[o] = hd[(size_t)i * km + j];
        }
    }
    if (offsets) offsets[n] = (int)o;
    for (int pass = 0; pass < 2; ++pass) {
        double* out = pass ? endTan : startTan;
        if (!out) continue;
        ht.resize(3 * tot);
        CU(cudaMemcpy(ht.data(), pass ? ctx->d_nbrTe : ctx->d_nbrTs, sizeof(double) * 3 * tot, cudaMemcpyDeviceToHost));
        o = 0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < cnt[i]; ++j, ++o)
                for (int d = 0; d < 3; ++d) out[3 * o + d] = ht[3 * ((size_t)i * km + j) + d];
    }
    return CSS_OK;
}

int css_compute_forces(css_ctx* ctx, int kind, const double* params, int zero)
{
    if (!ctx || !params) return CSS_EINVAL;
    BIND();
    double range;
    ForceParams fp = mkForce(kind, params, &range);
    std::vector<double> saved;
    if (!zero) { // a rerun must restart from the caller's forces
        saved.resize(3 * (size_t)std::max(ctx->nLocal, 1));
        CU(cudaMemcpy(saved.data(), ctx->d_frc, sizeof(double) * 3 * ctx->nLocal, cudaMemcpyDeviceToHost));
    }
    for (int attempt = 0; attempt < 8; ++attempt) {
        int rc = findNeighborsImpl(ctx, range, 1, fp, zero, 0.0);
        if (rc) return rc;
        bool rerun = false;
        if ((rc = checkCapacity(ctx, &rerun))) return rc;
        if (!rerun) break;
        if (!zero) CU(cudaMemcpy(ctx->d_frc, saved.data(), sizeof(double) * 3 * ctx->nLocal, cudaMemcpyHostToDevice));
    }
    if (ctx->timing) {
        cudaEventElapsedTime(&ctx->msCell, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&ctx->msGeo, ctx->ev[1], ctx->ev[2]);
    }
    return CSS_OK;
}

int css_compute_energy(css_ctx* ctx, int kind, const double* params, double* energy)
{
    if (!ctx || !params || !energy) return CSS_EINVAL;
    BIND();
    double range;
    ForceParams fp = mkForce(kind, params, &range);
    int rc = css_find_neighbors(ctx, range, nullptr);
    if (rc) return rc;
    launchEnergy(ctx->st, ctx->nLocal, ctx->kmax, ctx->d_nbrCount, ctx->d_nbrDist, fp, ctx->d_partial, ctx->d_red);
    ctx->hostKernels += 2;
    CU(cudaMemcpyAsync(energy, ctx->d_red, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return css_reduce(ctx, CSS_SUM, 1, energy);
}

static int reduceDevice(css_ctx* ctx, double out[5], bool* guardUp = nullptr);
// simulation::computeMonodisperseStress (simulation.cpp:104-173) for one force computer: virial + kinetic parts of the
// Euclidean 3x3 "stress" of the surface, density = N / area
int css_compute_stress(css_ctx* ctx, int kind, const double* params, double stress[9])
{
    if (!ctx || !params || !stress) return CSS_EINVAL;
    if (!ctx->nTotal) return fail(ctx, CSS_ESTATE, "state not set");
    BIND();
    double range;
    ForceParams fp = mkForce(kind, params, &range);
    int rc = css_find_neighbors(ctx, range, nullptr);
    if (rc) return rc;
    launchStress(ctx->st, ctx->nLocal, ctx->kmax, ctx->d_nbrCount, ctx->d_nbrDist, ctx->d_nbrTs, ctx->d_vel, fp, ctx->d_partial, ctx->d_red);
    ctx->hostKernels += 2;
    double h[24] = {0};
    CU(cudaMemcpyAsync(h, ctx->d_red, sizeof(double) * 18, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    for (int o = 0; o < 18 && !rc; o += 8) rc = css_reduce(ctx, CSS_SUM, std::min(8, 18 - o), h + o);
    if (rc) return rc;
    const double N = ctx->nTotal, A = ctx->area, density = N / A;
    for (int q = 0; q < 9; ++q) stress[q] = density * 1.0 * h[9 + q] / (2 * N) + h[q] / (2 * 2 * A * N);
    return CSS_OK;
}
// noseHooverNVT::getTemperatureFromKE (noseHooverNVT.cpp:141-150): sum v.v / (2 Ndof)
int css_temperature(css_ctx* ctx, double* temperature)
{
    if (!ctx || !temperature) return CSS_EINVAL;
    if (!ctx->nTotal) return fail(ctx, CSS_ESTATE, "state not set");
    BIND();
    double r[5]; // {f.f, v.v, f.v, max f.f, KE}, already folded over the ranks
    int rc = reduceDevice(ctx, r);
    if (rc) return rc;
    *temperature = r[1] / (2.0 * ctx->nTotal);
    return CSS_OK;
}

static int moveImpl(css_ctx* ctx, int transportForce, int transportVelocity, int mode, double dt)
{
    int rc = ensurePeerWindow(ctx);
    if (rc) return rc;
    const bool p2p = ctx->nranks > 1 && ctx->pw.n > 1;
    if (p2p) { // the staging copies of the previous exchange must have been drained everywhere before they are overwritten
        launchPeerWaitConsumed(ctx->st, ctx->pw, ctx->d_epoch, ctx->d_counters);
        ctx->hostKernels++;
    }
    if (ctx->timing) recordEvent(ctx, ctx->ev[3]);
    launchWalk(ctx->st, meshDev(ctx), ctx->nLocal, ctx->minIdx, ctx->d_face, ctx->d_bary, ctx->d_disp, ctx->d_vel, ctx->d_frc, transportForce,
               transportVelocity, mode, dt, ctx->d_walkFlags, ctx->d_counters, p2p ? ctx->pw : PeerWin{});
    ctx->hostKernels++;
    if (ctx->timing) recordEvent(ctx, ctx->ev[4]);
    ctx->nbrValid = false;
    if (p2p) {
        launchPeerBarrier(ctx->st, ctx->pw, ctx->d_epoch, ctx->d_counters);
        launchPeerCopy(ctx->st, ctx->pw, ctx->nTotal, ctx->minIdx, ctx->minIdx + ctx->nLocal, ctx->d_face, ctx->d_bary, ctx->d_epoch, ctx->d_ticket);
        ctx->hostKernels += 2;
        CU(cudaGetLastError());
        return CSS_OK;
    }
    if (ctx->nranks > 1) return css_gather_positions(ctx);
    return CSS_OK;
}

int css_move(css_ctx* ctx, const double* disp, int transportForce, int transportVelocity)
{
    if (!ctx || !ctx->nTotal) return fail(ctx, CSS_ESTATE, "state not set");
    BIND();
    if (disp) CU(cudaMemcpyAsync(ctx->d_disp, disp, sizeof(double) * 3 * ctx->nLocal, cudaMemcpyHostToDevice, ctx->st));
    int rc = moveImpl(ctx, transportForce, transportVelocity, 0, 0.0);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->st));
    if (ctx->timing) cudaEventElapsedTime(&ctx->msWalk, ctx->ev[3], ctx->ev[4]);
    return CSS_OK;
}
int css_get_walk_flags(css_ctx* ctx, int32_t* flags)
{
    if (!ctx || !ctx->nTotal || !flags) return fail(ctx, CSS_ESTATE, "state not set");
    BIND();
    CU(cudaMemcpy(flags, ctx->d_walkFlags, sizeof(int) * ctx->nLocal, cudaMemcpyDeviceToHost));
    return CSS_OK;
}

// ---------------------------------------------------------------------------------------- updaters
static int nveStepLaunches(css_ctx* ctx, const ForceParams& fp, double range, double dt)
{
    // first half step fused into the walker; second half kick fused into the geodesic/force kernel
    int rc = moveImpl(ctx, 0, 1, 1, dt);
    if (rc) return rc;
    if (ctx->ioFace) { // host-buffer step: the positions are final here, their download overlaps the neighbour / force phase
        CU(cudaEventRecord(ctx->evFork, ctx->st));
        CU(cudaStreamWaitEvent(ctx->stCopy, ctx->evFork, 0));
        CU(cudaMemcpyAsync(ctx->ioFace, ctx->d_face, sizeof(int) * ctx->nTotal, cudaMemcpyDeviceToHost, ctx->stCopy));
        CU(cudaMemcpyAsync(ctx->ioBary, ctx->d_bary, sizeof(double) * 3 * ctx->nTotal, cudaMemcpyDeviceToHost, ctx->stCopy));
        CU(cudaEventRecord(ctx->evJoin, ctx->stCopy));
    }
    rc = findNeighborsImpl(ctx, range, 1, fp, 1, 0.5 * dt);
    if (ctx->ioFace) CU(cudaStreamWaitEvent(ctx->st, ctx->evJoin, 0));
    return rc;
}

// everything that decides what the captured launches look like: a change of any of these re-captures the graph
static uint64_t nveGraphKey(css_ctx* ctx, const ForceParams& fp, double range, double dt)
{
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) {
        const unsigned char* b = (const unsigned char*)p;
        for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
    };
#define MIX(x) mix(&(x), sizeof(x))
    MIX(fp.kind), MIX(fp.a), MIX(fp.sigma), MIX(range), MIX(dt);
    MIX(ctx->nLocal), MIX(ctx->nTotal), MIX(ctx->minIdx), MIX(ctx->kmax), MIX(ctx->nranks), MIX(ctx->timing), MIX(ctx->submeshing),
        MIX(ctx->maxDist), MIX(ctx->boundaryMode), MIX(ctx->useCellList), MIX(ctx->wantEnd), MIX(ctx->winWpb), MIX(ctx->winLean), MIX(ctx->winHalf), MIX(ctx->useStencil), MIX(ctx->stencilUsable), MIX(ctx->stencilMissing), MIX(ctx->grid), MIX(ctx->nCells);
    void* ptrs[] = {ctx->d_vert, ctx->d_corner, ctx->d_adj, ctx->d_adjopp, ctx->d_geo, ctx->d_saddle, ctx->d_face, ctx->d_bary, ctx->d_eucl, ctx->d_vel, ctx->d_frc,
                    ctx->d_disp, ctx->d_walkFlags, ctx->d_cellOf, ctx->d_cellCount, ctx->d_cellStart, ctx->d_cellSlot, ctx->d_tmpItems,
                    ctx->d_items, ctx->d_nbrCount, ctx->d_nbrIdx, ctx->d_nbrDist, ctx->d_nbrTs, ctx->d_nbrTe, ctx->d_retry[0], ctx->d_retry[1],
                    ctx->d_retry[2], ctx->d_retry[3], ctx->d_gws, ctx->d_records, ctx->d_recordsL, ctx->d_stencil, ctx->d_spill, ctx->d_recvI, ctx->d_recvD, (void*)ctx->comm, ctx->winLocal,
                    ctx->winPeer[0], ctx->winPeer[1], ctx->winPeer[2], ctx->winPeer[3], ctx->winPeer[4], ctx->winPeer[5], ctx->winPeer[6],
                    ctx->winPeer[7]};
    MIX(ptrs);
    MIX(ctx->ioFace), MIX(ctx->ioBary);
#undef MIX
    return h ? h : 1;
}

static int nveSteps(css_ctx* ctx, const ForceParams& fp, double range, double dt, int nsteps)
{
    int rc = CSS_OK;
    for (int s = 0; s < nsteps && !rc; ++s) {
        // The first two steps of a configuration run as plain launches (they size every buffer); after that the
        // step is replayed from a CUDA graph as long as nothing it captured has changed.
        uint64_t key = ctx->useGraph && !(ctx->ioFace && ctx->ioNoGraph) && ctx->nveCalls >= 2 ? nveGraphKey(ctx, fp, range, dt) : 0;
        if (key && key == ctx->nveKey && ctx->nveExec) {
            CU(cudaGraphLaunch(ctx->nveExec, ctx->st));
            ctx->hostKernels += ctx->nveKernels;
            ctx->nbrValid = true;
            continue;
        }
        if (key) {
            if (ctx->nveExec) cudaGraphExecDestroy(ctx->nveExec);
            ctx->nveExec = nullptr, ctx->nveKey = 0;
            unsigned long long k0 = ctx->hostKernels;
            cudaGraph_t g = nullptr;
            CU(cudaStreamBeginCapture(ctx->st, cudaStreamCaptureModeRelaxed));
            ctx->capturing = true;
            rc = nveStepLaunches(ctx, fp, range, dt);
            ctx->capturing = false;
            cudaError_t ce = cudaStreamEndCapture(ctx->st, &g);
            ctx->hostKernels = k0;
            bool same = rc == CSS_OK && ce == cudaSuccess && g && nveGraphKey(ctx, fp, range, dt) == key; // nothing was re-allocated meanwhile
            if (same && cudaGraphInstantiate(&ctx->nveExec, g, 0) == cudaSuccess) {
                ctx->nveKey = key;
                ctx->nveKernels = 0;
                size_t nn = 0;
                if (cudaGraphGetNodes(g, nullptr, &nn) == cudaSuccess) {
                    std::vector<cudaGraphNode_t> nodes(nn);
                    cudaGraphGetNodes(g, nodes.data(), &nn);
                    for (auto nd : nodes) {
                        cudaGraphNodeType t;
                        if (cudaGraphNodeGetType(nd, &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ctx->nveKernels++;
                    }
                }
            } else {
                (void)cudaGetLastError();
                ctx->nveExec = nullptr;
            }
            if (g) cudaGraphDestroy(g);
            rc = CSS_OK;
            if (ctx->nveExec) {
                CU(cudaGraphLaunch(ctx->nveExec, ctx->st));
                ctx->hostKernels += ctx->nveKernels;
                ctx->nbrValid = true;
                continue;
            }
            if (ctx->ioFace) ctx->ioNoGraph = true; // e.g. pageable host buffers cannot be captured: the host-buffer step runs plain launches
            else ctx->useGraph = false;             // capture is not possible in this configuration: plain launches from now on
        }
        rc = nveStepLaunches(ctx, fp, range, dt);
        ctx->nveCalls++;
    }
    return rc;
}

// nveSteps + the host half of the stride guard (common.cuh): when the cell-list build of some step found a stencil fuller than
// the neighbour stride, the rest of the call froze behind the guard.  Regrow the stride, finish that step (forces + second
// half kick from the positions its walker produced) and run the remaining steps.  Ends with a synchronisation.
static int nveStepsGuarded(css_ctx* ctx, const ForceParams& fp, double range, double dt, int nsteps)
{
    int remaining = nsteps;
    for (int attempt = 0; attempt < 16; ++attempt) {
        CU(cudaMemsetAsync(ctx->d_counters + C_STEP_GUARD, 0, sizeof(unsigned long long), ctx->st));
        int rc = nveSteps(ctx, fp, range, dt, remaining);
        if (rc) return rc;
        bool rerun = false;
        int done = 0;
        if ((rc = checkCapacity(ctx, &rerun, &done))) return rc;
        if (!rerun) return CSS_OK;
        if (done < 1 || done > remaining) return fail(ctx, CSS_ESTATE, "stride guard: inconsistent step count %d of %d", done, remaining);
        for (int a2 = 0;; ++a2) { // second half of step `done - 1`
            if ((rc = findNeighborsImpl(ctx, range, 1, fp, 1, 0.5 * dt))) return rc;
            if ((rc = checkCapacity(ctx, &rerun))) return rc;
            if (!rerun) break;
            if (a2 == 8) return fail(ctx, CSS_ECAPACITY, "neighbour stride keeps overflowing");
        }
        remaining -= done;
        if (remaining == 0) return CSS_OK;
    }
    return fail(ctx, CSS_ECAPACITY, "neighbour stride keeps overflowing");
}

int css_step_nve(css_ctx* ctx, int kind, const double* params, double dt, int nsteps)
{
    if (!ctx || !params) return CSS_EINVAL;
    if (!ctx->nTotal) return fail(ctx, CSS_ESTATE, "state not set");
    BIND();
    double range;
    ForceParams fp = mkForce(kind, params, &range);
    int rc = nveStepsGuarded(ctx, fp, range, dt, nsteps);
    if (ctx->timing && nsteps > 0) {
        cudaEventElapsedTime(&ctx->msCell, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&ctx->msGeo, ctx->ev[1], ctx->ev[2]);
        cudaEventElapsedTime(&ctx->msWalk, ctx->ev[3], ctx->ev[4]);
    }
    return rc;
}

// One velocity-Verlet step of a HOST-resident state (the reference keeps positions / velocities / forces in std::vectors and
// every performTimestep reads and writes them): upload, step, download in one call.  The positions are downloaded on a
// second stream as soon as the walker and the exchange are done, overlapping the neighbour / force phase.  Buffers should
// be page-locked for the copies to be asynchronous.  face/bary hold all nTotal particles, vel/frc this rank's block.
int css_step_nve_host(css_ctx* ctx, int kind, const double* params, double dt, int32_t* face, double* bary, double* vel, double* frc)
{
    if (!ctx || !params || !face || !bary || !vel || !frc) return CSS_EINVAL;
    if (!ctx->nTotal) return fail(ctx, CSS_ESTATE, "state not set (css_set_state fixes the sharding)");
    BIND();
    const int nT = ctx->nTotal, nL = ctx->nLocal;
    if (!ctx->stCopy) {
        CU(cudaStreamCreateWithFlags(&ctx->stCopy, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming));
    }
    double range;
    ForceParams fp = mkForce(kind, params, &range);
    CU(cudaMemcpyAsync(ctx->d_bary, bary, sizeof(double) * 3 * nT, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(ctx->d_vel, vel, sizeof(double) * 3 * nL, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(ctx->d_frc, frc, sizeof(double) * 3 * nL, cudaMemcpyHostToDevice, ctx->st));
    int bad = 0; // the face indices are validated while the other uploads fly, and uploaded last: a rejected call leaves the old faces in place
    for (int i = 0; i < nT; ++i) bad |= (face[i] < 0) | (face[i] >= ctx->nF);
    if (bad) {
        CU(cudaStreamSynchronize(ctx->st));
        ctx->nbrValid = false;
        return fail(ctx, CSS_EINVAL, "css_step_nve_host: face index out of range");
    }
    CU(cudaMemcpyAsync(ctx->d_face, face, sizeof(int) * nT, cudaMemcpyHostToDevice, ctx->st));
    ctx->ioFace = face, ctx->ioBary = bary;
    CU(cudaMemsetAsync(ctx->d_counters + C_STEP_GUARD, 0, sizeof(unsigned long long), ctx->st));
    int rc = nveSteps(ctx, fp, range, dt, 1);
    ctx->ioFace = nullptr, ctx->ioBary = nullptr;
    if (rc) return rc;
    CU(cudaMemcpyAsync(vel, ctx->d_vel, sizeof(double) * 3 * nL, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(frc, ctx->d_frc, sizeof(double) * 3 * nL, cudaMemcpyDeviceToHost, ctx->st));
    bool rerun = false;
    if ((rc = checkCapacity(ctx, &rerun))) return rc; // synchronises the stream (which has joined the copy stream)
    for (int a2 = 0; rerun; ++a2) { // stride guard (common.cuh): the positions (already downloaded) are final, forces and the second kick are not
        if (a2 == 8) return fail(ctx, CSS_ECAPACITY, "neighbour stride keeps overflowing");
        if ((rc = findNeighborsImpl(ctx, range, 1, fp, 1, 0.5 * dt))) return rc;
        CU(cudaMemcpyAsync(vel, ctx->d_vel, sizeof(double) * 3 * nL, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(frc, ctx->d_frc, sizeof(double) * 3 * nL, cudaMemcpyDeviceToHost, ctx->st));
        if ((rc = checkCapacity(ctx, &rerun))) return rc;
    }
    return CSS_OK;
}

int css_step_gd(css_ctx* ctx, int kind, const double* params, double dt, int nsteps)
{
    if (!ctx || !params) return CSS_EINVAL;
    if (!ctx->nTotal) return fail(ctx, CSS_ESTATE, "state not set");
    BIND();
    double range;
    ForceParams fp = mkForce(kind, params, &range);
    // All steps are queued without a host round trip.  When a particle first exceeds the neighbour stride, the force phase of that
    // step and every move after it return at once behind the stride guard (common.cuh); the walker's move counter then tells how
    // many steps are complete, the stride is regrown and the rest of the call is queued again.
    int remaining = nsteps;
    for (int attempt = 0; remaining > 0; ++attempt) {
        if (attempt == 16) return fail(ctx, CSS_ECAPACITY, "neighbour stride keeps overflowing");
        CU(cudaMemsetAsync(ctx->d_counters + C_STEP_GUARD, 0, sizeof(unsigned long long), ctx->st));
        for (int s = 0; s < remaining; ++s) {
            int rc = findNeighborsImpl(ctx, range, 1, fp, 1, 0.0);
            if (rc) return rc;
            launchAxpy(ctx->st, 1, ctx->nLocal, dt, 0, ctx->d_vel, ctx->d_frc, ctx->d_disp);
            ctx->hostKernels++;
            if ((rc = moveImpl(ctx, 0, 0, 0, 0.0))) return rc;
        }
        bool rerun = false;
        int moves = 0;
        if (int rc = checkCapacity(ctx, &rerun, &moves)) return rc; // synchronises
        if (!rerun) return CSS_OK;
        if (moves < 0 || moves >= remaining) return fail(ctx, CSS_ESTATE, "stride guard: inconsistent move count %d of %d steps", moves, remaining);
        remaining -= moves; // the step that raised the guard did nothing: it starts again with its force phase
    }
    return checkCapacity(ctx, nullptr);
}

// guardUp (optional): the stride-guard flag (common.cuh), read in the same synchronisation
static int reduceDevice(css_ctx* ctx, double out[5], bool* guardUp)
{
    launchReduce(ctx->st, ctx->nLocal, ctx->d_vel, ctx->d_frc, ctx->d_partial, ctx->d_red);
    ctx->hostKernels += 2;
    unsigned long long ovf = 0;
    CU(cudaMemcpyAsync(out, ctx->d_red, 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    if (guardUp) CU(cudaMemcpyAsync(&ovf, ctx->d_counters + C_KMAX_OVERFLOW, sizeof ovf, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    if (guardUp) *guardUp = ovf != 0;
    if (ctx->nranks > 1) {
        double s[4] = {out[0], out[1], out[2], out[4]};
        int rc = css_reduce(ctx, CSS_SUM, 4, s);
        if (rc) return rc;
        out[0] = s[0], out[1] = s[1], out[2] = s[2], out[4] = s[3];
        if ((rc = css_reduce(ctx, CSS_MAX, 1, out + 3))) return rc;
    }
    return CSS_OK;
}

int css_max_force(css_ctx* ctx, double* maxForce)
{
    if (!ctx || !ctx->nTotal || !maxForce) return fail(ctx, CSS_ESTATE, "state not set");
    BIND();
    double r[5];
    int rc = reduceDevice(ctx, r);
    if (rc) return rc;
    *maxForce = std::sqrt(r[3]);
    return CSS_OK;
}
int css_force_norm(css_ctx* ctx, double* forceNorm)
{
    if (!ctx || !ctx->nTotal || !forceNorm) return fail(ctx, CSS_ESTATE, "state not set");
    BIND();
    double r[5];
    int rc = reduceDevice(ctx, r);
    if (rc) return rc;
    *forceNorm = std::sqrt(r[0]);
    return CSS_OK;
}

// noseHooverNVT (src/updaters/noseHooverNVT.cpp); the chain is a handful of scalars and stays on the host
int css_nvt_init(css_ctx* ctx, double dt, double T, double tau, int M)
{
    if (!ctx || M < 1) return CSS_EINVAL;
    if (!ctx->nTotal) return fail(ctx, CSS_ESTATE, "state not set");
    auto& h = ctx->nh;
    h.dt = dt, h.dt2 = 0.5 * dt, h.dt4 = 0.25 * dt, h.dt8 = 0.125 * dt, h.T = T, h.tau = tau, h.M = M;
    h.bx.assign(M + 1, 0), h.by.assign(M + 1, 0), h.bz.assign(M + 1, 0), h.bw.assign(M + 1, 0);
    int Ndof = ctx->nTotal; // global particle count: the chain is replicated on every rank
    h.bw[0] = 2.0 * (Ndof - 1) * T * tau * tau;
    for (int i = 1; i <= M; ++i) h.bw[i] = T * tau * tau;
    h.KE = h.bw[0];
    h.scale = 1.0;
    h.hostNewer = true, h.deviceNewer = false;
    return CSS_OK;
}
// host mirror <-> device block of the chain
static int nhUpload(css_ctx* ctx)
{
    auto& h = ctx->nh;
    const int M = h.M, n = 4 * (M + 1) + 6;
    if (!h.d || h.dM != M) {
        if (h.d) cudaFree(h.d);
        h.d = nullptr;
        CU(cudaMalloc(&h.d, sizeof(double) * n));
        h.dM = M;
    }
    std::vector<double> buf(n);
    for (int i = 0; i <= M; ++i) buf[i] = h.bx[i], buf[M + 1 + i] = h.by[i], buf[2 * (M + 1) + i] = h.bz[i], buf[3 * (M + 1) + i] = h.bw[i];
    double* sc = buf.data() + 4 * (M + 1);
    sc[0] = h.KE, sc[1] = h.scale, sc[2] = h.T, sc[3] = h.dt2, sc[4] = h.dt4, sc[5] = h.dt8;
    CU(cudaMemcpyAsync(h.d, buf.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaStreamSynchronize(ctx->st)); // (buf is a local)
    h.hostNewer = false, h.deviceNewer = false;
    return CSS_OK;
}
static int nhDownload(css_ctx* ctx)
{
    auto& h = ctx->nh;
    if (!h.deviceNewer || !h.d) return CSS_OK;
    const int M = h.M, n = 4 * (M + 1) + 6;
    std::vector<double> buf(n);
    CU(cudaMemcpyAsync(buf.data(), h.d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    for (int i = 0; i <= M; ++i) h.bx[i] = buf[i], h.by[i] = buf[M + 1 + i], h.bz[i] = buf[2 * (M + 1) + i], h.bw[i] = buf[3 * (M + 1) + i];
    h.KE = buf[4 * (M + 1)], h.scale = buf[4 * (M + 1) + 1];
    h.deviceNewer = false;
    return CSS_OK;
}
static void propagateChain(css_ctx* ctx) // noseHooverNVT.cpp:65-110
{
    auto& h = ctx->nh;
    int M = h.M;
    double ef = 0;
    for (int ii = M - 1; ii > 0; --ii) {
        h.bz[ii] = (h.bw[ii - 1] * h.by[ii - 1] * h.by[ii - 1] - h.T) / h.bw[ii];
        ef = std::exp(-h.dt8 * h.by[ii + 1]);
        h.by[ii] *= ef;
        h.by[ii] += h.bz[ii] * h.dt4;
        h.by[ii] *= ef;
    }
    h.bz[0] = (2.0 * h.KE / h.bw[0] - 1.0);
    ef = std::exp(-h.dt8 * h.by[1]);
    h.by[0] *= ef;
    h.by[0] += h.bz[0] * h.dt4;
    h.by[0] *= ef;
    for (int ii = 0; ii < M; ++ii) h.bx[ii] += h.dt2 * h.by[ii];
    h.scale = std::exp(-h.dt2 * h.by[0]);
    h.KE = h.scale * h.scale * h.KE;
    h.bz[0] = (2.0 * h.KE / h.bw[0] - 1.0);
    ef = std::exp(-h.dt8 * h.by[1]);
    h.by[0] *= ef;
    h.by[0] += h.bz[0] * h.dt4;
    h.by[0] *= ef;
    for (int ii = 1; ii < M; ++ii) {
        h.bz[ii] = (h.bw[ii - 1] * h.by[ii - 1] * h.by[ii - 1] - h.T) / h.bw[ii];
        ef = std::exp(-h.dt8 * h.by[ii + 1]);
        h.by[ii] *= ef;
        h.by[ii] += h.bz[ii] * h.dt4;
        h.by[ii] *= ef;
    }
}
// Single rank: every step of the call is queued without a host round trip.  The chain (k_nh_chain) and the velocity scaling read
// the kinetic energy / the scale factor from device memory.  Like the NVE path the queue freezes behind the stride guard when a
// particle first exceeds the neighbour stride (cell-list build of some step): the walker, the stage kernels, the chain and the
// scaling all return at once, the state stays "after the first move of that step", and the host finishes the step and goes on.
static int stepNvtDevice(css_ctx* ctx, const ForceParams& fp, double range, int nsteps)
{
    auto& h = ctx->nh;
    int rc;
    if (h.hostNewer && (rc = nhUpload(ctx))) return rc;
    const int M = h.M;
    double* scaleDev = h.d + 4 * (M + 1) + 1;
    auto firstHalf = [&]() -> int { // chain, scale, displacement, move
        launchNhChain(ctx->st, h.d, M, ctx->d_red, 0, ctx->d_counters);
        launchScaleDev(ctx->st, ctx->nLocal, scaleDev, ctx->d_vel, ctx->d_counters);
        launchAxpy(ctx->st, 3, ctx->nLocal, h.dt2, 0, ctx->d_vel, ctx->d_frc, ctx->d_disp);
        ctx->hostKernels += 3;
        return moveImpl(ctx, 0, 1, 0, 0.0);
    };
    auto secondHalf = [&]() -> int { // forces + kick, displacement, kinetic energy, move, chain, scale
        int rc2 = findNeighborsImpl(ctx, range, 1, fp, 1, h.dt);
        if (rc2) return rc2;
        launchAxpy(ctx->st, 3, ctx->nLocal, h.dt2, 0, ctx->d_vel, ctx->d_frc, ctx->d_disp);
        launchReduce(ctx->st, ctx->nLocal, ctx->d_vel, ctx->d_frc, ctx->d_partial, ctx->d_red);
        ctx->hostKernels += 3;
        if ((rc2 = moveImpl(ctx, 0, 1, 0, 0.0))) return rc2;
        launchNhChain(ctx->st, h.d, M, ctx->d_red, 1, ctx->d_counters);
        launchScaleDev(ctx->st, ctx->nLocal, scaleDev, ctx->d_vel, ctx->d_counters);
        ctx->hostKernels += 2;
        return CSS_OK;
    };
    h.deviceNewer = true;
    int remaining = nsteps;
    for (int attempt = 0; remaining > 0; ++attempt) {
        if (attempt == 16) return fail(ctx, CSS_ECAPACITY, "neighbour stride keeps overflowing");
        CU(cudaMemsetAsync(ctx->d_counters + C_STEP_GUARD, 0, sizeof(unsigned long long), ctx->st));
        for (int s = 0; s < remaining; ++s) {
            if ((rc = firstHalf())) return rc;
            if ((rc = secondHalf())) return rc;
        }
        bool rerun = false;
        int moves = 0;
        if ((rc = checkCapacity(ctx, &rerun, &moves))) return rc; // synchronises
        if (!rerun) break;
        // the guard went up in the force phase of step (moves - 1) / 2: its first move happened, nothing after it did
        const int done = (moves - 1) / 2;
        if (moves < 1 || (moves & 1) == 0 || done >= remaining) return fail(ctx, CSS_ESTATE, "stride guard: inconsistent move count %d of %d steps", moves, remaining);
        for (int a2 = 0;; ++a2) {
            if ((rc = secondHalf())) return rc;
            if ((rc = checkCapacity(ctx, &rerun))) return rc;
            if (!rerun) break;
            if (a2 == 8) return fail(ctx, CSS_ECAPACITY, "neighbour stride keeps overflowing");
        }
        remaining -= done + 1;
    }
    if (nsteps > 0) readPhaseTimes(ctx, true);
    return CSS_OK;
}

int css_step_nvt(css_ctx* ctx, int kind, const double* params, int nsteps)
{
    if (!ctx || !params) return CSS_EINVAL;
    if (!ctx->nTotal || !ctx->nh.M) return fail(ctx, CSS_ESTATE, "css_nvt_init not called");
    BIND();
    double range;
    ForceParams fp = mkForce(kind, params, &range);
    auto& h = ctx->nh;
    if (ctx->nranks == 1 && ctx->nvtOnDevice) return stepNvtDevice(ctx, fp, range, nsteps);
    if (int rcd = nhDownload(ctx)) return rcd;
    h.hostNewer = true;
    for (int s = 0; s < nsteps; ++s) { // noseHooverNVT::performUpdate :42-59 and propagatePositionsVelocities :116-139
        propagateChain(ctx);
        launchAxpy(ctx->st, 2, ctx->nLocal, h.scale, 0, ctx->d_vel, ctx->d_frc, ctx->d_disp);
        launchAxpy(ctx->st, 3, ctx->nLocal, h.dt2, 0, ctx->d_vel, ctx->d_frc, ctx->d_disp);
        ctx->hostKernels += 2;
        int rc = moveImpl(ctx, 0, 1, 0, 0.0);
        if (rc) return rc;
        double r[5];
        for (int attempt = 0;; ++attempt) { // repeated with a larger neighbour stride when the stride guard went up (common.cuh)
            if ((rc = findNeighborsImpl(ctx, range, 1, fp, 1, h.dt))) return rc; // v += (dt/m) f fused as the kick
            launchAxpy(ctx->st, 3, ctx->nLocal, h.dt2, 0, ctx->d_vel, ctx->d_frc, ctx->d_disp);
            ctx->hostKernels++;
            bool guardUp = false, rerun = false;
            if ((rc = reduceDevice(ctx, r, &guardUp))) return rc;
            if (!guardUp) break;
            if (attempt == 8) return fail(ctx, CSS_ECAPACITY, "neighbour stride keeps overflowing");
            if ((rc = checkCapacity(ctx, &rerun))) return rc;
        }
        h.KE = r[4];
        if ((rc = moveImpl(ctx, 0, 1, 0, 0.0))) return rc;
        propagateChain(ctx);
        launchAxpy(ctx->st, 2, ctx->nLocal, h.scale, 0, ctx->d_vel, ctx->d_frc, ctx->d_disp);
        ctx->hostKernels++;
    }
    int rc = checkCapacity(ctx, nullptr);
    if (nsteps > 0) readPhaseTimes(ctx, true);
    return rc;
}
int css_nvt_state(css_ctx* ctx, double* bath, double* ke, double* scale)
{
    if (!ctx || !ctx->nh.M) return fail(ctx, CSS_ESTATE, "css_nvt_init not called");
    BIND();
    if (int rcd = nhDownload(ctx)) return rcd;
    auto& h = ctx->nh;
    for (int i = 0; i <= h.M && bath; ++i) bath[4 * i] = h.bx[i], bath[4 * i + 1] = h.by[i], bath[4 * i + 2] = h.bz[i], bath[4 * i + 3] = h.bw[i];
    if (ke) *ke = h.KE;
    if (scale) *scale = h.scale;
    return CSS_OK;
}

// fireMinimization (src/updaters/fireMinimization.{h,cpp})
int css_fire_init(css_ctx* ctx, const double* p, double dt0, double alpha0)
{
    if (!ctx) return CSS_EINVAL;
    auto& f = ctx->fire;
    f = css_ctx::FireState();
    f.dt = dt0, f.alpha = alpha0;
    if (p) { // setFIREParameters ignores its deltaT argument (fireMinimization.cpp:74-90)
        f.maximumIterations = (int)p[0];
        f.alphaStart = p[2], f.deltaTMax = p[3], f.deltaTMin = p[4], f.deltaTInc = p[5], f.deltaTDec = p[6], f.alphaDec = p[7];
        f.nMin = (int)p[8], f.forceCutoff = p[9], f.alphaMin = p[10];
        f.alpha = f.alphaStart;
    }
    return CSS_OK;
}
int css_fire_minimize(css_ctx* ctx, int kind, const double* params, double* out)
{
    if (!ctx || !params) return CSS_EINVAL;
    if (!ctx->nTotal) return fail(ctx, CSS_ESTATE, "state not set");
    BIND();
    double range;
    ForceParams fp = mkForce(kind, params, &range);
    auto& f = ctx->fire;
    int rc;
    double r[5];
    // force phase + reductions; repeated with a larger neighbour stride when the stride guard went up (common.cuh)
    auto forcesAndReduce = [&](double kick) -> int {
        for (int attempt = 0;; ++attempt) {
            int rc2 = findNeighborsImpl(ctx, range, 1, fp, 1, kick);
            if (rc2) return rc2;
            bool guardUp = false, rerun = false;
            if ((rc2 = reduceDevice(ctx, r, &guardUp))) return rc2;
            if (!guardUp) return CSS_OK;
            if (attempt == 8) return fail(ctx, CSS_ECAPACITY, "neighbour stride keeps overflowing");
            if ((rc2 = checkCapacity(ctx, &rerun))) return rc2;
        }
    };
    if ((rc = forcesAndReduce(0.0))) return rc;
    f.forceMax = std::sqrt(r[3]);
    f.iterations = 0;
    while (f.iterations < f.maximumIterations && f.forceMax > f.forceCutoff) { // minimizeByFire :3-21
        f.iterations += 1;
        if ((rc = moveImpl(ctx, 1, 1, 1, f.dt))) return rc;                       // first half + move, transporting [force, velocity]
        if ((rc = forcesAndReduce(0.5 * f.dt))) return rc;                         // forces + second half kick; fireStep :36-72
        double forceNorm = r[0], velocityNorm = r[1], power = r[2];
        double scaling = 0.0;
        if (forceNorm > 0) scaling = std::sqrt(velocityNorm / forceNorm);
        launchAxpy(ctx->st, 4, ctx->nLocal, scaling, f.alpha, ctx->d_vel, ctx->d_frc, ctx->d_disp);
        ctx->hostKernels++;
        if (power > 0) {
            if (f.nSinceNegativePower > f.nMin) {
                f.dt = std::min(f.dt * f.deltaTInc, f.deltaTMax);
                f.alpha = f.alpha * f.alphaDec;
                f.alpha = std::max(f.alpha, f.alphaMin);
            }
            f.nSinceNegativePower += 1;
        } else {
            f.nSinceNegativePower = 0;
            f.dt = f.dt * f.deltaTDec;
            f.dt = std::max(f.dt, f.deltaTMin);
            f.alpha = f.alphaStart;
            launchAxpy(ctx->st, 5, ctx->nLocal, 0, 0, ctx->d_vel, ctx->d_frc, ctx->d_disp);
            ctx->hostKernels++;
        }
        f.forceMax = std::sqrt(r[3]);
    }
    if (out) out[0] = f.iterations, out[1] = f.forceMax, out[2] = f.dt, out[3] = f.alpha;
    rc = checkCapacity(ctx, nullptr);
    readPhaseTimes(ctx, f.iterations > 0);
    return rc;
}

// --------------------------------------------------------------------------------------- multi-GPU
int css_comm_unique_id(void* id128)
{
    if (!id128) return CSS_EINVAL;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return CSS_ENCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    std::memcpy(id128, &id, 128);
    return CSS_OK;
}
int css_comm_init(css_ctx* ctx, int rank, int nranks, const void* id128)
{
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return CSS_EINVAL;
    BIND();
    releasePeerWindow(ctx);
    ctx->p2pFailed = false;
    ctx->rank = rank, ctx->nranks = nranks;
    ctx->nveCalls = 0;
    if (nranks == 1) return CSS_OK;
    if (!id128) return CSS_EINVAL;
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    NC(ncclCommInitRank(&ctx->comm, nranks, id, rank));
    CU(regrow(ctx->d_redBuf, 64 * (size_t)nranks));
    return CSS_OK;
}
int css_comm_info(css_ctx* ctx, int* rank, int* nranks, int* peerExchange)
{
    if (!ctx) return CSS_EINVAL;
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    if (peerExchange) *peerExchange = ctx->nranks > 1 && ctx->pw.n > 1;
    return CSS_OK;
}
// mpiSimulation::synchronizeAndTransferBuffers: every rank contributes a block padded to per = ceil(N/R)
int css_gather_positions(css_ctx* ctx)
{
    if (!ctx || !ctx->nTotal) return fail(ctx, CSS_ESTATE, "state not set");
    if (ctx->nranks == 1) return CSS_OK;
    if (!ctx->comm) return fail(ctx, CSS_ENCCL, "communicator not initialised");
    BIND();
    int per = (ctx->nTotal + ctx->nranks - 1) / ctx->nranks;
    if (ctx->minIdx != ctx->rank * per) return fail(ctx, CSS_EINVAL, "sharding does not follow mpiModel::determineIndexBounds");
    if ((size_t)per * ctx->nranks > (size_t)ctx->nTotal + CSS_GATHER_PAD) return fail(ctx, CSS_EINVAL, "css_gather_positions: more than %d ranks", CSS_GATHER_PAD);
    if (per * ctx->nranks == ctx->nTotal) { // equal blocks: gather in place, no scratch and no copies
        NC(ncclGroupStart());
        NC(ncclAllGather(ctx->d_face + ctx->minIdx, ctx->d_face, per, ncclInt32, ctx->comm, ctx->st));
        NC(ncclAllGather(ctx->d_bary + 3 * (size_t)ctx->minIdx, ctx->d_bary, 3 * (size_t)per, ncclFloat64, ctx->comm, ctx->st));
        NC(ncclGroupEnd());
        return CSS_OK;
    }
    // the replicated arrays are already laid out rank-block by rank-block; only the last block is short,
    // so gather into padded scratch and copy back the first nTotal entries
    if (per > ctx->capComm) {
        CU(regrow(ctx->d_recvI, (size_t)per * ctx->nranks));
        CU(regrow(ctx->d_recvD, 3 * (size_t)per * ctx->nranks));
        ctx->capComm = per;
    }
    NC(ncclGroupStart());
    NC(ncclAllGather(ctx->d_face + ctx->minIdx, ctx->d_recvI, per, ncclInt32, ctx->comm, ctx->st));
    NC(ncclAllGather(ctx->d_bary + 3 * (size_t)ctx->minIdx, ctx->d_recvD, 3 * (size_t)per, ncclFloat64, ctx->comm, ctx->st));
    NC(ncclGroupEnd());
    CU(cudaMemcpyAsync(ctx->d_face, ctx->d_recvI, sizeof(int) * ctx->nTotal, cudaMemcpyDeviceToDevice, ctx->st));
    CU(cudaMemcpyAsync(ctx->d_bary, ctx->d_recvD, sizeof(double) * 3 * ctx->nTotal, cudaMemcpyDeviceToDevice, ctx->st));
    return CSS_OK;
}
// mpiSimulation::manipulateUpdaterData: all-gather k doubles per rank, fold in rank order
int css_reduce(css_ctx* ctx, int op, int k, double* data)
{
    if (!ctx || k < 0 || k > 8 || !data) return CSS_EINVAL;
    if (ctx->nranks == 1) return CSS_OK;
    if (!ctx->comm) return fail(ctx, CSS_ENCCL, "communicator not initialised");
    BIND();
    // d_red[0..8) is the send block, d_redBuf[0 .. 8*nranks) receives
    CU(cudaMemcpyAsync(ctx->d_red, data, sizeof(double) * k, cudaMemcpyHostToDevice, ctx->st));
    NC(ncclAllGather(ctx->d_red, ctx->d_redBuf, 8, ncclFloat64, ctx->comm, ctx->st));
    std::vector<double> h(8 * (size_t)ctx->nranks);
    CU(cudaMemcpyAsync(h.data(), ctx->d_redBuf, sizeof(double) * 8 * ctx->nranks, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    for (int i = 0; i < k; ++i) {
        double acc = 0.0; // mpiSimulation.cpp:78-88 starts from 0 for both sum and max
        for (int r = 0; r < ctx->nranks; ++r) {
            double y = h[8 * (size_t)r + i];
            acc = op == CSS_MAX ? std::max(acc, y) : acc + y;
        }
        data[i] = acc;
    }
    return CSS_OK;
}

// ----------------------------------------------------------------------------------- diagnostics
int css_counters(css_ctx* ctx, uint64_t* out, int reset)
{
    if (!ctx || !out) return CSS_EINVAL;
    BIND();
    unsigned long long h[NUM_COUNTERS];
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaMemcpy(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    for (int i = 0; i < CSS_NUM_COUNTERS; ++i) out[i] = h[i];
    out[CSS_C_KERNELS] = ctx->hostKernels;
    if (reset) {
        CU(cudaMemset(ctx->d_counters, 0, sizeof h));
        ctx->hostKernels = 0;
    }
    return CSS_OK;
}
int css_synchronize(css_ctx* ctx)
{
    if (!ctx) return CSS_EINVAL;
    BIND();
    CU(cudaStreamSynchronize(ctx->st));
    return CSS_OK;
}
int css_device_positions(css_ctx* ctx, void** face_dev, void** bary_dev)
{
    if (!ctx || !ctx->nTotal) return fail(ctx, CSS_ESTATE, "state not set");
    if (face_dev) *face_dev = ctx->d_face;
    if (bary_dev) *bary_dev = ctx->d_bary;
    return CSS_OK;
}
int css_set_timing(css_ctx* ctx, int enabled)
{
    if (!ctx) return CSS_EINVAL;
    if ((enabled != 0) != ctx->timing) ctx->nveCalls = 0;
    ctx->timing = enabled != 0;
    return CSS_OK;
}
int css_last_kernel_ms(css_ctx* ctx, float* geodesic_ms, float* walk_ms, float* celllist_ms)
{
    if (!ctx) return CSS_EINVAL;
    if (geodesic_ms) *geodesic_ms = ctx->msGeo;
    if (walk_ms) *walk_ms = ctx->msWalk;
    if (celllist_ms) *celllist_ms = ctx->msCell;
    return CSS_OK;
}

int css_last_stage_ms(css_ctx* ctx, float* patch_ms, float* window_ms, float* retry_ms, float* gather_ms)
{
    if (!ctx) return CSS_EINVAL;
    if (!ctx->timing) return fail(ctx, CSS_ESTATE, "stage timing needs css_set_timing(1)");
    BIND();
    CU(cudaEventSynchronize(ctx->ev[2]));
    float a = 0, b = 0, c = 0;
    CU(cudaEventElapsedTime(&a, ctx->evS[0], ctx->evS[1]));
    CU(cudaEventElapsedTime(&b, ctx->evS[1], ctx->evS[2]));
    CU(cudaEventElapsedTime(&c, ctx->evS[2], ctx->ev[2]));
    if (patch_ms) *patch_ms = a;
    if (window_ms) *window_ms = b;
    if (retry_ms) *retry_ms = c;
    if (gather_ms) { // walker end -> start of the neighbour phase: the position all-gather (0 on one rank)
        float g = 0;
        if (cudaEventElapsedTime(&g, ctx->ev[4], ctx->ev[0]) != cudaSuccess) g = 0, (void)cudaGetLastError();
        *gather_ms = g > 0 ? g : 0; // (an NVT step ends with a move: its last walker event follows the neighbour phase)
    }
    return CSS_OK;
}

int css_timer_record(css_ctx* ctx, int slot)
{
    if (!ctx || slot < 0 || slot >= 8) return CSS_EINVAL;
    BIND();
    CU(cudaEventRecord(ctx->tev[slot], ctx->st));
    return CSS_OK;
}
int css_timer_elapsed_ms(css_ctx* ctx, int slotA, int slotB, float* ms)
{
    if (!ctx || slotA < 0 || slotA >= 8 || slotB < 0 || slotB >= 8 || !ms) return CSS_EINVAL;
    BIND();
    CU(cudaEventSynchronize(ctx->tev[slotB]));
    CU(cudaEventElapsedTime(ms, ctx->tev[slotA], ctx->tev[slotB]));
    return CSS_OK;
}

int css_microbench(css_ctx* ctx, int what, int reps, double* value)
{
    if (!ctx || !value || what < 0 || what > 1 || reps < 1) return CSS_EINVAL;
    BIND();
    CU(cudaStreamSynchronize(ctx->st));
    double v = runMicrobench(ctx->st, ctx->numSMs, what, reps);
    ctx->hostKernels += (unsigned long long)reps + 1;
    if (v < 0) return fail(ctx, CSS_ECUDA, "css_microbench: launch failed");
    *value = v;
    return CSS_OK;
}

} // extern "C"
#pragma GCC visibility pop
