// ORACLE (test infrastructure only): C entry points so tests/ and bench.py's cpu_baseline leg can
// drive the CPU restatement through ctypes.  Nothing in curvedspacesim_b200/ may link or call this.
#include "locate.hpp"
#include "sim.hpp"
#include <chrono>
#include <cstring>

using namespace orc;

#define ORC_API extern "C" __attribute__((visibility("default")))

static thread_local char g_err[512];

ORC_API const char* orc_last_error() { return g_err; }

ORC_API void* orc_create(int nV, const double* xyz, int nF, const int* corners)
{
    try {
        Sim* s = new Sim();
        s->setMesh(nV, xyz, nF, corners);
        return s;
    } catch (std::exception& e) {
        std::snprintf(g_err, sizeof g_err, "%s", e.what());
        return nullptr;
    }
}
ORC_API void orc_destroy(void* h) { delete (Sim*)h; }

ORC_API void orc_mesh_info(void* h, double* bbmin, double* bbmax, double* area)
{
    Sim* s = (Sim*)h;
    for (int d = 0; d < 3; ++d) {
        bbmin[d] = s->mesh.bbmin[d];
        bbmax[d] = s->mesh.bbmax[d];
    }
    *area = s->mesh.area();
}
ORC_API void orc_get_adjacency(void* h, int* adj, int* adjk)
{
    Sim* s = (Sim*)h;
    std::memcpy(adj, s->mesh.adj.data(), sizeof(int) * s->mesh.adj.size());
    std::memcpy(adjk, s->mesh.adjk.data(), sizeof(int) * s->mesh.adjk.size());
}
ORC_API void orc_get_saddle(void* h, char* out)
{
    Sim* s = (Sim*)h;
    std::memcpy(out, s->saddle.data(), s->saddle.size());
}

// simpleModel::R3PositionsToMeshPositions (src/models/simpleModel.cpp:136-154), brute force over the faces
ORC_API void orc_locate(void* h, int n, const double* xyz, double clampTol, int* face, double* bary)
{
    Sim* s = (Sim*)h;
    for (int i = 0; i < n; ++i) locatePoint(s->mesh, V3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, clampTol, face[i], bary + 3 * i);
}

ORC_API void orc_set_submeshing(void* h, int enabled, double maxDist)
{
    Sim* s = (Sim*)h;
    s->submeshing = enabled != 0;
    s->maximumDistance = maxDist;
}
ORC_API void orc_set_options(void* h, int useCellList, int strictTrig, int nThreads)
{
    Sim* s = (Sim*)h;
    s->useCellList = useCellList != 0;
    s->strictTrig = strictTrig != 0;
    s->nThreads = nThreads;
}
// override of the cell-list domain (cellListNeighborStructure ctor, cellListNeighborStructure.cpp:4-15)
ORC_API void orc_set_boundary(void* h, int mode) { ((Sim*)h)->boundaryMode = mode; }

ORC_API void orc_set_cell_domain(void* h, const double* mn, const double* mx)
{
    Sim* s = (Sim*)h;
    s->cl.setDomain(V3{mn[0], mn[1], mn[2]}, V3{mx[0], mx[1], mx[2]});
}

ORC_API void orc_set_state(void* h, int N, const int* face, const double* bary, const double* vel, const double* frc)
{
    Sim* s = (Sim*)h;
    if (s->N != N) s->resize(N);
    std::memcpy(s->face.data(), face, sizeof(int) * N);
    std::memcpy(s->bary.data(), bary, sizeof(double) * 3 * N);
    for (int i = 0; i < N; ++i) {
        if (vel) s->vel[i] = V3{vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]};
        if (frc) s->frc[i] = V3{frc[3 * i], frc[3 * i + 1], frc[3 * i + 2]};
    }
}
ORC_API void orc_get_state(void* h, int* face, double* bary, double* vel, double* frc)
{
    Sim* s = (Sim*)h;
    int N = s->N;
    if (face) std::memcpy(face, s->face.data(), sizeof(int) * N);
    if (bary) std::memcpy(bary, s->bary.data(), sizeof(double) * 3 * N);
    for (int i = 0; i < N; ++i) {
        if (vel) vel[3 * i] = s->vel[i].x, vel[3 * i + 1] = s->vel[i].y, vel[3 * i + 2] = s->vel[i].z;
        if (frc) frc[3 * i] = s->frc[i].x, frc[3 * i + 1] = s->frc[i].y, frc[3 * i + 2] = s->frc[i].z;
    }
}

ORC_API void orc_euclidean(void* h, int n, const int* face, const double* bary, double* out)
{
    Sim* s = (Sim*)h;
    for (int i = 0; i < n; ++i) {
        V3 p = s->mesh.point(face[i], bary + 3 * i);
        out[3 * i] = p.x, out[3 * i + 1] = p.y, out[3 * i + 2] = p.z;
    }
}

// cell-list candidates for the current state: CSR (offsets has N+1 entries); returns total, fills <= cap
ORC_API long orc_candidates(void* h, double range, int* offsets, int* idx, long cap, double* maxDist)
{
    Sim* s = (Sim*)h;
    s->fillEuclidean();
    s->cl.setRange(range);
    s->cl.build(s->eucl);
    long tot = 0;
    std::vector<int> c;
    for (int i = 0; i < s->N; ++i) {
        offsets[i] = (int)tot;
        double R = s->cl.candidates(i, c);
        if (maxDist) maxDist[i] = R;
        for (int j : c) {
            if (tot < cap) idx[tot] = j;
            tot++;
        }
    }
    offsets[s->N] = (int)tot;
    return tot;
}
ORC_API void orc_cell_grid(void* h, double range, int* n, double* cs)
{
    Sim* s = (Sim*)h;
    s->cl.setRange(range);
    for (int d = 0; d < 3; ++d) n[d] = s->cl.n[d], cs[d] = s->cl.cs[d];
}

ORC_API int orc_patch(void* h, int sf, const double* sb, int K, const int* tf, double R, int* out, int cap)
{
    Sim* s = (Sim*)h;
    std::vector<int> t(tf, tf + K);
    std::vector<int> pf = patchFaces(s->mesh, sf, sb, t, R);
    for (size_t i = 0; i < pf.size() && (int)i < cap; ++i) out[i] = pf[i];
    return (int)pf.size();
}

// triangulatedMeshSpace::distance; stats[0..4] = windows created, processed, pseudo-sources, patch faces, patch verts
ORC_API int orc_distance(void* h, int sf, const double* sb, int K, const int* tf, const double* tb, double threshold,
                         double* dist, double* ts, double* te, int* tie, long* stats)
{
    Sim* s = (Sim*)h;
    std::vector<GeoTarget> tg(K);
    for (int i = 0; i < K; ++i) {
        tg[i].face = tf[i];
        tg[i].b[0] = tb[3 * i], tg[i].b[1] = tb[3 * i + 1], tg[i].b[2] = tb[3 * i + 2];
    }
    std::vector<GeoResult> res;
    GeoStats st;
    s->distance(sf, sb, tg, threshold, res, &st);
    for (int i = 0; i < K; ++i) {
        dist[i] = res[i].dist;
        if (ts) ts[3 * i] = res[i].ts.x, ts[3 * i + 1] = res[i].ts.y, ts[3 * i + 2] = res[i].ts.z;
        if (te) te[3 * i] = res[i].te.x, te[3 * i + 1] = res[i].te.y, te[3 * i + 2] = res[i].te.z;
        if (tie) tie[i] = res[i].tie;
    }
    if (stats) {
        stats[0] = st.windowsCreated, stats[1] = st.windowsProcessed, stats[2] = st.pseudoSources, stats[3] = st.faces,
        stats[4] = st.verts;
    }
    return 0;
}

// triangulatedMeshSpace::transportParticleAndVectors for n independent particles; vecs is [n][nVec][3]
ORC_API void orc_transport(void* h, int n, int* face, double* bary, double* disp, int nVec, double* vecs, int* flags, int* crossings)
{
    Sim* s = (Sim*)h;
    for (int i = 0; i < n; ++i) {
        V3 d{disp[3 * i], disp[3 * i + 1], disp[3 * i + 2]};
        V3 T[8];
        for (int j = 0; j < nVec && j < 8; ++j)
            T[j] = V3{vecs[3 * (i * nVec + j)], vecs[3 * (i * nVec + j) + 1], vecs[3 * (i * nVec + j) + 2]};
        int cr = 0;
        int fl = transport(s->mesh, face[i], bary + 3 * i, d, T, nVec, s->strictTrig, &cr, s->boundaryMode);
        if (flags) flags[i] = fl;
        if (crossings) crossings[i] = cr;
        disp[3 * i] = d.x, disp[3 * i + 1] = d.y, disp[3 * i + 2] = d.z;
        for (int j = 0; j < nVec && j < 8; ++j)
            vecs[3 * (i * nVec + j)] = T[j].x, vecs[3 * (i * nVec + j) + 1] = T[j].y, vecs[3 * (i * nVec + j) + 2] = T[j].z;
    }
}

static PairForce mkForce(int kind, const double* p)
{
    PairForce f;
    f.kind = kind;
    if (kind == FORCE_HARMONIC) {
        f.k = p[0], f.sigma = p[1], f.range = p[2];
    } else {
        f.alpha = p[0], f.gsigma = p[1], f.range = p[2];
    }
    return f;
}

ORC_API long orc_find_neighbors(void* h, double range)
{
    Sim* s = (Sim*)h;
    s->findNeighbors(range);
    long tot = 0;
    for (auto& v : s->nbr) tot += (long)v.size();
    return tot;
}
ORC_API void orc_get_neighbors(void* h, int* offsets, int* idx, double* dist, double* ts, double* te)
{
    Sim* s = (Sim*)h;
    long tot = 0;
    for (int i = 0; i < s->N; ++i) {
        offsets[i] = (int)tot;
        for (size_t jj = 0; jj < s->nbr[i].size(); ++jj, ++tot) {
            if (idx) idx[tot] = s->nbr[i][jj];
            if (dist) dist[tot] = s->nbrDist[i][jj];
            if (ts) ts[3 * tot] = s->nbrStart[i][jj].x, ts[3 * tot + 1] = s->nbrStart[i][jj].y, ts[3 * tot + 2] = s->nbrStart[i][jj].z;
            if (te) te[3 * tot] = s->nbrEnd[i][jj].x, te[3 * tot + 1] = s->nbrEnd[i][jj].y, te[3 * tot + 2] = s->nbrEnd[i][jj].z;
        }
    }
    offsets[s->N] = (int)tot;
}
ORC_API void orc_compute_forces(void* h, int kind, const double* params, int zero)
{
    Sim* s = (Sim*)h;
    s->computeForces(mkForce(kind, params), zero != 0);
}
ORC_API double orc_compute_energy(void* h, int kind, const double* params)
{
    Sim* s = (Sim*)h;
    return s->computeEnergy(mkForce(kind, params));
}
ORC_API void orc_compute_stress(void* h, int kind, const double* params, double* stress9)
{
    Sim* s = (Sim*)h;
    s->computeStress(mkForce(kind, params), s->mesh.area(), stress9);
}
ORC_API double orc_temperature(void* h) { return ((Sim*)h)->temperature(); }
ORC_API void orc_move(void* h, double* disp, int transportForce, int transportVelocity)
{
    Sim* s = (Sim*)h;
    s->transportForce = transportForce, s->transportVelocity = transportVelocity;
    std::vector<V3> d(s->N);
    for (int i = 0; i < s->N; ++i) d[i] = V3{disp[3 * i], disp[3 * i + 1], disp[3 * i + 2]};
    s->moveParticles(d);
    for (int i = 0; i < s->N; ++i) disp[3 * i] = d[i].x, disp[3 * i + 1] = d[i].y, disp[3 * i + 2] = d[i].z;
}
ORC_API void orc_get_walk_flags(void* h, int* flags)
{
    Sim* s = (Sim*)h;
    std::memcpy(flags, s->walkFlags.data(), sizeof(int) * s->N);
}

// returns seconds spent in the step loop
ORC_API double orc_run_nve(void* h, int kind, const double* params, double dt, int steps)
{
    Sim* s = (Sim*)h;
    PairForce pf = mkForce(kind, params);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; ++i) s->stepNVE(dt, pf);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
ORC_API double orc_run_gd(void* h, int kind, const double* params, double dt, int steps)
{
    Sim* s = (Sim*)h;
    PairForce pf = mkForce(kind, params);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; ++i) s->stepGD(dt, pf);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
ORC_API void orc_nvt_init(void* h, double dt, double T, double tau, int M)
{
    Sim* s = (Sim*)h;
    s->nh.init(dt, T, tau, M, s->N);
}
ORC_API double orc_run_nvt(void* h, int kind, const double* params, int steps)
{
    Sim* s = (Sim*)h;
    PairForce pf = mkForce(kind, params);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; ++i) s->stepNVT(pf);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
ORC_API void orc_nvt_state(void* h, double* bath /* 4*(M+1) */, double* ke, double* scale)
{
    Sim* s = (Sim*)h;
    for (int i = 0; i <= s->nh.M; ++i)
        bath[4 * i] = s->nh.bx[i], bath[4 * i + 1] = s->nh.by[i], bath[4 * i + 2] = s->nh.bz[i], bath[4 * i + 3] = s->nh.bw[i];
    *ke = s->nh.KE;
    *scale = s->nh.scale;
}
// p = {maximumIterations, deltaT, alphaStart, deltaTMax, deltaTMin, deltaTInc, deltaTDec, alphaDec, nMin, forceCutoff, alphaMin}
// note fireMinimization::setFIREParameters ignores its deltaT argument (fireMinimization.cpp:74-90); dt0 is separate
ORC_API void orc_fire_init(void* h, const double* p, double dt0, double alpha0)
{
    Sim* s = (Sim*)h;
    auto& f = s->fire;
    f = Sim::Fire();
    f.dt = dt0, f.alpha = alpha0;
    if (p) {
        f.maximumIterations = (int)p[0];
        f.alphaStart = p[2], f.deltaTMax = p[3], f.deltaTMin = p[4], f.deltaTInc = p[5], f.deltaTDec = p[6], f.alphaDec = p[7];
        f.nMin = (int)p[8], f.forceCutoff = p[9], f.alphaMin = p[10];
        f.alpha = f.alphaStart;
    }
}
ORC_API double orc_run_fire(void* h, int kind, const double* params, double* out /* iterations, forceMax, dt, alpha */)
{
    Sim* s = (Sim*)h;
    PairForce pf = mkForce(kind, params);
    auto t0 = std::chrono::steady_clock::now();
    s->minimizeByFire(pf);
    double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (out) out[0] = s->fire.iterations, out[1] = s->fire.forceMax, out[2] = s->fire.dt, out[3] = s->fire.alpha;
    return el;
}
// counters: [0..4] walker flags (vertex,nohit,itercap,nan,border) [5] disconnected [6] ties [7] crossings
//           [8] windows created [9] windows processed [10] pseudo-sources [11] sum patch faces [12] sum patch verts
ORC_API void orc_counters(void* h, long* out, int reset)
{
    Sim* s = (Sim*)h;
    for (int i = 0; i < 7; ++i) out[i] = s->flagCounts[i];
    out[7] = s->crossings;
    out[8] = s->stats.windowsCreated, out[9] = s->stats.windowsProcessed, out[10] = s->stats.pseudoSources;
    out[11] = s->stats.faces, out[12] = s->stats.verts;
    if (reset) {
        for (auto& c : s->flagCounts) c = 0;
        s->crossings = 0;
        s->stats = GeoStats();
    }
}
