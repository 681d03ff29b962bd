// ORACLE (test infrastructure only). Model + pair forces + updaters, restating
//   simpleModel::findNeighbors / moveParticles        src/models/simpleModel.cpp:44-112
//   mpiModel sharding                                 src/models/mpiModel.cpp:20-31,75-157
//   triangulatedMeshSpace::distance[WithSubmeshing]   src/models/triangulatedMeshSpace.cpp:155-238
//   force::computeForces / computeEnergy              src/forces/baseForce.cpp:12-44
//   harmonicRepulsion / gaussianRepulsion             src/forces/harmonicRepulsion.cpp:3-33, gaussianRepulsion.{h,cpp}
//   velocityVerletNVE / noseHooverNVT / FIRE / GD     src/updaters/*.cpp
// "Ranks" are emulated in one address space: positions are replicated by construction, so the
// MPI_Allgather of src/simulation/mpiSimulation.cpp:11-42 is the identity; worker threads shard
// particles exactly like mpiModel::determineIndexBounds.
#pragma once
#include "celllist.hpp"
#include "geodesic.hpp"
#include "patch.hpp"
#include "walker.hpp"
#include <functional>
#include <thread>

namespace orc {

enum ForceKind { FORCE_HARMONIC = 0, FORCE_GAUSSIAN = 1 };

struct PairForce {
    int kind = FORCE_HARMONIC;
    double k = 1, sigma = 1;      // harmonic: stiffness, range
    double alpha = 1, gsigma = 1; // gaussian: strength, "variance" (as named by the reference)
    double range = 1;             // force::maximumInteractionRange (baseForce.h:57 default 1)
    V3 force(const V3& sep, double d) const
    {
        if (kind == FORCE_HARMONIC) { // harmonicRepulsion.cpp:19-33
            V3 ans{0, 0, 0};
            if (d <= sigma) ans = (-k * (sigma - d)) * sep;
            return ans;
        }
        // gaussianRepulsion.h:16-24, gaussianRepulsion.cpp:8-12 (sigma^{3/2} as coded)
        const double sqrtTwoPi = 2.50662827463100050241576528481104525300698674061;
        double twoSigmaSquared = 2.0 * gsigma * gsigma;
        double sigmaThreeHalvesSqrtTwoPi = (sqrtTwoPi * gsigma) * std::sqrt(gsigma);
        double pre = d * alpha * std::exp(-d * d / twoSigmaSquared) / (sigmaThreeHalvesSqrtTwoPi);
        return (-pre) * sep;
    }
    double energy(const V3&, double d) const
    {
        if (kind == FORCE_HARMONIC) { // harmonicRepulsion.cpp:3-16
            double ans = 0;
            if (d < sigma) ans = 0.5 * k * (sigma - d) * (sigma - d);
            return ans;
        }
        const double sqrtTwoPi = 2.50662827463100050241576528481104525300698674061;
        double twoSigmaSquared = 2.0 * gsigma * gsigma;
        return alpha * std::exp(-d * d / twoSigmaSquared) / (sqrtTwoPi * gsigma);
    }
};

// mpiModel::determineIndexBounds (mpiModel.cpp:20-31)
inline void indexBounds(int nTotal, int rank, int nranks, int& lo, int& hi)
{
    int per = (int)std::ceil((double)nTotal / (double)nranks);
    lo = rank * per;
    hi = (rank + 1) * per;
    if (rank == nranks - 1) hi = nTotal;
    if (lo > nTotal) lo = nTotal; // guard for the reference's latent (R-1)*per >= N failure
    if (hi > nTotal) hi = nTotal;
}

struct Sim {
    Mesh mesh;
    std::vector<char> saddle;
    bool submeshing = false;
    double maximumDistance = 0;   // triangulatedMeshSpace::maximumDistance (triangulatedMeshSpace.h:122)
    bool useCellList = true;
    CellList cl;
    int N = 0;
    std::vector<int> face;
    std::vector<double> bary;     // 3N
    std::vector<V3> vel, frc, eucl;
    std::vector<std::vector<int>> nbr;
    std::vector<std::vector<double>> nbrDist;
    std::vector<std::vector<V3>> nbrStart, nbrEnd;
    std::vector<int> walkFlags;
    bool transportForce = false, transportVelocity = true;
    int boundaryMode = 0; // openMeshSpace variants: 0 closed, 1 absorbing, 2 tangential (walker.hpp BoundaryMode)
    bool strictTrig = false;
    int nThreads = 1;
    GeoStats stats;
    long flagCounts[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // vertex, nohit, itercap, nan, border, disconnected, tie, -
    long crossings = 0;

    void setMesh(int nV, const double* xyz, int nF, const int* corners)
    {
        mesh.set(nV, xyz, nF, corners);
        saddle = saddleFlags(mesh);
        cl.setDomain(mesh.bbmin, mesh.bbmax);
    }
    void resize(int n)
    {
        N = n;
        face.assign(n, 0);
        bary.assign(3 * (size_t)n, 1.0 / 3);
        vel.assign(n, V3{0, 0, 0});
        frc.assign(n, V3{0, 0, 0});
        nbr.assign(n, {});
        nbrDist.assign(n, {});
        nbrStart.assign(n, {});
        nbrEnd.assign(n, {});
        walkFlags.assign(n, 0);
    }

    template <class F> void parallelFor(F fn)
    {
        int T = std::max(1, nThreads);
        if (T == 1) {
            fn(0, N, 0);
            return;
        }
        std::vector<std::thread> th;
        for (int r = 0; r < T; ++r) {
            int lo, hi;
            indexBounds(N, r, T, lo, hi);
            th.emplace_back([=] { fn(lo, hi, r); });
        }
        for (auto& t : th) t.join();
    }

    // triangulatedMeshSpace::meshPositionToEuclideanLocation (triangulatedMeshSpace.cpp:82-106)
    void fillEuclidean()
    {
        eucl.resize(N);
        for (int i = 0; i < N; ++i) eucl[i] = mesh.point(face[i], &bary[3 * i]);
    }

    // triangulatedMeshSpace::distance (:212-238) and distanceWithSubmeshing (:155-205)
    void distance(int sf, const double* sb, const std::vector<GeoTarget>& tg, double threshold, std::vector<GeoResult>& out,
                  GeoStats* st, std::vector<int>* patchOut = nullptr) const
    {
        if (submeshing) {
            double thr = maximumDistance;
            if (threshold < maximumDistance) thr = threshold; // :167-169
            std::vector<int> tf(tg.size());
            for (size_t i = 0; i < tg.size(); ++i) tf[i] = tg[i].face;
            std::vector<int> pf = patchFaces(mesh, sf, sb, tf, thr);
            if (patchOut) *patchOut = pf;
            PatchGeodesic pg(mesh, saddle, &pf);
            pg.solve(sf, sb, tg, out, st);
            for (auto& r : out)
                if (r.dist < 0) { // :198-203 disconnected-patch sentinel
                    r.dist = 2.0 * maximumDistance;
                    r.ts = V3{0, 0, 1};
                    r.te = V3{0, 0, 1};
                    r.tie = -1;
                }
        } else {
            PatchGeodesic pg(mesh, saddle, nullptr);
            pg.solve(sf, sb, tg, out, st);
        }
    }

    // simpleModel::findNeighbors (simpleModel.cpp:68-112); [lo,hi) = the rank's particles (mpiModel.cpp:75-122)
    void findNeighbors(double range)
    {
        fillEuclidean();
        if (useCellList) {
            cl.setRange(range);
            cl.build(eucl);
        }
        std::vector<GeoStats> tst(std::max(1, nThreads));
        std::vector<long> tdis(std::max(1, nThreads), 0), ttie(std::max(1, nThreads), 0);
        parallelFor([&](int lo, int hi, int tid) {
            std::vector<GeoTarget> tg;
            std::vector<GeoResult> res;
            for (int i = lo; i < hi; ++i) {
                double R;
                if (useCellList)
                    R = cl.candidates(i, nbr[i]);
                else { // baseNeighborStructure.cpp:17-36: everybody else, VERYLARGEDOUBLE
                    nbr[i].clear();
                    for (int j = 0; j < N; ++j)
                        if (j != i) nbr[i].push_back(j);
                    R = 1e20;
                }
                int K = (int)nbr[i].size();
                tg.resize(K);
                for (int jj = 0; jj < K; ++jj) {
                    int j = nbr[i][jj];
                    tg[jj].face = face[j];
                    tg[jj].b[0] = bary[3 * j], tg[jj].b[1] = bary[3 * j + 1], tg[jj].b[2] = bary[3 * j + 2];
                }
                nbrDist[i].resize(K);
                nbrStart[i].resize(K);
                nbrEnd[i].resize(K);
                if (K == 0) continue; // q9: the reference still calls distance(); nothing observable changes
                distance(face[i], &bary[3 * i], tg, R, res, &tst[tid]);
                for (int jj = 0; jj < K; ++jj) {
                    nbrDist[i][jj] = res[jj].dist;
                    nbrStart[i][jj] = res[jj].ts;
                    nbrEnd[i][jj] = res[jj].te;
                    if (res[jj].tie < 0) tdis[tid]++;
                    if (res[jj].tie > 0) ttie[tid]++;
                }
            }
        });
        for (size_t t = 0; t < tst.size(); ++t) {
            stats.add(tst[t]);
            flagCounts[5] += tdis[t];
            flagCounts[6] += ttie[t];
        }
    }

    // force::computeForces (baseForce.cpp:12-28)
    void computeForces(const PairForce& pf, bool zero)
    {
        findNeighbors(pf.range);
        for (int i = 0; i < N; ++i) {
            if (zero) frc[i] = V3{0, 0, 0};
            for (size_t jj = 0; jj < nbr[i].size(); ++jj) frc[i] = frc[i] + pf.force(nbrStart[i][jj], nbrDist[i][jj]);
        }
    }
    double computeEnergy(const PairForce& pf) // baseForce.cpp:33-44
    {
        findNeighbors(pf.range);
        double e = 0;
        for (int i = 0; i < N; ++i)
            for (size_t jj = 0; jj < nbr[i].size(); ++jj) e += pf.energy(nbrStart[i][jj], nbrDist[i][jj]);
        return e;
    }

    // simulation::computeMonodisperseStress (simulation.cpp:104-173), one force computer.  NB v (x) v sits inside the
    // neighbour loop there, so it is accumulated once per neighbour; kept as is.
    void computeStress(const PairForce& pf, double area, double out[9])
    {
        findNeighbors(pf.range);
        double fdr[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, vv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < N; ++i)
            for (size_t jj = 0; jj < nbr[i].size(); ++jj) {
                const V3& sep = nbrStart[i][jj];
                V3 f = pf.force(sep, nbrDist[i][jj]);
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) fdr[3 * a + b] += f[a] * sep[b], vv[3 * a + b] += vel[i][a] * vel[i][b];
            }
        double density = N / area;
        for (int q = 0; q < 9; ++q) out[q] = density * 1.0 * vv[q] / (2 * N) + fdr[q] / (2 * 2 * area * N);
    }
    double temperature() const // noseHooverNVT::getTemperatureFromKE (noseHooverNVT.cpp:141-150)
    {
        double v2 = 0;
        for (int i = 0; i < N; ++i) v2 = v2 + dot(vel[i], vel[i]);
        return v2 / (2 * N);
    }

    // simpleModel::moveParticles (simpleModel.cpp:44-66): transports = [force?][velocity?]
    void moveParticles(std::vector<V3>& disp)
    {
        std::vector<long> tc(std::max(1, nThreads), 0);
        parallelFor([&](int lo, int hi, int tid) {
            for (int i = lo; i < hi; ++i) {
                V3 T[2];
                int nT = 0;
                if (transportForce) T[nT++] = frc[i];
                if (transportVelocity) T[nT++] = vel[i];
                int cr = 0;
                walkFlags[i] = transport(mesh, face[i], &bary[3 * i], disp[i], T, nT, strictTrig, &cr, boundaryMode);
                tc[tid] += cr;
                nT = 0;
                if (transportForce) frc[i] = T[nT++];
                if (transportVelocity) vel[i] = T[nT++];
            }
        });
        for (long c : tc) crossings += c;
        for (int i = 0; i < N; ++i)
            for (int b = 0; b < 5; ++b)
                if (walkFlags[i] & (1 << b)) flagCounts[b]++;
    }

    // ---------------- updaters ----------------
    std::vector<V3> disp;

    // velocityVerletNVE.cpp:14-29
    void nveFirstHalf(double dt)
    {
        disp.resize(N);
        for (int i = 0; i < N; ++i) {
            disp[i] = dt * vel[i] + (0.5 * dt * dt) * frc[i];
            vel[i] = vel[i] + (0.5 * dt) * frc[i];
        }
    }
    void nveSecondHalf(double dt, const PairForce& pf)
    {
        moveParticles(disp);
        computeForces(pf, true);
        for (int i = 0; i < N; ++i) vel[i] = vel[i] + (0.5 * dt) * frc[i];
    }
    void stepNVE(double dt, const PairForce& pf)
    {
        transportForce = false;
        transportVelocity = true;
        nveFirstHalf(dt);
        nveSecondHalf(dt, pf);
    }
    // gradientDescent.cpp:6-18
    void stepGD(double dt, const PairForce& pf)
    {
        // gradientDescent never asks for vector transport (simpleModel.h defaults: both flags false)
        transportForce = false;
        transportVelocity = false;
        disp.resize(N);
        computeForces(pf, true);
        for (int i = 0; i < N; ++i) disp[i] = dt * frc[i];
        moveParticles(disp);
    }

    // noseHooverNVT.cpp
    struct NoseHoover {
        double dt, dt2, dt4, dt8, T, tau;
        int M;
        std::vector<double> bx, by, bz, bw; // double4 bathVariables {x,y,z,w}
        double KE = 0, scale = 1;
        void init(double dt_, double T_, double tau_, int M_, int Ndof) // :3-36
        {
            dt = dt_, dt2 = 0.5 * dt_, dt4 = 0.25 * dt_, dt8 = 0.125 * dt_;
            T = T_, tau = tau_, M = M_;
            bx.assign(M + 1, 0), by.assign(M + 1, 0), bz.assign(M + 1, 0), bw.assign(M + 1, 0);
            bw[0] = 2.0 * (Ndof - 1) * T * tau * tau;
            for (int i = 1; i <= M; ++i) bw[i] = T * tau * tau;
            KE = bw[0];
            scale = 1.0;
        }
        void propagateChain() // :65-110
        {
            double ef = 0;
            for (int ii = M - 1; ii > 0; --ii) {
                bz[ii] = (bw[ii - 1] * by[ii - 1] * by[ii - 1] - T) / bw[ii];
                ef = std::exp(-dt8 * by[ii + 1]);
                by[ii] *= ef;
                by[ii] += bz[ii] * dt4;
                by[ii] *= ef;
            }
            bz[0] = (2.0 * KE / bw[0] - 1.0);
            ef = std::exp(-dt8 * by[1]);
            by[0] *= ef;
            by[0] += bz[0] * dt4;
            by[0] *= ef;
            for (int ii = 0; ii < M; ++ii) bx[ii] += dt2 * by[ii];
            scale = std::exp(-dt2 * by[0]);
            KE = scale * scale * KE;
            bz[0] = (2.0 * KE / bw[0] - 1.0);
            ef = std::exp(-dt8 * by[1]);
            by[0] *= ef;
            by[0] += bz[0] * dt4;
            by[0] *= ef;
            for (int ii = 1; ii < M; ++ii) {
                bz[ii] = (bw[ii - 1] * by[ii - 1] * by[ii - 1] - T) / bw[ii];
                ef = std::exp(-dt8 * by[ii + 1]);
                by[ii] *= ef;
                by[ii] += bz[ii] * dt4;
                by[ii] *= ef;
            }
        }
    } nh;
    void stepNVT(const PairForce& pf) // :42-59, :116-139 (unit masses)
    {
        transportForce = false;
        transportVelocity = true;
        disp.resize(N);
        nh.propagateChain();
        for (int i = 0; i < N; ++i) vel[i] = nh.scale * vel[i];
        nh.KE = 0.0;
        for (int i = 0; i < N; ++i) disp[i] = nh.dt2 * vel[i];
        moveParticles(disp);
        computeForces(pf, true);
        for (int i = 0; i < N; ++i) {
            vel[i] = vel[i] + (nh.dt / 1.0) * frc[i];
            disp[i] = nh.dt2 * vel[i];
            double vv = dot(vel[i], vel[i]);
            nh.KE += 0.5 * (1.0) * vv;
        }
        moveParticles(disp);
        nh.propagateChain();
        for (int i = 0; i < N; ++i) vel[i] = nh.scale * vel[i];
    }

    // fireMinimization.{h,cpp}
    struct Fire {
        double dt = 0.001, alpha = 0.99;
        int maximumIterations = 1000, nMin = 4, nSinceNegativePower = 0, iterations = 0;
        double alphaStart = 0.99, deltaTMax = 0.1, deltaTInc = 1.1, deltaTMin = 1e-5, deltaTDec = 0.95, alphaDec = 0.9,
               forceCutoff = 1e-12, alphaMin = 0.0;
        double forceMax = 0, power = 0, forceNorm = 0, velocityNorm = 0;
    } fire;
    double maxForce() const // baseUpdater.cpp:39-54
    {
        double mx = 0;
        for (int i = 0; i < N; ++i) {
            double s = sqlen(frc[i]);
            if (s > mx) mx = s;
        }
        return std::sqrt(mx);
    }
    void fireStep() // fireMinimization.cpp:36-72
    {
        double fn = 0, vn = 0, pw = 0;
        for (int i = 0; i < N; ++i) fn += dot(frc[i], frc[i]);
        for (int i = 0; i < N; ++i) vn += dot(vel[i], vel[i]);
        for (int i = 0; i < N; ++i) pw += dot(frc[i], vel[i]);
        fire.forceNorm = fn, fire.velocityNorm = vn, fire.power = pw;
        double scaling = 0.0;
        if (fn > 0) scaling = std::sqrt(vn / fn);
        for (int i = 0; i < N; ++i) vel[i] = (1 - fire.alpha) * vel[i] + (fire.alpha * scaling) * frc[i];
        if (pw > 0) {
            if (fire.nSinceNegativePower > fire.nMin) {
                fire.dt = std::min(fire.dt * fire.deltaTInc, fire.deltaTMax);
                fire.alpha = fire.alpha * fire.alphaDec;
                fire.alpha = std::max(fire.alpha, fire.alphaMin);
            }
            fire.nSinceNegativePower += 1;
        } else {
            fire.nSinceNegativePower = 0;
            fire.dt = fire.dt * fire.deltaTDec;
            fire.dt = std::max(fire.dt, fire.deltaTMin);
            fire.alpha = fire.alphaStart;
            for (int i = 0; i < N; ++i) vel[i] = V3{0, 0, 0};
        }
    }
    void minimizeByFire(const PairForce& pf) // fireMinimization.cpp:3-21
    {
        transportForce = true;
        transportVelocity = true;
        computeForces(pf, true);
        fire.forceMax = maxForce();
        fire.iterations = 0;
        while (fire.iterations < fire.maximumIterations && fire.forceMax > fire.forceCutoff) {
            fire.iterations += 1;
            nveFirstHalf(fire.dt);
            nveSecondHalf(fire.dt, pf);
            fireStep();
            fire.forceMax = maxForce();
        }
    }
};

} // namespace orc
