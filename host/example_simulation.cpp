// example_simulation.cpp — the wiring of the reference's curvedSpaceSimulation.cpp:72-147 / curvedSpaceNVTSim.cpp:71-122
// on top of host/css_host.hpp (GPU space + GPU model, unchanged force / updater / simulation plumbing).
//
//   example_simulation <mesh.off> <N> <iterations> <programBranch> <fused> <dump.bin> [dt] [temperature]
//     programBranch: 0 FIRE, 1 gradient descent, 2 velocity-Verlet NVE, 3 Nose-Hoover NVT
//     fused: 0 = host-driven updaters (reference control flow, one ABI call per moveParticles/computeForces)
//            1 = gpuSimulation (device-resident fused step)
//   CSS_EXAMPLE_DB=<dir> in the environment: the initial and final states are also written as two records of a
//   simpleModelDatabase (host/css_database.hpp), as the reference's mains do with their HDF5 trajectory file.
//   CSS_EXAMPLE_USERFORCE=1: the pair potential is a user-written force subclass (host callback path) instead of the stock one.
//   CSS_EXAMPLE_R3=<file>: the final R^3 coordinates are written to <file> and imported again (setMeshPositionsFromR3File).
// The dump holds N, then the initial (face, bary, velocity) and the final (face, bary, velocity, force) arrays in raw
// little-endian form; tests/test_gpu_parity.py replays the same initial state through the ctypes binding and the oracle.
#include "css_database.hpp"

#include <chrono>
#include <cstdlib>

// A potential written by a USER of the plugin surface (src/forces/baseForce.h:30-57): only pairwiseForce / pairwiseEnergy, no
// device functor.  With a gpuModel the base class downloads the neighbour lists and calls these virtuals on the host.
// Selected with CSS_EXAMPLE_USERFORCE=1; it states the harmonic law, so the run must reproduce the stock potential's.
class userWrittenRepulsion : public force
    {
public:
    userWrittenRepulsion(double stiffness, double range) : k(stiffness), sigma(range) { maximumInteractionRange = range; }
    virtual string reportSelfName() { return "user-written repulsion"; }
    virtual double pairwiseEnergy(vector3 separation, double distance)
        {
        (void)separation;
        return distance < sigma ? 0.5 * k * (sigma - distance) * (sigma - distance) : 0.0;
        }
    virtual vector3 pairwiseForce(vector3 separation, double distance)
        {
        return distance <= sigma ? (-k * (sigma - distance)) * separation : vector3(0, 0, 0);
        }
    double k, sigma;
    };

static void dumpState(FILE* f, simpleModel& m, bool withForces)
{
    for (int i = 0; i < m.N; ++i) fwrite(&m.positions[i].faceIndex, sizeof(int), 1, f);
    for (int i = 0; i < m.N; ++i) fwrite(m.positions[i].x.c, sizeof(double), 3, f);
    for (int i = 0; i < m.N; ++i) fwrite(m.velocities[i].c, sizeof(double), 3, f);
    if (withForces)
        for (int i = 0; i < m.N; ++i) fwrite(m.forces[i].c, sizeof(double), 3, f);
}

int main(int argc, char** argv)
{
    if (argc < 7)
        {
        fprintf(stderr, "usage: %s mesh.off N iterations programBranch fused dump.bin [dt] [T]\n", argv[0]);
        return 2;
        }
    string meshName = argv[1];
    int N = atoi(argv[2]), maximumIterations = atoi(argv[3]), programBranch = atoi(argv[4]);
    bool fused = atoi(argv[5]) != 0;
    double dt = argc > 7 ? atof(argv[7]) : 0.01, temperature = argc > 8 ? atof(argv[8]) : 0.2, areaFraction = 0.9;
    try
        {
        shared_ptr<closedMeshSpace> meshSpace = make_shared<closedMeshSpace>();
        meshSpace->loadMeshFromFile(meshName, true);
        double area = totalArea(*meshSpace);
        double maximumInteractionRange = 2 * sqrt(areaFraction * area / (N * M_PI));
        meshSpace->useSubmeshingRoutines(true, maximumInteractionRange);

        shared_ptr<gpuModel> configuration = make_shared<gpuModel>(N);
        configuration->setSpace(meshSpace);
        shared_ptr<cellListNeighborStructure> cellList
            = make_shared<cellListNeighborStructure>(meshSpace->minVertexPosition, meshSpace->maxVertexPosition, maximumInteractionRange);
        configuration->setNeighborStructure(cellList);

        noiseSource noise(true);
        configuration->setRandomParticlePositions(noise);
        configuration->setMaxwellBoltzmannVelocities(noise, temperature);

        shared_ptr<force> pairwiseForce;
        if (getenv("CSS_EXAMPLE_USERFORCE")) pairwiseForce = make_shared<userWrittenRepulsion>(1.0, maximumInteractionRange);
        else pairwiseForce = make_shared<harmonicRepulsion>(1.0, maximumInteractionRange);
        pairwiseForce->setModel(configuration);

        shared_ptr<simulation> simulator = fused ? make_shared<gpuSimulation>() : make_shared<simulation>();
        simulator->setConfiguration(configuration);
        simulator->addForce(pairwiseForce);

        shared_ptr<updater> eom;
        if (programBranch >= 3) eom = make_shared<noseHooverNVT>(dt, temperature, 1.0, 2);
        else if (programBranch >= 2) eom = make_shared<velocityVerletNVE>(dt);
        else if (programBranch >= 1) eom = make_shared<gradientDescent>(dt);
        else
            {
            auto fire = make_shared<fireMinimization>();
            fire->setFIREParameters(maximumIterations, dt, 0.99, 0.1, 1e-5, 1.1, 0.95, 0.9, 4, 1e-12, 0.0);
            fire->setDeltaT(dt);
            eom = fire;
            maximumIterations = 1; // one performTimestep is a whole minimisation
            }
        simulator->addUpdater(eom, configuration);

        FILE* f = fopen(argv[6], "wb");
        if (!f) ERRORERROR("cannot open the dump file");
        fwrite(&N, sizeof(int), 1, f);
        dumpState(f, *configuration, false);

        if (const char* dbDir = getenv("CSS_EXAMPLE_DB"))
            {
            simpleModelDatabase db(N, dbDir, fileMode::replace);
            db.writeState(configuration, 0.0);
            }
        auto t0 = std::chrono::steady_clock::now();
        for (int ii = 0; ii < maximumIterations; ++ii) simulator->performTimestep();
        if (fused) std::static_pointer_cast<gpuSimulation>(simulator)->syncHost();
        double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

        dumpState(f, *configuration, true);
        fclose(f);
        vector<double> stress;
        simulator->computeMonodisperseStress(stress); // curvedSpaceNVTSim.cpp:113-118 prints the same observable
        printf("stress trace %.17g temperature %.17g\n", stress[0] + stress[4] + stress[8], configuration->temperatureOnDevice());
        if (const char* dbDir = getenv("CSS_EXAMPLE_DB"))
            {
            simpleModelDatabase db(N, dbDir, fileMode::readwrite);
            db.writeState(configuration, simulator->Time);
            }
        if (const char* r3file = getenv("CSS_EXAMPLE_R3"))
            { // simpleModel::setMeshPositionsFromR3File: write the final R^3 coordinates as "x,y,z" lines, import them into a second
              // model on the same space and compare with the mesh positions they came from
            configuration->fillEuclideanLocations();
            FILE* r3 = fopen(r3file, "w");
            if (!r3) ERRORERROR("cannot open the R3 file");
            for (int i = 0; i < N; ++i)
                fprintf(r3, "%.17g,%.17g,%.17g\n", configuration->euclideanLocations[i].x, configuration->euclideanLocations[i].y,
                        configuration->euclideanLocations[i].z);
            fclose(r3);
            shared_ptr<gpuModel> imported = make_shared<gpuModel>(1);
            imported->setSpace(meshSpace);
            imported->setMeshPositionsFromR3File(r3file);
            int sameFace = 0;
            double maxDiff = 0;
            for (int i = 0; i < N && imported->N == N; ++i)
                if (imported->positions[i].faceIndex == configuration->positions[i].faceIndex)
                    {
                    sameFace++;
                    for (int k = 0; k < 3; ++k) maxDiff = std::max(maxDiff, std::fabs(imported->positions[i].x[k] - configuration->positions[i].x[k]));
                    }
            printf("R3 import: %d positions, %d on the same face, max weight difference %.3e\n", imported->N, sameFace, maxDiff);
            }
        printf("%d particles, %d timesteps in %.4f s (%s): %.3e particle-timesteps/s; fN %g fM %g\n", N, maximumIterations, secs,
               fused ? "fused device step" : "host-driven updaters", N * (double)maximumIterations / secs, eom->getForceNorm(), eom->getMaxForce());
        }
    catch (const std::exception&)
        {
        fprintf(stderr, "example_simulation: aborted by an error\n");
        return 1;
        }
    return 0;
}
