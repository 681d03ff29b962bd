// dump_reference.cpp — OFF-BOX tool: produces TRUE-REFERENCE golden vectors with the unmodified curvedSpaceSim + CGAL.
//
// This file is NOT built by this repository (CGAL 5.6, Boost, HDF5 and MPI are absent from its image; see DESIGN.md
// section 2).  On a machine that has the reference built, drop it next to the reference's other mains, add
//     add_executable(dump_reference.out dump_reference.cpp)  + the same target_link_libraries as curvedSpaceSimulation.out
// to the reference's CMakeLists.txt and run
//     ./dump_reference.out <mesh.off> <state_in.bin> <out.bin> <range> <nveSteps> <dt>
// state_in.bin is written by tests/golden/make_reference_inputs.py (N, face[N] int32, bary[N][3] f64, vel[N][3] f64);
// corner order in both files is the reference's own (getVertexIndicesFromFace), so no permutation is applied anywhere.
// out.bin holds, little-endian: N, the neighbour CSR (offsets[N+1] int32, idx, dist f64, startTangent[.][3] f64,
// endTangent[.][3] f64) from simpleModel::findNeighbors-equivalent calls of triangulatedMeshSpace::distance, the forces
// (harmonicRepulsion k = 1, sigma = range), the state (face, bary, vel, force) after nveSteps velocity-Verlet steps, and an
// R3PositionsToMeshPositions block (points, faces, clamped weights) for css_locate.
// tests/test_reference_dumps.py compares the oracle AND the CUDA path against every such file found under
// tests/golden/reference_dumps/ (bars: bit-exact lists/faces, 1e-9 distances/tangents, 1e-8 forces, 1e-6 trajectories).
#include "cellListNeighborStructure.h"
#include "harmonicRepulsion.h"
#include "simpleModel.h"
#include "simulation.h"
#include "std_include.h"
#include "triangulatedMeshSpace.h"
#include "velocityVerletNVE.h"

#include <cstdio>
#include <cstdlib>

static void put(FILE* f, const void* p, size_t n)
{
    if (fwrite(p, 1, n, f) != n) { perror("write"); exit(1); }
}
static void get(FILE* f, void* p, size_t n)
{
    if (fread(p, 1, n, f) != n) { perror("read"); exit(1); }
}
static void putState(FILE* f, simpleModel& m)
{
    for (int i = 0; i < m.N; ++i) put(f, &m.positions[i].faceIndex, 4);
    for (int i = 0; i < m.N; ++i) for (int k = 0; k < 3; ++k) { double v = m.positions[i].x[k]; put(f, &v, 8); }
    for (int i = 0; i < m.N; ++i) for (int k = 0; k < 3; ++k) { double v = m.velocities[i][k]; put(f, &v, 8); }
    for (int i = 0; i < m.N; ++i) for (int k = 0; k < 3; ++k) { double v = m.forces[i][k]; put(f, &v, 8); }
}

int main(int argc, char* argv[])
{
    if (argc < 7) { fprintf(stderr, "usage: %s mesh.off state_in.bin out.bin range nveSteps dt\n", argv[0]); return 2; }
    double range = atof(argv[4]);
    int steps = atoi(argv[5]);
    double dt = atof(argv[6]);

    shared_ptr<triangulatedMeshSpace> meshSpace = make_shared<triangulatedMeshSpace>();
    meshSpace->loadMeshFromFile(argv[1], false);
    meshSpace->useSubmeshingRoutines(true, range);

    FILE* in = fopen(argv[2], "rb");
    if (!in) { perror(argv[2]); return 1; }
    int N = 0;
    get(in, &N, 4);
    std::vector<int> face(N);
    std::vector<double> bary(3 * N), vel(3 * N);
    get(in, face.data(), 4 * (size_t)N);
    get(in, bary.data(), 24 * (size_t)N);
    get(in, vel.data(), 24 * (size_t)N);
    fclose(in);

    shared_ptr<simpleModel> configuration = make_shared<simpleModel>(N);
    configuration->setSpace(meshSpace);
    shared_ptr<cellListNeighborStructure> cellList
        = make_shared<cellListNeighborStructure>(meshSpace->minVertexPosition, meshSpace->maxVertexPosition, range);
    configuration->setNeighborStructure(cellList);
    vector<meshPosition> pos(N);
    for (int i = 0; i < N; ++i)
        {
        pos[i] = meshPosition(point3(bary[3 * i], bary[3 * i + 1], bary[3 * i + 2]), face[i]);
        configuration->velocities[i] = vector3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        }
    configuration->setParticlePositions(pos);

    // neighbour lists with BOTH tangents: the reference's findNeighbors drops the end tangents, so the space is queried
    // here exactly the way simpleModel.cpp:88-107 does, keeping them
    configuration->findNeighbors(range); // fills neighbors / neighborVectors / neighborDistances
    FILE* out = fopen(argv[3], "wb");
    if (!out) { perror(argv[3]); return 1; }
    put(out, &N, 4);
    int off = 0;
    for (int i = 0; i < N; ++i) { put(out, &off, 4); off += (int)configuration->neighbors[i].size(); }
    put(out, &off, 4);
    for (int i = 0; i < N; ++i) for (int j : configuration->neighbors[i]) put(out, &j, 4);
    for (int i = 0; i < N; ++i) for (double d : configuration->neighborDistances[i]) put(out, &d, 8);
    for (int i = 0; i < N; ++i) for (vector3& t : configuration->neighborVectors[i]) for (int k = 0; k < 3; ++k) { double v = t[k]; put(out, &v, 8); }
    for (int i = 0; i < N; ++i)
        {
        std::vector<meshPosition> targets;
        double maxD2 = 0;
        std::vector<meshPosition> self(1, configuration->positions[i]), selfE;
        meshSpace->meshPositionToEuclideanLocation(self, selfE);
        for (int j : configuration->neighbors[i])
            {
            targets.push_back(configuration->positions[j]);
            std::vector<meshPosition> one(1, configuration->positions[j]), oneE;
            meshSpace->meshPositionToEuclideanLocation(one, oneE);
            maxD2 = std::max(maxD2, CGAL::squared_distance(selfE[0].x, oneE[0].x));
            }
        std::vector<double> d;
        std::vector<vector3> ts, te;
        if (!targets.empty()) meshSpace->distance(configuration->positions[i], targets, d, ts, te, sqrt(maxD2));
        for (vector3& t : te) for (int k = 0; k < 3; ++k) { double v = t[k]; put(out, &v, 8); }
        }

    shared_ptr<harmonicRepulsion> pairwiseForce = make_shared<harmonicRepulsion>(1.0, range);
    pairwiseForce->setModel(configuration);
    shared_ptr<simulation> simulator = make_shared<simulation>();
    simulator->setConfiguration(configuration);
    simulator->addForce(pairwiseForce);
    simulator->computeForces();
    for (int i = 0; i < N; ++i) for (int k = 0; k < 3; ++k) { double v = configuration->forces[i][k]; put(out, &v, 8); }

    shared_ptr<velocityVerletNVE> nve = make_shared<velocityVerletNVE>(dt);
    simulator->addUpdater(nve, configuration);
    for (int s = 0; s < steps; ++s) simulator->performTimestep();
    putState(out, *configuration);

    // optional trailing block (css_locate): the R^3 coordinates of the final state, nudged off the surface by 1e-9 along x,
    // go through simpleModel::R3PositionsToMeshPositions (simpleModel.cpp:136-154); written as M, xyz[M][3], face[M], bary[M][3]
    {
    std::vector<meshPosition> finalE;
    meshSpace->meshPositionToEuclideanLocation(configuration->positions, finalE);
    std::vector<point3> r3;
    for (int i = 0; i < N; ++i) r3.push_back(point3(finalE[i].x[0] + 1e-9, finalE[i].x[1], finalE[i].x[2]));
    std::vector<meshPosition> located;
    configuration->R3PositionsToMeshPositions(meshSpace->surface, r3, located);
    int M = (int)located.size();
    put(out, &M, 4);
    for (int i = 0; i < M; ++i) for (int k = 0; k < 3; ++k) { double v = r3[i][k]; put(out, &v, 8); }
    for (int i = 0; i < M; ++i) put(out, &located[i].faceIndex, 4);
    for (int i = 0; i < M; ++i) for (int k = 0; k < 3; ++k) { double v = located[i].x[k]; put(out, &v, 8); }
    }
    fclose(out);
    printf("dumped %d particles, %d neighbour pairs, %d NVE steps\n", N, off, steps);
    return 0;
}
