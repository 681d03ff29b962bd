// CPU-only check of host/css_database.hpp: writes a simpleModelDatabase and a valueVectorDatabase that
// tests/test_trajectory_io.py reads back with curvedspacesim_b200/trajectory.py, and reads one that Python wrote.
// No CUDA call is made: the model below is a plain simpleModel on a one-triangle "space".
#include "../../host/css_database.hpp"

struct flatSpace : public baseSpace
    {
    virtual void displaceParticle(meshPosition&, vector3&) {}
    virtual void transportParticleAndVectors(meshPosition&, vector3&, vector<vector3>&) {}
    virtual void distance(meshPosition&, vector<meshPosition>&, vector<double>&, vector<vector3>&, vector<vector3>&, double) {}
    virtual void meshPositionToEuclideanLocation(vector<meshPosition>& p1, vector<double3>& result)
        {
        result.resize(p1.size()); // triangle (0,0,0) (1,0,0) (0,1,0) shifted by the face index along z
        for (size_t i = 0; i < p1.size(); ++i) result[i] = double3{p1[i].x[1], p1[i].x[2], (double)p1[i].faceIndex};
        }
    virtual void meshPositionToEuclideanLocation(vector<meshPosition>&, vector<meshPosition>&) {}
    virtual double getArea() { return 0.5; }
    virtual void randomPosition(meshPosition&, noiseSource&) {}
    virtual void randomVectorAtPosition(meshPosition&, vector3&, noiseSource&) {}
    };
struct plainModel : public simpleModel
    {
    using simpleModel::simpleModel;
    virtual void findNeighbors(double) {}
    };

int main(int argc, char** argv)
    {
    if (argc < 3) return 2;
    string dir = argv[1], mode = argv[2];
    const int N = 5;
    auto m = make_shared<plainModel>(N);
    m->setSpace(make_shared<flatSpace>());
    if (mode == "write")
        {
        simpleModelDatabase db(N, dir + "/traj.cssdb", fileMode::replace);
        for (int rec = 0; rec < 3; ++rec)
            {
            for (int i = 0; i < N; ++i)
                {
                double a = 0.1 * (i + 1) + 0.01 * rec, b = 0.05 * (i + 1);
                m->positions[i] = meshPosition(point3(1 - a - b, a, b), 10 * rec + i);
                m->velocities[i] = vector3(rec + 0.5, i, -1.25);
                m->forces[i] = vector3(-rec, 2.0 * i, 1e-3);
                m->types[i] = i % 2;
                }
            db.writeState(m, 0.25 * rec);
            }
        if (db.currentNumberOfRecords() != 3) return 3;
        valueVectorDatabase vv(dir + "/series.cssdb", 4, fileMode::replace);
        vector<double> row{1, 2, 3, 4};
        vv.writeState(7.5, row);
        row[0] = -1;
        vv.writeState(8.5, row);
        return vv.currentNumberOfRecords() == 2 ? 0 : 4;
        }
    // read what Python wrote: record 1 of 2
    simpleModelDatabase db(N, dir + "/pytraj.cssdb", fileMode::readonly);
    if (db.currentNumberOfRecords() != 2) return 5;
    db.readState(m, 1);
    bool ok = db.lastTime == 2.5;
    for (int i = 0; i < N; ++i)
        ok = ok && m->positions[i].faceIndex == 100 + i && m->positions[i].x[0] == 0.5 && m->velocities[i][1] == 3.0 * i && m->forces[i][2] == -2.0
             && m->types[i] == 7;
    db.readState(m, -1);
    ok = ok && db.lastTime == 2.5;
    return ok ? 0 : 6;
    }
