// css_host.hpp — header-only C++17 host layer above the C ABI (include/css_api.h).
//
// It mirrors, name for name, the plugin surface of curvedSpaceSim that sits on the per-timestep path, so the
// reference's mains keep their wiring (shared_ptr's of space / model / force / updater / simulation) and only
// swap their includes:
//
//   point3, vector3, double3, double4, meshPosition      inc/cgalIncludesAndTypedefs.h:9-16, inc/dataTypes.h, inc/pointDataType.h:10-28
//   THRESHOLD, VERYLARGEDOUBLE, ERRORERROR, UNWRITTENCODE inc/std_include.h:26-27, inc/debuggingHelp.h:12-31
//   noiseSource                                           src/utility/noiseSource.{h,cpp}
//   baseSpace                                             src/models/baseSpace.h:24-53
//   gpuMeshSpace (= triangulatedMeshSpace, closedMeshSpace) src/models/triangulatedMeshSpace.h:28-87, closedMeshSpace.h
//   baseNeighborStructure, cellListNeighborStructure      src/utility/baseNeighborStructure.h, cellListNeighborStructure.h
//   simpleModel, gpuModel                                 src/models/simpleModel.h:36-98
//   force, harmonicRepulsion, gaussianRepulsion           src/forces/*.h
//   updater, velocityVerletNVE, gradientDescent, noseHooverNVT, fireMinimization   src/updaters/*.h
//   basicSimulation, simulation, gpuSimulation            src/simulation/basicSimulation.h, simulation.h
//
// All geometry runs on the GPU behind the C ABI: gpuMeshSpace forwards the per-call baseSpace interface
// (distance / displaceParticle / transportParticleAndVectors / meshPositionToEuclideanLocation) and gpuModel the
// batched model interface (findNeighbors / moveParticles) to libcurvedspacesim_b200.so.  Host std::vector's stay
// the authoritative public arrays at every API boundary, exactly as in the reference; gpuSimulation adds a
// device-resident fused step for the stock NVE / NVT / FIRE / GD updaters.  A non-zero ABI status becomes the
// reference's error convention: message on stderr + throw std::exception().  There is no CPU implementation here.
#pragma once
#include "../include/css_api.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <exception>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>

using std::function;
using std::make_shared;
using std::shared_ptr;
using std::string;
using std::vector;
using std::weak_ptr;

#define THRESHOLD 1e-14
#define VERYLARGEDOUBLE 1e20
#define ERRORERROR(msg)                                                                        \
    do {                                                                                       \
        std::cerr << "\nError in file " << __FILE__ << " on line " << __LINE__ << ": " << msg << std::endl; \
        throw std::exception();                                                                \
    } while (0)
#define UNWRITTENCODE(msg) ERRORERROR(msg)

// ---------------------------------------------------------------------------------------------- PODs
struct double3 {
    double x, y, z;
};
struct double4 {
    double x, y, z, w;
};
struct int3 {
    int x, y, z;
};

class vector3;
class point3
    {
public:
    point3() : c{0, 0, 0} {}
    point3(double x, double y, double z) : c{x, y, z} {}
    double x() const { return c[0]; }
    double y() const { return c[1]; }
    double z() const { return c[2]; }
    double operator[](int i) const { return c[i]; }
    double c[3];
    };
class vector3
    {
public:
    vector3() : c{0, 0, 0} {}
    vector3(double x, double y, double z) : c{x, y, z} {}
    vector3(const point3& p, const point3& q) : c{q[0] - p[0], q[1] - p[1], q[2] - p[2]} {} // q - p, as CGAL
    double x() const { return c[0]; }
    double y() const { return c[1]; }
    double z() const { return c[2]; }
    double operator[](int i) const { return c[i]; }
    double squared_length() const { return c[0] * c[0] + c[1] * c[1] + c[2] * c[2]; }
    vector3& operator+=(const vector3& o) { c[0] += o.c[0], c[1] += o.c[1], c[2] += o.c[2]; return *this; }
    vector3& operator-=(const vector3& o) { c[0] -= o.c[0], c[1] -= o.c[1], c[2] -= o.c[2]; return *this; }
    vector3& operator*=(double s) { c[0] *= s, c[1] *= s, c[2] *= s; return *this; }
    vector3& operator/=(double s) { c[0] /= s, c[1] /= s, c[2] /= s; return *this; }
    double c[3];
    };
inline vector3 operator+(const vector3& a, const vector3& b) { return vector3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline vector3 operator-(const vector3& a, const vector3& b) { return vector3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline vector3 operator-(const vector3& a) { return vector3(-a[0], -a[1], -a[2]); }
inline vector3 operator*(double s, const vector3& a) { return vector3(s * a[0], s * a[1], s * a[2]); }
inline vector3 operator*(const vector3& a, double s) { return s * a; }
inline vector3 operator/(const vector3& a, double s) { return vector3(a[0] / s, a[1] / s, a[2] / s); }
inline double operator*(const vector3& a, const vector3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; } // dot, as CGAL
inline point3 operator+(const point3& p, const vector3& v) { return point3(p[0] + v[0], p[1] + v[1], p[2] + v[2]); }
inline point3 operator-(const point3& p, const vector3& v) { return point3(p[0] - v[0], p[1] - v[1], p[2] - v[2]); }
inline vector3 operator-(const point3& p, const point3& q) { return vector3(q, p); }
inline vector3 cross_product(const vector3& a, const vector3& b)
    {
    return vector3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
    }
inline double squared_distance(const point3& p, const point3& q) { return vector3(p, q).squared_length(); }

class meshPosition
    {
public:
    meshPosition() : x(0, 0, 0), faceIndex(-1) {}
    meshPosition(point3 _x, int _fIdx) : x(_x), faceIndex(_fIdx) {}
    point3 x; // barycentric weights in a mesh space (or R^3 coordinates in a Euclidean one)
    int faceIndex;
    };

// ---------------------------------------------------------------------------------------------- noise
class noiseSource
    {
public:
    noiseSource(bool rep = false) : Reproducible(rep), gen(13377), genrd(std::random_device{}()) {}
    int getInt(int minimum, int maximum) { return std::uniform_int_distribution<int>(minimum, maximum)(engine()); }
    double getRealUniform(double minimum = 0., double maximum = 1.) { return std::uniform_real_distribution<double>(minimum, maximum)(engine()); }
    double getRealNormal(double mean = 0., double sd = 1.) { return std::normal_distribution<>(mean, sd)(engine()); }
    double3 getRandomBarycentricSet()
        {
        double u = getRealUniform(0, 1), v = getRealUniform(0, 1 - u);
        return double3{u, v, 1 - u - v};
        }
    void setReproducible(bool _rep) { Reproducible = _rep; }
    void setReproducibleSeed(int _seed) { RNGSeed = _seed, gen = std::mt19937(_seed); }
    bool Reproducible;
    int RNGSeed = 13377;
    std::mt19937 gen, genrd;

private:
    std::mt19937& engine() { return Reproducible ? gen : genrd; }
    };

// ---------------------------------------------------------------------------------------------- ABI plumbing
namespace cssHost {
inline void check(css_ctx* ctx, int status, const char* what)
    {
    if (status != CSS_OK) ERRORERROR(what << " failed with status " << status << ": " << css_last_error(ctx));
    }
struct contextHolder // one GPU context shared by the space and the model built on it
    {
    explicit contextHolder(int device)
        {
        if (css_create(&ctx, device) != CSS_OK) ERRORERROR("css_create: no usable CUDA device " << device << " (this library has no CPU path)");
        }
    ~contextHolder() { css_destroy(ctx); }
    contextHolder(const contextHolder&) = delete;
    css_ctx* ctx = nullptr;
    };
} // namespace cssHost

// ---------------------------------------------------------------------------------------------- spaces
class baseSpace
    {
public:
    virtual ~baseSpace() = default;
    virtual void displaceParticle(meshPosition& pos, vector3& displacementVector) = 0;
    virtual void transportParticleAndVectors(meshPosition& pos, vector3& displacementVector, vector<vector3>& transportVectors) = 0;
    virtual void distance(meshPosition& p1, vector<meshPosition>& p2, vector<double>& distances, vector<vector3>& startPathTangent,
                          vector<vector3>& endPathTangent, double distanceThreshold) = 0;
    virtual void meshPositionToEuclideanLocation(vector<meshPosition>& p1, vector<double3>& result) = 0;
    virtual void meshPositionToEuclideanLocation(vector<meshPosition>& p1, vector<meshPosition>& result) = 0;
    //! closest mesh positions of points of R^3 (the mesh half of simpleModel::R3PositionsToMeshPositions, simpleModel.cpp:136-154)
    virtual void R3PositionsToMeshPositions(vector<point3>& r3positions, vector<meshPosition>& simPositions, double clampTolerance)
        {
        (void)r3positions, (void)simPositions, (void)clampTolerance;
        ERRORERROR("this space cannot locate R3 positions");
        }
    virtual double getArea() = 0;
    bool positionsAreEuclidean = true;
    virtual void randomPosition(meshPosition& p, noiseSource& noise) = 0;
    virtual void randomVectorAtPosition(meshPosition& p, vector3& v, noiseSource& noise) = 0;
    };

//! The triangulated-mesh space of the reference with every geometric query executed on the GPU.
class gpuMeshSpace : public baseSpace
    {
public:
    explicit gpuMeshSpace(int device = 0) : holder(make_shared<cssHost::contextHolder>(device)) { positionsAreEuclidean = false; }

    //! OFF reader.  Corner order follows the reference: an OFF face "3 a b c" has corners (c, a, b).
    virtual void loadMeshFromFile(string filename, bool verbose = false)
        {
        std::ifstream in(filename);
        if (!in) ERRORERROR("Invalid input file.");
        string tok;
        in >> tok;
        if (tok != "OFF") ERRORERROR("Invalid input file.");
        auto nextLine = [&](std::istringstream& ss) {
            string line;
            while (std::getline(in, line))
                {
                auto h = line.find('#');
                if (h != string::npos) line.erase(h);
                if (line.find_first_not_of(" \t\r\n") != string::npos)
                    {
                    ss.clear();
                    ss.str(line);
                    return true;
                    }
                }
            return false;
        };
        std::istringstream ss;
        int nV = 0, nF = 0, nE = 0;
        if (!nextLine(ss) || !(ss >> nV >> nF >> nE)) ERRORERROR("Invalid input file.");
        vector<double> xyz(3 * (size_t)nV);
        for (int i = 0; i < nV; ++i)
            if (!nextLine(ss) || !(ss >> xyz[3 * i] >> xyz[3 * i + 1] >> xyz[3 * i + 2])) ERRORERROR("Invalid input file.");
        vector<int32_t> tri(3 * (size_t)nF);
        for (int f = 0; f < nF; ++f)
            {
            int k = 0, a, b, c;
            if (!nextLine(ss) || !(ss >> k >> a >> b >> c) || k != 3) ERRORERROR("Non-triangular mesh. Exiting.");
            tri[3 * f] = c, tri[3 * f + 1] = a, tri[3 * f + 2] = b;
            }
        setMesh(xyz, tri);
        if (verbose) printf("mesh loaded: %d vertices, %d faces, area %g\n", nV, nF, area);
        }

    //! vertices [nV][3] and faces [nF][3] already in the reference's corner order
    void setMesh(const vector<double>& xyz, const vector<int32_t>& corners)
        {
        vertices = xyz;
        faces = corners;
        cssHost::check(ctx(), css_set_mesh(ctx(), (int)(xyz.size() / 3), xyz.data(), (int)(corners.size() / 3), corners.data()), "css_set_mesh");
        double mn[3], mx[3];
        cssHost::check(ctx(), css_mesh_info(ctx(), mn, mx, &area), "css_mesh_info");
        minVertexPosition = double3{mn[0], mn[1], mn[2]};
        maxVertexPosition = double3{mx[0], mx[1], mx[2]};
        }

    virtual void displaceParticle(meshPosition& pos, vector3& displacementVector)
        {
        vector<vector3> none;
        transportParticleAndVectors(pos, displacementVector, none);
        }
    virtual void transportParticleAndVectors(meshPosition& pos, vector3& displacementVector, vector<vector3>& transportVectors)
        {
        int32_t f = pos.faceIndex, flags = 0;
        double b[3] = {pos.x[0], pos.x[1], pos.x[2]}, d[3] = {displacementVector[0], displacementVector[1], displacementVector[2]};
        vector<double> v(3 * transportVectors.size());
        for (size_t i = 0; i < transportVectors.size(); ++i)
            for (int k = 0; k < 3; ++k) v[3 * i + k] = transportVectors[i][k];
        cssHost::check(ctx(), css_transport(ctx(), 1, &f, b, d, (int)transportVectors.size(), v.empty() ? nullptr : v.data(), &flags), "css_transport");
        if ((flags & 16) && boundaryMode == 0) ERRORERROR("a border edge was met in a closed mesh space"); // triangulatedMeshSpace.cpp:522-523
        if (flags & 2) ERRORERROR("no edge intersection found although the target lies outside the face");
        pos = meshPosition(point3(b[0], b[1], b[2]), f);
        displacementVector = vector3(d[0], d[1], d[2]);
        for (size_t i = 0; i < transportVectors.size(); ++i) transportVectors[i] = vector3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        lastWalkFlags = flags;
        }
    virtual void distance(meshPosition& p1, vector<meshPosition>& p2, vector<double>& distances, vector<vector3>& startPathTangent,
                          vector<vector3>& endPathTangent, double distanceThreshold = VERYLARGEDOUBLE)
        {
        int K = (int)p2.size();
        distances.resize(K);
        startPathTangent.resize(K);
        endPathTangent.resize(K);
        if (K == 0) return;
        vector<int32_t> tf(K);
        vector<double> tb(3 * (size_t)K), ts(3 * (size_t)K), te(3 * (size_t)K);
        for (int i = 0; i < K; ++i)
            {
            tf[i] = p2[i].faceIndex;
            for (int k = 0; k < 3; ++k) tb[3 * i + k] = p2[i].x[k];
            }
        double sb[3] = {p1.x[0], p1.x[1], p1.x[2]};
        cssHost::check(ctx(), css_distance(ctx(), p1.faceIndex, sb, K, tf.data(), tb.data(), distanceThreshold, distances.data(), ts.data(), te.data()),
                       "css_distance");
        for (int i = 0; i < K; ++i)
            {
            startPathTangent[i] = vector3(ts[3 * i], ts[3 * i + 1], ts[3 * i + 2]);
            endPathTangent[i] = vector3(te[3 * i], te[3 * i + 1], te[3 * i + 2]);
            }
        }
    //! PMP::locate_with_AABB_tree + simpleModel::clampBarycentricCoordinatesToFace on the device (css_locate)
    virtual void R3PositionsToMeshPositions(vector<point3>& r3positions, vector<meshPosition>& simPositions, double clampTolerance)
        {
        size_t n = r3positions.size();
        vector<double> xyz(3 * n), b(3 * n);
        vector<int32_t> f(n);
        for (size_t i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) xyz[3 * i + k] = r3positions[i][k];
        if (n) cssHost::check(ctx(), css_locate(ctx(), (int)n, xyz.data(), clampTolerance, f.data(), b.data()), "css_locate");
        for (size_t i = 0; i < n; ++i) simPositions.push_back(meshPosition(point3(b[3 * i], b[3 * i + 1], b[3 * i + 2]), f[i]));
        }
    virtual void meshPositionToEuclideanLocation(vector<meshPosition>& p1, vector<double3>& result)
        {
        vector<double> xyz;
        euclid(p1, xyz);
        result.resize(p1.size());
        for (size_t i = 0; i < p1.size(); ++i) result[i] = double3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        }
    virtual void meshPositionToEuclideanLocation(vector<meshPosition>& p1, vector<meshPosition>& result)
        {
        vector<double> xyz;
        euclid(p1, xyz);
        result.resize(p1.size());
        for (size_t i = 0; i < p1.size(); ++i) result[i] = meshPosition(point3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), p1[i].faceIndex);
        }
    virtual double getArea() { return area; }

    //! uniform over faces, not over area, as the reference samples (triangulatedMeshSpace.cpp:108-116)
    virtual void randomPosition(meshPosition& p, noiseSource& noise)
        {
        double u = noise.getRealUniform();
        double v = noise.getRealUniform(0, 1 - u);
        p.x = point3(u, v, 1 - u - v);
        p.faceIndex = noise.getInt(0, (int)(faces.size() / 3) - 1);
        }
    //! in-plane Gaussian vector (triangulatedMeshSpace.cpp:118-138)
    virtual void randomVectorAtPosition(meshPosition& p, vector3& v, noiseSource& noise)
        {
        point3 q[3];
        for (int k = 0; k < 3; ++k)
            {
            const double* s = &vertices[3 * (size_t)faces[3 * (size_t)p.faceIndex + k]];
            q[k] = point3(s[0], s[1], s[2]);
            }
        vector3 n = cross_product(vector3(q[0], q[1]), vector3(q[0], q[2]));
        n /= std::sqrt(n.squared_length());
        vector3 o = (n[0] == 0 && n[1] == 0) ? vector3(n[1] - n[2], n[2] - n[0], n[0] - n[1]) : vector3(n[1], -n[0], 0);
        o /= std::sqrt(o.squared_length());
        vector3 t2 = cross_product(n, o);
        double g1 = noise.getRealNormal(), g2 = noise.getRealNormal();
        v = g1 * o + g2 * t2;
        }

    void useSubmeshingRoutines(bool _useSubMesh, double maxDist = 1.0, bool _danger = false)
        {
        (void)_danger;
        submeshingActivated = _useSubMesh;
        maximumDistance = maxDist;
        cssHost::check(ctx(), css_set_submeshing(ctx(), _useSubMesh ? 1 : 0, maxDist), "css_set_submeshing");
        }
    void setNewSubmeshCutoff(double newCutoff) { useSubmeshingRoutines(true, newCutoff); }

    //! 0 closed (border edges are an error), 1 absorbing, 2 tangential: the rule the walker applies at border edges
    void setBoundaryMode(int mode)
        {
        boundaryMode = mode;
        cssHost::check(ctx(), css_set_boundary(ctx(), mode), "css_set_boundary");
        }
    int getBoundaryMode() const { return boundaryMode; }

    css_ctx* ctx() const { return holder->ctx; }
    double3 minVertexPosition{0, 0, 0}, maxVertexPosition{0, 0, 0};
    vector<double> vertices;  // [nV][3]
    vector<int32_t> faces;    // [nF][3], reference corner order
    int lastWalkFlags = 0;

protected:
    void euclid(vector<meshPosition>& p1, vector<double>& xyz)
        {
        size_t n = p1.size();
        vector<int32_t> f(n);
        vector<double> b(3 * n);
        xyz.assign(3 * n, 0.0);
        for (size_t i = 0; i < n; ++i)
            {
            f[i] = p1[i].faceIndex;
            for (int k = 0; k < 3; ++k) b[3 * i + k] = p1[i].x[k];
            }
        if (n) cssHost::check(ctx(), css_euclidean(ctx(), (int)n, f.data(), b.data(), xyz.data()), "css_euclidean");
        }
    shared_ptr<cssHost::contextHolder> holder;
    double area = 0;
    bool submeshingActivated = false;
    double maximumDistance = 0;
    int boundaryMode = 0;
    };
typedef gpuMeshSpace triangulatedMeshSpace;
typedef gpuMeshSpace closedMeshSpace;

//! openMeshSpace subclasses (src/models/absorbingOpenMeshSpace.h, tangentialOpenMeshSpace.h): the same walker kernel with the
//! border-edge rule switched on the device (css_set_boundary); distance() already treats boundary vertices as pseudo-sources.
class absorbingOpenMeshSpace : public gpuMeshSpace
    {
public:
    explicit absorbingOpenMeshSpace(int device = 0) : gpuMeshSpace(device) { setBoundaryMode(1); }
    };
class tangentialOpenMeshSpace : public gpuMeshSpace
    {
public:
    explicit tangentialOpenMeshSpace(int device = 0) : gpuMeshSpace(device) { setBoundaryMode(2); }
    };

inline double totalArea(gpuMeshSpace& space) { return space.getArea(); }

// ---------------------------------------------------------------------------------------------- neighbour structures
class baseNeighborStructure // all-to-all candidates (src/utility/baseNeighborStructure.cpp:17-36)
    {
public:
    virtual ~baseNeighborStructure() = default;
    virtual bool usesCellList() const { return false; }
    double3 domainMin{0, 0, 0}, domainMax{0, 0, 0};
    };
class cellListNeighborStructure : public baseNeighborStructure
    {
public:
    cellListNeighborStructure(double3 minPos, double3 maxPos, double gridSize) : interactionRange(gridSize)
        {
        domainMin = minPos;
        domainMax = maxPos;
        }
    virtual bool usesCellList() const { return true; }
    void setInteractionRange(double range) { interactionRange = range; }
    double interactionRange;
    };

// ---------------------------------------------------------------------------------------------- models
class simpleModel
    {
public:
    simpleModel() {}
    explicit simpleModel(int n) { initializeSimpleModel(n); }
    virtual ~simpleModel() = default;
    void initializeSimpleModel(int n)
        {
        N = n;
        positions.assign(n, meshPosition());
        velocities.assign(n, vector3(0, 0, 0));
        forces.assign(n, vector3(0, 0, 0));
        neighbors.assign(n, {});
        neighborVectors.assign(n, {});
        neighborDistances.assign(n, {});
        types.assign(n, 0);
        masses.assign(n, 1.0);
        }
    virtual void setSpace(shared_ptr<baseSpace> _space) { space = _space; }
    virtual void setNeighborStructure(shared_ptr<baseNeighborStructure> _structure) { neighborStructure = _structure; }
    virtual int getNumberOfParticles() { return N; }
    virtual void setParticlePositions(vector<meshPosition>& newPositions)
        {
        if (N != (int)newPositions.size()) initializeSimpleModel((int)newPositions.size());
        positions = newPositions;
        positionsChanged();
        }
    virtual void setRandomParticlePositions(noiseSource& noise)
        {
        for (int pp = 0; pp < N; ++pp) space->randomPosition(positions[pp], noise);
        positionsChanged();
        }
    //! simpleModel::R3PositionsToMeshPositions (simpleModel.cpp:136-154); the mesh is the one the space was loaded with
    virtual void R3PositionsToMeshPositions(vector<point3> r3positions, vector<meshPosition>& simPositions)
        {
        space->R3PositionsToMeshPositions(r3positions, simPositions, clampTolerance);
        }
    //! simpleModel::setMeshPositionsFromR3File (simpleModel.cpp:156-203): one "x,y,z" per line, malformed lines are skipped
    virtual void setMeshPositionsFromR3File(string filename)
        {
        std::ifstream file(filename);
        if (!file.is_open()) std::cerr << "Failed to open position file." << std::endl;
        vector<point3> points;
        string line;
        while (std::getline(file, line))
            {
            std::stringstream ss(line);
            string entry;
            vector<double> entries;
            while (std::getline(ss, entry, ','))
                {
                double value = 0;
                std::istringstream(entry) >> value;
                entries.push_back(value);
                }
            if (entries.size() != 3)
                {
                std::cerr << "Error: input file had invalid number of entries on a line. Skipping line." << std::endl;
                continue;
                }
            points.push_back(point3(entries[0], entries[1], entries[2]));
            }
        vector<meshPosition> simPositions;
        simPositions.reserve(points.size());
        R3PositionsToMeshPositions(points, simPositions);
        setParticlePositions(simPositions);
        }
    double clampTolerance = 0.00000000000001; // simpleModel.h:101
    virtual void setMaxwellBoltzmannVelocities(noiseSource& noise, double T)
        {
        for (int pp = 0; pp < N; ++pp)
            {
            space->randomVectorAtPosition(positions[pp], velocities[pp], noise);
            velocities[pp] *= std::sqrt(T);
            }
        }
    //! per-particle transport through the space, one call per particle (simpleModel.cpp:44-66)
    virtual void moveParticles(vector<vector3>& displacements)
        {
        for (int ii = 0; ii < N; ++ii)
            {
            vector<vector3> transports;
            if (particleShiftsRequireForceTransport) transports.push_back(forces[ii]);
            if (particleShiftsRequireVelocityTransport) transports.push_back(velocities[ii]);
            space->transportParticleAndVectors(positions[ii], displacements[ii], transports);
            size_t k = 0;
            if (particleShiftsRequireForceTransport) forces[ii] = transports[k++];
            if (particleShiftsRequireVelocityTransport) velocities[ii] = transports[k++];
            }
        }
    virtual void findNeighbors(double maximumInteractionRange) = 0;
    virtual void positionsChanged() {}
    void setVerbose(bool v) { verbose = v; }
    //! simpleModel::fillEuclideanLocations (simpleModel.cpp:28-31)
    virtual void fillEuclideanLocations() { space->meshPositionToEuclideanLocation(positions, euclideanLocations); }
    vector<double3> euclideanLocations;

    shared_ptr<baseSpace> space;
    int N = 0;
    vector<meshPosition> positions;
    vector<vector3> velocities, forces;
    vector<vector<int>> neighbors;
    vector<vector<vector3>> neighborVectors;
    vector<vector<double>> neighborDistances;
    vector<int> types;
    vector<double> masses;
    bool particleShiftsRequireVelocityTransport = false;
    bool particleShiftsRequireForceTransport = false;

protected:
    shared_ptr<baseNeighborStructure> neighborStructure;
    bool verbose = false;
    };
typedef shared_ptr<simpleModel> ConfigPtr;
typedef weak_ptr<simpleModel> WeakConfigPtr;

//! simpleModel whose findNeighbors / moveParticles run batched on the GPU (cell list, patches, exact geodesics,
//! walker).  The public host vectors are refreshed after every call, so user code reads them as before.
class gpuModel : public simpleModel
    {
public:
    explicit gpuModel(int n) : simpleModel(n) {}
    virtual void setSpace(shared_ptr<baseSpace> _space)
        {
        meshSpace = std::dynamic_pointer_cast<gpuMeshSpace>(_space);
        if (!meshSpace) ERRORERROR("gpuModel needs a gpuMeshSpace");
        space = _space;
        stateOnDevice = false;
        }
    virtual void setNeighborStructure(shared_ptr<baseNeighborStructure> _structure)
        {
        neighborStructure = _structure;
        bool cells = _structure && _structure->usesCellList();
        if (cells)
            {
            double mn[3] = {_structure->domainMin.x, _structure->domainMin.y, _structure->domainMin.z};
            double mx[3] = {_structure->domainMax.x, _structure->domainMax.y, _structure->domainMax.z};
            cssHost::check(ctx(), css_set_cell_domain(ctx(), mn, mx), "css_set_cell_domain");
            }
        cssHost::check(ctx(), css_set_options(ctx(), cells ? 1 : 0, 0), "css_set_options");
        }
    virtual void positionsChanged() { stateOnDevice = false; }

    virtual void moveParticles(vector<vector3>& displacements)
        {
        pushState();
        if (particleShiftsRequireVelocityTransport) pushVectors(velocities, css_set_velocities, "css_set_velocities");
        if (particleShiftsRequireForceTransport) pushVectors(forces, css_set_forces, "css_set_forces");
        vector<double> d(3 * (size_t)N);
        for (int i = 0; i < N; ++i)
            for (int k = 0; k < 3; ++k) d[3 * i + k] = displacements[i][k];
        cssHost::check(ctx(), css_move(ctx(), d.data(), particleShiftsRequireForceTransport, particleShiftsRequireVelocityTransport), "css_move");
        pullState(true, particleShiftsRequireVelocityTransport, particleShiftsRequireForceTransport);
        }
    virtual void findNeighbors(double maximumInteractionRange)
        {
        pushState();
        int64_t total = 0;
        cssHost::check(ctx(), css_find_neighbors(ctx(), maximumInteractionRange, &total), "css_find_neighbors");
        vector<int32_t> off(N + 1), idx((size_t)std::max<int64_t>(total, 1));
        vector<double> dist(idx.size()), ts(3 * idx.size());
        cssHost::check(ctx(), css_get_neighbors(ctx(), off.data(), idx.data(), dist.data(), ts.data(), nullptr), "css_get_neighbors");
        for (int i = 0; i < N; ++i)
            {
            int a = off[i], b = off[i + 1];
            neighbors[i].assign(idx.begin() + a, idx.begin() + b);
            neighborDistances[i].assign(dist.begin() + a, dist.begin() + b);
            neighborVectors[i].resize(b - a);
            for (int j = a; j < b; ++j) neighborVectors[i][j - a] = vector3(ts[3 * j], ts[3 * j + 1], ts[3 * j + 2]);
            }
        }
    //! fused neighbour search + pair force on the device (used by the stock pair potentials)
    void computeForcesOnDevice(int kind, const double params[3], bool zeroOutForces)
        {
        pushState();
        if (!zeroOutForces) pushVectors(forces, css_set_forces, "css_set_forces");
        cssHost::check(ctx(), css_compute_forces(ctx(), kind, params, zeroOutForces ? 1 : 0), "css_compute_forces");
        pullState(false, false, true);
        }
    double computeEnergyOnDevice(int kind, const double params[3])
        {
        pushState();
        double e = 0;
        cssHost::check(ctx(), css_compute_energy(ctx(), kind, params, &e), "css_compute_energy");
        return e;
        }
    void computeStressOnDevice(int kind, const double params[3], double stress[9])
        {
        pushState();
        pushVectors(velocities, css_set_velocities, "css_set_velocities"); // host-driven updaters kick the host copy
        cssHost::check(ctx(), css_compute_stress(ctx(), kind, params, stress), "css_compute_stress");
        }
    double temperatureOnDevice()
        {
        pushState();
        pushVectors(velocities, css_set_velocities, "css_set_velocities");
        double t = 0;
        cssHost::check(ctx(), css_temperature(ctx(), &t), "css_temperature");
        return t;
        }

    //! upload positions / velocities / forces if the host copies are newer
    void pushState()
        {
        if (stateOnDevice) return;
        vector<int32_t> f(N);
        vector<double> b(3 * (size_t)N), v(3 * (size_t)N), fr(3 * (size_t)N);
        for (int i = 0; i < N; ++i)
            {
            f[i] = positions[i].faceIndex;
            for (int k = 0; k < 3; ++k) b[3 * i + k] = positions[i].x[k], v[3 * i + k] = velocities[i][k], fr[3 * i + k] = forces[i][k];
            }
        cssHost::check(ctx(), css_set_state(ctx(), N, N, 0, f.data(), b.data(), v.data(), fr.data()), "css_set_state");
        stateOnDevice = true;
        }
    //! refresh the public host vectors from the device
    void pullState(bool pos = true, bool vel = true, bool frc = true)
        {
        vector<int32_t> f(pos ? N : 0);
        vector<double> b(pos ? 3 * (size_t)N : 0), v(vel ? 3 * (size_t)N : 0), fr(frc ? 3 * (size_t)N : 0);
        cssHost::check(ctx(), css_get_state(ctx(), pos ? f.data() : nullptr, pos ? b.data() : nullptr, vel ? v.data() : nullptr, frc ? fr.data() : nullptr),
                       "css_get_state");
        for (int i = 0; i < N; ++i)
            {
            if (pos) positions[i] = meshPosition(point3(b[3 * i], b[3 * i + 1], b[3 * i + 2]), f[i]);
            if (vel) velocities[i] = vector3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
            if (frc) forces[i] = vector3(fr[3 * i], fr[3 * i + 1], fr[3 * i + 2]);
            }
        }
    void pushVelocities() { pushState(), pushVectors(velocities, css_set_velocities, "css_set_velocities"); }
    css_ctx* ctx() const { return meshSpace->ctx(); }
    shared_ptr<gpuMeshSpace> meshSpace;

protected:
    void pushVectors(const vector<vector3>& src, int (*setter)(css_ctx*, const double*), const char* what)
        {
        vector<double> v(3 * (size_t)N);
        for (int i = 0; i < N; ++i)
            for (int k = 0; k < 3; ++k) v[3 * i + k] = src[i][k];
        cssHost::check(ctx(), setter(ctx(), v.data()), what);
        }
    bool stateOnDevice = false;
    };

// ---------------------------------------------------------------------------------------------- simulation base
class basicSimulation
    {
public:
    basicSimulation() : integerTimestep(0), Time(0.), integrationTimestep(0.), myRank(0), totalRanks(1), sortPeriod(-1) {}
    virtual ~basicSimulation() = default;
    virtual void computeForces() = 0;
    virtual void moveParticles(vector<vector3>& displacements) = 0;
    virtual double computePotentialEnergy(bool verbose = false) { (void)verbose; return 0.0; }
    WeakConfigPtr configuration;
    int integerTimestep;
    double Time;
    double integrationTimestep;
    void setSortPeriod(int sp) { sortPeriod = sp; }
    virtual void setCurrentTime(double _cTime) { Time = _cTime; }
    virtual void setCurrentTimestep(int _cTime) { integerTimestep = _cTime; }
    virtual void manipulateUpdaterData(vector<double>& data, function<double(double, double)> manipulatingFunction) { (void)data, (void)manipulatingFunction; }
    int myRank, totalRanks;

protected:
    int sortPeriod;
    };
typedef shared_ptr<basicSimulation> SimPtr;

// ---------------------------------------------------------------------------------------------- forces
class force
    {
public:
    virtual ~force() = default;
    virtual string reportSelfName() { return "base force"; }
    //! neighbour search, then the pair potential summed in neighbour order (baseForce.cpp:12-28).  This generic form
    //! calls the virtual pairwiseForce on the host, so user-written potentials keep working with a gpuModel.
    virtual void computeForces(vector<vector3>& forces, bool zeroOutForce = true, int type = 0)
        {
        (void)type;
        if ((int)forces.size() != model->N) forces.resize(model->N);
        model->findNeighbors(maximumInteractionRange);
        for (int ii = 0; ii < model->N; ++ii)
            {
            if (zeroOutForce) forces[ii] = vector3(0.0, 0.0, 0.0);
            for (size_t jj = 0; jj < model->neighbors[ii].size(); ++jj)
                forces[ii] += pairwiseForce(model->neighborVectors[ii][jj], model->neighborDistances[ii][jj]);
            }
        if (auto g = std::dynamic_pointer_cast<gpuModel>(model)) g->positionsChanged(); // host forces are now newer than the device's
        }
    virtual double computeEnergy(bool verbose = false)
        {
        (void)verbose;
        energy = 0.0;
        model->findNeighbors(maximumInteractionRange);
        for (int ii = 0; ii < model->N; ++ii)
            for (size_t jj = 0; jj < model->neighbors[ii].size(); ++jj)
                energy += pairwiseEnergy(model->neighborVectors[ii][jj], model->neighborDistances[ii][jj]);
        return energy;
        }
    virtual void setForceParameters(vector<double>& params) { (void)params; }
    virtual double pairwiseEnergy(vector3 separation, double distance) = 0;
    virtual vector3 pairwiseForce(vector3 separation, double distance) = 0;
    //! device functor id and parameter block, when the potential has one (otherwise the host callback path is used)
    virtual bool deviceKind(int& kind, double params[3]) { (void)kind, (void)params; return false; }
    void setSimulation(shared_ptr<basicSimulation> _sim) { sim = _sim; }
    virtual void setModel(shared_ptr<simpleModel> _model) { model = _model; }
    SimPtr sim;
    shared_ptr<simpleModel> model;
    double energy = 0;
    double maximumInteractionRange = 1;

protected:
    //! stock potentials on a gpuModel: fused neighbour search + force kernel
    bool tryDevice(vector<vector3>& forces, bool zeroOutForce)
        {
        auto g = std::dynamic_pointer_cast<gpuModel>(model);
        int kind;
        double p[3];
        if (!g || &forces != &g->forces || !deviceKind(kind, p)) return false;
        g->computeForcesOnDevice(kind, p, zeroOutForce);
        return true;
        }
    };
typedef shared_ptr<force> ForcePtr;
typedef weak_ptr<force> WeakForcePtr;

class harmonicRepulsion : public force
    {
public:
    harmonicRepulsion(double stiffness = 1, double monodisperseRange = 1, bool _monodisperse = true)
        : k(stiffness), sigma(monodisperseRange), monodisperse(_monodisperse) { maximumInteractionRange = monodisperseRange; }
    virtual string reportSelfName() { return "harmonic repulsion"; }
    virtual double pairwiseEnergy(vector3 separation, double distance)
        {
        (void)separation;
        return distance < sigma ? 0.5 * k * (sigma - distance) * (sigma - distance) : 0.0;
        }
    virtual vector3 pairwiseForce(vector3 separation, double distance)
        {
        return distance <= sigma ? (-k * (sigma - distance)) * separation : vector3(0, 0, 0);
        }
    virtual bool deviceKind(int& kind, double params[3])
        {
        kind = CSS_FORCE_HARMONIC, params[0] = k, params[1] = sigma, params[2] = maximumInteractionRange;
        return monodisperse;
        }
    virtual void computeForces(vector<vector3>& forces, bool zeroOutForce = true, int type = 0)
        {
        if (!tryDevice(forces, zeroOutForce)) force::computeForces(forces, zeroOutForce, type);
        }

protected:
    double k, sigma;
    bool monodisperse;
    };

class gaussianRepulsion : public force
    {
public:
    gaussianRepulsion(double strength, double variance) : alpha(strength), sigma(variance) {}
    virtual string reportSelfName() { return "gaussian repulsion"; }
    virtual double pairwiseEnergy(vector3 separation, double distance)
        {
        (void)separation;
        return alpha * std::exp(-distance * distance / (2.0 * sigma * sigma)) / (sqrtTwoPi * sigma);
        }
    virtual vector3 pairwiseForce(vector3 separation, double distance)
        { // sigma^{3/2} in the prefactor, as the reference codes it (gaussianRepulsion.h:21-23)
        double pre = distance * alpha * std::exp(-distance * distance / (2.0 * sigma * sigma)) / ((sqrtTwoPi * sigma) * std::sqrt(sigma));
        return (-pre) * separation;
        }
    virtual bool deviceKind(int& kind, double params[3])
        {
        kind = CSS_FORCE_GAUSSIAN, params[0] = alpha, params[1] = sigma, params[2] = maximumInteractionRange;
        return true;
        }
    virtual void computeForces(vector<vector3>& forces, bool zeroOutForce = true, int type = 0)
        {
        if (!tryDevice(forces, zeroOutForce)) force::computeForces(forces, zeroOutForce, type);
        }

protected:
    double alpha, sigma;
    const double sqrtTwoPi = 2.50662827463100050241576528481104525300698674061;
    };

// ---------------------------------------------------------------------------------------------- updaters
class updater
    {
public:
    updater() : Period(-1), Phase(0), reproducible(true) {}
    explicit updater(int _p) : Period(_p), Phase(0), reproducible(true) {}
    virtual ~updater() = default;
    virtual void Update(int timestep)
        {
        iterations = timestep;
        if (maxIterations > 0 && maxIterations < iterations) return;
        if (Period <= 0 || (timestep + Phase) % Period == 0) performUpdate();
        }
    virtual void performUpdate() { sim->computeForces(); }
    void setSimulation(shared_ptr<basicSimulation> _sim) { sim = _sim; }
    virtual void setModel(shared_ptr<simpleModel> _model) { model = _model, initializeFromModel(); }
    virtual void initializeFromModel() { Ndof = model->getNumberOfParticles(); }
    void setPeriod(int _p) { Period = _p; }
    void setPhase(int _p) { Phase = _p; }
    virtual void setReproducible(bool rep) { reproducible = rep; }
    int getNdof() { return Ndof; }
    void setNdof(int _n) { Ndof = _n; }
    virtual void setDeltaT(double dt) { deltaT = dt; }
    double getDeltaT() const { return deltaT; }
    virtual double getMaxForce()
        {
        vector<double> m(1, 0.0);
        for (int ii = 0; ii < Ndof; ++ii) m[0] = std::max(m[0], model->forces[ii].squared_length());
        sim->manipulateUpdaterData(m, [](double x, double y) { return std::max(x, y); });
        return maximumForceNorm = std::sqrt(m[0]);
        }
    virtual double getForceNorm()
        {
        vector<double> s(1, 0.0);
        for (int ii = 0; ii < Ndof; ++ii) s[0] += model->forces[ii].squared_length();
        sim->manipulateUpdaterData(s, [](double x, double y) { return x + y; });
        squaredTotalForceNorm = s[0];
        return std::sqrt(s[0]);
        }
    void setMaximumIterations(int maxIt = -1) { maxIterations = maxIt; }
    int getCurrentIterations() { return iterations; }
    //! which fused device step implements this updater (0 none, 1 NVE, 2 GD, 3 NVT, 4 FIRE); see gpuSimulation
    virtual int fusedKind() const { return 0; }
    shared_ptr<basicSimulation> sim;
    shared_ptr<simpleModel> model;
    int iterations = 0;
    double squaredTotalForceNorm = 0, maximumForceNorm = 0;

protected:
    int Period, Phase, Ndof = 0;
    bool reproducible;
    double deltaT = 0.001;
    int maxIterations = -1;
    };
typedef shared_ptr<updater> UpdaterPtr;
typedef weak_ptr<updater> WeakUpdaterPtr;

class velocityVerletNVE : public updater
    {
public:
    velocityVerletNVE() { deltaT = 0.001; }
    explicit velocityVerletNVE(double _dt) { deltaT = _dt; }
    virtual void performUpdate()
        {
        velocityVerletFirstHalfStep();
        velocityVerletSecondHalfStep();
        }
    void velocityVerletFirstHalfStep()
        {
        displacements.resize(Ndof);
        for (int ii = 0; ii < Ndof; ++ii)
            {
            displacements[ii] = deltaT * model->velocities[ii] + 0.5 * deltaT * deltaT * model->forces[ii];
            model->velocities[ii] += 0.5 * deltaT * model->forces[ii];
            }
        }
    void velocityVerletSecondHalfStep()
        {
        sim->moveParticles(displacements);
        sim->computeForces();
        for (int ii = 0; ii < Ndof; ++ii) model->velocities[ii] += 0.5 * deltaT * model->forces[ii];
        }
    virtual void setModel(shared_ptr<simpleModel> _model)
        {
        model = _model;
        model->particleShiftsRequireVelocityTransport = true;
        initializeFromModel();
        }
    virtual int fusedKind() const { return 1; }

protected:
    vector<vector3> displacements;
    };

class gradientDescent : public updater
    {
public:
    explicit gradientDescent(double _dt) { deltaT = _dt; }
    virtual void performUpdate()
        {
        displacements.resize(Ndof);
        sim->computeForces();
        for (int ii = 0; ii < Ndof; ++ii) displacements[ii] = deltaT * model->forces[ii];
        sim->moveParticles(displacements);
        }
    virtual int fusedKind() const { return 2; }

protected:
    vector<vector3> displacements;
    };

class noseHooverNVT : public updater
    {
public:
    noseHooverNVT(double _dt, double _T, double _tau = 1.0, int _M = 2) : temperature(_T), chainLength(_M), tau(_tau)
        {
        deltaT = _dt, dt2 = 0.5 * _dt, dt4 = 0.25 * _dt, dt8 = 0.125 * _dt;
        bathVariables.assign(_M + 1, double4{0, 0, 0, 0});
        }
    virtual void setModel(shared_ptr<simpleModel> _model)
        {
        model = _model;
        model->particleShiftsRequireVelocityTransport = true;
        initializeFromModel();
        setBathVariables();
        }
    virtual void performUpdate()
        {
        displacements.resize(Ndof);
        propagateChain();
        for (int ii = 0; ii < Ndof; ++ii) model->velocities[ii] = kineticEnergyScaleFactor * model->velocities[ii];
        propagatePositionsVelocities();
        propagateChain();
        for (int ii = 0; ii < Ndof; ++ii) model->velocities[ii] = kineticEnergyScaleFactor * model->velocities[ii];
        }
    double getTemperatureFromKE()
        {
        double v2 = 0;
        for (int i = 0; i < Ndof; ++i) v2 += model->velocities[i] * model->velocities[i];
        return v2 / (2 * Ndof);
        }
    virtual int fusedKind() const { return 3; }
    double temperature;
    int chainLength;
    double tau;
    vector<double4> bathVariables;
    double kineticEnergy = 0, kineticEnergyScaleFactor = 1;

protected:
    void setBathVariables()
        { // first bath mass 2 (Ndof - 1) T tau^2, the others T tau^2 (noseHooverNVT.cpp:28-36)
        bathVariables[0].w = 2.0 * (Ndof - 1) * temperature * tau * tau;
        for (size_t ii = 1; ii < bathVariables.size(); ++ii) bathVariables[ii].w = temperature * tau * tau;
        kineticEnergy = bathVariables[0].w;
        kineticEnergyScaleFactor = 1.0;
        }
    void bathKick(int ii)
        {
        double ef = std::exp(-dt8 * bathVariables[ii + 1].y);
        bathVariables[ii].y *= ef;
        bathVariables[ii].y += bathVariables[ii].z * dt4;
        bathVariables[ii].y *= ef;
        }
    void propagateChain()
        { // Martyna-Tuckerman-Tobias-Klein chain, quarter / half / quarter step (noseHooverNVT.cpp:65-110)
        auto& b = bathVariables;
        for (int ii = chainLength - 1; ii > 0; --ii)
            {
            b[ii].z = (b[ii - 1].w * b[ii - 1].y * b[ii - 1].y - temperature) / b[ii].w;
            bathKick(ii);
            }
        b[0].z = 2.0 * kineticEnergy / b[0].w - 1.0;
        bathKick(0);
        for (int ii = 0; ii < chainLength; ++ii) b[ii].x += dt2 * b[ii].y;
        kineticEnergyScaleFactor = std::exp(-dt2 * b[0].y);
        kineticEnergy = kineticEnergyScaleFactor * kineticEnergyScaleFactor * kineticEnergy;
        b[0].z = 2.0 * kineticEnergy / b[0].w - 1.0;
        bathKick(0);
        for (int ii = 1; ii < chainLength; ++ii)
            {
            b[ii].z = (b[ii - 1].w * b[ii - 1].y * b[ii - 1].y - temperature) / b[ii].w;
            bathKick(ii);
            }
        }
    void propagatePositionsVelocities()
        { // half move, forces, full kick, half move (noseHooverNVT.cpp:116-139)
        kineticEnergy = 0.0;
        for (int ii = 0; ii < Ndof; ++ii) displacements[ii] = dt2 * model->velocities[ii];
        sim->moveParticles(displacements);
        sim->computeForces();
        for (int ii = 0; ii < Ndof; ++ii)
            {
            model->velocities[ii] = model->velocities[ii] + (deltaT / model->masses[ii]) * model->forces[ii];
            displacements[ii] = dt2 * model->velocities[ii];
            kineticEnergy += 0.5 * model->masses[ii] * (model->velocities[ii] * model->velocities[ii]);
            }
        sim->moveParticles(displacements);
        }
    vector<vector3> displacements;
    double dt2, dt4, dt8;
    };

class fireMinimization : public velocityVerletNVE
    {
public:
    fireMinimization() { deltaT = 0.001, alpha = 0.99; }
    virtual void performUpdate() { minimizeByFire(); }
    void minimizeByFire()
        {
        sim->computeForces();
        forceMax = getMaxForce();
        iterations = 0;
        while (iterations < maximumIterations && forceMax > forceCutoff)
            {
            iterations += 1;
            velocityVerletFirstHalfStep();
            velocityVerletSecondHalfStep();
            fireStep();
            forceMax = getMaxForce();
            }
        }
    void fireStep()
        {
        forceNorm = dotProduct(model->forces, model->forces);
        velocityNorm = dotProduct(model->velocities, model->velocities);
        power = dotProduct(model->forces, model->velocities);
        double scaling = forceNorm > 0 ? std::sqrt(velocityNorm / forceNorm) : 0.0;
        for (int ii = 0; ii < Ndof; ++ii) model->velocities[ii] = (1 - alpha) * model->velocities[ii] + alpha * scaling * model->forces[ii];
        if (power > 0)
            {
            if (nSinceNegativePower > nMin)
                {
                deltaT = std::min(deltaT * deltaTInc, deltaTMax);
                alpha = std::max(alpha * alphaDec, alphaMin);
                }
            nSinceNegativePower += 1;
            }
        else
            {
            nSinceNegativePower = 0;
            deltaT = std::max(deltaT * deltaTDec, deltaTMin);
            alpha = alphaStart;
            model->velocities.assign(Ndof, vector3(0., 0., 0.));
            }
        }
    //! NB: as in the reference (fireMinimization.cpp:74-90) the _deltaT argument is accepted and ignored
    void setFIREParameters(int _maximumIterations, double _deltaT, double _alphaStart, double _deltaTMax, double _deltaTMin, double _deltaTInc,
                           double _deltaTDec, double _alphaDec, int _nMin, double _forceCutoff, double _alphaMin = 0.75)
        {
        (void)_deltaT;
        maximumIterations = _maximumIterations, alphaStart = _alphaStart, deltaTMax = _deltaTMax, deltaTInc = _deltaTInc;
        deltaTMin = _deltaTMin, deltaTDec = _deltaTDec, alphaDec = _alphaDec, forceCutoff = _forceCutoff, alphaMin = _alphaMin, nMin = _nMin;
        alpha = alphaStart;
        }
    virtual void setModel(shared_ptr<simpleModel> _model)
        {
        model = _model;
        model->particleShiftsRequireVelocityTransport = true;
        model->particleShiftsRequireForceTransport = true;
        initializeFromModel();
        }
    virtual int fusedKind() const { return 4; }
    void fireParameterBlock(double p[11]) const
        {
        double q[11] = {(double)maximumIterations, deltaT, alphaStart, deltaTMax, deltaTMin, deltaTInc, deltaTDec, alphaDec, (double)nMin, forceCutoff, alphaMin};
        std::copy(q, q + 11, p);
        }
    double forceMax = 0, power = 0, forceNorm = 0, velocityNorm = 0, alpha;
    int nSinceNegativePower = 0;

protected:
    double dotProduct(vector<vector3>& a, vector<vector3>& b)
        {
        vector<double> s(1, 0.0);
        for (size_t ii = 0; ii < a.size(); ++ii) s[0] += a[ii] * b[ii];
        sim->manipulateUpdaterData(s, [](double x, double y) { return x + y; });
        return s[0];
        }
    int maximumIterations = 1000, nMin = 4;
    double alphaStart = 0.99, deltaTMax = 0.1, deltaTInc = 1.1, deltaTMin = 1e-5, deltaTDec = 0.95, alphaDec = 0.9, forceCutoff = 1e-12, alphaMin = 0.0;
    };

// ---------------------------------------------------------------------------------------------- simulation
class simulation : public basicSimulation, public std::enable_shared_from_this<simulation>
    {
public:
    void setConfiguration(ConfigPtr _config) { configuration = _config; }
    virtual void computeForces()
        {
        auto Conf = configuration.lock();
        for (size_t f = 0; f < forceComputers.size(); ++f) forceComputers[f].lock()->computeForces(Conf->forces, f == 0);
        }
    virtual void moveParticles(vector<vector3>& displacements) { configuration.lock()->moveParticles(displacements); }
    virtual void performTimestep()
        {
        integerTimestep += 1;
        Time += integrationTimestep;
        for (auto& u : updaters) u.lock()->Update(integerTimestep);
        }
    shared_ptr<simulation> getPointer() { return shared_from_this(); }
    void addUpdater(UpdaterPtr _upd) { updaters.push_back(_upd); }
    void addUpdater(UpdaterPtr _upd, ConfigPtr _config)
        {
        _upd->setModel(_config);
        _upd->setSimulation(getPointer());
        updaters.push_back(_upd);
        }
    void addForce(ForcePtr _force) { forceComputers.push_back(_force); }
    void addForce(ForcePtr _force, ConfigPtr _config)
        {
        _force->setModel(_config);
        forceComputers.push_back(_force);
        }
    void clearForceComputers() { forceComputers.clear(); }
    void clearUpdaters() { updaters.clear(); }
    //! virial + kinetic "stress" of a monodisperse system (simulation.cpp:104-173), flattened 3x3.  A stock pair potential on
    //! a gpuModel is evaluated on the device; anything else runs the reference's double loop over the host neighbour lists.
    void computeMonodisperseStress(vector<double>& stress)
        {
        auto conf = configuration.lock();
        stress.assign(9, 0.0);
        int kind = 0;
        double p[3];
        auto g = std::dynamic_pointer_cast<gpuModel>(conf);
        if (g && forceComputers.size() == 1 && forceComputers[0].lock()->deviceKind(kind, p))
            {
            g->computeStressOnDevice(kind, p, stress.data());
            return;
            }
        auto f0 = forceComputers[0].lock();
        double area = conf->space->getArea();
        int Ndof = conf->N;
        double density = Ndof / area;
        conf->findNeighbors(f0->maximumInteractionRange);
        double fOuterDR[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, vOuterv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        bool first = true;
        for (auto& wf : forceComputers)
            {
            auto frc = wf.lock();
            for (int ii = 0; ii < Ndof; ++ii)
                for (size_t jj = 0; jj < conf->neighbors[ii].size(); ++jj)
                    {
                    vector3 sep = conf->neighborVectors[ii][jj];
                    vector3 force = frc->pairwiseForce(sep, conf->neighborDistances[ii][jj]);
                    for (int a = 0; a < 3; ++a)
                        for (int b = 0; b < 3; ++b)
                            {
                            fOuterDR[3 * a + b] += force[a] * sep[b];
                            if (first) vOuterv[3 * a + b] += conf->velocities[ii][a] * conf->velocities[ii][b];
                            }
                    }
            first = false;
            }
        for (int q = 0; q < 9; ++q) stress[q] = density * vOuterv[q] / (2 * Ndof) + fOuterDR[q] / (2 * 2 * area * Ndof);
        }
    void setIntegrationTimestep(double dt)
        {
        integrationTimestep = dt;
        for (auto& u : updaters) u.lock()->setDeltaT(dt);
        }
    void setReproducible(bool reproducible)
        {
        for (auto& u : updaters) u.lock()->setReproducible(reproducible);
        }
    vector<WeakUpdaterPtr> updaters;
    vector<WeakForcePtr> forceComputers;
    };
typedef shared_ptr<simulation> SimulationPtr;

//! simulation whose performTimestep keeps the state on the GPU: with one stock pair potential and one stock
//! updater the whole step (walker, cell list, patches, geodesics, forces, kicks) is a single ABI call and the
//! host vectors are refreshed only by syncHost().  Anything else falls back to simulation::performTimestep.
class gpuSimulation : public simulation
    {
public:
    virtual void performTimestep()
        {
        auto g = std::dynamic_pointer_cast<gpuModel>(configuration.lock());
        int kind = 0;
        double p[3];
        UpdaterPtr u = updaters.size() == 1 ? updaters[0].lock() : nullptr;
        bool fused = g && u && u->fusedKind() != 0 && forceComputers.size() == 1 && forceComputers[0].lock()->deviceKind(kind, p);
        if (!fused)
            {
            simulation::performTimestep();
            return;
            }
        integerTimestep += 1;
        Time += integrationTimestep;
        g->pushState();
        css_ctx* c = g->ctx();
        switch (u->fusedKind())
            {
            case 1: // the first step uses whatever forces the model holds (zero unless the caller computed them), as the reference does
                cssHost::check(c, css_step_nve(c, kind, p, u->getDeltaT(), 1), "css_step_nve");
                break;
            case 2: cssHost::check(c, css_step_gd(c, kind, p, u->getDeltaT(), 1), "css_step_gd"); break;
            case 3:
                {
                auto nh = std::static_pointer_cast<noseHooverNVT>(u);
                if (!forcesPrimed)
                    {
                    cssHost::check(c, css_nvt_init(c, nh->getDeltaT(), nh->temperature, nh->tau, nh->chainLength), "css_nvt_init");
                    forcesPrimed = true;
                    }
                cssHost::check(c, css_step_nvt(c, kind, p, 1), "css_step_nvt");
                break;
                }
            case 4:
                {
                auto fire = std::static_pointer_cast<fireMinimization>(u);
                double fp[11], out[4];
                fire->fireParameterBlock(fp);
                if (!forcesPrimed) cssHost::check(c, css_fire_init(c, fp, fire->getDeltaT(), fire->alpha), "css_fire_init"), forcesPrimed = true;
                cssHost::check(c, css_fire_minimize(c, kind, p, out), "css_fire_minimize");
                fire->iterations = (int)out[0], fire->forceMax = out[1], fire->setDeltaT(out[2]), fire->alpha = out[3];
                break;
                }
            }
        hostStale = true;
        }
    //! refresh positions / velocities / forces of the model from the device (call before reading them on the host)
    void syncHost()
        {
        auto g = std::dynamic_pointer_cast<gpuModel>(configuration.lock());
        if (g && hostStale) g->pullState(true, true, true);
        hostStale = false;
        }

protected:
    bool forcesPrimed = false, hostStale = false;
    };
