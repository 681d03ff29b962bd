// Trajectory and scalar/vector time-series databases in the reference's LOGICAL layout
// (src/databases/simpleModelDatabase.cpp:37-50, src/databases/vectorValueDatabase.cpp:25-29): the same dataset
// names, element types, row widths and append-one-record-per-write semantics as its extendable HDF5 datasets.
// HDF5 is not part of this image, so a database is a directory `<name>/` holding `meta.txt` and one raw
// little-endian `<dataset>.bin` per dataset, rows appended in record order; curvedspacesim_b200/trajectory.py reads
// and writes the same files and converts them to the reference's HDF5 files where h5py exists.
//
//   meta.txt:   cssdb 1
//               N <particles>                         (simpleModelDatabase only)
//               dataset <name> <f64|i32> <row width>  (one line per dataset, in registration order)
#pragma once
#include "css_host.hpp"
#include <cstdio>
#include <sys/stat.h>

namespace fileMode
    {
enum Enum
    {
    readonly,  //!< we just want to read
    readwrite, //!< we intend to both read and write
    replace    //!< we will completely overwrite all of the data
    };
    }

class baseRawDatabase
    {
public:
    baseRawDatabase(string fn, fileMode::Enum _mode) : filename(fn), mode(_mode)
        {
        struct stat st;
        bool exists = stat((filename + "/meta.txt").c_str(), &st) == 0;
        if (mode == fileMode::readonly && !exists) ERRORERROR("database does not exist");
        if (mode != fileMode::readonly) mkdir(filename.c_str(), 0755);
        if (exists && mode != fileMode::replace) readMeta();
        }
    virtual ~baseRawDatabase() = default;
    string filename;
    fileMode::Enum mode;

protected:
    struct dataset
        {
        string name;
        bool isDouble;
        size_t width;
        };
    vector<dataset> datasets;
    long metaN = -1;
    size_t elementSize(const dataset& d) const { return d.isDouble ? sizeof(double) : sizeof(int32_t); }
    string path(const string& name) const { return filename + "/" + name + ".bin"; }
    const dataset& find(const string& name) const
        {
        for (auto& d : datasets)
            if (d.name == name) return d;
        ERRORERROR("unknown dataset");
        }
    template <typename T> void registerExtendableDataset(const string& name, size_t width)
        {
        datasets.push_back(dataset{name, std::is_same<T, double>::value, width});
        FILE* f = fopen(path(name).c_str(), "wb"); // created empty (replace) or on first registration
        if (!f) ERRORERROR("cannot create dataset file");
        fclose(f);
        writeMeta();
        }
    void writeMeta()
        {
        FILE* f = fopen((filename + "/meta.txt").c_str(), "w");
        if (!f) ERRORERROR("cannot write database meta file");
        fprintf(f, "cssdb 1\n");
        if (metaN >= 0) fprintf(f, "N %ld\n", metaN);
        for (auto& d : datasets) fprintf(f, "dataset %s %s %zu\n", d.name.c_str(), d.isDouble ? "f64" : "i32", d.width);
        fclose(f);
        }
    void readMeta()
        {
        std::ifstream in(filename + "/meta.txt");
        string tok;
        int version = 0;
        if (!(in >> tok >> version) || tok != "cssdb") ERRORERROR("not a cssdb database");
        datasets.clear();
        while (in >> tok)
            {
            if (tok == "N") in >> metaN;
            else if (tok == "dataset")
                {
                dataset d;
                string ty;
                in >> d.name >> ty >> d.width;
                d.isDouble = ty == "f64";
                datasets.push_back(d);
                }
            }
        }
    unsigned long getDatasetDimensions(const string& name) const
        {
        for (auto& d : datasets)
            if (d.name == name)
                {
                struct stat st;
                if (stat(path(name).c_str(), &st) != 0) return 0;
                return (unsigned long)(st.st_size / (elementSize(d) * d.width));
                }
        return 0;
        }
    template <typename T> void extendDataset(const string& name, const vector<T>& row)
        {
        const dataset& d = find(name);
        if (row.size() != d.width || std::is_same<T, double>::value != d.isDouble) ERRORERROR("row does not match the dataset");
        if (mode == fileMode::readonly) ERRORERROR("database opened read-only");
        FILE* f = fopen(path(name).c_str(), "ab");
        if (!f || fwrite(row.data(), sizeof(T), row.size(), f) != row.size()) ERRORERROR("dataset write failed");
        fclose(f);
        }
    template <typename T> void readDataset(const string& name, vector<T>& row, int record) const
        {
        const dataset& d = find(name);
        long n = (long)getDatasetDimensions(name);
        if (record < 0) record += (int)n; // -1 = last record
        if (record < 0 || record >= n) ERRORERROR("record out of range");
        row.resize(d.width);
        FILE* f = fopen(path(name).c_str(), "rb");
        if (!f || fseek(f, (long)(sizeof(T) * d.width) * record, SEEK_SET) != 0 || fread(row.data(), sizeof(T), d.width, f) != d.width)
            ERRORERROR("dataset read failed");
        fclose(f);
        }
    };

//! time, R3position, barycentricPosition, faceIndex [, velocity] [, force] [, type]  (simpleModelDatabase.cpp:37-50)
class simpleModelDatabase : public baseRawDatabase
    {
public:
    typedef shared_ptr<simpleModel> STATE;
    simpleModelDatabase(int numberOfParticles, string fn = "temp.cssdb", fileMode::Enum _mode = fileMode::readonly, bool saveVelocities = true,
                        bool saveTypes = true, bool saveForces = true)
        : baseRawDatabase(fn, _mode), N(numberOfParticles), velocity(saveVelocities), type(saveTypes), force(saveForces)
        {
        if (metaN >= 0 && metaN != N && mode != fileMode::replace) ERRORERROR("database holds a different number of particles");
        metaN = N;
        if (mode == fileMode::replace || (mode == fileMode::readwrite && datasets.empty()))
            {
            datasets.clear();
            registerDatasets();
            }
        }
    unsigned long currentNumberOfRecords() { return getDatasetDimensions("time"); }
    //! appends one record (rec must stay -1, as in the reference: "overwriting specific records not implemented")
    virtual void writeState(STATE s, double time = -1.0, int rec = -1)
        {
        if (rec >= 0) ERRORERROR("overwriting specific records not implemented at the moment");
        extendDataset("time", vector<double>{time});
        vector<double> r3pos(3 * (size_t)N), baryPos(3 * (size_t)N), vel(3 * (size_t)N), forceVector(3 * (size_t)N);
        vector<int32_t> typeVector(N), faceIdx(N);
        s->fillEuclideanLocations();
        for (int ii = 0; ii < N; ++ii)
            {
            r3pos[3 * ii] = s->euclideanLocations[ii].x, r3pos[3 * ii + 1] = s->euclideanLocations[ii].y, r3pos[3 * ii + 2] = s->euclideanLocations[ii].z;
            for (int k = 0; k < 3; ++k)
                {
                baryPos[3 * ii + k] = s->positions[ii].x[k];
                vel[3 * ii + k] = s->velocities[ii][k];
                forceVector[3 * ii + k] = s->forces[ii][k];
                }
            faceIdx[ii] = s->positions[ii].faceIndex;
            typeVector[ii] = s->types[ii];
            }
        extendDataset("R3position", r3pos);
        extendDataset("barycentricPosition", baryPos);
        extendDataset("faceIndex", faceIdx);
        if (velocity) extendDataset("velocity", vel);
        if (force) extendDataset("force", forceVector);
        if (type) extendDataset("type", typeVector);
        }
    //! restores positions (and velocities / forces / types when stored) of record rec (-1 = last)
    virtual void readState(STATE s, int rec)
        {
        vector<double> t, baryPos, vel, forceVector;
        vector<int32_t> typeVector, faceIdx;
        readDataset("time", t, rec);
        lastTime = t[0];
        readDataset("barycentricPosition", baryPos, rec);
        readDataset("faceIndex", faceIdx, rec);
        if (velocity) readDataset("velocity", vel, rec);
        if (force) readDataset("force", forceVector, rec);
        if (type) readDataset("type", typeVector, rec);
        for (int ii = 0; ii < N; ++ii)
            {
            s->positions[ii] = meshPosition(point3(baryPos[3 * ii], baryPos[3 * ii + 1], baryPos[3 * ii + 2]), faceIdx[ii]);
            if (velocity) s->velocities[ii] = vector3(vel[3 * ii], vel[3 * ii + 1], vel[3 * ii + 2]);
            if (force) s->forces[ii] = vector3(forceVector[3 * ii], forceVector[3 * ii + 1], forceVector[3 * ii + 2]);
            if (type) s->types[ii] = typeVector[ii];
            }
        s->positionsChanged();
        }
    double lastTime = 0;

protected:
    void registerDatasets()
        {
        registerExtendableDataset<double>("time", 1);
        registerExtendableDataset<double>("R3position", 3 * (size_t)N);
        registerExtendableDataset<double>("barycentricPosition", 3 * (size_t)N);
        registerExtendableDataset<int32_t>("faceIndex", N);
        if (velocity) registerExtendableDataset<double>("velocity", 3 * (size_t)N);
        if (force) registerExtendableDataset<double>("force", 3 * (size_t)N);
        if (type) registerExtendableDataset<int32_t>("type", N);
        }
    int N;
    bool velocity, type, force;
    };

//! value, vector  (vectorValueDatabase.cpp:25-29)
class valueVectorDatabase : public baseRawDatabase
    {
public:
    valueVectorDatabase(string fn, unsigned long vectorSize, fileMode::Enum _mode = fileMode::readonly)
        : baseRawDatabase(fn, _mode), maximumVectorSize(vectorSize)
        {
        valueVector.resize(1);
        dataVector.resize(maximumVectorSize);
        if (mode == fileMode::replace || (mode == fileMode::readwrite && datasets.empty()))
            {
            datasets.clear();
            registerExtendableDataset<double>("value", 1);
            registerExtendableDataset<double>("vector", maximumVectorSize);
            }
        }
    unsigned long currentNumberOfRecords() { return getDatasetDimensions("value"); }
    void writeState(double val, vector<double>& data)
        {
        extendDataset("vector", data);
        valueVector[0] = val;
        extendDataset("value", valueVector);
        }
    void readState(int record)
        {
        readDataset("value", valueVector, record);
        readDataset("vector", dataVector, record);
        }
    vector<double> valueVector, dataVector;

protected:
    unsigned long maximumVectorSize;
    };
