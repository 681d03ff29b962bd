"""Trajectory / time-series databases in the reference's logical layout.

The reference stores trajectories in HDF5 files of extendable datasets, one row per record:
`simpleModelDatabase` -> time[1], R3position[3N], barycentricPosition[3N], faceIndex[N], velocity[3N], force[3N], type[N]
(src/databases/simpleModelDatabase.cpp:37-50) and `valueVectorDatabase` -> value[1], vector[M]
(src/databases/vectorValueDatabase.cpp:25-29).  HDF5 is not available in this image, so the same datasets (names, element
types, row widths, append-per-record semantics) live in a directory of raw little-endian files described by `meta.txt`;
host/css_database.hpp writes and reads the identical files from C++.  `to_hdf5` / `from_hdf5` convert to and from the
reference's own files wherever h5py exists.
"""
from __future__ import annotations

import os

import numpy as np

_DT = {"f64": np.dtype("<f8"), "i32": np.dtype("<i4")}


class RawDatabase:
    def __init__(self, path: str, mode: str = "r"):
        """mode: 'r' (fileMode::readonly), 'a' (readwrite), 'w' (replace)."""
        if mode not in ("r", "a", "w"):
            raise ValueError(mode)
        self.path, self.mode = path, mode
        self.datasets: list[tuple[str, str, int]] = []
        self.N = None
        meta = os.path.join(path, "meta.txt")
        if mode == "r" and not os.path.exists(meta):
            raise FileNotFoundError(meta)
        if mode != "r":
            os.makedirs(path, exist_ok=True)
        if os.path.exists(meta) and mode != "w":
            self._read_meta()

    # ---- meta ----
    def _read_meta(self):
        with open(os.path.join(self.path, "meta.txt")) as fh:
            head = fh.readline().split()
            if head[:1] != ["cssdb"]:
                raise ValueError("not a cssdb database: %s" % self.path)
            for line in fh:
                t = line.split()
                if not t:
                    continue
                if t[0] == "N":
                    self.N = int(t[1])
                elif t[0] == "dataset":
                    self.datasets.append((t[1], t[2], int(t[3])))

    def _write_meta(self):
        with open(os.path.join(self.path, "meta.txt"), "w") as fh:
            fh.write("cssdb 1\n")
            if self.N is not None:
                fh.write("N %d\n" % self.N)
            for name, ty, width in self.datasets:
                fh.write("dataset %s %s %d\n" % (name, ty, width))

    def _register(self, name, ty, width):
        self.datasets.append((name, ty, int(width)))
        open(self._file(name), "wb").close()
        self._write_meta()

    def _file(self, name):
        return os.path.join(self.path, name + ".bin")

    def _spec(self, name):
        for n, ty, width in self.datasets:
            if n == name:
                return _DT[ty], width
        raise KeyError(name)

    # ---- rows ----
    def records(self, name):
        dt, width = self._spec(name)
        return os.path.getsize(self._file(name)) // (dt.itemsize * width) if os.path.exists(self._file(name)) else 0

    def extend(self, name, row):
        if self.mode == "r":
            raise IOError("database opened read-only")
        dt, width = self._spec(name)
        row = np.ascontiguousarray(row, dt).reshape(-1)
        if row.size != width:
            raise ValueError("row of %d values for dataset %s of width %d" % (row.size, name, width))
        with open(self._file(name), "ab") as fh:
            fh.write(row.tobytes())

    def read(self, name, record=None):
        """One record (negative = from the end) or, with record=None, the whole [records, width] array."""
        dt, width = self._spec(name)
        n = self.records(name)
        if record is None:
            return np.fromfile(self._file(name), dt, n * width).reshape(n, width)
        if record < 0:
            record += n
        if not 0 <= record < n:
            raise IndexError("record out of range")
        return np.fromfile(self._file(name), dt, width, offset=record * width * dt.itemsize)

    # ---- HDF5 interchange (needs h5py; the reference's baseHDF5Database uses unlimited first dimension, chunked rows) ----
    def to_hdf5(self, filename):
        import h5py  # noqa: PLC0415  (absent in the build image; present wherever the reference's files are used)

        with h5py.File(filename, "w") as f:
            for name, ty, width in self.datasets:
                f.create_dataset(name, data=self.read(name), maxshape=(None, width), chunks=(1, width))

    @classmethod
    def from_hdf5(cls, filename, path):
        import h5py  # noqa: PLC0415

        db = cls(path, "w")
        with h5py.File(filename, "r") as f:
            if "faceIndex" in f:
                db.N = int(f["faceIndex"].shape[1])
            for name in f:
                arr = np.asarray(f[name])
                ty = "f64" if arr.dtype.kind == "f" else "i32"
                db._register(name, ty, arr.shape[1])
                with open(db._file(name), "wb") as fh:
                    fh.write(np.ascontiguousarray(arr, _DT[ty]).tobytes())
        return db


class SimpleModelDatabase(RawDatabase):
    """simpleModelDatabase (src/databases/simpleModelDatabase.cpp): one record per write_state."""

    def __init__(self, n_particles, path="temp.cssdb", mode="r", save_velocities=True, save_types=True, save_forces=True):
        super().__init__(path, mode)
        if self.N is not None and self.N != n_particles and mode != "w":
            raise ValueError("database holds %d particles, not %d" % (self.N, n_particles))
        self.N = int(n_particles)
        self.velocity, self.type, self.force = save_velocities, save_types, save_forces
        if mode == "w" or (mode == "a" and not self.datasets):
            self.datasets = []
            N = self.N
            self._register("time", "f64", 1)
            self._register("R3position", "f64", 3 * N)
            self._register("barycentricPosition", "f64", 3 * N)
            self._register("faceIndex", "i32", N)
            if save_velocities:
                self._register("velocity", "f64", 3 * N)
            if save_forces:
                self._register("force", "f64", 3 * N)
            if save_types:
                self._register("type", "i32", N)
        else:
            names = {d[0] for d in self.datasets}
            self.velocity, self.force, self.type = "velocity" in names, "force" in names, "type" in names

    def current_number_of_records(self):
        return self.records("time")

    def write_state(self, time, r3, face, bary, vel=None, frc=None, types=None):
        N = self.N
        self.extend("time", [time])
        self.extend("R3position", np.asarray(r3).reshape(3 * N))
        self.extend("barycentricPosition", np.asarray(bary).reshape(3 * N))
        self.extend("faceIndex", face)
        if self.velocity:
            self.extend("velocity", np.zeros(3 * N) if vel is None else np.asarray(vel).reshape(3 * N))
        if self.force:
            self.extend("force", np.zeros(3 * N) if frc is None else np.asarray(frc).reshape(3 * N))
        if self.type:
            self.extend("type", np.zeros(N, np.int32) if types is None else types)

    def write_context(self, ctx, time):
        """Append the state of a binding.Context (all ranks hold all positions; velocities / forces are this rank's block)."""
        face, bary, vel, frc = ctx.get_state()
        if len(vel) != self.N:
            raise ValueError("write_context needs the unsharded state (gather velocities and forces first)")
        self.write_state(time, ctx.euclidean(face, bary), face, bary, vel, frc)

    def read_state(self, record=-1):
        N = self.N
        out = {"time": float(self.read("time", record)[0]), "R3position": self.read("R3position", record).reshape(N, 3),
               "barycentricPosition": self.read("barycentricPosition", record).reshape(N, 3), "faceIndex": self.read("faceIndex", record)}
        if self.velocity:
            out["velocity"] = self.read("velocity", record).reshape(N, 3)
        if self.force:
            out["force"] = self.read("force", record).reshape(N, 3)
        if self.type:
            out["type"] = self.read("type", record)
        return out


class ValueVectorDatabase(RawDatabase):
    """valueVectorDatabase (src/databases/vectorValueDatabase.cpp): (value, vector) records."""

    def __init__(self, path, vector_size, mode="r"):
        super().__init__(path, mode)
        self.vector_size = int(vector_size)
        if mode == "w" or (mode == "a" and not self.datasets):
            self.datasets = []
            self._register("value", "f64", 1)
            self._register("vector", "f64", self.vector_size)

    def current_number_of_records(self):
        return self.records("value")

    def write_state(self, value, vector):
        self.extend("vector", vector)
        self.extend("value", [value])

    def read_state(self, record=-1):
        return float(self.read("value", record)[0]), self.read("vector", record)
