"""Trajectory I/O in the reference's logical layout: the Python and C++ writers/readers agree on
the files, the datasets carry the reference's names, types and row widths, records append."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

from curvedspacesim_b200 import trajectory

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_python_round_trip_and_layout(tmp_path):
    N = 7
    rng = np.random.default_rng(3)
    db = trajectory.SimpleModelDatabase(N, str(tmp_path / "t.cssdb"), "w")
    # dataset names / types / widths of simpleModelDatabase::registerDatasets (simpleModelDatabase.cpp:37-50)
    assert db.datasets == [("time", "f64", 1), ("R3position", "f64", 3 * N), ("barycentricPosition", "f64", 3 * N), ("faceIndex", "i32", N),
                           ("velocity", "f64", 3 * N), ("force", "f64", 3 * N), ("type", "i32", N)]
    frames = []
    for rec in range(4):
        fr = dict(r3=rng.random((N, 3)), face=rng.integers(0, 100, N).astype(np.int32), bary=rng.random((N, 3)), vel=rng.random((N, 3)),
                  frc=rng.random((N, 3)), types=np.full(N, rec, np.int32))
        db.write_state(0.5 * rec, **fr)
        frames.append(fr)
    assert db.current_number_of_records() == 4
    rd = trajectory.SimpleModelDatabase(N, str(tmp_path / "t.cssdb"), "r")
    for rec in (0, 2, -1):
        s = rd.read_state(rec)
        fr = frames[rec]
        assert s["time"] == 0.5 * (rec % 4)
        assert np.array_equal(s["R3position"], fr["r3"]) and np.array_equal(s["barycentricPosition"], fr["bary"])
        assert np.array_equal(s["faceIndex"], fr["face"]) and np.array_equal(s["velocity"], fr["vel"])
        assert np.array_equal(s["force"], fr["frc"]) and np.array_equal(s["type"], fr["types"])
    assert rd.read("R3position").shape == (4, 3 * N)
    with pytest.raises(IOError):
        rd.write_state(9.0, **frames[0])
    with pytest.raises(IndexError):
        rd.read_state(4)
    # append mode keeps the records; optional datasets can be left out
    ap = trajectory.SimpleModelDatabase(N, str(tmp_path / "t.cssdb"), "a")
    ap.write_state(9.0, **frames[1])
    assert ap.current_number_of_records() == 5
    lean = trajectory.SimpleModelDatabase(N, str(tmp_path / "lean.cssdb"), "w", save_velocities=False, save_types=False, save_forces=False)
    lean.write_state(0.0, frames[0]["r3"], frames[0]["face"], frames[0]["bary"])
    assert [d[0] for d in lean.datasets] == ["time", "R3position", "barycentricPosition", "faceIndex"]
    with pytest.raises(ValueError):
        trajectory.SimpleModelDatabase(N + 1, str(tmp_path / "t.cssdb"), "r")
    # valueVectorDatabase (vectorValueDatabase.cpp:25-29)
    vv = trajectory.ValueVectorDatabase(str(tmp_path / "v.cssdb"), 3, "w")
    vv.write_state(1.5, [1, 2, 3])
    vv.write_state(2.5, [4, 5, 6])
    assert vv.datasets == [("value", "f64", 1), ("vector", "f64", 3)]
    val, vec = trajectory.ValueVectorDatabase(str(tmp_path / "v.cssdb"), 3, "r").read_state(-1)
    assert val == 2.5 and np.array_equal(vec, [4, 5, 6])


def test_cpp_and_python_databases_interoperate(tmp_path):
    exe = str(tmp_path / "db_roundtrip")
    subprocess.check_call([os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "db_roundtrip.cpp"), "-L" + os.path.join(ROOT, "curvedspacesim_b200"),
                           "-lcurvedspacesim_b200", "-Wl,-rpath," + os.path.join(ROOT, "curvedspacesim_b200")])
    N = 5
    subprocess.check_call([exe, str(tmp_path), "write"])
    db = trajectory.SimpleModelDatabase(N, str(tmp_path / "traj.cssdb"), "r")
    assert db.current_number_of_records() == 3 and db.N == N
    for rec in range(3):
        s = db.read_state(rec)
        assert s["time"] == 0.25 * rec
        for i in range(N):
            a, b = 0.1 * (i + 1) + 0.01 * rec, 0.05 * (i + 1)
            assert np.array_equal(s["barycentricPosition"][i], [1 - a - b, a, b]) and s["faceIndex"][i] == 10 * rec + i
            assert np.array_equal(s["R3position"][i], [a, b, 10 * rec + i])               # fillEuclideanLocations through the space
            assert np.array_equal(s["velocity"][i], [rec + 0.5, i, -1.25]) and np.array_equal(s["force"][i], [-rec, 2.0 * i, 1e-3])
            assert s["type"][i] == i % 2
    val, vec = trajectory.ValueVectorDatabase(str(tmp_path / "series.cssdb"), 4, "r").read_state(1)
    assert val == 8.5 and np.array_equal(vec, [-1, 2, 3, 4])
    # and the other way round
    py = trajectory.SimpleModelDatabase(N, str(tmp_path / "pytraj.cssdb"), "w")
    for rec, t in enumerate((1.5, 2.5)):
        py.write_state(t, np.zeros((N, 3)), 100 * rec + np.arange(N, dtype=np.int32), np.tile([0.5, 0.25, 0.25], (N, 1)),
                       np.stack([np.zeros(N), 3.0 * np.arange(N), np.zeros(N)], 1), np.tile([0.0, 0.0, -2.0], (N, 1)), np.full(N, 7, np.int32))
    subprocess.check_call([exe, str(tmp_path), "read"])
