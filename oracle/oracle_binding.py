"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: imported by
tests/ (tests/conftest.py puts this directory on sys.path), __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repository root (this file lives in oracle/)
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_long)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_last_error.restype = C.c_char_p
        L.orc_candidates.restype = C.c_long
        L.orc_find_neighbors.restype = C.c_long
        for n in ("orc_run_nve", "orc_run_gd", "orc_run_nvt", "orc_run_fire", "orc_compute_energy", "orc_temperature"):
            getattr(L, n).restype = C.c_double
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(c_dp)


def _i(a):
    return a.ctypes.data_as(c_ip)


def force_params(kind, **kw):
    if kind == "harmonic":
        return 0, np.array([kw.get("k", 1.0), kw["sigma"], kw.get("range", kw["sigma"])], dtype=np.float64)
    if kind == "gaussian":
        return 1, np.array([kw.get("alpha", 1.0), kw["sigma"], kw.get("range", 1.0)], dtype=np.float64)
    raise ValueError(kind)


class Oracle:
    def __init__(self, V, corners):
        self.V = np.ascontiguousarray(V, dtype=np.float64)
        self.corners = np.ascontiguousarray(corners, dtype=np.int32)
        self.nV, self.nF = len(self.V), len(self.corners)
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create(self.nV, _d(self.V), self.nF, _i(self.corners)))
        if not self.h:
            raise RuntimeError(self.L.orc_last_error().decode())
        self.N = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def mesh_info(self):
        mn, mx, a = np.zeros(3), np.zeros(3), C.c_double()
        self.L.orc_mesh_info(self.h, _d(mn), _d(mx), C.byref(a))
        return mn, mx, a.value

    def adjacency(self):
        adj = np.zeros((self.nF, 3), np.int32)
        adjk = np.zeros((self.nF, 3), np.int32)
        self.L.orc_get_adjacency(self.h, _i(adj), _i(adjk))
        return adj, adjk

    def saddle(self):
        out = np.zeros(self.nV, np.int8)
        self.L.orc_get_saddle(self.h, out.ctypes.data_as(C.c_char_p))
        return out

    def set_submeshing(self, enabled, max_dist=1.0):
        self.L.orc_set_submeshing(self.h, int(enabled), C.c_double(max_dist))

    def set_options(self, use_cell_list=True, strict_trig=False, threads=1):
        self.L.orc_set_options(self.h, int(use_cell_list), int(strict_trig), int(threads))

    def set_boundary(self, mode):
        """0 closed (triangulatedMeshSpace), 1 absorbing, 2 tangential (openMeshSpace variants)."""
        self.L.orc_set_boundary(self.h, int(mode))

    def set_cell_domain(self, mn, mx):
        mn = np.ascontiguousarray(mn, np.float64)
        mx = np.ascontiguousarray(mx, np.float64)
        self.L.orc_set_cell_domain(self.h, _d(mn), _d(mx))

    def set_state(self, face, bary, vel=None, frc=None):
        face = np.ascontiguousarray(face, np.int32)
        bary = np.ascontiguousarray(bary, np.float64)
        self.N = len(face)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float64)
        frc = None if frc is None else np.ascontiguousarray(frc, np.float64)
        self.L.orc_set_state(self.h, self.N, _i(face), _d(bary), _d(vel) if vel is not None else None,
                             _d(frc) if frc is not None else None)

    def get_state(self):
        face = np.zeros(self.N, np.int32)
        bary = np.zeros((self.N, 3))
        vel = np.zeros((self.N, 3))
        frc = np.zeros((self.N, 3))
        self.L.orc_get_state(self.h, _i(face), _d(bary), _d(vel), _d(frc))
        return face, bary, vel, frc

    def euclidean(self, face, bary):
        face = np.ascontiguousarray(face, np.int32)
        bary = np.ascontiguousarray(bary, np.float64)
        out = np.zeros((len(face), 3))
        self.L.orc_euclidean(self.h, len(face), _i(face), _d(bary), _d(out))
        return out

    def locate(self, xyz, clamp_tol=1e-14):
        """simpleModel::R3PositionsToMeshPositions (brute force over the faces)."""
        xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        face = np.zeros(len(xyz), np.int32)
        bary = np.zeros((len(xyz), 3))
        self.L.orc_locate(self.h, len(xyz), _d(xyz), C.c_double(clamp_tol), _i(face), _d(bary))
        return face, bary

    def candidates(self, rng):
        off = np.zeros(self.N + 1, np.int32)
        maxd = np.zeros(self.N)
        tot = self.L.orc_candidates(self.h, C.c_double(rng), _i(off), None, C.c_long(0), _d(maxd))
        idx = np.zeros(max(tot, 1), np.int32)
        self.L.orc_candidates(self.h, C.c_double(rng), _i(off), _i(idx), C.c_long(tot), _d(maxd))
        return off, idx[:tot], maxd

    def cell_grid(self, rng):
        n = np.zeros(3, np.int32)
        cs = np.zeros(3)
        self.L.orc_cell_grid(self.h, C.c_double(rng), _i(n), _d(cs))
        return n, cs

    def patch(self, sf, sb, tfaces, R):
        sb = np.ascontiguousarray(sb, np.float64)
        tf = np.ascontiguousarray(tfaces, np.int32)
        cap = self.nF
        out = np.zeros(cap, np.int32)
        n = self.L.orc_patch(self.h, int(sf), _d(sb), len(tf), _i(tf), C.c_double(R), _i(out), cap)
        return out[:n]

    def distance(self, sf, sb, tf, tb, threshold=1e20):
        sb = np.ascontiguousarray(sb, np.float64)
        tf = np.ascontiguousarray(tf, np.int32)
        tb = np.ascontiguousarray(tb, np.float64)
        K = len(tf)
        dist = np.zeros(K)
        ts = np.zeros((K, 3))
        te = np.zeros((K, 3))
        tie = np.zeros(K, np.int32)
        st = np.zeros(5, np.int64)
        self.L.orc_distance(self.h, int(sf), _d(sb), K, _i(tf), _d(tb), C.c_double(threshold), _d(dist), _d(ts), _d(te),
                            _i(tie), st.ctypes.data_as(c_lp))
        return dist, ts, te, tie, st

    def transport(self, face, bary, disp, vecs=None):
        face = np.array(face, np.int32)
        bary = np.array(bary, np.float64)
        disp = np.array(disp, np.float64)
        n = len(face)
        if vecs is None:
            vecs = np.zeros((n, 0, 3))
        vecs = np.array(vecs, np.float64).reshape(n, -1, 3)
        nvec = vecs.shape[1]
        flags = np.zeros(n, np.int32)
        cr = np.zeros(n, np.int32)
        self.L.orc_transport(self.h, n, _i(face), _d(bary), _d(disp), nvec, _d(vecs) if nvec else None, _i(flags), _i(cr))
        return face, bary, disp, vecs, flags, cr

    def find_neighbors(self, rng):
        tot = self.L.orc_find_neighbors(self.h, C.c_double(rng))
        off = np.zeros(self.N + 1, np.int32)
        idx = np.zeros(max(tot, 1), np.int32)
        dist = np.zeros(max(tot, 1))
        ts = np.zeros((max(tot, 1), 3))
        te = np.zeros((max(tot, 1), 3))
        self.L.orc_get_neighbors(self.h, _i(off), _i(idx), _d(dist), _d(ts), _d(te))
        return off, idx[:tot], dist[:tot], ts[:tot], te[:tot]

    def compute_forces(self, kind, params, zero=True):
        self.L.orc_compute_forces(self.h, kind, _d(params), int(zero))
        return self.get_state()[3]

    def compute_energy(self, kind, params):
        return self.L.orc_compute_energy(self.h, kind, _d(params))

    def compute_stress(self, kind, params):
        out = np.zeros(9)
        self.L.orc_compute_stress(self.h, kind, _d(params), _d(out))
        return out.reshape(3, 3)

    def temperature(self):
        return self.L.orc_temperature(self.h)

    def move(self, disp, transport_force=False, transport_velocity=True):
        disp = np.array(disp, np.float64)
        self.L.orc_move(self.h, _d(disp), int(transport_force), int(transport_velocity))
        return disp

    def walk_flags(self):
        f = np.zeros(self.N, np.int32)
        self.L.orc_get_walk_flags(self.h, _i(f))
        return f

    def run_nve(self, kind, params, dt, steps):
        return self.L.orc_run_nve(self.h, kind, _d(params), C.c_double(dt), int(steps))

    def run_gd(self, kind, params, dt, steps):
        return self.L.orc_run_gd(self.h, kind, _d(params), C.c_double(dt), int(steps))

    def nvt_init(self, dt, T, tau=1.0, M=2):
        self.L.orc_nvt_init(self.h, C.c_double(dt), C.c_double(T), C.c_double(tau), int(M))
        self._M = M

    def run_nvt(self, kind, params, steps):
        return self.L.orc_run_nvt(self.h, kind, _d(params), int(steps))

    def nvt_state(self):
        bath = np.zeros((self._M + 1, 4))
        ke, sc = C.c_double(), C.c_double()
        self.L.orc_nvt_state(self.h, _d(bath), C.byref(ke), C.byref(sc))
        return bath, ke.value, sc.value

    def fire_init(self, p=None, dt0=0.001, alpha0=0.99):
        pp = None if p is None else np.ascontiguousarray(p, np.float64)
        self.L.orc_fire_init(self.h, _d(pp) if pp is not None else None, C.c_double(dt0), C.c_double(alpha0))

    def run_fire(self, kind, params):
        out = np.zeros(4)
        t = self.L.orc_run_fire(self.h, kind, _d(params), _d(out))
        return t, out

    def counters(self, reset=False):
        out = np.zeros(13, np.int64)
        self.L.orc_counters(self.h, out.ctypes.data_as(c_lp), int(reset))
        names = ["vertex", "nohit", "itercap", "nan", "border", "disconnected", "ties", "crossings", "windows_created",
                 "windows_processed", "pseudo_sources", "patch_faces", "patch_verts"]
        return dict(zip(names, out.tolist()))
