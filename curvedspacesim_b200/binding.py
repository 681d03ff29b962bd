"""ctypes binding of the C ABI declared in include/css_api.h (libcurvedspacesim_b200.so).

This is plumbing only: every method forwards to one ``css_*`` entry point, which launches CUDA kernels.
There is no Python or CPU implementation behind it; if the library is missing or no B200 is visible the
calls fail loudly (RuntimeError)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CSS_LIB_PATH") or os.path.join(_HERE, "libcurvedspacesim_b200.so")  # env: developer variants only

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)

FORCE_HARMONIC, FORCE_GAUSSIAN = 0, 1
SUM, MAX = 0, 1

COUNTER_NAMES = ["walk_vertex", "walk_nohit", "walk_itercap", "walk_nan", "walk_border", "disconnected", "ties", "crossings",
                 "windows", "pseudo_sources", "patch_faces", "patch_verts", "queries", "sources", "tier_retry", "overflow",
                 "kernels", "kmax_overflow", "ovf_candidates", "ovf_faces", "ovf_verts", "ovf_ring", "clk_batch", "clk_fan", "clk_prop",
                 "clk_patch", "clk_total", "peer_timeout", "spilled"]
NUM_COUNTERS = 32

# every symbol include/css_api.h declares (tests check that the library exports all of them)
API_SYMBOLS = [
    "css_create", "css_destroy", "css_last_error", "css_set_mesh", "css_mesh_info", "css_set_submeshing", "css_set_cell_domain",
    "css_set_options", "css_set_boundary", "css_euclidean", "css_locate", "css_distance", "css_transport", "css_set_state", "css_get_state", "css_set_velocities",
    "css_set_forces", "css_find_neighbors", "css_get_neighbors", "css_compute_forces", "css_compute_energy", "css_compute_stress",
    "css_temperature", "css_move",
    "css_get_walk_flags", "css_step_nve", "css_step_nve_host", "css_step_gd", "css_nvt_init", "css_step_nvt", "css_nvt_state", "css_fire_init",
    "css_fire_minimize", "css_max_force", "css_force_norm", "css_comm_unique_id", "css_comm_init", "css_comm_info", "css_gather_positions",
    "css_reduce", "css_counters", "css_synchronize", "css_device_positions", "css_last_kernel_ms", "css_set_timing", "css_last_stage_ms",
    "css_timer_record", "css_timer_elapsed_ms", "css_microbench",
]

_lib = None


def _preload_nccl():
    """The library links libnccl.so.2.  In a Python process that also imports torch, the NCCL that gets
    mapped first wins for both (same soname), and torch needs its own bundled, newer build; so map that
    one first when it exists.  A C++ host simply uses the system libnccl."""
    import importlib.util

    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for loc in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
        p = os.path.join(loc, "lib", "libnccl.so.2")
        if os.path.exists(p):
            C.CDLL(p, mode=C.RTLD_GLOBAL)
            return


def load_library():
    """Load the CUDA library.  Raises if it has not been built (python -m curvedspacesim_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libcurvedspacesim_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        _preload_nccl()
        L = C.CDLL(LIB_PATH)
        L.css_last_error.restype = C.c_char_p
        L.css_last_error.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def force_params(kind: str, **kw):
    """('harmonic', k=, sigma=[, range=]) or ('gaussian', alpha=, sigma=[, range=]) -> (kind id, params[3])."""
    if kind == "harmonic":
        return FORCE_HARMONIC, np.array([kw.get("k", 1.0), kw["sigma"], kw.get("range", kw["sigma"])], dtype=np.float64)
    if kind == "gaussian":
        # force::maximumInteractionRange stays at the base default 1 unless set (baseForce.h:57)
        return FORCE_GAUSSIAN, np.array([kw.get("alpha", 1.0), kw["sigma"], kw.get("range", 1.0)], dtype=np.float64)
    raise ValueError(kind)


def _d(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(c_ip)


class CssError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("css error %d: %s" % (code, msg))
        self.code = code


class Context:
    """One GPU context = one mesh space + one (sharded) model state."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.css_create(C.byref(h), int(device))
        if rc:
            raise CssError(rc, "css_create failed (no usable CUDA device %d); this library has no CPU path" % device)
        self.h = h
        self.n_local = self.n_total = self.min_idx = 0
        self._M = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.css_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise CssError(rc, self.L.css_last_error(self.h).decode())

    # ---- space ----
    def set_mesh(self, V, corners):
        V = np.ascontiguousarray(V, np.float64)
        corners = np.ascontiguousarray(corners, np.int32)
        self._ck(self.L.css_set_mesh(self.h, len(V), _d(V), len(corners), _i(corners)))
        self.nV, self.nF = len(V), len(corners)

    def mesh_info(self):
        mn, mx, a = np.zeros(3), np.zeros(3), C.c_double()
        self._ck(self.L.css_mesh_info(self.h, _d(mn), _d(mx), C.byref(a)))
        return mn, mx, a.value

    def set_submeshing(self, enabled, max_dist=1.0):
        self._ck(self.L.css_set_submeshing(self.h, int(bool(enabled)), C.c_double(max_dist)))

    def set_cell_domain(self, mn, mx):
        mn = np.ascontiguousarray(mn, np.float64)
        mx = np.ascontiguousarray(mx, np.float64)
        self._ck(self.L.css_set_cell_domain(self.h, _d(mn), _d(mx)))

    def set_boundary(self, mode):
        """0 closed, 1 absorbing, 2 tangential (open-mesh variants of the walker)."""
        self._ck(self.L.css_set_boundary(self.h, int(mode)))

    def set_options(self, use_cell_list=True, want_end_tangents=False):
        self._ck(self.L.css_set_options(self.h, int(bool(use_cell_list)), int(bool(want_end_tangents))))

    def euclidean(self, face, bary):
        face = np.ascontiguousarray(face, np.int32)
        bary = np.ascontiguousarray(bary, np.float64)
        out = np.zeros((len(face), 3))
        self._ck(self.L.css_euclidean(self.h, len(face), _i(face), _d(bary), _d(out)))
        return out

    def locate(self, xyz, clamp_tol=1e-14):
        """simpleModel::R3PositionsToMeshPositions: closest mesh position (face, clamped barycentric weights) of points of R^3."""
        xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        face = np.zeros(len(xyz), np.int32)
        bary = np.zeros((len(xyz), 3))
        self._ck(self.L.css_locate(self.h, len(xyz), _d(xyz), C.c_double(clamp_tol), _i(face), _d(bary)))
        return face, bary

    def distance(self, src_face, src_bary, tgt_face, tgt_bary, threshold=1e20):
        sb = np.ascontiguousarray(src_bary, np.float64)
        tf = np.ascontiguousarray(tgt_face, np.int32)
        tb = np.ascontiguousarray(tgt_bary, np.float64)
        K = len(tf)
        dist, ts, te = np.zeros(K), np.zeros((K, 3)), np.zeros((K, 3))
        self._ck(self.L.css_distance(self.h, int(src_face), _d(sb), K, _i(tf), _d(tb), C.c_double(threshold), _d(dist), _d(ts), _d(te)))
        return dist, ts, te

    def transport(self, face, bary, disp, vecs=None):
        face = np.array(face, np.int32)
        bary = np.array(bary, np.float64)
        disp = np.array(disp, np.float64)
        n = len(face)
        vecs = np.zeros((n, 0, 3)) if vecs is None else np.array(vecs, np.float64).reshape(n, -1, 3)
        nvec = vecs.shape[1]
        flags = np.zeros(n, np.int32)
        self._ck(self.L.css_transport(self.h, n, _i(face), _d(bary), _d(disp), nvec, _d(vecs) if nvec else None, _i(flags)))
        return face, bary, disp, vecs, flags

    # ---- model ----
    def set_state(self, face, bary, vel=None, frc=None, n_local=None, min_idx=0):
        face = np.ascontiguousarray(face, np.int32)
        bary = np.ascontiguousarray(bary, np.float64)
        n_total = len(face)
        n_local = n_total if n_local is None else int(n_local)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float64)
        frc = None if frc is None else np.ascontiguousarray(frc, np.float64)
        self._ck(self.L.css_set_state(self.h, n_local, n_total, int(min_idx), _i(face), _d(bary), _d(vel), _d(frc)))
        self.n_local, self.n_total, self.min_idx = n_local, n_total, int(min_idx)

    def get_state(self):
        face = np.zeros(self.n_total, np.int32)
        bary = np.zeros((self.n_total, 3))
        vel = np.zeros((self.n_local, 3))
        frc = np.zeros((self.n_local, 3))
        self._ck(self.L.css_get_state(self.h, _i(face), _d(bary), _d(vel), _d(frc)))
        return face, bary, vel, frc

    def get_state_into(self, face=None, bary=None, vel=None, frc=None):
        """css_get_state straight into caller-owned (e.g. pinned) arrays; any of them may be None."""
        for a, n, dt in ((face, self.n_total, np.int32), (bary, 3 * self.n_total, np.float64), (vel, 3 * self.n_local, np.float64),
                         (frc, 3 * self.n_local, np.float64)):
            if a is not None and (a.dtype != dt or a.size != n or not a.flags["C_CONTIGUOUS"]):
                raise ValueError("get_state_into: wrong dtype / size / layout")
        self._ck(self.L.css_get_state(self.h, _i(face), _d(bary), _d(vel), _d(frc)))

    def set_velocities(self, vel):
        vel = np.ascontiguousarray(vel, np.float64)
        self._ck(self.L.css_set_velocities(self.h, _d(vel)))

    def set_forces(self, frc):
        frc = np.ascontiguousarray(frc, np.float64)
        self._ck(self.L.css_set_forces(self.h, _d(frc)))

    def find_neighbors(self, rng, want_end=False):
        tot = C.c_int64()
        self._ck(self.L.css_find_neighbors(self.h, C.c_double(rng), C.byref(tot)))
        n = tot.value
        off = np.zeros(self.n_local + 1, np.int32)
        idx = np.zeros(max(n, 1), np.int32)
        dist = np.zeros(max(n, 1))
        ts = np.zeros((max(n, 1), 3))
        te = np.zeros((max(n, 1), 3)) if want_end else None
        self._ck(self.L.css_get_neighbors(self.h, _i(off), _i(idx), _d(dist), _d(ts), _d(te)))
        return off, idx[:n], dist[:n], ts[:n], (None if te is None else te[:n])

    def compute_forces(self, kind, params, zero=True):
        self._ck(self.L.css_compute_forces(self.h, int(kind), _d(params), int(bool(zero))))

    def compute_energy(self, kind, params):
        e = C.c_double()
        self._ck(self.L.css_compute_energy(self.h, int(kind), _d(params), C.byref(e)))
        return e.value

    def compute_stress(self, kind, params):
        out = np.zeros(9)
        self._ck(self.L.css_compute_stress(self.h, int(kind), _d(params), _d(out)))
        return out.reshape(3, 3)

    def temperature(self):
        t = C.c_double()
        self._ck(self.L.css_temperature(self.h, C.byref(t)))
        return t.value

    def move(self, disp=None, transport_force=False, transport_velocity=True):
        d = None if disp is None else np.ascontiguousarray(disp, np.float64)
        self._ck(self.L.css_move(self.h, _d(d), int(bool(transport_force)), int(bool(transport_velocity))))

    def walk_flags(self):
        f = np.zeros(self.n_local, np.int32)
        self._ck(self.L.css_get_walk_flags(self.h, _i(f)))
        return f

    # ---- updaters ----
    def step_nve_host(self, kind, params, dt, face, bary, vel, frc):
        """One NVE step of a host-resident state, in place (upload, step, download with the position download overlapped)."""
        for a, n, dt_ in ((face, self.n_total, np.int32), (bary, 3 * self.n_total, np.float64), (vel, 3 * self.n_local, np.float64),
                          (frc, 3 * self.n_local, np.float64)):
            if a.dtype != dt_ or a.size != n or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("step_nve_host: wrong dtype / size / layout")
        self._ck(self.L.css_step_nve_host(self.h, int(kind), _d(params), C.c_double(dt), _i(face), _d(bary), _d(vel), _d(frc)))

    def step_nve(self, kind, params, dt, nsteps=1):
        self._ck(self.L.css_step_nve(self.h, int(kind), _d(params), C.c_double(dt), int(nsteps)))

    def step_gd(self, kind, params, dt, nsteps=1):
        self._ck(self.L.css_step_gd(self.h, int(kind), _d(params), C.c_double(dt), int(nsteps)))

    def nvt_init(self, dt, T, tau=1.0, M=2):
        self._ck(self.L.css_nvt_init(self.h, C.c_double(dt), C.c_double(T), C.c_double(tau), int(M)))
        self._M = M

    def step_nvt(self, kind, params, nsteps=1):
        self._ck(self.L.css_step_nvt(self.h, int(kind), _d(params), int(nsteps)))

    def nvt_state(self):
        bath = np.zeros((self._M + 1, 4))
        ke, sc = C.c_double(), C.c_double()
        self._ck(self.L.css_nvt_state(self.h, _d(bath), C.byref(ke), C.byref(sc)))
        return bath, ke.value, sc.value

    def fire_init(self, p=None, dt0=0.001, alpha0=0.99):
        pp = None if p is None else np.ascontiguousarray(p, np.float64)
        self._ck(self.L.css_fire_init(self.h, _d(pp), C.c_double(dt0), C.c_double(alpha0)))

    def fire_minimize(self, kind, params):
        out = np.zeros(4)
        self._ck(self.L.css_fire_minimize(self.h, int(kind), _d(params), _d(out)))
        return out

    def max_force(self):
        v = C.c_double()
        self._ck(self.L.css_max_force(self.h, C.byref(v)))
        return v.value

    def force_norm(self):
        v = C.c_double()
        self._ck(self.L.css_force_norm(self.h, C.byref(v)))
        return v.value

    # ---- multi-GPU ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = load_library().css_comm_unique_id(buf)
        if rc:
            raise CssError(rc, "ncclGetUniqueId failed")
        return buf.raw

    def comm_init(self, rank, nranks, uid: bytes | None):
        buf = C.create_string_buffer(uid, 128) if uid is not None else None
        self._ck(self.L.css_comm_init(self.h, int(rank), int(nranks), buf))

    def gather_positions(self):
        self._ck(self.L.css_gather_positions(self.h))

    def comm_info(self):
        """(rank, nranks, peer_exchange): peer_exchange is True when the exchange after every move runs over peer memory."""
        r, n, p = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.L.css_comm_info(self.h, C.byref(r), C.byref(n), C.byref(p)))
        return r.value, n.value, bool(p.value)

    def reduce(self, op, data):
        data = np.array(data, np.float64)
        self._ck(self.L.css_reduce(self.h, int(op), len(data), _d(data)))
        return data

    # ---- diagnostics ----
    def counters(self, reset=False):
        out = np.zeros(NUM_COUNTERS, np.uint64)
        self._ck(self.L.css_counters(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)), int(bool(reset))))
        return {n: int(out[i]) for i, n in enumerate(COUNTER_NAMES)}

    def synchronize(self):
        self._ck(self.L.css_synchronize(self.h))

    def set_timing(self, enabled=True):
        self._ck(self.L.css_set_timing(self.h, int(bool(enabled))))

    def last_kernel_ms(self):
        g, w, c = C.c_float(), C.c_float(), C.c_float()
        self._ck(self.L.css_last_kernel_ms(self.h, C.byref(g), C.byref(w), C.byref(c)))
        return {"geodesic_ms": g.value, "walk_ms": w.value, "celllist_ms": c.value}

    def last_stage_ms(self):
        p, w, r, g = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        self._ck(self.L.css_last_stage_ms(self.h, C.byref(p), C.byref(w), C.byref(r), C.byref(g)))
        return {"patch_ms": p.value, "window_ms": w.value, "retry_ms": r.value, "gather_ms": g.value}

    def microbench(self, what, reps=5):
        """Measured device ceiling: what = 0 FP64 FMA TFLOP/s, 1 L2 read GB/s."""
        v = C.c_double()
        self._ck(self.L.css_microbench(self.h, int(what), int(reps), C.byref(v)))
        return v.value

    def timer_record(self, slot):
        self._ck(self.L.css_timer_record(self.h, int(slot)))

    def timer_elapsed_ms(self, a, b):
        ms = C.c_float()
        self._ck(self.L.css_timer_elapsed_ms(self.h, int(a), int(b), C.byref(ms)))
        return ms.value
