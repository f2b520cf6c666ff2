"""ctypes binding of the CPU oracle (oracle/libsf_oracle.so) and of the reference's own scene
generator (oracle/_ref/libsf_refscene.so).  TEST INFRASTRUCTURE: imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libsf_oracle.so")
REFSCENE_SO = os.path.join(ORACLE_DIR, "_ref", "libsf_refscene.so")

SCENES = {"SphereDrop": 0, "CubeDrop": 1, "Dambreak": 2, "DoubleDambreak": 3}


class OracleParams(C.Structure):
    _fields_ = [
        ("scene", C.c_int32), ("numThreads", C.c_int32), ("stopTime", C.c_float), ("defaultTimestep", C.c_float),
        ("boxMin", C.c_float * 3), ("boxMax", C.c_float * 3),
        ("pressureStiffness", C.c_float), ("viscosity", C.c_float), ("kernelRadius", C.c_float),
        ("bCorrectDensity", C.c_int32), ("bUseBoundaryParticles", C.c_int32), ("bUseAttractivePressure", C.c_int32),
        ("boundaryRestitution", C.c_float), ("attractivePressureRatio", C.c_float), ("restDensity", C.c_float),
        ("particleMass", C.c_float), ("particleRadius", C.c_float), ("kernelRadiusSqr", C.c_float),
        ("restDensitySqr", C.c_float),
    ]


def build_oracle(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists).  Building the checker is
    not using it."""
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "sf_oracle.c")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "--no-print-directory"], stdout=subprocess.DEVNULL)
    elif os.path.isdir("/root/reference") and not os.path.exists(REFSCENE_SO):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "--no-print-directory", "ref"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        vp, u32, u64, f32p, u32p = C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        PP = C.POINTER(OracleParams)
        L.sfo_params_default.argtypes = [PP]
        L.sfo_params_set_resolution.argtypes = [PP, C.c_float]
        L.sfo_params_update.argtypes = [PP]
        L.sfo_scene_generate.argtypes = [PP, C.c_int, vp, u64]
        L.sfo_scene_generate.restype = u64
        L.sfo_create.argtypes = [PP]
        L.sfo_create.restype = vp
        L.sfo_destroy.argtypes = [vp]
        L.sfo_set_threads.argtypes = [vp, C.c_int]
        L.sfo_set_traversal.argtypes = [vp, C.c_int]
        L.sfo_set_particles.argtypes = [vp, vp, vp, u32]
        L.sfo_generate_boundary.argtypes = [vp, u32]
        L.sfo_set_boundary.argtypes = [vp, C.c_int, vp, u32]
        L.sfo_get_boundary.argtypes = [vp, C.c_int, vp, u32]
        L.sfo_get_boundary.restype = u32
        L.sfo_make_ready.argtypes = [vp]
        L.sfo_advance_frame.argtypes = [vp]
        L.sfo_advance_frame.restype = C.c_float
        L.sfo_last_timing.argtypes = [vp, C.POINTER(C.c_double)]
        L.sfo_num_particles.argtypes = [vp]
        L.sfo_num_particles.restype = u32
        for name in ("sfo_positions", "sfo_velocities", "sfo_density", "sfo_accel"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = f32p
        L.sfo_pressure.argtypes = [vp, vp]
        L.sfo_grid_dims.argtypes = [vp, C.POINTER(C.c_int32)]
        L.sfo_cell_index.argtypes = [vp, vp]
        L.sfo_neighbors.argtypes = [vp, vp, vp, u64]
        L.sfo_neighbors.restype = u64
        L.sfo_table.argtypes = [vp, C.c_int, vp]
        L.sfo_kernel_consts.argtypes = [vp, vp]
        _lib = L
    return _lib


def default_params(resolution=24.0, scene="Dambreak", **overrides):
    p = OracleParams()
    lib().sfo_params_default(C.byref(p))
    p.scene = SCENES[scene] if isinstance(scene, str) else int(scene)
    lib().sfo_params_set_resolution(C.byref(p), C.c_float(resolution))
    for k, v in overrides.items():
        setattr(p, k, v)
    lib().sfo_params_update(C.byref(p))
    return p


def scene(params, scene_id=None):
    sid = params.scene if scene_id is None else (SCENES[scene_id] if isinstance(scene_id, str) else scene_id)
    n = lib().sfo_scene_generate(C.byref(params), sid, None, 0)
    out = np.empty((n, 3), np.float32)
    lib().sfo_scene_generate(C.byref(params), sid, out.ctypes.data, n)
    return out


def ref_scene(particle_radius, scene_id):
    """The reference's own Source/SceneManager.cpp (compiled unmodified into oracle/_ref)."""
    if not os.path.exists(REFSCENE_SO):
        build_oracle()
    if not os.path.exists(REFSCENE_SO):
        return None
    L = C.CDLL(REFSCENE_SO)
    L.ref_scene_generate.argtypes = [C.c_float, C.c_int, C.c_void_p, C.c_uint64]
    L.ref_scene_generate.restype = C.c_uint64
    sid = SCENES[scene_id] if isinstance(scene_id, str) else scene_id
    n = L.ref_scene_generate(C.c_float(particle_radius), sid, None, 0)
    out = np.empty((n, 3), np.float32)
    L.ref_scene_generate(C.c_float(particle_radius), sid, out.ctypes.data, n)
    return out


class Oracle:
    """Thin object wrapper: mirrors the solver-facing surface (makeReady / advanceFrame / getters)."""

    def __init__(self, params, pos, vel=None, boundary_seed=0, threads=0, reversed_traversal=False):
        self.L = lib()
        self.params = params
        self.h = self.L.sfo_create(C.byref(params))
        pos = np.ascontiguousarray(pos, np.float32)
        self.n = pos.shape[0]
        velp = None if vel is None else np.ascontiguousarray(vel, np.float32).ctypes.data
        self.L.sfo_set_particles(self.h, pos.ctypes.data, velp, self.n)
        self.L.sfo_set_threads(self.h, threads)
        self.L.sfo_set_traversal(self.h, 1 if reversed_traversal else 0)
        if boundary_seed is not None:
            self.L.sfo_generate_boundary(self.h, boundary_seed)
        self.L.sfo_make_ready(self.h)

    def close(self):
        if self.h:
            self.L.sfo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def advance(self):
        return float(self.L.sfo_advance_frame(self.h))

    def timing(self):
        t = (C.c_double * 6)()
        self.L.sfo_last_timing(self.h, t)
        return list(t)

    def _arr(self, fn, cols):
        ptr = fn(self.h)
        shape = (self.n, cols) if cols > 1 else (self.n,)
        return np.ctypeslib.as_array(ptr, shape=shape).copy()

    def positions(self):
        return self._arr(self.L.sfo_positions, 3)

    def velocities(self):
        return self._arr(self.L.sfo_velocities, 3)

    def density(self):
        return self._arr(self.L.sfo_density, 1)

    def accel(self):
        return self._arr(self.L.sfo_accel, 3)

    def pressure(self):
        out = np.empty(self.n, np.float32)
        self.L.sfo_pressure(self.h, out.ctypes.data)
        return out

    def grid_dims(self):
        d = (C.c_int32 * 3)()
        self.L.sfo_grid_dims(self.h, d)
        return tuple(d)

    def cell_index(self):
        out = np.empty(self.n, np.uint32)
        self.L.sfo_cell_index(self.h, out.ctypes.data)
        return out

    def neighbors(self):
        """(counts[n], ids[total]) for the CURRENT positions; ids ascending per particle."""
        counts = np.empty(self.n, np.uint32)
        total = self.L.sfo_neighbors(self.h, counts.ctypes.data, None, 0)
        ids = np.empty(total, np.uint32)
        self.L.sfo_neighbors(self.h, counts.ctypes.data, ids.ctypes.data, total)
        return counts, ids

    def boundary(self, wall):
        n = self.L.sfo_get_boundary(self.h, wall, None, 0)
        out = np.empty((n, 3), np.float32)
        self.L.sfo_get_boundary(self.h, wall, out.ctypes.data, n)
        return out

    def table(self, which):
        out = np.empty(10001, np.float32)
        self.L.sfo_table(self.h, which, out.ctypes.data)
        return out

    def kernel_consts(self):
        out = np.empty(4, np.float32)
        self.L.sfo_kernel_consts(self.h, out.ctypes.data)
        return out
