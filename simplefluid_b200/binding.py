"""ctypes binding of include/sf_b200.h (libsf_b200.so).  Host-side mirror of the reference's
solver-facing interface: `SPHSolver` carries the method names of QtSPHSolver
(/root/reference/Include/QtSPHSolver.h:27-36) and SPHSolver<float> (Source/Simulator.cpp:42,49)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_LIB_PATH = os.environ.get("SF_B200_LIB") or os.path.join(_HERE, "lib", "libsf_b200.so")  # SF_B200_LIB: kernel experiments (tools/exp_bench.py)

SCENES = {"SphereDrop": 0, "CubeDrop": 1, "Dambreak": 2, "DoubleDambreak": 3}  # Include/Common.h:52-58

FIELD_DENSITY, FIELD_PRESSURE, FIELD_ACCEL, FIELD_CELL_INDEX, FIELD_NEIGHBOR_COUNT, FIELD_NEIGHBOR_IDS, \
    FIELD_SORT_PERM, FIELD_TABLE_CUBIC_W, FIELD_TABLE_SPIKY_GRAD, FIELD_LIST_COUNTS, FIELD_LIST_IDS, FIELD_LIST_TABLE_INDEX = range(12)

EXPORTS = [
    "sf_params_default", "sf_params_set_resolution", "sf_params_update", "sf_scene_generate", "sf_build_tables",
    "sf_boundary_generate", "sf_wall_candidate_masks", "sf_wall_subcell", "sf_create", "sf_destroy",
    "sf_set_params", "sf_get_params", "sf_last_error", "sf_set_stream", "sf_upload_particles", "sf_num_particles",
    "sf_download_positions", "sf_download_velocities", "sf_generate_boundary", "sf_set_boundary_particles",
    "sf_get_boundary_particles", "sf_make_ready", "sf_advance_frame", "sf_advance_steps", "sf_advance_frame_time",
    "sf_synchronize", "sf_step_host", "sf_set_capture", "sf_field_size", "sf_download_field", "sf_grid_dims",
    "sf_profile_enable", "sf_profile_reset", "sf_profile_get", "sf_launch_count", "sf_timer_start", "sf_timer_stop",
    "sf_comm_unique_id", "sf_comm_init", "sf_upload_particles_global", "sf_slab_info", "sf_download_owned",
    "sf_slab_plan", "sf_slab_rebalance", "sf_cell_layers", "sf_download_local", "sf_upload_local",
    "sf_snapshot_positions_async", "sf_snapshot_wait", "sf_checkpoint_write", "sf_checkpoint_read",
    "sf_host_alloc", "sf_host_free", "sf_diagnostics", "sf_set_list_capacity",
    "sf_slab_axis", "sf_step_host_owned", "sf_checkpoint_read_slab", "sf_debug_counters",
]


class SFError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sf_b200 error {code}: {msg}")
        self.code = code


class SFParams(C.Structure):
    """struct sf_params (= SPHParameters<float>, Source/Controller.cpp:54-63)."""
    _fields_ = [
        ("scene", C.c_int32), ("numThreads", C.c_int32), ("stopTime", C.c_float), ("defaultTimestep", C.c_float),
        ("boxMin", C.c_float * 3), ("boxMax", C.c_float * 3),
        ("pressureStiffness", C.c_float), ("viscosity", C.c_float), ("kernelRadius", C.c_float),
        ("bCorrectDensity", C.c_int32), ("bUseBoundaryParticles", C.c_int32), ("bUseAttractivePressure", C.c_int32),
        ("boundaryRestitution", C.c_float), ("attractivePressureRatio", C.c_float), ("restDensity", C.c_float),
        ("particleMass", C.c_float), ("particleRadius", C.c_float), ("kernelRadiusSqr", C.c_float),
        ("restDensitySqr", C.c_float),
    ]

    def updateParams(self):
        library().sf_params_update(C.byref(self))


def library_path():
    return _LIB_PATH


def build_library(force=False, verbose=False):
    """Compile libsf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    args = ["make", "-C", _CSRC, "--no-print-directory", "all"]
    if force:
        args.append("-B")
    r = subprocess.run(args, capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("building libsf_b200.so failed")
    return _LIB_PATH


_lib = None


def library():
    """Load the C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise SFError(-2, f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(_LIB_PATH)
    vp, i32, u32, u64, f32 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_float
    PP = C.POINTER(SFParams)
    sig = {
        "sf_params_default": [PP], "sf_params_set_resolution": [PP, f32], "sf_params_update": [PP],
        "sf_scene_generate": [PP, C.c_int, vp, u64, C.POINTER(u64)],
        "sf_build_tables": [PP, vp, vp, vp], "sf_boundary_generate": [PP, u32, C.c_int, vp, u32, C.POINTER(u32)],
        "sf_wall_candidate_masks": [PP, C.c_int, vp, u32, u32, vp], "sf_wall_subcell": [PP, C.c_int, vp, C.POINTER(u32)],
        "sf_create": [PP, C.c_int, C.POINTER(vp)], "sf_set_params": [vp, PP], "sf_get_params": [vp, PP],
        "sf_set_stream": [vp, vp], "sf_upload_particles": [vp, vp, vp, u32], "sf_num_particles": [vp, C.POINTER(u32)],
        "sf_download_positions": [vp, vp], "sf_download_velocities": [vp, vp], "sf_generate_boundary": [vp, u32],
        "sf_set_boundary_particles": [vp, C.c_int, vp, u32],
        "sf_get_boundary_particles": [vp, C.c_int, vp, u32, C.POINTER(u32)],
        "sf_make_ready": [vp], "sf_advance_frame": [vp, C.POINTER(f32)], "sf_advance_steps": [vp, u32, C.POINTER(f32)],
        "sf_advance_frame_time": [vp, C.c_double, C.POINTER(f32), C.POINTER(u32)], "sf_synchronize": [vp],
        "sf_step_host": [vp, vp, vp, u32, C.POINTER(f32)], "sf_set_capture": [vp, C.c_int],
        "sf_field_size": [vp, C.c_int, C.POINTER(u64)], "sf_download_field": [vp, C.c_int, vp, u64],
        "sf_grid_dims": [vp, C.POINTER(i32)], "sf_profile_enable": [vp, C.c_int], "sf_profile_reset": [vp],
        "sf_profile_get": [vp, vp, C.c_size_t, vp, vp, u32, C.POINTER(u32)], "sf_launch_count": [vp, C.POINTER(u64)],
        "sf_timer_start": [vp], "sf_timer_stop": [vp, C.POINTER(f32)],
        "sf_comm_unique_id": [vp], "sf_comm_init": [vp, C.c_int, C.c_int, vp],
        "sf_upload_particles_global": [vp, vp, vp, u32],
        "sf_slab_info": [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(u32), C.POINTER(u32)],
        "sf_download_owned": [vp, vp, vp, vp, u32, C.POINTER(u32)],
        "sf_download_local": [vp, vp, vp, vp, u32, C.POINTER(u32)], "sf_upload_local": [vp, vp, vp, vp, u32],
        "sf_snapshot_positions_async": [vp, vp], "sf_snapshot_wait": [vp],
        "sf_checkpoint_write": [vp, C.c_char_p, f32], "sf_checkpoint_read": [C.c_char_p, C.c_int, C.POINTER(vp), C.POINTER(f32)],
        "sf_host_alloc": [u64, C.POINTER(vp)], "sf_host_free": [vp],
        "sf_diagnostics": [vp, vp], "sf_set_list_capacity": [vp, C.c_int],
        "sf_slab_axis": [vp, C.POINTER(i32)], "sf_debug_counters": [vp, vp],
        "sf_step_host_owned": [vp, vp, vp, vp, u32, u32, C.POINTER(u32), C.POINTER(f32)],
        "sf_checkpoint_read_slab": [C.c_char_p, C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp), C.POINTER(f32)],
        "sf_slab_plan": [vp, i32, i32, vp], "sf_slab_rebalance": [vp, i32, i32, vp], "sf_cell_layers": [PP, vp, u32, vp],
    }
    for name, argtypes in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    L.sf_destroy.argtypes = [vp]
    L.sf_destroy.restype = None
    L.sf_last_error.argtypes = [vp]
    L.sf_last_error.restype = C.c_char_p
    _lib = L
    return L


class PinnedArray:
    """numpy view of a page-locked host buffer from sf_host_alloc (freed on close() / garbage collection)."""

    def __init__(self, shape, dtype=np.float32):
        self.L = library()
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = C.c_void_p()
        rc = self.L.sf_host_alloc(max(nbytes, 1), C.byref(self.ptr))
        if rc:
            raise SFError(rc, (self.L.sf_last_error(None) or b"").decode())
        buf = (C.c_char * max(nbytes, 1)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self):
        if self.ptr is not None and self.ptr.value:
            self.array = None
            self.L.sf_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_params(resolution=24.0, scene="Dambreak", **overrides):
    """SPHParameters defaults + Controller::updateSimParams (Source/Controller.cpp:52-64)."""
    p = SFParams()
    L = library()
    L.sf_params_default(C.byref(p))
    p.scene = SCENES[scene] if isinstance(scene, str) else int(scene)
    L.sf_params_set_resolution(C.byref(p), C.c_float(resolution))
    for k, v in overrides.items():
        setattr(p, k, v)
    L.sf_params_update(C.byref(p))
    return p


def scene_generate(params, scene=None):
    """SceneManager::setupScene (Source/SceneManager.cpp:21-38): (N,3) float32 positions; velocities are zero."""
    L = library()
    sid = params.scene if scene is None else (SCENES[scene] if isinstance(scene, str) else int(scene))
    n = C.c_uint64(0)
    L.sf_scene_generate(C.byref(params), sid, None, 0, C.byref(n))
    out = np.empty((n.value, 3), np.float32)
    L.sf_scene_generate(C.byref(params), sid, out.ctypes.data, n.value, C.byref(n))
    return out


def build_tables(params):
    """(cubicW[10001], spikyGrad[10001], (W_zero, radius2, invStep)) as makeReady() builds them."""
    w, g, c = np.empty(10001, np.float32), np.empty(10001, np.float32), np.empty(3, np.float32)
    library().sf_build_tables(C.byref(params), w.ctypes.data, g.ctypes.data, c.ctypes.data)
    return w, g, c


def boundary_generate(params, seed, wall):
    n = C.c_uint32(0)
    library().sf_boundary_generate(C.byref(params), seed, wall, None, 0, C.byref(n))
    out = np.empty((n.value, 3), np.float32)
    library().sf_boundary_generate(C.byref(params), seed, wall, out.ctypes.data, n.value, C.byref(n))
    return out


WALL_SUBCELLS = 64


def wall_candidate_masks(params, wall, xyz):
    """Candidate masks of a wall list as the density pass uses them: (WALL_SUBCELLS + 1, words) uint32."""
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    words = max((len(xyz) + 31) // 32, 1)
    out = np.zeros((WALL_SUBCELLS + 1, words), np.uint32)
    rc = library().sf_wall_candidate_masks(C.byref(params), wall, xyz.ctypes.data, len(xyz), words, out.ctypes.data)
    if rc:
        raise SFError(rc, "sf_wall_candidate_masks: invalid arguments")
    return out


def wall_subcell(params, wall, pos):
    pos = np.ascontiguousarray(pos, np.float32).reshape(3)
    e = C.c_uint32(0)
    rc = library().sf_wall_subcell(C.byref(params), wall, pos.ctypes.data, C.byref(e))
    if rc:
        raise SFError(rc, "sf_wall_subcell: invalid arguments")
    return e.value


def comm_unique_id():
    """128-byte ncclUniqueId (rank 0 creates it, the caller distributes it to the other ranks)."""
    buf = C.create_string_buffer(128)
    rc = library().sf_comm_unique_id(buf)
    if rc:
        raise SFError(rc, (library().sf_last_error(None) or b"").decode())
    return buf.raw


def slab_plan(layer_counts, nranks):
    lc = np.ascontiguousarray(layer_counts, np.uint64)
    cuts = np.zeros(nranks + 1, np.int32)
    rc = library().sf_slab_plan(lc.ctypes.data, len(lc), nranks, cuts.ctypes.data)
    if rc:
        raise SFError(rc, "sf_slab_plan: invalid arguments (too few layers for this many slabs?)")
    return cuts


def slab_rebalance(table, nz, cuts):
    t = np.ascontiguousarray(table, np.uint32)
    c = np.ascontiguousarray(cuts, np.int32).copy()
    rc = library().sf_slab_rebalance(t.ctypes.data, len(c) - 1, nz, c.ctypes.data)
    if rc:
        raise SFError(rc, "sf_slab_rebalance: invalid arguments")
    return c


def cell_layers(params, pos):
    pos = np.ascontiguousarray(pos, np.float32)
    out = np.empty(len(pos), np.int32)
    library().sf_cell_layers(C.byref(params), pos.ctypes.data, len(pos), out.ctypes.data)
    return out


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor (pinned host memory in bench.py)
        return a.data_ptr()
    return a


class SPHSolver:
    """Mirror of QtSPHSolver : SPHSolver<float>.  Method names follow the reference."""

    def __init__(self, params, device=0):
        self.L = library()
        self.params = params
        h = C.c_void_p()
        rc = self.L.sf_create(C.byref(params), device, C.byref(h))
        if rc:
            raise SFError(rc, (self.L.sf_last_error(None) or b"").decode())
        self.h = h

    # -- plumbing
    def _ck(self, rc):
        if rc:
            raise SFError(rc, (self.L.sf_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.sf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- reference surface
    def setParticles(self, pos, vel=None):
        """What SceneManager::setupScene does to getParticles()/getVelocity() (Source/Simulator.cpp:95)."""
        pos = np.ascontiguousarray(pos, np.float32)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32)
        self._ck(self.L.sf_upload_particles(self.h, _ptr(pos), _ptr(vel), pos.shape[0]))

    def setupScene(self, scene=None):
        pos = scene_generate(self.params, scene)
        self.setParticles(pos)
        return pos

    def getNumParticles(self):
        n = C.c_uint32(0)
        self._ck(self.L.sf_num_particles(self.h, C.byref(n)))
        return n.value

    def getParticles(self):
        out = np.empty((self.getNumParticles(), 3), np.float32)
        self._ck(self.L.sf_download_positions(self.h, out.ctypes.data))
        return out

    def getVelocity(self):
        out = np.empty((self.getNumParticles(), 3), np.float32)
        self._ck(self.L.sf_download_velocities(self.h, out.ctypes.data))
        return out

    def generateBoundaryParticles(self, seed=0):
        self._ck(self.L.sf_generate_boundary(self.h, seed))

    def setBoundaryParticles(self, wall, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32)
        self._ck(self.L.sf_set_boundary_particles(self.h, wall, xyz.ctypes.data, xyz.shape[0]))

    def getBoundaryParticles(self, wall):
        n = C.c_uint32(0)
        self._ck(self.L.sf_get_boundary_particles(self.h, wall, None, 0, C.byref(n)))
        out = np.empty((n.value, 3), np.float32)
        self._ck(self.L.sf_get_boundary_particles(self.h, wall, out.ctypes.data, n.value, C.byref(n)))
        return out

    def makeReady(self):
        self._ck(self.L.sf_make_ready(self.h))

    def advanceFrame(self):
        """One substep; returns the dt advanced (Source/Simulator.cpp:49)."""
        dt = C.c_float(0)
        self._ck(self.L.sf_advance_frame(self.h, C.byref(dt)))
        return dt.value

    # -- extensions of the C-ABI
    def advanceSteps(self, n, want_time=False):
        if want_time:
            t = C.c_float(0)
            self._ck(self.L.sf_advance_steps(self.h, n, C.byref(t)))
            return t.value
        self._ck(self.L.sf_advance_steps(self.h, n, None))
        return None

    def advanceFrameTime(self, frame_time=0.0333333333):
        t, k = C.c_float(0), C.c_uint32(0)
        self._ck(self.L.sf_advance_frame_time(self.h, frame_time, C.byref(t), C.byref(k)))
        return t.value, k.value

    def synchronize(self):
        self._ck(self.L.sf_synchronize(self.h))

    def stepHost(self, pos, vel):
        dt = C.c_float(0)
        n = pos.shape[0]
        self._ck(self.L.sf_step_host(self.h, _ptr(pos), _ptr(vel), n, C.byref(dt)))
        return dt.value

    def setStream(self, cuda_stream_ptr):
        self._ck(self.L.sf_set_stream(self.h, cuda_stream_ptr))

    def setCapture(self, on=True):
        self._ck(self.L.sf_set_capture(self.h, 1 if on else 0))

    def gridDims(self):
        d = (C.c_int32 * 3)()
        self._ck(self.L.sf_grid_dims(self.h, d))
        return tuple(d)

    def field(self, field):
        nbytes = C.c_uint64(0)
        self._ck(self.L.sf_field_size(self.h, field, C.byref(nbytes)))
        dtype = np.uint32 if field in (FIELD_CELL_INDEX, FIELD_NEIGHBOR_COUNT, FIELD_NEIGHBOR_IDS, FIELD_SORT_PERM, FIELD_LIST_COUNTS,
                                       FIELD_LIST_IDS, FIELD_LIST_TABLE_INDEX) else np.float32
        out = np.empty(nbytes.value // 4, dtype)
        if nbytes.value:
            self._ck(self.L.sf_download_field(self.h, field, out.ctypes.data, nbytes.value))
        if field == FIELD_ACCEL:
            out = out.reshape(-1, 3)
        return out

    def density(self):
        return self.field(FIELD_DENSITY)

    def pressure(self):
        return self.field(FIELD_PRESSURE)

    def accel(self):
        return self.field(FIELD_ACCEL)

    def cellIndex(self):
        return self.field(FIELD_CELL_INDEX)

    def neighbors(self):
        return self.field(FIELD_NEIGHBOR_COUNT), self.field(FIELD_NEIGHBOR_IDS)

    def productionLists(self):
        """The neighbour list the density pass built and the force / viscosity passes walked, decoded:
        (fluid counts[n] with -1 where the particle had no list, packed raw counts[n], ids[sum] in list order,
        table indices[sum])."""
        raw = self.field(FIELD_LIST_COUNTS)
        nolist = raw == 0xFFFFFFFF
        cnt = np.where(nolist, np.int64(-1), (raw & 16383).astype(np.int64))
        return cnt, raw, self.field(FIELD_LIST_IDS), self.field(FIELD_LIST_TABLE_INDEX)

    def setListCapacity(self, kmax):
        self._ck(self.L.sf_set_list_capacity(self.h, int(kmax)))

    def debugCounters(self):
        out = (C.c_uint64 * 8)()
        self._ck(self.L.sf_debug_counters(self.h, out))
        return [int(x) for x in out]

    def diagnostics(self):
        out = (C.c_uint64 * 8)()
        self._ck(self.L.sf_diagnostics(self.h, out))
        keys = ("substeps", "bricks", "fallback_bricks", "fallback_particles", "nbr_max", "nbr_sum", "particles_without_list", "particles")
        d = dict(zip(keys, (int(x) for x in out)))
        d["nbr_mean"] = d["nbr_sum"] / max(d["particles"], 1)
        return d

    # -- renderer hand-off / checkpoint
    def snapshotPositionsAsync(self, host_xyz):
        self._ck(self.L.sf_snapshot_positions_async(self.h, _ptr(host_xyz)))

    def snapshotWait(self):
        self._ck(self.L.sf_snapshot_wait(self.h))

    def checkpointWrite(self, path, sim_time=0.0):
        self._ck(self.L.sf_checkpoint_write(self.h, str(path).encode(), sim_time))

    @classmethod
    def fromCheckpoint(cls, path, device=0):
        L = library()
        h, t = C.c_void_p(), C.c_float(0)
        rc = L.sf_checkpoint_read(str(path).encode(), device, C.byref(h), C.byref(t))
        if rc:
            raise SFError(rc, (L.sf_last_error(None) or b"").decode())
        self = cls.__new__(cls)
        self.L, self.h = L, h
        self.params = SFParams()
        L.sf_get_params(h, C.byref(self.params))
        return self, t.value

    # -- multi-GPU (z-slabs; one SPHSolver per rank / GPU)
    def commInit(self, rank, nranks, unique_id_bytes):
        buf = C.create_string_buffer(bytes(unique_id_bytes), 128)
        self._ck(self.L.sf_comm_init(self.h, rank, nranks, buf))

    def setParticlesGlobal(self, pos, vel=None):
        pos = np.ascontiguousarray(pos, np.float32)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32)
        self._ck(self.L.sf_upload_particles_global(self.h, _ptr(pos), _ptr(vel), pos.shape[0]))

    def slabInfo(self):
        zb, ze, no, ng = C.c_int32(0), C.c_int32(0), C.c_uint32(0), C.c_uint32(0)
        self._ck(self.L.sf_slab_info(self.h, C.byref(zb), C.byref(ze), C.byref(no), C.byref(ng)))
        return zb.value, ze.value, no.value, ng.value

    def downloadOwned(self):
        n = C.c_uint32(0)
        self._ck(self.L.sf_download_owned(self.h, None, None, None, 0, C.byref(n)))
        ids = np.empty(n.value, np.uint32)
        x = np.empty((n.value, 3), np.float32)
        v = np.empty((n.value, 3), np.float32)
        self._ck(self.L.sf_download_owned(self.h, ids.ctypes.data, x.ctypes.data, v.ctypes.data, n.value, C.byref(n)))
        return ids, x, v

    def downloadOwnedInto(self, ids, pos, vel):
        """Owned particles into caller-provided (ideally pinned) buffers; returns the count."""
        n = C.c_uint32(0)
        self._ck(self.L.sf_download_owned(self.h, _ptr(ids), _ptr(pos), _ptr(vel), ids.shape[0], C.byref(n)))
        if n.value > ids.shape[0]:
            raise SFError(-1, f"buffers hold {ids.shape[0]} particles, the rank owns {n.value}")
        return n.value

    def stepHostOwned(self, ids, pos, vel, m):
        """sf_step_host_owned: upload the m owned particles in the buffers, one substep, download the new owned set
        into the same buffers; returns its size."""
        out, dt = C.c_uint32(0), C.c_float(0)
        self._ck(self.L.sf_step_host_owned(self.h, _ptr(ids), _ptr(pos), _ptr(vel), int(m), ids.shape[0], C.byref(out), C.byref(dt)))
        if out.value > ids.shape[0]:
            raise SFError(-1, f"buffers hold {ids.shape[0]} particles, the rank owns {out.value}")
        self.last_dt = dt.value
        return out.value

    def slabAxis(self):
        a = C.c_int32(2)
        self._ck(self.L.sf_slab_axis(self.h, C.byref(a)))
        return a.value

    @classmethod
    def fromCheckpointSlab(cls, path, device, rank, nranks, unique_id_bytes):
        L = library()
        h, t = C.c_void_p(), C.c_float(0)
        buf = C.create_string_buffer(bytes(unique_id_bytes), 128) if unique_id_bytes is not None else None
        rc = L.sf_checkpoint_read_slab(str(path).encode(), device, rank, nranks, buf, C.byref(h), C.byref(t))
        if rc:
            raise SFError(rc, (L.sf_last_error(None) or b"").decode())
        self = cls.__new__(cls)
        self.L, self.h = L, h
        self.params = SFParams()
        L.sf_get_params(h, C.byref(self.params))
        return self, t.value

    def localSlots(self):
        n = C.c_uint32(0)
        self._ck(self.L.sf_download_local(self.h, None, None, None, 0, C.byref(n)))
        return n.value

    def downloadLocal(self, pos4, vel4, ids):
        n = C.c_uint32(0)
        self._ck(self.L.sf_download_local(self.h, _ptr(pos4), _ptr(vel4), _ptr(ids), ids.shape[0], C.byref(n)))
        return n.value

    def uploadLocal(self, pos4, vel4, ids, n):
        self._ck(self.L.sf_upload_local(self.h, _ptr(pos4), _ptr(vel4), _ptr(ids), n))

    # -- measurement
    def profileEnable(self, on=True, every=1):
        """Per-kernel event timing of every `every`-th substep (the others replay the CUDA graph)."""
        self._ck(self.L.sf_profile_enable(self.h, int(every) if on else 0))

    def profileReset(self):
        self._ck(self.L.sf_profile_reset(self.h))

    def profile(self):
        cap = 64
        names = C.create_string_buffer(4096)
        ms = (C.c_double * cap)()
        launches = (C.c_uint64 * cap)()
        count = C.c_uint32(0)
        self._ck(self.L.sf_profile_get(self.h, names, 4096, ms, launches, cap, C.byref(count)))
        parts = names.raw.split(b"\0")
        return {parts[i].decode(): (ms[i], launches[i]) for i in range(count.value)}

    def launchCount(self):
        n = C.c_uint64(0)
        self._ck(self.L.sf_launch_count(self.h, C.byref(n)))
        return n.value

    def timerStart(self):
        self._ck(self.L.sf_timer_start(self.h))

    def timerStop(self):
        ms = C.c_float(0)
        self._ck(self.L.sf_timer_stop(self.h, C.byref(ms)))
        return ms.value
