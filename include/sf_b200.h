/*
 * sf_b200.h -- C-ABI of the B200-native SPH solver step (libsf_b200.so).
 *
 * Drop-in boundary for the simulation core of ttnghia/SimpleFluid.  The reference has no FFI
 * layer: its boundary is the C++ class surface `QtSPHSolver : SPHSolver<float>`
 * (Include/QtSPHSolver.h:27-36) as driven by `Simulator` (Source/Simulator.cpp:42,49,95-102) and
 * filled by `SceneManager` (Source/SceneManager.cpp:21-173), parameterised by
 * `SPHParameters<float>` (Source/Controller.cpp:52-64).  Each entry point below names the
 * reference interface it replaces; `EXE@0x...` are virtual addresses in the reference's shipped
 * binary Prebuild/SimpleFluid.exe (the solver source itself is not in the reference tree, see
 * SURVEY.md section 0 and Appendix A).
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every call returns 0 (SF_OK) or a
 * negative sf_status; nothing throws across the boundary; one host thread per solver at a time
 * (the reference drives its solver from the single std::async worker of Source/Simulator.cpp:29).
 * There is NO CPU fallback: every compute entry point fails with SF_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef SF_B200_H
#define SF_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum sf_status {
    SF_OK = 0,
    SF_ERR_INVALID = -1,   /* bad argument / call order */
    SF_ERR_CUDA = -2,      /* CUDA runtime or launch failure (see sf_last_error) */
    SF_ERR_DOMAIN = -3,    /* a particle lies outside [boxMin, boxMax] or is not finite */
    SF_ERR_OOM = -4,
    SF_ERR_COMM = -5,      /* NCCL failure */
    SF_ERR_STATE = -6      /* device-side consistency check failed (e.g. slab migration overflow) */
} sf_status;

/* Include/Common.h:52-58 */
enum { SF_SCENE_SPHERE_DROP = 0, SF_SCENE_CUBE_DROP = 1, SF_SCENE_DAMBREAK = 2, SF_SCENE_DOUBLE_DAMBREAK = 3 };

/* SPHParameters<float> / SimulationParameters.  Field names follow Source/Controller.cpp:54-63;
 * defaults and the derived block follow the constructor EXE@0x140011db0 and updateParams()
 * EXE@0x140006ac6. */
typedef struct sf_params {
    int32_t scene;                   /* Source/Simulator.cpp:80 */
    int32_t numThreads;              /* Controller.cpp:54; kept for API parity, ignored (CUDA grid launches replace TBB) */
    float   stopTime;                /* Controller.cpp:60, default 5.0 */
    float   defaultTimestep;         /* 1e-4; dt is clamped to [0.1, 10] x this */
    float   boxMin[3];               /* (-1,-1,-1) */
    float   boxMax[3];               /* ( 1, 1, 1) */
    float   pressureStiffness;       /* Controller.cpp:57, default 50000 */
    float   viscosity;               /* Controller.cpp:59, default 0.05 */
    float   kernelRadius;            /* Controller.cpp:55, h = 2/resolution */
    int32_t bCorrectDensity;         /* default 0 (no GUI control) */
    int32_t bUseBoundaryParticles;   /* default 1 */
    int32_t bUseAttractivePressure;  /* Controller.cpp:61, default 0 */
    float   boundaryRestitution;     /* Controller.cpp:58, default 0.1 */
    float   attractivePressureRatio; /* 0.1 */
    float   restDensity;             /* 1000 */
    /* derived by sf_params_update (= updateParams(), Controller.cpp:63) */
    float   particleMass;
    float   particleRadius;
    float   kernelRadiusSqr;
    float   restDensitySqr;
} sf_params;

typedef struct sf_solver sf_solver;

/* ---- parameters: SPHParameters ctor + updateParams() (Source/Controller.cpp:52-64) ---------- */
int sf_params_default(sf_params* p);
int sf_params_set_resolution(sf_params* p, float resolution);   /* kernelRadius = 2/res, then update */
int sf_params_update(sf_params* p);

/* ---- scenes: SceneManager::setupScene (Source/SceneManager.cpp:21-173) --------------------- */
/* Writes up to cap particles (xyz AoS, byte-compatible with Vec_Vec3<float>::data()) and the total
 * count to *n_out.  pos_xyz may be NULL to query the count.  Host-side, no GPU needed. */
int sf_scene_generate(const sf_params* p, int scene, float* pos_xyz, uint64_t cap, uint64_t* n_out);

/* ---- host-side setup pieces of makeReady(), exposed for inspection (no GPU needed) ------------- */
/* PrecomputedKernel<Cubic|Spiky,10000>::setRadius (EXE@0x14001a4e0, EXE@0x14001a2d0): the two tables the
 * step reads, 10001 floats each; consts3 = {W_zero, radius^2, invStep}. */
int sf_build_tables(const sf_params* p, float* cubic_w10001, float* spiky_grad10001, float* consts3);
/* generateBoundaryParticles (EXE@0x140016d80) for one wall with an explicit seed. */
int sf_boundary_generate(const sf_params* p, uint32_t seed, int wall, float* xyz, uint32_t cap, uint32_t* n_out);
/* Candidate masks of a wall list, as the density pass uses them (no counterpart in the reference, whose loop
 * over a wall list -- EXE@0x1400179c0, SURVEY A.6 -- tests all of its particles): masks[(SF_WALL_SUBCELLS + 1) * words],
 * words >= (n + 31) / 32; bit b of entry c marks wall particle b as possibly within h of sub-cell c of the h-cube a
 * shifted near-wall position lies in; entry SF_WALL_SUBCELLS marks every particle.  sf_wall_subcell: the entry a
 * position takes (same fp32 operations as the device).  Exposed so that the bound can be checked without a GPU. */
#define SF_WALL_SUBCELLS 64
int sf_wall_candidate_masks(const sf_params* p, int wall, const float* xyz, uint32_t n, uint32_t words, uint32_t* masks);
int sf_wall_subcell(const sf_params* p, int wall, const float pos_xyz[3], uint32_t* entry_out);

/* ---- lifecycle: QtSPHSolver ctor/dtor (Include/QtSPHSolver.h:30-31) ------------------------- */
int  sf_create(const sf_params* p, int device, sf_solver** out);
void sf_destroy(sf_solver* s);
int  sf_set_params(sf_solver* s, const sf_params* p);            /* Controller::updateSimParams */
int  sf_get_params(sf_solver* s, sf_params* p);
const char* sf_last_error(sf_solver* s);                          /* s may be NULL: last create error */
/* Launch on an externally owned CUDA stream (cudaStream_t as void*); NULL restores the solver's own. */
int  sf_set_stream(sf_solver* s, void* cuda_stream);

/* ---- particle buffers: getParticles()/getVelocity()/getNumParticles() (QtSPHSolver.h:33-35) - */
/* Positions/velocities are N x 3 fp32 AoS in ORIGINAL particle order (the renderer's colour VBOs
 * are per original index, Source/FluidRenderWidget.cpp:292-312).  vel_xyz may be NULL (zeros, as
 * SceneManager.cpp:63). */
int sf_upload_particles(sf_solver* s, const float* pos_xyz, const float* vel_xyz, uint32_t n);
int sf_num_particles(sf_solver* s, uint32_t* n_out);
int sf_download_positions(sf_solver* s, float* pos_xyz);
int sf_download_velocities(sf_solver* s, float* vel_xyz);

/* ---- wall boundary particles: generateBoundaryParticles EXE@0x140016d80 --------------------- */
/* The reference seeds std::mt19937 from std::random_device (not reproducible); here the seed is
 * explicit.  wall: 0 LX, 1 UX, 2 LY, 3 UY, 4 LZ, 5 UZ. */
int sf_generate_boundary(sf_solver* s, uint32_t seed);
int sf_set_boundary_particles(sf_solver* s, int wall, const float* xyz, uint32_t n);
int sf_get_boundary_particles(sf_solver* s, int wall, float* xyz, uint32_t cap, uint32_t* n_out);

/* ---- the step: makeReady() / advanceFrame() (Source/Simulator.cpp:42,49) -------------------- */
int sf_make_ready(sf_solver* s);                                  /* EXE@0x140016650 */
/* One reference substep (advanceFrame EXE@0x140016810); *dt_out = the dt it advanced by.
 * Synchronises (dt is read back), like the reference's blocking call. */
int sf_advance_frame(sf_solver* s, float* dt_out);
/* nsteps substeps enqueued without host synchronisation (dt stays on the device).  time_out may
 * be NULL; when given, the call synchronises and returns the simulated time advanced. */
int sf_advance_steps(sf_solver* s, uint32_t nsteps, float* time_out);
/* The inner loop of Simulator::doSimulation (Source/Simulator.cpp:46-51): substeps until the
 * accumulated frame time reaches frame_time (0.0333333333 there). */
int sf_advance_frame_time(sf_solver* s, double frame_time, float* time_out, uint32_t* nsteps_out);
int sf_synchronize(sf_solver* s);
/* Stateless host-buffer step (what bench.py's e2e leg times): upload pos/vel, one substep,
 * download pos/vel into the same buffers. */
int sf_step_host(sf_solver* s, float* pos_xyz, float* vel_xyz, uint32_t n, float* dt_out);
/* Page-locked host buffers for the calls above that take host pointers (sf_step_host, sf_upload_*, sf_download_*,
 * sf_snapshot_positions_async): with them the copies run at full PCIe rate and beside the kernels; pageable
 * memory works too, slower.  The reference keeps its arrays in std::vector (Include/QtSPHSolver.h:33-34). */
int sf_host_alloc(uint64_t bytes, void** out);
int sf_host_free(void* p);

/* ---- renderer hand-off and checkpoint/restart (SURVEY.md section 8 f) ----------------------------- */
/* Asynchronous snapshot of the positions (original order) into host_xyz (pinned memory for a truly asynchronous
 * copy): what FluidRenderWidget::updateParticleData uploads per particleChanged (Source/FluidRenderWidget.cpp:204-221).
 * The solver may keep stepping; host_xyz is valid after sf_snapshot_wait. */
int sf_snapshot_positions_async(sf_solver* s, float* host_xyz);
int sf_snapshot_wait(sf_solver* s);
/* {params, wall particles, simulated time, positions, velocities}; a restarted run continues bit-identically.
 * A slab run (sf_comm_init) writes one part per rank, `<path>.<rank>`: the particles that rank owns with their global
 * ids; every rank calls sf_checkpoint_write with the same path.  sf_checkpoint_read restores either form on one GPU;
 * sf_checkpoint_read_slab restores it as rank `rank` of `nranks` (any number of ranks, not necessarily the number that
 * wrote it: the cut planes are re-planned, results do not depend on them).  Counts in the file are validated against
 * its size; a truncated or inconsistent file gives SF_ERR_INVALID. */
int sf_checkpoint_write(sf_solver* s, const char* path, float sim_time);
int sf_checkpoint_read(const char* path, int device, sf_solver** out, float* sim_time);
int sf_checkpoint_read_slab(const char* path, int device, int rank, int nranks, const void* nccl_unique_id128, sf_solver** out, float* sim_time);

/* ---- parity / inspection fields (state of the LAST substep, original particle order) -------- */
typedef enum sf_field {
    SF_FIELD_DENSITY = 0,        /* float[n]      rho after computeDensity (A.8) */
    SF_FIELD_PRESSURE = 1,       /* float[n]      Pr(rho) of computePressureForces (A.11) */
    SF_FIELD_ACCEL = 2,          /* float[3n]     pressure acceleration (A.11); needs sf_set_capture(1) */
    SF_FIELD_CELL_INDEX = 3,     /* uint32[n]     (cz*ny+cy)*nx+cx as binned by collectParticlesToCells (A.7) */
    SF_FIELD_NEIGHBOR_COUNT = 4, /* uint32[n]     |{q != p : d2 <= h^2}| */
    SF_FIELD_NEIGHBOR_IDS = 5,   /* uint32[sum]   original ids, ascending per particle, concatenated in particle order */
    SF_FIELD_SORT_PERM = 6,      /* uint32[n]     original id of the particle at each sorted slot */
    SF_FIELD_TABLE_CUBIC_W = 7,  /* float[10001]  PrecomputedKernel<Cubic>  W table (A.2) */
    SF_FIELD_TABLE_SPIKY_GRAD = 8,/* float[10001] PrecomputedKernel<Spiky> gradW/r table (A.2) */
    /* The PRODUCTION neighbour list of the last substep -- the very arrays the density pass wrote and the pressure
     * and viscosity passes walked -- decoded to original ids (fields 4/5 above re-traverse the cells instead): */
    SF_FIELD_LIST_COUNTS = 9,    /* uint32[n]     packed count per particle: fluid (14 bit) | wall X (6) | wall Y (6) | wall Z (6);
                                                  0xffffffff = no list (capacity exceeded, the particle took the traversal path) */
    SF_FIELD_LIST_IDS = 10,      /* uint32[sum]   fluid neighbours in LIST order (= the reference's traversal order A.6:
                                                  cells z->y->x, ascending id inside a cell), concatenated in particle order */
    SF_FIELD_LIST_TABLE_INDEX = 11 /* uint32[sum] kernel-table index min(trunc(sqrt(d2)*invStep), 10000) per list entry (A.2) */
} sf_field;
int sf_set_capture(sf_solver* s, int on);                         /* extra per-step stores for ACCEL */
/* Entries per particle of the neighbour list the density pass builds (default 64 = 256 bytes per particle; a
 * rest-density particle has 28-35 fluid neighbours, the fullest of any developed BASELINE state 48; the run time does not
 * depend on it; rounded up to a multiple of four).  Particles with more take the cell-traversal path in all three
 * passes: slower, same bits.  Before the first upload only. */
int sf_set_list_capacity(sf_solver* s, int kmax);
int sf_field_size(sf_solver* s, int field, uint64_t* bytes_out);
int sf_download_field(sf_solver* s, int field, void* out, uint64_t bytes);
int sf_grid_dims(sf_solver* s, int32_t n3[3]);                     /* Grid3D::setGrid EXE@0x14001ab20 */

/* Diagnostics of the last substep: out = {substeps done, non-empty bricks, bricks whose halo did not fit the staging
 * buffers (cumulative), particles whose list overflowed (cumulative), max fluid neighbour count, sum of fluid neighbour
 * counts, particles without a list, particles resident on this rank}.  Synchronises. */
int sf_diagnostics(sf_solver* s, uint64_t out[8]);

/* Development counters of instrumented builds (make EXTRA=-DSF_EXP_WAITSTAT), cumulative: density consumer-warp cycles
 * waiting for a staged brick, refills counted, density consumer cycles in the exact phase, density consumer cycles in
 * total, producer cycles waiting for a free staging buffer, (unused x3).  Zeros in a normal build. */
int sf_debug_counters(sf_solver* s, uint64_t out[8]);

/* ---- measurement ---------------------------------------------------------------------------- */
/* Per-kernel CUDA-event timing on the launching stream.  sf_profile_enable(s, N): every N-th substep is timed
 * kernel by kernel (direct launches bracketed by events; the others keep the CUDA-graph replay), N = 1: every
 * substep, 0: off.  names: NUL-separated list written into buf; ms/launches arrays of length cap.  Returns the
 * number of kernels via *count_out. */
int sf_profile_enable(sf_solver* s, int on);
int sf_profile_reset(sf_solver* s);
int sf_profile_get(sf_solver* s, char* names_buf, size_t names_cap, double* ms, uint64_t* launches, uint32_t cap, uint32_t* count_out);
/* Number of kernel launches issued since creation (for bench.py's gpu_launches). */
int sf_launch_count(sf_solver* s, uint64_t* n_out);
/* Elapsed ms between two in-stream markers (CUDA events owned by the solver). */
int sf_timer_start(sf_solver* s);
int sf_timer_stop(sf_solver* s, float* ms_out);

/* ---- multi-GPU: z-slab decomposition, one solver (= one process, one GPU) per slab ---------- */
/* Nothing like this exists in the reference.  nccl_unique_id is the 128-byte
 * ncclUniqueId obtained from sf_comm_unique_id on rank 0 and distributed by the caller. */
int sf_comm_unique_id(void* id128);
int sf_comm_init(sf_solver* s, int rank, int nranks, const void* id128);
/* Scatter a global particle set (identical on every rank) into slabs: each rank keeps the particles
 * of its own z-range of cell layers, cut so that counts balance. */
int sf_upload_particles_global(sf_solver* s, const float* pos_xyz, const float* vel_xyz, uint32_t n_global);
int sf_slab_info(sf_solver* s, int32_t* z_begin, int32_t* z_end, uint32_t* n_owned, uint32_t* n_ghost);
/* Slow axis of the cell key = the axis the slabs are cut along: 2 = z (single GPU, and slab runs by default), 1 = y. */
int sf_slab_axis(sf_solver* s, int32_t* axis_out);
/* Owned particles with their global (original) ids, compacted on the device; any of ids / pos_xyz / vel_xyz may be
 * NULL, at most cap particles are written, *n_out = the owned count.  Order: unspecified (the ids say who is who). */
int sf_download_owned(sf_solver* s, uint32_t* ids, float* pos_xyz, float* vel_xyz, uint32_t cap, uint32_t* n_out);
/* Host-buffer step of a slab run (the multi-GPU counterpart of sf_step_host, what bench.py's e2e leg times at N > 1):
 * upload this rank's m_in OWNED particles {id, position, velocity} -- the resident copies are dropped, the ghost
 * particles of the coming substep stay on the device (they belong to the neighbours and arrived with the last
 * exchange) -- one substep incl. halo exchange and migration, download the particles this rank owns afterwards
 * (*m_out, at most cap written).  SF_ERR_DOMAIN when an uploaded particle lies outside the box or this rank's layers. */
int sf_step_host_owned(sf_solver* s, uint32_t* ids, float* pos_xyz, float* vel_xyz, uint32_t m_in, uint32_t cap, uint32_t* m_out, float* dt_out);
/* Raw resident state of this rank (float4 positions/velocities + ids, all slots) to / from host buffers. */
int sf_download_local(sf_solver* s, float* pos4, float* vel4, uint32_t* ids, uint32_t cap, uint32_t* n_out);
int sf_upload_local(sf_solver* s, const float* pos4, const float* vel4, const uint32_t* ids, uint32_t n);
/* Host-side pieces of the decomposition (no GPU needed): count-balanced cut planes (cuts[nranks+1]) from a
 * per-layer particle histogram; the one-layer-per-substep rebalancing rule applied to the all-gathered table
 * (8 uint32 per rank: sendLo, sendHi, nOwn, firstLayerCount, lastLayerCount, ...); global cell layer per particle. */
int sf_slab_plan(const uint64_t* layer_counts, int32_t nz, int32_t nranks, int32_t* cuts);
int sf_slab_rebalance(const uint32_t* table, int32_t nranks, int32_t nz, int32_t* cuts);
int sf_cell_layers(const sf_params* p, const float* pos_xyz, uint32_t n, int32_t* layers);

#ifdef __cplusplus
}
#endif
#endif
