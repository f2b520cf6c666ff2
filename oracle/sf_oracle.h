/*
 * sf_oracle.h -- CPU oracle for the SimpleFluid SPH step.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a restatement (plain C + OpenMP) of the reference's SPH substep.  The arithmetic of
 * that step is NOT in the reference source tree: it lives in the author's un-vendored, un-pinned
 * private library "Banana" (SimpleFluid.pro:27-30, Include/QtSPHSolver.h:22).  The restatement
 * therefore follows SURVEY.md Appendix A, which was reconstructed from the disassembly of
 * /root/reference/Prebuild/SimpleFluid.exe (virtual addresses are cited per function below),
 * plus Source/SceneManager.cpp:41-173 (scenes), Source/Simulator.cpp:44-56 (substep loop) and
 * Source/Controller.cpp:52-64 (parameters).
 *
 * PARITY STATUS: PINNED against outputs of the reference itself.  The reference ships no tests or golden vectors and
 * its solver source is absent, but its compiled solver is not: oracle/exe/sf_exe_harness.c maps the shipped binary
 * Prebuild/SimpleFluid.exe and calls its own SPHSolver::makeReady (EXE@0x140016650) and SPHSolver::advanceFrame
 * (EXE@0x140016810) natively (imports bound to libm / malloc / a single-threaded stand-in for the TBB scheduler).
 * This oracle reproduces that code's results BIT FOR BIT -- kernel tables, wall particles (same seed), dt, cell
 * indices, density, acceleration, positions, velocities; every scene, the Shepard / attractive-pressure / no-wall
 * variants, an odd grid, random moving states, 1000 substeps of the reference default: tests/test_oracle_vs_exe.py
 * (live runs in the build container + the committed fixtures tests/golden/exe_*.npz everywhere).  Also: (1) the scene
 * generator is checked against the reference's own Source/SceneManager.cpp compiled unmodified (oracle/_ref, see
 * oracle/Makefile) and the particle counts in the reference's screenshots (Captured/1.png..3.png); (2) the analytic
 * known answers of SURVEY.md section 8c (tests/test_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (simplefluid_b200/, include/) never does.
 */
#ifndef SF_ORACLE_H
#define SF_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* SimulationParameters, binary layout EXE@0x140011db0; names per
 * Source/Controller.cpp:54-63 where the source names them. */
typedef struct sfo_params {
    int32_t scene;               /* Include/Common.h:52-58: SphereDrop 0, CubeDrop 1, Dambreak 2, DoubleDambreak 3 */
    int32_t numThreads;          /* 0 = automatic (Source/Controller.cpp:54) */
    float   stopTime;            /* 5.0 */
    float   defaultTimestep;     /* 1e-4 */
    float   boxMin[3];           /* -1 */
    float   boxMax[3];           /* +1 */
    float   pressureStiffness;   /* 50000 */
    float   viscosity;           /* 0.05 */
    float   kernelRadius;        /* h = 2/resolution (Source/Controller.cpp:55) */
    int32_t bCorrectDensity;     /* false */
    int32_t bUseBoundaryParticles; /* true */
    int32_t bUseAttractivePressure; /* false */
    float   boundaryRestitution; /* 0.1 */
    float   attractivePressureRatio; /* 0.1 */
    float   restDensity;         /* 1000 */
    /* derived by updateParams() */
    float   particleMass;
    float   particleRadius;
    float   kernelRadiusSqr;
    float   restDensitySqr;
} sfo_params;

typedef struct sfo_solver sfo_solver;

void sfo_params_default(sfo_params* p);                 /* ctor EXE@0x140011db0 */
void sfo_params_set_resolution(sfo_params* p, float resolution); /* Controller.cpp:55 + updateParams */
void sfo_params_update(sfo_params* p);                  /* updateParams EXE@0x140006ac6 */

/* SceneManager::setupScene (Source/SceneManager.cpp:21-173).  Returns particle count; writes at
 * most cap particles (xyz AoS) if pos != NULL. */
uint64_t sfo_scene_generate(const sfo_params* p, int scene, float* pos_xyz, uint64_t cap);

sfo_solver* sfo_create(const sfo_params* p);
void        sfo_destroy(sfo_solver* s);
void        sfo_set_threads(sfo_solver* s, int nthreads);       /* 0 = OpenMP default */
void        sfo_set_traversal(sfo_solver* s, int reversed);     /* 1: reversed neighbour order (self-divergence study) */
int         sfo_set_particles(sfo_solver* s, const float* pos_xyz, const float* vel_xyz, uint32_t n);
void        sfo_generate_boundary(sfo_solver* s, uint32_t seed); /* EXE@0x140016d80, seeded */
int         sfo_set_boundary(sfo_solver* s, int wall, const float* xyz, uint32_t n);
uint32_t    sfo_get_boundary(sfo_solver* s, int wall, float* xyz, uint32_t cap);
void        sfo_make_ready(sfo_solver* s);                       /* EXE@0x140016650 */
float       sfo_advance_frame(sfo_solver* s);                    /* EXE@0x140016810, one substep, returns dt */
/* wall-clock seconds spent in the last advance, split: [0] dt/maxvel [1] collect [2] density
 * [3] gravity+pressure+updateVelocity [4] viscosity [5] updatePosition */
void        sfo_last_timing(sfo_solver* s, double* t6);

uint32_t     sfo_num_particles(sfo_solver* s);
const float* sfo_positions(sfo_solver* s);
const float* sfo_velocities(sfo_solver* s);
const float* sfo_density(sfo_solver* s);
const float* sfo_accel(sfo_solver* s);
void         sfo_pressure(sfo_solver* s, float* out);            /* Pr(rho) of A.11 per particle */
void         sfo_grid_dims(sfo_solver* s, int32_t* n3);
void         sfo_cell_index(sfo_solver* s, uint32_t* out);       /* cell of each particle as binned by the last step (A.7) */
/* neighbour sets of the last step's binning: counts[n], ids (ascending per particle) up to cap
 * entries; returns total number of neighbour entries */
uint64_t     sfo_neighbors(sfo_solver* s, uint32_t* counts, uint32_t* ids, uint64_t cap);
/* kernel tables (A.2): which 0 = cubic W[10001], 1 = spiky gradW/r [10001] */
void         sfo_table(sfo_solver* s, int which, float* out10001);
void         sfo_kernel_consts(sfo_solver* s, float* out4);      /* W_zero, radius2, invStep, spiky radius2 */

#ifdef __cplusplus
}
#endif
#endif
