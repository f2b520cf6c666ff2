// sf_device.cuh -- device-side data model of the B200 SPH step.
//
// HBM layout (all SoA, one entry per particle slot, fp32 / uint32):
//   posA/velA/idA : state at the start of a substep, in the sorted order of the PREVIOUS substep
//                   (float4 {x,y,z,-}, float4 {vx,vy,vz,-}, original particle id)
//   keys/vals[2]  : radix-sort ping-pong (cell key, slot in A)
//   posB/velB/idB : state re-sorted by cell key for THIS substep.  The .w lanes carry the
//                   per-particle terms the pair loops need so that every neighbour costs one
//                   128-bit load:  posB.w = P(rho)/rho^2 (pressure term, written by the density pass together
//                   with x, y, z as one 16-byte element), velB.w = 1/rho (written by the force pass with v*)
//   keyB          : cell key per sorted slot ((cz*ny+cy)*nx+cx, reference A.7)
//   cellTab       : uint2 {begin,end} sorted-slot range per cell, {0,0} when empty
//   cellCnt       : particles per cell, counted for the next binning (by k_hash_count, or by k_visc_brick as it writes
//                   the new positions); the scan that turns it into cellTab zeroes it again
//   rho           : density per sorted slot
//   brickList     : ids of the non-empty cell bricks (BX x BY x BZ cells) of this substep, in z-major
//                   order; the three pair kernels are persistent CTAs that pull bricks from it
//   nbrL          : neighbour list built once per substep by the density pass and reused by the
//                   force and viscosity passes: [slot/32][k][slot%32], so a warp reads/writes 128 B
//                   rows and the rows of 32 slots are one contiguous block; entry = brick-local halo index of the neighbour (16 bit,
//                   or wall-particle index) | kernel-table index min(trunc(sqrt(d2)*invStep), 10000)
//                   << 16, the index being shared by the cubic-W and spiky-grad tables
//   nbrCnt        : packed counts: fluid (14 bit) | wallX (6) | wallY (6) | wallZ (6);
//                   0xffffffff = no list (capacity exceeded): the later passes re-traverse the cells
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sf
{
struct DevParams {
    float bmin[3], bmax[3];
    float h, h2, r, mass, stiffness, viscosity, restitution, rho0, attractRatio;
    float rhoMin, rhoMax; // f((double)rho0*0.1), f((double)rho0*10.0)
    float Wzero, invStep, radius2;
    float dtMin, dtMax;   // defaultTimestep*0.1f, defaultTimestep*10.0f
    int   nx, ny, nz;
    int   useBoundary, attractive, correctDensity, capture;
    uint32_t n, npad;
    int      kmax;
    uint32_t kmaxBytes;  // kmax * 128: bytes from the first to one past the last row of a list column
    int      nbx, nby, nbz; // brick grid
    uint32_t numBricks;
    // z-slab decomposition (single GPU: z0 = 0, nzGlobal = nz, every range = [0, nz)).  nz above is the LOCAL
    // number of cell layers; local layer l is global layer l + z0 (z0 may be negative at the bottom slab).
    int      z0, nzGlobal;
    int      zDensLo, zDensHi;   // local layers whose particles get a density / list
    int      zShepLo, zShepHi;   // ... a Shepard-corrected density (bCorrectDensity only)
    int      zForceLo, zForceHi; // ... a pressure force and v*
    int      zOwnLo, zOwnHi;     // ... are integrated (owned by this rank)
    int      zEdge;              // own layers within zEdge of a slab face are "boundary" for the overlapped exchange
    int      slab;               // 1: slab mode
    // Key order is (slow * ny + mid) * nx + x.  The slow axis is z (axisS = 2, the reference's own order) or, for
    // slab runs whose particles span more layers in y, y (axisS = 1); then ny/nz above hold the MID / SLOW axis
    // cell counts and every "z"/"layer" of the slab logic means the slow axis.  The 3x3 rows of a neighbourhood
    // are always visited in the reference's order (dz outer, dy inner).
    int      axisS;
    uint32_t nbnd[6];     // wall particle counts
    uint32_t bndStride;   // slots per wall in the bnd array
    uint32_t wallWords;   // words per candidate mask of a wall list (wallMask)
    float    wallSubInv;  // kWallSub / h: sub-cell index of a shifted near-wall position (host: wall_sub_inv)
};

struct DevState {
    unsigned maxv2Bits[2]; // max |v|^2 as float bits, ping-pong by step parity
    float    dt;
    unsigned step;
    float    frameTime;   // accumulated like `frameTime += advanceFrame()` (Simulator.cpp:49)
    unsigned skip;        // 1: frame target reached, the kernels of this substep are no-ops
    double   frameTarget; // 0: no target
    unsigned errFlags;    // SF_DEVERR_*
    unsigned nbrMax;      // largest neighbour count seen (diagnostics)
    unsigned long long stepsDone;
    unsigned brickCount;  // non-empty bricks of this substep
    unsigned cursor[6];   // work cursors of the persistent pair kernels (density, force, viscosity edge / interior, Shepard)
    unsigned fallbackBricks, fallbackParticles; // diagnostics: halos that did not fit smem / lists that overflowed
    // SF_EXP_WAITSTAT builds: density consumer-warp cycles waiting for a brick / refills counted (producers of all pair
    // kernels) / density consumer cycles in the exact phase / density consumer cycles in total / producer cycles waiting
    // for a free staging buffer / (unused x3)
    unsigned long long dbg[8];
};

enum : unsigned { SF_DEVERR_NBR_OVERFLOW = 1u, SF_DEVERR_WALL_OVERFLOW = 2u, SF_DEVERR_DOMAIN = 4u };

struct DevBuffers {
    float4 *posA, *velA, *posB, *velB;
    uint32_t *idA, *idB;
    uint32_t *keys[2], *vals[2];
    uint32_t *keyB;       // aliases the sorted keys buffer
    uint2*    cellTab;
    uint32_t* cellCnt;    // counting sort: particles per cell of the NEXT binning; all zero between a scan and the next count
    float*    rho;
    float*    rho2;       // correctDensity scratch
    float4*   accel;      // capture only
    uint32_t* nbrL;
    uint32_t* nbrCnt;
    uint32_t* brickFlag;
    uint32_t* brickList;
    float *tabW, *tabG;   // kTableEntries each
    float4*   bnd;        // [6][bndStride]
    uint32_t* wallMask;   // [6][kWallSubCells + 1][wallWords]: wall particles that can be within h of a sub-cell (sf_host.cpp)
    uint32_t* radixCounts;
    uint32_t* radixTotals;
    DevState* state;
};
} // namespace sf
