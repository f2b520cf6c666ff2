// sf_internal.h -- shared declarations of the B200 SPH solver (host setup <-> device pipeline).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/sf_b200.h"

namespace sf
{
constexpr int kTableSize = 10000;             // PrecomputedKernel<*, 10000>
constexpr int kTableEntries = kTableSize + 1; // index 10000 reads 0 (W[10000] aliases gradW[0])

// Precomputed kernel tables, built on the host exactly once per makeReady (A.2, A.3).
struct KernelTables {
    std::vector<float> cubicW;    // kTableEntries
    std::vector<float> spikyGrad; // kTableEntries, gradW_x(X)/X
    float Wzero = 0.f;            // cubic W_zero
    float radius2 = 0.f;          // h*h (identical for both kernels)
    float invStep = 0.f;          // f(1/(double)(h/10000))
};

// ---- host-side setup (sf_host.cpp) ------------------------------------------------------------
void params_default(sf_params& p);
void params_update(sf_params& p);
uint64_t scene_generate(const sf_params& p, int scene, float* out_xyz, uint64_t cap);
void build_tables(float h, KernelTables& t);
void grid_dims(const sf_params& p, int32_t n[3]);
// unclamped cell coordinates of a position (A.6); returns false if outside the grid / not finite
bool cell_coords_checked(const sf_params& p, const int32_t n[3], const float* x, int32_t c[3]);
void generate_boundary(const sf_params& p, uint32_t seed, std::vector<float> walls[6]);
// Candidate masks of a wall list for the density pass (see sf_host.cpp): [kWallSubCells + 1][words] words, bit b = wall
// particle b; wall_subcell is the host mirror of the device's sub-cell index (same fp32 operations)
constexpr int kWallSub = 4, kWallSubCells = kWallSub * kWallSub * kWallSub;
float wall_sub_inv(const sf_params& p);
int   wall_subcell(const sf_params& p, int wall, const float x[3]);
void  wall_candidate_masks(const sf_params& p, int wall, const float* xyz, uint32_t n, uint32_t words, uint32_t* masks);
// global cell layer (A.7 z index, clamped) of a z coordinate -- same float ops as the device hash
int32_t cell_layer(const sf_params& p, int32_t n, float coord, int axis = 2);
// z-slab decomposition (no counterpart in the reference): count-balanced cut planes with a minimum thickness, and
// the one-layer-per-substep rebalancing rule every rank evaluates identically from the all-gathered table
// (row r = {sendLo, sendHi, nOwn, firstLayerCount, lastLayerCount, ...}, rowWords words per row)
void slab_plan(const uint64_t* layerCounts, int32_t nz, int32_t nranks, int32_t minThick, int32_t* cuts);
void slab_rebalance(const uint32_t* table, int32_t rowWords, int32_t nranks, int32_t nz, int32_t minThick, int32_t* cuts);
} // namespace sf
