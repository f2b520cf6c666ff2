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
} // namespace sf
