// sf_host.cpp -- host-side setup of the B200 SPH solver: parameters, scenes, kernel tables,
// grid dimensions and wall boundary particles.  Pure C++ (no CUDA); compiled without FMA
// contraction because several results (particle counts, table entries) depend on float32
// rounding of separately rounded operations, as in the reference's MSVC build.
//
// Reference interfaces restated here (see SURVEY.md Appendix A/B for the EXE@ addresses):
//   SPHParameters ctor / updateParams()      Source/Controller.cpp:52-64, EXE@0x140011db0, EXE@0x140006ac6
//   SceneManager::setupScene*                 Source/SceneManager.cpp:21-173
//   PrecomputedKernel<Cubic|Spiky>::setRadius EXE@0x14001a4e0, EXE@0x14001a2d0
//   Grid3D::setGrid                           EXE@0x14001ab20
//   SPHSolver::generateBoundaryParticles      EXE@0x140016d80
#include "sf_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>

namespace sf
{
// ------------------------------------------------------------------------------------------------
void params_update(sf_params& p)
{
    const float h = p.kernelRadius;
    const float r = h * 0.25f;
    p.particleRadius  = r;
    p.kernelRadiusSqr = h * h;
    // mass = f(pow(2r, 3) * rho0 * 0.9): the 0.9 makes the rest lattice sit at 0.9 rho0
    const double diameter = static_cast<double>(r) + static_cast<double>(r);
    p.particleMass   = static_cast<float>(std::pow(diameter, 3.0) * static_cast<double>(p.restDensity) * 0.9);
    p.restDensitySqr = p.restDensity * p.restDensity;
}

void params_default(sf_params& p)
{
    std::memset(&p, 0, sizeof(p));
    p.scene           = SF_SCENE_SPHERE_DROP;
    p.numThreads      = 0;
    p.stopTime        = 5.0f;
    p.defaultTimestep = 1.0e-4f;
    for(int d = 0; d < 3; ++d) {
        p.boxMin[d] = -1.0f;
        p.boxMax[d] = 1.0f;
    }
    p.pressureStiffness       = 50000.0f; // DEFAULT_PRESSURE_STIFFNESS (Controller.cpp:136)
    p.viscosity               = 0.05f;    // DEFAULT_VISCOSITY (Controller.cpp:140)
    p.kernelRadius            = 2.0f / 24.0f; // DEFAULT_RESOLUTION 24 (Controller.cpp:128)
    p.bCorrectDensity         = 0;
    p.bUseBoundaryParticles   = 1;
    p.bUseAttractivePressure  = 0;
    p.boundaryRestitution     = 0.1f; // DEFAULT_BOUNDARY_RESTITUTION (Controller.cpp:144)
    p.attractivePressureRatio = 0.1f;
    p.restDensity             = 1000.0f;
    params_update(p);
}

// ------------------------------------------------------------------------------------------------
// Scenes.  Every scene of Source/SceneManager.cpp is a union of lattice blocks with pitch 2r,
// emitted x-outermost / z-innermost; a block is anchored either at its lower corner and grows
// upward, or at its upper corner and grows downward (second block of DoubleDambreak, :165).
namespace
{
struct LatticeBlock {
    float anchor[3];
    int   count[3];
    float dir;     // +1 grow upward from anchor, -1 grow downward
    bool  ballCut; // keep only points with |p| <= 0.5 (SphereDrop, :83)
};

struct Emitter {
    float*   out;
    uint64_t cap;
    uint64_t n = 0;
    void put(float x, float y, float z)
    {
        if(out && n < cap) {
            float* o = out + 3 * n;
            o[0]     = x;
            o[1]     = y;
            o[2]     = z;
        }
        ++n;
    }
};

inline int lattice_count(float lo, float hi, float pitch) { return static_cast<int>((hi - lo) / pitch); }

void emit_block(const LatticeBlock& b, float pitch, Emitter& e)
{
    for(int ix = 0; ix < b.count[0]; ++ix) {
        const float sx = pitch * static_cast<float>(ix);
        const float x  = b.dir > 0 ? b.anchor[0] + sx : b.anchor[0] - sx;
        for(int iy = 0; iy < b.count[1]; ++iy) {
            const float sy = pitch * static_cast<float>(iy);
            const float y  = b.dir > 0 ? b.anchor[1] + sy : b.anchor[1] - sy;
            for(int iz = 0; iz < b.count[2]; ++iz) {
                const float sz = pitch * static_cast<float>(iz);
                const float z  = b.dir > 0 ? b.anchor[2] + sz : b.anchor[2] - sz;
                if(b.ballCut) {
                    // glm::length(ppos - center), center = 0: sqrt((x*x + y*y) + z*z)
                    const float dx = x - 0.0f, dy = y - 0.0f, dz = z - 0.0f;
                    if(std::sqrt((dx * dx + dy * dy) + dz * dz) > 0.5f) continue;
                }
                e.put(x, y, z);
            }
        }
    }
}

LatticeBlock block_from_box(const float lo[3], const float hi[3], float pitch, float dir)
{
    LatticeBlock b{};
    for(int d = 0; d < 3; ++d) {
        b.count[d]  = lattice_count(lo[d], hi[d], pitch);
        b.anchor[d] = dir > 0 ? lo[d] : hi[d];
    }
    b.dir     = dir;
    b.ballCut = false;
    return b;
}
} // namespace

uint64_t scene_generate(const sf_params& p, int scene, float* out_xyz, uint64_t cap)
{
    const float r     = p.particleRadius;
    const float pitch = 2.0f * r;
    Emitter     e{ out_xyz, cap };
    switch(scene) {
        case SF_SCENE_CUBE_DROP: { // :41-64
            const float lo[3] = { -0.5f, -0.5f, -0.5f }, hi[3] = { 0.5f, 0.5f, 0.5f };
            emit_block(block_from_box(lo, hi, pitch, +1.f), pitch, e);
            break;
        }
        case SF_SCENE_SPHERE_DROP: { // :67-91, cubic lattice of int(2*radius/pitch)^3 culled to the ball
            LatticeBlock b{};
            const int    g = static_cast<int>(2.0f * 0.5f / pitch);
            for(int d = 0; d < 3; ++d) {
                b.anchor[d] = 0.0f - 0.5f;
                b.count[d]  = g;
            }
            b.dir     = +1.f;
            b.ballCut = true;
            emit_block(b, pitch, e);
            break;
        }
        case SF_SCENE_DAMBREAK:
        case SF_SCENE_DOUBLE_DAMBREAK: { // :94-119 and :122-173
            const float lo[3] = { -1.0f + r, -1.0f + r, -1.0f + r }, hi[3] = { 0.4f, 0.4f, -0.5f };
            emit_block(block_from_box(lo, hi, pitch, +1.f), pitch, e);
            if(scene == SF_SCENE_DOUBLE_DAMBREAK) {
                const float lo2[3] = { -0.4f + 0.0f, -1.0f + r, 0.5f + 0.0f };
                const float hi2[3] = { 1.0f - r, 0.4f - 0.0f, 1.0f - r };
                emit_block(block_from_box(lo2, hi2, pitch, -1.f), pitch, e);
            }
            break;
        }
        default: break;
    }
    return e.n;
}

// ------------------------------------------------------------------------------------------------
// Kernel tables (A.2).  Lookup in the step: tab[min((uint32)(int64)(sqrtf(d2)*invStep), 10000)],
// truncating index, no interpolation -- so the device must use these very tables.
void build_tables(float h, KernelTables& t)
{
    constexpr float pi_f = 3.14159274f;
    t.cubicW.assign(kTableEntries, 0.f);
    t.spikyGrad.assign(kTableEntries, 0.f);
    t.radius2 = h * h;
    const float step = h / 10000.0f;
    t.invStep        = static_cast<float>(1.0 / static_cast<double>(step));

    // cubic spline: k = 8/(pi h^3)
    const float  h3     = (h * h) * h;
    const float  kCubic = static_cast<float>(8.0 / static_cast<double>(h3 * pi_f));
    // spiky gradient: l = -45/(pi h^6)
    const float  h6pi   = std::pow(h, 6.0f) * pi_f;
    const float  lSpiky = static_cast<float>(-45.0 / static_cast<double>(h6pi));
    const double kd     = static_cast<double>(kCubic);

    for(int i = 0; i < kTableSize; ++i) {
        const float X = static_cast<float>(i) * step;
        // --- cubic W(X)
        const float q = X / h;
        float       w = 0.f;
        if(1.0f >= q) {
            if(0.5f >= q) {
                const float q2 = q * q, q3 = q2 * q;
                w = static_cast<float>(((static_cast<double>(q3) * 6.0 - static_cast<double>(q2) * 6.0) + 1.0) * kd);
            } else {
                w = static_cast<float>((2.0 * std::pow(1.0 - static_cast<double>(q), 3.0)) * kd);
            }
        }
        t.cubicW[i] = w;
        // --- spiky gradW_x(X, 0, 0) / X
        float g = 0.f;
        if(static_cast<double>(X) > 1e-6) {
            const float r2 = X * X + 0.0f;
            float       gx = 0.f;
            if(h * h >= r2) {
                const float rl = std::sqrt(r2);
                gx             = ((((h - rl) * (h - rl)) * lSpiky) * X) * (1.0f / rl);
            }
            g = gx / X;
        }
        t.spikyGrad[i] = g;
    }
    t.cubicW[kTableSize]    = 0.f; // aliases the (unused) cubic gradW[0] = 0
    t.spikyGrad[kTableSize] = 0.f;
    t.spikyGrad[0]          = 0.f;
    t.Wzero                 = t.cubicW[0]; // W[min((uint)(invStep*0), 10000)]
}

// ------------------------------------------------------------------------------------------------
void grid_dims(const sf_params& p, int32_t n[3])
{
    for(int d = 0; d < 3; ++d) n[d] = static_cast<int32_t>(std::ceil((p.boxMax[d] - p.boxMin[d]) / p.kernelRadius));
}

bool cell_coords_checked(const sf_params& p, const int32_t n[3], const float* x, int32_t c[3])
{
    for(int d = 0; d < 3; ++d) {
        if(!std::isfinite(x[d])) return false;
        const float t = (x[d] - p.boxMin[d]) / p.kernelRadius;
        if(!(t >= 0.0f) || !(t < static_cast<float>(n[d]))) return false;
        c[d] = static_cast<int32_t>(t);
        if(c[d] < 0 || c[d] >= n[d]) return false;
    }
    return true;
}

int32_t cell_layer(const sf_params& p, int32_t n, float coord, int axis)
{
    int32_t c = static_cast<int32_t>((coord - p.boxMin[axis]) / p.kernelRadius);
    c         = c < n - 1 ? c : n - 1;
    return c > 0 ? c : 0;
}

// ------------------------------------------------------------------------------------------------
// z-slab planning
void slab_plan(const uint64_t* layerCounts, int32_t nz, int32_t nranks, int32_t minThick, int32_t* cuts)
{
    uint64_t total = 0;
    for(int32_t z = 0; z < nz; ++z) total += layerCounts[z];
    cuts[0]      = 0;
    cuts[nranks] = nz;
    uint64_t acc = 0;
    int32_t  z   = 0;
    for(int32_t r = 1; r < nranks; ++r) {
        const uint64_t want = total * static_cast<uint64_t>(r) / static_cast<uint64_t>(nranks);
        while(z < nz && acc + layerCounts[z] <= want) acc += layerCounts[z++];
        // z = first layer that would overshoot: cut below or above it, whichever is closer to the target
        int32_t cut = z;
        if(z < nz && (want - acc) * 2 > layerCounts[z]) cut = z + 1;
        const int32_t lo = cuts[r - 1] + minThick, hi = nz - (nranks - r) * minThick;
        cut     = cut < lo ? lo : (cut > hi ? hi : cut);
        cuts[r] = cut;
        while(z < cut) acc += layerCounts[z++];
    }
}

void slab_rebalance(const uint32_t* table, int32_t rowWords, int32_t nranks, int32_t nz, int32_t minThick, int32_t* cuts)
{
    (void)nz;
    std::vector<int32_t> old(cuts, cuts + nranks + 1);
    bool lostBottom = false; // did the slab above the previous cut (= the lower slab of this one) just lose its bottom layer?
    for(int32_t b = 1; b < nranks; ++b) {
        const uint32_t* lower = table + static_cast<size_t>(b - 1) * rowWords;
        const uint32_t* upper = table + static_cast<size_t>(b) * rowWords;
        const int64_t   A = lower[2], Bc = upper[2];
        const int64_t   topA = lower[4], botB = upper[3];
        const int32_t   thickA = old[b] - old[b - 1], thickB = old[b + 1] - old[b];
        // moving one layer changes the difference by twice that layer's population: move only past that hysteresis.
        // A slab never loses two layers in one substep: cuts are decided bottom-up, and a slab that has just given
        // its bottom layer to the slab below keeps its top layer for this substep.
        bool upperLosesBottom = false;
        if(A > Bc + 2 * topA && thickA > minThick + 1 && !lostBottom) cuts[b] = old[b] - 1;
        else if(Bc > A + 2 * botB && thickB > minThick + 1) {
            cuts[b]          = old[b] + 1;
            upperLosesBottom = true;
        }
        lostBottom = upperLosesBottom;
    }
}

// ------------------------------------------------------------------------------------------------
// Wall patches (A.4): 6 walls x (nA x nA x nB) jittered points hanging outside the box; the patch
// is expressed in wall-local tangential coordinates and follows the particle (A.6).
// Per wall the binary draws three numbers (EXE@0x14001705a-0x1400172e0) and gives them these roles -- J = tangential
// jitter d*(hi-lo)+lo, D = depth jitter d*(lo-0)+0, ti / tj = lattice coordinate of the outer / middle loop:
//   LX, UX: x = normal + D(d3)   y = J(d2) + ti       z = J(d1) + tj
//   LY, UY: x = J(d3) + ti       y = normal + D(d2)   z = J(d1) + tj
//   LZ, UZ: x = J(d3) + ti       y = J(d2) + tj       z = normal + D(d1)
// with normal = boxMin - depth (lower walls) or depth + boxMax (upper walls).  tests/test_oracle_vs_exe.py checks this
// table against the disassembly.
void generate_boundary(const sf_params& p, uint32_t seed, std::vector<float> walls[6])
{
    std::mt19937 gen(seed);
    // MSVC's std::generate_canonical<float, 24> over mt19937 (EXE@0x1400101a0): one 32-bit draw converted to float
    // (round to nearest) and divided by 2^32f -- so it can return exactly 1.0f, as in the reference binary
    auto U = [&gen]() { return static_cast<float>(static_cast<uint32_t>(gen())) / 4294967296.0f; };

    const float r = p.particleRadius, h = p.kernelRadius;
    const float jitLo = static_cast<float>(static_cast<double>(r) * 0.1);
    const float jitHi = static_cast<float>(static_cast<double>(r) * 0.3);
    const float pitch = r * 1.7f;
    const int   nA    = static_cast<int>(std::ceil(h * 3.0f / pitch)) + 1;
    const int   nB    = static_cast<int>(std::ceil(h / pitch));
    const float base  = r - h;
    for(int w = 0; w < 6; ++w) {
        walls[w].clear();
        walls[w].reserve(static_cast<size_t>(nA) * nA * nB * 3);
    }
    // role of draw 1 / 2 / 3 per axis of the wall normal: 0 = normal (depth jitter), 1 = x, 2 = y, 3 = z tangential
    auto J = [&](float d) { return d * (jitHi - jitLo) + jitLo; };
    auto D = [&](float d) { return d * (jitLo - 0.0f) + 0.0f; };
    for(int i = 0; i < nA; ++i) {
        for(int j = 0; j < nA; ++j) {
            for(int k = 0; k < nB; ++k) {
                const float ti    = base + static_cast<float>(i) * pitch;
                const float tj    = base + static_cast<float>(j) * pitch;
                const float depth = static_cast<float>(k) * pitch + r;
                for(int w = 0; w < 6; ++w) { // LX UX LY UY LZ UZ, three fresh draws each
                    const float d1 = U(), d2 = U(), d3 = U();
                    const int   axis = w / 2;
                    const float nrm  = (w & 1) ? depth + p.boxMax[axis] : p.boxMin[axis] - depth;
                    float       q[3];
                    if(axis == 0) {
                        q[2] = J(d1) + tj;
                        q[1] = J(d2) + ti;
                        q[0] = nrm + D(d3);
                    } else if(axis == 1) {
                        q[2] = J(d1) + tj;
                        q[1] = nrm + D(d2);
                        q[0] = J(d3) + ti;
                    } else {
                        q[2] = nrm + D(d1);
                        q[1] = J(d2) + tj;
                        q[0] = J(d3) + ti;
                    }
                    walls[w].insert(walls[w].end(), q, q + 3);
                }
            }
        }
    }
}
// ------------------------------------------------------------------------------------------------
// Candidate masks of the wall particles (density pass).  The reference tests every particle of a wall list against
// every near-wall fluid particle (243 per wall with the generated walls, of which about 9 are within h).  A fluid
// particle within h of wall w meets the list through its SHIFTED position (A.6): the two tangential coordinates
// wrapped into [0, h), the normal coordinate within h of the wall plane -- a cube of edge h.  That cube is cut into
// kWallSub^3 sub-cells, and for each sub-cell the wall particles that can be within h of SOME point of it are marked
// in a bit mask (bit b of word b / 32 = wall particle b): the density pass walks the set bits of its sub-cell's mask
// in ascending order -- the list order -- and applies the exact test to each, so sums, list entries and their order
// are those of the full loop.  Works for any wall list (sf_set_boundary_particles), not only the generated ones.
// Conservative by construction: the sub-cell boxes overlap by 1 % of h (the device derives the sub-cell from its
// own fp32 shifted position; its rounding is eight orders of magnitude below that), the distance bound carries a
// relative 1e-4, and positions beyond the wall plane (possible for an upload in the last, partly outside, cell
// layer) take entry kWallSubCells: every particle of the list.
int wall_subcell(const sf_params& p, int wall, const float x[3])
{
    const int   A = wall / 2, t1 = A == 0 ? 1 : 0, t2 = A == 2 ? 1 : 2;
    const float h = p.kernelRadius, inv = wall_sub_inv(p);
    const float s1 = x[t1] - h * std::floor(x[t1] / h), s2 = x[t2] - h * std::floor(x[t2] / h); // wall_shift
    const float dn = (wall & 1) ? p.boxMax[A] - x[A] : x[A] - p.boxMin[A];
    if(dn < 0.0f) return kWallSubCells;
    auto idx = [inv](float v) {
        const int i = static_cast<int>(std::floor(v * inv));
        return i < 0 ? 0 : (i > kWallSub - 1 ? kWallSub - 1 : i);
    };
    return (idx(dn) * kWallSub + idx(s2)) * kWallSub + idx(s1);
}

float wall_sub_inv(const sf_params& p) { return static_cast<float>(kWallSub) / p.kernelRadius; }

void wall_candidate_masks(const sf_params& p, int wall, const float* xyz, uint32_t n, uint32_t words, uint32_t* masks)
{
    const int    A = wall / 2, t1 = A == 0 ? 1 : 0, t2 = A == 2 ? 1 : 2;
    const double h = p.kernelRadius, inv = wall_sub_inv(p), m = 0.01 * h, reach = h * (1.0 + 1e-4);
    const double wallPlane = (wall & 1) ? p.boxMax[A] : p.boxMin[A], sgn = (wall & 1) ? -1.0 : 1.0;
    std::fill(masks, masks + static_cast<size_t>(kWallSubCells + 1) * words, 0u);
    for(uint32_t b = 0; b < n; ++b) masks[static_cast<size_t>(kWallSubCells) * words + b / 32u] |= 1u << (b % 32u);
    auto gap = [](double q, double lo, double hi) { return q < lo ? lo - q : (q > hi ? q - hi : 0.0); };
    for(int in = 0; in < kWallSub; ++in)
        for(int i2 = 0; i2 < kWallSub; ++i2)
            for(int i1 = 0; i1 < kWallSub; ++i1) {
                uint32_t* mk = masks + static_cast<size_t>((in * kWallSub + i2) * kWallSub + i1) * words;
                // normal coordinate: distance dn from the wall plane into the box, absolute position = plane + sgn * dn
                const double nLoD = in / inv - m, nHiD = (in + 1) / inv + m;
                const double nLo = sgn > 0 ? wallPlane + nLoD : wallPlane - nHiD, nHi = sgn > 0 ? wallPlane + nHiD : wallPlane - nLoD;
                for(uint32_t b = 0; b < n; ++b) {
                    const double g1 = gap(xyz[3 * b + t1], i1 / inv - m, (i1 + 1) / inv + m);
                    const double g2 = gap(xyz[3 * b + t2], i2 / inv - m, (i2 + 1) / inv + m);
                    const double gn = gap(xyz[3 * b + A], nLo, nHi);
                    if(g1 * g1 + g2 * g2 + gn * gn <= reach * reach) mk[b / 32u] |= 1u << (b % 32u);
                }
            }
}
} // namespace sf
