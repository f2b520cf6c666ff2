// sf_kernels.cuh -- hand-written sm_100a kernels of the SPH substep (included by sf_api.cu).
//
// Numerics contract: this translation unit is compiled with -fmad=false,
// default -prec-div=true -prec-sqrt=true -ftz=false, so every * + - / sqrt is a separately and
// correctly rounded IEEE-754 fp32 operation, like the reference's SSE scalar code.  All pair loops
// accumulate in the reference's traversal order (cells z -> y -> x, ascending particle id inside a
// cell, then the wall lists X, Y, Z), one scalar accumulator per thread, which makes every field
// bit-identical to a faithful CPU evaluation -- not merely within tolerance.
#pragma once
#include <cfloat>
#include "sf_device.cuh"

namespace sf
{
constexpr int kTab = 10000;

// ------------------------------------------------------------------------------------------------
// small helpers
__device__ __forceinline__ float dist2(float dx, float dy, float dz) { return (dx * dx + dy * dy) + dz * dz; }

// A.2: tab[min((uint32)(int64)(sqrtf(d2) * invStep), 10000)]
__device__ __forceinline__ uint32_t table_index(float d2, float invStep)
{
    const float    t = sqrtf(d2) * invStep;
    const uint32_t i = static_cast<uint32_t>(__float2ll_rz(t));
    return i > static_cast<uint32_t>(kTab) ? static_cast<uint32_t>(kTab) : i;
}

// A.11 Pr(rho)
__device__ __forceinline__ float pressure_of(const DevParams& P, float rho)
{
    const float x = rho / P.rho0;
    float       t = x * x;
    t             = t * x;
    t             = t * t;
    t             = t * x;
    const float pp = static_cast<float>(static_cast<double>(t) - 1.0);
    if(P.attractive) return fmaxf(pp, pp * P.attractRatio);
    return static_cast<float>(fmax(static_cast<double>(pp), 0.0));
}

__device__ __forceinline__ float comp(const float4& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

// A.6 wall patch: shifted position x~ = x - h*floorf(x/h) on the tangential axes, x - h*0 on axis A
template<int A>
__device__ __forceinline__ float3 wall_shift(const DevParams& P, const float4& x)
{
    float3 s;
    s.x = x.x - P.h * (A == 0 ? 0.0f : floorf(x.x / P.h));
    s.y = x.y - P.h * (A == 1 ? 0.0f : floorf(x.y / P.h));
    s.z = x.z - P.h * (A == 2 ? 0.0f : floorf(x.z / P.h));
    return s;
}

// returns -1 when the particle is not within h of a wall of axis A, else the wall id 2A / 2A+1
template<int A>
__device__ __forceinline__ int wall_of(const DevParams& P, const float4& x)
{
    const float xa = comp(x, A);
    const float lo = P.h + P.bmin[A], hi = P.bmax[A] - P.h;
    if(lo > xa) return 2 * A;
    if(xa > hi) return 2 * A + 1;
    return -1;
}

// ------------------------------------------------------------------------------------------------
// Substep prologue: computeTimeStep (A.5) from the max |v|^2 the previous substep's integrate
// kernel reduced, and the frame-time bookkeeping of Simulator::doSimulation (Simulator.cpp:46-51).
__global__ void k_begin_step(DevState* st, DevParams P)
{
    if(st->frameTarget > 0.0 && !(static_cast<double>(st->frameTime) < st->frameTarget)) {
        st->skip = 1;
        return;
    }
    st->skip          = 0;
    const unsigned pr = st->step & 1u;
    const float    M  = __uint_as_float(st->maxv2Bits[pr]);
    const float    maxv = sqrtf(M);
    float          dt   = (static_cast<double>(maxv) > 1e-8) ? ((P.r + P.r) / maxv) * 0.2f : 1e10f;
    dt                  = fmaxf(dt, P.dtMin);
    dt                  = fminf(dt, P.dtMax);
    st->dt              = dt;
    st->maxv2Bits[pr ^ 1u] = __float_as_uint(FLT_MIN);
    st->step += 1u;
    st->frameTime = st->frameTime + dt;
    st->stepsDone += 1ull;
}

// max |v|^2 of the uploaded velocities (computeMaxVel A.5): m = (vy*vy + vx*vx) + vz*vz
__global__ void k_init_maxvel(const float4* __restrict__ vel, uint32_t n, DevState* st)
{
    float m = FLT_MIN;
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 v = vel[i];
        m              = fmaxf(m, (v.y * v.y + v.x * v.x) + v.z * v.z);
    }
    for(int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if((threadIdx.x & 31) == 0) atomicMax(&st->maxv2Bits[st->step & 1u], __float_as_uint(m));
}

// ------------------------------------------------------------------------------------------------
// (1) cell hashing: collectParticlesToCells' index expression (A.7), bit-exact
__device__ __forceinline__ uint32_t cell_key(const DevParams& P, const float4& x)
{
    int cx = static_cast<int>((x.x - P.bmin[0]) / P.h);
    int cy = static_cast<int>((x.y - P.bmin[1]) / P.h);
    int cz = static_cast<int>((x.z - P.bmin[2]) / P.h);
    cx     = max(min(cx, P.nx - 1), 0);
    cy     = max(min(cy, P.ny - 1), 0);
    cz     = max(min(cz, P.nz - 1), 0);
    return static_cast<uint32_t>((cz * P.ny + cy) * P.nx + cx);
}

__global__ void k_hash(const float4* __restrict__ pos, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                       DevParams P, const DevState* st)
{
    if(st->skip) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= P.n) return;
    keys[i] = cell_key(P, pos[i]);
    vals[i] = i;
}

// ------------------------------------------------------------------------------------------------
// (1) stable LSD radix sort of (cell key, slot), one digit per pass:
//     k_radix_hist -> k_radix_scan -> k_radix_scatter.   Tile = RS_THREADS * RS_ITEMS keys per CTA;
//     each warp owns 32*RS_ITEMS consecutive keys and ranks them 32 at a time with match.any, so
//     equal digits keep their input order (stability) without any shared-memory sort.
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS   = 8;
constexpr int RS_TILE    = RS_THREADS * RS_ITEMS;
constexpr int RS_MAXRADIX = 256;

__global__ void __launch_bounds__(RS_THREADS)
k_radix_hist(const uint32_t* __restrict__ keys, uint32_t n, int shift, int radix, uint32_t* __restrict__ counts,
             uint32_t numBlocks, const DevState* st)
{
    if(st->skip) return;
    __shared__ uint32_t hist[RS_MAXRADIX];
    for(int d = threadIdx.x; d < radix; d += RS_THREADS) hist[d] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    const uint32_t mask = static_cast<uint32_t>(radix - 1);
#pragma unroll
    for(int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = base + r * RS_THREADS + threadIdx.x;
        if(i < n) atomicAdd(&hist[(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    for(int d = threadIdx.x; d < radix; d += RS_THREADS) counts[static_cast<size_t>(d) * numBlocks + blockIdx.x] = hist[d];
}

// one CTA per digit: exclusive scan of that digit's per-tile counts, total to totals[digit]
__global__ void __launch_bounds__(1024)
k_radix_scan(uint32_t* __restrict__ counts, uint32_t numBlocks, uint32_t* __restrict__ totals, const DevState* st)
{
    if(st->skip) return;
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t carry;
    uint32_t*           row = counts + static_cast<size_t>(blockIdx.x) * numBlocks;
    if(threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for(uint32_t base = 0; base < numBlocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < numBlocks ? row[i] : 0u;
        uint32_t       x = v;
        for(int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if(lane >= o) x += y;
        }
        if(lane == 31) warpSums[wid] = x;
        __syncthreads();
        if(wid == 0) {
            uint32_t s = warpSums[lane];
            for(int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
                if(lane >= o) s += y;
            }
            warpSums[lane] = s; // inclusive over warps
        }
        __syncthreads();
        const uint32_t warpOff = wid ? warpSums[wid - 1] : 0u;
        const uint32_t c       = carry;
        if(i < numBlocks) row[i] = c + warpOff + x - v;
        __syncthreads();
        if(threadIdx.x == 1023) carry = c + warpOff + x;
        __syncthreads();
    }
    if(threadIdx.x == 0) totals[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(RS_THREADS)
k_radix_scatter(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn, uint32_t* __restrict__ keysOut,
                uint32_t* __restrict__ valsOut, uint32_t n, int shift, int radix, const uint32_t* __restrict__ counts,
                uint32_t numBlocks, const uint32_t* __restrict__ totals, const DevState* st)
{
    if(st->skip) return;
    constexpr int NW = RS_THREADS / 32;
    __shared__ uint32_t warpCnt[NW][RS_MAXRADIX];
    __shared__ uint32_t digitBase[RS_MAXRADIX];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for(int i = threadIdx.x; i < NW * RS_MAXRADIX; i += RS_THREADS) (&warpCnt[0][0])[i] = 0;
    // exclusive scan of the digit totals (radix <= 256 = RS_THREADS): Hillis-Steele in smem
    {
        uint32_t v = threadIdx.x < radix ? totals[threadIdx.x] : 0u;
        digitBase[threadIdx.x] = v;
        __syncthreads();
        for(int o = 1; o < RS_MAXRADIX; o <<= 1) {
            const uint32_t y = threadIdx.x >= o ? digitBase[threadIdx.x - o] : 0u;
            __syncthreads();
            digitBase[threadIdx.x] += y;
            __syncthreads();
        }
        const uint32_t incl = digitBase[threadIdx.x];
        __syncthreads();
        digitBase[threadIdx.x] = incl - v;
        __syncthreads();
    }
    const uint32_t mask     = static_cast<uint32_t>(radix - 1);
    const uint32_t warpBase = blockIdx.x * RS_TILE + wid * (32 * RS_ITEMS);
    uint32_t       key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for(int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = warpBase + r * 32 + lane;
        key[r]           = i < n ? keysIn[i] : 0xffffffffu;
        val[r]           = i < n ? valsIn[i] : 0u;
    }
#pragma unroll
    for(int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t d     = (key[r] >> shift) & mask;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t below = __popc(peers & ((1u << lane) - 1u));
        const uint32_t prev  = warpCnt[wid][d];
        rank[r]              = prev + below;
        __syncwarp();
        if(below == 0) warpCnt[wid][d] = prev + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    if(threadIdx.x < radix) {
        uint32_t run = digitBase[threadIdx.x] + counts[static_cast<size_t>(threadIdx.x) * numBlocks + blockIdx.x];
#pragma unroll
        for(int w = 0; w < NW; ++w) {
            const uint32_t c       = warpCnt[w][threadIdx.x];
            warpCnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for(int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = warpBase + r * 32 + lane;
        if(i < n) {
            const uint32_t d   = (key[r] >> shift) & mask;
            const uint32_t dst = warpCnt[wid][d] + rank[r];
            keysOut[dst]       = key[r];
            valsOut[dst]       = val[r];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// (1) cell start/end tables + particle reorder
__global__ void k_clear_cells(uint4* __restrict__ tab, size_t nvec, const DevState* st)
{
    if(st->skip) return;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for(size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<size_t>(gridDim.x) * blockDim.x) tab[i] = z;
}

__global__ void k_cell_bounds(const uint32_t* __restrict__ keys, uint32_t n, uint2* __restrict__ cellTab, const DevState* st)
{
    if(st->skip) return;
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= n) return;
    const uint32_t k = keys[p];
    if(p == 0 || keys[p - 1] != k) cellTab[k].x = p;
    if(p == n - 1 || keys[p + 1] != k) cellTab[k].y = p + 1;
}

// Gather A -> B in key order.  The radix sort is stable with respect to the previous substep's
// order, not the original ids, so inside each cell the slot is re-ranked by original id: that
// reproduces the ascending-id cell lists of the reference's serial push_back (A.7) and with it the
// reference's summation order.
__global__ void k_reorder(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                          const uint2* __restrict__ cellTab, const float4* __restrict__ posA,
                          const float4* __restrict__ velA, const uint32_t* __restrict__ idA, float4* __restrict__ posB,
                          float4* __restrict__ velB, uint32_t* __restrict__ idB, uint32_t n, const DevState* st)
{
    if(st->skip) return;
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= n) return;
    const uint32_t src  = vals[p];
    const uint32_t myId = idA[src];
    const uint2    ce   = cellTab[keys[p]];
    uint32_t       rank = 0;
    for(uint32_t q = ce.x; q < ce.y; ++q) rank += (idA[vals[q]] < myId) ? 1u : 0u;
    const uint32_t dst = ce.x + rank;
    float4         x   = posA[src];
    float4         v   = velA[src];
    x.w                = 0.f;
    v.w                = 0.f;
    posB[dst]          = x;
    velB[dst]          = v;
    idB[dst]           = myId;
}

// ------------------------------------------------------------------------------------------------
// neighbour traversal shared by the density pass: the 27 cells in reference order collapse to 9
// contiguous slot runs because the 3 x-adjacent cells of a row are adjacent in key order.
struct Run {
    uint32_t b, e;
};
__device__ __forceinline__ Run row_run(const uint2* __restrict__ cellTab, int rowBase, int x0, int x1)
{
    Run r{ 0xffffffffu, 0u };
    for(int x = x0; x <= x1; ++x) {
        const uint2 ce = __ldg(&cellTab[rowBase + x]);
        if(ce.y > ce.x) {
            r.b = min(r.b, ce.x);
            r.e = max(r.e, ce.y);
        }
    }
    if(r.e == 0u) r.b = 0u;
    return r;
}

__device__ __forceinline__ void nbr_append(const DevBuffers& B, const DevParams& P, uint32_t p, uint32_t& k, uint32_t j, uint32_t idx)
{
    if(k < static_cast<uint32_t>(P.kmax)) {
        B.nbrJ[static_cast<size_t>(k) * P.npad + p]   = j;
        B.nbrIdx[static_cast<size_t>(k) * P.npad + p] = static_cast<uint16_t>(idx);
    }
    ++k;
}

// (2) density (A.8) + equation of state, and the neighbour list for the two later passes.
template<int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_density(DevBuffers B, DevParams P)
{
    if(B.state->skip) return;
    extern __shared__ float s_tab[]; // cubic W table
    for(int i = threadIdx.x; i <= kTab; i += BLOCK) s_tab[i] = B.tabW[i];
    __syncthreads();

    for(uint32_t base = blockIdx.x * BLOCK; base < P.n; base += gridDim.x * BLOCK) {
        const uint32_t p = base + threadIdx.x;
        if(p >= P.n) continue;
        const float4   xp  = B.posB[p];
        const uint32_t key = B.keyB[p];
        const int      cx  = static_cast<int>(key % static_cast<uint32_t>(P.nx));
        const int      t   = static_cast<int>(key / static_cast<uint32_t>(P.nx));
        const int      cy = t % P.ny, cz = t / P.ny;
        const int      x0 = max(cx - 1, 0), x1 = min(cx + 1, P.nx - 1);

        float    S = P.Wzero;
        uint32_t k = 0;
        for(int dz = -1; dz <= 1; ++dz) {
            const int z = cz + dz;
            if(z < 0 || z >= P.nz) continue;
            for(int dy = -1; dy <= 1; ++dy) {
                const int y = cy + dy;
                if(y < 0 || y >= P.ny) continue;
                const Run run = row_run(B.cellTab, (z * P.ny + y) * P.nx, x0, x1);
                for(uint32_t j = run.b; j < run.e; ++j) {
                    if(j == p) continue;
                    const float4 xq = B.posB[j];
                    const float  d2 = dist2(xq.x - xp.x, xq.y - xp.y, xq.z - xp.z);
                    if(P.radius2 >= d2) {
                        const uint32_t idx = table_index(d2, P.invStep);
                        S += s_tab[idx];
                        nbr_append(B, P, p, k, j, idx);
                    }
                }
            }
        }
        const uint32_t nFluid = k;
        uint32_t       nWall[3] = { 0, 0, 0 };
        if(P.useBoundary) {
#define SF_WALL_DENSITY(A)                                                                               \
    {                                                                                                    \
        const int w = wall_of<A>(P, xp);                                                                 \
        if(w >= 0) {                                                                                     \
            const float3   xs = wall_shift<A>(P, xp);                                                    \
            const float4*  bw = B.bnd + static_cast<size_t>(w) * P.bndStride;                            \
            const uint32_t nb = P.nbnd[w];                                                               \
            const uint32_t k0 = k;                                                                       \
            for(uint32_t b = 0; b < nb; ++b) {                                                           \
                const float4 xb = __ldg(&bw[b]);                                                         \
                const float  d2 = dist2(xb.x - xs.x, xb.y - xs.y, xb.z - xs.z);                          \
                if(P.radius2 >= d2) {                                                                    \
                    const uint32_t idx = table_index(d2, P.invStep);                                     \
                    S += s_tab[idx];                                                                     \
                    nbr_append(B, P, p, k, b, idx);                                                      \
                }                                                                                        \
            }                                                                                            \
            nWall[A] = k - k0;                                                                           \
        }                                                                                                \
    }
            SF_WALL_DENSITY(0)
            SF_WALL_DENSITY(1)
            SF_WALL_DENSITY(2)
#undef SF_WALL_DENSITY
        }
        unsigned err = 0;
        if(k > static_cast<uint32_t>(P.kmax) || nFluid > 16383u) err |= SF_DEVERR_NBR_OVERFLOW;
        if(nWall[0] > 63u || nWall[1] > 63u || nWall[2] > 63u) err |= SF_DEVERR_WALL_OVERFLOW;
        if(err) atomicOr(&B.state->errFlags, err);
        B.nbrCnt[p] = nFluid | (nWall[0] << 14) | (nWall[1] << 20) | (nWall[2] << 26);

        const float rho = (1.0f > S) ? 0.0f : fminf(fmaxf(S * P.mass, P.rhoMin), P.rhoMax);
        B.rho[p]        = rho;
        if(!P.correctDensity) {
            // pair-loop terms of A.11 / A.13 hoisted per particle: identical values, computed once
            float pterm, inv;
            if(1e-8 > static_cast<double>(rho)) {
                pterm = __int_as_float(0x7fc00000); // NaN marks "rho < 1e-8: skipped as a neighbour"
                inv   = 1.0f / rho;
            } else {
                pterm = pressure_of(P, rho) / (rho * rho);
                inv   = 1.0f / rho;
            }
            B.posB[p].w = pterm;
            B.velB[p].w = inv;
        }
    }
}

// correctDensity (A.9, default off): Shepard normalisation over the neighbour list, then the
// per-particle terms from the corrected density.
__global__ void k_correct_density(DevBuffers B, DevParams P)
{
    if(B.state->skip) return;
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= P.n) return;
    const float    rp  = B.rho[p];
    const uint32_t cnt = B.nbrCnt[p];
    const uint32_t nF = cnt & 16383u, nW = ((cnt >> 14) & 63u) + ((cnt >> 20) & 63u) + ((cnt >> 26) & 63u);
    float          T = P.Wzero / rp;
    for(uint32_t k = 0; k < nF; ++k) {
        const uint32_t j  = B.nbrJ[static_cast<size_t>(k) * P.npad + p];
        const float    rq = B.rho[j];
        if(!(static_cast<double>(rq) >= 1e-8)) continue;
        T += __ldg(&B.tabW[B.nbrIdx[static_cast<size_t>(k) * P.npad + p]]) / rq;
    }
    for(uint32_t k = nF; k < nF + nW; ++k) T += __ldg(&B.tabW[B.nbrIdx[static_cast<size_t>(k) * P.npad + p]]) / P.rho0;
    B.rho2[p] = (static_cast<double>(T) > 1e-8) ? rp / fminf(T * P.mass, P.rhoMax) : 0.0f;
}

__global__ void k_density_terms(DevBuffers B, DevParams P)
{
    if(B.state->skip) return;
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= P.n) return;
    const float rho = B.rho2[p];
    B.rho[p]        = rho;
    B.posB[p].w     = (1e-8 > static_cast<double>(rho)) ? __int_as_float(0x7fc00000) : pressure_of(P, rho) / (rho * rho);
    B.velB[p].w     = 1.0f / rho;
}

// ------------------------------------------------------------------------------------------------
// (3a) pressure acceleration (A.11) + gravity (A.10) + velocity update (A.12), over the list
template<int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_force(DevBuffers B, DevParams P)
{
    if(B.state->skip) return;
    extern __shared__ float s_tab[]; // spiky gradW/r table
    for(int i = threadIdx.x; i <= kTab; i += BLOCK) s_tab[i] = B.tabG[i];
    __syncthreads();
    const float dt = B.state->dt;

    for(uint32_t base = blockIdx.x * BLOCK; base < P.n; base += gridDim.x * BLOCK) {
        const uint32_t p = base + threadIdx.x;
        if(p >= P.n) continue;
        const float4 xp = B.posB[p]; // w = P_p / rho_p^2
        float4       vp = B.velB[p]; // w = 1 / rho_p
        const float  rp = B.rho[p];
        float        ax = 0.f, ay = 0.f, az = 0.f;
        if(!(1e-8 > static_cast<double>(rp))) {
            const uint32_t cnt = B.nbrCnt[p];
            const uint32_t nF  = min(cnt & 16383u, static_cast<uint32_t>(P.kmax));
            for(uint32_t k = 0; k < nF; ++k) {
                const uint32_t j   = B.nbrJ[static_cast<size_t>(k) * P.npad + p];
                const uint32_t idx = B.nbrIdx[static_cast<size_t>(k) * P.npad + p];
                const float4   xq  = B.posB[j];
                if(xq.w != xq.w) continue; // rho_q < 1e-8
                const float dx = xq.x - xp.x, dy = xq.y - xp.y, dz = xq.z - xp.z;
                const float g  = s_tab[idx];
                const float fp = xq.w + xp.w;
                ax += fp * (g * dx);
                ay += fp * (dy * g);
                az += fp * (g * dz);
            }
            if(P.useBoundary) {
                uint32_t k = nF;
#define SF_WALL_FORCE(A, SH)                                                                              \
    {                                                                                                    \
        const uint32_t nw = (cnt >> SH) & 63u;                                                           \
        if(nw) {                                                                                         \
            const int     w  = wall_of<A>(P, xp);                                                        \
            const float3  xs = wall_shift<A>(P, xp);                                                     \
            const float4* bw = B.bnd + static_cast<size_t>(w) * P.bndStride;                             \
            for(uint32_t e = 0; e < nw && k < static_cast<uint32_t>(P.kmax); ++e, ++k) {                 \
                const uint32_t b   = B.nbrJ[static_cast<size_t>(k) * P.npad + p];                        \
                const uint32_t idx = B.nbrIdx[static_cast<size_t>(k) * P.npad + p];                      \
                const float4   xb  = __ldg(&bw[b]);                                                      \
                const float    dx = xb.x - xs.x, dy = xb.y - xs.y, dz = xb.z - xs.z;                     \
                const float    g = s_tab[idx];                                                           \
                ax += xp.w * (g * dx);                                                                   \
                ay += xp.w * (dy * g);                                                                   \
                az += xp.w * (g * dz);                                                                   \
            }                                                                                            \
        }                                                                                                \
    }
                SF_WALL_FORCE(0, 14)
                SF_WALL_FORCE(1, 20)
                SF_WALL_FORCE(2, 26)
#undef SF_WALL_FORCE
            }
            ax = (ax * P.mass) * P.stiffness;
            ay = (ay * P.mass) * P.stiffness;
            az = (az * P.mass) * P.stiffness;
        }
        if(P.capture) B.accel[p] = make_float4(ax, ay, az, 0.f);
        // addGravity (A.10) then updateVelocity (A.12)
        vp.y = static_cast<float>(static_cast<double>(vp.y) - static_cast<double>(dt) * 9.8);
        vp.x = dt * ax + vp.x;
        vp.y = dt * ay + vp.y;
        vp.z = dt * az + vp.z;
        B.velB[p] = vp; // w still 1/rho_p: the viscosity pass reads {v*, 1/rho} of its neighbours in one load
    }
}

// (3b) XSPH viscosity (A.13) + updatePosition with wall clamp/restitution (A.14) + max |v|^2 for
// the next substep's computeTimeStep (A.5).  Writes the new state into A (sorted order of B).
template<int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_visc_integrate(DevBuffers B, DevParams P)
{
    if(B.state->skip) return;
    extern __shared__ float s_tab[]; // cubic W table
    __shared__ float        s_max[BLOCK / 32];
    for(int i = threadIdx.x; i <= kTab; i += BLOCK) s_tab[i] = B.tabW[i];
    __syncthreads();
    const float dt   = B.state->dt;
    float       vmax = FLT_MIN;

    for(uint32_t base = blockIdx.x * BLOCK; base < P.n; base += gridDim.x * BLOCK) {
        const uint32_t p = base + threadIdx.x;
        if(p >= P.n) continue;
        const float4   xp  = B.posB[p];
        const float4   vp  = B.velB[p];
        const uint32_t cnt = B.nbrCnt[p];
        const uint32_t nF  = min(cnt & 16383u, static_cast<uint32_t>(P.kmax));
        float          sx = 0.f, sy = 0.f, sz = 0.f;
        for(uint32_t k = 0; k < nF; ++k) {
            const uint32_t j   = B.nbrJ[static_cast<size_t>(k) * P.npad + p];
            const uint32_t idx = B.nbrIdx[static_cast<size_t>(k) * P.npad + p];
            const float4   vq  = B.velB[j]; // {v*, 1/rho_q}
            const float    w   = s_tab[idx];
            const float    dvx = vq.x - vp.x, dvy = vq.y - vp.y, dvz = vq.z - vp.z;
            sx += (vq.w * dvx) * w;
            sy += (dvy * vq.w) * w;
            sz += (dvz * vq.w) * w;
        }
        float v[3] = { P.viscosity * (sx * P.mass) + vp.x, P.viscosity * (sy * P.mass) + vp.y, P.viscosity * (sz * P.mass) + vp.z };
        float x[3] = { xp.x, xp.y, xp.z };
#pragma unroll
        for(int d = 0; d < 3; ++d) {
            const float lo = P.bmin[d] + P.r, hi = P.bmax[d] - P.r;
            float       xn = v[d] * dt + x[d];
            if(lo > xn) {
                xn   = lo;
                v[d] = -(v[d] * P.restitution);
            } else if(xn > hi) {
                xn   = hi;
                v[d] = -(v[d] * P.restitution);
            }
            x[d] = xn;
        }
        B.posA[p] = make_float4(x[0], x[1], x[2], 0.f);
        B.velA[p] = make_float4(v[0], v[1], v[2], 0.f);
        B.idA[p]  = B.idB[p]; // A now holds this substep's sorted order
        vmax      = fmaxf(vmax, (v[1] * v[1] + v[0] * v[0]) + v[2] * v[2]);
    }
    for(int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = vmax;
    __syncthreads();
    if(threadIdx.x < 32) {
        float m = threadIdx.x < BLOCK / 32 ? s_max[threadIdx.x] : FLT_MIN;
        for(int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if(threadIdx.x == 0) atomicMax(&B.state->maxv2Bits[B.state->step & 1u], __float_as_uint(m));
    }
}

// ------------------------------------------------------------------------------------------------
// host <-> device marshalling (original particle order on the host side)
__global__ void k_pack_upload(const float* __restrict__ posXYZ, const float* __restrict__ velXYZ, float4* __restrict__ pos,
                              float4* __restrict__ vel, uint32_t* __restrict__ id, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    pos[i] = make_float4(posXYZ[3 * i], posXYZ[3 * i + 1], posXYZ[3 * i + 2], 0.f);
    vel[i] = velXYZ ? make_float4(velXYZ[3 * i], velXYZ[3 * i + 1], velXYZ[3 * i + 2], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    id[i]  = i;
}

__global__ void k_unpack_xyz(const float4* __restrict__ src, const uint32_t* __restrict__ id, float* __restrict__ outXYZ, uint32_t n)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= n) return;
    const float4 v = src[p];
    const size_t o = 3 * static_cast<size_t>(id[p]);
    outXYZ[o]      = v.x;
    outXYZ[o + 1]  = v.y;
    outXYZ[o + 2]  = v.z;
}

__global__ void k_unpack_scalar(const float* __restrict__ src, const uint32_t* __restrict__ id, float* __restrict__ out, uint32_t n)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p < n) out[id[p]] = src[p];
}

__global__ void k_unpack_u32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ id, uint32_t* __restrict__ out, uint32_t n)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p < n) out[id[p]] = src[p];
}

__global__ void k_unpack_pressure(const float* __restrict__ rho, const uint32_t* __restrict__ id, float* __restrict__ out, DevParams P)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p < P.n) out[id[p]] = pressure_of(P, rho[p]);
}
} // namespace sf
