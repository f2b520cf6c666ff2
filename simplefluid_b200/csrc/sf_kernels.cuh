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
#include "sf_internal.h"

namespace sf
{
constexpr int kTab = 10000;
// work unit of the pair kernels (sf_pairs.cuh): a brick of BX x BY x BZ grid cells
#ifndef SF_BZ
#define SF_BZ 4
#endif
constexpr int BX = 8, BY = 4, BZ = SF_BZ;

// ------------------------------------------------------------------------------------------------
// small helpers
__device__ __forceinline__ float dist2(float dx, float dy, float dz) { return (dx * dx + dy * dy) + dz * dz; }

// A.2: tab[min((uint32)(int64)(sqrtf(d2) * invStep), 10000)]
__device__ __forceinline__ uint32_t table_index(float d2, float invStep)
{
    // t >= 0 and far below 2^32, so the reference's (uint32)(int64) truncation equals a direct u32 truncation
    const float t = sqrtf(d2) * invStep;
    return min(__float2uint_rz(t), static_cast<uint32_t>(kTab));
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

// Sub-cell of the h-cube a shifted near-wall position lies in -> entry of the wall's candidate masks (sf_host.cpp:
// wall_subcell is the host mirror, wall_candidate_masks builds the masks).  xs = wall_shift<A>(P, x), w = wall_of<A>.
template<int A>
__device__ __forceinline__ uint32_t wall_subcell(const DevParams& P, const float4& x, const float3& xs, int w)
{
    const float dn = (w & 1) ? P.bmax[A] - comp(x, A) : comp(x, A) - P.bmin[A];
    const float s1 = A == 0 ? xs.y : xs.x, s2 = A == 2 ? xs.y : xs.z;
    const int   i1 = min(max(__float2int_rd(s1 * P.wallSubInv), 0), kWallSub - 1);
    const int   i2 = min(max(__float2int_rd(s2 * P.wallSubInv), 0), kWallSub - 1);
    const int   in = min(max(__float2int_rd(dn * P.wallSubInv), 0), kWallSub - 1);
    return dn < 0.0f ? static_cast<uint32_t>(kWallSubCells) : static_cast<uint32_t>((in * kWallSub + i2) * kWallSub + i1);
}

// ------------------------------------------------------------------------------------------------
// Substep prologue: computeTimeStep (A.5) from the max |v|^2 the previous substep's integrate
// kernel reduced, and the frame-time bookkeeping of Simulator::doSimulation (Simulator.cpp:46-51).
// phases: bit 0 = work-queue resets (needed before the cell tables are built), bit 1 = dt and the frame clock
// (needed before the force pass).  The resident path launches both at once; sf_step_host launches them apart so
// that the velocity upload (which max |v|^2 and hence dt depend on) overlaps the sort and the density pass.
constexpr int kBeginResets = 1, kBeginClock = 2, kBeginAll = 3;
__global__ void k_begin_step(DevState* st, DevParams P, int phases)
{
    if(phases & kBeginResets) {
        if(st->frameTarget > 0.0 && !(static_cast<double>(st->frameTime) < st->frameTarget)) {
            st->skip = 1;
            return;
        }
        st->skip          = 0;
        st->brickCount    = 0u;
        st->cursor[0] = st->cursor[1] = st->cursor[2] = st->cursor[3] = st->cursor[4] = st->cursor[5] = 0u;
    }
    if(!(phases & kBeginClock) || st->skip) return;
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
    const int axisM = 3 - P.axisS;
    int cx = static_cast<int>((x.x - P.bmin[0]) / P.h);
    int cm = static_cast<int>((comp(x, axisM) - P.bmin[axisM]) / P.h);
    int cs = static_cast<int>((comp(x, P.axisS) - P.bmin[P.axisS]) / P.h);
    cx     = max(min(cx, P.nx - 1), 0);
    cm     = max(min(cm, P.ny - 1), 0);
    cs     = max(min(cs, P.nzGlobal - 1), 0);
    cs     = max(min(cs - P.z0, P.nz - 1), 0); // local layer of this rank's window (z0 = 0 on a single GPU)
    return static_cast<uint32_t>((cs * P.ny + cm) * P.nx + cx);
}

// global cell layer of a position along the slow axis (A.7 index, clamped), used by the slab exchange
__device__ __forceinline__ int cell_layer_global(const DevParams& P, const float4& x)
{
    const int cs = static_cast<int>((comp(x, P.axisS) - P.bmin[P.axisS]) / P.h);
    return max(min(cs, P.nzGlobal - 1), 0);
}

constexpr uint32_t kInvalidId = 0xffffffffu;

// nSlots >= P.n in slab mode: dead slots (id == kInvalidId: last step's ghosts) get the maximal key and sort behind
// the P.n live particles
__global__ void k_hash(const float4* __restrict__ pos, const uint32_t* __restrict__ id, uint32_t* __restrict__ keys,
                       uint32_t* __restrict__ vals, uint32_t nSlots, DevParams P, const DevState* st)
{
    if(st->skip) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= nSlots) return;
    keys[i] = id[i] == kInvalidId ? 0xffffffffu : cell_key(P, pos[i]);
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
// (1') counting sort by cell key -- the alternative to the radix passes (SF_SORT=count).  A cell key has up to 27
//      bits but only ~8 particles share one, and k_reorder re-ranks every cell by original id anyway, so the order
//      in which particles arrive inside a cell is irrelevant:
//        k_hash_count (key + arrival rank in its cell, warp-aggregated atomics on cellCnt[key])
//        -> k_cell_scan_reduce / k_radix_scan (tile sums) / k_cell_scan_apply (cellTab = {begin,end}, brick flags;
//           cellCnt zeroed for the next count)
//        -> k_count_scatter (slot permutation in key order)
//      One pass over the particles and one over the cells instead of three passes of (histogram, scan, scatter).
//      On a single GPU with device-resident state the first step is fused into the previous substep's integrate
//      kernel (k_visc_brick counts the cell of every new position as it writes it: count_into_cell below), so a
//      substep starts at the scan.
constexpr int CS_THREADS = 256;
constexpr int CS_ITEMS   = 8;
constexpr int CS_TILE    = CS_THREADS * CS_ITEMS;

// arrival rank of a particle in cell `key` (0xffffffff: not a live particle), one atomic per distinct key of the warp;
// called by all 32 lanes
__device__ __forceinline__ uint32_t count_into_cell(uint32_t* __restrict__ cellCnt, uint32_t key, bool live)
{
    const int      lane   = threadIdx.x & 31;
    const uint32_t peers  = __match_any_sync(0xffffffffu, key);
    const int      leader = __ffs(peers) - 1;
    uint32_t       base   = 0u;
    if(live && lane == leader) base = atomicAdd(&cellCnt[key], static_cast<uint32_t>(__popc(peers)));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}

__global__ void k_hash_count(const float4* __restrict__ pos, const uint32_t* __restrict__ id, uint32_t* __restrict__ keys,
                             uint32_t* __restrict__ ranks, uint32_t nSlots, DevParams P, uint32_t* __restrict__ cellCnt, const DevState* st)
{
    if(st->skip) return;
    const uint32_t i    = blockIdx.x * blockDim.x + threadIdx.x; // blockDim.x is a multiple of 32: whole warps reach the match
    const bool     live = i < nSlots && id[i] != kInvalidId;
    const uint32_t key  = live ? cell_key(P, pos[i]) : 0xffffffffu;
    const uint32_t rank = count_into_cell(cellCnt, key, live);
    if(i < nSlots) {
        keys[i]  = key;
        ranks[i] = rank;
    }
}

__global__ void __launch_bounds__(CS_THREADS)
k_cell_scan_reduce(const uint32_t* __restrict__ cellCnt, uint32_t ncells, uint32_t* __restrict__ tileSums, const DevState* st)
{
    if(st->skip) return;
    __shared__ uint32_t warpSum[CS_THREADS / 32];
    const uint32_t base = blockIdx.x * CS_TILE + threadIdx.x * CS_ITEMS;
    uint32_t       sum  = 0u;
#pragma unroll
    for(int r = 0; r < CS_ITEMS; ++r)
        if(base + r < ncells) sum += cellCnt[base + r];
    for(int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if((threadIdx.x & 31) == 0) warpSum[threadIdx.x >> 5] = sum;
    __syncthreads();
    if(threadIdx.x == 0) {
        uint32_t t = 0u;
        for(int w = 0; w < CS_THREADS / 32; ++w) t += warpSum[w];
        tileSums[blockIdx.x] = t;
    }
}

// tileSums have been exclusive-scanned in place (k_radix_scan with one row)
__global__ void __launch_bounds__(CS_THREADS)
k_cell_scan_apply(uint2* __restrict__ cellTab, uint32_t* __restrict__ cellCnt, uint32_t ncells, const uint32_t* __restrict__ tileSums,
                  uint32_t* __restrict__ brickFlag, DevParams P, const DevState* st)
{
    if(st->skip) return;
    __shared__ uint32_t warpSum[CS_THREADS / 32];
    const int      lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * CS_TILE + threadIdx.x * CS_ITEMS;
    uint32_t       cnt[CS_ITEMS];
    uint32_t       sum = 0u;
#pragma unroll
    for(int r = 0; r < CS_ITEMS; ++r) {
        cnt[r] = base + r < ncells ? cellCnt[base + r] : 0u;
        sum += cnt[r];
    }
    uint32_t x = sum; // inclusive scan of the per-thread sums over the warp, then over the warps
    for(int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if(lane >= o) x += y;
    }
    if(lane == 31) warpSum[wid] = x;
    __syncthreads();
    uint32_t off = tileSums[blockIdx.x] + x - sum;
    for(int w = 0; w < wid; ++w) off += warpSum[w];
#pragma unroll
    for(int r = 0; r < CS_ITEMS; ++r) {
        if(base + r < ncells) {
            const uint32_t c = cnt[r];
            cellTab[base + r] = c ? make_uint2(off, off + c) : make_uint2(0u, 0u);
            if(c) {
                cellCnt[base + r] = 0u; // ready for the next count
                const uint32_t key = base + r;
                const int      cx = static_cast<int>(key % static_cast<uint32_t>(P.nx));
                const int      t  = static_cast<int>(key / static_cast<uint32_t>(P.nx));
                const int      cy = t % P.ny, cz = t / P.ny;
                brickFlag[((cz / BZ) * P.nby + (cy / BY)) * P.nbx + (cx / BX)] = 1u; // == brick_of_key (sf_pairs.cuh)
            }
            off += c;
        }
    }
}

__global__ void k_count_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ ranks, const uint2* __restrict__ cellTab,
                                uint32_t* __restrict__ keysOut, uint32_t* __restrict__ slotsOut, uint32_t nSlots, const DevState* st)
{
    if(st->skip) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= nSlots) return;
    const uint32_t key = keys[i];
    if(key == 0xffffffffu) return; // dead slot (last substep's ghost)
    const uint32_t dst = cellTab[key].x + ranks[i];
    keysOut[dst]  = key;
    slotsOut[dst] = i;
}

// ------------------------------------------------------------------------------------------------
// (1) cell start/end tables + particle reorder
__global__ void k_clear_cells(uint4* __restrict__ tab, size_t nvec, const DevState* st)
{
    if(st->skip) return;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for(size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<size_t>(gridDim.x) * blockDim.x) tab[i] = z;
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

// sf_step_host variants (velocity upload overlapped with sort + density): k_reorder without the velocity gather ...
__global__ void k_reorder_pos(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                              const uint2* __restrict__ cellTab, const float4* __restrict__ posA,
                              const uint32_t* __restrict__ idA, float4* __restrict__ posB, uint32_t* __restrict__ idB,
                              uint32_t n, const DevState* st)
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
    x.w                = 0.f;
    posB[dst]          = x;
    idB[dst]           = myId;
}

// ... and the velocity gather straight from the uploaded xyz array (id = upload index) once it has arrived (the .w
// lane is filled by the force pass), fused with computeMaxVel (A.5) of the uploaded velocities
__global__ void k_gather_vel_host(const float* __restrict__ velXYZ, const uint32_t* __restrict__ idB, float4* __restrict__ velB,
                                  uint32_t n, DevState* st)
{
    if(st->skip) return;
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    float          m = FLT_MIN;
    if(p < n) {
        const size_t o = 3 * static_cast<size_t>(idB[p]);
        const float  vx = velXYZ[o], vy = velXYZ[o + 1], vz = velXYZ[o + 2];
        velB[p]        = make_float4(vx, vy, vz, 0.f);
        m              = fmaxf(m, (vy * vy + vx * vx) + vz * vz);
    }
    for(int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if((threadIdx.x & 31) == 0 && m > FLT_MIN) atomicMax(&st->maxv2Bits[st->step & 1u], __float_as_uint(m));
}

// the host's cell_coords_checked (sf_host.cpp) with the same float ops: inside the box on every axis, finite
__device__ __forceinline__ bool in_box(const DevParams& P, float x, float y, float z)
{
    const float c[3] = { x, y, z };
    const int   g[3] = { P.nx, P.axisS == 1 ? P.nzGlobal : P.ny, P.axisS == 2 ? P.nzGlobal : P.ny };
    for(int d = 0; d < 3; ++d) {
        const float t = (c[d] - P.bmin[d]) / P.h;
        if(!(t >= 0.0f) || !(t < static_cast<float>(g[d]))) return false; // also rejects NaN / inf
    }
    return true;
}

// sf_step_host, steady state: positions from the uploaded xyz array, id = upload index.  Validates the domain like
// sf_upload_particles does on the host (SF_DEVERR_DOMAIN -> SF_ERR_DOMAIN): the pair loops assume that a particle's
// unclamped cell equals its binned cell.
__global__ void k_pack_pos(const float* __restrict__ posXYZ, float4* __restrict__ pos, uint32_t* __restrict__ id, uint32_t n, DevParams P, DevState* st)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const float x = posXYZ[3 * static_cast<size_t>(i)], y = posXYZ[3 * static_cast<size_t>(i) + 1], z = posXYZ[3 * static_cast<size_t>(i) + 2];
    if(!in_box(P, x, y, z)) atomicOr(&st->errFlags, SF_DEVERR_DOMAIN);
    pos[i] = make_float4(x, y, z, 0.f);
    id[i]  = i;
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

__global__ void k_unpack_u32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ id, uint32_t* __restrict__ out, uint32_t n, uint32_t add)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p < n) out[id[p]] = src[p] + add;
}

__global__ void k_unpack_pressure(const float* __restrict__ rho, const uint32_t* __restrict__ id, float* __restrict__ out, DevParams P)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p < P.n) out[id[p]] = pressure_of(P, rho[p]);
}
} // namespace sf
