// sf_slab.cuh -- multi-GPU z-slab decomposition (included at the end of sf_api.cu; one translation unit).
//
// Nothing like this exists in the reference (single process, TBB only: SURVEY.md section 2b).  One solver = one
// process = one GPU = one slab of whole cell layers [zb, ze).  Because x is the fastest and z the slowest digit of
// the cell key, a slab and each of its ghost layers is a contiguous range of the sorted particle array.
//
// Every rank holds its own layers plus THREE ghost layers on each side: with positions/velocities of layers
// [zb-3, ze+3) a rank can compute density on [zb-2, ze+2), pressure force and v* on [zb-1, ze+1) and the XSPH sum +
// integration on its own [zb, ze) without any exchange inside the substep -- and bit-identically to a single-GPU
// run, because every particle sees the same neighbours in the same order (cells z->y->x, ascending global id).
// So there is exactly ONE exchange per substep, after integration:
//   * each rank sends the lower neighbour its own particles whose NEW layer is < zb'+3 (migrants and fresh ghosts in
//     one message) and the upper neighbour those with new layer >= ze'-3 (primes = next bounds);
//   * last step's ghosts are dropped (their slots get the invalid id and sort to the tail), received particles are
//     appended; the next substep's radix sort puts everything in place.
// Overlap: the XSPH/integrate kernel runs twice -- bricks within 5 layers of a slab face first, interior bricks
// second; the pack + NCCL exchange of the edge particles runs on a second stream while the interior bricks execute.
// The global dt needs max |v|^2 over all ranks: one in-stream ncclAllReduce(max) on the two uint32 slots.
// Load balance: cut planes start count-balanced and move by at most one layer per substep (decided identically on
// every rank from an all-gathered 8-word table), which the same exchange absorbs.
#pragma once
#include <dlfcn.h>
#include <nccl.h> // types and enums only; the library is dlopen'ed so that single-GPU users need no NCCL

namespace sf
{
constexpr int kGhostMax = 4; // ghost layers per side: 3 (density, force, XSPH), 4 with bCorrectDensity (+ Shepard pass)
inline int slab_edge(int ghost) { return ghost + 2; } // own layers per side whose particles may have to be exchanged (ghost + movement + cut shift)
constexpr int kMinThick = 6; // minimum slab thickness in layers
constexpr int kRowWords = 8;

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

inline NcclApi& nccl_api()
{
    static NcclApi api;
    if(api.lib) return api;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for(const char* n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if(api.lib) break;
    }
    if(!api.lib) return api;
#define SF_NCCL_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym))
    SF_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    SF_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    SF_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    SF_NCCL_SYM(Send, "ncclSend");
    SF_NCCL_SYM(Recv, "ncclRecv");
    SF_NCCL_SYM(GroupStart, "ncclGroupStart");
    SF_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    SF_NCCL_SYM(AllReduce, "ncclAllReduce");
    SF_NCCL_SYM(AllGather, "ncclAllGather");
    SF_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef SF_NCCL_SYM
    return api;
}

// ------------------------------------------------------------------------------------------------
// kernels
// layerStart[l] = first sorted slot whose local layer is >= l  (l = 0 .. nz), by binary search in the sorted keys
__global__ void k_layer_start(const uint32_t* __restrict__ keyB, uint32_t n, DevParams P, uint32_t* __restrict__ layerStart)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if(l > P.nz) return;
    const uint32_t target = static_cast<uint32_t>(l) * static_cast<uint32_t>(P.nx * P.ny);
    uint32_t       lo = 0, hi = n;
    while(lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if(keyB[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    layerStart[l] = lo;
}

// every slot is dead until the integrate kernel revives the particles this rank owns
__global__ void k_fill_u32(uint32_t* __restrict__ a, uint32_t n, uint32_t v)
{
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a[i] = v;
}

// Pack the freshly integrated own particles near the slab faces for the neighbours.  zbN/zeN = NEXT bounds (global).
// send buffers: 2 float4 per particle {x, y, z, id bits}, {vx, vy, vz, 0}.  counters[0/1] = lower/upper count.
__global__ void k_slab_pack(const float4* __restrict__ posA, const float4* __restrict__ velA, const uint32_t* __restrict__ idA,
                            const uint32_t* __restrict__ layerStart, DevParams P, int zbN, int zeN, int ghost, int hasLower, int hasUpper,
                            float4* __restrict__ sendLo, float4* __restrict__ sendHi, uint32_t cap, uint32_t* __restrict__ counters)
{
    const uint32_t ownB = layerStart[P.zOwnLo], ownE = layerStart[P.zOwnHi];
    const uint32_t loE  = layerStart[min(P.zOwnLo + P.zEdge, P.zOwnHi)];
    const uint32_t hiB  = max(layerStart[max(P.zOwnHi - P.zEdge, P.zOwnLo)], loE);
    const uint32_t nLo = loE - ownB, nHi = ownE - hiB;
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nLo + nHi; i += gridDim.x * blockDim.x) {
        const uint32_t p  = i < nLo ? ownB + i : hiB + (i - nLo);
        const uint32_t id = idA[p];
        if(id == kInvalidId) continue;
        const float4 x  = posA[p];
        const float4 v  = velA[p];
        const int    cz = cell_layer_global(P, x);
        if(hasLower && cz < zbN + ghost) {
            const uint32_t k = atomicAdd(&counters[0], 1u);
            if(k < cap) {
                sendLo[2 * k]     = make_float4(x.x, x.y, x.z, __uint_as_float(id));
                sendLo[2 * k + 1] = make_float4(v.x, v.y, v.z, 0.f);
            }
        }
        if(hasUpper && cz >= zeN - ghost) {
            const uint32_t k = atomicAdd(&counters[1], 1u);
            if(k < cap) {
                sendHi[2 * k]     = make_float4(x.x, x.y, x.z, __uint_as_float(id));
                sendHi[2 * k + 1] = make_float4(v.x, v.y, v.z, 0.f);
            }
        }
    }
}

// this rank's row of the all-gathered table: {sendLo, sendHi, nOwn, firstLayerCount, lastLayerCount, ownBegin, ownEnd, 0}
__global__ void k_slab_row(const uint32_t* __restrict__ layerStart, const uint32_t* __restrict__ counters, DevParams P, uint32_t* __restrict__ row)
{
    const uint32_t ownB = layerStart[P.zOwnLo], ownE = layerStart[P.zOwnHi];
    row[0] = counters[0];
    row[1] = counters[1];
    row[2] = ownE - ownB;
    row[3] = layerStart[P.zOwnLo + 1] - ownB;
    row[4] = ownE - layerStart[P.zOwnHi - 1];
    row[5] = ownB;
    row[6] = ownE;
    row[7] = 0u;
}

__global__ void k_slab_unpack(const float4* __restrict__ recv, uint32_t count, float4* __restrict__ pos, float4* __restrict__ vel, uint32_t* __restrict__ id)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= count) return;
    const float4 a = recv[2 * i], b = recv[2 * i + 1];
    pos[i] = make_float4(a.x, a.y, a.z, 0.f);
    vel[i] = make_float4(b.x, b.y, b.z, 0.f);
    id[i]  = __float_as_uint(a.w);
}

// ---- host-owned slab state (sf_step_host_owned / sf_download_owned): every rank's OWNED particles live on the host
// between substeps; the ghost particles of the next substep stay resident (they came with the last exchange).
// counters: [0] = resident owned particles dropped, [1] = uploaded particles outside the box or this rank's layers,
// [2] = owned particles gathered.
__global__ void k_owned_reset_maxvel(DevState* st) { st->maxv2Bits[st->step & 1u] = __float_as_uint(FLT_MIN); }

// the resident copies of this rank's own particles give way to the host's
__global__ void k_owned_kill(const float4* __restrict__ posA, uint32_t* __restrict__ idA, uint32_t nSlots, DevParams P, int zb, int ze,
                             uint32_t* __restrict__ counters)
{
    uint32_t killed = 0u;
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nSlots; i += gridDim.x * blockDim.x) {
        if(idA[i] == kInvalidId) continue;
        const int cz = cell_layer_global(P, posA[i]);
        if(cz >= zb && cz < ze) {
            idA[i] = kInvalidId;
            ++killed;
        }
    }
    for(int o = 16; o > 0; o >>= 1) killed += __shfl_xor_sync(0xffffffffu, killed, o);
    if((threadIdx.x & 31) == 0 && killed) atomicAdd(&counters[0], killed);
}

// the host's owned set appended behind the resident slots, with computeMaxVel (A.5) of the uploaded velocities
__global__ void k_owned_append(const float* __restrict__ posXYZ, const float* __restrict__ velXYZ, const uint32_t* __restrict__ ids, uint32_t m,
                               float4* __restrict__ pos, float4* __restrict__ vel, uint32_t* __restrict__ id, DevParams P, int zb, int ze,
                               uint32_t* __restrict__ counters, DevState* st)
{
    float    mx  = FLT_MIN;
    uint32_t bad = 0u;
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const float4 x = make_float4(posXYZ[3 * static_cast<size_t>(i)], posXYZ[3 * static_cast<size_t>(i) + 1], posXYZ[3 * static_cast<size_t>(i) + 2], 0.f);
        const float4 v = make_float4(velXYZ[3 * static_cast<size_t>(i)], velXYZ[3 * static_cast<size_t>(i) + 1], velXYZ[3 * static_cast<size_t>(i) + 2], 0.f);
        const int    cz = cell_layer_global(P, x);
        if(!in_box(P, x.x, x.y, x.z) || cz < zb || cz >= ze) ++bad;
        pos[i] = x;
        vel[i] = v;
        id[i]  = ids[i];
        mx     = fmaxf(mx, (v.y * v.y + v.x * v.x) + v.z * v.z);
    }
    for(int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if((threadIdx.x & 31) == 0) {
        if(mx > FLT_MIN) atomicMax(&st->maxv2Bits[st->step & 1u], __float_as_uint(mx));
        if(bad) atomicAdd(&counters[1], bad);
    }
}

// compaction of the particles this rank owns ([zb, ze) by the CURRENT cut planes) into xyz staging arrays; the
// order is whatever the atomics give (the ids travel with the particles)
__global__ void k_owned_gather(const float4* __restrict__ posA, const float4* __restrict__ velA, const uint32_t* __restrict__ idA, uint32_t nSlots,
                               DevParams P, int zb, int ze, float* __restrict__ posXYZ, float* __restrict__ velXYZ, uint32_t* __restrict__ ids,
                               uint32_t cap, uint32_t* __restrict__ counters)
{
    const int lane = threadIdx.x & 31;
    for(uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < nSlots; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + lane;
        bool           own = false;
        float4         x = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t       pid = kInvalidId;
        if(i < nSlots) {
            pid = idA[i];
            if(pid != kInvalidId) {
                x            = posA[i];
                const int cz = cell_layer_global(P, x);
                own          = cz >= zb && cz < ze;
            }
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, own);
        if(!mask) continue;
        uint32_t k0 = 0u;
        if(lane == 0) k0 = atomicAdd(&counters[2], static_cast<uint32_t>(__popc(mask)));
        k0 = __shfl_sync(0xffffffffu, k0, 0);
        if(own) {
            const uint32_t k = k0 + __popc(mask & ((1u << lane) - 1u));
            if(k < cap) {
                const float4 v = velA[i];
                posXYZ[3 * static_cast<size_t>(k)]     = x.x;
                posXYZ[3 * static_cast<size_t>(k) + 1] = x.y;
                posXYZ[3 * static_cast<size_t>(k) + 2] = x.z;
                velXYZ[3 * static_cast<size_t>(k)]     = v.x;
                velXYZ[3 * static_cast<size_t>(k) + 1] = v.y;
                velXYZ[3 * static_cast<size_t>(k) + 2] = v.z;
                ids[k] = pid;
            }
        }
    }
}

__global__ void k_pack_upload_ids(const float* __restrict__ posXYZ, const float* __restrict__ velXYZ, const uint32_t* __restrict__ ids,
                                  float4* __restrict__ pos, float4* __restrict__ vel, uint32_t* __restrict__ id, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    pos[i] = make_float4(posXYZ[3 * i], posXYZ[3 * i + 1], posXYZ[3 * i + 2], 0.f);
    vel[i] = velXYZ ? make_float4(velXYZ[3 * i], velXYZ[3 * i + 1], velXYZ[3 * i + 2], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    id[i]  = ids[i];
}
} // namespace sf
