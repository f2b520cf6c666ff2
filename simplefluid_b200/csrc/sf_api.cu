// sf_api.cu -- C-ABI (include/sf_b200.h) over the sm_100a SPH pipeline.
//
// One sf_solver = one GPU = the object behind QtSPHSolver (Include/QtSPHSolver.h:27-36).
// The substep (advanceFrame, EXE@0x140016810) is a fixed sequence of kernel launches on one stream
// with no host synchronisation: dt, the frame-time accumulator and the max-velocity reduction all
// live in a DevState block on the device.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "sf_internal.h"
#include "sf_pairs.cuh"

using namespace sf;

namespace
{
thread_local std::string g_createError;

enum KernelId {
    K_BEGIN = 0, K_HASH, K_RADIX_HIST, K_RADIX_SCAN, K_RADIX_SCATTER, K_CLEAR_CELLS, K_CELL_BOUNDS, K_BRICK_COMPACT, K_REORDER,
    K_DENSITY, K_CORRECT_DENSITY, K_FORCE, K_VISC_INTEGRATE, K_MARSHAL, K_COUNT
};
const char* const kKernelNames[K_COUNT] = {
    "k_begin_step", "k_hash", "k_radix_hist", "k_radix_scan", "k_radix_scatter", "k_clear_cells", "k_cell_bounds", "k_brick_compact",
    "k_reorder", "k_density", "k_correct_density", "k_force", "k_visc_integrate", "k_marshal"
};

struct PendingEvent {
    int         id;
    cudaEvent_t a, b;
};
} // namespace

struct sf_solver {
    sf_params    params{};
    int          device = 0;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    int          numSMs = 148;
    std::string  lastError;

    uint32_t   n = 0, cap = 0, npad = 0;
    int        kmax = 96;
    int32_t    grid[3] = { 0, 0, 0 };
    uint64_t   ncells = 0, cellCap = 0;
    bool       ready = false, uploaded = false, capture = false;
    DevBuffers B{};
    DevParams  P{};
    KernelTables tables;
    std::vector<float> walls[6];
    bool         wallsSet = false;
    uint32_t     bndStride = 0;
    float*       stage = nullptr; // device staging for host<->device marshalling (2 x 12 B x cap)
    size_t       stageBytes = 0;
    DevState*    hostState = nullptr; // pinned
    uint32_t     radixBlocks = 0;
    int          sortPasses = 0, sortBits[4] = { 0, 0, 0, 0 };
    int          occDensity = 1, occForce = 1, occVisc = 1;
    uint32_t     numBricks = 0, brickCap = 0;

    // measurement
    bool                      profiling = false;
    double                    profMs[K_COUNT] = {};
    uint64_t                  profLaunches[K_COUNT] = {};
    std::vector<PendingEvent> pending;
    std::vector<cudaEvent_t>  eventPool;
    uint64_t                  launches = 0;
    cudaEvent_t               timerA = nullptr, timerB = nullptr;
};

namespace
{
int fail(sf_solver* s, int code, const std::string& msg)
{
    if(s) s->lastError = msg;
    else g_createError = msg;
    return code;
}

#define SF_CUDA(s, call)                                                                                   \
    do {                                                                                                   \
        cudaError_t _e = (call);                                                                           \
        if(_e != cudaSuccess)                                                                              \
            return fail((s), _e == cudaErrorMemoryAllocation ? SF_ERR_OOM : SF_ERR_CUDA,                    \
                        std::string(#call) + ": " + cudaGetErrorString(_e));                               \
    } while(0)

template<class T>
cudaError_t dev_alloc(T*& p, size_t count)
{
    if(p) {
        cudaFree(p);
        p = nullptr;
    }
    if(count == 0) return cudaSuccess;
    return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
}

cudaEvent_t take_event(sf_solver* s)
{
    if(!s->eventPool.empty()) {
        cudaEvent_t e = s->eventPool.back();
        s->eventPool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void fold_pending(sf_solver* s)
{
    if(s->pending.empty()) return;
    cudaEventSynchronize(s->pending.back().b);
    for(auto& pe : s->pending) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, pe.a, pe.b);
        s->profMs[pe.id] += ms;
        s->profLaunches[pe.id] += 1;
        s->eventPool.push_back(pe.a);
        s->eventPool.push_back(pe.b);
    }
    s->pending.clear();
}

struct LaunchScope {
    sf_solver*  s;
    int         id;
    cudaEvent_t a = nullptr;
    LaunchScope(sf_solver* s_, int id_) : s(s_), id(id_)
    {
        s->launches++;
        if(s->profiling) {
            a = take_event(s);
            cudaEventRecord(a, s->stream);
        }
    }
    ~LaunchScope()
    {
        if(s->profiling) {
            cudaEvent_t b = take_event(s);
            cudaEventRecord(b, s->stream);
            s->pending.push_back({ id, a, b });
            if(s->pending.size() >= 8192) fold_pending(s);
        }
    }
};

inline uint32_t cdiv(uint64_t a, uint32_t b) { return static_cast<uint32_t>((a + b - 1) / b); }


void fill_dev_params(sf_solver* s)
{
    const sf_params& p = s->params;
    DevParams&       P = s->P;
    for(int d = 0; d < 3; ++d) {
        P.bmin[d] = p.boxMin[d];
        P.bmax[d] = p.boxMax[d];
    }
    P.h            = p.kernelRadius;
    P.h2           = p.kernelRadiusSqr;
    P.r            = p.particleRadius;
    P.mass         = p.particleMass;
    P.stiffness    = p.pressureStiffness;
    P.viscosity    = p.viscosity;
    P.restitution  = p.boundaryRestitution;
    P.rho0         = p.restDensity;
    P.attractRatio = p.attractivePressureRatio;
    P.rhoMin       = static_cast<float>(static_cast<double>(p.restDensity) * 0.1);
    P.rhoMax       = static_cast<float>(static_cast<double>(p.restDensity) * 10.0);
    P.Wzero        = s->tables.Wzero;
    P.invStep      = s->tables.invStep;
    P.radius2      = s->tables.radius2;
    P.dtMin        = p.defaultTimestep * 0.1f;
    P.dtMax        = p.defaultTimestep * 10.0f;
    P.nx = s->grid[0];
    P.ny = s->grid[1];
    P.nz = s->grid[2];
    P.useBoundary    = p.bUseBoundaryParticles ? 1 : 0;
    P.attractive     = p.bUseAttractivePressure ? 1 : 0;
    P.correctDensity = p.bCorrectDensity ? 1 : 0;
    P.capture        = s->capture ? 1 : 0;
    P.n    = s->n;
    P.npad = s->npad;
    P.kmax = s->kmax;
    P.nbx  = (s->grid[0] + BX - 1) / BX;
    P.nby  = (s->grid[1] + BY - 1) / BY;
    P.nbz  = (s->grid[2] + BZ - 1) / BZ;
    P.numBricks = s->numBricks;
    for(int w = 0; w < 6; ++w) P.nbnd[w] = P.useBoundary ? static_cast<uint32_t>(s->walls[w].size() / 3) : 0u;
    P.bndStride = s->bndStride;
}

int ensure_particle_capacity(sf_solver* s, uint32_t n)
{
    if(n <= s->cap && s->B.posA) return SF_OK;
    const uint32_t cap  = n;
    const uint32_t npad = (cap + 127u) & ~127u;
    SF_CUDA(s, dev_alloc(s->B.posA, npad));
    SF_CUDA(s, dev_alloc(s->B.velA, npad));
    SF_CUDA(s, dev_alloc(s->B.posB, npad));
    SF_CUDA(s, dev_alloc(s->B.velB, npad));
    SF_CUDA(s, dev_alloc(s->B.idA, npad));
    SF_CUDA(s, dev_alloc(s->B.idB, npad));
    for(int i = 0; i < 2; ++i) {
        SF_CUDA(s, dev_alloc(s->B.keys[i], npad));
        SF_CUDA(s, dev_alloc(s->B.vals[i], npad));
    }
    SF_CUDA(s, dev_alloc(s->B.rho, npad));
    SF_CUDA(s, dev_alloc(s->B.rho2, npad));
    SF_CUDA(s, dev_alloc(s->B.accel, npad));
    SF_CUDA(s, dev_alloc(s->B.nbrCnt, npad));
    SF_CUDA(s, dev_alloc(s->B.nbrL, static_cast<size_t>(npad) * s->kmax));
    s->radixBlocks = cdiv(npad, RS_TILE);
    SF_CUDA(s, dev_alloc(s->B.radixCounts, static_cast<size_t>(s->radixBlocks) * RS_MAXRADIX));
    SF_CUDA(s, dev_alloc(s->B.radixTotals, RS_MAXRADIX));
    s->stageBytes = static_cast<size_t>(npad) * 24;
    SF_CUDA(s, dev_alloc(s->stage, s->stageBytes / sizeof(float)));
    s->cap  = cap;
    s->npad = npad;
    return SF_OK;
}

int enqueue_substep(sf_solver* s)
{
    DevBuffers&     B  = s->B;
    const DevParams P  = s->P;
    cudaStream_t    st = s->stream;
    const uint32_t  n  = s->n;
    const uint32_t  gridN = cdiv(n, 256);
    {
        LaunchScope ls(s, K_BEGIN);
        k_begin_step<<<1, 1, 0, st>>>(B.state, P);
    }
    if(n == 0) {
        SF_CUDA(s, cudaGetLastError());
        return SF_OK;
    }
    {
        LaunchScope ls(s, K_HASH);
        k_hash<<<gridN, 256, 0, st>>>(B.posA, B.keys[0], B.vals[0], P, B.state);
    }
    const uint32_t nb = cdiv(n, RS_TILE);
    int            cur = 0, shift = 0;
    for(int pass = 0; pass < s->sortPasses; ++pass) {
        const int radix = 1 << s->sortBits[pass];
        {
            LaunchScope ls(s, K_RADIX_HIST);
            k_radix_hist<<<nb, RS_THREADS, 0, st>>>(B.keys[cur], n, shift, radix, B.radixCounts, nb, B.state);
        }
        {
            LaunchScope ls(s, K_RADIX_SCAN);
            k_radix_scan<<<radix, 1024, 0, st>>>(B.radixCounts, nb, B.radixTotals, B.state);
        }
        {
            LaunchScope ls(s, K_RADIX_SCATTER);
            k_radix_scatter<<<nb, RS_THREADS, 0, st>>>(B.keys[cur], B.vals[cur], B.keys[cur ^ 1], B.vals[cur ^ 1], n, shift, radix,
                                                        B.radixCounts, nb, B.radixTotals, B.state);
        }
        shift += s->sortBits[pass];
        cur ^= 1;
    }
    B.keyB = B.keys[cur];
    {
        LaunchScope  ls(s, K_CLEAR_CELLS);
        const size_t nvec = (s->ncells * sizeof(uint2) + 15) / 16;
        k_clear_cells<<<std::min<uint32_t>(cdiv(nvec, 256), s->numSMs * 16), 256, 0, st>>>(reinterpret_cast<uint4*>(B.cellTab), nvec, B.state);
    }
    {
        LaunchScope ls(s, K_CELL_BOUNDS);
        k_cell_bounds_bricks<<<gridN, 256, 0, st>>>(B.keyB, n, B.cellTab, B.brickFlag, P, B.state);
    }
    {
        LaunchScope ls(s, K_BRICK_COMPACT);
        k_brick_compact<<<1, 1024, 0, st>>>(B.brickFlag, B.brickList, s->numBricks, B.state);
    }
    {
        LaunchScope ls(s, K_REORDER);
        k_reorder<<<gridN, 256, 0, st>>>(B.keyB, B.vals[cur], B.cellTab, B.posA, B.velA, B.idA, B.posB, B.velB, B.idB, n, B.state);
    }
    const uint32_t pairGrid = std::min<uint32_t>(s->numBricks, static_cast<uint32_t>(s->numSMs) * 2u);
    {
        LaunchScope ls(s, K_DENSITY);
        k_density_brick<<<std::min<uint32_t>(pairGrid, s->numSMs * s->occDensity), kBrickThreads, kSmemDensity, st>>>(B, P);
    }
    if(P.correctDensity) {
        LaunchScope ls(s, K_CORRECT_DENSITY);
        k_correct_density<<<gridN, 256, 0, st>>>(B, P);
        k_density_terms<<<gridN, 256, 0, st>>>(B, P);
    }
    {
        LaunchScope ls(s, K_FORCE);
        k_force_brick<<<std::min<uint32_t>(pairGrid, s->numSMs * s->occForce), kBrickThreads, kSmemPair, st>>>(B, P);
    }
    {
        LaunchScope ls(s, K_VISC_INTEGRATE);
        k_visc_brick<<<std::min<uint32_t>(pairGrid, s->numSMs * s->occVisc), kBrickThreads, kSmemPair, st>>>(B, P);
    }
    SF_CUDA(s, cudaGetLastError());
    return SF_OK;
}

int read_state(sf_solver* s)
{
    SF_CUDA(s, cudaMemcpyAsync(s->hostState, s->B.state, sizeof(DevState), cudaMemcpyDeviceToHost, s->stream));
    SF_CUDA(s, cudaStreamSynchronize(s->stream));
    if(s->hostState->errFlags) {
        char buf[256];
        std::snprintf(buf, sizeof(buf),
                      "device consistency check failed (flags 0x%x): neighbour list capacity exceeded (kmax=%d per particle)",
                      s->hostState->errFlags, s->kmax);
        return fail(s, SF_ERR_STATE, buf);
    }
    return SF_OK;
}

int require_ready(sf_solver* s)
{
    if(!s) return SF_ERR_INVALID;
    if(!s->uploaded) return fail(s, SF_ERR_INVALID, "no particles uploaded");
    if(!s->ready) return fail(s, SF_ERR_INVALID, "sf_make_ready has not been called");
    return SF_OK;
}
} // namespace

// =================================================================================================
extern "C" {

int sf_params_default(sf_params* p)
{
    if(!p) return SF_ERR_INVALID;
    params_default(*p);
    return SF_OK;
}

int sf_params_update(sf_params* p)
{
    if(!p) return SF_ERR_INVALID;
    params_update(*p);
    return SF_OK;
}

int sf_params_set_resolution(sf_params* p, float resolution)
{
    if(!p || !(resolution > 0.f)) return SF_ERR_INVALID;
    p->kernelRadius = 2.0f / resolution; // Source/Controller.cpp:55
    params_update(*p);
    return SF_OK;
}

int sf_scene_generate(const sf_params* p, int scene, float* pos_xyz, uint64_t cap, uint64_t* n_out)
{
    if(!p || scene < 0 || scene > 3) return SF_ERR_INVALID;
    const uint64_t n = scene_generate(*p, scene, pos_xyz, cap);
    if(n_out) *n_out = n;
    return SF_OK;
}

int sf_build_tables(const sf_params* p, float* cubic_w10001, float* spiky_grad10001, float* consts3)
{
    if(!p) return SF_ERR_INVALID;
    KernelTables t;
    build_tables(p->kernelRadius, t);
    if(cubic_w10001) std::memcpy(cubic_w10001, t.cubicW.data(), sizeof(float) * kTableEntries);
    if(spiky_grad10001) std::memcpy(spiky_grad10001, t.spikyGrad.data(), sizeof(float) * kTableEntries);
    if(consts3) {
        consts3[0] = t.Wzero;
        consts3[1] = t.radius2;
        consts3[2] = t.invStep;
    }
    return SF_OK;
}

int sf_boundary_generate(const sf_params* p, uint32_t seed, int wall, float* xyz, uint32_t cap, uint32_t* n_out)
{
    if(!p || wall < 0 || wall > 5) return SF_ERR_INVALID;
    std::vector<float> walls[6];
    generate_boundary(*p, seed, walls);
    const uint32_t n = static_cast<uint32_t>(walls[wall].size() / 3);
    if(n_out) *n_out = n;
    if(xyz) std::memcpy(xyz, walls[wall].data(), sizeof(float) * 3 * std::min(n, cap));
    return SF_OK;
}

int sf_create(const sf_params* p, int device, sf_solver** out)
{
    if(!p || !out) return fail(nullptr, SF_ERR_INVALID, "null argument");
    *out = nullptr;
    int         count = 0;
    cudaError_t e     = cudaGetDeviceCount(&count);
    if(e != cudaSuccess || count == 0)
        return fail(nullptr, SF_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
    if(device < 0 || device >= count) return fail(nullptr, SF_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop{};
    SF_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    if(prop.major < 10) return fail(nullptr, SF_ERR_CUDA, std::string("device ") + prop.name + " is not sm_100 class; kernels are built for sm_100a only");
    SF_CUDA(nullptr, cudaSetDevice(device));
    sf_solver* s = new sf_solver();
    s->params    = *p;
    s->device    = device;
    s->numSMs    = prop.multiProcessorCount;
    e            = cudaStreamCreateWithFlags(&s->ownStream, cudaStreamNonBlocking);
    if(e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->B.state), sizeof(DevState));
    if(e == cudaSuccess) e = cudaMemset(s->B.state, 0, sizeof(DevState));
    if(e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&s->hostState), sizeof(DevState));
    if(e == cudaSuccess) e = cudaEventCreate(&s->timerA);
    if(e == cudaSuccess) e = cudaEventCreate(&s->timerB);
    if(e == cudaSuccess) e = cudaFuncSetAttribute(k_density_brick, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemDensity));
    if(e == cudaSuccess) e = cudaFuncSetAttribute(k_force_brick, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPair));
    if(e == cudaSuccess) e = cudaFuncSetAttribute(k_visc_brick, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPair));
    if(e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->occDensity, k_density_brick, kBrickThreads, kSmemDensity);
    if(e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->occForce, k_force_brick, kBrickThreads, kSmemPair);
    if(e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->occVisc, k_visc_brick, kBrickThreads, kSmemPair);
    if(e != cudaSuccess) {
        const std::string msg = std::string("sf_create: ") + cudaGetErrorString(e);
        sf_destroy(s);
        return fail(nullptr, SF_ERR_CUDA, msg);
    }
    s->stream = s->ownStream;
    s->occDensity = std::max(s->occDensity, 1);
    s->occForce   = std::max(s->occForce, 1);
    s->occVisc    = std::max(s->occVisc, 1);
    *out = s;
    return SF_OK;
}

void sf_destroy(sf_solver* s)
{
    if(!s) return;
    cudaSetDevice(s->device);
    if(s->stream) cudaStreamSynchronize(s->stream);
    fold_pending(s);
    for(auto e : s->eventPool) cudaEventDestroy(e);
    DevBuffers& B = s->B;
    cudaFree(B.posA); cudaFree(B.velA); cudaFree(B.posB); cudaFree(B.velB); cudaFree(B.idA); cudaFree(B.idB);
    for(int i = 0; i < 2; ++i) { cudaFree(B.keys[i]); cudaFree(B.vals[i]); }
    cudaFree(B.cellTab); cudaFree(B.rho); cudaFree(B.rho2); cudaFree(B.accel); cudaFree(B.nbrL); cudaFree(B.nbrCnt); cudaFree(B.brickFlag); cudaFree(B.brickList);
    cudaFree(B.tabW); cudaFree(B.tabG); cudaFree(B.bnd); cudaFree(B.radixCounts); cudaFree(B.radixTotals); cudaFree(B.state);
    cudaFree(s->stage);
    if(s->hostState) cudaFreeHost(s->hostState);
    if(s->timerA) cudaEventDestroy(s->timerA);
    if(s->timerB) cudaEventDestroy(s->timerB);
    if(s->ownStream) cudaStreamDestroy(s->ownStream);
    delete s;
}

const char* sf_last_error(sf_solver* s) { return s ? s->lastError.c_str() : g_createError.c_str(); }

int sf_set_params(sf_solver* s, const sf_params* p)
{
    if(!s || !p) return SF_ERR_INVALID;
    s->params = *p;
    s->ready  = false; // tables / grid depend on kernelRadius: makeReady again (Simulator.cpp:42)
    return SF_OK;
}

int sf_get_params(sf_solver* s, sf_params* p)
{
    if(!s || !p) return SF_ERR_INVALID;
    *p = s->params;
    return SF_OK;
}

int sf_set_stream(sf_solver* s, void* cuda_stream)
{
    if(!s) return SF_ERR_INVALID;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    s->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s->ownStream;
    return SF_OK;
}

int sf_upload_particles(sf_solver* s, const float* pos_xyz, const float* vel_xyz, uint32_t n)
{
    if(!s || (!pos_xyz && n)) return SF_ERR_INVALID;
    SF_CUDA(s, cudaSetDevice(s->device));
    // domain check: the pair loops assume the unclamped cell of A.6 equals the binned cell of A.7
    int32_t g[3];
    grid_dims(s->params, g);
    for(uint32_t i = 0; i < n; ++i) {
        int32_t c[3];
        if(!cell_coords_checked(s->params, g, pos_xyz + 3 * static_cast<size_t>(i), c)) {
            char buf[160];
            std::snprintf(buf, sizeof(buf), "particle %u at (%g, %g, %g) lies outside the simulation box", i, pos_xyz[3 * i], pos_xyz[3 * i + 1], pos_xyz[3 * i + 2]);
            return fail(s, SF_ERR_DOMAIN, buf);
        }
    }
    int rc = ensure_particle_capacity(s, n);
    if(rc) return rc;
    s->n = n;
    if(n) {
        float* dpos = s->stage;
        float* dvel = s->stage + 3 * static_cast<size_t>(s->npad);
        SF_CUDA(s, cudaMemcpyAsync(dpos, pos_xyz, static_cast<size_t>(n) * 12, cudaMemcpyHostToDevice, s->stream));
        if(vel_xyz) SF_CUDA(s, cudaMemcpyAsync(dvel, vel_xyz, static_cast<size_t>(n) * 12, cudaMemcpyHostToDevice, s->stream));
        LaunchScope ls(s, K_MARSHAL);
        k_pack_upload<<<cdiv(n, 256), 256, 0, s->stream>>>(dpos, vel_xyz ? dvel : nullptr, s->B.posA, s->B.velA, s->B.idA, n);
        SF_CUDA(s, cudaMemcpyAsync(s->B.idB, s->B.idA, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice, s->stream));
    }
    SF_CUDA(s, cudaStreamSynchronize(s->stream));
    s->uploaded = true;
    s->ready    = false;
    return SF_OK;
}

int sf_num_particles(sf_solver* s, uint32_t* n_out)
{
    if(!s || !n_out) return SF_ERR_INVALID;
    *n_out = s->n;
    return SF_OK;
}

static int download_xyz(sf_solver* s, const float4* src, float* out)
{
    if(!s || !out) return SF_ERR_INVALID;
    if(!s->uploaded) return fail(s, SF_ERR_INVALID, "no particles uploaded");
    SF_CUDA(s, cudaSetDevice(s->device));
    if(s->n == 0) return SF_OK;
    {
        LaunchScope ls(s, K_MARSHAL);
        k_unpack_xyz<<<cdiv(s->n, 256), 256, 0, s->stream>>>(src, s->B.idA, s->stage, s->n);
    }
    SF_CUDA(s, cudaMemcpyAsync(out, s->stage, static_cast<size_t>(s->n) * 12, cudaMemcpyDeviceToHost, s->stream));
    SF_CUDA(s, cudaStreamSynchronize(s->stream));
    return SF_OK;
}

int sf_download_positions(sf_solver* s, float* pos_xyz) { return download_xyz(s, s ? s->B.posA : nullptr, pos_xyz); }
int sf_download_velocities(sf_solver* s, float* vel_xyz) { return download_xyz(s, s ? s->B.velA : nullptr, vel_xyz); }

int sf_generate_boundary(sf_solver* s, uint32_t seed)
{
    if(!s) return SF_ERR_INVALID;
    generate_boundary(s->params, seed, s->walls);
    s->wallsSet = true;
    s->ready    = false;
    return SF_OK;
}

int sf_set_boundary_particles(sf_solver* s, int wall, const float* xyz, uint32_t n)
{
    if(!s || wall < 0 || wall > 5 || (!xyz && n)) return SF_ERR_INVALID;
    s->walls[wall].assign(xyz, xyz + 3 * static_cast<size_t>(n));
    s->wallsSet = true;
    s->ready    = false;
    return SF_OK;
}

int sf_get_boundary_particles(sf_solver* s, int wall, float* xyz, uint32_t cap, uint32_t* n_out)
{
    if(!s || wall < 0 || wall > 5) return SF_ERR_INVALID;
    const uint32_t n = static_cast<uint32_t>(s->walls[wall].size() / 3);
    if(n_out) *n_out = n;
    if(xyz) std::memcpy(xyz, s->walls[wall].data(), sizeof(float) * 3 * std::min(n, cap));
    return SF_OK;
}

int sf_make_ready(sf_solver* s)
{
    if(!s) return SF_ERR_INVALID;
    if(!s->uploaded) return fail(s, SF_ERR_INVALID, "sf_make_ready: upload particles first");
    SF_CUDA(s, cudaSetDevice(s->device));
    params_update(s->params);
    build_tables(s->params.kernelRadius, s->tables);
    grid_dims(s->params, s->grid);
    s->ncells = static_cast<uint64_t>(s->grid[0]) * s->grid[1] * s->grid[2];
    if(s->ncells == 0 || s->ncells >= (1ull << 31)) return fail(s, SF_ERR_INVALID, "grid has no cells or more than 2^31 cells");
    if(s->params.bUseBoundaryParticles && !s->wallsSet) {
        generate_boundary(s->params, 0u, s->walls); // the reference seeds from std::random_device; we default to seed 0
        s->wallsSet = true;
    }
    if(s->ncells > s->cellCap || !s->B.cellTab) {
        SF_CUDA(s, dev_alloc(s->B.cellTab, s->ncells + 2));
        s->cellCap = s->ncells;
    }
    {
        const uint32_t nb = static_cast<uint32_t>((s->grid[0] + BX - 1) / BX) * ((s->grid[1] + BY - 1) / BY) * ((s->grid[2] + BZ - 1) / BZ);
        if(nb > s->brickCap || !s->B.brickFlag) {
            SF_CUDA(s, dev_alloc(s->B.brickFlag, nb + 1));
            SF_CUDA(s, dev_alloc(s->B.brickList, nb + 1));
            s->brickCap = nb;
        }
        SF_CUDA(s, cudaMemsetAsync(s->B.brickFlag, 0, sizeof(uint32_t) * (nb + 1), s->stream));
        s->numBricks = nb;
    }
    if(!s->B.tabW) {
        SF_CUDA(s, dev_alloc(s->B.tabW, kTableEntries + 3));
        SF_CUDA(s, dev_alloc(s->B.tabG, kTableEntries + 3));
    }
    SF_CUDA(s, cudaMemcpyAsync(s->B.tabW, s->tables.cubicW.data(), sizeof(float) * kTableEntries, cudaMemcpyHostToDevice, s->stream));
    SF_CUDA(s, cudaMemcpyAsync(s->B.tabG, s->tables.spikyGrad.data(), sizeof(float) * kTableEntries, cudaMemcpyHostToDevice, s->stream));
    // wall particles as float4 [6][stride]
    uint32_t maxWall = 1;
    for(int w = 0; w < 6; ++w) maxWall = std::max<uint32_t>(maxWall, static_cast<uint32_t>(s->walls[w].size() / 3));
    s->bndStride = (maxWall + 3u) & ~3u;
    std::vector<float4> bnd(static_cast<size_t>(6) * s->bndStride, make_float4(0.f, 0.f, 0.f, 0.f));
    for(int w = 0; w < 6; ++w)
        for(size_t b = 0; b < s->walls[w].size() / 3; ++b)
            bnd[static_cast<size_t>(w) * s->bndStride + b] = make_float4(s->walls[w][3 * b], s->walls[w][3 * b + 1], s->walls[w][3 * b + 2], 0.f);
    SF_CUDA(s, dev_alloc(s->B.bnd, bnd.size()));
    SF_CUDA(s, cudaMemcpyAsync(s->B.bnd, bnd.data(), sizeof(float4) * bnd.size(), cudaMemcpyHostToDevice, s->stream));

    // radix-sort plan: ceil(log2(ncells)) key bits in passes of at most 8 bits
    int bits = 1;
    while((1ull << bits) < s->ncells) ++bits;
    s->sortPasses = (bits + 7) / 8;
    for(int i = 0, left = bits; i < s->sortPasses; ++i) {
        s->sortBits[i] = (left + (s->sortPasses - i) - 1) / (s->sortPasses - i);
        left -= s->sortBits[i];
    }
    fill_dev_params(s);
    // device state: step 0, both max-velocity slots at FLT_MIN, then computeMaxVel of the upload
    DevState init{};
    init.maxv2Bits[0] = init.maxv2Bits[1] = 0x00800000u; // FLT_MIN
    SF_CUDA(s, cudaMemcpyAsync(s->B.state, &init, sizeof(init), cudaMemcpyHostToDevice, s->stream));
    if(s->n) {
        LaunchScope ls(s, K_MARSHAL);
        k_init_maxvel<<<std::min<uint32_t>(cdiv(s->n, 256), s->numSMs * 8), 256, 0, s->stream>>>(s->B.velA, s->n, s->B.state);
    }
    SF_CUDA(s, cudaStreamSynchronize(s->stream));
    SF_CUDA(s, cudaGetLastError());
    s->ready = true;
    return SF_OK;
}

int sf_advance_frame(sf_solver* s, float* dt_out)
{
    int rc = require_ready(s);
    if(rc) return rc;
    SF_CUDA(s, cudaSetDevice(s->device));
    rc = enqueue_substep(s);
    if(rc) return rc;
    rc = read_state(s);
    if(rc) return rc;
    if(dt_out) *dt_out = s->hostState->dt;
    return SF_OK;
}

int sf_advance_steps(sf_solver* s, uint32_t nsteps, float* time_out)
{
    int rc = require_ready(s);
    if(rc) return rc;
    SF_CUDA(s, cudaSetDevice(s->device));
    float t0 = 0.f;
    if(time_out) {
        rc = read_state(s);
        if(rc) return rc;
        t0 = s->hostState->frameTime;
    }
    for(uint32_t i = 0; i < nsteps; ++i) {
        rc = enqueue_substep(s);
        if(rc) return rc;
    }
    if(time_out) {
        rc = read_state(s);
        if(rc) return rc;
        *time_out = s->hostState->frameTime - t0;
    }
    return SF_OK;
}

int sf_advance_frame_time(sf_solver* s, double frame_time, float* time_out, uint32_t* nsteps_out)
{
    int rc = require_ready(s);
    if(rc) return rc;
    if(!(frame_time > 0.0)) return fail(s, SF_ERR_INVALID, "frame_time must be positive");
    SF_CUDA(s, cudaSetDevice(s->device));
    rc = read_state(s);
    if(rc) return rc;
    const unsigned long long steps0 = s->hostState->stepsDone;
    // frameTime = 0; while(frameTime < frame_time) frameTime += advanceFrame();  (Simulator.cpp:46-51)
    DevState patch = *s->hostState;
    patch.frameTime   = 0.f;
    patch.frameTarget = frame_time;
    patch.skip        = 0;
    SF_CUDA(s, cudaMemcpyAsync(s->B.state, &patch, sizeof(patch), cudaMemcpyHostToDevice, s->stream));
    const double dtMax = static_cast<double>(s->P.dtMax);
    for(int guard = 0; guard < 100000; ++guard) {
        rc = read_state(s);
        if(rc) return rc;
        const double remaining = frame_time - static_cast<double>(s->hostState->frameTime);
        if(!(remaining > 0.0)) break;
        // dt <= dtMax, so at least this many more substeps are needed; each one re-checks the target on the device
        const uint32_t batch = static_cast<uint32_t>(std::max(1.0, std::ceil(remaining / dtMax)));
        for(uint32_t i = 0; i < batch; ++i) {
            rc = enqueue_substep(s);
            if(rc) return rc;
        }
    }
    rc = read_state(s);
    if(rc) return rc;
    if(time_out) *time_out = s->hostState->frameTime;
    if(nsteps_out) *nsteps_out = static_cast<uint32_t>(s->hostState->stepsDone - steps0);
    patch             = *s->hostState;
    patch.frameTarget = 0.0;
    patch.skip        = 0;
    SF_CUDA(s, cudaMemcpyAsync(s->B.state, &patch, sizeof(patch), cudaMemcpyHostToDevice, s->stream));
    SF_CUDA(s, cudaStreamSynchronize(s->stream));
    return SF_OK;
}

int sf_synchronize(sf_solver* s)
{
    if(!s) return SF_ERR_INVALID;
    SF_CUDA(s, cudaSetDevice(s->device));
    SF_CUDA(s, cudaStreamSynchronize(s->stream));
    return SF_OK;
}

int sf_step_host(sf_solver* s, float* pos_xyz, float* vel_xyz, uint32_t n, float* dt_out)
{
    if(!s || !pos_xyz || !vel_xyz) return SF_ERR_INVALID;
    const bool sameShape = s->uploaded && s->ready && n == s->n;
    int        rc;
    if(!sameShape) {
        rc = sf_upload_particles(s, pos_xyz, vel_xyz, n);
        if(rc) return rc;
        rc = sf_make_ready(s);
        if(rc) return rc;
    } else {
        // steady state: same particle count, overwrite the device state from the host buffers.
        // The state is re-packed in upload order (id = index), so the previous sort order is dropped.
        SF_CUDA(s, cudaSetDevice(s->device));
        float* dpos = s->stage;
        float* dvel = s->stage + 3 * static_cast<size_t>(s->npad);
        SF_CUDA(s, cudaMemcpyAsync(dpos, pos_xyz, static_cast<size_t>(n) * 12, cudaMemcpyHostToDevice, s->stream));
        SF_CUDA(s, cudaMemcpyAsync(dvel, vel_xyz, static_cast<size_t>(n) * 12, cudaMemcpyHostToDevice, s->stream));
        {
            LaunchScope ls(s, K_MARSHAL);
            k_pack_upload<<<cdiv(n, 256), 256, 0, s->stream>>>(dpos, dvel, s->B.posA, s->B.velA, s->B.idA, n);
        }
        DevState init{};
        init.maxv2Bits[0] = init.maxv2Bits[1] = 0x00800000u;
        SF_CUDA(s, cudaMemcpyAsync(s->B.state, &init, sizeof(init), cudaMemcpyHostToDevice, s->stream));
        {
            LaunchScope ls(s, K_MARSHAL);
            k_init_maxvel<<<std::min<uint32_t>(cdiv(n, 256), s->numSMs * 8), 256, 0, s->stream>>>(s->B.velA, n, s->B.state);
        }
    }
    rc = enqueue_substep(s);
    if(rc) return rc;
    {
        LaunchScope ls(s, K_MARSHAL);
        float*      dpos = s->stage;
        float*      dvel = s->stage + 3 * static_cast<size_t>(s->npad);
        k_unpack_xyz<<<cdiv(n, 256), 256, 0, s->stream>>>(s->B.posA, s->B.idA, dpos, n);
        k_unpack_xyz<<<cdiv(n, 256), 256, 0, s->stream>>>(s->B.velA, s->B.idA, dvel, n);
        SF_CUDA(s, cudaMemcpyAsync(pos_xyz, dpos, static_cast<size_t>(n) * 12, cudaMemcpyDeviceToHost, s->stream));
        SF_CUDA(s, cudaMemcpyAsync(vel_xyz, dvel, static_cast<size_t>(n) * 12, cudaMemcpyDeviceToHost, s->stream));
    }
    rc = read_state(s);
    if(rc) return rc;
    if(dt_out) *dt_out = s->hostState->dt;
    return SF_OK;
}

int sf_set_capture(sf_solver* s, int on)
{
    if(!s) return SF_ERR_INVALID;
    s->capture   = on != 0;
    s->P.capture = on ? 1 : 0;
    return SF_OK;
}

int sf_grid_dims(sf_solver* s, int32_t n3[3])
{
    if(!s || !n3) return SF_ERR_INVALID;
    grid_dims(s->params, n3);
    return SF_OK;
}

static int neighbor_lists_host(sf_solver* s, std::vector<uint32_t>& counts, std::vector<uint32_t>* ids)
{
    // neighbour sets of the last substep's binning, by traversal of the device cell tables (CSR in two passes)
    const uint32_t n = s->n;
    counts.assign(n, 0);
    if(n == 0) {
        if(ids) ids->clear();
        return SF_OK;
    }
    if(!s->B.keyB) return fail(s, SF_ERR_INVALID, "no substep has run yet");
    uint32_t* dCounts = nullptr;
    SF_CUDA(s, cudaMalloc(reinterpret_cast<void**>(&dCounts), sizeof(uint32_t) * n));
    k_neighbor_count<<<cdiv(n, 128), 128, 0, s->stream>>>(s->B, s->P, dCounts);
    cudaError_t e = cudaMemcpyAsync(counts.data(), dCounts, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s->stream);
    if(e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(dCounts);
    SF_CUDA(s, e);
    if(!ids) return SF_OK;
    std::vector<unsigned long long> offset(n + 1, 0);
    for(uint32_t i = 0; i < n; ++i) offset[i + 1] = offset[i] + counts[i];
    ids->assign(offset[n], 0);
    if(offset[n] == 0) return SF_OK;
    unsigned long long* dOff = nullptr;
    uint32_t*           dIds = nullptr;
    SF_CUDA(s, cudaMalloc(reinterpret_cast<void**>(&dOff), sizeof(unsigned long long) * (n + 1)));
    e = cudaMalloc(reinterpret_cast<void**>(&dIds), sizeof(uint32_t) * offset[n]);
    if(e == cudaSuccess) e = cudaMemcpyAsync(dOff, offset.data(), sizeof(unsigned long long) * (n + 1), cudaMemcpyHostToDevice, s->stream);
    if(e == cudaSuccess) {
        k_neighbor_fill<<<cdiv(n, 128), 128, 0, s->stream>>>(s->B, s->P, dOff, dIds);
        e = cudaMemcpyAsync(ids->data(), dIds, sizeof(uint32_t) * offset[n], cudaMemcpyDeviceToHost, s->stream);
    }
    if(e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(dOff);
    cudaFree(dIds);
    SF_CUDA(s, e);
    for(uint32_t i = 0; i < n; ++i) std::sort(ids->begin() + offset[i], ids->begin() + offset[i + 1]);
    return SF_OK;
}

int sf_field_size(sf_solver* s, int field, uint64_t* bytes_out)
{
    if(!s || !bytes_out) return SF_ERR_INVALID;
    const uint64_t n = s->n;
    switch(field) {
        case SF_FIELD_DENSITY:
        case SF_FIELD_PRESSURE:
        case SF_FIELD_CELL_INDEX:
        case SF_FIELD_NEIGHBOR_COUNT:
        case SF_FIELD_SORT_PERM: *bytes_out = 4 * n; return SF_OK;
        case SF_FIELD_ACCEL: *bytes_out = 12 * n; return SF_OK;
        case SF_FIELD_TABLE_CUBIC_W:
        case SF_FIELD_TABLE_SPIKY_GRAD: *bytes_out = 4ull * kTableEntries; return SF_OK;
        case SF_FIELD_NEIGHBOR_IDS: {
            int rc = require_ready(s);
            if(rc) return rc;
            SF_CUDA(s, cudaSetDevice(s->device));
            SF_CUDA(s, cudaStreamSynchronize(s->stream));
            std::vector<uint32_t> counts;
            rc = neighbor_lists_host(s, counts, nullptr);
            if(rc) return rc;
            uint64_t total = 0;
            for(uint32_t c : counts) total += c;
            *bytes_out = 4 * total;
            return SF_OK;
        }
        default: return fail(s, SF_ERR_INVALID, "unknown field");
    }
}

int sf_download_field(sf_solver* s, int field, void* out, uint64_t bytes)
{
    int rc = require_ready(s);
    if(rc) return rc;
    if(!out) return SF_ERR_INVALID;
    SF_CUDA(s, cudaSetDevice(s->device));
    SF_CUDA(s, cudaStreamSynchronize(s->stream));
    const uint32_t n = s->n;
    if(field == SF_FIELD_TABLE_CUBIC_W || field == SF_FIELD_TABLE_SPIKY_GRAD) {
        if(bytes < 4ull * kTableEntries) return fail(s, SF_ERR_INVALID, "buffer too small");
        SF_CUDA(s, cudaMemcpy(out, field == SF_FIELD_TABLE_CUBIC_W ? s->B.tabW : s->B.tabG, 4ull * kTableEntries, cudaMemcpyDeviceToHost));
        return SF_OK;
    }
    if(field == SF_FIELD_NEIGHBOR_COUNT || field == SF_FIELD_NEIGHBOR_IDS) {
        std::vector<uint32_t> counts, ids;
        rc = neighbor_lists_host(s, counts, field == SF_FIELD_NEIGHBOR_IDS ? &ids : nullptr);
        if(rc) return rc;
        const std::vector<uint32_t>& src = field == SF_FIELD_NEIGHBOR_IDS ? ids : counts;
        if(bytes < 4ull * src.size()) return fail(s, SF_ERR_INVALID, "buffer too small");
        std::memcpy(out, src.data(), 4ull * src.size());
        return SF_OK;
    }
    const uint64_t need = (field == SF_FIELD_ACCEL ? 12ull : 4ull) * n;
    if(bytes < need) return fail(s, SF_ERR_INVALID, "buffer too small");
    if(n == 0) return SF_OK;
    const uint32_t g = cdiv(n, 256);
    switch(field) {
        case SF_FIELD_DENSITY: k_unpack_scalar<<<g, 256, 0, s->stream>>>(s->B.rho, s->B.idA, s->stage, n); break;
        case SF_FIELD_PRESSURE: k_unpack_pressure<<<g, 256, 0, s->stream>>>(s->B.rho, s->B.idA, s->stage, s->P); break;
        case SF_FIELD_CELL_INDEX:
            if(!s->B.keyB) return fail(s, SF_ERR_INVALID, "no substep has run yet");
            k_unpack_u32<<<g, 256, 0, s->stream>>>(s->B.keyB, s->B.idA, reinterpret_cast<uint32_t*>(s->stage), n);
            break;
        case SF_FIELD_SORT_PERM:
            SF_CUDA(s, cudaMemcpy(out, s->B.idA, need, cudaMemcpyDeviceToHost));
            return SF_OK;
        case SF_FIELD_ACCEL:
            if(!s->capture) return fail(s, SF_ERR_INVALID, "SF_FIELD_ACCEL needs sf_set_capture(1) before the substep");
            k_unpack_xyz<<<g, 256, 0, s->stream>>>(s->B.accel, s->B.idA, s->stage, n);
            break;
        default: return fail(s, SF_ERR_INVALID, "unknown field");
    }
    SF_CUDA(s, cudaMemcpyAsync(out, s->stage, need, cudaMemcpyDeviceToHost, s->stream));
    SF_CUDA(s, cudaStreamSynchronize(s->stream));
    return SF_OK;
}

// ---- measurement -------------------------------------------------------------------------------
int sf_profile_enable(sf_solver* s, int on)
{
    if(!s) return SF_ERR_INVALID;
    cudaSetDevice(s->device);
    if(!on) fold_pending(s);
    s->profiling = on != 0;
    return SF_OK;
}

int sf_profile_reset(sf_solver* s)
{
    if(!s) return SF_ERR_INVALID;
    cudaSetDevice(s->device);
    fold_pending(s);
    std::memset(s->profMs, 0, sizeof(s->profMs));
    std::memset(s->profLaunches, 0, sizeof(s->profLaunches));
    return SF_OK;
}

int sf_profile_get(sf_solver* s, char* names_buf, size_t names_cap, double* ms, uint64_t* launches, uint32_t cap, uint32_t* count_out)
{
    if(!s) return SF_ERR_INVALID;
    cudaSetDevice(s->device);
    fold_pending(s);
    size_t off = 0;
    for(uint32_t i = 0; i < K_COUNT; ++i) {
        if(names_buf) {
            const size_t len = std::strlen(kKernelNames[i]) + 1;
            if(off + len <= names_cap) {
                std::memcpy(names_buf + off, kKernelNames[i], len);
                off += len;
            }
        }
        if(i < cap) {
            if(ms) ms[i] = s->profMs[i];
            if(launches) launches[i] = s->profLaunches[i];
        }
    }
    if(count_out) *count_out = K_COUNT;
    return SF_OK;
}

int sf_launch_count(sf_solver* s, uint64_t* n_out)
{
    if(!s || !n_out) return SF_ERR_INVALID;
    *n_out = s->launches;
    return SF_OK;
}

int sf_timer_start(sf_solver* s)
{
    if(!s) return SF_ERR_INVALID;
    SF_CUDA(s, cudaSetDevice(s->device));
    SF_CUDA(s, cudaEventRecord(s->timerA, s->stream));
    return SF_OK;
}

int sf_timer_stop(sf_solver* s, float* ms_out)
{
    if(!s || !ms_out) return SF_ERR_INVALID;
    SF_CUDA(s, cudaSetDevice(s->device));
    SF_CUDA(s, cudaEventRecord(s->timerB, s->stream));
    SF_CUDA(s, cudaEventSynchronize(s->timerB));
    SF_CUDA(s, cudaEventElapsedTime(ms_out, s->timerA, s->timerB));
    return SF_OK;
}

// ---- multi-GPU (slab decomposition) -- implemented in sf_slab.cu when built --------------------
#ifndef SF_WITH_SLAB
int sf_comm_unique_id(void*) { return SF_ERR_INVALID; }
int sf_comm_init(sf_solver* s, int, int, const void*) { return fail(s, SF_ERR_INVALID, "slab decomposition not built"); }
int sf_upload_particles_global(sf_solver* s, const float*, const float*, uint32_t) { return fail(s, SF_ERR_INVALID, "slab decomposition not built"); }
int sf_slab_info(sf_solver* s, int32_t*, int32_t*, uint32_t*, uint32_t*) { return fail(s, SF_ERR_INVALID, "slab decomposition not built"); }
int sf_download_owned(sf_solver* s, uint32_t*, float*, float*, uint32_t, uint32_t*) { return fail(s, SF_ERR_INVALID, "slab decomposition not built"); }
#endif

} // extern "C"
