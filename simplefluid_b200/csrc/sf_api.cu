// sf_api.cu -- C-ABI (include/sf_b200.h) over the sm_100a SPH pipeline.
//
// One sf_solver = one GPU = the object behind QtSPHSolver (Include/QtSPHSolver.h:27-36).
// The substep (advanceFrame, EXE@0x140016810) is a fixed sequence of kernel launches on one stream
// with no host synchronisation: dt, the frame-time accumulator and the max-velocity reduction all
// live in a DevState block on the device.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>
#include "sf_internal.h"
#include "sf_pairs.cuh"
#include "sf_slab.cuh"

using namespace sf;

namespace
{
thread_local std::string g_createError;

enum KernelId {
    K_BEGIN = 0, K_HASH, K_RADIX_HIST, K_RADIX_SCAN, K_RADIX_SCATTER, K_CLEAR_CELLS, K_CELL_BOUNDS, K_BRICK_COMPACT, K_REORDER,
    K_DENSITY, K_CORRECT_DENSITY, K_FORCE, K_VISC_INTEGRATE, K_MARSHAL, K_HASH_COUNT, K_CELL_SCAN, K_COUNT_SCATTER, K_COUNT
};
const char* const kKernelNames[K_COUNT] = {
    "k_begin_step", "k_hash", "k_radix_hist", "k_radix_scan", "k_radix_scatter", "k_clear_cells", "k_cell_bounds", "k_brick_compact",
    "k_reorder", "k_density", "k_correct_density", "k_force", "k_visc_integrate", "k_marshal", "k_hash_count", "k_cell_scan", "k_count_scatter"
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
    int        kmax = 64; // list rows per particle: 47-48 is the most any developed BASELINE state needs (profiles/r02_d_config_sweep.log)
    int32_t    grid[3] = { 0, 0, 0 };
    uint64_t   ncells = 0, cellCap = 0;
    bool       ready = false, uploaded = false, capture = false;
    DevBuffers B{};
    DevParams  P{};
    KernelTables tables;
    std::vector<float> walls[6];
    bool         wallsSet = false;
    uint32_t     bndStride = 0;
    cudaStream_t xferStream = nullptr; // sf_step_host: host<->device copies beside the compute stream
    cudaEvent_t  evPosUp = nullptr, evVelUp = nullptr, evPosOut = nullptr, evVelOut = nullptr;
    cudaStream_t snapStream = nullptr; // asynchronous position snapshots for the viewer
    cudaEvent_t  snapReady = nullptr, snapDone = nullptr;
    float*       snapBuf = nullptr;
    uint32_t     snapCap = 0;
    float*       stage = nullptr; // device staging for host<->device marshalling (2 x 12 B x cap)
    size_t       stageBytes = 0;
    DevState*    hostState = nullptr; // pinned
    uint32_t     radixBlocks = 0;
    // Kernel variants, all bit-identical (tools/variant_bench.py); the environment overrides exist for A/B timing:
    bool         countSort = true;       // SF_SORT=count (default): counting sort by cell; radix: three LSD radix passes
    uint32_t*    cellTileSums = nullptr; // counting sort: per-tile particle counts of the cell table
    uint64_t     cellTileCap = 0;
    int          sortPasses = 0, sortBits[4] = { 0, 0, 0, 0 };
    int          occDensity = 1, occForce = 1, occVisc = 1;
    uint32_t     numBricks = 0, brickCap = 0;
    uint32_t     nSlots = 0; // live + dead slots of the A arrays (== n on a single GPU)
    // captured substep (single GPU); dropped whenever the launch arguments change.  [1]: the steady form, whose cell
    // counts were left by the previous substep's integrate kernel; [0]: with k_hash_count up front
    cudaGraphExec_t stepGraph[2] = { nullptr, nullptr };
    bool            hashed = false;      // cellCnt / keys[0] / vals[0] hold the binning of the current posA (written by k_visc_brick<true>)
    uint64_t        graphLaunches[2] = { 0, 0 };
    bool            useGraph = true;
    int          axisS = 2;  // slow axis of the cell key: 2 = z (reference order), 1 = y (slab runs that are longer in y)
    int32_t      nS() const { return grid[axisS]; }
    int32_t      nM() const { return grid[3 - axisS]; }

    // z-slab decomposition (sf_comm_init)
    struct Slab {
        bool                 on = false;
        int                  rank = 0, nranks = 1;
        ncclComm_t           comm = nullptr;
        std::vector<int32_t> cur, next; // cut planes (nranks + 1): bounds of this substep / of the next one
        cudaStream_t         commStream = nullptr;
        cudaEvent_t          evEdge = nullptr, evExchanged = nullptr;
        float4 *             sendLo = nullptr, *sendHi = nullptr, *recvLo = nullptr, *recvHi = nullptr;
        uint32_t             sendCap = 0, recvCap = 0;
        uint64_t             regrowths = 0;
        uint32_t *           layerStart = nullptr, *counters = nullptr, *row = nullptr, *table = nullptr;
        uint32_t*            hostTable = nullptr; // pinned
        uint64_t             nGlobal = 0;
        uint32_t             nOwn = 0;
        uint64_t             ncellsMax = 0;
        uint64_t             exchangedParticles = 0;
        double               tWaitPack = 0, tPost = 0, tFront = 0; // host seconds: waiting for the table / posting the exchange / enqueueing
        double               tGroup = 0, tReduce = 0; // ... of which: the send/recv group, the dt all-reduce
        // SF_SLAB_TRACE=2: device-side timeline of the last kTlSteps substeps (CUDA events; printed at destruction):
        // 0 step start, 1 sort + reorder done, 2 density done, 3 dt known, 4 force done on the edge layers, 5 edge layers
        // integrated, 6 interior layers done (compute stream); 7 table gathered, 8 exchange done (communication stream)
        static constexpr int kTlSteps = 64, kTlMarks = 9;
        std::vector<cudaEvent_t> tl;
        int                      tlLevel = 0;
        uint64_t             steps = 0;
        // global dt: the all-reduce of max |v|^2 runs on the communication stream behind the exchange and is awaited
        // only before the force pass of the next substep (dtReduced: evDt covers the current maxv2Bits)
        cudaEvent_t          evInterior = nullptr, evDt = nullptr;
        bool                 dtReduced = false;
        int                  ghost = 3; // ghost layers per side this rank holds (fixed at sf_upload_particles_global)
    } slab;
    uint32_t *ownCounters = nullptr, *hostOwnCounters = nullptr; // device / pinned: sf_download_owned, sf_step_host_owned

    // measurement
    bool                      profiling = false; // the substep being enqueued is timed kernel by kernel
    uint32_t                  profEvery = 0, profCount = 0; // sf_profile_enable(N): every N-th substep is timed (0 = off)
    double                    profMs[K_COUNT] = {};
    uint64_t                  profLaunches[K_COUNT] = {};
    std::vector<PendingEvent> pending;
    std::vector<cudaEvent_t>  eventPool;
    uint64_t                  launches = 0;
    cudaEvent_t               timerA = nullptr, timerB = nullptr;
};

namespace
{
void drop_graph(sf_solver* s)
{
    for(cudaGraphExec_t& g : s->stepGraph) {
        if(g) cudaGraphExecDestroy(g);
        g = nullptr;
    }
}

// The list walkers request rows four at a time (walk_list, sf_pairs.cuh): the capacity is a multiple of four, so that
// "the next four rows lie inside the column" is one comparison
int round_list_capacity(int kmax) { return std::min((kmax + 3) & ~3, 16380); }

// posA is about to change behind the integrate kernel's back (or the grid is): forget the fused binning
int invalidate_binning(sf_solver* s)
{
    cudaError_t e = cudaSuccess;
    if(s->hashed && s->B.cellCnt) e = cudaMemsetAsync(s->B.cellCnt, 0, sizeof(uint32_t) * s->ncells, s->stream);
    s->hashed = false;
    return e == cudaSuccess ? SF_OK : SF_ERR_CUDA;
}

void tl_mark(sf_solver* s, int mark, cudaStream_t st)
{
    sf_solver::Slab& L = s->slab;
    if(L.tlLevel < 2) return;
    if(L.tl.empty()) {
        L.tl.resize(static_cast<size_t>(L.kTlSteps) * L.kTlMarks);
        for(auto& e : L.tl) cudaEventCreate(&e);
    }
    cudaEventRecord(L.tl[static_cast<size_t>(L.steps % L.kTlSteps) * L.kTlMarks + mark], st);
}

void slab_timeline_report(sf_solver* s)
{
    sf_solver::Slab& L = s->slab;
    if(!L.tl.empty() && L.steps > static_cast<uint64_t>(L.kTlSteps)) {
        // mean intervals over the recorded ring (skipping the step the ring currently points at)
        const char* names[] = { "sort+reorder", "density", "wait dt (all-reduce)", "force (edge layers)", "integrate (edge layers)", "force + integrate (interior layers)",
                                "edge->table gathered (comm)", "table->exchange done (comm, incl. host turn-around)",
                                "interior done->next step start (compute stream idle)", "exchange done->next step start" };
        double acc[10] = {};
        int    cnt = 0;
        auto ms = [&](int stepA, int a, int stepB, int b) {
            float t = 0.f;
            cudaEventElapsedTime(&t, L.tl[static_cast<size_t>(stepA) * L.kTlMarks + a], L.tl[static_cast<size_t>(stepB) * L.kTlMarks + b]);
            return static_cast<double>(t);
        };
        const int cur = static_cast<int>(L.steps % L.kTlSteps);
        for(int k = 0; k < L.kTlSteps - 2; ++k) {
            const int a = (cur + 1 + k) % L.kTlSteps, b = (a + 1) % L.kTlSteps; // a older than b, both complete
            for(int i = 0; i < 6; ++i) acc[i] += ms(a, i, a, i + 1);
            acc[6] += ms(a, 5, a, 7);
            acc[7] += ms(a, 7, a, 8);
            acc[8] += ms(a, 6, b, 0);
            acc[9] += ms(a, 8, b, 0);
            ++cnt;
        }
        std::string line = "[sf slab rank " + std::to_string(L.rank) + "] device timeline, mean ms over " + std::to_string(cnt) + " substeps:";
        for(int i = 0; i < 10; ++i) {
            char buf[128];
            std::snprintf(buf, sizeof(buf), " %s %.3f;", names[i], acc[i] / cnt);
            line += buf;
        }
        std::fprintf(stderr, "%s\n", line.c_str());
    }
}

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
    P.ny = s->nM();
    P.nz = s->nS();
    P.axisS = s->axisS;
    P.useBoundary    = p.bUseBoundaryParticles ? 1 : 0;
    P.attractive     = p.bUseAttractivePressure ? 1 : 0;
    P.correctDensity = p.bCorrectDensity ? 1 : 0;
    P.capture        = s->capture ? 1 : 0;
    P.n    = s->n;
    P.npad = s->npad;
    P.kmax = s->kmax;
    P.kmaxBytes = static_cast<uint32_t>(s->kmax) * 128u;
    P.nbx  = (s->grid[0] + BX - 1) / BX;
    P.nby  = (s->nM() + BY - 1) / BY;
    P.nbz  = (s->nS() + BZ - 1) / BZ;
    P.numBricks = s->numBricks;
    P.z0 = 0;
    P.nzGlobal = s->nS();
    P.zDensLo = P.zShepLo = P.zForceLo = P.zOwnLo = 0;
    P.zDensHi = P.zShepHi = P.zForceHi = P.zOwnHi = s->nS();
    P.zEdge = 0;
    P.slab  = 0;
    for(int w = 0; w < 6; ++w) P.nbnd[w] = P.useBoundary ? static_cast<uint32_t>(s->walls[w].size() / 3) : 0u;
    P.bndStride = s->bndStride;
    P.wallWords  = (s->bndStride + 31u) / 32u;
    P.wallSubInv = wall_sub_inv(p);
}

// preserve > 0: the first `preserve` slots of the state arrays (posA / velA / idA) survive the reallocation; every
// other per-particle array is scratch that the next substep rewrites.  The caller has synchronised the streams.
template<class T>
cudaError_t dev_grow(T*& p, size_t newCount, size_t preserve)
{
    T*          q = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&q), newCount * sizeof(T));
    if(e != cudaSuccess) return e;
    if(p && preserve) e = cudaMemcpy(q, p, preserve * sizeof(T), cudaMemcpyDeviceToDevice);
    if(p) cudaFree(p);
    p = q;
    return e;
}

int ensure_particle_capacity(sf_solver* s, uint32_t n, uint32_t preserve = 0)
{
    if(n <= s->cap && s->B.posA) return SF_OK;
    const uint32_t cap  = n;
    const uint32_t npad = (cap + 127u) & ~127u;
    if(preserve) {
        SF_CUDA(s, dev_grow(s->B.posA, npad, preserve));
        SF_CUDA(s, dev_grow(s->B.velA, npad, preserve));
        SF_CUDA(s, dev_grow(s->B.idA, npad, preserve));
    } else {
        SF_CUDA(s, dev_alloc(s->B.posA, npad));
        SF_CUDA(s, dev_alloc(s->B.velA, npad));
        SF_CUDA(s, dev_alloc(s->B.idA, npad));
    }
    SF_CUDA(s, dev_alloc(s->B.posB, npad));
    SF_CUDA(s, dev_alloc(s->B.velB, npad));
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

int slab_exchange(sf_solver* s, uint32_t n, uint32_t nSlots);
int slab_global_dt(sf_solver* s);

// One reference substep (advanceFrame, EXE@0x140016810) as a launch sequence on the solver's stream.
// n = live particles, nSlots >= n = slots of the A arrays to sort (slab mode keeps last step's dead ghost slots).
// velHostXYZ != nullptr (sf_step_host): the velocities of this substep are still on their way from the host; they
// are gathered from that xyz array (after evVelUp) between the density and the force pass, dt follows them.
int enqueue_substep_launches(sf_solver* s, const float* velHostXYZ = nullptr)
{
    DevBuffers&     B  = s->B;
    const DevParams P  = s->P;
    cudaStream_t    st = s->stream;
    const bool      slab   = s->slab.on;
    const uint32_t  n      = s->n;
    const uint32_t  nSlots = slab ? s->nSlots : n;
    const uint32_t  gridN  = cdiv(n, 256);
    // Slab mode: dt needs max |v|^2 over ALL slabs (uint32 order == float order for non-negative floats).  The
    // all-reduce was issued on the communication stream behind the previous substep's exchange (slab_exchange) and is
    // awaited only before the force pass: sorting and the density pass need no dt, so the skew between ranks -- every
    // rank would otherwise wait here for the slowest one to finish integrating -- hides behind them.
    const bool splitClock = velHostXYZ != nullptr || slab;
    if(slab) tl_mark(s, 0, st);
    {
        LaunchScope ls(s, K_BEGIN);
        k_begin_step<<<1, 1, 0, st>>>(B.state, P, splitClock ? kBeginResets : kBeginAll);
    }
    if(nSlots == 0 && !slab) {
        if(velHostXYZ) k_begin_step<<<1, 1, 0, st>>>(B.state, P, kBeginClock);
        SF_CUDA(s, cudaGetLastError());
        return SF_OK;
    }
    int cur = 0;
    // single GPU, device-resident state: the integrate kernel also bins the new positions for the next substep
    const bool fuseHash = s->countSort && !slab && !velHostXYZ;
    if(s->countSort) {
        // counting sort by cell (sf_kernels.cuh, section 1'): cell table first, then the slot permutation
        const uint32_t nc = static_cast<uint32_t>(s->ncells), ntiles = cdiv(nc, CS_TILE);
        if(nSlots && !(fuseHash && s->hashed)) {
            LaunchScope ls(s, K_HASH_COUNT);
            k_hash_count<<<cdiv(nSlots, 256), 256, 0, st>>>(B.posA, B.idA, B.keys[0], B.vals[0], nSlots, P, B.cellCnt, B.state);
        }
        {
            LaunchScope ls(s, K_CELL_SCAN);
            s->launches += 2; // three kernels under one timing scope
            k_cell_scan_reduce<<<ntiles, CS_THREADS, 0, st>>>(B.cellCnt, nc, s->cellTileSums, B.state);
            k_radix_scan<<<1, 1024, 0, st>>>(s->cellTileSums, ntiles, B.radixTotals, B.state);
            k_cell_scan_apply<<<ntiles, CS_THREADS, 0, st>>>(B.cellTab, B.cellCnt, nc, s->cellTileSums, B.brickFlag, P, B.state);
        }
        if(nSlots) {
            LaunchScope ls(s, K_COUNT_SCATTER);
            k_count_scatter<<<cdiv(nSlots, 256), 256, 0, st>>>(B.keys[0], B.vals[0], B.cellTab, B.keys[1], B.vals[1], nSlots, B.state);
        }
        cur    = 1;
        B.keyB = B.keys[1];
    } else {
        if(nSlots) {
            {
                LaunchScope ls(s, K_HASH);
                k_hash<<<cdiv(nSlots, 256), 256, 0, st>>>(B.posA, B.idA, B.keys[0], B.vals[0], nSlots, P, B.state);
            }
            const uint32_t nb    = cdiv(nSlots, RS_TILE);
            int            shift = 0;
            for(int pass = 0; pass < s->sortPasses; ++pass) {
                const int radix = 1 << s->sortBits[pass];
                {
                    LaunchScope ls(s, K_RADIX_HIST);
                    k_radix_hist<<<nb, RS_THREADS, 0, st>>>(B.keys[cur], nSlots, shift, radix, B.radixCounts, nb, B.state);
                }
                {
                    LaunchScope ls(s, K_RADIX_SCAN);
                    k_radix_scan<<<radix, 1024, 0, st>>>(B.radixCounts, nb, B.radixTotals, B.state);
                }
                {
                    LaunchScope ls(s, K_RADIX_SCATTER);
                    k_radix_scatter<<<nb, RS_THREADS, 0, st>>>(B.keys[cur], B.vals[cur], B.keys[cur ^ 1], B.vals[cur ^ 1], nSlots, shift, radix,
                                                                B.radixCounts, nb, B.radixTotals, B.state);
                }
                shift += s->sortBits[pass];
                cur ^= 1;
            }
        }
        B.keyB = B.keys[cur];
        {
            LaunchScope  ls(s, K_CLEAR_CELLS);
            const size_t nvec = (s->ncells * sizeof(uint2) + 15) / 16;
            k_clear_cells<<<std::min<uint32_t>(cdiv(nvec, 256), s->numSMs * 16), 256, 0, st>>>(reinterpret_cast<uint4*>(B.cellTab), nvec, B.state);
        }
        if(n) {
            LaunchScope ls(s, K_CELL_BOUNDS);
            k_cell_bounds_bricks<<<gridN, 256, 0, st>>>(B.keyB, n, B.cellTab, B.brickFlag, P, B.state);
        }
    }
    {
        LaunchScope ls(s, K_BRICK_COMPACT);
        k_brick_compact<<<cdiv(s->numBricks, 1024), 1024, 0, st>>>(B.brickFlag, B.brickList, s->numBricks, B.state);
    }
    if(slab) k_layer_start<<<cdiv(static_cast<uint32_t>(P.nz) + 1u, 128), 128, 0, st>>>(B.keyB, n, P, s->slab.layerStart);
    const uint32_t pairGrid = std::max<uint32_t>(1u, std::min<uint32_t>(s->numBricks, static_cast<uint32_t>(s->numSMs) * 2u));
    if(n) {
        {
            LaunchScope ls(s, K_REORDER);
            if(velHostXYZ) k_reorder_pos<<<gridN, 256, 0, st>>>(B.keyB, B.vals[cur], B.cellTab, B.posA, B.idA, B.posB, B.idB, n, B.state);
            else k_reorder<<<gridN, 256, 0, st>>>(B.keyB, B.vals[cur], B.cellTab, B.posA, B.velA, B.idA, B.posB, B.velB, B.idB, n, B.state);
        }
        if(slab) tl_mark(s, 1, st);
        {
            LaunchScope ls(s, K_DENSITY);
            k_density_brick<<<std::min<uint32_t>(pairGrid, s->numSMs * s->occDensity), kBrickThreads, kSmemDensity, st>>>(B, P);
        }
        if(slab) tl_mark(s, 2, st);
        if(P.correctDensity) {
            LaunchScope ls(s, K_CORRECT_DENSITY);
            k_shepard_brick<<<std::min<uint32_t>(pairGrid, s->numSMs * s->occForce), kBrickThreads, kSmemPair, st>>>(B, P);
            k_density_terms<<<gridN, 256, 0, st>>>(B, P);
        }
        if(velHostXYZ) {
            SF_CUDA(s, cudaStreamWaitEvent(st, s->evVelUp, 0));
            LaunchScope ls(s, K_MARSHAL);
            k_gather_vel_host<<<gridN, 256, 0, st>>>(velHostXYZ, B.idB, B.velB, n, B.state);
            k_begin_step<<<1, 1, 0, st>>>(B.state, P, kBeginClock);
        }
    }
    if(slab) {
        const int rc = slab_global_dt(s);
        if(rc) return rc;
        tl_mark(s, 3, st);
    }
    if(n) {
        {   // slab mode: the layers near the slab faces first; the interior layers follow the edge integrate (slab_exchange)
            LaunchScope ls(s, K_FORCE);
            k_force_brick<<<std::min<uint32_t>(pairGrid, s->numSMs * s->occForce), kBrickThreads, kSmemPair, st>>>(B, P, slab ? 1 : 0);
        }
    }
    if(slab) tl_mark(s, 4, st);
    if(!slab) {
        LaunchScope ls(s, K_VISC_INTEGRATE);
        if(fuseHash) k_visc_brick<true><<<std::min<uint32_t>(pairGrid, s->numSMs * s->occVisc), kBrickThreads, kSmemPair, st>>>(B, P, 0);
        else k_visc_brick<false><<<std::min<uint32_t>(pairGrid, s->numSMs * s->occVisc), kBrickThreads, kSmemPair, st>>>(B, P, 0);
        SF_CUDA(s, cudaGetLastError());
        s->hashed = fuseHash && n != 0;
        return SF_OK;
    }
    return slab_exchange(s, n, nSlots);
}

// Single-GPU substeps are replayed from a CUDA graph (the launch sequence and every kernel argument are fixed
// between two makeReady calls; dt and the work cursors live in device memory), which removes the launch gaps that
// dominate small scenes.  Profiling and slab mode use direct launches.
int enqueue_substep(sf_solver* s)
{
    // sampled per-kernel timing: the timed substeps use direct launches with events, the others the graph
    s->profiling = s->profEvery && (s->profCount++ % s->profEvery) == 0;
    if(s->slab.on || s->profiling || !s->useGraph || s->n == 0) return enqueue_substep_launches(s);
    const int gi = s->hashed ? 1 : 0; // the launch sequence depends on whether the previous substep left the binning
    if(!s->stepGraph[gi]) {
        const uint64_t before = s->launches;
        cudaGraph_t    graph  = nullptr;
        SF_CUDA(s, cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_substep_launches(s);
        cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
        if(rc) {
            if(graph) cudaGraphDestroy(graph);
            return rc;
        }
        SF_CUDA(s, e);
        e = cudaGraphInstantiate(&s->stepGraph[gi], graph, 0);
        cudaGraphDestroy(graph);
        SF_CUDA(s, e);
        s->graphLaunches[gi] = s->launches - before;
        s->launches          = before;
    }
    SF_CUDA(s, cudaGraphLaunch(s->stepGraph[gi], s->stream));
    s->launches += s->graphLaunches[gi];
    s->hashed = s->countSort && s->n != 0; // what enqueue_substep_launches recorded while capturing
    return SF_OK;
}

int read_state(sf_solver* s)
{
    SF_CUDA(s, cudaMemcpyAsync(s->hostState, s->B.state, sizeof(DevState), cudaMemcpyDeviceToHost, s->stream));
    SF_CUDA(s, cudaStreamSynchronize(s->stream));
    if(s->hostState->errFlags & SF_DEVERR_DOMAIN) {
        cudaMemsetAsync(&s->B.state->errFlags, 0, sizeof(unsigned), s->stream);
        return fail(s, SF_ERR_DOMAIN, "a particle of the host buffer lies outside the simulation box or is not finite");
    }
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

namespace
{
// P / sort plan / brick grid for the current cut planes: local window = own layers + L.ghost layers on each side
int slab_configure_window(sf_solver* s)
{
    sf_solver::Slab& L = s->slab;
    const int zb = L.cur[L.rank], ze = L.cur[L.rank + 1];
    const int g   = L.ghost;
    const int nzL = ze - zb + 2 * g;
    s->ncells    = static_cast<uint64_t>(s->grid[0]) * s->nM() * nzL;
    s->numBricks = static_cast<uint32_t>((s->grid[0] + BX - 1) / BX) * ((s->nM() + BY - 1) / BY) * ((nzL + BZ - 1) / BZ);
    fill_dev_params(s);
    DevParams& P = s->P;
    P.nz        = nzL;
    P.nbz       = (nzL + BZ - 1) / BZ;
    P.numBricks = s->numBricks;
    P.z0        = zb - g;
    P.nzGlobal  = s->nS();
    // every pass needs its neighbours' previous pass: density on all but the outermost ghost layer it needs, then
    // (Shepard,) force, and the XSPH sum + integration on the own layers
    const int dens = P.correctDensity ? g - 3 : g - 2;
    P.zDensLo  = dens;
    P.zDensHi  = nzL - dens;
    P.zShepLo  = g - 2;
    P.zShepHi  = nzL - (g - 2);
    P.zForceLo = g - 1;
    P.zForceHi = nzL - (g - 1);
    P.zOwnLo = g;
    P.zOwnHi = nzL - g;
    P.zEdge  = slab_edge(g);
    P.slab   = 1;
    int bits = 1;
    while((1ull << bits) <= s->ncells) ++bits;
    s->sortPasses = (bits + 7) / 8;
    for(int i = 0, left = bits; i < s->sortPasses; ++i) {
        s->sortBits[i] = (left + (s->sortPasses - i) - 1) / (s->sortPasses - i);
        left -= s->sortBits[i];
    }
    return SF_OK;
}

// dt of a slab substep (A.5 on the global max |v|^2) right before the force pass.  Normally the all-reduce is already
// in flight on the communication stream (issued by the previous substep's slab_exchange); after makeReady or a host
// upload it runs in-stream.  k_begin_step(kBeginClock) also resets the other max slot, which the integrate kernel of
// THIS substep accumulates into -- hence it must run after the all-reduce (in place on both slots) and before
// k_visc_brick, which the stream order guarantees.
int slab_global_dt(sf_solver* s)
{
    sf_solver::Slab& L = s->slab;
    if(L.dtReduced) {
        SF_CUDA(s, cudaStreamWaitEvent(s->stream, L.evDt, 0));
    } else if(nccl_api().AllReduce(s->B.state->maxv2Bits, s->B.state->maxv2Bits, 2, ncclUint32, ncclMax, L.comm, s->stream) != ncclSuccess) {
        return fail(s, SF_ERR_COMM, "ncclAllReduce(max |v|^2) failed");
    }
    L.dtReduced = false;
    LaunchScope ls(s, K_BEGIN);
    k_begin_step<<<1, 1, 0, s->stream>>>(s->B.state, s->P, kBeginClock);
    return SF_OK;
}

// Tail of a slab substep: integrate edge bricks, exchange them while the interior bricks integrate.
int slab_exchange(sf_solver* s, uint32_t n, uint32_t nSlots)
{
    sf_solver::Slab& L  = s->slab;
    DevBuffers&      B  = s->B;
    const DevParams  P  = s->P;
    cudaStream_t     cs = s->stream, ms = L.commStream;
    NcclApi&         nc = nccl_api();
    const uint32_t   extent = std::max(n, nSlots);
    const uint32_t   pairGrid = std::max<uint32_t>(1u, std::min<uint32_t>(s->numBricks, static_cast<uint32_t>(s->numSMs) * 2u));
    if(extent) k_fill_u32<<<std::min<uint32_t>(cdiv(extent, 256), s->numSMs * 8), 256, 0, cs>>>(B.idA, extent, kInvalidId);
    if(n) {
        LaunchScope ls(s, K_VISC_INTEGRATE);
        k_visc_brick<false><<<std::min<uint32_t>(pairGrid, s->numSMs * s->occVisc), kBrickThreads, kSmemPair, cs>>>(B, P, 1);
    }
    SF_CUDA(s, cudaEventRecord(L.evEdge, cs));
    tl_mark(s, 5, cs);
    static const int freeSlots = std::getenv("SF_SLAB_FREE_SLOTS") ? std::max(0, std::atoi(std::getenv("SF_SLAB_FREE_SLOTS"))) : 4;
    if(n) { // interior layers: force, then integrate; a few CTA slots stay free so that the pack / NCCL kernels can run beside them
        {
            LaunchScope    ls(s, K_FORCE);
            const uint32_t g = std::min<uint32_t>(pairGrid, s->numSMs * s->occForce);
            k_force_brick<<<g > 32 ? g - std::min<uint32_t>(freeSlots, g - 1) : g, kBrickThreads, kSmemPair, cs>>>(B, P, 2);
        }
        LaunchScope    ls(s, K_VISC_INTEGRATE);
        const uint32_t g = std::min<uint32_t>(pairGrid, s->numSMs * s->occVisc);
        k_visc_brick<false><<<g > 32 ? g - std::min<uint32_t>(freeSlots, g - 1) : g, kBrickThreads, kSmemPair, cs>>>(B, P, 2);
    }
    SF_CUDA(s, cudaEventRecord(L.evInterior, cs));
    tl_mark(s, 6, cs);
    // ---- communication stream
    const int hasLower = L.rank > 0, hasUpper = L.rank < L.nranks - 1;
    SF_CUDA(s, cudaStreamWaitEvent(ms, L.evEdge, 0));
    SF_CUDA(s, cudaMemsetAsync(L.counters, 0, 2 * sizeof(uint32_t), ms));
    if(n) k_slab_pack<<<std::min<uint32_t>(cdiv(n, 256), 64), 256, 0, ms>>>(B.posA, B.velA, B.idA, L.layerStart, P, L.next[L.rank], L.next[L.rank + 1],
                                                                               L.ghost, hasLower, hasUpper, L.sendLo, L.sendHi, L.sendCap, L.counters);
    k_slab_row<<<1, 1, 0, ms>>>(L.layerStart, L.counters, P, L.row);
    if(nc.AllGather(L.row, L.table, kRowWords, ncclUint32, L.comm, ms) != ncclSuccess) return fail(s, SF_ERR_COMM, "ncclAllGather failed");
    SF_CUDA(s, cudaMemcpyAsync(L.hostTable, L.table, sizeof(uint32_t) * kRowWords * L.nranks, cudaMemcpyDeviceToHost, ms));
    tl_mark(s, 7, ms);
    const auto tw0 = std::chrono::steady_clock::now();
    SF_CUDA(s, cudaStreamSynchronize(ms)); // the compute stream keeps integrating the interior bricks meanwhile
    const auto tw1 = std::chrono::steady_clock::now();
    L.tWaitPack += std::chrono::duration<double>(tw1 - tw0).count();
    const uint32_t* T      = L.hostTable;
    const uint32_t  sendLo = T[L.rank * kRowWords + 0], sendHi = T[L.rank * kRowWords + 1], nOwn = T[L.rank * kRowWords + 2];
    const uint32_t  recvLo = hasLower ? T[(L.rank - 1) * kRowWords + 1] : 0u;
    const uint32_t  recvHi = hasUpper ? T[(L.rank + 1) * kRowWords + 0] : 0u;
    // Capacities follow the load (a settling flow can concentrate in few slabs): everything below is rank-local.
    if(std::max(sendLo, sendHi) > L.sendCap) { // the pack kernel dropped the overflow: grow and pack again (same counts)
        L.sendCap = std::max(sendLo, sendHi) + std::max(sendLo, sendHi) / 4 + 4096u;
        SF_CUDA(s, dev_alloc(L.sendLo, static_cast<size_t>(L.sendCap) * 2));
        SF_CUDA(s, dev_alloc(L.sendHi, static_cast<size_t>(L.sendCap) * 2));
        SF_CUDA(s, cudaMemsetAsync(L.counters, 0, 2 * sizeof(uint32_t), ms));
        k_slab_pack<<<std::min<uint32_t>(cdiv(n, 256), 64), 256, 0, ms>>>(B.posA, B.velA, B.idA, L.layerStart, P, L.next[L.rank], L.next[L.rank + 1],
                                                                            L.ghost, hasLower, hasUpper, L.sendLo, L.sendHi, L.sendCap, L.counters);
        L.regrowths++;
    }
    if(std::max(recvLo, recvHi) > L.recvCap) {
        L.recvCap = std::max(recvLo, recvHi) + std::max(recvLo, recvHi) / 4 + 4096u;
        SF_CUDA(s, dev_alloc(L.recvLo, static_cast<size_t>(L.recvCap) * 2));
        SF_CUDA(s, dev_alloc(L.recvHi, static_cast<size_t>(L.recvCap) * 2));
        L.regrowths++;
    }
    if(static_cast<uint64_t>(n) + recvLo + recvHi > s->cap) {
        SF_CUDA(s, cudaStreamSynchronize(cs)); // the interior bricks still write the state arrays
        SF_CUDA(s, cudaStreamSynchronize(ms)); // a re-run pack kernel may still read them
        const uint64_t want = static_cast<uint64_t>(n) + recvLo + recvHi;
        if(want + want / 2 > 0xfffffff0ull) return fail(s, SF_ERR_OOM, "more than 2^32 particle slots on one rank");
        const int rc = ensure_particle_capacity(s, static_cast<uint32_t>(want + want / 2), n);
        if(rc) return rc;
        L.regrowths++;
    }
    const auto tg0 = std::chrono::steady_clock::now();
    if(nc.GroupStart() != ncclSuccess) return fail(s, SF_ERR_COMM, "ncclGroupStart failed");
    ncclResult_t r = ncclSuccess;
    if(hasLower) {
        if(sendLo && r == ncclSuccess) r = nc.Send(L.sendLo, static_cast<size_t>(sendLo) * 8, ncclFloat, L.rank - 1, L.comm, ms);
        if(recvLo && r == ncclSuccess) r = nc.Recv(L.recvLo, static_cast<size_t>(recvLo) * 8, ncclFloat, L.rank - 1, L.comm, ms);
    }
    if(hasUpper) {
        if(sendHi && r == ncclSuccess) r = nc.Send(L.sendHi, static_cast<size_t>(sendHi) * 8, ncclFloat, L.rank + 1, L.comm, ms);
        if(recvHi && r == ncclSuccess) r = nc.Recv(L.recvHi, static_cast<size_t>(recvHi) * 8, ncclFloat, L.rank + 1, L.comm, ms);
    }
    if(nc.GroupEnd() != ncclSuccess || r != ncclSuccess) return fail(s, SF_ERR_COMM, "halo send/recv failed");
    L.tGroup += std::chrono::duration<double>(std::chrono::steady_clock::now() - tg0).count();
    if(recvLo) k_slab_unpack<<<cdiv(recvLo, 256), 256, 0, ms>>>(L.recvLo, recvLo, B.posA + n, B.velA + n, B.idA + n);
    if(recvHi) k_slab_unpack<<<cdiv(recvHi, 256), 256, 0, ms>>>(L.recvHi, recvHi, B.posA + n + recvLo, B.velA + n + recvLo, B.idA + n + recvLo);
    SF_CUDA(s, cudaEventRecord(L.evExchanged, ms));
    tl_mark(s, 8, ms);
    SF_CUDA(s, cudaStreamWaitEvent(cs, L.evExchanged, 0));
    // max |v|^2 of this substep is complete once the interior bricks are integrated: reduce it over the ranks on the
    // communication stream; the next substep waits for it only before its force pass (slab_global_dt)
    SF_CUDA(s, cudaStreamWaitEvent(ms, L.evInterior, 0));
    const auto tr0 = std::chrono::steady_clock::now();
    if(nc.AllReduce(B.state->maxv2Bits, B.state->maxv2Bits, 2, ncclUint32, ncclMax, L.comm, ms) != ncclSuccess)
        return fail(s, SF_ERR_COMM, "ncclAllReduce(max |v|^2) failed");
    L.tReduce += std::chrono::duration<double>(std::chrono::steady_clock::now() - tr0).count();
    SF_CUDA(s, cudaEventRecord(L.evDt, ms));
    L.dtReduced = true;
    SF_CUDA(s, cudaGetLastError());
    // ---- bookkeeping for the next substep: live = my integrated particles + received; dead = this substep's ghosts
    s->nSlots = n + recvLo + recvHi;
    s->n      = nOwn + recvLo + recvHi;
    L.exchangedParticles += static_cast<uint64_t>(sendLo) + sendHi;
    std::vector<int32_t> following = L.next;
    slab_rebalance(T, kRowWords, L.nranks, s->nS(), kMinThick, following.data());
    L.cur  = L.next;
    L.next = following;
    L.nOwn = nOwn; // by the old bounds; exact count by the new bounds comes with the next table
    L.tPost += std::chrono::duration<double>(std::chrono::steady_clock::now() - tw1).count();
    L.steps++;
    return slab_configure_window(s);
}

} // namespace

// =================================================================================================
extern "C" {

// No C++ exception crosses the C boundary (include/sf_b200.h): every entry point that returns a status is a
// function-try-block; std::bad_alloc / std::length_error of a host container become SF_ERR_OOM, anything else
// SF_ERR_INVALID, with the text in sf_last_error.
#define SF_NOTHROW(S, NAME)                                                                                     \
    catch(const std::bad_alloc& e) { return fail((S), SF_ERR_OOM, std::string(NAME ": ") + e.what()); }         \
    catch(const std::length_error& e) { return fail((S), SF_ERR_OOM, std::string(NAME ": ") + e.what()); }      \
    catch(const std::exception& e) { return fail((S), SF_ERR_INVALID, std::string(NAME ": ") + e.what()); }     \
    catch(...) { return fail((S), SF_ERR_INVALID, NAME ": unknown exception"); }

int sf_params_default(sf_params* p)
try {
    if(!p) return SF_ERR_INVALID;
    params_default(*p);
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_params_default")

int sf_params_update(sf_params* p)
try {
    if(!p) return SF_ERR_INVALID;
    params_update(*p);
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_params_update")

int sf_params_set_resolution(sf_params* p, float resolution)
try {
    if(!p || !(resolution > 0.f)) return SF_ERR_INVALID;
    p->kernelRadius = 2.0f / resolution; // Source/Controller.cpp:55
    params_update(*p);
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_params_set_resolution")

int sf_scene_generate(const sf_params* p, int scene, float* pos_xyz, uint64_t cap, uint64_t* n_out)
try {
    if(!p || scene < 0 || scene > 3) return SF_ERR_INVALID;
    const uint64_t n = scene_generate(*p, scene, pos_xyz, cap);
    if(n_out) *n_out = n;
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_scene_generate")

int sf_build_tables(const sf_params* p, float* cubic_w10001, float* spiky_grad10001, float* consts3)
try {
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
SF_NOTHROW(nullptr, "sf_build_tables")

int sf_boundary_generate(const sf_params* p, uint32_t seed, int wall, float* xyz, uint32_t cap, uint32_t* n_out)
try {
    if(!p || wall < 0 || wall > 5) return SF_ERR_INVALID;
    std::vector<float> walls[6];
    generate_boundary(*p, seed, walls);
    const uint32_t n = static_cast<uint32_t>(walls[wall].size() / 3);
    if(n_out) *n_out = n;
    if(xyz) std::memcpy(xyz, walls[wall].data(), sizeof(float) * 3 * std::min(n, cap));
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_boundary_generate")

int sf_wall_candidate_masks(const sf_params* p, int wall, const float* xyz, uint32_t n, uint32_t words, uint32_t* masks)
try {
    if(!p || wall < 0 || wall > 5 || (!xyz && n) || !masks || words < (n + 31u) / 32u) return SF_ERR_INVALID;
    static_assert(SF_WALL_SUBCELLS == kWallSubCells, "public constant follows the internal one");
    wall_candidate_masks(*p, wall, xyz, n, words, masks);
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_wall_candidate_masks")

int sf_wall_subcell(const sf_params* p, int wall, const float pos_xyz[3], uint32_t* entry_out)
try {
    if(!p || wall < 0 || wall > 5 || !pos_xyz || !entry_out) return SF_ERR_INVALID;
    *entry_out = static_cast<uint32_t>(wall_subcell(*p, wall, pos_xyz));
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_wall_subcell")

int sf_create(const sf_params* p, int device, sf_solver** out)
try {
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
    sf_solver* s = new(std::nothrow) sf_solver();
    if(!s) return fail(nullptr, SF_ERR_OOM, "sf_create: out of host memory");
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
    if(e == cudaSuccess) e = cudaFuncSetAttribute(k_visc_brick<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPair));
    if(e == cudaSuccess) e = cudaFuncSetAttribute(k_visc_brick<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPair));
    if(e == cudaSuccess) e = cudaFuncSetAttribute(k_shepard_brick, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPair));
    if(e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->occDensity, k_density_brick, kBrickThreads, kSmemDensity);
    if(e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->occForce, k_force_brick, kBrickThreads, kSmemPair);
    if(e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->occVisc, k_visc_brick<true>, kBrickThreads, kSmemPair);
    if(e != cudaSuccess) {
        const std::string msg = std::string("sf_create: ") + cudaGetErrorString(e);
        sf_destroy(s);
        return fail(nullptr, SF_ERR_CUDA, msg);
    }
    s->stream = s->ownStream;
    s->useGraph = std::getenv("SF_NO_GRAPH") == nullptr;
    if(const char* m = std::getenv("SF_SORT")) s->countSort = std::strcmp(m, "radix") != 0;
    if(const char* m = std::getenv("SF_KMAX")) s->kmax = round_list_capacity(std::min(std::max(std::atoi(m), 8), 16380));
    s->occDensity = std::max(s->occDensity, 1);
    s->occForce   = std::max(s->occForce, 1);
    s->occVisc    = std::max(s->occVisc, 1);
    *out = s;
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_create")

void sf_destroy(sf_solver* s)
{
    if(!s) return;
    cudaSetDevice(s->device);
    if(s->stream) cudaStreamSynchronize(s->stream);
    drop_graph(s);
    fold_pending(s);
    for(auto e : s->eventPool) cudaEventDestroy(e);
    DevBuffers& B = s->B;
    cudaFree(B.posA); cudaFree(B.velA); cudaFree(B.posB); cudaFree(B.velB); cudaFree(B.idA); cudaFree(B.idB);
    for(int i = 0; i < 2; ++i) { cudaFree(B.keys[i]); cudaFree(B.vals[i]); }
    cudaFree(B.cellTab); cudaFree(B.cellCnt); cudaFree(B.rho); cudaFree(B.rho2); cudaFree(B.accel); cudaFree(B.nbrL); cudaFree(B.nbrCnt); cudaFree(B.brickFlag); cudaFree(B.brickList);
    cudaFree(B.tabW); cudaFree(B.tabG); cudaFree(B.bnd); cudaFree(B.wallMask); cudaFree(B.radixCounts); cudaFree(B.radixTotals); cudaFree(B.state);
    cudaFree(s->stage);
    cudaFree(s->cellTileSums);
    if(s->xferStream) {
        cudaStreamSynchronize(s->xferStream);
        cudaStreamDestroy(s->xferStream);
        cudaEventDestroy(s->evPosUp);
        cudaEventDestroy(s->evVelUp);
        cudaEventDestroy(s->evPosOut);
        cudaEventDestroy(s->evVelOut);
    }
    if(s->snapStream) {
        cudaStreamSynchronize(s->snapStream);
        cudaStreamDestroy(s->snapStream);
        cudaEventDestroy(s->snapReady);
        cudaEventDestroy(s->snapDone);
    }
    cudaFree(s->snapBuf);
    {
        sf_solver::Slab& L = s->slab;
        if(L.on && L.steps && std::getenv("SF_SLAB_TRACE"))
            std::fprintf(stderr, "[sf slab rank %d] %llu substeps: host wait for edge+pack+allgather %.3f ms/step, post exchange %.3f ms/step (send/recv group %.3f, dt all-reduce %.3f), exchanged %.1f particles/step, %llu capacity regrowths, axis %c\n",
                         L.rank, (unsigned long long)L.steps, L.tWaitPack / L.steps * 1e3, L.tPost / L.steps * 1e3, L.tGroup / L.steps * 1e3, L.tReduce / L.steps * 1e3, double(L.exchangedParticles) / L.steps,
                         (unsigned long long)L.regrowths, s->axisS == 1 ? 'y' : 'z');
        if(L.commStream) cudaStreamSynchronize(L.commStream); // the last substep's dt all-reduce may still be in flight
        slab_timeline_report(s);
        for(auto& e : L.tl) cudaEventDestroy(e);
        if(L.comm && nccl_api().CommDestroy) nccl_api().CommDestroy(L.comm);
        cudaFree(L.sendLo); cudaFree(L.sendHi); cudaFree(L.recvLo); cudaFree(L.recvHi);
        cudaFree(L.layerStart); cudaFree(L.counters); cudaFree(L.row); cudaFree(L.table);
        if(L.hostTable) cudaFreeHost(L.hostTable);
        if(L.evEdge) cudaEventDestroy(L.evEdge);
        if(L.evExchanged) cudaEventDestroy(L.evExchanged);
        if(L.evInterior) cudaEventDestroy(L.evInterior);
        if(L.evDt) cudaEventDestroy(L.evDt);
        if(L.commStream) cudaStreamDestroy(L.commStream);
    }
    if(s->hostState) cudaFreeHost(s->hostState);
    cudaFree(s->ownCounters);
    if(s->hostOwnCounters) cudaFreeHost(s->hostOwnCounters);
    if(s->timerA) cudaEventDestroy(s->timerA);
    if(s->timerB) cudaEventDestroy(s->timerB);
    if(s->ownStream) cudaStreamDestroy(s->ownStream);
    delete s;
}

const char* sf_last_error(sf_solver* s) { return s ? s->lastError.c_str() : g_createError.c_str(); }

int sf_set_params(sf_solver* s, const sf_params* p)
try {
    if(!s || !p) return SF_ERR_INVALID;
    s->params = *p;
    s->ready  = false; // tables / grid depend on kernelRadius: makeReady again (Simulator.cpp:42)
    return SF_OK;
}
SF_NOTHROW(s, "sf_set_params")

int sf_get_params(sf_solver* s, sf_params* p)
try {
    if(!s || !p) return SF_ERR_INVALID;
    *p = s->params;
    return SF_OK;
}
SF_NOTHROW(s, "sf_get_params")

int sf_set_stream(sf_solver* s, void* cuda_stream)
try {
    if(!s) return SF_ERR_INVALID;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    s->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s->ownStream;
    drop_graph(s);
    return SF_OK;
}
SF_NOTHROW(s, "sf_set_stream")

int sf_upload_particles(sf_solver* s, const float* pos_xyz, const float* vel_xyz, uint32_t n)
try {
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
    if(s->slab.on) return fail(s, SF_ERR_INVALID, "slab mode: use sf_upload_particles_global");
    int rc = ensure_particle_capacity(s, n);
    if(rc) return rc;
    s->n = s->nSlots = n;
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
SF_NOTHROW(s, "sf_upload_particles")

int sf_num_particles(sf_solver* s, uint32_t* n_out)
try {
    if(!s || !n_out) return SF_ERR_INVALID;
    *n_out = s->n;
    return SF_OK;
}
SF_NOTHROW(s, "sf_num_particles")

static int download_xyz(sf_solver* s, const float4* src, float* out)
{
    if(!s || !out) return SF_ERR_INVALID;
    if(!s->uploaded) return fail(s, SF_ERR_INVALID, "no particles uploaded");
    if(s->slab.on) return fail(s, SF_ERR_INVALID, "slab mode: use sf_download_owned");
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
try {
    if(!s) return SF_ERR_INVALID;
    generate_boundary(s->params, seed, s->walls);
    s->wallsSet = true;
    s->ready    = false;
    return SF_OK;
}
SF_NOTHROW(s, "sf_generate_boundary")

int sf_set_boundary_particles(sf_solver* s, int wall, const float* xyz, uint32_t n)
try {
    if(!s || wall < 0 || wall > 5 || (!xyz && n)) return SF_ERR_INVALID;
    s->walls[wall].assign(xyz, xyz + 3 * static_cast<size_t>(n));
    s->wallsSet = true;
    s->ready    = false;
    return SF_OK;
}
SF_NOTHROW(s, "sf_set_boundary_particles")

int sf_get_boundary_particles(sf_solver* s, int wall, float* xyz, uint32_t cap, uint32_t* n_out)
try {
    if(!s || wall < 0 || wall > 5) return SF_ERR_INVALID;
    const uint32_t n = static_cast<uint32_t>(s->walls[wall].size() / 3);
    if(n_out) *n_out = n;
    if(xyz) std::memcpy(xyz, s->walls[wall].data(), sizeof(float) * 3 * std::min(n, cap));
    return SF_OK;
}
SF_NOTHROW(s, "sf_get_boundary_particles")

int sf_make_ready(sf_solver* s)
try {
    if(!s) return SF_ERR_INVALID;
    if(!s->uploaded) return fail(s, SF_ERR_INVALID, "sf_make_ready: upload particles first");
    SF_CUDA(s, cudaSetDevice(s->device));
    params_update(s->params);
    build_tables(s->params.kernelRadius, s->tables);
    grid_dims(s->params, s->grid);
    const int zPad = s->slab.on ? 2 * kGhostMax : 0; // slab mode: room for any window [zb - ghost, ze + ghost)
    if(!s->slab.on) s->axisS = 2;
    s->ncells = static_cast<uint64_t>(s->grid[0]) * s->nM() * (s->nS() + zPad);
    if(s->ncells == 0 || s->ncells >= (1ull << 31)) return fail(s, SF_ERR_INVALID, "grid has no cells or more than 2^31 cells");
    if(s->slab.on && s->params.bCorrectDensity && s->slab.ghost < 4)
        return fail(s, SF_ERR_INVALID, "bCorrectDensity needs a fourth ghost layer: set it before sf_upload_particles_global");
    if(s->slab.on && s->slab.cur.empty()) return fail(s, SF_ERR_INVALID, "slab mode: use sf_upload_particles_global");
    if(s->params.bUseBoundaryParticles && !s->wallsSet) {
        generate_boundary(s->params, 0u, s->walls); // the reference seeds from std::random_device; we default to seed 0
        s->wallsSet = true;
    }
    if(s->ncells > s->cellCap || !s->B.cellTab) {
        SF_CUDA(s, dev_alloc(s->B.cellTab, s->ncells + 2));
        SF_CUDA(s, dev_alloc(s->B.cellCnt, s->ncells + 2));
        s->cellCap = s->ncells;
    }
    SF_CUDA(s, cudaMemsetAsync(s->B.cellCnt, 0, sizeof(uint32_t) * (s->ncells + 2), s->stream)); // the counting sort expects zeros
    s->hashed = false;
    if(s->countSort && (s->ncells > s->cellTileCap || !s->cellTileSums)) {
        SF_CUDA(s, dev_alloc(s->cellTileSums, static_cast<size_t>(cdiv(s->ncells, CS_TILE)) + 1));
        s->cellTileCap = s->ncells;
    }
    {
        const uint32_t nb = static_cast<uint32_t>((s->grid[0] + BX - 1) / BX) * ((s->nM() + BY - 1) / BY) * ((s->nS() + zPad + BZ - 1) / BZ);
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
    {   // candidate masks of the wall lists for the density pass (sf_host.cpp: wall_candidate_masks)
        const uint32_t        words = (s->bndStride + 31u) / 32u;
        const size_t          perWall = static_cast<size_t>(kWallSubCells + 1) * words;
        std::vector<uint32_t> masks(6 * perWall, 0u);
        for(int w = 0; w < 6; ++w)
            wall_candidate_masks(s->params, w, s->walls[w].data(), static_cast<uint32_t>(s->walls[w].size() / 3), words, masks.data() + w * perWall);
        SF_CUDA(s, dev_alloc(s->B.wallMask, masks.size()));
        SF_CUDA(s, cudaMemcpy(s->B.wallMask, masks.data(), sizeof(uint32_t) * masks.size(), cudaMemcpyHostToDevice)); // (synchronous: `masks` is a local)
    }

    // radix-sort plan: ceil(log2(ncells)) key bits in passes of at most 8 bits
    int bits = 1;
    while((1ull << bits) <= s->ncells) ++bits; // strictly more than ncells codes: the all-ones key marks dead slots
    s->sortPasses = (bits + 7) / 8;
    for(int i = 0, left = bits; i < s->sortPasses; ++i) {
        s->sortBits[i] = (left + (s->sortPasses - i) - 1) / (s->sortPasses - i);
        left -= s->sortBits[i];
    }
    fill_dev_params(s);
    if(s->slab.on) slab_configure_window(s);
    drop_graph(s);
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
    s->slab.dtReduced = false; // max |v|^2 was recomputed locally: the first substep all-reduces it in-stream
    s->ready = true;
    return SF_OK;
}
SF_NOTHROW(s, "sf_make_ready")

int sf_advance_frame(sf_solver* s, float* dt_out)
try {
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
SF_NOTHROW(s, "sf_advance_frame")

int sf_advance_steps(sf_solver* s, uint32_t nsteps, float* time_out)
try {
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
SF_NOTHROW(s, "sf_advance_steps")

int sf_advance_frame_time(sf_solver* s, double frame_time, float* time_out, uint32_t* nsteps_out)
try {
    int rc = require_ready(s);
    if(rc) return rc;
    if(!(frame_time > 0.0)) return fail(s, SF_ERR_INVALID, "frame_time must be positive");
    SF_CUDA(s, cudaSetDevice(s->device));
    if(s->slab.on) {
        // slab substeps synchronise with the host anyway (exchange sizes), so the loop of Simulator.cpp:46-51 runs on
        // the host; dt is identical on every rank (all-reduced max |v|^2), hence so is the number of substeps
        float    frameTime = 0.f;
        uint32_t k         = 0;
        while(static_cast<double>(frameTime) < frame_time) {
            rc = enqueue_substep(s);
            if(rc) return rc;
            rc = read_state(s);
            if(rc) return rc;
            frameTime = frameTime + s->hostState->dt;
            ++k;
        }
        if(time_out) *time_out = frameTime;
        if(nsteps_out) *nsteps_out = k;
        return SF_OK;
    }
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
SF_NOTHROW(s, "sf_advance_frame_time")

int sf_synchronize(sf_solver* s)
try {
    if(!s) return SF_ERR_INVALID;
    SF_CUDA(s, cudaSetDevice(s->device));
    SF_CUDA(s, cudaStreamSynchronize(s->stream));
    return SF_OK;
}
SF_NOTHROW(s, "sf_synchronize")

int sf_step_host(sf_solver* s, float* pos_xyz, float* vel_xyz, uint32_t n, float* dt_out)
try {
    if(!s || !pos_xyz || !vel_xyz) return SF_ERR_INVALID;
    if(s->slab.on) return fail(s, SF_ERR_INVALID, "slab mode: use sf_upload_local / sf_advance_frame / sf_download_local");
    const bool sameShape = s->uploaded && s->ready && n == s->n && n > 0;
    int        rc = invalidate_binning(s); // the positions come from the host
    if(rc) return rc;
    if(!sameShape) {
        rc = sf_upload_particles(s, pos_xyz, vel_xyz, n);
        if(rc) return rc;
        rc = sf_make_ready(s);
        if(rc) return rc;
        rc = enqueue_substep(s);
        if(rc) return rc;
        if(n) {
            LaunchScope ls(s, K_MARSHAL);
            float*      dpos = s->stage;
            float*      dvel = s->stage + 3 * static_cast<size_t>(s->npad);
            k_unpack_xyz<<<cdiv(n, 256), 256, 0, s->stream>>>(s->B.posA, s->B.idA, dpos, n);
            k_unpack_xyz<<<cdiv(n, 256), 256, 0, s->stream>>>(s->B.velA, s->B.idA, dvel, n);
            SF_CUDA(s, cudaMemcpyAsync(pos_xyz, dpos, static_cast<size_t>(n) * 12, cudaMemcpyDeviceToHost, s->stream));
            SF_CUDA(s, cudaMemcpyAsync(vel_xyz, dvel, static_cast<size_t>(n) * 12, cudaMemcpyDeviceToHost, s->stream));
        }
    } else {
        // Steady state: same particle count; the device state is overwritten from the host buffers (upload order,
        // id = index, so the previous sort order is dropped).  The copies run on a second stream:
        //   copy stream   : H2D positions | H2D velocities ........................ | D2H positions | D2H velocities
        //   compute stream:               | sort, cell tables, density | v gather, dt, force, XSPH + integrate | unpack
        // Sorting and the density pass need positions only, so the velocity upload hides behind them.
        SF_CUDA(s, cudaSetDevice(s->device));
        if(!s->xferStream) {
            SF_CUDA(s, cudaStreamCreateWithFlags(&s->xferStream, cudaStreamNonBlocking));
            SF_CUDA(s, cudaEventCreateWithFlags(&s->evPosUp, cudaEventDisableTiming));
            SF_CUDA(s, cudaEventCreateWithFlags(&s->evVelUp, cudaEventDisableTiming));
            SF_CUDA(s, cudaEventCreateWithFlags(&s->evPosOut, cudaEventDisableTiming));
            SF_CUDA(s, cudaEventCreateWithFlags(&s->evVelOut, cudaEventDisableTiming));
        }
        cudaStream_t cs = s->stream, xs = s->xferStream;
        float*       dpos = s->stage;
        float*       dvel = s->stage + 3 * static_cast<size_t>(s->npad);
        const size_t bytes = static_cast<size_t>(n) * 12;
        SF_CUDA(s, cudaMemcpyAsync(dpos, pos_xyz, bytes, cudaMemcpyHostToDevice, xs));
        SF_CUDA(s, cudaEventRecord(s->evPosUp, xs));
        SF_CUDA(s, cudaMemcpyAsync(dvel, vel_xyz, bytes, cudaMemcpyHostToDevice, xs));
        SF_CUDA(s, cudaEventRecord(s->evVelUp, xs));
        DevState init{};
        init.maxv2Bits[0] = init.maxv2Bits[1] = 0x00800000u; // FLT_MIN
        *s->hostState = init; // pinned: the asynchronous copy below reads it when the stream gets there
        SF_CUDA(s, cudaMemcpyAsync(s->B.state, s->hostState, sizeof(DevState), cudaMemcpyHostToDevice, cs));
        SF_CUDA(s, cudaStreamWaitEvent(cs, s->evPosUp, 0));
        {
            LaunchScope ls(s, K_MARSHAL);
            k_pack_pos<<<cdiv(n, 256), 256, 0, cs>>>(dpos, s->B.posA, s->B.idA, n, s->P, s->B.state);
        }
        rc = enqueue_substep_launches(s, dvel);
        if(rc) return rc;
        {
            LaunchScope ls(s, K_MARSHAL);
            k_unpack_xyz<<<cdiv(n, 256), 256, 0, cs>>>(s->B.posA, s->B.idA, dpos, n);
            SF_CUDA(s, cudaEventRecord(s->evPosOut, cs));
            k_unpack_xyz<<<cdiv(n, 256), 256, 0, cs>>>(s->B.velA, s->B.idA, dvel, n);
            SF_CUDA(s, cudaEventRecord(s->evVelOut, cs));
        }
        SF_CUDA(s, cudaStreamWaitEvent(xs, s->evPosOut, 0));
        SF_CUDA(s, cudaMemcpyAsync(pos_xyz, dpos, bytes, cudaMemcpyDeviceToHost, xs));
        SF_CUDA(s, cudaStreamWaitEvent(xs, s->evVelOut, 0));
        SF_CUDA(s, cudaMemcpyAsync(vel_xyz, dvel, bytes, cudaMemcpyDeviceToHost, xs));
    }
    rc = read_state(s);
    if(rc) return rc;
    if(s->xferStream) SF_CUDA(s, cudaStreamSynchronize(s->xferStream));
    if(dt_out) *dt_out = s->hostState->dt;
    return SF_OK;
}
SF_NOTHROW(s, "sf_step_host")

int sf_host_alloc(uint64_t bytes, void** out)
try {
    if(!out) return SF_ERR_INVALID;
    *out = nullptr;
    if(bytes == 0) return SF_OK;
    const cudaError_t e = cudaMallocHost(out, bytes);
    if(e != cudaSuccess) return fail(nullptr, e == cudaErrorMemoryAllocation ? SF_ERR_OOM : SF_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_host_alloc")

int sf_host_free(void* p)
try {
    if(!p) return SF_OK;
    const cudaError_t e = cudaFreeHost(p);
    if(e != cudaSuccess) return fail(nullptr, SF_ERR_CUDA, std::string("cudaFreeHost: ") + cudaGetErrorString(e));
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_host_free")

int sf_debug_counters(sf_solver* s, uint64_t out[8])
try {
    int rc = require_ready(s);
    if(rc) return rc;
    if(!out) return SF_ERR_INVALID;
    SF_CUDA(s, cudaSetDevice(s->device));
    rc = read_state(s);
    if(rc) return rc;
    for(int i = 0; i < 8; ++i) out[i] = s->hostState->dbg[i];
    return SF_OK;
}
SF_NOTHROW(s, "sf_debug_counters")

int sf_set_list_capacity(sf_solver* s, int kmax)
try {
    if(!s || kmax < 8 || kmax > 16383) return SF_ERR_INVALID;
    if(s->B.posA) return fail(s, SF_ERR_INVALID, "sf_set_list_capacity: call before the first upload");
    s->kmax = round_list_capacity(kmax);
    return SF_OK;
}
SF_NOTHROW(s, "sf_set_list_capacity")

int sf_set_capture(sf_solver* s, int on)
try {
    if(!s) return SF_ERR_INVALID;
    s->capture   = on != 0;
    s->P.capture = on ? 1 : 0;
    drop_graph(s);
    return SF_OK;
}
SF_NOTHROW(s, "sf_set_capture")

int sf_grid_dims(sf_solver* s, int32_t n3[3])
try {
    if(!s || !n3) return SF_ERR_INVALID;
    grid_dims(s->params, n3);
    return SF_OK;
}
SF_NOTHROW(s, "sf_grid_dims")

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

// The production neighbour list of the last substep (what k_density_brick wrote and the two list walkers read),
// decoded to original ids: counts = raw packed nbrCnt per original id (0xffffffff: no list, the particle took the
// traversal path), ids / tabIdx = CSR in list order.
static int production_lists_host(sf_solver* s, std::vector<uint32_t>& counts, std::vector<uint32_t>* ids, std::vector<uint32_t>* tabIdx)
{
    const uint32_t n = s->n;
    counts.assign(n, 0);
    if(ids) ids->clear();
    if(tabIdx) tabIdx->clear();
    if(n == 0) return SF_OK;
    if(!s->B.keyB) return fail(s, SF_ERR_INVALID, "no substep has run yet");
    if(s->params.bCorrectDensity) return fail(s, SF_ERR_INVALID, "production list download is not available with bCorrectDensity");
    k_unpack_u32<<<cdiv(n, 256), 256, 0, s->stream>>>(s->B.nbrCnt, s->B.idA, reinterpret_cast<uint32_t*>(s->stage), n, 0u);
    SF_CUDA(s, cudaMemcpyAsync(counts.data(), s->stage, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s->stream));
    SF_CUDA(s, cudaStreamSynchronize(s->stream));
    if(!ids && !tabIdx) return SF_OK;
    std::vector<unsigned long long> offset(static_cast<size_t>(n) + 1, 0);
    for(uint32_t i = 0; i < n; ++i) offset[i + 1] = offset[i] + (counts[i] == kCntNoList ? 0u : (counts[i] & 16383u));
    const unsigned long long total = offset[n];
    if(ids) ids->assign(total, 0);
    if(tabIdx) tabIdx->assign(total, 0);
    if(total == 0) return SF_OK;
    unsigned long long* dOff = nullptr;
    uint32_t *          dIds = nullptr, *dIdx = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&dOff), sizeof(unsigned long long) * (static_cast<size_t>(n) + 1));
    if(e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&dIds), sizeof(uint32_t) * total);
    if(e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&dIdx), sizeof(uint32_t) * total);
    if(e == cudaSuccess) e = cudaMemcpyAsync(dOff, offset.data(), sizeof(unsigned long long) * (static_cast<size_t>(n) + 1), cudaMemcpyHostToDevice, s->stream);
    if(e == cudaSuccess) {
        k_list_decode<<<std::max<uint32_t>(1u, s->numBricks), kDecodeThreads, 0, s->stream>>>(s->B, s->P, dOff, dIds, dIdx);
        if(ids) e = cudaMemcpyAsync(ids->data(), dIds, sizeof(uint32_t) * total, cudaMemcpyDeviceToHost, s->stream);
        if(e == cudaSuccess && tabIdx) e = cudaMemcpyAsync(tabIdx->data(), dIdx, sizeof(uint32_t) * total, cudaMemcpyDeviceToHost, s->stream);
    }
    if(e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(dOff);
    cudaFree(dIds);
    cudaFree(dIdx);
    SF_CUDA(s, e);
    return SF_OK;
}

int sf_field_size(sf_solver* s, int field, uint64_t* bytes_out)
try {
    if(!s || !bytes_out) return SF_ERR_INVALID;
    const uint64_t n = s->n;
    switch(field) {
        case SF_FIELD_DENSITY:
        case SF_FIELD_PRESSURE:
        case SF_FIELD_CELL_INDEX:
        case SF_FIELD_NEIGHBOR_COUNT:
        case SF_FIELD_LIST_COUNTS:
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
        case SF_FIELD_LIST_IDS:
        case SF_FIELD_LIST_TABLE_INDEX: {
            int rc = require_ready(s);
            if(rc) return rc;
            SF_CUDA(s, cudaSetDevice(s->device));
            SF_CUDA(s, cudaStreamSynchronize(s->stream));
            std::vector<uint32_t> counts;
            rc = production_lists_host(s, counts, nullptr, nullptr);
            if(rc) return rc;
            uint64_t total = 0;
            for(uint32_t c : counts) total += c == kCntNoList ? 0u : (c & 16383u);
            *bytes_out = 4 * total;
            return SF_OK;
        }
        default: return fail(s, SF_ERR_INVALID, "unknown field");
    }
}
SF_NOTHROW(s, "sf_field_size")

int sf_download_field(sf_solver* s, int field, void* out, uint64_t bytes)
try {
    int rc = require_ready(s);
    if(rc) return rc;
    if(!out) return SF_ERR_INVALID;
    if(s->slab.on && field != SF_FIELD_TABLE_CUBIC_W && field != SF_FIELD_TABLE_SPIKY_GRAD)
        return fail(s, SF_ERR_INVALID, "per-particle fields are not gathered in slab mode");
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
    if(field == SF_FIELD_LIST_COUNTS || field == SF_FIELD_LIST_IDS || field == SF_FIELD_LIST_TABLE_INDEX) {
        std::vector<uint32_t> counts, ids, idx;
        rc = production_lists_host(s, counts, field == SF_FIELD_LIST_IDS ? &ids : nullptr, field == SF_FIELD_LIST_TABLE_INDEX ? &idx : nullptr);
        if(rc) return rc;
        const std::vector<uint32_t>& src = field == SF_FIELD_LIST_IDS ? ids : (field == SF_FIELD_LIST_TABLE_INDEX ? idx : counts);
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
            k_unpack_u32<<<g, 256, 0, s->stream>>>(s->B.keyB, s->B.idA, reinterpret_cast<uint32_t*>(s->stage), n, 0u);
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
SF_NOTHROW(s, "sf_download_field")

// ---- diagnostics of the last substep ---------------------------------------------------------------
int sf_diagnostics(sf_solver* s, uint64_t out[8])
try {
    int rc = require_ready(s);
    if(rc) return rc;
    if(!out) return SF_ERR_INVALID;
    SF_CUDA(s, cudaSetDevice(s->device));
    rc = read_state(s);
    if(rc) return rc;
    const DevState& st = *s->hostState;
    unsigned long long h[3] = { 0ull, 0ull, 0ull };
    if(s->n && s->B.keyB && st.stepsDone) {
        unsigned long long* d = nullptr;
        SF_CUDA(s, cudaMalloc(reinterpret_cast<void**>(&d), sizeof(h)));
        cudaError_t e = cudaMemsetAsync(d, 0, sizeof(h), s->stream);
        if(e == cudaSuccess) {
            k_nbr_stats<<<std::min<uint32_t>(cdiv(s->n, 256), s->numSMs * 8), 256, 0, s->stream>>>(s->B.nbrCnt, s->B.keyB, s->P, d);
            e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, s->stream);
        }
        if(e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        cudaFree(d);
        SF_CUDA(s, e);
    }
    out[0] = st.stepsDone;
    out[1] = st.brickCount;
    out[2] = st.fallbackBricks;
    out[3] = st.fallbackParticles;
    out[4] = h[0];
    out[5] = h[2];
    out[6] = h[1];
    out[7] = s->n;
    return SF_OK;
}
SF_NOTHROW(s, "sf_diagnostics")

// ---- measurement -------------------------------------------------------------------------------
int sf_profile_enable(sf_solver* s, int on)
try {
    if(!s) return SF_ERR_INVALID;
    cudaSetDevice(s->device);
    if(!on) fold_pending(s);
    s->profEvery = on > 0 ? static_cast<uint32_t>(on) : 0u;
    s->profCount = 0;
    s->profiling = false;
    return SF_OK;
}
SF_NOTHROW(s, "sf_profile_enable")

int sf_profile_reset(sf_solver* s)
try {
    if(!s) return SF_ERR_INVALID;
    cudaSetDevice(s->device);
    fold_pending(s);
    std::memset(s->profMs, 0, sizeof(s->profMs));
    std::memset(s->profLaunches, 0, sizeof(s->profLaunches));
    return SF_OK;
}
SF_NOTHROW(s, "sf_profile_reset")

int sf_profile_get(sf_solver* s, char* names_buf, size_t names_cap, double* ms, uint64_t* launches, uint32_t cap, uint32_t* count_out)
try {
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
SF_NOTHROW(s, "sf_profile_get")

int sf_launch_count(sf_solver* s, uint64_t* n_out)
try {
    if(!s || !n_out) return SF_ERR_INVALID;
    *n_out = s->launches;
    return SF_OK;
}
SF_NOTHROW(s, "sf_launch_count")

int sf_timer_start(sf_solver* s)
try {
    if(!s) return SF_ERR_INVALID;
    SF_CUDA(s, cudaSetDevice(s->device));
    SF_CUDA(s, cudaEventRecord(s->timerA, s->stream));
    return SF_OK;
}
SF_NOTHROW(s, "sf_timer_start")

int sf_timer_stop(sf_solver* s, float* ms_out)
try {
    if(!s || !ms_out) return SF_ERR_INVALID;
    SF_CUDA(s, cudaSetDevice(s->device));
    SF_CUDA(s, cudaEventRecord(s->timerB, s->stream));
    SF_CUDA(s, cudaEventSynchronize(s->timerB));
    SF_CUDA(s, cudaEventElapsedTime(ms_out, s->timerA, s->timerB));
    if(s->slab.on && s->slab.tlLevel >= 2) slab_timeline_report(s); // SF_SLAB_TRACE=2: the region just timed
    return SF_OK;
}
SF_NOTHROW(s, "sf_timer_stop")

// ---- renderer hand-off: asynchronous position snapshot -------------------
// FluidRenderWidget::updateParticleData uploads the "Position" array once per particleChanged signal
// (Source/FluidRenderWidget.cpp:204-221).  The snapshot is scattered to original order on the compute stream,
// copied to the (ideally pinned) host buffer on a separate copy stream, and the solver keeps stepping meanwhile.
int sf_snapshot_positions_async(sf_solver* s, float* host_xyz)
try {
    if(!s || !host_xyz) return SF_ERR_INVALID;
    if(!s->uploaded) return fail(s, SF_ERR_INVALID, "no particles uploaded");
    if(s->slab.on) return fail(s, SF_ERR_INVALID, "slab mode: use sf_download_owned");
    SF_CUDA(s, cudaSetDevice(s->device));
    if(!s->snapStream) {
        SF_CUDA(s, cudaStreamCreateWithFlags(&s->snapStream, cudaStreamNonBlocking));
        SF_CUDA(s, cudaEventCreateWithFlags(&s->snapReady, cudaEventDisableTiming));
        SF_CUDA(s, cudaEventCreateWithFlags(&s->snapDone, cudaEventDisableTiming));
    }
    if(s->snapCap < s->n || !s->snapBuf) {
        SF_CUDA(s, cudaStreamSynchronize(s->snapStream));
        SF_CUDA(s, dev_alloc(s->snapBuf, static_cast<size_t>(s->npad) * 3));
        s->snapCap = s->npad;
    }
    if(s->n) {
        SF_CUDA(s, cudaStreamWaitEvent(s->stream, s->snapDone, 0)); // previous snapshot has left the buffer
        {
            LaunchScope ls(s, K_MARSHAL);
            k_unpack_xyz<<<cdiv(s->n, 256), 256, 0, s->stream>>>(s->B.posA, s->B.idA, s->snapBuf, s->n);
        }
        SF_CUDA(s, cudaEventRecord(s->snapReady, s->stream));
        SF_CUDA(s, cudaStreamWaitEvent(s->snapStream, s->snapReady, 0));
        SF_CUDA(s, cudaMemcpyAsync(host_xyz, s->snapBuf, static_cast<size_t>(s->n) * 12, cudaMemcpyDeviceToHost, s->snapStream));
    }
    SF_CUDA(s, cudaEventRecord(s->snapDone, s->snapStream));
    return SF_OK;
}
SF_NOTHROW(s, "sf_snapshot_positions_async")

int sf_snapshot_wait(sf_solver* s)
try {
    if(!s) return SF_ERR_INVALID;
    if(!s->snapStream) return SF_OK;
    SF_CUDA(s, cudaSetDevice(s->device));
    SF_CUDA(s, cudaEventSynchronize(s->snapDone));
    return SF_OK;
}
SF_NOTHROW(s, "sf_snapshot_wait")

// ---- checkpoint / restart ------------------------------------------------
// State = {params, wall particle sets, simulated time, particles {id, position, velocity}}.  A single-GPU run
// writes one file (ids implicit: original order); a slab run writes one part per rank, `<path>.<rank>`, holding the
// particles that rank owns with their global ids.  Restarting continues bit-identically: dt is recomputed from max
// |v|^2 (an exact max), the sort re-derives the cell order from positions and ids alone, and the result of a slab
// run does not depend on where the cut planes lie -- so a checkpoint may be restarted on ANY number of ranks
// (sf_checkpoint_read: one GPU; sf_checkpoint_read_slab: a rank of a slab run), the cut planes are re-planned.
namespace
{
struct CheckpointHeader {
    char     magic[8]; // "SFCKPT2\0"
    uint32_t paramsBytes, n, wallCount[6];
    float    simTime;
    uint32_t hasWalls;   // 0: written before the wall particles were generated / set
    uint32_t hasIds;     // 1: a part of a slab checkpoint, ids follow the velocities
    uint32_t part, parts; // this part / number of parts
    uint64_t nGlobal;
};

struct CheckpointData {
    sf_params             p{};
    std::vector<float>    walls[6], x, v; // x, v: global arrays in id order once all parts are merged
    bool                  hasWalls = false;
    float                 simTime = 0.f;
    uint64_t              nGlobal = 0;
};

long file_size(FILE* f)
{
    const long at = std::ftell(f);
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, at, SEEK_SET);
    return sz;
}

// reads one file (whole checkpoint or one part) and merges its particles into d; returns false on any inconsistency
bool checkpoint_read_file(const std::string& path, CheckpointData& d, uint32_t expectPart, uint32_t* partsOut, std::vector<unsigned char>* seen)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if(!f) return false;
    CheckpointHeader h{};
    sf_params        p{};
    bool ok = std::fread(&h, sizeof(h), 1, f) == 1 && std::memcmp(h.magic, "SFCKPT2", 8) == 0 && h.paramsBytes == sizeof(sf_params) &&
              std::fread(&p, sizeof(p), 1, f) == 1;
    if(ok) { // every count is checked against the file size before anything is allocated from it
        uint64_t need = sizeof(h) + sizeof(p) + static_cast<uint64_t>(h.n) * (h.hasIds ? 28u : 24u);
        for(int w = 0; w < 6; ++w) need += static_cast<uint64_t>(h.wallCount[w]) * 12u;
        ok = need == static_cast<uint64_t>(file_size(f)) && h.parts >= 1 && h.part == expectPart && h.part < h.parts &&
             (h.hasIds ? h.nGlobal >= h.n && h.nGlobal <= 0xfffffff0ull : (h.parts == 1 && h.nGlobal == h.n));
    }
    std::vector<float> walls[6];
    for(int w = 0; w < 6 && ok; ++w) {
        walls[w].resize(static_cast<size_t>(h.wallCount[w]) * 3);
        if(h.wallCount[w]) ok = std::fread(walls[w].data(), 12, h.wallCount[w], f) == h.wallCount[w];
    }
    if(ok && expectPart == 0) {
        d.p        = p;
        d.hasWalls = h.hasWalls != 0;
        d.simTime  = h.simTime;
        d.nGlobal  = h.nGlobal;
        for(int w = 0; w < 6; ++w) d.walls[w] = walls[w];
        d.x.assign(static_cast<size_t>(h.nGlobal) * 3, 0.f);
        d.v.assign(static_cast<size_t>(h.nGlobal) * 3, 0.f);
        if(seen) seen->assign(static_cast<size_t>(h.nGlobal), 0);
    }
    if(ok) ok = h.nGlobal == d.nGlobal && std::memcmp(&p, &d.p, sizeof(p)) == 0 && h.simTime == d.simTime;
    if(ok && h.n) {
        if(!h.hasIds) {
            ok = std::fread(d.x.data(), 12, h.n, f) == h.n && std::fread(d.v.data(), 12, h.n, f) == h.n;
            if(seen) std::fill(seen->begin(), seen->end(), 1);
        } else {
            std::vector<float>    x(static_cast<size_t>(h.n) * 3), v(static_cast<size_t>(h.n) * 3);
            std::vector<uint32_t> id(h.n);
            ok = std::fread(x.data(), 12, h.n, f) == h.n && std::fread(v.data(), 12, h.n, f) == h.n && std::fread(id.data(), 4, h.n, f) == h.n;
            for(uint32_t i = 0; i < h.n && ok; ++i) {
                ok = id[i] < d.nGlobal && !(*seen)[id[i]]; // every global id exactly once over all parts
                if(!ok) break;
                (*seen)[id[i]] = 1;
                std::memcpy(&d.x[3 * static_cast<size_t>(id[i])], &x[3 * static_cast<size_t>(i)], 12);
                std::memcpy(&d.v[3 * static_cast<size_t>(id[i])], &v[3 * static_cast<size_t>(i)], 12);
            }
        }
    }
    std::fclose(f);
    if(partsOut) *partsOut = h.parts;
    return ok;
}

// single file `path`, or the parts `path.0 .. path.<parts-1>` of a slab checkpoint
int checkpoint_load(const char* path, CheckpointData& d)
{
    try {
        std::vector<unsigned char> seen;
        uint32_t                   parts = 1;
        FILE*                      probe = std::fopen(path, "rb");
        if(probe) {
            std::fclose(probe);
            if(!checkpoint_read_file(path, d, 0, &parts, &seen) || parts != 1) return fail(nullptr, SF_ERR_INVALID, std::string("not a valid checkpoint: ") + path);
            return SF_OK;
        }
        const std::string base(path);
        if(!checkpoint_read_file(base + ".0", d, 0, &parts, &seen)) return fail(nullptr, SF_ERR_INVALID, std::string("not a valid checkpoint: ") + path + "[.0]");
        for(uint32_t k = 1; k < parts; ++k)
            if(!checkpoint_read_file(base + "." + std::to_string(k), d, k, nullptr, &seen))
                return fail(nullptr, SF_ERR_INVALID, std::string("missing or inconsistent checkpoint part ") + std::to_string(k) + " of " + path);
        for(unsigned char c : seen)
            if(!c) return fail(nullptr, SF_ERR_INVALID, std::string("checkpoint parts do not cover every particle id: ") + path);
        return SF_OK;
    } catch(const std::exception& e) { // bad_alloc / length_error never cross the C boundary
        return fail(nullptr, SF_ERR_OOM, std::string("sf_checkpoint_read: ") + e.what());
    }
}
} // namespace

int sf_checkpoint_write(sf_solver* s, const char* path, float sim_time)
try {
    if(!s || !path) return SF_ERR_INVALID;
    if(!s->uploaded) return fail(s, SF_ERR_INVALID, "no particles uploaded");
    try {
        const bool            slab = s->slab.on;
        uint32_t              n    = s->n;
        std::vector<float>    x, v;
        std::vector<uint32_t> id;
        int                   rc;
        if(slab) {
            if(!s->ready) return fail(s, SF_ERR_INVALID, "slab mode: sf_make_ready before sf_checkpoint_write");
            rc = sf_download_owned(s, nullptr, nullptr, nullptr, 0, &n);
            if(rc) return rc;
            x.resize(static_cast<size_t>(n) * 3);
            v.resize(static_cast<size_t>(n) * 3);
            id.resize(n);
            uint32_t n2 = 0;
            rc = sf_download_owned(s, id.data(), x.data(), v.data(), n, &n2);
            if(rc) return rc;
            if(n2 != n) return fail(s, SF_ERR_STATE, "owned particle count changed during the checkpoint");
        } else {
            x.resize(static_cast<size_t>(n) * 3);
            v.resize(static_cast<size_t>(n) * 3);
            rc = sf_download_positions(s, x.data());
            if(rc) return rc;
            rc = sf_download_velocities(s, v.data());
            if(rc) return rc;
        }
        CheckpointHeader h{};
        std::memcpy(h.magic, "SFCKPT2", 8);
        h.paramsBytes = sizeof(sf_params);
        h.n           = n;
        h.simTime     = sim_time;
        h.hasWalls    = s->wallsSet ? 1u : 0u;
        h.hasIds      = slab ? 1u : 0u;
        h.part        = slab ? static_cast<uint32_t>(s->slab.rank) : 0u;
        h.parts       = slab ? static_cast<uint32_t>(s->slab.nranks) : 1u;
        h.nGlobal     = slab ? s->slab.nGlobal : n;
        for(int w = 0; w < 6; ++w) h.wallCount[w] = s->wallsSet ? static_cast<uint32_t>(s->walls[w].size() / 3) : 0u;
        const std::string file = slab ? std::string(path) + "." + std::to_string(s->slab.rank) : std::string(path);
        FILE* f = std::fopen(file.c_str(), "wb");
        if(!f) return fail(s, SF_ERR_INVALID, std::string("cannot open ") + file);
        bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 && std::fwrite(&s->params, sizeof(sf_params), 1, f) == 1;
        for(int w = 0; w < 6 && ok; ++w)
            if(h.wallCount[w]) ok = std::fwrite(s->walls[w].data(), 12, h.wallCount[w], f) == h.wallCount[w];
        if(ok && n) ok = std::fwrite(x.data(), 12, n, f) == n && std::fwrite(v.data(), 12, n, f) == n;
        if(ok && n && slab) ok = std::fwrite(id.data(), 4, n, f) == n;
        ok = (std::fclose(f) == 0) && ok;
        return ok ? SF_OK : fail(s, SF_ERR_INVALID, std::string("short write to ") + file);
    } catch(const std::exception& e) {
        return fail(s, SF_ERR_OOM, std::string("sf_checkpoint_write: ") + e.what());
    }
}
SF_NOTHROW(s, "sf_checkpoint_write")

static int checkpoint_restore(const char* path, int device, int rank, int nranks, const void* id128, sf_solver** out, float* sim_time)
{
    if(!path || !out) return fail(nullptr, SF_ERR_INVALID, "null argument");
    *out = nullptr;
    CheckpointData d;
    int            rc = checkpoint_load(path, d);
    if(rc) return rc;
    sf_solver* s = nullptr;
    rc           = sf_create(&d.p, device, &s);
    if(rc) return rc;
    if(nranks > 1) rc = sf_comm_init(s, rank, nranks, id128);
    // a checkpoint written before the walls existed restores none: sf_make_ready then generates the default ones
    for(int w = 0; w < 6 && !rc && d.hasWalls; ++w) rc = sf_set_boundary_particles(s, w, d.walls[w].data(), static_cast<uint32_t>(d.walls[w].size() / 3));
    if(!rc) rc = sf_upload_particles_global(s, d.x.data(), d.v.data(), static_cast<uint32_t>(d.nGlobal));
    if(!rc) rc = sf_make_ready(s);
    if(rc) {
        g_createError = s->lastError;
        sf_destroy(s);
        return rc;
    }
    if(sim_time) *sim_time = d.simTime;
    *out = s;
    return SF_OK;
}

int sf_checkpoint_read(const char* path, int device, sf_solver** out, float* sim_time)
try {
    return checkpoint_restore(path, device, 0, 1, nullptr, out, sim_time);
}
SF_NOTHROW(nullptr, "sf_checkpoint_read")

int sf_checkpoint_read_slab(const char* path, int device, int rank, int nranks, const void* id128, sf_solver** out, float* sim_time)
try {
    if(nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) return fail(nullptr, SF_ERR_INVALID, "bad rank / nranks / id");
    return checkpoint_restore(path, device, rank, nranks, id128, out, sim_time);
}
SF_NOTHROW(nullptr, "sf_checkpoint_read_slab")

// ---- multi-GPU: z-slab decomposition (design notes in sf_slab.cuh) -----------------------------
int sf_comm_unique_id(void* id128)
try {
    if(!id128) return SF_ERR_INVALID;
    NcclApi& nc = nccl_api();
    if(!nc.lib || !nc.GetUniqueId) return fail(nullptr, SF_ERR_COMM, "libnccl.so.2 not found");
    ncclUniqueId id;
    if(nc.GetUniqueId(&id) != ncclSuccess) return fail(nullptr, SF_ERR_COMM, "ncclGetUniqueId failed");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(id128, &id, 128);
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_comm_unique_id")

int sf_comm_init(sf_solver* s, int rank, int nranks, const void* id128)
try {
    if(!s || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return SF_ERR_INVALID;
    if(nranks == 1) return SF_OK;
    NcclApi& nc = nccl_api();
    if(!nc.lib || !nc.CommInitRank) return fail(s, SF_ERR_COMM, "libnccl.so.2 not found");
    SF_CUDA(s, cudaSetDevice(s->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclResult_t r = nc.CommInitRank(&s->slab.comm, nranks, id, rank);
    if(r != ncclSuccess) return fail(s, SF_ERR_COMM, std::string("ncclCommInitRank: ") + nc.GetErrorString(r));
    sf_solver::Slab& L = s->slab;
    L.on     = true;
    L.rank   = rank;
    L.nranks = nranks;
    if(const char* t = std::getenv("SF_SLAB_TRACE")) L.tlLevel = std::atoi(t);
    int lo = 0, hi = 0;
    SF_CUDA(s, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SF_CUDA(s, cudaStreamCreateWithPriority(&L.commStream, cudaStreamNonBlocking, hi));
    SF_CUDA(s, cudaEventCreateWithFlags(&L.evEdge, cudaEventDisableTiming));
    SF_CUDA(s, cudaEventCreateWithFlags(&L.evExchanged, cudaEventDisableTiming));
    SF_CUDA(s, cudaEventCreateWithFlags(&L.evInterior, cudaEventDisableTiming));
    SF_CUDA(s, cudaEventCreateWithFlags(&L.evDt, cudaEventDisableTiming));
    SF_CUDA(s, dev_alloc(L.counters, 2));
    SF_CUDA(s, dev_alloc(L.row, kRowWords));
    SF_CUDA(s, dev_alloc(L.table, static_cast<size_t>(kRowWords) * nranks));
    SF_CUDA(s, cudaMallocHost(reinterpret_cast<void**>(&L.hostTable), sizeof(uint32_t) * kRowWords * nranks));
    return SF_OK;
}
SF_NOTHROW(s, "sf_comm_init")

static int upload_particles_global(sf_solver* s, const float* pos_xyz, const float* vel_xyz, uint32_t n_global);
int sf_upload_particles_global(sf_solver* s, const float* pos_xyz, const float* vel_xyz, uint32_t n_global)
try {
    if(!s || (!pos_xyz && n_global)) return SF_ERR_INVALID;
    try {
        return upload_particles_global(s, pos_xyz, vel_xyz, n_global);
    } catch(const std::exception& e) { // host vectors of the scatter: nothing throws across the C boundary
        return fail(s, SF_ERR_OOM, std::string("sf_upload_particles_global: ") + e.what());
    }
}
SF_NOTHROW(s, "sf_upload_particles_global")

static int upload_particles_global(sf_solver* s, const float* pos_xyz, const float* vel_xyz, uint32_t n_global)
{
    if(!s->slab.on) return sf_upload_particles(s, pos_xyz, vel_xyz, n_global);
    SF_CUDA(s, cudaSetDevice(s->device));
    sf_solver::Slab& L = s->slab;
    int32_t g[3];
    grid_dims(s->params, g);
    std::vector<uint64_t> histY(g[1], 0), histZ(g[2], 0);
    std::vector<int32_t>  layerY(n_global), layerZ(n_global);
    for(uint32_t i = 0; i < n_global; ++i) {
        int32_t c[3];
        if(!cell_coords_checked(s->params, g, pos_xyz + 3 * static_cast<size_t>(i), c)) return fail(s, SF_ERR_DOMAIN, "particle outside the simulation box");
        layerY[i] = c[1];
        layerZ[i] = c[2];
        histY[c[1]]++;
        histZ[c[2]]++;
    }
    // slab axis: z (orthogonal to gravity, so a settling flow keeps spanning it) unless the particle set spans at
    // least twice as many cell layers in y: thicker slabs, so the 3 + 3 ghost layers weigh less (64 M dambreak: 283 y
    // layers against 101 z layers).  Every rank sees the same input and takes the same decision.
    auto occupied = [](const std::vector<uint64_t>& h) { size_t k = 0; for(uint64_t v : h) k += v ? 1 : 0; return k; };
    if(const char* force = std::getenv("SF_SLAB_AXIS")) s->axisS = (force[0] == 'y' || force[0] == 'Y' || force[0] == '1') ? 1 : 2;
    else s->axisS = occupied(histY) >= 2 * occupied(histZ) ? 1 : 2;
    const std::vector<uint64_t>& hist  = s->axisS == 1 ? histY : histZ;
    const std::vector<int32_t>&  layer = s->axisS == 1 ? layerY : layerZ;
    s->grid[0] = g[0]; s->grid[1] = g[1]; s->grid[2] = g[2];
    if(s->nS() < L.nranks * kMinThick) return fail(s, SF_ERR_INVALID, "grid has too few cell layers for this many slabs");
    L.cur.assign(L.nranks + 1, 0);
    slab_plan(hist.data(), s->nS(), L.nranks, kMinThick, L.cur.data());
    L.next = L.cur;
    const int zb = L.cur[L.rank], ze = L.cur[L.rank + 1];
    uint64_t inWindow = 0; // particles of [zb - 3, ze + 3): from the layer histogram
    L.ghost = s->params.bCorrectDensity ? 4 : 3;
    const int kGhost = L.ghost;
    for(int l = std::max(zb - kGhost, 0); l < std::min(ze + kGhost, s->nS()); ++l) inWindow += hist[l];
    if(inWindow > 0xfffffff0ull) return fail(s, SF_ERR_OOM, "more than 2^32 particles on one rank");
    std::vector<float>    hp(3 * inWindow), hv(vel_xyz ? 3 * inWindow : 0);
    std::vector<uint32_t> hid(inWindow);
    uint32_t              nOwn = 0, n = 0;
    for(uint32_t i = 0; i < n_global; ++i) {
        if(layer[i] < zb - kGhost || layer[i] >= ze + kGhost) continue;
        std::memcpy(&hp[3 * static_cast<size_t>(n)], pos_xyz + 3 * static_cast<size_t>(i), 12);
        if(vel_xyz) std::memcpy(&hv[3 * static_cast<size_t>(n)], vel_xyz + 3 * static_cast<size_t>(i), 12);
        hid[n++] = i;
        nOwn += (layer[i] >= zb && layer[i] < ze) ? 1u : 0u;
    }
    // room for load-balance drift, ghosts and the dead slots of one substep
    // SF_SLAB_TIGHT (tests): start with almost no headroom so that the on-demand growth paths are exercised
    const bool     tight = std::getenv("SF_SLAB_TIGHT") != nullptr;
    const uint32_t want  = tight ? n + 64u : n + n / 2 + (1u << 18);
    int rc = ensure_particle_capacity(s, want);
    if(rc) return rc;
    L.sendCap = L.recvCap = tight ? 256u : std::max<uint32_t>(want / 4, 1u << 16);
    SF_CUDA(s, dev_alloc(L.sendLo, static_cast<size_t>(L.sendCap) * 2));
    SF_CUDA(s, dev_alloc(L.sendHi, static_cast<size_t>(L.sendCap) * 2));
    SF_CUDA(s, dev_alloc(L.recvLo, static_cast<size_t>(L.recvCap) * 2));
    SF_CUDA(s, dev_alloc(L.recvHi, static_cast<size_t>(L.recvCap) * 2));
    SF_CUDA(s, dev_alloc(L.layerStart, static_cast<size_t>(std::max(g[1], g[2])) + 2 * kGhostMax + 2));
    if(n) {
        float*    dpos = s->stage;
        float*    dvel = s->stage + 3 * static_cast<size_t>(s->npad);
        uint32_t* dids = s->B.keys[0]; // scratch until the first substep
        SF_CUDA(s, cudaMemcpyAsync(dpos, hp.data(), static_cast<size_t>(n) * 12, cudaMemcpyHostToDevice, s->stream));
        if(vel_xyz) SF_CUDA(s, cudaMemcpyAsync(dvel, hv.data(), static_cast<size_t>(n) * 12, cudaMemcpyHostToDevice, s->stream));
        SF_CUDA(s, cudaMemcpyAsync(dids, hid.data(), static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, s->stream));
        k_pack_upload_ids<<<cdiv(n, 256), 256, 0, s->stream>>>(dpos, vel_xyz ? dvel : nullptr, dids, s->B.posA, s->B.velA, s->B.idA, n);
    }
    SF_CUDA(s, cudaStreamSynchronize(s->stream));
    s->n = s->nSlots = n;
    L.nGlobal = n_global;
    L.nOwn    = nOwn;
    s->uploaded = true;
    s->ready    = false;
    return SF_OK;
}

int sf_slab_info(sf_solver* s, int32_t* z_begin, int32_t* z_end, uint32_t* n_owned, uint32_t* n_ghost)
try {
    if(!s) return SF_ERR_INVALID;
    if(!s->slab.on || s->slab.cur.empty()) {
        if(z_begin) *z_begin = 0;
        if(z_end) *z_end = s->nS();
        if(n_owned) *n_owned = s->n;
        if(n_ghost) *n_ghost = 0;
        return SF_OK;
    }
    if(z_begin) *z_begin = s->slab.cur[s->slab.rank];
    if(z_end) *z_end = s->slab.cur[s->slab.rank + 1];
    if(n_owned) *n_owned = s->slab.nOwn;
    if(n_ghost) *n_ghost = s->n - std::min(s->n, s->slab.nOwn);
    return SF_OK;
}
SF_NOTHROW(s, "sf_slab_info")

namespace
{
int owned_scratch(sf_solver* s)
{
    if(!s->ownCounters) {
        SF_CUDA(s, cudaMalloc(reinterpret_cast<void**>(&s->ownCounters), 4 * sizeof(uint32_t)));
        SF_CUDA(s, cudaMallocHost(reinterpret_cast<void**>(&s->hostOwnCounters), 4 * sizeof(uint32_t)));
    }
    return SF_OK;
}

// own layer range of this rank by the CURRENT cut planes (global layers along the slow axis)
void owned_range(const sf_solver* s, int& zb, int& ze)
{
    zb = s->slab.on ? s->slab.cur[s->slab.rank] : 0;
    ze = s->slab.on ? s->slab.cur[s->slab.rank + 1] : s->nS();
}

// Compacts the owned particles on the device (k_owned_gather) into the xyz staging arrays + B.keys[0] (free between
// substeps) and copies the first min(count, cap) to the host buffers that are given.  *n_out = owned count.
int gather_owned(sf_solver* s, uint32_t* ids, float* pos_xyz, float* vel_xyz, uint32_t cap, uint32_t* n_out)
{
    int rc = owned_scratch(s);
    if(rc) return rc;
    const uint32_t m = s->slab.on ? s->nSlots : s->n;
    int            zb, ze;
    owned_range(s, zb, ze);
    cudaStream_t st   = s->stream;
    float*       dpos = s->stage;
    float*       dvel = s->stage + 3 * static_cast<size_t>(s->npad);
    SF_CUDA(s, cudaMemsetAsync(s->ownCounters, 0, 4 * sizeof(uint32_t), st));
    if(m) {
        LaunchScope ls(s, K_MARSHAL);
        k_owned_gather<<<std::min<uint32_t>(cdiv(m, 256), s->numSMs * 16), 256, 0, st>>>(s->B.posA, s->B.velA, s->B.idA, m, s->P, zb, ze, dpos, dvel,
                                                                                           s->B.keys[0], s->npad, s->ownCounters);
    }
    SF_CUDA(s, cudaMemcpyAsync(s->hostOwnCounters, s->ownCounters, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SF_CUDA(s, cudaStreamSynchronize(st));
    const uint32_t k = s->hostOwnCounters[2];
    *n_out           = k;
    const uint32_t c = std::min(k, cap);
    if(c && (ids || pos_xyz || vel_xyz)) {
        if(pos_xyz) SF_CUDA(s, cudaMemcpyAsync(pos_xyz, dpos, static_cast<size_t>(c) * 12, cudaMemcpyDeviceToHost, st));
        if(vel_xyz) SF_CUDA(s, cudaMemcpyAsync(vel_xyz, dvel, static_cast<size_t>(c) * 12, cudaMemcpyDeviceToHost, st));
        if(ids) SF_CUDA(s, cudaMemcpyAsync(ids, s->B.keys[0], static_cast<size_t>(c) * 4, cudaMemcpyDeviceToHost, st));
        SF_CUDA(s, cudaStreamSynchronize(st));
    }
    return SF_OK;
}
} // namespace

int sf_download_owned(sf_solver* s, uint32_t* ids, float* pos_xyz, float* vel_xyz, uint32_t cap, uint32_t* n_out)
try {
    if(!s || !n_out) return SF_ERR_INVALID;
    if(!s->uploaded) return fail(s, SF_ERR_INVALID, "no particles uploaded");
    if(!s->ready) return fail(s, SF_ERR_INVALID, "sf_make_ready has not been called");
    SF_CUDA(s, cudaSetDevice(s->device));
    return gather_owned(s, ids, pos_xyz, vel_xyz, cap, n_out);
}
SF_NOTHROW(s, "sf_download_owned")

int sf_slab_axis(sf_solver* s, int32_t* axis_out)
try {
    if(!s || !axis_out) return SF_ERR_INVALID;
    *axis_out = s->axisS;
    return SF_OK;
}
SF_NOTHROW(s, "sf_slab_axis")

// Host-owned slab step: the multi-GPU counterpart of sf_step_host.  Between two calls the host holds this rank's
// OWNED particles {id, position, velocity} (28 B each); the ghost particles of the coming substep stay resident on
// the device -- they arrived with the previous exchange and belong to the neighbours.  One call = upload the m_in
// owned particles (the resident copies are dropped), one substep incl. the halo exchange and migration, download
// the particles this rank owns afterwards (*m_out of them, any order; at most cap are written).
int sf_step_host_owned(sf_solver* s, uint32_t* ids, float* pos_xyz, float* vel_xyz, uint32_t m_in, uint32_t cap, uint32_t* m_out, float* dt_out)
try {
    int rc = require_ready(s);
    if(rc) return rc;
    if(!ids || !pos_xyz || !vel_xyz || !m_out) return SF_ERR_INVALID;
    if(!s->slab.on) return fail(s, SF_ERR_INVALID, "sf_step_host_owned is the slab-mode call: use sf_step_host on a single GPU");
    SF_CUDA(s, cudaSetDevice(s->device));
    rc = owned_scratch(s);
    if(rc) return rc;
    cudaStream_t st = s->stream;
    SF_CUDA(s, cudaStreamSynchronize(st));
    SF_CUDA(s, cudaStreamSynchronize(s->slab.commStream)); // the in-flight all-reduce writes the max |v|^2 slots reset below
    if(static_cast<uint64_t>(s->nSlots) + m_in > s->cap) {
        const uint64_t want = static_cast<uint64_t>(s->nSlots) + m_in;
        if(want + want / 4 > 0xfffffff0ull) return fail(s, SF_ERR_OOM, "more than 2^32 particle slots on one rank");
        SF_CUDA(s, cudaStreamSynchronize(s->slab.commStream));
        rc = ensure_particle_capacity(s, static_cast<uint32_t>(want + want / 4), s->nSlots);
        if(rc) return rc;
        s->slab.regrowths++;
    }
    int zb, ze;
    owned_range(s, zb, ze);
    float*    dpos = s->stage;
    float*    dvel = s->stage + 3 * static_cast<size_t>(s->npad);
    uint32_t* dids = s->B.vals[0]; // free between substeps
    if(m_in) {
        SF_CUDA(s, cudaMemcpyAsync(dpos, pos_xyz, static_cast<size_t>(m_in) * 12, cudaMemcpyHostToDevice, st));
        SF_CUDA(s, cudaMemcpyAsync(dvel, vel_xyz, static_cast<size_t>(m_in) * 12, cudaMemcpyHostToDevice, st));
        SF_CUDA(s, cudaMemcpyAsync(dids, ids, static_cast<size_t>(m_in) * 4, cudaMemcpyHostToDevice, st));
    }
    SF_CUDA(s, cudaMemsetAsync(s->ownCounters, 0, 4 * sizeof(uint32_t), st));
    {
        LaunchScope ls(s, K_MARSHAL);
        // computeMaxVel (A.5) restarts from the uploaded velocities: this rank's slot goes back to FLT_MIN first
        k_owned_reset_maxvel<<<1, 1, 0, st>>>(s->B.state);
        if(s->nSlots) k_owned_kill<<<std::min<uint32_t>(cdiv(s->nSlots, 256), s->numSMs * 16), 256, 0, st>>>(s->B.posA, s->B.idA, s->nSlots, s->P, zb, ze, s->ownCounters);
        if(m_in) k_owned_append<<<std::min<uint32_t>(cdiv(m_in, 256), s->numSMs * 16), 256, 0, st>>>(dpos, dvel, dids, m_in, s->B.posA + s->nSlots, s->B.velA + s->nSlots,
                                                                                                     s->B.idA + s->nSlots, s->P, zb, ze, s->ownCounters, s->B.state);
    }
    SF_CUDA(s, cudaMemcpyAsync(s->hostOwnCounters, s->ownCounters, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SF_CUDA(s, cudaStreamSynchronize(st));
    if(s->hostOwnCounters[1])
        return fail(s, SF_ERR_DOMAIN, std::to_string(s->hostOwnCounters[1]) + " uploaded particles lie outside the box or outside this rank's cell layers");
    const uint32_t killed = s->hostOwnCounters[0];
    s->n      = s->n - std::min(s->n, killed) + m_in;
    s->nSlots = s->nSlots + m_in;
    s->P.n    = s->n;
    s->slab.dtReduced = false; // max |v|^2 changed: all-reduce in-stream before the force pass
    rc = enqueue_substep(s);
    if(rc) return rc;
    rc = read_state(s);
    if(rc) return rc;
    if(dt_out) *dt_out = s->hostState->dt;
    return gather_owned(s, ids, pos_xyz, vel_xyz, cap, m_out);
}
SF_NOTHROW(s, "sf_step_host_owned")

// raw local state (live + dead slots) <-> host: what bench.py's multi-GPU e2e leg moves every substep
int sf_download_local(sf_solver* s, float* pos4, float* vel4, uint32_t* ids, uint32_t cap, uint32_t* n_out)
try {
    if(!s || !n_out) return SF_ERR_INVALID;
    SF_CUDA(s, cudaSetDevice(s->device));
    const uint32_t m = s->slab.on ? s->nSlots : s->n;
    *n_out = m;
    if(!pos4 || !vel4 || !ids) return SF_OK; // count query
    if(m > cap) return fail(s, SF_ERR_INVALID, "buffer too small");
    SF_CUDA(s, cudaMemcpyAsync(pos4, s->B.posA, sizeof(float4) * m, cudaMemcpyDeviceToHost, s->stream));
    SF_CUDA(s, cudaMemcpyAsync(vel4, s->B.velA, sizeof(float4) * m, cudaMemcpyDeviceToHost, s->stream));
    SF_CUDA(s, cudaMemcpyAsync(ids, s->B.idA, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, s->stream));
    SF_CUDA(s, cudaStreamSynchronize(s->stream));
    return SF_OK;
}
SF_NOTHROW(s, "sf_download_local")

int sf_upload_local(sf_solver* s, const float* pos4, const float* vel4, const uint32_t* ids, uint32_t n)
try {
    if(!s || !pos4 || !vel4 || !ids) return SF_ERR_INVALID;
    SF_CUDA(s, cudaSetDevice(s->device));
    const uint32_t m = s->slab.on ? s->nSlots : s->n;
    if(n != m) return fail(s, SF_ERR_INVALID, "sf_upload_local: slot count must match the resident state");
    const int rc = invalidate_binning(s);
    if(rc) return rc;
    SF_CUDA(s, cudaMemcpyAsync(s->B.posA, pos4, sizeof(float4) * m, cudaMemcpyHostToDevice, s->stream));
    SF_CUDA(s, cudaMemcpyAsync(s->B.velA, vel4, sizeof(float4) * m, cudaMemcpyHostToDevice, s->stream));
    SF_CUDA(s, cudaMemcpyAsync(s->B.idA, ids, sizeof(uint32_t) * m, cudaMemcpyHostToDevice, s->stream));
    return SF_OK;
}
SF_NOTHROW(s, "sf_upload_local")

// host-side pieces of the decomposition, exposed for tests (no GPU needed)
int sf_slab_plan(const uint64_t* layer_counts, int32_t nz, int32_t nranks, int32_t* cuts)
try {
    if(!layer_counts || !cuts || nranks < 1 || nz < nranks * kMinThick) return SF_ERR_INVALID;
    slab_plan(layer_counts, nz, nranks, kMinThick, cuts);
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_slab_plan")

int sf_slab_rebalance(const uint32_t* table, int32_t nranks, int32_t nz, int32_t* cuts)
try {
    if(!table || !cuts || nranks < 1) return SF_ERR_INVALID;
    slab_rebalance(table, kRowWords, nranks, nz, kMinThick, cuts);
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_slab_rebalance")

int sf_cell_layers(const sf_params* p, const float* pos_xyz, uint32_t n, int32_t* layers)
try {
    if(!p || (!pos_xyz && n) || !layers) return SF_ERR_INVALID;
    int32_t g[3];
    grid_dims(*p, g);
    for(uint32_t i = 0; i < n; ++i) layers[i] = cell_layer(*p, g[2], pos_xyz[3 * static_cast<size_t>(i) + 2]);
    return SF_OK;
}
SF_NOTHROW(nullptr, "sf_cell_layers")

} // extern "C"
