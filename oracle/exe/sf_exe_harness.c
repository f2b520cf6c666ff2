/* sf_exe_harness.c -- TEST INFRASTRUCTURE (oracle/): runs the REFERENCE'S OWN compiled SPH step natively on Linux.
 *
 * The reference's solver arithmetic is not in its source tree (it lives in the un-vendored Banana library), but its
 * shipped binary /root/reference/Prebuild/SimpleFluid.exe holds the compiled code (x86-64, MSVC, Windows ABI).  That
 * code is plain SSE arithmetic plus a handful of imports, so it can be executed here without Windows:
 *   - the PE image is mapped, unmodified, at its preferred base (0x140000000; no relocations needed);
 *   - the import table is bound to small `ms_abi` shims: libm (sqrtf, floorf, ceilf, fminf, fmaxf, fmax, pow, powf,
 *     log2f), malloc/free/memcpy/memmove/memset, std::_Random_device (returns the seed given on the command line --
 *     the reference seeds its wall-particle jitter from std::random_device), and the eleven TBB entry points, behind
 *     which sits a 60-line single-threaded task scheduler (allocate_root/continuation/child, spawn,
 *     spawn_root_and_wait): parallel_for / parallel_reduce of the binary run their own partitioner code and their
 *     bodies over the whole range, sequentially.  Every other import is bound to a trap that names it and exits;
 *   - gs: points at a fake TEB (stack limit 0 for __chkstk; a TLS block holding `_Init_thread_epoch` = INT_MIN, so the
 *     function-static scratch vectors / constants of computeMaxVel, correctDensity and computeViscosity run their own
 *     thread-safe initialisers on first use; Enter/LeaveCriticalSection, SetEvent/ResetEvent and the atexit
 *     registration are no-op shims, and the CRT's `_Tss_event` startup variable is set non-null).
 * Entry points called: SPHSolver::makeReady EXE@0x140016650 and SPHSolver::advanceFrame
 * EXE@0x140016810 on a solver object laid out as in A.1 (params per Appendix B).  No code byte of the image is patched.
 *
 * usage: sf_exe_harness <SimpleFluid.exe> <in.bin> <out.bin>
 *   in.bin : u32 n, u32 nsteps, u32 seed, u32 flags (1 correctDensity, 2 boundary particles, 4 attractive, 8 velocities given),
 *            f32 h, stiffness, viscosity, restitution, attractiveRatio, restDensity, defaultTimestep, pad; f32 pos[3n]; [f32 vel[3n]]
 *   out.bin: see dump() below.
 * Only tests/ and oracle/Makefile use this; the product never does.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#define MS __attribute__((ms_abi))
#define IMAGE_BASE 0x140000000ull
#define VA(x) ((void*)(uintptr_t)(x))

static void die(const char* msg)
{
    fprintf(stderr, "sf_exe_harness: %s\n", msg);
    exit(2);
}

/* ------------------------------------------------------------------------------------------------ PE loader */
static uint8_t* g_file;
static size_t   g_fileSize;
static uint32_t rd32(size_t o) { uint32_t v; memcpy(&v, g_file + o, 4); return v; }
static uint16_t rd16(size_t o) { uint16_t v; memcpy(&v, g_file + o, 2); return v; }
static uint64_t rd64(size_t o) { uint64_t v; memcpy(&v, g_file + o, 8); return v; }

static uint32_t g_importRva;

static void map_image(const char* path)
{
    FILE* f = fopen(path, "rb");
    if (!f) die("cannot open the EXE");
    fseek(f, 0, SEEK_END);
    g_fileSize = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    g_file = (uint8_t*)malloc(g_fileSize);
    if (fread(g_file, 1, g_fileSize, f) != g_fileSize) die("short read");
    fclose(f);
    const size_t pe = rd32(0x3c);
    if (rd32(pe) != 0x00004550 || rd16(pe + 4) != 0x8664) die("not an x86-64 PE image");
    const int    nsec  = rd16(pe + 6);
    const size_t opt   = pe + 24;
    if (rd16(opt) != 0x20b) die("not PE32+");
    if (rd64(opt + 24) != IMAGE_BASE) die("unexpected image base");
    const uint32_t sizeOfImage = rd32(opt + 56);
    g_importRva = rd32(opt + 112 + 8 * 1);
    void* m = mmap(VA(IMAGE_BASE), sizeOfImage, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS | MAP_FIXED_NOREPLACE, -1, 0);
    if (m != VA(IMAGE_BASE)) die("cannot map the image at its preferred base");
    const size_t sec = opt + rd16(pe + 20);
    for (int i = 0; i < nsec; ++i) {
        const size_t   s = sec + 40 * (size_t)i;
        const uint32_t va = rd32(s + 12), rawSize = rd32(s + 16), rawPtr = rd32(s + 20), vsize = rd32(s + 8);
        const uint32_t n = rawSize < vsize || vsize == 0 ? rawSize : vsize;
        if (rawPtr + (size_t)n > g_fileSize || va + (size_t)n > sizeOfImage) die("bad section");
        memcpy((uint8_t*)VA(IMAGE_BASE) + va, g_file + rawPtr, n);
    }
}

/* ------------------------------------------------------------------------------------------------ import shims */
static uint32_t g_seed;
static MS float  w_sqrtf(float x) { return sqrtf(x); }
static MS float  w_floorf(float x) { return floorf(x); }
static MS float  w_ceilf(float x) { return ceilf(x); }
static MS float  w_fminf(float a, float b) { return fminf(a, b); }
static MS float  w_fmaxf(float a, float b) { return fmaxf(a, b); }
static MS double w_fmax(double a, double b) { return fmax(a, b); }
static MS double w_pow(double a, double b) { return pow(a, b); }
static MS float  w_powf(float a, float b) { return powf(a, b); }
static MS float  w_log2f(float a) { return log2f(a); }
static MS void*  w_malloc(size_t n) { return malloc(n ? n : 1); }
static MS void*  w_calloc(size_t a, size_t b) { return calloc(a ? a : 1, b ? b : 1); }
static MS void   w_free(void* p) { free(p); }
static MS int    w_callnewh(size_t n) { (void)n; return 0; }
static MS void*  w_memcpy(void* d, const void* s, size_t n) { return memcpy(d, s, n); }
static MS void*  w_memmove(void* d, const void* s, size_t n) { return memmove(d, s, n); }
static MS void*  w_memset(void* d, int c, size_t n) { return memset(d, c, n); }
static MS unsigned w_random_device(void) { return g_seed; }
static MS void   w_nop(void) {}
static MS int    w_ret0(void) { return 0; }
static MS int    w_ret1(void) { return 1; }

/* ---- TBB at its import boundary: a single-threaded scheduler ---------------------------------------------
 * task_prefix (64 bytes in front of every task), offsets relative to the task: -0x38 context, -0x30 origin,
 * -0x28 owner (scheduler*), -0x20 parent, -0x18 ref_count, -0x10 depth, -0xc state, -0xb extra_state, -0xa affinity,
 * -0x8 next.  The binary reaches the scheduler only through owner->vtable: slot 0 spawn(first, next), slot 2
 * spawn_root_and_wait(first, next) (EXE@0x14001c3e8, EXE@0x140017982). */
typedef struct { void** vtbl; } FakeSched;
static FakeSched g_sched;
static void**    g_queue;
static size_t    g_qn, g_qcap;
static uint64_t  g_tasksRun;

#define PFX(t, off, T) (*(T*)((uint8_t*)(t) + (off)))
static void* task_alloc(size_t bytes, void* context, void* parent)
{
    uint8_t* base = (uint8_t*)calloc(1, 0x40 + bytes + 16);
    void*    t    = base + 0x40;
    PFX(t, -0x38, void*) = context;
    PFX(t, -0x28, void*) = &g_sched;
    PFX(t, -0x20, void*) = parent;
    PFX(t, -0xc, uint8_t) = 3; /* allocated */
    return t;
}
static void queue_push(void* t)
{
    if (g_qn == g_qcap) {
        g_qcap  = g_qcap ? 2 * g_qcap : 64;
        g_queue = (void**)realloc(g_queue, g_qcap * sizeof(void*));
    }
    g_queue[g_qn++] = t;
}
static MS void* tbb_allocate_root(void* proxy, size_t bytes) { return task_alloc(bytes, *(void**)proxy, NULL); }
static MS void* tbb_allocate_continuation(void* self_task, size_t bytes)
{
    void* c = task_alloc(bytes, PFX(self_task, -0x38, void*), PFX(self_task, -0x20, void*));
    PFX(self_task, -0x20, void*) = NULL;
    return c;
}
static MS void* tbb_allocate_child(void* parent, size_t bytes) { return task_alloc(bytes, PFX(parent, -0x38, void*), parent); }
static MS void  tbb_free_root(void* proxy, void* task) { (void)proxy; free((uint8_t*)task - 0x40); }
static MS size_t tbb_initial_divisor(void) { return 4; } /* one "thread": divisor 4 * 1 */
static MS void  tbb_context_init(void* ctx)
{   /* keeps my_kind (+0) and my_version_and_traits (+0x80), which the caller sets before init() */
    memset((uint8_t*)ctx + 8, 0, 0x78);
    memset((uint8_t*)ctx + 0x88, 0, 0x78);
}
static MS void tbb_context_dtor(void* ctx) { (void)ctx; }
static MS int  tbb_is_cancelled(void* ctx) { (void)ctx; return 0; }
static MS void tbb_note_affinity(void* task, unsigned short id) { (void)task; (void)id; }

typedef MS void* (*ExecuteFn)(void* task);
static void run_task(void* t)
{
    while (t) {
        PFX(t, -0xc, uint8_t) = 0; /* executing */
        void** vt   = *(void***)t;
        void*  next = ((ExecuteFn)vt[1])(t);
        ++g_tasksRun;
        if (PFX(t, -0xc, uint8_t) != 0) die("a task recycled itself: not supported by the harness scheduler");
        void* parent = PFX(t, -0x20, void*);
        free((uint8_t*)t - 0x40);
        if (parent && --PFX(parent, -0x18, int64_t) == 0) queue_push(parent); /* continuation is ready */
        t = next;
    }
}
static MS void sched_spawn(FakeSched* s, void* first, void** next)
{
    (void)s;
    if (next != &PFX(first, -0x8, void*)) die("spawn of a task list: not supported");
    queue_push(first);
}
static MS void sched_spawn_root_and_wait(FakeSched* s, void* first, void** next)
{
    (void)s;
    if (next != &PFX(first, -0x8, void*)) die("spawn_root_and_wait of a task list: not supported");
    const size_t mark = g_qn;
    run_task(first);
    while (g_qn > mark) run_task(g_queue[--g_qn]);
}
static MS void sched_unexpected(void) { die("unexpected scheduler virtual call"); }
static void* g_schedVtbl[16];

/* ---- traps for every other import ------------------------------------------------------------------------- */
static char** g_importNames;
static MS void trap(uint64_t idx)
{
    fprintf(stderr, "sf_exe_harness: the reference code called an import the harness does not provide: %s\n", g_importNames[idx]);
    fflush(stderr);
    _exit(3);
}
static uint8_t* g_stubPage;
static size_t   g_stubUsed;
static void* make_trap_stub(uint64_t idx)
{
    if (!g_stubPage) g_stubPage = (uint8_t*)mmap(NULL, 1 << 20, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    uint8_t* p = g_stubPage + g_stubUsed;
    uint64_t h = (uint64_t)(uintptr_t)&trap;
    p[0] = 0x48; p[1] = 0xB9; memcpy(p + 2, &idx, 8);  /* mov rcx, idx */
    p[10] = 0x49; p[11] = 0xBB; memcpy(p + 12, &h, 8); /* mov r11, trap */
    p[20] = 0x41; p[21] = 0xFF; p[22] = 0xE3;          /* jmp r11 */
    g_stubUsed += 32;
    if (g_stubUsed + 32 > (1 << 20)) die("too many imports");
    return p;
}

static const struct { const char* name; void* fn; } kShims[] = {
    { "sqrtf", (void*)w_sqrtf }, { "floorf", (void*)w_floorf }, { "ceilf", (void*)w_ceilf }, { "fminf", (void*)w_fminf },
    { "fmaxf", (void*)w_fmaxf }, { "fmax", (void*)w_fmax }, { "pow", (void*)w_pow }, { "powf", (void*)w_powf }, { "log2f", (void*)w_log2f },
    { "malloc", (void*)w_malloc }, { "calloc", (void*)w_calloc }, { "free", (void*)w_free }, { "_callnewh", (void*)w_callnewh },
    { "memcpy", (void*)w_memcpy }, { "memmove", (void*)w_memmove }, { "memset", (void*)w_memset },
    { "?_Random_device@std@@YAIXZ", (void*)w_random_device },
    { "?allocate@allocate_root_with_context_proxy@internal@tbb@@QEBAAEAVtask@3@_K@Z", (void*)tbb_allocate_root },
    { "?allocate@allocate_continuation_proxy@internal@tbb@@QEBAAEAVtask@3@_K@Z", (void*)tbb_allocate_continuation },
    { "?allocate@allocate_child_proxy@internal@tbb@@QEBAAEAVtask@3@_K@Z", (void*)tbb_allocate_child },
    { "?free@allocate_root_with_context_proxy@internal@tbb@@QEBAXAEAVtask@3@@Z", (void*)tbb_free_root },
    { "?get_initial_auto_partitioner_divisor@internal@tbb@@YA_KXZ", (void*)tbb_initial_divisor },
    { "?init@task_group_context@tbb@@IEAAXXZ", (void*)tbb_context_init },
    { "??1task_group_context@tbb@@QEAA@XZ", (void*)tbb_context_dtor },
    { "?is_group_execution_cancelled@task_group_context@tbb@@QEBA_NXZ", (void*)tbb_is_cancelled },
    { "?note_affinity@task@tbb@@UEAAXG@Z", (void*)tbb_note_affinity },
    /* thread-safe static initialisation (_Init_thread_header / _Init_thread_footer / atexit of the static's dtor) */
    { "EnterCriticalSection", (void*)w_nop }, { "LeaveCriticalSection", (void*)w_nop }, { "SetEvent", (void*)w_ret1 },
    { "ResetEvent", (void*)w_ret1 }, { "_register_onexit_function", (void*)w_ret0 }, { "_crt_atexit", (void*)w_ret0 },
};

static void bind_imports(void)
{
    uint8_t* img = (uint8_t*)VA(IMAGE_BASE);
    size_t   count = 0, cap = 4096;
    g_importNames = (char**)calloc(cap, sizeof(char*));
    for (uint8_t* d = img + g_importRva;; d += 20) {
        uint32_t ilt, nameRva, iat;
        memcpy(&ilt, d, 4);
        memcpy(&nameRva, d + 12, 4);
        memcpy(&iat, d + 16, 4);
        if (!ilt && !iat) break;
        const char* dll = (const char*)img + nameRva;
        uint64_t*   lookup = (uint64_t*)(img + (ilt ? ilt : iat));
        uint64_t*   slot   = (uint64_t*)(img + iat);
        for (; *lookup; ++lookup, ++slot) {
            char label[512];
            void* fn = NULL;
            if (*lookup >> 63) {
                snprintf(label, sizeof(label), "%s!#%u", dll, (unsigned)(*lookup & 0xffff));
            } else {
                const char* name = (const char*)img + (uint32_t)*lookup + 2;
                snprintf(label, sizeof(label), "%s!%s", dll, name);
                for (size_t k = 0; k < sizeof(kShims) / sizeof(kShims[0]); ++k)
                    if (!strcmp(kShims[k].name, name)) fn = kShims[k].fn;
            }
            if (count == cap) die("import name table overflow");
            g_importNames[count] = strdup(label);
            *slot = (uint64_t)(uintptr_t)(fn ? fn : make_trap_stub(count));
            ++count;
        }
    }
}

static void fake_teb(void)
{
    static uint8_t  teb[4096] __attribute__((aligned(64)));
    static uint64_t tlsArray[64];
    static int32_t  tlsBlock[64];
    for (int i = 0; i < 64; ++i) {
        tlsBlock[i] = INT32_MIN; /* _Init_thread_epoch starts at INT_MIN (the .tls template): function statics initialise on first use */
        tlsArray[i] = (uint64_t)(uintptr_t)tlsBlock;
    }
    *(uint64_t*)(teb + 0x10) = 0;                               /* StackLimit: __chkstk never probes */
    *(uint64_t*)(teb + 0x30) = (uint64_t)(uintptr_t)teb;        /* NtCurrentTeb() */
    *(uint64_t*)(teb + 0x58) = (uint64_t)(uintptr_t)tlsArray;   /* ThreadLocalStoragePointer */
    if (syscall(SYS_arch_prctl, 0x1001 /* ARCH_SET_GS */, teb) != 0) die("arch_prctl(ARCH_SET_GS) failed");
    /* CRT startup state (the CRT's own startup, which never runs here, creates this event): with a non-null handle
     * _Init_thread_notify (EXE@0x14003e8d0) signals through SetEvent / ResetEvent -- shims above -- instead of through
     * an encoded WakeAllConditionVariable pointer that only the startup code could have filled in */
    *(uint64_t*)VA(0x14007ef50) = 1;
}

/* ------------------------------------------------------------------------------------------------ the solver object */
/* SimulationParameters, SURVEY.md Appendix B (binary layout EXE@0x140011db0) */
typedef struct {
    int32_t scene, numThreads;
    float   stopTime, defaultTimestep, boxMin[3], boxMax[3], pressureStiffness, stiffness2, viscosity, kernelRadius;
    uint8_t bCorrectDensity, bUseBoundaryParticles, bUseAttractivePressure, pad0;
    float   boundaryRestitution, attractivePressureRatio, restDensity, particleMass, particleRadius, kernelRadiusSqr, r25, restDensitySqr;
} ExeParams;
typedef struct { uint8_t *begin, *end, *cap; } MsvcVector; /* release-mode std::vector<T> */

#define SOLVER_BYTES 0x3ab48 /* EXE@0x140014d6b */
typedef MS void  (*MakeReadyFn)(void* solver);
typedef MS float (*AdvanceFrameFn)(void* solver);
typedef MS void  (*UpdateParamsFn)(void* params);

static void put(FILE* f, const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) die("short write"); }

int main(int argc, char** argv)
{
    if (argc != 4) die("usage: sf_exe_harness <SimpleFluid.exe> <in.bin> <out.bin>");
    FILE* fi = fopen(argv[2], "rb");
    if (!fi) die("cannot open the input");
    uint32_t hdr[4];
    float    par[8];
    if (fread(hdr, 4, 4, fi) != 4 || fread(par, 4, 8, fi) != 8) die("short input header");
    const uint32_t n = hdr[0], nsteps = hdr[1], flags = hdr[3];
    g_seed = hdr[2];
    float* pos0 = (float*)malloc((size_t)n * 12 + 16);
    float* vel0 = (float*)calloc((size_t)n * 3 + 4, 4);
    if (fread(pos0, 12, n, fi) != n) die("short input positions");
    if ((flags & 8u) && fread(vel0, 12, n, fi) != n) die("short input velocities");
    fclose(fi);

    map_image(argv[1]);
    g_schedVtbl[0] = (void*)sched_spawn;
    g_schedVtbl[1] = (void*)sched_unexpected;
    g_schedVtbl[2] = (void*)sched_spawn_root_and_wait;
    for (int i = 3; i < 16; ++i) g_schedVtbl[i] = (void*)sched_unexpected;
    g_sched.vtbl = g_schedVtbl;
    bind_imports();
    fake_teb();

    /* parameters: the fields Controller.cpp:54-63 sets, then the binary's OWN derived block.  updateParams() is inlined
     * in the binary's constructor/GUI code, so the derived fields are restated here (Appendix B) -- their values are part
     * of the dump and the tests compare them with the oracle's. */
    ExeParams* P = (ExeParams*)calloc(1, sizeof(ExeParams) + 64);
    P->scene = 2;
    P->stopTime = 5.0f;
    P->defaultTimestep = par[6];
    for (int d = 0; d < 3; ++d) { P->boxMin[d] = -1.0f; P->boxMax[d] = 1.0f; }
    P->pressureStiffness = P->stiffness2 = par[1];
    P->viscosity = par[2];
    P->kernelRadius = par[0];
    P->bCorrectDensity = (flags & 1u) ? 1 : 0;
    P->bUseBoundaryParticles = (flags & 2u) ? 1 : 0;
    P->bUseAttractivePressure = (flags & 4u) ? 1 : 0;
    P->boundaryRestitution = par[3];
    P->attractivePressureRatio = par[4];
    P->restDensity = par[5];
    {
        const float h = P->kernelRadius, r = h * 0.25f;
        P->particleRadius  = r;
        P->kernelRadiusSqr = h * h;
        P->r25             = r * 2.5f;
        P->particleMass    = (float)(pow((double)r + (double)r, 3.0) * (double)P->restDensity * 0.9);
        P->restDensitySqr  = P->restDensity * P->restDensity;
    }

    uint8_t* S = (uint8_t*)calloc(1, SOLVER_BYTES + 256);
    MsvcVector* positions = (MsvcVector*)calloc(1, sizeof(MsvcVector)); /* ParticleSystemData "Position" array, aliased at +0x78 */
    positions->begin = (uint8_t*)pos0;
    positions->end = positions->cap = (uint8_t*)pos0 + (size_t)n * 12;
    *(void**)(S + 0x08) = P;
    *(void**)(S + 0x78) = positions;
    /* SceneManager fills both arrays (Source/SceneManager.cpp:50-63: velocity.assign(n, 0)); makeReady itself only
     * "resizes" the velocity array to its own size (EXE@0x140016762-0x1400167a0) */
    MsvcVector* velocity = (MsvcVector*)(S + 0x80);
    velocity->begin = (uint8_t*)vel0;
    velocity->end = velocity->cap = (uint8_t*)vel0 + (size_t)n * 12;

    ((MakeReadyFn)VA(0x140016650))(S);
    if ((size_t)(velocity->end - velocity->begin) != (size_t)n * 12) die("makeReady changed the size of the velocity array");

    FILE* fo = fopen(argv[3], "wb");
    if (!fo) die("cannot open the output");
    /* ---- dump: header */
    const uint64_t dims[3] = { *(uint64_t*)(S + 0x18), *(uint64_t*)(S + 0x20), *(uint64_t*)(S + 0x28) };
    uint32_t out[16] = { 0x45584553u /* "SEXE" */, n, nsteps, (uint32_t)dims[0], (uint32_t)dims[1], (uint32_t)dims[2] };
    MsvcVector* walls = (MsvcVector*)(S + 0x98);
    for (int w = 0; w < 6; ++w) out[6 + w] = (uint32_t)((walls[w].end - walls[w].begin) / 12);
    put(fo, out, sizeof(out));
    put(fo, P, 0x5c);
    /* kernel objects (A.1): +0 h, k, l, W(0); +0x10 W[10000]; +0x9c50 gradW[10001]; +0x13894 radius, radius2, invStep, W_zero */
    const uint8_t* cubic = S + 0x158;
    const uint8_t* spiky = S + 0x139fc;
    put(fo, cubic, 16);
    put(fo, cubic + 0x13894, 16);
    put(fo, cubic + 0x10, 40000);
    put(fo, spiky, 16);
    put(fo, spiky + 0x13894, 16);
    put(fo, spiky + 0x9c50, 40004);
    for (int w = 0; w < 6; ++w) put(fo, walls[w].begin, (size_t)out[6 + w] * 12);

    MsvcVector* cells   = (MsvcVector*)(S + 0x30);
    MsvcVector* accel   = (MsvcVector*)(S + 0x128);
    MsvcVector* density = (MsvcVector*)(S + 0x140);
    uint32_t*   cellOf  = (uint32_t*)malloc((size_t)n * 4 + 4);
    for (uint32_t k = 0; k < nsteps; ++k) {
        const float dt = ((AdvanceFrameFn)VA(0x140016810))(S);
        /* cell index of every particle from the binary's own cell lists; list order must be ascending particle id */
        const MsvcVector* cl = (const MsvcVector*)cells->begin;
        const size_t      nc = (size_t)(cells->end - cells->begin) / sizeof(MsvcVector);
        uint32_t          listed = 0, ordered = 1;
        memset(cellOf, 0xff, (size_t)n * 4);
        for (size_t c = 0; c < nc; ++c) {
            const uint32_t* ids = (const uint32_t*)cl[c].begin;
            const size_t    m   = (size_t)(cl[c].end - cl[c].begin) / 4;
            for (size_t i = 0; i < m; ++i) {
                if (ids[i] >= n) die("cell list holds an invalid particle id");
                if (i && ids[i] <= ids[i - 1]) ordered = 0;
                cellOf[ids[i]] = (uint32_t)c;
                ++listed;
            }
        }
        uint32_t rec[4] = { 0, listed, ordered, (uint32_t)nc };
        memcpy(&rec[0], &dt, 4);
        put(fo, rec, sizeof(rec));
        put(fo, cellOf, (size_t)n * 4);
        put(fo, density->begin, (size_t)n * 4);
        put(fo, accel->begin, (size_t)n * 12);
        put(fo, positions->begin, (size_t)n * 12);
        put(fo, velocity->begin, (size_t)n * 12);
    }
    fclose(fo);
    fprintf(stderr, "sf_exe_harness: n=%u steps=%u grid=%llux%llux%llu walls=%u tasks=%llu ok\n", n, nsteps, (unsigned long long)dims[0],
            (unsigned long long)dims[1], (unsigned long long)dims[2], out[6], (unsigned long long)g_tasksRun);
    return 0;
}
