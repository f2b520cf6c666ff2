/*
 * sf_oracle.c -- CPU oracle for the SimpleFluid SPH step.  TEST INFRASTRUCTURE ONLY
 * (see sf_oracle.h for scope, provenance and how it is pinned to the reference binary's own outputs).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp   (the MSVC reference binary uses
 * separately rounded SSE scalar ops, no FMA -- SURVEY.md section 0 item 5).
 *
 * Every function cites the place it follows: `path:line` under /root/reference, or
 * EXE@0x14001xxxx = virtual address in /root/reference/Prebuild/SimpleFluid.exe as listed in
 * SURVEY.md Appendix A.
 */
#include "sf_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SFO_TAB 10000
#define PI_F 3.14159274f

typedef struct {
    float h, k, l;
    float tab[2 * SFO_TAB + 1]; /* W[0..9999] then gradW[0..10000]; W[10000] aliases gradW[0] (A.2) */
    float radius, radius2, invStep, Wzero;
} sfo_kernel;

struct sfo_solver {
    sfo_params P;
    uint32_t n;
    float *pos, *vel, *acc, *rho, *rho2, *visc, *m2;
    int32_t nc[3];
    float gmin[3], gmax[3], cell;
    uint64_t ncells;
    uint32_t *cellStart; /* ncells+1 */
    uint32_t *cellIds;   /* n, ascending particle id inside each cell = push_back order of A.7 */
    uint32_t *cellOf;    /* n */
    float *bnd[6];       /* LX UX LY UY LZ UZ */
    uint32_t nbnd[6];
    sfo_kernel cubic, spiky;
    int nthreads, reversed, ready;
    double timing[6];
};

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---------------------------------------------------------------- parameters (Appendix B) */
void sfo_params_update(sfo_params* p)
{ /* EXE@0x140006ac6-0x140006b2c */
    float h = p->kernelRadius;
    float r = h * 0.25f;
    p->particleRadius = r;
    p->kernelRadiusSqr = h * h;
    p->particleMass = (float)(pow((double)r + (double)r, 3.0) * (double)p->restDensity * 0.9);
    p->restDensitySqr = p->restDensity * p->restDensity;
}

void sfo_params_default(sfo_params* p)
{ /* EXE@0x140011db0; GUI defaults Captured/1.png */
    memset(p, 0, sizeof(*p));
    p->scene = 0;
    p->numThreads = 0;
    p->stopTime = 5.0f;
    p->defaultTimestep = 1.0e-4f;
    for (int d = 0; d < 3; ++d) { p->boxMin[d] = -1.0f; p->boxMax[d] = 1.0f; }
    p->pressureStiffness = 50000.0f;
    p->viscosity = 0.05f;
    p->kernelRadius = 2.0f / 24.0f;
    p->bCorrectDensity = 0;
    p->bUseBoundaryParticles = 1;
    p->bUseAttractivePressure = 0;
    p->boundaryRestitution = 0.1f;
    p->attractivePressureRatio = 0.1f;
    p->restDensity = 1000.0f;
    sfo_params_update(p);
}

void sfo_params_set_resolution(sfo_params* p, float resolution)
{ /* Source/Controller.cpp:55,63 */
    p->kernelRadius = 2.0f / resolution;
    sfo_params_update(p);
}

/* ---------------------------------------------------------------- scenes */
typedef struct { float* out; uint64_t cap, n; } sfo_sink;
static void sink_push(sfo_sink* s, float x, float y, float z)
{
    if (s->out && s->n < s->cap) { s->out[3 * s->n] = x; s->out[3 * s->n + 1] = y; s->out[3 * s->n + 2] = z; }
    s->n++;
}

/* lattice block: Source/SceneManager.cpp:46-62 (count = float->int truncation, i->j->k order).
 * sign=+1: ppos = origin + spacing*(i,j,k) (:57,:112,:144); sign=-1: ppos = origin - spacing*(i,j,k) (:165) */
static void scene_block(sfo_sink* s, const float bmin[3], const float bmax[3], float spacing, int sign, int sphere)
{
    int g[3];
    for (int d = 0; d < 3; ++d) g[d] = (int)((bmax[d] - bmin[d]) / spacing);
    const float* o = (sign > 0) ? bmin : bmax;
    for (int i = 0; i < g[0]; ++i)
        for (int j = 0; j < g[1]; ++j)
            for (int k = 0; k < g[2]; ++k) {
                float ox = spacing * (float)i, oy = spacing * (float)j, oz = spacing * (float)k;
                float x = (sign > 0) ? o[0] + ox : o[0] - ox;
                float y = (sign > 0) ? o[1] + oy : o[1] - oy;
                float z = (sign > 0) ? o[2] + oz : o[2] - oz;
                if (sphere) { /* :83  glm::length(ppos - center) > radius, center = 0, radius = 0.5 */
                    float len = sqrtf((x * x + y * y) + z * z);
                    if (len > 0.5f) continue;
                }
                sink_push(s, x, y, z);
            }
}

uint64_t sfo_scene_generate(const sfo_params* p, int scene, float* pos_xyz, uint64_t cap)
{
    sfo_sink s = { pos_xyz, cap, 0 };
    float r = p->particleRadius;
    float spacing = 2.0f * r;
    if (scene == 1) { /* CubeDrop :41-64 */
        float a[3] = { -0.5f, -0.5f, -0.5f }, b[3] = { 0.5f, 0.5f, 0.5f };
        scene_block(&s, a, b, spacing, +1, 0);
    } else if (scene == 0) { /* SphereDrop :67-91; grid = int(2*radius/spacing) on every axis (:73) */
        float a[3] = { 0.0f - 0.5f, 0.0f - 0.5f, 0.0f - 0.5f };
        int g = (int)(2.0f * 0.5f / spacing);
        /* scene_block derives the count from (bmax-bmin)/spacing; here the count comes from :73, so loop directly */
        for (int i = 0; i < g; ++i)
            for (int j = 0; j < g; ++j)
                for (int k = 0; k < g; ++k) {
                    float x = a[0] + spacing * (float)i, y = a[1] + spacing * (float)j, z = a[2] + spacing * (float)k;
                    float dx = x - 0.0f, dy = y - 0.0f, dz = z - 0.0f;
                    float len = sqrtf((dx * dx + dy * dy) + dz * dz);
                    if (len > 0.5f) continue;
                    sink_push(&s, x, y, z);
                }
    } else if (scene == 2 || scene == 3) { /* Dambreak :94-119, first block of DoubleDambreak :133-149 */
        float a[3] = { -1.0f + r, -1.0f + r, -1.0f + r }, b[3] = { 0.4f, 0.4f, -0.5f };
        scene_block(&s, a, b, spacing, +1, 0);
        if (scene == 3) { /* second block :152-169, filled downward from bMax */
            float a2[3] = { -0.4f + 0.0f, -1.0f + r, 0.5f + 0.0f };
            float b2[3] = { 1.0f - r, 0.4f - 0.0f, 1.0f - r };
            scene_block(&s, a2, b2, spacing, -1, 0);
        }
    }
    return s.n;
}

/* ---------------------------------------------------------------- kernel tables (A.2) */
static float cubic_W(const sfo_kernel* K, float r)
{ /* EXE@0x14001aa60 */
    float q = r / K->h;
    if (!(1.0f >= q)) return 0.0f;
    if (0.5f >= q) {
        float q2 = q * q, q3 = q2 * q;
        return (float)((((double)q3 * 6.0 - (double)q2 * 6.0) + 1.0) * (double)K->k);
    }
    return (float)((2.0 * pow(1.0 - (double)q, 3.0)) * (double)K->k);
}

static void kernel_common(sfo_kernel* K, float h, int spiky)
{ /* setRadius EXE@0x14001a4e0 (cubic) / 0x14001a2d0 (spiky) */
    K->h = h;
    if (!spiky) {
        float h3 = (h * h) * h;
        K->k = (float)(8.0 / (double)(h3 * PI_F));
        K->l = (float)(48.0 / (double)(h3 * PI_F));
    } else { /* EXE@0x14001a850 */
        float d = powf(h, 6.0f) * PI_F;
        K->k = (float)(15.0 / (double)d);
        K->l = (float)(-45.0 / (double)d);
    }
    K->radius = h;
    K->radius2 = h * h;
    float step = h / 10000.0f;
    K->invStep = (float)(1.0 / (double)step);
    float* W = K->tab;
    float* G = K->tab + SFO_TAB;
    for (int i = 0; i < SFO_TAB; ++i) {
        float X = (float)i * step;
        float w, g = 0.0f;
        if (!spiky) {
            w = cubic_W(K, X);
            /* the cubic gradW table is built by the reference but never read by the step
             * (A.2); only gradW[0] = 0 matters because W[10000] aliases it. */
        } else {
            w = (h * h >= X * X) ? powf(h - sqrtf(X * X), 3.0f) * K->k : 0.0f; /* unused by the step */
            if ((double)X > 1e-6) {
                float r2 = X * X + 0.0f;
                float gx;
                if (h * h >= r2) {
                    float rl = sqrtf(r2);
                    gx = ((((h - rl) * (h - rl)) * K->l) * X) * (1.0f / rl);
                } else gx = 0.0f;
                g = gx / X;
            }
        }
        W[i] = w;
        G[i] = g;
    }
    G[0] = 0.0f;
    G[SFO_TAB] = 0.0f;
    {
        uint32_t i0 = (uint32_t)(int64_t)(K->invStep * 0.0f);
        if (i0 > SFO_TAB) i0 = SFO_TAB;
        K->Wzero = W[i0];
    }
}

/* table lookup used everywhere in the step (A.2): tab[min((uint32)(int64)(sqrtf(d2)*invStep), 10000)] */
static inline uint32_t tab_index(const sfo_kernel* K, float d2)
{
    uint32_t i = (uint32_t)(int64_t)(sqrtf(d2) * K->invStep);
    return i > SFO_TAB ? SFO_TAB : i;
}
static inline float W_lookup(const sfo_kernel* K, float d2)
{ /* guarded by radius2 >= d2, else contributes +0 */
    if (K->radius2 >= d2) return K->tab[tab_index(K, d2)];
    return 0.0f;
}
static inline float G_lookup(const sfo_kernel* K, float d2)
{
    if (K->radius2 >= d2) return K->tab[SFO_TAB + tab_index(K, d2)];
    return 0.0f;
}

/* ---------------------------------------------------------------- lifecycle */
sfo_solver* sfo_create(const sfo_params* p)
{
    sfo_solver* s = (sfo_solver*)calloc(1, sizeof(sfo_solver));
    s->P = *p;
    return s;
}

void sfo_destroy(sfo_solver* s)
{
    if (!s) return;
    free(s->pos); free(s->vel); free(s->acc); free(s->rho); free(s->rho2); free(s->visc); free(s->m2);
    free(s->cellStart); free(s->cellIds); free(s->cellOf);
    for (int w = 0; w < 6; ++w) free(s->bnd[w]);
    free(s);
}

void sfo_set_threads(sfo_solver* s, int nthreads) { s->nthreads = nthreads; }
void sfo_set_traversal(sfo_solver* s, int reversed) { s->reversed = reversed; }

int sfo_set_particles(sfo_solver* s, const float* pos_xyz, const float* vel_xyz, uint32_t n)
{
    free(s->pos); free(s->vel);
    s->pos = (float*)malloc((size_t)n * 12 + 16);
    s->vel = (float*)calloc((size_t)n * 3 + 4, 4);
    memcpy(s->pos, pos_xyz, (size_t)n * 12);
    if (vel_xyz) memcpy(s->vel, vel_xyz, (size_t)n * 12);
    s->n = n;
    s->ready = 0;
    return 0;
}

/* std::mt19937 (the reference seeds it from std::random_device, i.e. is itself not
 * reproducible; we seed explicitly -- BASELINE.md section 2) */
typedef struct { uint32_t mt[624]; int idx; } mt19937;
static void mt_seed(mt19937* g, uint32_t seed)
{
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}
static uint32_t mt_next(mt19937* g)
{
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
/* MSVC std::generate_canonical<float,24> over mt19937, EXE@0x1400101a0: k = max(1, ceil(24 / log2(2^32))) = 1 draw;
 * ans = 0 + (float)(uint32 draw) * 1.0f (cvtsi2ss of the zero-extended draw: round to nearest);
 * return ans / (1.0f * 4294967296.0f).  May return exactly 1.0f. */
static float mt_canon(mt19937* g) { return (float)mt_next(g) / 4294967296.0f; }

void sfo_generate_boundary(sfo_solver* s, uint32_t seed)
{ /* EXE@0x140016d80 (A.4); roles of the three draws per wall: EXE@0x14001705a-0x1400172e0 */
    const sfo_params* P = &s->P;
    mt19937 gen;
    mt_seed(&gen, seed);
    float r = P->particleRadius, h = P->kernelRadius;
    float lo = (float)((double)r * 0.1), hi = (float)((double)r * 0.3);
    float sp = r * 1.7f;
    int nA = (int)ceilf(h * 3.0f / sp) + 1;
    int nB = (int)ceilf(h / sp);
    float base = r - h;
    uint32_t cnt = (uint32_t)(nA * nA * nB);
    for (int w = 0; w < 6; ++w) {
        free(s->bnd[w]);
        s->bnd[w] = (float*)malloc((size_t)cnt * 12);
        s->nbnd[w] = 0;
    }
    for (int i = 0; i < nA; ++i)
        for (int j = 0; j < nA; ++j)
            for (int k = 0; k < nB; ++k) {
                float ti = base + (float)i * sp;   /* xmm15: outer loop */
                float tj = base + (float)j * sp;   /* xmm11: middle loop */
                float depth = (float)k * sp + r;   /* xmm8 */
                float u[3], *o;
                /* LX 0x14001705a: jit, jit, depth-jit */
                u[0] = mt_canon(&gen) * (hi - lo) + lo; u[1] = mt_canon(&gen) * (hi - lo) + lo; u[2] = mt_canon(&gen) * (lo - 0.0f) + 0.0f;
                o = s->bnd[0] + 3 * (size_t)s->nbnd[0]++;
                o[0] = (P->boxMin[0] - depth) + u[2]; o[1] = u[1] + ti; o[2] = u[0] + tj;
                /* UX 0x1400170c7 */
                u[0] = mt_canon(&gen) * (hi - lo) + lo; u[1] = mt_canon(&gen) * (hi - lo) + lo; u[2] = mt_canon(&gen) * (lo - 0.0f) + 0.0f;
                o = s->bnd[1] + 3 * (size_t)s->nbnd[1]++;
                o[0] = u[2] + (depth + P->boxMax[0]); o[1] = u[1] + ti; o[2] = u[0] + tj;
                /* LY 0x140017133: jit, depth-jit, jit */
                u[0] = mt_canon(&gen) * (hi - lo) + lo; u[1] = mt_canon(&gen) * (lo - 0.0f) + 0.0f; u[2] = mt_canon(&gen) * (hi - lo) + lo;
                o = s->bnd[2] + 3 * (size_t)s->nbnd[2]++;
                o[0] = u[2] + ti; o[1] = (P->boxMin[1] - depth) + u[1]; o[2] = u[0] + tj;
                /* UY 0x1400171a0 */
                u[0] = mt_canon(&gen) * (hi - lo) + lo; u[1] = mt_canon(&gen) * (lo - 0.0f) + 0.0f; u[2] = mt_canon(&gen) * (hi - lo) + lo;
                o = s->bnd[3] + 3 * (size_t)s->nbnd[3]++;
                o[0] = u[2] + ti; o[1] = (depth + P->boxMax[1]) + u[1]; o[2] = u[0] + tj;
                /* LZ 0x14001720c: depth-jit, jit, jit */
                u[0] = mt_canon(&gen) * (lo - 0.0f) + 0.0f; u[1] = mt_canon(&gen) * (hi - lo) + lo; u[2] = mt_canon(&gen) * (hi - lo) + lo;
                o = s->bnd[4] + 3 * (size_t)s->nbnd[4]++;
                o[0] = u[2] + ti; o[1] = u[1] + tj; o[2] = (P->boxMin[2] - depth) + u[0];
                /* UZ 0x140017278 */
                u[0] = mt_canon(&gen) * (lo - 0.0f) + 0.0f; u[1] = mt_canon(&gen) * (hi - lo) + lo; u[2] = mt_canon(&gen) * (hi - lo) + lo;
                o = s->bnd[5] + 3 * (size_t)s->nbnd[5]++;
                o[0] = u[2] + ti; o[1] = u[1] + tj; o[2] = (depth + P->boxMax[2]) + u[0];
            }
}

int sfo_set_boundary(sfo_solver* s, int wall, const float* xyz, uint32_t n)
{
    if (wall < 0 || wall > 5) return -1;
    free(s->bnd[wall]);
    s->bnd[wall] = (float*)malloc((size_t)n * 12 + 4);
    memcpy(s->bnd[wall], xyz, (size_t)n * 12);
    s->nbnd[wall] = n;
    return 0;
}

uint32_t sfo_get_boundary(sfo_solver* s, int wall, float* xyz, uint32_t cap)
{
    if (wall < 0 || wall > 5) return 0;
    uint32_t n = s->nbnd[wall];
    if (xyz) memcpy(xyz, s->bnd[wall], (size_t)(n < cap ? n : cap) * 12);
    return n;
}

void sfo_make_ready(sfo_solver* s)
{ /* EXE@0x140016650 (A.3) */
    const sfo_params* P = &s->P;
    kernel_common(&s->cubic, P->kernelRadius, 0);
    kernel_common(&s->spiky, P->kernelRadius, 1);
    /* Grid3D::setGrid EXE@0x14001ab20 */
    s->cell = P->kernelRadius;
    s->ncells = 1;
    for (int d = 0; d < 3; ++d) {
        s->gmin[d] = P->boxMin[d];
        s->gmax[d] = P->boxMax[d];
        s->nc[d] = (int)ceilf((P->boxMax[d] - P->boxMin[d]) / s->cell);
        s->ncells *= (uint64_t)s->nc[d];
    }
    free(s->cellStart); free(s->cellIds); free(s->cellOf);
    s->cellStart = (uint32_t*)calloc(s->ncells + 1, 4);
    s->cellIds = (uint32_t*)malloc((size_t)s->n * 4 + 4);
    s->cellOf = (uint32_t*)malloc((size_t)s->n * 4 + 4);
    if (P->bUseBoundaryParticles && s->nbnd[0] == 0 && s->bnd[0] == NULL) sfo_generate_boundary(s, 0);
    free(s->acc); free(s->rho); free(s->rho2); free(s->visc); free(s->m2);
    s->acc = (float*)calloc((size_t)s->n * 3 + 4, 4);
    s->rho = (float*)calloc((size_t)s->n + 4, 4);
    s->rho2 = (float*)calloc((size_t)s->n + 4, 4);
    s->visc = (float*)calloc((size_t)s->n * 3 + 4, 4);
    s->m2 = (float*)calloc((size_t)s->n + 4, 4);
    s->ready = 1;
}

/* ---------------------------------------------------------------- the step */
static float compute_time_step(sfo_solver* s)
{ /* computeMaxVel EXE@0x140016a40 + computeTimeStep EXE@0x140016d00 (A.5) */
    const sfo_params* P = &s->P;
    const float* v = s->vel;
    uint32_t n = s->n;
    float M = FLT_MIN;
#pragma omp parallel for reduction(max : M) schedule(static)
    for (int64_t p = 0; p < (int64_t)n; ++p) {
        float vx = v[3 * p], vy = v[3 * p + 1], vz = v[3 * p + 2];
        float m = (vy * vy + vx * vx) + vz * vz;
        s->m2[p] = m;
        if (m > M) M = m;
    }
    float maxv = sqrtf(M);
    float r = P->particleRadius;
    float dt = ((double)maxv > 1e-8) ? ((r + r) / maxv) * 0.2f : (float)1e10;
    dt = fmaxf(dt, P->defaultTimestep * 0.1f);
    dt = fminf(dt, P->defaultTimestep * 10.0f);
    return dt;
}

static inline void cell_coords(const sfo_solver* s, const float* x, int c[3])
{ /* A.6: truncation, not clamped */
    for (int d = 0; d < 3; ++d) c[d] = (int)((x[d] - s->gmin[d]) / s->cell);
}

static void collect_particles_to_cells(sfo_solver* s)
{ /* EXE@0x140016890 (A.7) -- SERIAL in the reference (per-cell push_back in ascending p).
   * A counting sort gives the identical per-cell order. */
    uint32_t n = s->n;
    memset(s->cellStart, 0, (s->ncells + 1) * 4);
    for (uint32_t p = 0; p < n; ++p) {
        int c[3];
        cell_coords(s, s->pos + 3 * (size_t)p, c);
        for (int d = 0; d < 3; ++d) {
            c[d] = c[d] < s->nc[d] - 1 ? c[d] : s->nc[d] - 1;
            c[d] = c[d] > 0 ? c[d] : 0;
        }
        uint32_t key = (uint32_t)(((int64_t)c[2] * s->nc[1] + c[1]) * s->nc[0] + c[0]);
        s->cellOf[p] = key;
        s->cellStart[key + 1]++;
    }
    for (uint64_t c = 0; c < s->ncells; ++c) s->cellStart[c + 1] += s->cellStart[c];
    /* fill: ascending p inside each cell */
    uint32_t* cur = (uint32_t*)malloc(s->ncells * 4 + 4);
    memcpy(cur, s->cellStart, s->ncells * 4);
    for (uint32_t p = 0; p < n; ++p) s->cellIds[cur[s->cellOf[p]]++] = p;
    free(cur);
}

/* Iteration helpers shared by A.8/A.9/A.11/A.13: traversal lk(z) -> lj(y) -> li(x), cell lists in
 * stored order (A.6).  `reversed` flips every loop (self-divergence study only). */
#define FOR_NEIGHBOR_CELLS(s, c, ...)                                                             \
    for (int _a = 0; _a < 3; ++_a)                                                                  \
        for (int _b = 0; _b < 3; ++_b)                                                              \
            for (int _c = 0; _c < 3; ++_c) {                                                        \
                int lk = (s)->reversed ? 1 - _a : _a - 1;                                           \
                int lj = (s)->reversed ? 1 - _b : _b - 1;                                           \
                int li = (s)->reversed ? 1 - _c : _c - 1;                                           \
                int cx = (c)[0] + li, cy = (c)[1] + lj, cz = (c)[2] + lk;                           \
                if (cx < 0 || cy < 0 || cz < 0 || cx >= (s)->nc[0] || cy >= (s)->nc[1] || cz >= (s)->nc[2]) continue; \
                uint64_t _cell = ((uint64_t)cz * (s)->nc[1] + cy) * (s)->nc[0] + cx;                \
                uint32_t _b0 = (s)->cellStart[_cell], _e0 = (s)->cellStart[_cell + 1];              \
                for (uint32_t _t = _b0; _t < _e0; ++_t) {                                           \
                    uint32_t q = (s)->cellIds[(s)->reversed ? (_e0 - 1 - (_t - _b0)) : _t];         \
                    __VA_ARGS__                                                                     \
                }                                                                                   \
            }

/* wall lists (A.6): for axis A: if lo[A] > x[A] || x[A] > hi[A]: list = lower/upper wall of A;
 * xs = x - h*floorf(x/h) on the two tangential axes, x - h*0 on axis A. */
typedef struct { int nwalls; int wall[3]; float xs[3][3]; } wall_ctx;
static inline void wall_setup(const sfo_solver* s, const float* x, wall_ctx* w)
{
    const sfo_params* P = &s->P;
    float h = P->kernelRadius;
    w->nwalls = 0;
    if (!P->bUseBoundaryParticles) return;
    for (int A = 0; A < 3; ++A) {
        float lo = h + P->boxMin[A], hi = P->boxMax[A] - h;
        if (lo > x[A] || x[A] > hi) {
            int k = w->nwalls++;
            w->wall[k] = 2 * A + ((lo > x[A]) ? 0 : 1);
            for (int d = 0; d < 3; ++d) {
                float f = (d == A) ? 0.0f : floorf(x[d] / h);
                w->xs[k][d] = x[d] - h * f;
            }
        }
    }
}

static void compute_density(sfo_solver* s)
{ /* EXE@0x140017770 / lambda 0x1400179c0 (A.8) */
    const sfo_params* P = &s->P;
    const float rmin = (float)((double)P->restDensity * 0.1), rmax = (float)((double)P->restDensity * 10.0);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t p = 0; p < (int64_t)s->n; ++p) {
        const float* xp = s->pos + 3 * p;
        int c[3];
        cell_coords(s, xp, c);
        float S = s->cubic.Wzero;
        FOR_NEIGHBOR_CELLS(s, c, {
            if (q == (uint32_t)p) continue;
            const float* xq = s->pos + 3 * (size_t)q;
            float dx = xq[0] - xp[0], dy = xq[1] - xp[1], dz = xq[2] - xp[2];
            float d2 = (dx * dx + dy * dy) + dz * dz;
            S += W_lookup(&s->cubic, d2);
        })
        wall_ctx w;
        wall_setup(s, xp, &w);
        for (int k = 0; k < w.nwalls; ++k) {
            const float* B = s->bnd[w.wall[k]];
            uint32_t nb = s->nbnd[w.wall[k]];
            for (uint32_t b = 0; b < nb; ++b) {
                float dx = B[3 * b] - w.xs[k][0], dy = B[3 * b + 1] - w.xs[k][1], dz = B[3 * b + 2] - w.xs[k][2];
                float d2 = (dx * dx + dy * dy) + dz * dz;
                S += W_lookup(&s->cubic, d2);
            }
        }
        s->rho[p] = (1.0f > S) ? 0.0f : fminf(fmaxf(S * P->particleMass, rmin), rmax);
    }
}

static void correct_density(sfo_solver* s)
{ /* EXE@0x140018150 / lambda 0x140018420 (A.9), default off */
    const sfo_params* P = &s->P;
    const float rmax = (float)((double)P->restDensity * 10.0);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t p = 0; p < (int64_t)s->n; ++p) {
        const float* xp = s->pos + 3 * p;
        int c[3];
        cell_coords(s, xp, c);
        float T = s->cubic.Wzero / s->rho[p];
        FOR_NEIGHBOR_CELLS(s, c, {
            if (q == (uint32_t)p) continue;
            const float* xq = s->pos + 3 * (size_t)q;
            float dx = xq[0] - xp[0], dy = xq[1] - xp[1], dz = xq[2] - xp[2];
            float d2 = (dx * dx + dy * dy) + dz * dz;
            float rq = s->rho[q];
            if (!((double)rq >= 1e-8)) continue;
            T += W_lookup(&s->cubic, d2) / rq;
        })
        wall_ctx w;
        wall_setup(s, xp, &w);
        for (int k = 0; k < w.nwalls; ++k) {
            const float* B = s->bnd[w.wall[k]];
            uint32_t nb = s->nbnd[w.wall[k]];
            for (uint32_t b = 0; b < nb; ++b) {
                float dx = B[3 * b] - w.xs[k][0], dy = B[3 * b + 1] - w.xs[k][1], dz = B[3 * b + 2] - w.xs[k][2];
                float d2 = (dx * dx + dy * dy) + dz * dz;
                T += W_lookup(&s->cubic, d2) / P->restDensity;
            }
        }
        s->rho2[p] = ((double)T > 1e-8) ? s->rho[p] / fminf(T * P->particleMass, rmax) : 0.0f;
    }
    memcpy(s->rho, s->rho2, (size_t)s->n * 4);
}

static inline float pressure_of(const sfo_params* P, float rho)
{ /* A.11 Pr(rho) */
    float x = rho / P->restDensity;
    float t = x * x;
    t = t * x;
    t = t * t;
    t = t * x;
    float pp = (float)((double)t - 1.0);
    if (P->bUseAttractivePressure) return fmaxf(pp, pp * P->attractivePressureRatio);
    return (float)fmax((double)pp, 0.0);
}

static void add_gravity(sfo_solver* s, float dt)
{ /* EXE@0x140018bc0 / body 0x14001df70 (A.10) */
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < (int64_t)s->n; ++p) s->vel[3 * p + 1] = (float)((double)s->vel[3 * p + 1] - (double)dt * 9.8);
}

static void compute_pressure_forces(sfo_solver* s)
{ /* EXE@0x140018d10 / lambda 0x140018f20 (A.11) */
    const sfo_params* P = &s->P;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t p = 0; p < (int64_t)s->n; ++p) {
        float* a = s->acc + 3 * p;
        float rp = s->rho[p];
        if (1e-8 > (double)rp) { a[0] = a[1] = a[2] = 0.0f; continue; }
        const float* xp = s->pos + 3 * p;
        int c[3];
        cell_coords(s, xp, c);
        float Pp = pressure_of(P, rp);
        float ax = 0.0f, ay = 0.0f, az = 0.0f;
        FOR_NEIGHBOR_CELLS(s, c, {
            if (q == (uint32_t)p) continue;
            float rq = s->rho[q];
            if (1e-8 > (double)rq) continue;
            const float* xq = s->pos + 3 * (size_t)q;
            float dx = xq[0] - xp[0], dy = xq[1] - xp[1], dz = xq[2] - xp[2];
            float d2 = (dx * dx + dy * dy) + dz * dz;
            if (d2 > P->kernelRadiusSqr) continue;
            float Pq = pressure_of(P, rq);
            float g = G_lookup(&s->spiky, d2);
            float Gx = g * dx, Gy = dy * g, Gz = g * dz;
            float fp = Pq / (rq * rq) + Pp / (rp * rp);
            ax += fp * Gx; ay += fp * Gy; az += fp * Gz;
        })
        wall_ctx w;
        wall_setup(s, xp, &w);
        for (int k = 0; k < w.nwalls; ++k) {
            const float* B = s->bnd[w.wall[k]];
            uint32_t nb = s->nbnd[w.wall[k]];
            float fb = Pp / (rp * rp);
            for (uint32_t b = 0; b < nb; ++b) {
                float dx = B[3 * b] - w.xs[k][0], dy = B[3 * b + 1] - w.xs[k][1], dz = B[3 * b + 2] - w.xs[k][2];
                float d2 = (dx * dx + dy * dy) + dz * dz;
                float g = G_lookup(&s->spiky, d2);
                float Gx = g * dx, Gy = dy * g, Gz = g * dz;
                ax += fb * Gx; ay += fb * Gy; az += fb * Gz;
            }
        }
        a[0] = (ax * P->particleMass) * P->pressureStiffness;
        a[1] = (ay * P->particleMass) * P->pressureStiffness;
        a[2] = (az * P->particleMass) * P->pressureStiffness;
    }
}

static void update_velocity(sfo_solver* s, float dt)
{ /* EXE@0x140019a60 / body 0x14001dd20 (A.12) */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)s->n * 3; ++i) s->vel[i] = dt * s->acc[i] + s->vel[i];
}

static void compute_viscosity(sfo_solver* s)
{ /* EXE@0x140019bb0 / lambda 0x140019e60 + body 0x14001dae0 (A.13) */
    const sfo_params* P = &s->P;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t p = 0; p < (int64_t)s->n; ++p) {
        const float* xp = s->pos + 3 * p;
        const float* vp = s->vel + 3 * p;
        int c[3];
        cell_coords(s, xp, c);
        float sx = 0.0f, sy = 0.0f, sz = 0.0f;
        FOR_NEIGHBOR_CELLS(s, c, {
            if (q == (uint32_t)p) continue;
            const float* xq = s->pos + 3 * (size_t)q;
            float dx = xq[0] - xp[0], dy = xq[1] - xp[1], dz = xq[2] - xp[2];
            float d2 = (dx * dx + dy * dy) + dz * dz;
            if (d2 > P->kernelRadiusSqr) continue;
            float w = W_lookup(&s->cubic, d2);
            float inv = 1.0f / s->rho[q];
            const float* vq = s->vel + 3 * (size_t)q;
            float dvx = vq[0] - vp[0], dvy = vq[1] - vp[1], dvz = vq[2] - vp[2];
            sx += (inv * dvx) * w; sy += (dvy * inv) * w; sz += (dvz * inv) * w;
        })
        s->visc[3 * p] = sx * P->particleMass;
        s->visc[3 * p + 1] = sy * P->particleMass;
        s->visc[3 * p + 2] = sz * P->particleMass;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)s->n * 3; ++i) s->vel[i] = P->viscosity * s->visc[i] + s->vel[i];
}

static void update_position(sfo_solver* s, float dt)
{ /* EXE@0x140017400 / lambda 0x1400175d0 (A.14) */
    const sfo_params* P = &s->P;
    float r = P->particleRadius;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < (int64_t)s->n; ++p) {
        float* x = s->pos + 3 * p;
        float* v = s->vel + 3 * p;
        float nv[3] = { v[0], v[1], v[2] };
        int ch = 0;
        for (int d = 0; d < 3; ++d) {
            float lo = P->boxMin[d] + r, hi = P->boxMax[d] - r;
            float xn = v[d] * dt + x[d];
            if (lo > xn) { xn = lo; nv[d] = -(v[d] * P->boundaryRestitution); ch = 1; }
            else if (xn > hi) { xn = hi; nv[d] = -(v[d] * P->boundaryRestitution); ch = 1; }
            x[d] = xn;
        }
        if (ch) { v[0] = nv[0]; v[1] = nv[1]; v[2] = nv[2]; }
    }
}

float sfo_advance_frame(sfo_solver* s)
{ /* EXE@0x140016810 (A.15) */
    if (!s->ready) sfo_make_ready(s);
#ifdef _OPENMP
    omp_set_num_threads(s->nthreads > 0 ? s->nthreads : omp_get_num_procs()); /* 0 = all host cores (TBB "automatic") */
#endif
    double t0 = now_s();
    float dt = compute_time_step(s);
    double t1 = now_s();
    collect_particles_to_cells(s);
    double t2 = now_s();
    compute_density(s);
    if (s->P.bCorrectDensity) correct_density(s);
    double t3 = now_s();
    add_gravity(s, dt);
    compute_pressure_forces(s);
    update_velocity(s, dt);
    double t4 = now_s();
    compute_viscosity(s);
    double t5 = now_s();
    update_position(s, dt);
    double t6 = now_s();
    s->timing[0] = t1 - t0; s->timing[1] = t2 - t1; s->timing[2] = t3 - t2;
    s->timing[3] = t4 - t3; s->timing[4] = t5 - t4; s->timing[5] = t6 - t5;
    return dt;
}

void sfo_last_timing(sfo_solver* s, double* t6) { memcpy(t6, s->timing, sizeof(s->timing)); }

/* ---------------------------------------------------------------- accessors */
uint32_t sfo_num_particles(sfo_solver* s) { return s->n; }
const float* sfo_positions(sfo_solver* s) { return s->pos; }
const float* sfo_velocities(sfo_solver* s) { return s->vel; }
const float* sfo_density(sfo_solver* s) { return s->rho; }
const float* sfo_accel(sfo_solver* s) { return s->acc; }
void sfo_pressure(sfo_solver* s, float* out)
{
    for (uint32_t p = 0; p < s->n; ++p) out[p] = pressure_of(&s->P, s->rho[p]);
}
void sfo_grid_dims(sfo_solver* s, int32_t* n3) { n3[0] = s->nc[0]; n3[1] = s->nc[1]; n3[2] = s->nc[2]; }
void sfo_cell_index(sfo_solver* s, uint32_t* out) { memcpy(out, s->cellOf, (size_t)s->n * 4); }

static int cmp_u32(const void* a, const void* b)
{
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return (x > y) - (x < y);
}

/* Neighbour sets of the binning of the last step, evaluated on the positions that were binned.
 * Call it after sfo_collect_only() / before the positions move, or use sfo_neighbors_now(). */
uint64_t sfo_neighbors(sfo_solver* s, uint32_t* counts, uint32_t* ids, uint64_t cap)
{ /* test helper (not part of the reference step): the A.6 traversal with the A.13 / A.11 range test, two parallel passes */
    if (!s->ready) sfo_make_ready(s);
    int rev = s->reversed;
    s->reversed = 0;
    collect_particles_to_cells(s);
    uint64_t* off = (uint64_t*)malloc(((size_t)s->n + 1) * sizeof(uint64_t));
    omp_set_num_threads(s->nthreads > 0 ? s->nthreads : omp_get_num_procs());
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t pi = 0; pi < (int64_t)s->n; ++pi) {
        uint32_t p = (uint32_t)pi;
        const float* xp = s->pos + 3 * (size_t)p;
        int c[3];
        cell_coords(s, xp, c);
        uint32_t cnt = 0;
        FOR_NEIGHBOR_CELLS(s, c, {
            if (q == p) continue;
            const float* xq = s->pos + 3 * (size_t)q;
            float dx = xq[0] - xp[0], dy = xq[1] - xp[1], dz = xq[2] - xp[2];
            float d2 = (dx * dx + dy * dy) + dz * dz;
            if (d2 > s->P.kernelRadiusSqr) continue;
            cnt++;
        })
        off[p + 1] = cnt;
        if (counts) counts[p] = cnt;
    }
    off[0] = 0;
    for (uint32_t p = 0; p < s->n; ++p) off[p + 1] += off[p];
    uint64_t total = off[s->n];
    if (ids && total <= cap) {
#pragma omp parallel for schedule(dynamic, 1024)
        for (int64_t pi = 0; pi < (int64_t)s->n; ++pi) {
            uint32_t p = (uint32_t)pi;
            const float* xp = s->pos + 3 * (size_t)p;
            int c[3];
            cell_coords(s, xp, c);
            uint64_t o = off[p];
            FOR_NEIGHBOR_CELLS(s, c, {
                if (q == p) continue;
                const float* xq = s->pos + 3 * (size_t)q;
                float dx = xq[0] - xp[0], dy = xq[1] - xp[1], dz = xq[2] - xp[2];
                float d2 = (dx * dx + dy * dy) + dz * dz;
                if (d2 > s->P.kernelRadiusSqr) continue;
                ids[o++] = q;
            })
            qsort(ids + off[p], (size_t)(off[p + 1] - off[p]), 4, cmp_u32);
        }
    }
    free(off);
    s->reversed = rev;
    return total;
}

void sfo_table(sfo_solver* s, int which, float* out10001)
{
    const sfo_kernel* K = which ? &s->spiky : &s->cubic;
    memcpy(out10001, K->tab + (which ? SFO_TAB : 0), (SFO_TAB + 1) * 4);
}

void sfo_kernel_consts(sfo_solver* s, float* out4)
{
    out4[0] = s->cubic.Wzero; out4[1] = s->cubic.radius2; out4[2] = s->cubic.invStep; out4[3] = s->spiky.radius2;
}
