/*
 * libmnv_host.so -- native host-side half of VecMarineNavEnv.step_host (no counterpart in the reference, which steps one
 * environment and returns a numpy row): expands the compact observation packet written by mnv_pack_obs
 * (csrc/mnv_pack.cu) into the dense row-major f32 [E][obs_dim] array the API returns (marinenav_env.py:273-326 layout),
 * IN PLACE.  Packet: head f32 [E][4]; mask u32 [E][W] (bit b: beam b has a sonar return, W = ceil(n_beams / 32));
 * dir u32 [ceil(E / 32)] (start of the returns of environments 32 g .. 32 g + 31 in vals); vals f32 [.][2] (x, y) of a
 * group's returns, environment by environment, beam by beam.  Per row the expander rewrites the 4 head values, zeroes the
 * beam slots that held a return in the previous packet and hold none now ("no return" = (0, 0), marinenav_env.py:318-320)
 * and writes the new returns: one sequential sweep over the rows, touching only what changes.  Rows flagged in `skip`
 * (finished environments, whose row the GPU has already overwritten with the first observation of the next episode) are
 * left alone and re-scanned so that the bookkeeping (the mask of the previous packet, kept per environment) follows them.
 * A small pool of worker threads (the caller is worker 0) splits the groups into contiguous ranges, so no two threads
 * write the same row.  Workers spin briefly between jobs (a step loop keeps them hot) and then sleep.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    float* obs; const float* head; const uint8_t* skip; const uint32_t* mask; const uint32_t* dir; const float* vals; int rescan_only;
} job_t;

typedef struct mnvh_pool {
    int n_threads; int64_t E; int D, n_beams, W;
    pthread_t* tid;
    uint32_t* prev_mask;                        /* [E][W]: beam slots currently non-zero in the dense array */
    int32_t* skipped; int64_t* n_skipped;       /* rows mnvh_expand_early left for mnvh_rescan_skipped: worker t's rows start at
                                                   skipped[32 * g_lo[t]], n_skipped[8 * t] of them */
    int64_t* g_lo; int64_t* g_hi;               /* group range of every worker */
    int cpu_first, pin;
    job_t job;
    char pad0[64];
    atomic_int generation; char pad1[60];
    atomic_int finished; char pad2[60];
    atomic_int stop;
} mnvh_pool;

typedef struct { mnvh_pool* p; int t; } arg_t;

static inline uint32_t row_mask(const float* r, int b0, int b1)
{
    uint32_t m = 0;
    for (int b = b0; b < b1; ++b) m |= (r[2 * b] != 0.0f || r[2 * b + 1] != 0.0f) ? 1u << (b & 31) : 0u;
    return m;
}

static void work(mnvh_pool* p, int t)
{
    const job_t* j = &p->job;
    const int D = p->D, nb = p->n_beams, W = p->W;
    const int64_t E = p->E;
    float* obs = j->obs;
    uint32_t* pm = p->prev_mask;
    int32_t* my_skipped = p->skipped + 32 * p->g_lo[t];
    int64_t n_my = 0;
    for (int64_t g = p->g_lo[t]; g < p->g_hi[t]; ++g) {
        const int64_t e0 = g * 32, e1 = e0 + 32 < E ? e0 + 32 : E;
        if (j->rescan_only == 1) {                                /* job kinds: 0 expand, 1 rescan every row, 2 expand early */
            for (int64_t e = e0; e < e1; ++e)
                for (int w = 0; w < W; ++w) pm[e * W + w] = row_mask(obs + e * D + 4, 32 * w, 32 * w + 32 < nb ? 32 * w + 32 : nb);
            continue;
        }
        const float* v = j->vals + 2 * (size_t)j->dir[g];
        for (int64_t e = e0; e < e1; ++e) {
            float* row = obs + e * D;
            if (j->skip != NULL && j->skip[e]) {                  /* written by the GPU: pick up what the row holds */
                for (int w = 0; w < W; ++w) {
                    v += 2 * __builtin_popcount(j->mask[e * W + w]);
                    if (j->rescan_only != 2)                     /* 2: the GPU's row arrives later, mnvh_rescan_skipped follows */
                        pm[e * W + w] = row_mask(row + 4, 32 * w, 32 * w + 32 < nb ? 32 * w + 32 : nb);
                }
                if (j->rescan_only == 2) my_skipped[n_my++] = (int32_t)e;
                continue;
            }
            memcpy(row, j->head + 4 * e, 16);
            for (int w = 0; w < W; ++w) {
                const uint32_t m = j->mask[e * W + w];
                uint32_t clr = pm[e * W + w] & ~m, set = m;
                float* beams = row + 4 + 64 * w;
                while (clr) { const int b = __builtin_ctz(clr); clr &= clr - 1; beams[2 * b] = 0.0f; beams[2 * b + 1] = 0.0f; }
                while (set) { const int b = __builtin_ctz(set); set &= set - 1; memcpy(beams + 2 * b, v, 8); v += 2; }
                pm[e * W + w] = m;
            }
        }
    }
    if (j->rescan_only == 2) p->n_skipped[8 * t] = n_my;
}

static void* worker(void* a_)
{
    arg_t* a = (arg_t*)a_;
    mnvh_pool* p = a->p;
    const int t = a->t;
    free(a);
    if (p->pin) {
        cpu_set_t set; CPU_ZERO(&set); CPU_SET(p->cpu_first + t, &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    int seen = 0;
    for (;;) {
        int spins = 0;
        while (atomic_load_explicit(&p->generation, memory_order_acquire) == seen) {
            if (atomic_load_explicit(&p->stop, memory_order_relaxed)) return NULL;
            if (++spins < 30000) { __builtin_ia32_pause(); }           /* ~1 ms: a step loop (200 - 400 us per step) keeps them hot */
            else { struct timespec ts = {0, 20000}; nanosleep(&ts, NULL); }
        }
        seen = atomic_load_explicit(&p->generation, memory_order_acquire);
        work(p, t);
        atomic_fetch_add_explicit(&p->finished, 1, memory_order_release);
    }
}

static void run(mnvh_pool* p)
{
    atomic_store_explicit(&p->finished, 0, memory_order_relaxed);
    atomic_fetch_add_explicit(&p->generation, 1, memory_order_release);
    work(p, 0);
    while (atomic_load_explicit(&p->finished, memory_order_acquire) != p->n_threads - 1) __builtin_ia32_pause();
}

mnvh_pool* mnvh_create(int n_threads, int64_t E, int obs_dim, int cpu_first)
{
    if (n_threads < 1 || E <= 0 || obs_dim < 6 || (obs_dim & 1)) return NULL;
    const int64_t groups = (E + 31) / 32;
    if ((int64_t)n_threads > groups) n_threads = (int)groups;
    mnvh_pool* p = (mnvh_pool*)calloc(1, sizeof(*p));
    p->n_threads = n_threads; p->E = E; p->D = obs_dim; p->n_beams = (obs_dim - 4) / 2; p->W = (p->n_beams + 31) / 32;
    p->pin = cpu_first >= 0; p->cpu_first = cpu_first;
    p->prev_mask = (uint32_t*)calloc((size_t)E * p->W, sizeof(uint32_t));
    p->skipped = (int32_t*)calloc((size_t)groups * 32, sizeof(int32_t));
    p->n_skipped = (int64_t*)calloc((size_t)n_threads * 8, sizeof(int64_t));
    p->g_lo = (int64_t*)calloc(n_threads, sizeof(int64_t)); p->g_hi = (int64_t*)calloc(n_threads, sizeof(int64_t));
    for (int t = 0; t < n_threads; ++t) { p->g_lo[t] = groups * t / n_threads; p->g_hi[t] = groups * (t + 1) / n_threads; }
    p->tid = (pthread_t*)calloc(n_threads, sizeof(pthread_t));
    if (p->pin) {                                       /* the caller is worker 0 */
        cpu_set_t set; CPU_ZERO(&set); CPU_SET(cpu_first, &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    for (int t = 1; t < n_threads; ++t) {
        arg_t* a = (arg_t*)malloc(sizeof(arg_t)); a->p = p; a->t = t;
        pthread_create(&p->tid[t], NULL, worker, a);
    }
    return p;
}

void mnvh_destroy(mnvh_pool* p)
{
    if (p == NULL) return;
    atomic_store(&p->stop, 1);
    for (int t = 1; t < p->n_threads; ++t) pthread_join(p->tid[t], NULL);
    free(p->prev_mask); free(p->skipped); free(p->n_skipped); free(p->g_lo); free(p->g_hi); free(p->tid); free(p);
}

int mnvh_threads(const mnvh_pool* p) { return p ? p->n_threads : 0; }

/* obs holds a complete dense block (e.g. after a reset): rebuild the bookkeeping from it */
void mnvh_rescan(mnvh_pool* p, float* obs)
{
    p->job = (job_t){obs, NULL, NULL, NULL, NULL, NULL, 1};
    run(p);
}

/* one packet -> dense block, in place; skip (u8 [E], may be NULL): rows to leave alone */
void mnvh_expand(mnvh_pool* p, float* obs, const float* head, const uint8_t* skip, const uint32_t* mask, const uint32_t* dir,
                 const float* vals)
{
    p->job = (job_t){obs, head, skip, mask, dir, vals, 0};
    run(p);
}

/* mnvh_expand for a packet that arrives BEFORE the caller's own writes of the rows flagged in skip have landed (the hybrid
 * transport expands while the GPU still works): those rows are neither written nor re-scanned here ... */
void mnvh_expand_early(mnvh_pool* p, float* obs, const float* head, const uint8_t* skip, const uint32_t* mask, const uint32_t* dir,
                       const float* vals)
{
    p->job = (job_t){obs, head, skip, mask, dir, vals, 2};
    run(p);
}

/* ... and this picks them up once they have: re-scan of the rows flagged in skip. */
void mnvh_rescan_skipped(mnvh_pool* p, float* obs, const uint8_t* skip)
{
    (void)skip;
    /* the rows mnvh_expand_early skipped (each worker listed its own): a few hundred, not worth waking the pool */
    const int D = p->D, nb = p->n_beams, W = p->W;
    for (int t = 0; t < p->n_threads; ++t) {
        const int32_t* rows = p->skipped + 32 * p->g_lo[t];
        for (int64_t k = 0; k < p->n_skipped[8 * t]; ++k) {
            const int64_t e = rows[k];
            for (int w = 0; w < W; ++w) p->prev_mask[e * W + w] = row_mask(obs + e * D + 4, 32 * w, 32 * w + 32 < nb ? 32 * w + 32 : nb);
        }
        p->n_skipped[8 * t] = 0;
    }
}
