/*
 * libmnv_host.so -- native host-side half of VecMarineNavEnv.step_host (no counterpart in the reference, which steps one
 * environment and returns a numpy row): expands the compact observation packet written by mnv_pack_obs
 * (csrc/mnv_pack.cu: head f32 [E][4] + a list of (env << 8 | beam, x, y) for the beams with a sonar return) into the dense
 * row-major f32 [E][obs_dim] array the API returns (marinenav_env.py:273-326 layout), IN PLACE:
 *   - every row's 4 head values are rewritten;
 *   - the beam slots set by the previous packet are zeroed ("no return" = (0, 0), marinenav_env.py:318-320) and the new
 *     returns are written: only the slots that change are touched;
 *   - rows of environments flagged in `skip` (finished environments, whose row the GPU has already overwritten with the
 *     first observation of the next episode) are left alone and re-scanned so that the bookkeeping follows them.
 * A small pool of worker threads (the caller is worker 0) splits the environments into contiguous ranges; a worker applies
 * the list entries of its own range only, so no two threads write the same row.  Workers spin briefly between jobs (a
 * step loop keeps them hot) and then sleep.
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
    float* obs; const float* head; const uint8_t* skip; const uint32_t* hits; uint32_t n_hits; int rescan_only;
} job_t;

typedef struct mnvh_pool {
    int n_threads; int64_t E; int D, n_beams;
    pthread_t* tid;
    atomic_int generation, finished, stop;
    job_t job;
    uint32_t** prev; uint32_t* n_prev;          /* per worker: the (env << 8 | beam) slots currently non-zero in its range */
    int64_t* lo; int64_t* hi;
    int cpu_first, pin;
} mnvh_pool;

typedef struct { mnvh_pool* p; int t; } arg_t;

static void work(mnvh_pool* p, int t)
{
    const job_t* j = &p->job;
    const int D = p->D, nb = p->n_beams;
    const int64_t lo = p->lo[t], hi = p->hi[t];
    float* obs = j->obs;
    uint32_t* prev = p->prev[t];
    uint32_t n = 0;
    if (j->rescan_only) {
        for (int64_t e = lo; e < hi; ++e) {
            const float* r = obs + e * D + 4;
            for (int b = 0; b < nb; ++b)
                if (r[2 * b] != 0.0f || r[2 * b + 1] != 0.0f) prev[n++] = ((uint32_t)e << 8) | (uint32_t)b;
        }
        p->n_prev[t] = n;
        return;
    }
    const uint8_t* skip = j->skip;
    const uint32_t np = p->n_prev[t];
    for (uint32_t k = 0; k < np; ++k) {                 /* last packet's returns -> "no return" */
        const uint32_t eb = prev[k];
        const int64_t e = eb >> 8;
        if (skip != NULL && skip[e]) continue;
        float* s = obs + e * D + 4 + 2 * (eb & 255u);
        s[0] = 0.0f; s[1] = 0.0f;
    }
    const float* head = j->head;
    if (skip == NULL) {
        for (int64_t e = lo; e < hi; ++e) memcpy(obs + e * D, head + 4 * e, 16);
    } else {
        for (int64_t e = lo; e < hi; ++e)
            if (!skip[e]) memcpy(obs + e * D, head + 4 * e, 16);
    }
    const uint32_t* h = j->hits;
    const uint32_t nh = j->n_hits;
    for (uint32_t k = 0; k < nh; ++k) {                  /* this packet's returns that fall into my range */
        const uint32_t eb = h[3 * k];
        const int64_t e = eb >> 8;
        if (e < lo || e >= hi || (skip != NULL && skip[e])) continue;
        float* s = obs + e * D + 4 + 2 * (eb & 255u);
        memcpy(s, h + 3 * k + 1, 8);
        prev[n++] = eb;
    }
    if (skip != NULL) {
        for (int64_t e = lo; e < hi; ++e) {              /* rows written by the GPU: pick up what they hold */
            if (!skip[e]) continue;
            const float* r = obs + e * D + 4;
            for (int b = 0; b < nb; ++b)
                if (r[2 * b] != 0.0f || r[2 * b + 1] != 0.0f) prev[n++] = ((uint32_t)e << 8) | (uint32_t)b;
        }
    }
    p->n_prev[t] = n;
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
            if (++spins < 20000) { __builtin_ia32_pause(); }
            else { struct timespec ts = {0, 50000}; nanosleep(&ts, NULL); }
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
    if (n_threads < 1 || E <= 0 || E >= (1ll << 24) || obs_dim < 6 || (obs_dim & 1)) return NULL;
    if ((int64_t)n_threads > E) n_threads = (int)E;
    mnvh_pool* p = (mnvh_pool*)calloc(1, sizeof(*p));
    p->n_threads = n_threads; p->E = E; p->D = obs_dim; p->n_beams = (obs_dim - 4) / 2;
    p->pin = cpu_first >= 0; p->cpu_first = cpu_first;
    p->prev = (uint32_t**)calloc(n_threads, sizeof(uint32_t*));
    p->n_prev = (uint32_t*)calloc(n_threads, sizeof(uint32_t));
    p->lo = (int64_t*)calloc(n_threads, sizeof(int64_t)); p->hi = (int64_t*)calloc(n_threads, sizeof(int64_t));
    for (int t = 0; t < n_threads; ++t) {
        p->lo[t] = E * t / n_threads; p->hi[t] = E * (t + 1) / n_threads;
        p->prev[t] = (uint32_t*)malloc((size_t)(p->hi[t] - p->lo[t]) * p->n_beams * sizeof(uint32_t) + 64);
    }
    p->tid = (pthread_t*)calloc(n_threads, sizeof(pthread_t));
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
    for (int t = 0; t < p->n_threads; ++t) free(p->prev[t]);
    free(p->prev); free(p->n_prev); free(p->lo); free(p->hi); free(p->tid); free(p);
}

int mnvh_threads(const mnvh_pool* p) { return p ? p->n_threads : 0; }

/* obs holds a complete dense block (e.g. after a reset): rebuild the bookkeeping from it */
void mnvh_rescan(mnvh_pool* p, float* obs)
{
    p->job = (job_t){obs, NULL, NULL, NULL, 0, 1};
    run(p);
}

/* one packet -> dense block, in place; skip (u8 [E], may be NULL): rows to leave alone */
void mnvh_expand(mnvh_pool* p, float* obs, const float* head, const uint8_t* skip, const uint32_t* hits, uint32_t n_hits)
{
    p->job = (job_t){obs, head, skip, hits, n_hits, 0};
    run(p);
}
