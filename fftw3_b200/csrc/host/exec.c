/* exec.c -- run a plan: bind buffers, launch the recorded passes.
 *
 * Reference counterpart: api/execute.c:23-27 and the new-array variants
 * api/execute-dft.c:25-32, execute-dft-r2c.c, execute-dft-c2r.c, execute-r2r.c,
 * which call the root plan's apply() with (possibly new) pointers.
 *
 * Pointers may be device pointers (zero-copy: the passes run directly on them)
 * or host pointers (FFTW's classic contract): then the touched byte ranges are
 * staged to the GPU, transformed there and copied back, synchronously.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "b2_internal.h"

int b2_async_mode = 0;

static size_t real_size(int prec) { return prec == B2D_F32 ? 4 : 8; }

static void *resolve(const b2_plan *p, b2_ref r, void *const user[4], size_t rs)
{
    switch (r.buf) {
    case BUF_IN0: case BUF_IN1: case BUF_OUT0: case BUF_OUT1:
        return user[r.buf - BUF_IN0] ? (char *)user[r.buf - BUF_IN0] + r.off * (int64_t)rs : NULL;
    case BUF_SCRATCH0: case BUF_SCRATCH1: case BUF_SCRATCH2: case BUF_SCRATCH3: case BUF_SCRATCH4: case BUF_SCRATCH5:
        return (char *)p->scratch[r.buf - BUF_SCRATCH0] + r.off * (int64_t)rs;
    default:
        return NULL;
    }
}

/* Steps carry a lane: lane 0 runs on the caller's stream, lane k > 0 on side stream k - 1.  Steps of
   different lanes between two lane-0 steps are independent (the planner guarantees it: L2-resident
   groups of a multi-dimensional transform), so their kernels overlap: one lane's tail is filled by
   the next lane's head.  A lane is forked from the caller's stream the first time it is used after a
   join and joined back before the next lane-0 step (and at the end). */
#define B2_MAX_LANES 6      /* side streams 2..7 (0 and 1 belong to the distributed stages, dist.c) */

static void join_lanes(void *mainst, unsigned *active)
{
    int k;
    for (k = 1; k <= B2_MAX_LANES; ++k)
        if (*active & (1u << k)) { void *aux = b2d_aux_stream(k + 1); if (aux) b2d_stream_wait_stream(mainst, aux); }
    *active = 0;
}

static int run_steps(const b2_plan *p, void *const user[4])
{
    int i;
    size_t rs = real_size(p->prob.prec);
    void *mainst = b2d_get_stream();
    unsigned active = 0;
    for (i = 0; i < p->nsteps; ++i) {
        const b2_step *s = &p->steps[i];
        int rc = 0;
        void *aux = NULL, *prev = NULL;
        if (s->lane > 0 && s->lane <= B2_MAX_LANES) aux = b2d_aux_stream(s->lane + 1);
        if (s->fence && active) join_lanes(mainst, &active);
        if (aux) {
            if (!(active & (1u << s->lane))) { b2d_stream_wait_stream(aux, mainst); active |= 1u << s->lane; }
            prev = b2d_push_stream(aux);
        } else if (active) join_lanes(mainst, &active);
        if (s->kind == STEP_FFT) {
            b2d_fft_pass f = s->u.fft;
            f.in_re = resolve(p, s->r[0], user, rs);
            f.in_im = resolve(p, s->r[1], user, rs);
            f.out_re = resolve(p, s->r[2], user, rs);
            f.out_im = resolve(p, s->r[3], user, rs);
            rc = b2d_launch_fft_pass(&f);
        } else if (s->kind == STEP_COPY) {
            b2d_copy c = s->u.copy;
            c.in = resolve(p, s->r[0], user, rs);
            c.out = resolve(p, s->r[1], user, rs);
            rc = b2d_launch_copy(&c);
        } else if (s->kind == STEP_SPLIT) {
            b2d_split_pass sp = s->u.split;
            sp.user_re = resolve(p, s->r[0], user, rs);
            sp.user_im = resolve(p, s->r[1], user, rs);
            sp.work = resolve(p, s->r[4], user, rs);
            rc = b2d_launch_split_pass(&sp);
        } else {
            b2d_realop r = s->u.rop;
            r.x_re = resolve(p, s->r[0], user, rs);
            r.x_im = resolve(p, s->r[1], user, rs);
            r.y_re = resolve(p, s->r[2], user, rs);
            r.y_im = resolve(p, s->r[3], user, rs);
            r.work = resolve(p, s->r[4], user, rs);
            rc = b2d_launch_realop(&r);
        }
        if (aux) b2d_pop_stream(prev);
        if (rc) {
            fprintf(stderr, "fftw3_b200: pass %d failed: %s\n", i, b2d_last_error());
            if (active) join_lanes(mainst, &active);
            return rc;
        }
    }
    if (active) join_lanes(mainst, &active);
    return 0;
}

/* ---- host staging ---- */
typedef struct { char *lo, *hi; int has_in, has_out; } region;

/* tensor of everything one user pointer touches */
static void touched_span(const b2_problem *q, int which /*0..3*/, int64_t *lo, int64_t *hi)
{
    int is_out = which >= 2;
    b2_tensor t = q->sz;
    int i;
    /* half-complex side of r2c / c2r: last dim has n/2+1 entries */
    if (t.rnk > 0 && ((q->kind == B2_R2C && is_out) || (q->kind == B2_C2R && !is_out)))
        t.d[t.rnk - 1].n = t.d[t.rnk - 1].n / 2 + 1;
    for (i = 0; i < q->vecsz.rnk && t.rnk < B2_MAXRANK; ++i) t.d[t.rnk++] = q->vecsz.d[i];
    b2_tensor_span(&t, is_out, lo, hi);
}

void b2_problem_span(const b2_problem *q, int which, int64_t *lo, int64_t *hi) { touched_span(q, which, lo, hi); }

static int64_t touched_count(const b2_problem *q, int which)
{
    int is_out = which >= 2;
    b2_tensor t = q->sz;
    if (t.rnk > 0 && ((q->kind == B2_R2C && is_out) || (q->kind == B2_C2R && !is_out)))
        t.d[t.rnk - 1].n = t.d[t.rnk - 1].n / 2 + 1;
    return b2_tensor_count(&t) * b2_tensor_count(&q->vecsz);
}

/* async: enqueue only (upload, passes, download all on the current stream); the caller synchronises */
static int execute_host(b2_plan *p, void *const user[4], int async)
{
    const b2_problem *q = &p->prob;
    size_t rs = real_size(q->prec);
    region reg[4];
    int nreg = 0, map[4], i, j, rc = 0;
    void *dev_user[4];
    int64_t written_reals = 0;

    for (i = 0; i < 4; ++i) {
        int64_t lo, hi;
        char *a, *b;
        map[i] = -1;
        if (!user[i]) continue;
        touched_span(q, i, &lo, &hi);
        a = (char *)user[i] + lo * (int64_t)rs;
        b = (char *)user[i] + (hi + 1) * (int64_t)rs;
        if (i >= 2) written_reals += touched_count(q, i);
        /* merge with an existing region only if the two really overlap or touch: a gap between two user
           arrays is not the user's memory (it may be allocator metadata) and must be neither read nor
           written back */
        for (j = 0; j < nreg; ++j) {
            if (a <= reg[j].hi && b >= reg[j].lo) {
                if (a < reg[j].lo) reg[j].lo = a;
                if (b > reg[j].hi) reg[j].hi = b;
                break;
            }
        }
        if (j == nreg) { reg[nreg].lo = a; reg[nreg].hi = b; reg[nreg].has_in = reg[nreg].has_out = 0; ++nreg; }
        map[i] = j;
        if (i < 2) reg[j].has_in = 1; else reg[j].has_out = 1;
    }
    /* regions may have become overlapping after growth: merge again */
    for (i = 0; i < nreg; ++i)
        for (j = i + 1; j < nreg; ++j)
            if (reg[j].lo <= reg[i].hi && reg[j].hi >= reg[i].lo) {
                int k;
                if (reg[j].lo < reg[i].lo) reg[i].lo = reg[j].lo;
                if (reg[j].hi > reg[i].hi) reg[i].hi = reg[j].hi;
                reg[i].has_in |= reg[j].has_in; reg[i].has_out |= reg[j].has_out;
                for (k = 0; k < 4; ++k) { if (map[k] == j) map[k] = i; else if (map[k] > j) map[k]--; }
                for (k = j; k + 1 < nreg; ++k) reg[k] = reg[k + 1];
                --nreg; j = i;
            }
    {
        /* is the output dense?  then output-only regions need no upload */
        size_t out_bytes = 0;
        int dense;
        for (i = 0; i < nreg; ++i) if (reg[i].has_out && !reg[i].has_in) out_bytes += (size_t)(reg[i].hi - reg[i].lo);
        dense = (out_bytes == (size_t)written_reals * rs);
        for (i = 0; i < nreg; ++i) {
            size_t bytes = (size_t)(reg[i].hi - reg[i].lo);
            if (p->stage_cap[i] < bytes) {
                b2d_free(p->stage_dev[i]);
                p->stage_dev[i] = b2d_malloc(bytes);
                p->stage_cap[i] = p->stage_dev[i] ? bytes : 0;
                if (!p->stage_dev[i]) { fprintf(stderr, "fftw3_b200: staging alloc failed: %s\n", b2d_last_error()); return -1; }
            }
            if (reg[i].has_in || !dense)
                rc |= b2d_memcpy_h2d(p->stage_dev[i], reg[i].lo, bytes);
        }
    }
    for (i = 0; i < 4; ++i)
        dev_user[i] = (map[i] >= 0) ? (char *)p->stage_dev[map[i]] + ((char *)user[i] - reg[map[i]].lo) : NULL;
    if (!rc) rc = run_steps(p, dev_user);
    for (i = 0; i < nreg && !rc; ++i)
        if (reg[i].has_out) rc |= (async ? b2d_memcpy_d2h_async : b2d_memcpy_d2h)(reg[i].lo, p->stage_dev[i], (size_t)(reg[i].hi - reg[i].lo));
    if (!rc && !async) rc = b2d_sync();
    if (rc) fprintf(stderr, "fftw3_b200: execute failed: %s\n", b2d_last_error());
    return rc;
}

/* batched problem on host arrays: chunk c goes through chunk plan c % 3 on stream c % 3 (upload, passes, download in
   stream order), so consecutive chunks overlap each other's transfers and passes; every chunk plan has its own
   staging and scratch buffers, and a stream reuses them only after its previous chunk has left */
static int execute_host_pipelined(b2_plan *p, void *const user[4])
{
    size_t rs = real_size(p->prob.prec);
    int c, k, rc = 0;
    for (c = 0; c < p->pipe_chunks && !rc; ++c) {
        void *u[4], *st = b2d_pipe_stream(c % 3), *prev = NULL;
        for (k = 0; k < 4; ++k)
            u[k] = user[k] ? (char *)user[k] + (int64_t)c * (k < 2 ? p->pipe_in_off : p->pipe_out_off) * (int64_t)rs : NULL;
        if (st) prev = b2d_push_stream(st);
        rc = execute_host(p->pipe[c % 3], u, st != NULL);
        if (st) b2d_pop_stream(prev);
    }
    for (k = 0; k < 3; ++k) {
        void *st = b2d_pipe_stream(k), *prev;
        if (!st) continue;
        prev = b2d_push_stream(st);
        rc |= b2d_sync();
        b2d_pop_stream(prev);
    }
    return rc;
}

/* nosync: return after enqueueing (device pointers only).  fftw_execute* may be called from several
   threads on the same plan (doc/threads.texi:225-270): a plan that keeps intermediate data in its own
   scratch (or staging) buffers serialises its executes on the plan mutex -- held until the passes are
   complete, or, when only enqueueing, until they are all in the stream (stream order then keeps two
   executes of one plan apart). */
void b2_execute_ex(b2_plan *p, void *in0, void *in1, void *out0, void *out1, int nosync)
{
    void *user[4];
    int dev, rc, i, shared = 0;
    pthread_mutex_t *m;
    if (!p || p->is_nop) return;
    m = (pthread_mutex_t *)p->lock;
    user[0] = in0; user[1] = in1; user[2] = out0; user[3] = out1;
    dev = b2d_pointer_is_device(in0 ? in0 : out0);
    if (dev < 0) { fprintf(stderr, "fftw3_b200: no CUDA device: %s\n", b2d_last_error()); abort(); }
    if (dev == 1) {
        for (i = 0; i < B2_NSCRATCH; ++i) if (p->scratch[i]) shared = 1;
        for (i = 0; i < p->nsteps && !shared; ++i) if (p->steps[i].lane) shared = 1;   /* side streams are shared too */
        if (shared && m) pthread_mutex_lock(m);
        rc = run_steps(p, user);
        if (!rc && !nosync) rc = b2d_sync();
        if (shared && m) pthread_mutex_unlock(m);
        if (rc) { fprintf(stderr, "fftw3_b200: %s\n", b2d_last_error()); abort(); }
    } else {
        if (m) pthread_mutex_lock(m);
        rc = p->pipe_chunks > 1 ? execute_host_pipelined(p, user) : execute_host(p, user, 0);
        if (m) pthread_mutex_unlock(m);
        if (rc) abort();
    }
}

void b2_execute(b2_plan *p, void *in0, void *in1, void *out0, void *out1)
{
    b2_execute_ex(p, in0, in1, out0, out1, b2_async_mode);
}

void b2_plan_lock_init(b2_plan *p)
{
    pthread_mutex_t *m = (pthread_mutex_t *)malloc(sizeof *m);
    if (m) pthread_mutex_init(m, NULL);
    p->lock = m;
}

void b2_plan_lock_destroy(b2_plan *p)
{
    if (p->lock) { pthread_mutex_destroy((pthread_mutex_t *)p->lock); free(p->lock); p->lock = NULL; }
}
