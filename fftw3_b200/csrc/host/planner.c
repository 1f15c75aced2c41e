/* planner.c -- the GPU plan builder.
 *
 * Turns a canonical problem into a flat list of device passes.  It plays the
 * role of the reference's planner + solvers (kernel/planner.c:518-747,
 * dft/ct.c, dft/rank-geq2.c:42-52, dft/vrank-geq1.c:54-65, dft/bluestein.c,
 * rdft/rank-geq2-rdft2.c:40-66, rdft/ct-hc2c.c:59-82, reodft/ *.c) but the
 * search space is the GPU one: per pass, the radix factorisation, how many
 * transforms a CTA stages, threads per transform and the specialised-vs-generic
 * kernel; candidates are timed with CUDA events (FFTW_MEASURE and above) or
 * chosen by a closed-form heuristic (FFTW_ESTIMATE), and the winner is
 * remembered as wisdom keyed on the pass signature.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "b2_internal.h"

double b2_timelimit = -1.0;

/* working-set size for L2-blocked pass pairs (B200: 126 MB L2 over two dies);
   FFTW3_B200_L2_BLOCK_MB overrides, 0 disables */
#define B2_DEFAULT_L2_BLOCK_MB 0    /* measured on B200: per-group launches lose more to tails than L2 reuse wins */

/* ------------------------------------------------------------------ plan options */
static void opts_from_env(b2_plan_opts *o)
{
    const char *e;
    memset(o, 0, sizeof *o);
    o->l2_block_bytes = (size_t)B2_DEFAULT_L2_BLOCK_MB << 20;
    o->l2_lanes = 2; o->l2_keep = 4;
    o->split_mode = 0; o->split_bytes = (size_t)16 << 20; o->split_lanes = 3;
    if ((e = getenv("FFTW3_B200_L2_BLOCK_MB"))) o->l2_block_bytes = (size_t)atol(e) << 20;
    if ((e = getenv("FFTW3_B200_L2_BLOCK_KB"))) o->l2_block_bytes = (size_t)atol(e) << 10;      /* finer unit, for tests */
    if ((e = getenv("FFTW3_B200_L2_LANES"))) o->l2_lanes = atoi(e);
    if ((e = getenv("FFTW3_B200_L2_KEEP"))) o->l2_keep = atoi(e);
    if ((e = getenv("FFTW3_B200_L2_PAIR"))) o->l2_pair_outer = !strcmp(e, "outer");
    if ((e = getenv("FFTW3_B200_SPLIT"))) o->split_mode = atoi(e);
    if ((e = getenv("FFTW3_B200_SPLIT_MB"))) o->split_bytes = (size_t)atol(e) << 20;
    if ((e = getenv("FFTW3_B200_SPLIT_KB"))) o->split_bytes = (size_t)atol(e) << 10;
    if ((e = getenv("FFTW3_B200_SPLIT_LANES"))) o->split_lanes = atoi(e);
    if (getenv("FFTW3_B200_R2C_UNFUSED")) o->real_unfused |= 1;
    if (getenv("FFTW3_B200_C2R_UNFUSED")) o->real_unfused |= 2;
    if (getenv("FFTW3_B200_R2R_TRANSPOSES")) o->r2r_transposes = 1;
    if ((e = getenv("FFTW3_B200_PRIME"))) o->prime_mode = !strcmp(e, "rader") ? 1 : (!strcmp(e, "bluestein") ? 2 : 0);
}

static int opts_pinned_by_env(void)
{
    static const char *names[] = { "FFTW3_B200_L2_BLOCK_MB", "FFTW3_B200_L2_BLOCK_KB", "FFTW3_B200_L2_LANES",
        "FFTW3_B200_L2_KEEP", "FFTW3_B200_L2_PAIR", "FFTW3_B200_SPLIT", "FFTW3_B200_SPLIT_MB", "FFTW3_B200_SPLIT_KB",
        "FFTW3_B200_SPLIT_LANES", "FFTW3_B200_FORCE_VARIANT", "FFTW3_B200_R2C_UNFUSED", "FFTW3_B200_C2R_UNFUSED",
        "FFTW3_B200_R2R_TRANSPOSES", "FFTW3_B200_R2R_UNFUSED", "FFTW3_B200_PRIME" };
    size_t i;
    for (i = 0; i < sizeof names / sizeof names[0]; ++i) if (getenv(names[i])) return 1;
    return 0;
}

/* planning clock: fftw_set_timelimit (api/apiplan.c:92-136, kernel/planner.c:493-516) bounds the time
   spent MEASURING; once it has run out every remaining choice is made by the estimator */
#include <time.h>
static double g_plan_t0 = 0.0;
static double now_seconds(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static int time_is_up(void)
{
    return b2_timelimit >= 0.0 && now_seconds() - g_plan_t0 >= b2_timelimit;
}

/* ------------------------------------------------------------------ helpers */
typedef struct {           /* where a complex (or real) line lives */
    b2_ref re, im;
    int64_t stride;        /* element stride, reals */
} b2_view;

static size_t real_size(int prec) { return prec == B2D_F32 ? 4 : 8; }

static b2_step *new_step(b2_plan *p, b2_step_kind kind)
{
    b2_step *s;
    if (p->nsteps == p->cap) {
        int ncap = p->cap ? 2 * p->cap : 8;
        b2_step *ns = (b2_step *)realloc(p->steps, (size_t)ncap * sizeof(b2_step));
        if (!ns) return NULL;
        p->steps = ns; p->cap = ncap;
    }
    s = &p->steps[p->nsteps++];
    memset(s, 0, sizeof *s);
    s->kind = kind;
    return s;
}

static const void *plan_table(b2_plan *p, int prec, int kind, int64_t n, int64_t aux)
{
    b2_table *t = b2_table_get(prec, kind, n, aux);
    if (!t) return NULL;
    if (p->ntables == p->tcap) {
        int ncap = p->tcap ? 2 * p->tcap : 8;
        b2_table **nt = (b2_table **)realloc(p->tables, (size_t)ncap * sizeof(b2_table *));
        if (!nt) { b2_table_release(t); return NULL; }
        p->tables = nt; p->tcap = ncap;
    }
    p->tables[p->ntables++] = t;
    return t->dev;
}

static void need_scratch(b2_plan *p, int slot, size_t bytes)
{
    if (bytes > p->scratch_bytes[slot]) p->scratch_bytes[slot] = bytes;
}

static b2_ref mkref(int buf, int64_t off) { b2_ref r; r.buf = buf; r.off = off; return r; }

/* --------------------------------------------------------- radix selection */
int64_t b2_max_single_pass(int prec)
{
    return prec == B2D_F32 ? 8192 : 4096;
}

static int single_pass_fits(int64_t n, int prec)
{
    /* generic kernel: two padded rows of n complex + offsets must fit 227 KB */
    int64_t pitch = n + (n >> 4) + 9;
    size_t esz = 2 * real_size(prec);
    return (size_t)(2 * pitch) * esz + 256 <= (size_t)232448;
}

/* Factor n into supported radices.  variant selects among orderings/groupings;
   returns the number of stages, 0 if n has a prime factor > 13 or variant is
   out of range. */
int b2_factorize(int64_t n, int prec, int variant, int *radix)
{
    static const int odd[] = { 13, 11, 7, 5, 3 };
    int cnt = 0, i, a = 0, n3 = 0;
    int tmp[64];
    int64_t m = n;
    (void)prec;
    if (n < 1) return 0;
    if (n == 1) return (variant == 0) ? -1 : 0;   /* -1: zero stages, valid */
    while (m % 2 == 0) { m /= 2; ++a; }
    for (i = 0; i < 5; ++i)
        while (m % odd[i] == 0) {
            m /= odd[i];
            if (odd[i] == 3) ++n3; else tmp[cnt++] = odd[i];
            if (cnt > 40) return 0;
        }
    if (m != 1) return 0;
    while (n3 >= 2) { tmp[cnt++] = 9; n3 -= 2; }
    if (n3) tmp[cnt++] = 3;
    /* power-of-two part 2^a */
    {
        int q = a / 4, r = a % 4;
        switch (variant) {
        case 0:    /* as many radix-16 as possible, remainder as one stage */
            for (i = 0; i < q; ++i) tmp[cnt++] = 16;
            if (r == 1) {
                if (q > 0) { tmp[cnt - 1] = 8; tmp[cnt++] = 4; } else tmp[cnt++] = 2;
            } else if (r == 2) tmp[cnt++] = 4;
            else if (r == 3) tmp[cnt++] = 8;
            break;
        case 1:    /* radix-8 flavoured */
            q = a / 3; r = a % 3;
            for (i = 0; i < q; ++i) tmp[cnt++] = 8;
            if (r == 1) { if (q > 0) tmp[cnt - 1] = 16; else tmp[cnt++] = 2; }
            else if (r == 2) tmp[cnt++] = 4;
            break;
        case 2:    /* radix-4 flavoured */
            for (i = 0; i < a / 2; ++i) tmp[cnt++] = 4;
            if (a % 2) tmp[cnt++] = 2;
            break;
        default:
            return 0;
        }
    }
    if (cnt > B2D_MAX_STAGES) return 0;
    /* descending: the first (twiddle-free) stage gets the largest radix */
    {
        int j;
        for (i = 0; i < cnt; ++i)
            for (j = i + 1; j < cnt; ++j)
                if (tmp[j] > tmp[i]) { int t = tmp[i]; tmp[i] = tmp[j]; tmp[j] = t; }
    }
    for (i = 0; i < cnt; ++i) radix[i] = tmp[i];
    return cnt;
}

static int64_t next_pow2(int64_t n) { int64_t m = 1; while (m < n) m <<= 1; return m; }

/* ------------------------------------------------ single shared-memory pass */
typedef struct {
    int pre_op, post_op;
    int n_in, n_out;          /* 0 = n */
    int64_t big_n, tw4_split; /* STORE_TWIDDLE4: enclosing size and lo-table length */
    int cache;                /* L2 residency hints (b2d_fft_pass.cache) */
    int r2r_kind;             /* LOAD_R2R / STORE_R2R */
    int64_t idx_mul;          /* four-step halves: logical index rule for HERMCONJ / TRUNC (b2d_fft_pass.idx_mul) */
    int force_kernel;         /* pass shapes served by exactly one specialised kernel (STORE_R2C_SPLIT): its code */
    int64_t tw4_off;          /* STORE_TWIDDLE4: global index of batch column 0 (b2d_fft_pass.tw4_off) */
    int64_t merge_n;          /* LOAD_C2R_MERGE: the real length n = 2m whose roots of unity the merge uses */
} b2_ops;

static void fill_geometry(b2d_fft_pass *f, int variant)
{
    /* heuristic CTA shape; `variant` perturbs it for the measuring planner:
       variant = fvar + 3 * tvar  (fvar: factorisation, tvar: tile size class) */
    int tvar = variant / 3;
    int rmax = 1, i, tpx, tpb;
    int col = f->load_col || f->store_col;
    size_t esz = 2 * real_size(f->prec);
    int64_t pitch = f->n + (f->n >> 4) + 9;
    size_t per_xform = (size_t)(2 * pitch) * esz + 24;
    size_t budget = 110 * 1024;            /* aim for two CTAs per SM */
    for (i = 0; i < f->nstages; ++i) if (f->radix[i] > rmax) rmax = f->radix[i];
    tpx = f->n / rmax;
    if (tpx < 1) tpx = 1;
    if (tpx > 256) tpx = 256;
    if (col) tpb = (f->prec == B2D_F32) ? 16 : 8;
    else { tpb = 128 / tpx; if (tpb < 1) tpb = 1; }
    if (tvar == 1) tpb *= 2;
    else if (tvar == 2) { tpb /= 2; if (tpb < 1) tpb = 1; }
    else if (tvar == 3) { tpb *= 4; budget = 220 * 1024; }
    while (tpb > 1 && (size_t)tpb * per_xform > budget) tpb /= 2;
    if ((size_t)tpb * per_xform > (size_t)232448 - 64) tpb = 1;
    if (f->bn[0] < tpb) tpb = (int)(f->bn[0] > 0 ? f->bn[0] : 1);
    while (tpb * tpx > 1024 && tpx > 1) tpx /= 2;
    while (tpb * tpx > 1024 && tpb > 1) tpb /= 2;
    while (tpb * tpx < 64 && tpx < f->n && tpx < 256) tpx *= 2;
    f->tpb = tpb;
    f->tpx = tpx;
}

#define NVARIANTS 12          /* generic-kernel variants: factorisation x tile class */
#define NWARP 3               /* 32 x 32 kernels: rows (warp per transform), strided with 64 B / 128 B segments */
#define NFAST 42              /* specialised-kernel variants: tile width 1,2,4,8,16,32 x flavor 0..6
                                 (flavor 0 plain, 1 register-capped, 4-6 L2 prefetch-size loads for narrow COL tiles) */

static int configure_variant(b2d_fft_pass *f, int variant)
{
    int ns;
    f->kernel = 0;
    if (variant >= NVARIANTS + NFAST && variant < NVARIANTS + NFAST + NWARP) {
        /* 32 x 32 kernels for 1024-point lines (device/fft_warp.cuh): warp-per-transform for contiguous lines,
           one-exchange CTA kernels with 64- or 128-byte row segments for strided ones */
        int k = variant - (NVARIANTS + NFAST);
        int esz = (int)(2 * real_size(f->prec));
        int code = k == 0 ? 3001 : 3100 + (k == 1 ? 64 : 128) / esz;
        if (!b2d_fast_available(f, code)) return -1;
        ns = b2_factorize(f->n, f->prec, 0, f->radix);
        if (ns == 0) return -1;
        f->nstages = ns < 0 ? 0 : ns;
        fill_geometry(f, 0);            /* generic geometry stays configured: fallback for misaligned new arrays */
        f->kernel = code;
        return 0;
    }
    if (variant >= NVARIANTS + NFAST + NWARP) return -1;
    if (variant >= NVARIANTS) {
        int tpb = 1 << ((variant - NVARIANTS) % 6);
        int flavor = (variant - NVARIANTS) / 6;
        int code;
        if (flavor == 2 || flavor == 3) return -1;        /* derived from the pass shape below, not selectable */
        if (f->npeer && f->peer_rows > 0) { if (flavor) return -1; flavor = 8; }      /* row-split peer stores */
        else
        /* pass shapes with their own specialised flavour (see device/fft_fast.cuh) */
        if (f->pre_op == B2D_LOAD_C2R_MERGE) {
            if (flavor || f->post_op != B2D_STORE_TWIDDLE4 || !f->load_col || !f->store_col) return -1;
            flavor = 11;
        }
        else if (f->post_op == B2D_STORE_TWIDDLE4 && f->load_col && f->store_col) { if (flavor) return -1; flavor = 2; }
        else if (f->pre_op == B2D_LOAD_R2R && f->post_op == B2D_STORE_R2R) { if (flavor) return -1; flavor = 9; }
        else if (!f->load_col && f->store_col) { if (flavor) return -1; flavor = 3; }
        else if (f->bluestein) { if (flavor) return -1; flavor = 7; }
        code = ((f->load_col) ? 1000 : 0) + 100 * flavor + tpb;
        if (flavor == 11) code = 2100 + tpb;
        if (!b2d_fast_available(f, code)) return -1;
        /* generic geometry stays configured: it is the fallback for misaligned new arrays */
        ns = b2_factorize(f->n, f->prec, 0, f->radix);
        if (ns == 0) return -1;
        f->nstages = ns < 0 ? 0 : ns;
        fill_geometry(f, 0);
        f->kernel = code;
        return 0;
    }
    ns = b2_factorize(f->n, f->prec, variant % 3, f->radix);
    if (ns == 0) return -1;
    f->nstages = ns < 0 ? 0 : ns;
    fill_geometry(f, variant);
    /* twiddle table in shared memory when it fits next to the data (and leaves room for a second CTA
       when the data alone would) */
    f->tw_smem = 0;
    {
        size_t base = b2d_fft_pass_smem(f), tws = (size_t)f->n * 2 * real_size(f->prec);
        size_t cap = b2d_max_smem_per_block();
        if (base + tws <= cap && (base > cap / 2 || base + tws <= cap / 2)) f->tw_smem = 1;
    }
    if (b2d_fft_pass_smem(f) > b2d_max_smem_per_block()) return -1;
    return 0;
}

/* closed-form choice (FFTW_ESTIMATE): a specialised kernel when one exists.  Strided (COL) passes
   measured on B200 (profiles/r01_*): 64-byte tiles with L2::256B loads win while a pencil's stride stays
   inside a few 2 MiB pages (two CTAs per SM overlap load and compute), 128-byte tiles win beyond that
   (every row of the tile is in a page of its own). */
static int estimate_variant(b2d_fft_pass *f)
{
    static const int col_pref[] = { 3, 4, 2, 5 }, row_pref[] = { 1, 2, 0, 3, 4, 5 };
    int col = f->load_col || f->store_col, i;       /* wide tiles whenever a side is strided */
    const int *pref = col ? col_pref : row_pref;
    int npref = col ? 4 : 6;
    if (!col) {
        /* contiguous 1024-point lines: the warp-per-transform kernel (measured 6.1 TB/s vs 5.3 for the
           block-cooperative one, profiles/r02_row_experiment.log) */
        b2d_fft_pass t = *f;
        if (!configure_variant(&t, NVARIANTS + NFAST)) return NVARIANTS + NFAST;
    }
    if (f->load_col && f->store_col && !f->pre_op && !f->post_op && !f->npeer) {
        size_t esz = 2 * real_size(f->prec);
        int64_t stride_bytes = llabs(f->is) * (int64_t)real_size(f->prec);
        int want = stride_bytes >= (1 << 20) ? 128 : 64, lg = 0, v;
        while (((size_t)1 << lg) * esz < (size_t)want) ++lg;
        v = NVARIANTS + ((want == 64 && !(f->cache & 1)) ? 6 * 6 : 0) + lg;
        { b2d_fft_pass t = *f; if (!configure_variant(&t, v)) return v; }
    }
    if (!f->load_col && (f->pre_op & B2D_LOAD_R2R) && f->prec == B2D_F64) {
        /* double r2r lines so long that two transforms per CTA leave room for one CTA per SM only: one transform
           per CTA, two resident CTAs overlap each other's phases (160 vs 168 us per 4096 x 4096 pass, 182 vs 204
           with transposed stores, profiles/r02_c5b_pieces.txt) */
        int64_t pitch = f->n + (f->n >> 4) + 9;
        if ((size_t)(2 * pitch) * 16 > 113 * 1024) {
            b2d_fft_pass t = *f;
            if (!configure_variant(&t, NVARIANTS + 0)) return NVARIANTS + 0;
        }
    }
    if (!f->load_col && f->store_col && !f->pre_op && f->prec == B2D_F32) {
        /* contiguous lines stored transposed (second pass of a four-step): 128-byte store segments, i.e. 16 single
           precision lines per CTA (C2 c2r 1.60 vs 1.68 ms with 8, profiles/r02_c2_fused.log) */
        static const int s_pref[] = { 4, 3, 2, 5 };
        pref = s_pref; npref = 4;
    }
    if (!f->load_col && f->store_col && (f->pre_op & B2D_LOAD_R2R)) {
        /* long r2r lines stored transposed: the line kernels, widest tile first (more adjacent lines per store) */
        static const int t_pref[] = { 2, 1, 0 };
        pref = t_pref; npref = 3;
    }
    for (i = 0; i < npref; ++i) {
        b2d_fft_pass t = *f;
        if (!configure_variant(&t, NVARIANTS + pref[i])) return NVARIANTS + pref[i];
    }
    return 0;
}

/* offsets reached by a pass on its input / output side (reals) */
static void pass_span(const b2d_fft_pass *f, int out, int64_t *lo, int64_t *hi)
{
    int i;
    int64_t mn = 0, mx = 0, s = out ? f->os : f->is;
    int64_t len = out ? (f->n_out ? f->n_out : f->n) : (f->n_in ? f->n_in : f->n);
    if (f->idx_mul) len = f->n;      /* half of a four-step line: n_in / n_out are lengths of the whole line */
    int64_t e = (len - 1) * s;
    if (e < 0) mn += e; else mx += e;
    if (f->r2r_pair) { e = out ? f->pair_os : f->pair_is; if (e < 0) mn += e; else mx += e; }
    if (!out && (f->pre_op & B2D_LOAD_C2R_MERGE)) {      /* one element past the m the pass walks: X[m] */
        e = f->idx_mul ? f->bis[0] : f->is;
        if (e < 0) mn += e; else mx += e;
    }
    for (i = 0; i < B2D_MAX_BATCH_DIMS; ++i) {
        int64_t bs = out ? f->bos[i] : f->bis[i];
        e = (f->bn[i] - 1) * bs;
        if (e < 0) mn += e; else mx += e;
    }
    *lo = mn; *hi = mx;
}

/* time one configured pass on scratch buffers; returns ms or <0 */
static double time_pass(b2d_fft_pass *f, int inplace, int64_t im_minus_re_in, int64_t im_minus_re_out)
{
    int64_t ilo, ihi, olo, ohi;
    size_t rs = real_size(f->prec);
    size_t ibytes, obytes;
    void *ibuf, *obuf;
    float ms = 0, best = 1e30f;
    int rep, reps;
    b2d_fft_pass g = *f;
    pass_span(f, 0, &ilo, &ihi);
    pass_span(f, 1, &olo, &ohi);
    ibytes = (size_t)(ihi - ilo + 4 + llabs(im_minus_re_in)) * rs;
    obytes = (size_t)(ohi - olo + 4 + llabs(im_minus_re_out)) * rs;
    ibuf = b2d_malloc(ibytes);
    if (!ibuf) return -1;
    if (inplace) { obuf = ibuf; if (obytes > ibytes) { b2d_free(ibuf); return -1; } }
    else { obuf = b2d_malloc(obytes); if (!obuf) { b2d_free(ibuf); return -1; } }
    b2d_memset(ibuf, 0, ibytes);
    if (!inplace) b2d_memset(obuf, 0, obytes);
    {
        char *ib = (char *)ibuf + (size_t)(-ilo + (im_minus_re_in < 0 ? -im_minus_re_in : 0)) * rs;
        char *ob = (char *)obuf + (size_t)(-olo + (im_minus_re_out < 0 ? -im_minus_re_out : 0)) * rs;
        g.in_re = ib; g.in_im = ib + im_minus_re_in * (int64_t)rs;
        g.out_re = ob; g.out_im = ob + im_minus_re_out * (int64_t)rs;
    }
    if (b2d_launch_fft_pass(&g) || b2d_sync()) { best = -1; goto done; }
    /* the reference's protocol (kernel/timer.c:142-181): repeat a batch of launches, doubling the batch until
       it lasts long enough to be measurable (1 ms of CUDA-event time here), and keep the minimum over the
       repeats -- 8 of them for short passes, 3 for passes that already take milliseconds */
    {
        int iters = 1, i;
        for (;;) {
            b2d_timer_start();
            for (i = 0; i < iters; ++i) if (b2d_launch_fft_pass(&g)) { best = -1; goto done; }
            if (b2d_timer_stop(&ms)) { best = -1; goto done; }
            if (ms >= 1.0f || iters >= 128) break;
            iters *= 2;
        }
        best = ms / (float)iters;
        reps = (ms >= 4.0f && iters == 1) ? 2 : 7;
        for (rep = 0; rep < reps; ++rep) {
            b2d_timer_start();
            for (i = 0; i < iters; ++i) if (b2d_launch_fft_pass(&g)) { best = -1; goto done; }
            if (b2d_timer_stop(&ms)) { best = -1; goto done; }
            if (ms / (float)iters < best) best = ms / (float)iters;
        }
    }
done:
    if (!inplace) b2d_free(obuf);
    b2d_free(ibuf);
    return best;
}

static unsigned patience_of(unsigned flags)
{
    if (flags & B2F_ESTIMATE) return 0;
    if (flags & B2F_EXHAUSTIVE) return 3;
    if (flags & B2F_PATIENT) return 2;
    return 1;
}

/* Emit one single-pass FFT step (n fits shared memory, smooth).  Batch dims
   already canonical (<= 3). */
static int emit_single(b2_plan *p, int prec, int64_t n, b2_view in, b2_view out,
                       const b2_dim *bd, int brank, b2_ops ops, int bluestein_m, const char *note)
{
    b2_step *s = new_step(p, STEP_FFT);
    b2d_fft_pass *f;
    int i, variant = 0, have = 0;
    unsigned pat = patience_of(p->prob.flags);
    int inplace;
    if (!s) return -1;
    f = &s->u.fft;
    f->prec = prec;
    /* bluestein_m > 0: Bluestein with padded length M; bluestein_m < 0: Rader (work length n - 1) */
    f->n = (int)(bluestein_m > 0 ? bluestein_m : (bluestein_m < 0 ? n - 1 : n));
    f->pre_op = ops.pre_op; f->post_op = ops.post_op;
    f->cache = ops.cache;
    f->idx_mul = ops.idx_mul;
    f->n_in = ops.n_in ? ops.n_in : (int)n;
    f->n_out = ops.n_out ? ops.n_out : (int)n;
    f->is = in.stride; f->os = out.stride;
    for (i = 0; i < B2D_MAX_BATCH_DIMS; ++i) { f->bn[i] = 1; f->bis[i] = 0; f->bos[i] = 0; }
    for (i = 0; i < brank; ++i) { f->bn[i] = bd[i].n; f->bis[i] = bd[i].is; f->bos[i] = bd[i].os; }
    f->load_col = (brank > 0 && llabs(bd[0].is) < llabs(in.stride)) ? 1 : 0;
    f->store_col = (brank > 0 && llabs(bd[0].os) < llabs(out.stride)) ? 1 : 0;
    f->scale = 1.0;
    f->tw = plan_table(p, prec, TAB_TWIDDLE, f->n, 0);
    if (!f->tw) return -1;
    if (bluestein_m < 0) {
        /* Rader (dft/rader.c:95-165): x permuted by generator powers, cyclic convolution of length n - 1 with
           the permuted roots of unity (two FFTs of that length in the same CTA), outputs permuted back */
        f->bluestein = 2;
        f->pre_op |= B2D_LOAD_RADER;
        f->post_op |= B2D_STORE_RADER;
        f->scale = 1.0 / (double)(n - 1);
        f->aux0 = plan_table(p, prec, TAB_RADER_PERM, n, 0);
        f->aux1 = plan_table(p, prec, TAB_RADER_B, n, 0);
        if (!f->aux0 || !f->aux1) return -1;
    } else if (bluestein_m) {
        f->bluestein = 1;
        f->pre_op |= B2D_LOAD_PAD | B2D_LOAD_CHIRP;
        f->post_op |= B2D_STORE_TRUNC | B2D_STORE_CHIRP_SCALE;
        f->scale = 1.0 / (double)bluestein_m;
        f->aux0 = plan_table(p, prec, TAB_CHIRP, n, 0);
        f->aux1 = plan_table(p, prec, TAB_BLUE_B, n, bluestein_m);
        if (!f->aux0 || !f->aux1) return -1;
    }
    if (f->pre_op & B2D_LOAD_R2R) {
        int k = ops.r2r_kind;
        f->r2r_kind = k;
        /* kinds with a real PRE sequence (R2HC DHT REDFT00 RODFT00 REDFT10 RODFT10): two neighbouring lines
           share one complex transform -- half the arithmetic */
        if ((k == 0 || k == 2 || k == 3 || k == 7 || k == 5 || k == 9) && brank >= 1 && bd[0].n >= 2 &&
            bd[0].n % 2 == 0 && !getenv("FFTW3_B200_R2R_UNPAIRED")) {
            f->r2r_pair = 1;
            f->pair_is = bd[0].is; f->pair_os = bd[0].os;
            f->bn[0] = bd[0].n / 2; f->bis[0] = 2 * bd[0].is; f->bos[0] = 2 * bd[0].os;
        }
        if (k == 4 || k == 5 || k == 6 || k == 8 || k == 9 || k == 10) {     /* types 2, 3, 4: quarter-wave table */
            f->aux0 = plan_table(p, prec, TAB_QUARTER, f->n_in, 0);
            if (!f->aux0) return -1;
        }
    }
    if (f->post_op & B2D_STORE_TWIDDLE4) {
        f->big_n = ops.big_n; f->aux_split = ops.tw4_split;
        f->tw4_off = ops.tw4_off;
        f->tw4_shift = -1;
        if ((ops.big_n & (ops.big_n - 1)) == 0 && (ops.tw4_split & (ops.tw4_split - 1)) == 0) {
            int sh = 0;
            while (((int64_t)1 << sh) < ops.tw4_split) ++sh;
            f->tw4_shift = sh;
        }
        f->aux0 = plan_table(p, prec, TAB_TW4_LO, ops.big_n, ops.tw4_split);
        f->aux1 = plan_table(p, prec, TAB_TW4_HI, ops.big_n, ops.tw4_split);
        if (!f->aux0 || !f->aux1) return -1;
    }
    if (f->post_op & B2D_STORE_R2C_SPLIT) {
        f->aux0 = plan_table(p, prec, TAB_R2C, ops.big_n, 0);      /* exp(-2 pi i q / n), q <= n / 2 */
        if (!f->aux0) return -1;
    }
    if (f->pre_op & B2D_LOAD_C2R_MERGE) {
        f->aux2 = plan_table(p, prec, TAB_R2C, ops.merge_n, 0);
        if (!f->aux2) return -1;
    }
    s->r[0] = in.re; s->r[1] = in.im; s->r[2] = out.re; s->r[3] = out.im;
    snprintf(s->note, sizeof s->note, "%s", note);
    if (ops.force_kernel) {
        int ns = b2_factorize(f->n, f->prec, 0, f->radix);
        if (ns == 0) return -1;
        f->nstages = ns < 0 ? 0 : ns;
        fill_geometry(f, 0);
        f->kernel = ops.force_kernel;
        if (!b2d_fast_available(f, f->kernel)) return -1;
        goto counted;
    }
    inplace = (in.re.buf == out.re.buf && in.re.off == out.re.off) ||
              (in.re.buf == BUF_IN0 && out.re.buf == BUF_OUT0 && p->inplace);

    /* choose the variant: wisdom -> measure -> heuristic */
    {
        b2_sig sig = b2_sig_of_pass(f, inplace);
        const char *force = getenv("FFTW3_B200_FORCE_VARIANT");   /* tests: pin a kernel variant */
        if (force) {
            b2d_fft_pass trial = *f;
            if (!configure_variant(&trial, atoi(force))) { variant = atoi(force); have = 2; }
        }
        if (!have && b2_wisdom_lookup(sig, pat, &variant)) have = 1;
        if (!have && (p->prob.flags & B2F_WISDOM_ONLY)) return -2;
        if (!have && pat >= 1 && !time_is_up()) {
            int v, nv = NVARIANTS + NFAST + NWARP, bestv = -1, timed_any = 0;
            double bestt = 1e30;
            int64_t dri = (in.im.buf == in.re.buf) ? in.im.off - in.re.off : 1;
            int64_t dro = (out.im.buf == out.re.buf) ? out.im.off - out.re.off : 1;
            if (f->pre_op & (B2D_LOAD_REAL | B2D_LOAD_R2R)) dri = 0;
            if (f->post_op & (B2D_STORE_REALPART | B2D_STORE_R2R)) dro = 0;
            /* split arrays living in different buffers: time as interleaved-adjacent */
            if (llabs(dri) > 64) dri = 1;
            if (llabs(dro) > 64) dro = 1;
            for (v = 0; v < nv; ++v) {
                double t;
                b2d_fft_pass trial = *f;
                if (v >= 6 && v < NVARIANTS && pat < 2) continue;   /* extra generic shapes: PATIENT only */
                if (configure_variant(&trial, v)) continue;
                if (timed_any && time_is_up()) break;     /* fftw_set_timelimit: keep the best so far */
                t = time_pass(&trial, inplace, dri, dro);
                timed_any = 1;
                if (getenv("FFTW3_B200_VERBOSE")) {
                    double bytes = 4.0 * real_size(prec) * (double)f->n * (double)(f->bn[0] * f->bn[1] * f->bn[2]);
                    fprintf(stderr, "[b200 planner] n=%d %s->%s batch=%lldx%lldx%lld variant %2d %s tile=%d: %.4f ms  %.0f GB/s\n",
                            f->n, f->load_col ? "col" : "row", f->store_col ? "col" : "row", (long long)f->bn[0],
                            (long long)f->bn[1], (long long)f->bn[2], v,
                            trial.kernel >= 3000 ? "warp32x32" : (trial.kernel ? "codelet" : "generic"),
                            trial.kernel ? trial.kernel % 100 : trial.tpb, t, t > 0 ? bytes / t / 1e6 : 0.0);
                }
                if (t >= 0 && t < bestt) { bestt = t; bestv = v; }
            }
            if (bestv >= 0) { variant = bestv; have = 1; p->cost += bestt; }
        }
        if (!have) { variant = estimate_variant(f); pat = 0; }      /* nothing was timed: remembered as an estimate */
        if (configure_variant(f, variant)) {
            if (configure_variant(f, 0)) return -1;
            variant = 0;
        }
        if (have != 2) b2_wisdom_store(sig, pat, variant);
    }
counted:
    /* op count estimate (reference convention is per-plan add/mul/fma) */
    {
        double nb = (double)(f->bn[0] * f->bn[1] * f->bn[2]);
        double lg = log2((double)(f->n > 1 ? f->n : 2));
        double reps = bluestein_m ? 2.0 : 1.0;
        p->est_flops_add += reps * nb * 3.0 * f->n * lg;
        p->est_flops_mul += reps * nb * 0.5 * f->n * lg;
        p->est_flops_fma += reps * nb * 1.0 * f->n * lg;
    }
    return 0;
}

/* iterate a canonical batch tensor of arbitrary rank: the 3 innermost dims go
   into the kernel, outer dims are looped here (rare) */
typedef int (*emit_fn)(b2_plan *p, void *ctx, const b2_dim *bd, int brank, int64_t ioff, int64_t ooff);

static int for_outer_dims(b2_plan *p, const b2_tensor *batch, emit_fn fn, void *ctx)
{
    int inner = batch->rnk < B2D_MAX_BATCH_DIMS ? batch->rnk : B2D_MAX_BATCH_DIMS;
    int64_t idx[B2_MAXRANK];
    int d, rc;
    if (batch->rnk <= B2D_MAX_BATCH_DIMS) return fn(p, ctx, batch->d, inner, 0, 0);
    for (d = 0; d < B2_MAXRANK; ++d) idx[d] = 0;
    for (;;) {
        int64_t io = 0, oo = 0;
        for (d = inner; d < batch->rnk; ++d) { io += idx[d] * batch->d[d].is; oo += idx[d] * batch->d[d].os; }
        rc = fn(p, ctx, batch->d, inner, io, oo);
        if (rc) return rc;
        for (d = inner; d < batch->rnk; ++d) {
            if (++idx[d] < batch->d[d].n) break;
            idx[d] = 0;
        }
        if (d == batch->rnk) break;
    }
    return 0;
}

static b2_view view_shift(b2_view v, int64_t off)
{
    v.re.off += off; v.im.off += off;
    return v;
}

/* real-op steps (defined further down) are also used by the large-Bluestein path */
typedef struct {
    int prec, op, n, m; int64_t xs; b2_ref x_re, x_im, y_re, y_im; int work_slot;
    int64_t wdist; const void *tw; int user_is_out; int64_t *line_base;
    const void *aux; int flags, n_lim; double scale;
} rop_ctx;
static int emit_realop(b2_plan *p, rop_ctx *c, const b2_tensor *wb);
static void dense_work_strides(b2_tensor *t, int64_t wdist);

/* ----------------------------------------------------- batched 1-D complex FFT */
typedef struct {
    int prec; int64_t n; b2_view in, out; b2_ops ops; int scratch_slot; const char *note;
} fft1d_ctx;

static int emit_fft1d(b2_plan *p, int prec, int64_t n, b2_view in, b2_view out,
                      const b2_tensor *batch_in, b2_ops ops, int scratch_slot, const char *note);

/* ---- strided transform as two register-only sub-passes through an L2-resident buffer (device/fft_split.cuh)
   Applicable to a plain c2c pass along a strided dimension whose pencils are contiguous interleaved complex
   numbers on both sides.  The pencils are cut into groups of FFTW3_B200_SPLIT_MB MiB (default 16); a group's
   phase A and phase B are consecutive launches on one lane, consecutive groups rotate over the lanes, and
   every lane owns its slice of scratch slot 1 -- small enough that the work data never leaves L2. */
static int split_radices(int64_t n, int prec, int *ra, int *rb)
{
    static const int cand[][2] = { {32, 32}, {16, 32}, {16, 16}, {8, 16}, {8, 8} };
    size_t i;
    for (i = 0; i < sizeof cand / sizeof cand[0]; ++i)
        if ((int64_t)cand[i][0] * cand[i][1] == n && b2d_split_supported(prec, cand[i][0], cand[i][1])) {
            *ra = cand[i][0]; *rb = cand[i][1];
            return 1;
        }
    return 0;
}

/* does the view address interleaved complex numbers (im = re +- 1 real), vector-aligned at plan time? */
static int view_interleaved(const b2_plan *p, b2_view v)
{
    int64_t rs = (int64_t)real_size(p->prob.prec);
    const char *re = NULL, *im = NULL;
    if (v.re.buf == v.im.buf) return llabs(v.im.off - v.re.off) == 1 && v.re.buf >= BUF_SCRATCH0 &&
                                     ((v.re.off < v.im.off ? v.re.off : v.im.off) % 2) == 0;
    if (v.re.buf == BUF_IN0 && v.im.buf == BUF_IN1) { re = (const char *)p->prob.in0; im = (const char *)p->prob.in1; }
    else if (v.re.buf == BUF_IN1 && v.im.buf == BUF_IN0) { re = (const char *)p->prob.in1; im = (const char *)p->prob.in0; }
    else if (v.re.buf == BUF_OUT0 && v.im.buf == BUF_OUT1) { re = (const char *)p->prob.out0; im = (const char *)p->prob.out1; }
    else if (v.re.buf == BUF_OUT1 && v.im.buf == BUF_OUT0) { re = (const char *)p->prob.out1; im = (const char *)p->prob.out0; }
    else return 0;
    if (!re || !im || v.re.off != v.im.off) return 0;
    if (im - re != rs && im - re != -rs) return 0;
    return ((uintptr_t)((im < re ? im : re) + v.re.off * rs) % (size_t)(2 * rs)) == 0;
}

static int split_wanted(const b2_plan *p, const fft1d_ctx *c, b2_view in, b2_view out, const b2_dim *bd, int brank,
                        int *ra, int *rb)
{
    int mode = p->opt.split_mode;
    int64_t rs = (int64_t)real_size(c->prec), sb_in, sb_out;
    if (!mode || brank < 1) return 0;
    if (c->ops.pre_op || c->ops.post_op || c->ops.cache) return 0;
    if (p->prob.flags & B2F_UNALIGNED) return 0;
    if (bd[0].is != 2 || bd[0].os != 2 || bd[0].n < 64) return 0;
    if ((in.stride & 1) || (out.stride & 1) || llabs(in.stride) <= 2 || llabs(out.stride) <= 2) return 0;
    if (!view_interleaved(p, in) || !view_interleaved(p, out)) return 0;
    if (brank >= 2 && ((bd[1].is & 1) || (bd[1].os & 1))) return 0;
    if (brank >= 3 && ((bd[2].is & 1) || (bd[2].os & 1))) return 0;
    if (!split_radices(c->n, c->prec, ra, rb)) return 0;
    sb_in = llabs(in.stride) * rs; sb_out = llabs(out.stride) * rs;
    if (mode == 1 && sb_in < (1 << 20) && sb_out < (1 << 20)) return 0;     /* rows share 2 MiB pages: one kernel does it */
    return 1;
}

static int emit_split(b2_plan *p, const fft1d_ctx *c, b2_view in, b2_view out, const b2_dim *bd, int brank, int ra, int rb)
{
    int64_t n = c->n, esz = 2 * (int64_t)real_size(c->prec);
    int64_t gbytes = (int64_t)p->opt.split_bytes;
    int nlanes = p->opt.split_lanes;
    int64_t pmax = gbytes / (n * esz), nc0 = bd[0].n, nb1 = brank >= 2 ? bd[1].n : 1, n2 = brank >= 3 ? bd[2].n : 1;
    int64_t cstep, bstep, i2, b0, c0, gi = 0;
    const void *tw = plan_table(p, c->prec, TAB_TWIDDLE, n, 0);
    if (!tw) return -1;
    if (nlanes < 1) nlanes = 1;
    if (nlanes > 6) nlanes = 6;
    pmax -= pmax % 128;
    if (pmax < 128) pmax = 128;
    if (nc0 > pmax) { cstep = pmax; bstep = 1; }
    else { cstep = nc0; bstep = pmax / nc0; if (bstep < 1) bstep = 1; }
    need_scratch(p, 1, (size_t)nlanes * (size_t)(cstep * bstep) * (size_t)n * (size_t)esz);
    for (i2 = 0; i2 < n2; ++i2)
        for (b0 = 0; b0 < nb1; b0 += bstep)
            for (c0 = 0; c0 < nc0; c0 += cstep, ++gi) {
                int64_t nc = nc0 - c0 < cstep ? nc0 - c0 : cstep, nb = nb1 - b0 < bstep ? nb1 - b0 : bstep;
                int lane = (int)(gi % nlanes), ph;
                for (ph = 0; ph < 2; ++ph) {
                    b2_step *s = new_step(p, STEP_SPLIT);
                    b2d_split_pass *sp;
                    b2_view v = ph ? out : in;
                    int64_t off;
                    if (!s) return -1;
                    sp = &s->u.split;
                    sp->prec = c->prec; sp->ra = ra; sp->rb = rb; sp->phase = ph;
                    sp->row_stride = v.stride;
                    sp->nc = nc; sp->nb = nb;
                    sp->bs = brank >= 2 ? (ph ? bd[1].os : bd[1].is) : 0;
                    sp->tw = tw;
                    off = 2 * c0 + b0 * sp->bs + (brank >= 3 ? i2 * (ph ? bd[2].os : bd[2].is) : 0);
                    s->r[0] = v.re; s->r[0].off += off;
                    s->r[1] = v.im; s->r[1].off += off;
                    s->r[4] = mkref(BUF_SCRATCH1, (int64_t)lane * cstep * bstep * n * 2);
                    s->lane = nlanes > 1 ? 1 + lane : 0;
                    snprintf(s->note, sizeof s->note, "split %s %dx%d", ph ? "B" : "A", ra, rb);
                }
            }
    {
        double nbt = (double)(nc0 * nb1 * n2), lg = log2((double)n);
        p->est_flops_add += nbt * 3.0 * n * lg; p->est_flops_mul += nbt * 0.5 * n * lg; p->est_flops_fma += nbt * 1.0 * n * lg;
    }
    return 0;
}

/* n = n1 * n2 with both halves smooth and one-pass sized: the n1 closest to sqrt(n) from below (or the one
   FFTW3_B200_FOURSTEP_N1 pins), -1 if there is none */
static int64_t two_factor_split(int64_t n, int prec)
{
    int64_t d, best = -1;
    int tmp[64];
    const char *fs = getenv("FFTW3_B200_FOURSTEP_N1");       /* tests / tuning: pin the split */
    for (d = 2; d * d <= n; ++d) {
        if (n % d) continue;
        if (!single_pass_fits(n / d, prec) || !single_pass_fits(d, prec)) continue;
        if (!b2_factorize(d, prec, 0, tmp) || !b2_factorize(n / d, prec, 0, tmp)) continue;
        if (d > best) best = d;
        if (fs && d == atol(fs)) return d;
    }
    return best;
}

static int fft1d_inner(b2_plan *p, void *vctx, const b2_dim *bd, int brank, int64_t ioff, int64_t ooff)
{
    fft1d_ctx *c = (fft1d_ctx *)vctx;
    b2_view in = view_shift(c->in, ioff), out = view_shift(c->out, ooff);
    int64_t n = c->n;
    int radix[64];
    int smooth = b2_factorize(n, c->prec, 0, radix) != 0;

    if ((c->ops.post_op & B2D_STORE_R2C_SPLIT) && (!smooth || single_pass_fits(n, c->prec))) return -4;   /* only on a four-step's second pass */
    /* ... and the c2r merge only on the load of a (two-pass) four-step's first pass */
    if ((c->ops.pre_op & B2D_LOAD_C2R_MERGE) &&
        (!smooth || single_pass_fits(n, c->prec) || two_factor_split(n, c->prec) < 0 || (in.stride & 1))) return -4;
    if (smooth && single_pass_fits(n, c->prec)) {
        int ra, rb;
        if (split_wanted(p, c, in, out, bd, brank, &ra, &rb)) return emit_split(p, c, in, out, bd, brank, ra, rb);
        return emit_single(p, c->prec, n, in, out, bd, brank, c->ops, 0, c->note);
    }

    /* index-dependent fused ops cannot follow a second level of digit reversal: such lines (odd smooth
       real transforms beyond ~1.6e7 points) take the chirp-z route, whose pre/post maps carry the ops */
    if (smooth && two_factor_split(n, c->prec) < 0 &&
        ((c->ops.pre_op & B2D_LOAD_HERMCONJ) || (c->ops.post_op & B2D_STORE_TRUNC))) smooth = 0;

    if (!smooth) {
        /* Bluestein (dft/bluestein.c:82-128): one CTA does chirp, FFT_M, x B, IFFT_M, chirp */
        int64_t m = next_pow2(2 * n - 1);
        /* Rader for primes whose n - 1 is smooth and fits one CTA: half the transform length of
           Bluestein.  Taken when no register-resident Bluestein kernel exists for M (those beat the
           generic kernel by more than the length ratio); FFTW3_B200_PRIME=rader|bluestein overrides. */
        {
            const char *force = p->opt.prime_mode == 1 ? "rader" : (p->opt.prime_mode == 2 ? "bluestein" : NULL);
            int rr[64];
            int can = !c->ops.pre_op && !c->ops.post_op && n >= 5 && n < 2000000000 && b2_is_prime(n) &&
                      b2_factorize(n - 1, c->prec, 0, rr) != 0 && single_pass_fits(n - 1, c->prec);
            int want = can;
            if (can && !(force && !strcmp(force, "rader"))) {
                b2d_fft_pass probe;
                memset(&probe, 0, sizeof probe);
                probe.prec = c->prec; probe.n = (int)m; probe.bluestein = 1;
                probe.pre_op = B2D_LOAD_PAD | B2D_LOAD_CHIRP; probe.post_op = B2D_STORE_TRUNC | B2D_STORE_CHIRP_SCALE;
                probe.is = in.stride; probe.os = out.stride;
                if (single_pass_fits(m, c->prec) && in.stride == 2 && out.stride == 2 &&
                    (b2d_fast_available(&probe, 701) || b2d_fast_available(&probe, 702))) want = 0;
            }
            if (force && !strcmp(force, "bluestein")) want = 0;
            if (want) return emit_single(p, c->prec, n, in, out, bd, brank, c->ops, -1, "rader");
        }
        if (single_pass_fits(m, c->prec))
            return emit_single(p, c->prec, n, in, out, bd, brank, c->ops, (int)m, "bluestein");
        /* padded length too long for one CTA: chirp.pad kernel, FFT_M (four-step), x B, FFT_M,
           chirp kernel -- the same five steps as dft/bluestein.c:82-128, through scratch */
        {
            b2_tensor ub;
            rop_ctx rc_;
            b2_view wv;
            b2_ops none;
            int i, rc;
            int64_t lines = 1;
            size_t esz = 2 * real_size(c->prec);
            if (c->ops.pre_op & (B2D_LOAD_PAD | B2D_LOAD_CHIRP)) return -1;
            if (c->ops.post_op & ~(B2D_STORE_REALPART | B2D_STORE_TRUNC)) return -1;
            for (i = 0; i < brank; ++i) lines *= bd[i].n;
            need_scratch(p, 3, (size_t)lines * (size_t)m * esz);
            memset(&none, 0, sizeof none);
            /* PRE: user line -> work [line][m] */
            b2_tensor_init(&ub, 0);
            for (i = 0; i < brank; ++i) { ub.d[i].n = bd[i].n; ub.d[i].is = bd[i].is; }
            ub.rnk = brank;
            dense_work_strides(&ub, m);
            memset(&rc_, 0, sizeof rc_);
            rc_.prec = c->prec; rc_.op = B2D_ROP_BLUE_PRE; rc_.n = (int)n; rc_.m = (int)m; rc_.xs = in.stride;
            rc_.x_re = in.re; rc_.x_im = in.im; rc_.work_slot = 3; rc_.wdist = m; rc_.user_is_out = 0;
            rc_.flags = c->ops.pre_op; rc_.n_lim = (int)n;
            rc_.tw = plan_table(p, c->prec, TAB_CHIRP, n, 0);
            rc_.aux = plan_table(p, c->prec, TAB_BLUE_B, n, m);
            if (!rc_.tw || !rc_.aux) return -1;
            rc = emit_realop(p, &rc_, &ub);
            if (rc) return rc;
            /* FFT_M, x B (conj), FFT_M on the work lines */
            wv.re = mkref(BUF_SCRATCH3, 0); wv.im = mkref(BUF_SCRATCH3, 1); wv.stride = 2;
            for (i = 0; i < 2; ++i) {
                b2_tensor fb;
                b2_tensor_init(&fb, 1);
                fb.d[0].n = lines; fb.d[0].is = 2 * m; fb.d[0].os = 2 * m;
                rc = emit_fft1d(p, c->prec, m, wv, wv, &fb, none, 1, "bluestein FFT_M");
                if (rc) return rc;
                if (i == 0) {
                    b2_tensor mb;
                    rop_ctx mid = rc_;
                    mid.op = B2D_ROP_BLUE_MID;
                    b2_tensor_init(&mb, 1);
                    mb.d[0].n = lines; mb.d[0].is = 0; mb.d[0].os = 2 * m;
                    rc = emit_realop(p, &mid, &mb);
                    if (rc) return rc;
                }
            }
            /* POST: work -> user line */
            for (i = 0; i < brank; ++i) ub.d[i].is = bd[i].os;
            rc_.op = B2D_ROP_BLUE_POST; rc_.xs = out.stride;
            rc_.y_re = out.re; rc_.y_im = out.im; rc_.user_is_out = 1;
            rc_.x_re = rc_.x_im = mkref(BUF_NONE, 0);
            rc_.flags = c->ops.post_op;
            rc_.n_lim = c->ops.n_out ? c->ops.n_out : (int)n;
            rc_.scale = 1.0 / (double)m;
            return emit_realop(p, &rc_, &ub);
        }
    }

    /* four-step: n = n1 * n2, pass A strided length-n1 FFTs + twiddle into scratch, pass B contiguous
       length-n2 FFTs with transposed store.  n2 itself may need the same treatment (n beyond the square of
       the one-pass limit): pass B is then planned recursively through the next scratch slot.  Fused real-data
       ops ride on the passes: LOAD_REAL / LOAD_HERMCONJ on the load of pass A, STORE_REALPART / STORE_TRUNC
       on the store of pass B (the index-dependent ones through the pass's logical-index rule, idx_mul). */
    if (c->ops.pre_op & (B2D_LOAD_PAD | B2D_LOAD_CHIRP | B2D_LOAD_R2R | B2D_LOAD_RADER)) return -1;
    if (c->ops.post_op & ~(B2D_STORE_REALPART | B2D_STORE_TRUNC | B2D_STORE_R2C_SPLIT)) return -1;
    if (brank > 2) {
        int64_t k;
        for (k = 0; k < bd[brank - 1].n; ++k) {
            int rc2 = fft1d_inner(p, vctx, bd, brank - 1, ioff + k * bd[brank - 1].is, ooff + k * bd[brank - 1].os);
            if (rc2) return rc2;
        }
        return 0;
    }
    {
        int64_t n1 = 0, n2 = 0, d, best = two_factor_split(n, c->prec);
        int64_t lines = 1, L;
        b2_dim ba[3];
        b2_view sv;
        b2_ops oa, ob;
        int i, rc, tmp[64], slot = c->scratch_slot, nested = 0, next_slot = -1, r2c_code = 0;
        size_t rs = real_size(c->prec);
        if (c->ops.post_op & B2D_STORE_R2C_SPLIT) {
            /* the split rides on pass B only through a dedicated kernel (fft_fast.cuh flavour 10): find its tile */
            int tpb, code = 0;
            if (best < 0 || (out.stride & 1)) return -4;
            for (tpb = 16; tpb >= 2 && !code; tpb >>= 1) {
                b2d_fft_pass probe;
                memset(&probe, 0, sizeof probe);
                probe.prec = c->prec; probe.n = (int)(n / best); probe.post_op = B2D_STORE_R2C_SPLIT;
                probe.store_col = 1; probe.is = 2; probe.os = best * out.stride;
                probe.bn[0] = best; probe.bis[0] = 2 * (n / best); probe.bos[0] = out.stride;
                probe.bn[1] = probe.bn[2] = 1;
                for (i = 0; i < brank; ++i) { probe.bis[i + 1] = 2 * n; probe.bos[i + 1] = bd[i].os; }
                if (b2d_fast_available(&probe, 2000 + tpb)) code = 2000 + tpb;
            }
            if (!code) return -4;
            r2c_code = code;
        }
        if (best < 0) {
            /* no two-factor split fits: peel off the largest one-pass factor and recurse on the rest */
            for (d = 2; d <= 16384; ++d) {
                if (n % d || !single_pass_fits(d, c->prec) || !b2_factorize(d, c->prec, 0, tmp)) continue;
                if (d > best) best = d;
            }
            if (best < 0) return -1;
            nested = 1;
            next_slot = (slot == 1) ? 4 : (slot == 4 ? 5 : -1);
            if (next_slot < 0) return -1;
        }
        n1 = best; n2 = n / best;        /* n1 <= n2: rows of pass B are the long contiguous ones */
        for (i = 0; i < brank; ++i) lines *= bd[i].n;
        need_scratch(p, slot, (size_t)lines * (size_t)n * 2 * rs);
        /* scratch layout [line][k1][j2], interleaved complex */
        sv.re = mkref(BUF_SCRATCH0 + slot, 0);
        sv.im = mkref(BUF_SCRATCH0 + slot, 1);
        /* pass A */
        ba[0].n = n2; ba[0].is = in.stride; ba[0].os = 2;
        {
            int64_t ld = 2 * n;
            for (i = 0; i < brank; ++i) { ba[i + 1].n = bd[i].n; ba[i + 1].is = bd[i].is; ba[i + 1].os = ld; ld *= bd[i].n; }
        }
        {
            b2_view ia = in, oa_v = sv;
            ia.stride = n2 * in.stride;
            oa_v.stride = 2 * n2;
            L = 1; while (L * L < n) L <<= 1;
            oa = c->ops; oa.post_op = B2D_STORE_TWIDDLE4; oa.n_out = 0;
            if (oa.pre_op & (B2D_LOAD_HERMCONJ | B2D_LOAD_C2R_MERGE)) { oa.n_in = (int)n; oa.idx_mul = n2; }
            else oa.n_in = 0;
            oa.big_n = n; oa.tw4_split = L;
            rc = emit_single(p, c->prec, n1, ia, oa_v, ba, brank + 1, oa, 0, "four-step A");
            if (rc) return rc;
        }
        /* pass B */
        {
            b2_view ib = sv, ob_v = out;
            b2_tensor bb;
            int64_t ld = 2 * n;
            ib.stride = 2;
            ob_v.stride = n1 * out.stride;
            memset(&ob, 0, sizeof ob);
            ob.post_op = c->ops.post_op;
            if (ob.post_op & B2D_STORE_TRUNC) { ob.n_out = c->ops.n_out; ob.idx_mul = n1; }
            if (ob.post_op & B2D_STORE_R2C_SPLIT) { ob.force_kernel = r2c_code; ob.big_n = c->ops.big_n; }
            b2_tensor_init(&bb, 0);
            bb.d[0].n = n1; bb.d[0].is = 2 * n2; bb.d[0].os = out.stride;
            for (i = 0; i < brank; ++i) { bb.d[i + 1].n = bd[i].n; bb.d[i + 1].is = ld; bb.d[i + 1].os = bd[i].os; ld *= bd[i].n; }
            bb.rnk = brank + 1;
            if (!nested) rc = emit_single(p, c->prec, n2, ib, ob_v, bb.d, brank + 1, ob, 0, "four-step B");
            else {
                /* the nested transform sorts and merges its own batch; its k1 dimension must stay apart from
                   the lines (different output strides), which sort_merge guarantees by checking both sides */
                fft1d_ctx c2 = *c;
                c2.n = n2; c2.in = ib; c2.out = ob_v; c2.ops = ob; c2.scratch_slot = next_slot; c2.note = "four-step B (nested)";
                b2_tensor_drop_unit(&bb);
                rc = for_outer_dims(p, &bb, fft1d_inner, &c2);
            }
            if (rc) return rc;
        }
    }
    return 0;
}

static int emit_fft1d(b2_plan *p, int prec, int64_t n, b2_view in, b2_view out,
                      const b2_tensor *batch_in, b2_ops ops, int scratch_slot, const char *note)
{
    b2_tensor batch = *batch_in;
    fft1d_ctx c;
    int s0 = p->nsteps, rc;
    if (batch.rnk == B2_RNK_MINFTY) return 0;
    b2_tensor_drop_unit(&batch);
    b2_tensor_sort_merge(&batch);
    if (b2_tensor_count(&batch) == 0) return 0;
    c.prec = prec; c.n = n; c.in = in; c.out = out; c.ops = ops; c.scratch_slot = scratch_slot; c.note = note;
    rc = for_outer_dims(p, &batch, fft1d_inner, &c);
    /* a pass reads what the previous pass wrote, possibly through several lanes: join them first */
    if (!rc && p->nsteps > s0 && !p->no_fence) p->steps[s0].fence = 1;
    return rc;
}

/* every b2_ref offset, user buffer or scratch, is in units of the real scalar type */

/* plain contiguous in-place FFT on a device buffer (used to build Bluestein's B table):
   an ordinary plan of this library, executed once */
int b2_run_contig_fft(int prec, int64_t n, void *dev)
{
    b2_problem q;
    b2_plan *pl;
    int rc;
    memset(&q, 0, sizeof q);
    q.prec = prec; q.kind = B2_C2C; q.flags = B2F_ESTIMATE;
    b2_tensor_init(&q.sz, 1); b2_tensor_init(&q.vecsz, 0);
    q.sz.d[0].n = n; q.sz.d[0].is = 2; q.sz.d[0].os = 2;
    q.in0 = dev; q.in1 = (char *)dev + real_size(prec);
    q.out0 = q.in0; q.out1 = q.in1;
    pl = b2_mkplan(&q);
    if (!pl) return -1;
    b2_execute(pl, q.in0, q.in1, q.out0, q.out1);
    rc = b2d_sync();
    b2_plan_destroy(pl);
    return rc;
}

/* ------------------------------------------------------------ copy (rank 0) */
typedef struct { int prec; b2_ref in, out; int elem; } copy_ctx;

static int copy_inner(b2_plan *p, void *vctx, const b2_dim *bd, int brank, int64_t ioff, int64_t ooff)
{
    copy_ctx *c = (copy_ctx *)vctx;
    b2_step *s = new_step(p, STEP_COPY);
    int i;
    if (!s) return -1;
    s->u.copy.prec = c->prec;
    s->u.copy.elem_reals = c->elem;
    s->u.copy.rank = brank;
    for (i = 0; i < 4; ++i) { s->u.copy.n[i] = 1; s->u.copy.is[i] = 0; s->u.copy.os[i] = 0; }
    for (i = 0; i < brank; ++i) { s->u.copy.n[i] = bd[i].n; s->u.copy.is[i] = bd[i].is; s->u.copy.os[i] = bd[i].os; }
    s->r[0] = c->in; s->r[0].off += ioff;
    s->r[1] = c->out; s->r[1].off += ooff;
    snprintf(s->note, sizeof s->note, "copy");
    return 0;
}

static int emit_copy(b2_plan *p, int prec, b2_ref in, b2_ref out, const b2_tensor *t, int elem)
{
    b2_tensor c = *t;
    copy_ctx ctx;
    if (c.rnk == B2_RNK_MINFTY) return 0;
    b2_tensor_drop_unit(&c);
    b2_tensor_sort_merge(&c);
    if (b2_tensor_count(&c) == 0) return 0;
    ctx.prec = prec; ctx.in = in; ctx.out = out; ctx.elem = elem;
    return for_outer_dims(p, &c, copy_inner, &ctx);
}

/* --------------------------------------------------------------- real ops */


static int rop_inner(b2_plan *p, void *vctx, const b2_dim *bd, int brank, int64_t ioff, int64_t ooff)
{
    rop_ctx *c = (rop_ctx *)vctx;
    b2_step *s = new_step(p, STEP_REALOP);
    b2d_realop *r;
    int i;
    int64_t lines = 1;
    if (!s) return -1;
    r = &s->u.rop;
    r->prec = c->prec; r->op = c->op; r->n = c->n; r->m = c->m; r->xs = c->xs;
    for (i = 0; i < B2D_MAX_BATCH_DIMS; ++i) { r->bn[i] = 1; r->bxs[i] = 0; }
    /* the tensor handed to us has user strides in .is and dense work strides in .os */
    for (i = 0; i < brank; ++i) { r->bn[i] = bd[i].n; r->bxs[i] = bd[i].is; lines *= bd[i].n; }
    r->wdist = c->wdist;
    r->tw = c->tw;
    r->aux = c->aux; r->flags = c->flags; r->n_lim = c->n_lim; r->scale = c->scale;
    s->r[0] = c->x_re; s->r[1] = c->x_im; s->r[2] = c->y_re; s->r[3] = c->y_im;
    if (c->user_is_out) { s->r[2].off += ioff; s->r[3].off += ioff; }
    else { s->r[0].off += ioff; s->r[1].off += ioff; }
    s->r[4] = mkref(BUF_SCRATCH0 + c->work_slot, ooff);   /* dense work strides are in reals */
    if (c->op == B2D_ROP_BLUE_MID) { s->r[0] = s->r[1] = s->r[2] = s->r[3] = mkref(BUF_NONE, 0); }
    snprintf(s->note, sizeof s->note, "realop %d", c->op);
    (void)lines;
    return 0;
}

/* user-side batch `ub` (strides in .is) gets dense work strides in .os (reals,
   2*wdist per line) in ascending order of the sorted user strides */
static void dense_work_strides(b2_tensor *t, int64_t wdist)
{
    int i;
    int64_t ld = 2 * wdist;
    for (i = 0; i < t->rnk; ++i) { t->d[i].os = ld; ld *= t->d[i].n; }
}

static int cmp_is(const void *a, const void *b)
{
    int64_t x = llabs(((const b2_dim *)a)->is), y = llabs(((const b2_dim *)b)->is);
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* Build the canonical user batch (sorted by user stride, unit dims dropped)
   with dense work strides.  Both the FFT passes and the real ops that share a
   work buffer use this same tensor so their line numbering agrees. */
static void make_work_batch(b2_tensor *t, int64_t wdist)
{
    b2_tensor_drop_unit(t);
    if (t->rnk > 1) qsort(t->d, (size_t)t->rnk, sizeof(b2_dim), cmp_is);
    dense_work_strides(t, wdist);
}

static int emit_realop(b2_plan *p, rop_ctx *c, const b2_tensor *wb)
{
    /* no merging here: the dense strides already make merged == unmerged numbering,
       but sort_merge orders by .os which is ascending by construction */
    b2_tensor t = *wb;
    if (t.rnk == B2_RNK_MINFTY || b2_tensor_count(&t) == 0) return 0;
    b2_tensor_sort_merge(&t);
    return for_outer_dims(p, &t, rop_inner, c);
}

/* ------------------------------------------------------------------ c2c */
static void other_dims(const b2_problem *q, int skip, int use_out_for_in, b2_tensor *t)
{
    /* batch = vecsz + all sz dims except `skip`; when use_out_for_in, the pass
       runs in place on the output so both sides use .os */
    int i;
    b2_tensor_init(t, 0);
    for (i = 0; i < q->sz.rnk; ++i) {
        if (i == skip) continue;
        t->d[t->rnk] = q->sz.d[i];
        if (use_out_for_in) t->d[t->rnk].is = q->sz.d[i].os;
        t->rnk++;
    }
    for (i = 0; i < q->vecsz.rnk; ++i) {
        t->d[t->rnk] = q->vecsz.d[i];
        if (use_out_for_in) t->d[t->rnk].is = q->vecsz.d[i].os;
        t->rnk++;
    }
}

static int plan_c2c(b2_plan *p)
{
    const b2_problem *q = &p->prob;
    b2_ops none;
    int d, first = 1, rc;
    memset(&none, 0, sizeof none);
    if (q->sz.rnk == 0) {
        /* rank-0: copy (rdft/rank0.c); re and im planes separately */
        int64_t di = (char *)q->in1 - (char *)q->in0, dout = (char *)q->out1 - (char *)q->out0;
        int64_t rs = (int64_t)real_size(q->prec);
        int even = 1, i;
        if (p->inplace) return 0;
        for (i = 0; i < q->vecsz.rnk; ++i) if ((q->vecsz.d[i].is | q->vecsz.d[i].os) & 1) even = 0;
        if (even && di == dout && (di == rs || di == -rs)) {
            /* interleaved on both sides: move whole complex numbers (128-bit for double) */
            int lo_in = di > 0 ? BUF_IN0 : BUF_IN1, lo_out = di > 0 ? BUF_OUT0 : BUF_OUT1;
            return emit_copy(p, q->prec, mkref(lo_in, 0), mkref(lo_out, 0), &q->vecsz, 2);
        }
        rc = emit_copy(p, q->prec, mkref(BUF_IN0, 0), mkref(BUF_OUT0, 0), &q->vecsz, 1);
        if (rc) return rc;
        return emit_copy(p, q->prec, mkref(BUF_IN1, 0), mkref(BUF_OUT1, 0), &q->vecsz, 1);
    }
    if (q->tw_big_n) {
        /* one strided rank-1 pass with the six-step's twiddle in its store (dist_api.c) */
        b2_tensor batch;
        b2_view in, out;
        b2_ops tw;
        int64_t L = 1;
        int radix[64];
        if (q->sz.rnk != 1 || !b2_factorize(q->sz.d[0].n, q->prec, 0, radix) || !single_pass_fits(q->sz.d[0].n, q->prec)) return -1;
        memset(&tw, 0, sizeof tw);
        while (L * L < q->tw_big_n) L <<= 1;
        tw.post_op = B2D_STORE_TWIDDLE4; tw.big_n = q->tw_big_n; tw.tw4_split = L; tw.tw4_off = q->tw_off;
        other_dims(q, 0, 0, &batch);
        in.re = mkref(BUF_IN0, 0); in.im = mkref(BUF_IN1, 0); in.stride = q->sz.d[0].is;
        out.re = mkref(BUF_OUT0, 0); out.im = mkref(BUF_OUT1, 0); out.stride = q->sz.d[0].os;
        return emit_fft1d(p, q->prec, q->sz.d[0].n, in, out, &batch, tw, 1, "dft + six-step twiddle");
    }
    /* L2-resident pass pairs.  The pass over the last (contiguous) dim and the pass over one other dim
       only couple elements that share every remaining index, so the two can be run group by group over a
       third dim: with a group small enough to stay in the 126 MB L2, the second pass reads what the first
       just wrote from L2 and the pair costs one HBM read and one HBM write instead of two of each ("fused
       multi-pass": the role of the reference's cache-oblivious rank-geq2 / buffered recursion,
       dft/rank-geq2.c:42-52, dft/buffered.c:41-69, on a cache shared by all SMs).  Consecutive groups go to
       different lanes (side streams, exec.c): they are independent, so the tail of one group's kernels is
       filled by the head of the next group's.  The first pass keeps its output in L2 (ordinary or
       evict_last stores), the second reads it with ordinary loads and streams its result out. */
    {
        int done[B2_MAXRANK];
        int i, last = q->sz.rnk - 1;
        for (i = 0; i < B2_MAXRANK; ++i) done[i] = 0;
        if (q->sz.rnk >= 2 && p->opt.l2_block_bytes > 0) {
            int pair = (p->opt.l2_pair_outer && q->sz.rnk >= 3) ? 0 : last - 1;
            int nlanes = p->opt.l2_lanes, keep = p->opt.l2_keep;
            int oi = -1, ovec = 0;
            int64_t best = 0, per = 2 * (int64_t)real_size(q->prec), nout, G;
            if (nlanes < 0) nlanes = 0;
            if (nlanes > 6) nlanes = 6;
            for (i = 0; i < q->sz.rnk; ++i)
                if (i != last && i != pair && llabs(q->sz.d[i].os) > best) { best = llabs(q->sz.d[i].os); oi = i; ovec = 0; }
            for (i = 0; i < q->vecsz.rnk; ++i)
                if (llabs(q->vecsz.d[i].os) > best) { best = llabs(q->vecsz.d[i].os); oi = i; ovec = 1; }
            if (oi >= 0) {
                const b2_dim *od = ovec ? &q->vecsz.d[oi] : &q->sz.d[oi];
                nout = od->n;
                for (i = 0; i < q->sz.rnk; ++i) if (ovec || i != oi) per *= q->sz.d[i].n;
                for (i = 0; i < q->vecsz.rnk; ++i) if (!ovec || i != oi) per *= q->vecsz.d[i].n;
                if (!p->inplace) per *= 2;
                G = (int64_t)p->opt.l2_block_bytes / (per > 0 ? per : 1);
                if (G >= 1 && G < nout && nout / G <= 8192) {
                    int64_t g0, gi = 0;
                    for (g0 = 0; g0 < nout; g0 += G, ++gi) {
                        int64_t cnt = (nout - g0 < G) ? nout - g0 : G;
                        int pass, s0 = p->nsteps;
                        p->no_fence = gi > 0;      /* groups are independent of each other */
                        for (pass = 0; pass < 2; ++pass) {
                            b2_problem qq = *q;
                            b2_tensor batch;
                            b2_view in, out;
                            int dd = pass ? pair : last;
                            b2_dim *md = ovec ? &qq.vecsz.d[oi] : &qq.sz.d[oi];
                            md->n = cnt;
                            other_dims(&qq, dd, pass, &batch);
                            out.re = mkref(BUF_OUT0, g0 * od->os); out.im = mkref(BUF_OUT1, g0 * od->os);
                            out.stride = q->sz.d[dd].os;
                            if (pass == 0) {
                                in.re = mkref(BUF_IN0, g0 * od->is); in.im = mkref(BUF_IN1, g0 * od->is);
                                in.stride = q->sz.d[dd].is;
                            } else in = out;
                            none.cache = pass ? 1 : keep;   /* first pass leaves its output in L2 for the second */
                            rc = emit_fft1d(p, q->prec, q->sz.d[dd].n, in, out, &batch, none, 1,
                                            pass ? "dft(in place, L2 group)" : "dft(L2 group)");
                            none.cache = 0;
                            if (rc) return rc;
                        }
                        p->no_fence = 0;
                        if (nlanes > 0) for (i = s0; i < p->nsteps; ++i) p->steps[i].lane = 1 + (int)(gi % nlanes);
                    }
                    first = 0;
                    done[last] = done[pair] = 1;
                }
            }
        }
        for (d = q->sz.rnk - 1; d >= 0; --d) {
            b2_tensor batch;
            b2_view in, out;
            if (done[d]) continue;
            other_dims(q, d, !first, &batch);
            out.re = mkref(BUF_OUT0, 0); out.im = mkref(BUF_OUT1, 0); out.stride = q->sz.d[d].os;
            if (first) { in.re = mkref(BUF_IN0, 0); in.im = mkref(BUF_IN1, 0); in.stride = q->sz.d[d].is; }
            else in = out;
            rc = emit_fft1d(p, q->prec, q->sz.d[d].n, in, out, &batch, none, 1, first ? "dft" : "dft(in place)");
            if (rc) return rc;
            first = 0;
        }
    }
    return 0;
}

/* In-place real transforms whose single-pass (odd n) kernel would read and
   write the same memory are only safe when every line owns its bytes: all
   outer/batch strides equal on both sides and at least one row long.  Other
   layouts (e.g. interleaved vectors) get their input staged through scratch,
   the counterpart of the reference's buffered solvers (rdft/buffered2.c). */
static int rows_self_contained(const b2_problem *q, int last)
{
    int i;
    int64_t n = q->sz.d[last].n;
    int64_t nr = (q->kind == B2_R2C) ? n : n / 2 + 1, nw = (q->kind == B2_R2C) ? n / 2 + 1 : n;
    int64_t ext_in = (nr - 1) * llabs(q->sz.d[last].is) + 2;
    int64_t ext_out = (nw - 1) * llabs(q->sz.d[last].os) + 2;
    int64_t ext = ext_in > ext_out ? ext_in : ext_out;
    for (i = 0; i < q->sz.rnk; ++i) {
        if (i == last) continue;
        if (q->sz.d[i].is != q->sz.d[i].os || llabs(q->sz.d[i].is) < ext) return 0;
    }
    for (i = 0; i < q->vecsz.rnk; ++i)
        if (q->vecsz.d[i].is != q->vecsz.d[i].os || llabs(q->vecsz.d[i].is) < ext) return 0;
    return 1;
}

/* copy the whole input of a real transform into scratch slot 2 (dense,
   row-major over [vecsz..., sz...]) and rewrite the problem's input strides */
static int stage_input(b2_plan *p, b2_problem *q, int last, int *src0, int *src1)
{
    b2_tensor t;
    int i, k = 0, rc;
    int cmplx = (q->kind == B2_C2R);
    int64_t ld = cmplx ? 2 : 1;
    int64_t nlast = cmplx ? q->sz.d[last].n / 2 + 1 : q->sz.d[last].n;
    b2_tensor_init(&t, 0);
    /* dense strides: last transform dim fastest, then the other sz dims, then vecsz */
    for (i = q->sz.rnk - 1; i >= 0; --i) {
        t.d[k].n = (i == last) ? nlast : q->sz.d[i].n;
        t.d[k].is = q->sz.d[i].is; t.d[k].os = ld;
        q->sz.d[i].is = ld;
        ld *= t.d[k].n; ++k;
    }
    for (i = q->vecsz.rnk - 1; i >= 0; --i) {
        t.d[k].n = q->vecsz.d[i].n;
        t.d[k].is = q->vecsz.d[i].is; t.d[k].os = ld;
        q->vecsz.d[i].is = ld;
        ld *= t.d[k].n; ++k;
    }
    t.rnk = k;
    need_scratch(p, 2, (size_t)ld * real_size(q->prec));
    rc = emit_copy(p, q->prec, mkref(BUF_IN0, 0), mkref(BUF_SCRATCH2, 0), &t, 1);
    if (rc) return rc;
    if (cmplx) {
        rc = emit_copy(p, q->prec, mkref(BUF_IN1, 0), mkref(BUF_SCRATCH2, 1), &t, 1);
        if (rc) return rc;
    }
    *src0 = BUF_SCRATCH2; *src1 = BUF_SCRATCH2;
    return 0;
}

/* ------------------------------------------------------------------ r2c */
static int plan_r2c(b2_plan *p)
{
    b2_problem qq = p->prob;
    const b2_problem *q = &qq;
    int last = q->sz.rnk - 1, d, rc;
    int64_t n, is, os;
    b2_tensor batch;
    b2_ops ops;
    b2_view in, out;
    int src0 = BUF_IN0, src1 = BUF_IN0;
    if (q->sz.rnk < 1) {
        /* rank 0 (rdft/rank0-rdft2.c): out = in + 0i for every element of the batch -- a length-1
           transform with a real load */
        memset(&ops, 0, sizeof ops);
        ops.pre_op = B2D_LOAD_REAL;
        in.re = in.im = mkref(BUF_IN0, 0); in.stride = 1;
        out.re = mkref(BUF_OUT0, 0); out.im = mkref(BUF_OUT1, 0); out.stride = 2;
        return emit_fft1d(p, q->prec, 1, in, out, &q->vecsz, ops, 1, "r2c rank 0");
    }
    n = q->sz.d[last].n;
    if (p->inplace && (n % 2) && !rows_self_contained(q, last)) {
        rc = stage_input(p, &qq, last, &src0, &src1);
        if (rc) return rc;
    }
    is = q->sz.d[last].is; os = q->sz.d[last].os;
    other_dims(q, last, 0, &batch);
    memset(&ops, 0, sizeof ops);
    out.re = mkref(BUF_OUT0, 0); out.im = mkref(BUF_OUT1, 0); out.stride = os;
    if (n % 2 == 0 && n >= 2) {
        int64_t m = n / 2;
        b2_tensor wb = batch, fb;
        rop_ctx c;
        b2_view wv;
        size_t esz = 2 * real_size(q->prec);
        int i;
        in.re = mkref(BUF_IN0, 0); in.im = mkref(BUF_IN0, is); in.stride = 2 * is;
        if (!(p->opt.real_unfused & 1) && !(q->flags & B2F_UNALIGNED) && view_interleaved(p, out) && src0 == BUF_IN0) {
            /* long lines (the half-size transform is a four-step): the split rides on the store of its second
               pass, so the line crosses HBM twice, not three times (rdft/ct-hc2c.c:146-273 fuses it the same way
               into the last twiddle codelet) */
            b2_ops fo;
            int s0 = p->nsteps;
            memset(&fo, 0, sizeof fo);
            fo.post_op = B2D_STORE_R2C_SPLIT; fo.big_n = n;
            rc = emit_fft1d(p, q->prec, m, in, out, &batch, fo, 1, "r2c half-size dft + split");
            if (rc == 0 && p->nsteps > s0) goto leading_dims;
            if (rc != -4 && rc != 0) return rc;
            p->nsteps = s0;
        }
        /* FFT_m of (even, odd) samples as (re, im): user -> work */
        make_work_batch(&wb, m);      /* .is = user real strides, .os = dense work */
        need_scratch(p, 0, (size_t)(b2_tensor_count(&wb) > 0 ? b2_tensor_count(&wb) : 1) * (size_t)m * esz);
        in.re = mkref(BUF_IN0, 0); in.im = mkref(BUF_IN0, is); in.stride = 2 * is;
        wv.re = mkref(BUF_SCRATCH0, 0); wv.im = mkref(BUF_SCRATCH0, 1); wv.stride = 2;
        rc = emit_fft1d(p, q->prec, m, in, wv, &wb, ops, 1, "r2c half-size dft");
        if (rc) return rc;
        /* split: work -> user complex; user strides are now the OUTPUT strides */
        fb = wb;
        {
            /* replace .is by output strides of the same dims, keeping order */
            b2_tensor ob;
            other_dims(q, last, 1, &ob);   /* .is == .os == output strides */
            b2_tensor_drop_unit(&ob);
            /* wb was sorted by input stride; re-create in the same order: match by position
               is fragile, so rebuild: sort a copy of (in,out) pairs by input stride */
            {
                b2_tensor pair;
                other_dims(q, last, 0, &pair);
                b2_tensor_drop_unit(&pair);
                if (pair.rnk > 1) qsort(pair.d, (size_t)pair.rnk, sizeof(b2_dim), cmp_is);
                for (i = 0; i < pair.rnk; ++i) { fb.d[i].n = pair.d[i].n; fb.d[i].is = pair.d[i].os; }
                fb.rnk = pair.rnk;
                dense_work_strides(&fb, m);
            }
        }
        memset(&c, 0, sizeof c);
        c.prec = q->prec; c.op = B2D_ROP_R2C_POST; c.n = (int)n; c.m = (int)m; c.xs = os;
        c.y_re = mkref(BUF_OUT0, 0); c.y_im = mkref(BUF_OUT1, 0);
        c.work_slot = 0; c.wdist = m; c.user_is_out = 1;
        c.tw = plan_table(p, q->prec, TAB_R2C, n, 0);
        if (!c.tw) return -1;
        rc = emit_realop(p, &c, &fb);
        if (rc) return rc;
    } else {
        in.re = mkref(src0, 0); in.im = mkref(src0, 0); in.stride = is;
        ops.pre_op = B2D_LOAD_REAL; ops.post_op = B2D_STORE_TRUNC; ops.n_out = (int)(n / 2 + 1);
        rc = emit_fft1d(p, q->prec, n, in, out, &batch, ops, 1, "r2c odd");
        if (rc) return rc;
    }
leading_dims:
    /* remaining dims: complex, in place on the output, last dim now n/2+1 long */
    memset(&ops, 0, sizeof ops);
    for (d = last - 1; d >= 0; --d) {
        b2_tensor b2;
        int i, k = 0;
        other_dims(q, d, 1, &b2);
        /* the entry that came from the last dim has n -> n/2+1 */
        for (i = 0; i < q->sz.rnk; ++i) {
            if (i == d) continue;
            if (i == last) b2.d[k].n = n / 2 + 1;
            ++k;
        }
        out.stride = q->sz.d[d].os;
        rc = emit_fft1d(p, q->prec, q->sz.d[d].n, out, out, &b2, ops, 1, "r2c outer dft");
        if (rc) return rc;
    }
    return 0;
}

/* ------------------------------------------------------------------ c2r */
static int plan_c2r(b2_plan *p)
{
    b2_problem qq = p->prob;
    const b2_problem *q = &qq;
    int last = q->sz.rnk - 1, d, rc;
    int64_t n, is, os;
    b2_tensor batch;
    b2_ops ops;
    b2_view in, out;
    int src_re = BUF_IN0, src_im = BUF_IN1;
    int64_t im_off = 0;
    if (q->sz.rnk < 1) {
        /* rank 0 (rdft/rank0-rdft2.c): out = Re(in) */
        memset(&ops, 0, sizeof ops);
        ops.post_op = B2D_STORE_REALPART;
        in.re = mkref(BUF_IN0, 0); in.im = mkref(BUF_IN1, 0); in.stride = 2;
        out.re = out.im = mkref(BUF_OUT0, 0); out.stride = 1;
        return emit_fft1d(p, q->prec, 1, in, out, &q->vecsz, ops, 1, "c2r rank 0");
    }
    n = q->sz.d[last].n;
    if (p->inplace && (n % 2) && !rows_self_contained(q, last)) {
        rc = stage_input(p, &qq, last, &src_re, &src_im);
        if (rc) return rc;
        im_off = 1;
    }
    is = q->sz.d[last].is; os = q->sz.d[last].os;
    memset(&ops, 0, sizeof ops);

    if (q->sz.rnk > 1) {
        /* leading dims: backward complex passes in place on the INPUT (this is why
           multi-dimensional c2r destroys its input: api/plan-many-dft-c2r.c:41-42) */
        if (!p->inplace) p->destroys_input = 1;
        for (d = 0; d < last; ++d) {
            b2_tensor b2;
            int i, k = 0;
            /* in place on input: both sides use .is */
            b2_tensor_init(&b2, 0);
            for (i = 0; i < q->sz.rnk; ++i) {
                if (i == d) continue;
                b2.d[k] = q->sz.d[i];
                b2.d[k].os = q->sz.d[i].is;
                if (i == last) b2.d[k].n = n / 2 + 1;
                ++k;
            }
            for (i = 0; i < q->vecsz.rnk; ++i) { b2.d[k] = q->vecsz.d[i]; b2.d[k].os = q->vecsz.d[i].is; ++k; }
            b2.rnk = k;
            /* backward = forward with re/im swapped */
            in.re = mkref(src_im, im_off); in.im = mkref(src_re, 0); in.stride = q->sz.d[d].is;
            rc = emit_fft1d(p, q->prec, q->sz.d[d].n, in, in, &b2, ops, 1, "c2r outer dft");
            if (rc) return rc;
        }
    }
    other_dims(q, last, 0, &batch);
    out.re = mkref(BUF_OUT0, 0); out.im = mkref(BUF_OUT0, 0); out.stride = os;
    if (n % 2 == 0 && n >= 2) {
        int64_t m = n / 2;
        b2_tensor wb = batch, fb;
        rop_ctx c;
        b2_view wv;
        size_t esz = 2 * real_size(q->prec);
        int i;
        {
            b2_view iv;
            iv.re = mkref(src_re, 0); iv.im = mkref(src_im, im_off); iv.stride = is;
            if (!(p->opt.real_unfused & 2) && !(q->flags & B2F_UNALIGNED) && src_re == BUF_IN0 && im_off == 0 &&
                view_interleaved(p, iv) && (const char *)p->prob.in1 - (const char *)p->prob.in0 == (ptrdiff_t)real_size(q->prec)) {
                /* long lines (the half-size transform is a four-step): the merge rides on the load of its first
                   pass, so the line crosses HBM twice, not three times -- the mirror image of the r2c split that
                   rides on the last store */
                b2_ops mo;
                b2_view ov;
                int s0 = p->nsteps;
                memset(&mo, 0, sizeof mo);
                mo.pre_op = B2D_LOAD_C2R_MERGE; mo.merge_n = n;
                ov.re = mkref(BUF_OUT0, os); ov.im = mkref(BUF_OUT0, 0); ov.stride = 2 * os;
                rc = emit_fft1d(p, q->prec, m, iv, ov, &batch, mo, 1, "c2r merge + half-size dft");
                if (rc == 0 && p->nsteps > s0) return 0;
                if (rc != -4 && rc != 0) return rc;
                p->nsteps = s0;
            }
        }
        make_work_batch(&wb, m);       /* sorted by INPUT (complex) strides */
        need_scratch(p, 0, (size_t)(b2_tensor_count(&wb) > 0 ? b2_tensor_count(&wb) : 1) * (size_t)m * esz);
        memset(&c, 0, sizeof c);
        c.prec = q->prec; c.op = B2D_ROP_C2R_PRE; c.n = (int)n; c.m = (int)m; c.xs = is;
        c.x_re = mkref(src_re, 0); c.x_im = mkref(src_im, im_off);
        c.work_slot = 0; c.wdist = m; c.user_is_out = 0;
        c.tw = plan_table(p, q->prec, TAB_R2C, n, 0);
        if (!c.tw) return -1;
        rc = emit_realop(p, &c, &wb);
        if (rc) return rc;
        /* forward FFT_m on the swapped work, scattered as (odd, even) reals */
        fb = wb;
        for (i = 0; i < fb.rnk; ++i) { int64_t t = fb.d[i].is; fb.d[i].is = fb.d[i].os; fb.d[i].os = t; }
        /* fb: .is = dense work strides, .os = must be the user's OUTPUT strides of the same dims */
        {
            b2_tensor pair;
            other_dims(q, last, 0, &pair);
            b2_tensor_drop_unit(&pair);
            if (pair.rnk > 1) qsort(pair.d, (size_t)pair.rnk, sizeof(b2_dim), cmp_is);
            for (i = 0; i < pair.rnk; ++i) fb.d[i].os = pair.d[i].os;
        }
        wv.re = mkref(BUF_SCRATCH0, 0); wv.im = mkref(BUF_SCRATCH0, 1); wv.stride = 2;
        out.re = mkref(BUF_OUT0, os); out.im = mkref(BUF_OUT0, 0); out.stride = 2 * os;
        rc = emit_fft1d(p, q->prec, m, wv, out, &fb, ops, 1, "c2r half-size dft");
        if (rc) return rc;
    } else {
        in.re = mkref(src_re, 0); in.im = mkref(src_im, im_off); in.stride = is;
        ops.pre_op = B2D_LOAD_HERMCONJ; ops.post_op = B2D_STORE_REALPART;
        ops.n_in = (int)n; ops.n_out = (int)n;
        rc = emit_fft1d(p, q->prec, n, in, out, &batch, ops, 1, "c2r odd");
        if (rc) return rc;
    }
    return 0;
}

/* ------------------------------------------------------------------ r2r */
static int r2r_work_len(int kind, int64_t n, int64_t *m)
{
    switch (kind) {
    case 3: if (n < 2) return -1; *m = 2 * (n - 1); return 0;     /* REDFT00 */
    case 7: *m = 2 * (n + 1); return 0;                           /* RODFT00 */
    case 6: case 10: *m = 2 * n; return 0;                        /* REDFT11, RODFT11 */
    default: *m = n; return 0;
    }
}

static int plan_r2r(b2_plan *p)
{
    const b2_problem *q = &p->prob;
    int d, first = 1, rc;
    b2_ops none;
    memset(&none, 0, sizeof none);
    if (q->sz.rnk == 0) {
        if (p->inplace) return 0;
        return emit_copy(p, q->prec, mkref(BUF_IN0, 0), mkref(BUF_OUT0, 0), &q->vecsz, 1);
    }
    /* Dense row-major 2-d array whose columns are too long for a tile of them to share a CTA: both passes read
       contiguous lines and store them transposed (pass 1 into scratch as [k1][i0], pass 2 from there into the
       user's [k0][k1]), instead of bracketing the column pass with two transposes -- two launches, each array
       read and written once per dimension */
    if (q->sz.rnk == 2 && b2_tensor_count(&q->vecsz) == 1 && !getenv("FFTW3_B200_R2R_UNFUSED") &&
        !p->opt.r2r_transposes) {
        int64_t n0 = q->sz.d[0].n, n1 = q->sz.d[1].n, m0, m1;
        size_t esz = 2 * real_size(q->prec);
        int radix[64], k0 = q->r2r_kind[0], k1 = q->r2r_kind[1];
        if (k0 >= 0 && k0 <= 10 && k1 >= 0 && k1 <= 10 && !r2r_work_len(k0, n0, &m0) && !r2r_work_len(k1, n1, &m1) &&
            q->sz.d[1].is == 1 && q->sz.d[1].os == 1 && q->sz.d[0].is == n1 && q->sz.d[0].os == n1 &&
            m0 >= 2 && m1 >= 2 && b2_factorize(m0, q->prec, 0, radix) != 0 && b2_factorize(m1, q->prec, 0, radix) != 0 &&
            single_pass_fits(m0, q->prec) && single_pass_fits(m1, q->prec) && (size_t)m0 * 4 * esz > 131072) {
            b2_view vin, vout;
            b2_tensor ub;
            b2_ops ops;
            int mark = p->nsteps;
            need_scratch(p, 0, (size_t)(n0 * n1) * real_size(q->prec));
            memset(&ops, 0, sizeof ops);
            ops.pre_op = B2D_LOAD_R2R; ops.post_op = B2D_STORE_R2R;
            ops.n_in = ops.n_out = (int)n1; ops.r2r_kind = k1;
            vin.re = vin.im = mkref(BUF_IN0, 0); vin.stride = 1;
            vout.re = vout.im = mkref(BUF_SCRATCH0, 0); vout.stride = n0;
            b2_tensor_init(&ub, 1); ub.d[0].n = n0; ub.d[0].is = n1; ub.d[0].os = 1;
            rc = emit_fft1d(p, q->prec, m1, vin, vout, &ub, ops, 1, "r2r (maps fused, lines stored transposed)");
            if (!rc) {
                ops.n_in = ops.n_out = (int)n0; ops.r2r_kind = k0;
                vin.re = vin.im = mkref(BUF_SCRATCH0, 0); vin.stride = 1;
                vout.re = vout.im = mkref(BUF_OUT0, 0); vout.stride = n1;
                ub.d[0].n = n1; ub.d[0].is = n0; ub.d[0].os = 1;
                rc = emit_fft1d(p, q->prec, m0, vin, vout, &ub, ops, 1, "r2r (maps fused, lines stored transposed)");
            }
            if (!rc) return 0;
            p->nsteps = mark;       /* could not be emitted that way: the general path below */
        }
    }
    for (d = q->sz.rnk - 1; d >= 0; --d) {
        int kind = q->r2r_kind[d];
        int64_t n = q->sz.d[d].n, m;
        b2_tensor ub, wb_in, wb_out, fb;
        rop_ctx c;
        b2_view wv;
        size_t esz = 2 * real_size(q->prec);
        int i;
        if (kind < 0 || kind > 10) return -1;
        if (r2r_work_len(kind, n, &m)) return -1;
        /* work transform fits one CTA: PRE and POST maps ride in the load and the store of ONE pass
           over the user's arrays (device/r2r_maps.cuh) -- one read and one write per dimension */
        {
            int radix[64];
            if (m >= 2 && b2_factorize(m, q->prec, 0, radix) != 0 && single_pass_fits(m, q->prec) &&
                !getenv("FFTW3_B200_R2R_UNFUSED")) {
                b2_view vin, vout;
                b2_ops ops;
                memset(&ops, 0, sizeof ops);
                ops.pre_op = B2D_LOAD_R2R; ops.post_op = B2D_STORE_R2R;
                ops.n_in = (int)n; ops.n_out = (int)n; ops.r2r_kind = kind;
                vin.re = vin.im = first ? mkref(BUF_IN0, 0) : mkref(BUF_OUT0, 0);
                vin.stride = first ? q->sz.d[d].is : q->sz.d[d].os;
                vout.re = vout.im = mkref(BUF_OUT0, 0);
                vout.stride = q->sz.d[d].os;
                other_dims(q, d, !first, &ub);
                {
                    int col = 0;
                    for (i = 0; i < ub.rnk; ++i)
                        if (ub.d[i].n > 1 && llabs(ub.d[i].is) < llabs(vin.stride)) col = 1;
                    if (!col || (size_t)m * 4 * esz <= 131072) {
                        rc = emit_fft1d(p, q->prec, m, vin, vout, &ub, ops, 1, "r2r (maps fused)");
                    } else {
                        /* strided lines too long for a tile of them to share a CTA: make them contiguous
                           with a tiled transpose, run the fused pass there, transpose back
                           (the reference's dft/indirect-transpose.c:38-59 strategy) */
                        b2_tensor srt = ub, tin, tout, tb;
                        b2_view sv;
                        int64_t dense = n;
                        b2_tensor_drop_unit(&srt);
                        if (srt.rnk > 1) qsort(srt.d, (size_t)srt.rnk, sizeof(b2_dim), cmp_is);
                        b2_tensor_init(&tin, 0); b2_tensor_init(&tout, 0); b2_tensor_init(&tb, 0);
                        tin.d[0].n = n; tin.d[0].is = vin.stride; tin.d[0].os = 1;
                        tout.d[0].n = n; tout.d[0].is = 1; tout.d[0].os = vout.stride;
                        tin.rnk = tout.rnk = 1;
                        for (i = 0; i < srt.rnk; ++i) {
                            tin.d[tin.rnk].n = srt.d[i].n; tin.d[tin.rnk].is = srt.d[i].is; tin.d[tin.rnk].os = dense;
                            tout.d[tout.rnk].n = srt.d[i].n; tout.d[tout.rnk].is = dense; tout.d[tout.rnk].os = srt.d[i].os;
                            tb.d[tb.rnk].n = srt.d[i].n; tb.d[tb.rnk].is = dense; tb.d[tb.rnk].os = dense;
                            tin.rnk++; tout.rnk++; tb.rnk++;
                            dense *= srt.d[i].n;
                        }
                        need_scratch(p, 0, (size_t)dense * real_size(q->prec));
                        sv.re = sv.im = mkref(BUF_SCRATCH0, 0); sv.stride = 1;
                        rc = emit_copy(p, q->prec, vin.re, mkref(BUF_SCRATCH0, 0), &tin, 1);
                        if (!rc) rc = emit_fft1d(p, q->prec, m, sv, sv, &tb, ops, 1, "r2r (maps fused, lines transposed)");
                        if (!rc) rc = emit_copy(p, q->prec, mkref(BUF_SCRATCH0, 0), vout.re, &tout, 1);
                    }
                }
                if (rc) return rc;
                first = 0;
                continue;
            }
        }
        /* user batch for this dim: input side strides for PRE, output side for POST */
        other_dims(q, d, !first, &ub);      /* .is = source strides, .os = out strides */
        wb_in = ub;
        make_work_batch(&wb_in, m);         /* sorted by source stride; .os := dense */
        /* POST tensor: same dim order, user strides = output strides */
        {
            b2_tensor pair = ub;
            b2_tensor_drop_unit(&pair);
            if (pair.rnk > 1) qsort(pair.d, (size_t)pair.rnk, sizeof(b2_dim), cmp_is);
            wb_out = wb_in;
            for (i = 0; i < pair.rnk; ++i) wb_out.d[i].is = pair.d[i].os;
        }
        need_scratch(p, 0, (size_t)(b2_tensor_count(&wb_in) > 0 ? b2_tensor_count(&wb_in) : 1) * (size_t)m * esz);
        memset(&c, 0, sizeof c);
        c.prec = q->prec; c.n = (int)n; c.m = (int)m; c.work_slot = 0; c.wdist = m;
        c.tw = NULL;
        if (kind == 4 || kind == 5 || kind == 6 || kind == 8 || kind == 9 || kind == 10) {
            c.tw = plan_table(p, q->prec, TAB_QUARTER, n, 0);
            if (!c.tw) return -1;
        }
        /* PRE */
        c.op = B2D_ROP_R2R_PRE | kind;
        c.xs = first ? q->sz.d[d].is : q->sz.d[d].os;
        c.x_re = first ? mkref(BUF_IN0, 0) : mkref(BUF_OUT0, 0);
        c.user_is_out = 0;
        rc = emit_realop(p, &c, &wb_in);
        if (rc) return rc;
        /* FFT_m in place on the work lines */
        fb = wb_in;
        for (i = 0; i < fb.rnk; ++i) fb.d[i].is = fb.d[i].os;
        wv.re = mkref(BUF_SCRATCH0, 0); wv.im = mkref(BUF_SCRATCH0, 1); wv.stride = 2;
        rc = emit_fft1d(p, q->prec, m, wv, wv, &fb, none, 1, "r2r core dft");
        if (rc) return rc;
        /* POST */
        c.op = B2D_ROP_R2R_POST | kind;
        c.xs = q->sz.d[d].os;
        c.y_re = mkref(BUF_OUT0, 0);
        c.user_is_out = 1;
        rc = emit_realop(p, &c, &wb_out);
        if (rc) return rc;
        first = 0;
    }
    return 0;
}

/* ------------------------------------------------------------------ entry */
static int tensor_valid(const b2_tensor *t, int allow_minfty)
{
    int i;
    if (t->rnk == B2_RNK_MINFTY) return allow_minfty;
    if (t->rnk < 0 || t->rnk > B2_MAXRANK) return 0;
    for (i = 0; i < t->rnk; ++i) if (t->d[i].n < (allow_minfty ? 0 : 1)) return 0;
    return 1;
}

static pthread_mutex_t g_planner_mu = PTHREAD_RECURSIVE_MUTEX_INITIALIZER_NP;
void b2_planner_lock(void) { pthread_mutex_lock(&g_planner_mu); }
void b2_planner_unlock(void) { pthread_mutex_unlock(&g_planner_mu); }

static b2_plan *mkplan_locked(const b2_problem *prob);
static void plan_destroy_locked(b2_plan *p);

static b2_plan *build_plan(const b2_problem *prob, const b2_plan_opts *opt, int alt);

/* Host arrays and a batch: cut the outermost batch dimension into chunks and build three chunk-sized plans for
   exec.c's pipeline (upload of chunk c + 1 | passes of chunk c | download of chunk c - 1).  Only when consecutive
   chunks touch disjoint, consecutive pieces of the user's arrays on both sides; a transform without a batch (the 3-d
   headline) depends on all of its input and cannot be pipelined this way. */
static void make_pipeline(b2_plan *p)
{
    const b2_problem *q = &p->prob;
    static const int tries[] = { 8, 6, 5, 7, 4, 3, 2 };
    const char *e;
    b2_problem sub;
    int d = -1, i, k, K = 0;
    int64_t n, lo, hi, span_in = 0, span_out = 0, in_off, out_off, min_bytes;
    size_t rs = real_size(q->prec);
    if (p->is_nop || q->vecsz.rnk < 1 || q->vecsz.rnk == B2_RNK_MINFTY || q->vecsz.rnk > B2_MAXRANK) return;
    if ((e = getenv("FFTW3_B200_PIPELINE")) && !atoi(e)) return;
    if (b2d_pointer_is_device(q->in0 ? q->in0 : q->out0) != 0) return;
    for (i = 0; i < q->vecsz.rnk; ++i)
        if (q->vecsz.d[i].n > 1 && (d < 0 || q->vecsz.d[i].is > q->vecsz.d[d].is)) d = i;
    if (d < 0 || q->vecsz.d[d].is <= 0 || q->vecsz.d[d].os <= 0) return;
    n = q->vecsz.d[d].n;
    for (i = 0; i < (int)(sizeof tries / sizeof tries[0]) && !K; ++i) if (n % tries[i] == 0) K = tries[i];
    if (!K) return;
    sub = *q;
    sub.vecsz.d[d].n = n / K;
    for (k = 0; k < 4; ++k) {
        const void *ptr = k == 0 ? q->in0 : k == 1 ? q->in1 : k == 2 ? q->out0 : q->out1;
        int64_t sp;
        if (!ptr) continue;
        b2_problem_span(&sub, k, &lo, &hi);
        sp = hi - lo + 1;
        /* the re and im pointers of one interleaved array together reach one real further than either alone */
        if (k < 2 ? (q->in0 && q->in1 && q->in0 != q->in1) : (q->out0 && q->out1 && q->out0 != q->out1)) ++sp;
        if (k < 2) { if (sp > span_in) span_in = sp; } else { if (sp > span_out) span_out = sp; }
    }
    in_off = q->vecsz.d[d].is * (n / K);
    out_off = q->vecsz.d[d].os * (n / K);
    if (in_off < span_in || out_off < span_out) return;          /* chunks would interleave in memory */
    min_bytes = (e = getenv("FFTW3_B200_PIPE_MIN_KB")) ? (int64_t)atol(e) << 10 : (int64_t)32 << 20;
    if ((span_in + span_out) * (int64_t)rs * K < min_bytes) return;
    for (k = 0; k < 3; ++k) {
        p->pipe[k] = build_plan(&sub, &p->opt, p->alt);
        if (!p->pipe[k] || p->pipe[k]->is_nop) {
            for (i = 0; i <= k; ++i) { plan_destroy_locked(p->pipe[i]); p->pipe[i] = NULL; }
            return;
        }
    }
    p->pipe_chunks = K; p->pipe_in_off = in_off; p->pipe_out_off = out_off;
}

b2_plan *b2_mkplan(const b2_problem *prob)
{
    b2_plan *p;
    b2_planner_lock();
    p = mkplan_locked(prob);
    if (p) make_pipeline(p);
    b2_planner_unlock();
    return p;
}

void b2_plan_destroy(b2_plan *p)
{
    if (!p) return;
    b2_planner_lock();
    plan_destroy_locked(p);
    b2_planner_unlock();
}

static b2_plan *build_plan(const b2_problem *prob, const b2_plan_opts *opt, int alt);

/* time a built plan on the user's own arrays (they are overwritten, as with the reference's measuring
   planner: kernel/timer.c:148-149, doc/reference.texi:405-416): one warm-up, then the minimum of `reps` runs */
static double time_plan(b2_plan *pl, int reps)
{
    const b2_problem *q = &pl->prob;
    float ms = 0, best = 1e30f;
    int r;
    b2_execute_ex(pl, q->in0, q->in1, q->out0, q->out1, 1);
    if (b2d_sync()) return -1;
    for (r = 0; r < reps; ++r) {
        if (b2d_timer_start()) return -1;
        b2_execute_ex(pl, q->in0, q->in1, q->out0, q->out1, 1);
        if (b2d_timer_stop(&ms)) return -1;
        if (ms < best) best = ms;
    }
    return best;
}

/* Whole-plan alternatives (the role of the reference planner's search over solver trees,
   kernel/planner.c:518-615): for multi-dimensional complex transforms that are much larger than L2, the
   plain plan (one HBM pass per dimension) competes with L2-resident pass pairs and with register-only
   sub-pass pairs.  FFTW_MEASURE builds each alternative, times it on the user's arrays and keeps the
   fastest; the choice is wisdom under the problem's signature, so FFTW_ESTIMATE / WISDOM_ONLY plans of
   the same problem reuse it. */
static int plan_alternatives(const b2_problem *prob, const b2_plan_opts *base, b2_plan_opts *alts, int max)
{
    int n = 0, i, last = prob->sz.rnk - 1;
    int64_t bytes = 2 * (int64_t)real_size(prob->prec), min_bytes;
    const char *e = getenv("FFTW3_B200_ALT_MIN_KB");      /* tests: let small problems have alternatives */
    alts[n++] = *base;
    if (prob->sz.rnk < 1 || opts_pinned_by_env()) return n;
    for (i = 0; i < prob->sz.rnk; ++i) bytes *= prob->sz.d[i].n;
    for (i = 0; i < prob->vecsz.rnk; ++i) bytes *= prob->vecsz.d[i].n > 0 ? prob->vecsz.d[i].n : 1;
    if (prob->kind != B2_C2C) bytes /= 2;
    /* the decompositions below only differ on arrays of some size; timing them costs a few executes each */
    min_bytes = e ? (int64_t)atol(e) << 10 : (int64_t)32 << 20;
    if (prob->kind == B2_R2C || prob->kind == B2_C2R) {
        /* even last dimension whose half-size transform is a four-step: split / merge fused into its outer
           pass (the default) or as a pass of its own */
        int64_t nl = prob->sz.d[last].n;
        int tmp[64];
        if (bytes >= min_bytes && nl % 2 == 0 && b2_factorize(nl / 2, prob->prec, 0, tmp) != 0 &&
            !single_pass_fits(nl / 2, prob->prec) && n < max) {
            alts[n] = *base; alts[n].real_unfused = 3; ++n;
        }
        return n;
    }
    if (prob->kind == B2_R2R) {
        /* dense 2-d array with long columns: two line passes with transposed stores (the default) or transposes
           around the column pass */
        if (bytes >= min_bytes && prob->sz.rnk == 2 && b2_tensor_count(&prob->vecsz) == 1 &&
            (size_t)prob->sz.d[0].n * 8 * real_size(prob->prec) > 131072 && n < max) {
            alts[n] = *base; alts[n].r2r_transposes = 1; ++n;
        }
        return n;
    }
    if (prob->kind != B2_C2C) return n;
    /* prime dimensions: the rule's choice, Rader, Bluestein (dft/rader.c vs dft/bluestein.c -- the reference's
       planner times both solvers too) */
    if (bytes >= min_bytes) {
        int has_prime = 0;
        for (i = 0; i < prob->sz.rnk; ++i) if (prob->sz.d[i].n > 13 && b2_is_prime(prob->sz.d[i].n)) has_prime = 1;
        if (has_prime) {
            if (n < max) { alts[n] = *base; alts[n].prime_mode = 1; ++n; }
            if (n < max) { alts[n] = *base; alts[n].prime_mode = 2; ++n; }
            return n;
        }
    }
    if (prob->sz.rnk < 2) return n;
    if (bytes < (e ? (int64_t)atol(e) << 10 : (int64_t)512 << 20)) return n;   /* arrays that (nearly) fit L2 gain nothing */
    {
        /* group sizes relative to the array so that the same code paths run on the tests' small problems */
        size_t g32 = (size_t)32 << 20, g16 = (size_t)16 << 20;
        while ((int64_t)g32 * 8 > bytes && g32 > 4096) { g32 >>= 1; g16 >>= 1; }
        if (n < max) { alts[n] = *base; alts[n].l2_block_bytes = g32; alts[n].l2_lanes = 2; alts[n].l2_keep = 4; ++n; }
        if (n < max) { alts[n] = *base; alts[n].l2_block_bytes = g16; alts[n].l2_lanes = 4; alts[n].l2_keep = 4; ++n; }
    }
    if (n < max && (prob->flags & (B2F_PATIENT | B2F_EXHAUSTIVE))) {
        alts[n] = *base; alts[n].split_mode = 1; alts[n].split_bytes = (size_t)16 << 20; alts[n].split_lanes = 3; ++n;
    }
    return n;
}

static b2_plan *mkplan_locked(const b2_problem *prob)
{
    b2_plan_opts base, alts[6];
    b2_plan *best = NULL;
    int nalt, a, chosen = 0, have = 0, inplace = (prob->in0 == prob->out0);
    unsigned pat = patience_of(prob->flags);
    b2_sig sig;
    double bestt = 1e30;
    if (!tensor_valid(&prob->sz, 0) || !tensor_valid(&prob->vecsz, 1)) return NULL;
    g_plan_t0 = now_seconds();
    opts_from_env(&base);
    nalt = plan_alternatives(prob, &base, alts, 6);
    if (nalt == 1) return build_plan(prob, &alts[0], 0);
    sig = b2_sig_of_problem(prob, inplace);
    b2_wisdom_set_prec(prob->prec);
    if (b2_wisdom_lookup(sig, pat, &chosen) && chosen >= 0 && chosen < nalt) have = 1;
    /* no wisdom for the decomposition is not a reason to fail a WISDOM_ONLY plan: the plain plan's passes
       decide that (api/apiplan.c:102-107 asks for wisdom of the problem, which its passes carry) */
    if (!have && (pat == 0 || (prob->flags & B2F_WISDOM_ONLY) || b2d_pointer_is_device(prob->in0 ? prob->in0 : prob->out0) != 1))
        return build_plan(prob, &alts[0], 0);
    if (have) {
        best = build_plan(prob, &alts[chosen], chosen);
        return best ? best : build_plan(prob, &alts[0], 0);
    }
    for (a = 0; a < nalt; ++a) {
        b2_plan *pl;
        double t;
        if (a > 0 && time_is_up()) break;
        pl = build_plan(prob, &alts[a], a);
        if (!pl) continue;
        t = pl->is_nop ? 0.0 : time_plan(pl, 3);
        if (getenv("FFTW3_B200_VERBOSE"))
            fprintf(stderr, "[b200 planner] whole-plan alternative %d (l2 %zu MiB x %d lanes, split %d, real-unfused %d, "
                            "r2r-transposes %d, prime-mode %d): %d steps, %.3f ms\n", a,
                    alts[a].l2_block_bytes >> 20, alts[a].l2_lanes, alts[a].split_mode, alts[a].real_unfused,
                    alts[a].r2r_transposes, alts[a].prime_mode, pl->nsteps, t);
        if (t >= 0 && t < bestt) { if (best) plan_destroy_locked(best); best = pl; bestt = t; chosen = a; }
        else plan_destroy_locked(pl);
    }
    if (best) { best->cost = bestt; b2_wisdom_store(sig, pat, chosen); }
    return best;
}

static b2_plan *build_plan(const b2_problem *prob, const b2_plan_opts *opt, int alt)
{
    b2_plan *p;
    int rc = 0, i;
    if (b2d_device_count() <= 0) return NULL;      /* no GPU, no plan: there is no CPU fallback */
    p = (b2_plan *)calloc(1, sizeof *p);
    if (!p) return NULL;
    p->refcnt = 1;
    p->prob = *prob;
    p->opt = *opt;
    p->alt = alt;
    b2_plan_lock_init(p);
    b2_wisdom_set_prec(prob->prec);
    b2_tensor_drop_unit(&p->prob.vecsz);
    /* unit transform dims are identities; keep their r2r kinds aligned */
    {
        b2_tensor *t = &p->prob.sz;
        int k = 0;
        int keep_last = (prob->kind == B2_R2C || prob->kind == B2_C2R);
        for (i = 0; i < t->rnk; ++i) {
            int is_last = (i == t->rnk - 1);
            if (t->d[i].n != 1 || (keep_last && is_last) || prob->kind == B2_R2R) {
                p->prob.r2r_kind[k] = prob->r2r_kind[i];
                t->d[k++] = t->d[i];
            }
        }
        t->rnk = k;
    }
    /* internal tensors hold B2_MAXRANK dims: transform + batch dims together must fit (the reference has
       no such bound, kernel/tensor.c:29-50 allocates; 16 combined non-unit dims is ample in practice) */
    if (p->prob.sz.rnk + (p->prob.vecsz.rnk > 0 ? p->prob.vecsz.rnk : 0) > B2_MAXRANK) { b2_plan_destroy(p); return NULL; }
    p->inplace = (prob->in0 == prob->out0);
    if (prob->kind == B2_C2C && prob->in0 == prob->out1 && prob->in1 == prob->out0) p->inplace = 0;
    if (b2_tensor_count(&p->prob.vecsz) == 0) { p->is_nop = 1; return p; }
    {
        /* In place with different input and output strides (e.g. an in-place transpose
           expressed as a rank-0 transform, rdft/rank0.c:345-381, rdft/vrank3-transpose.c):
           the reference marks the plain problem unsolvable (dft/problem.c:95-99) and solves
           it through buffered/indirect solvers (dft/indirect.c:55-108).  Same idea here:
           transform into dense plan-owned scratch, then one strided copy to the output. */
        int via_scratch = 0, k;
        b2_problem saved = p->prob;
        int64_t dense = 0;
        if (p->inplace && prob->kind != B2_R2C && prob->kind != B2_C2R &&
            (!b2_tensor_inplace_ok(&p->prob.sz) || !b2_tensor_inplace_ok(&p->prob.vecsz))) {
            int64_t ld = (prob->kind == B2_C2C) ? 2 : 1;
            via_scratch = 1;
            for (i = p->prob.sz.rnk - 1; i >= 0; --i) { p->prob.sz.d[i].os = ld; ld *= p->prob.sz.d[i].n; }
            for (i = p->prob.vecsz.rnk - 1; i >= 0; --i) { p->prob.vecsz.d[i].os = ld; ld *= p->prob.vecsz.d[i].n; }
            dense = ld;
            p->inplace = 0;
        }
        switch (prob->kind) {
        case B2_C2C: rc = plan_c2c(p); break;
        case B2_R2C: rc = plan_r2c(p); break;
        case B2_C2R: rc = plan_c2r(p); break;
        case B2_R2R: rc = plan_r2r(p); break;
        }
        if (!rc && via_scratch) {
            b2_tensor t;
            size_t rsz = (prob->prec == B2D_F32) ? 4 : 8;
            for (i = 0; i < p->nsteps; ++i)
                for (k = 0; k < 6; ++k) {
                    b2_ref *r = &p->steps[i].r[k];
                    if (r->buf == BUF_OUT0) r->buf = BUF_SCRATCH2;
                    else if (r->buf == BUF_OUT1) { r->buf = BUF_SCRATCH2; r->off += 1; }
                }
            if ((size_t)dense * rsz > p->scratch_bytes[2]) p->scratch_bytes[2] = (size_t)dense * rsz;
            /* copy back: dense strides (now in .os of the rewritten problem) -> the user's */
            b2_tensor_init(&t, 0);
            for (i = 0; i < p->prob.sz.rnk; ++i) {
                t.d[t.rnk].n = p->prob.sz.d[i].n; t.d[t.rnk].is = p->prob.sz.d[i].os;
                t.d[t.rnk].os = saved.sz.d[i].os; t.rnk++;
            }
            for (i = 0; i < p->prob.vecsz.rnk; ++i) {
                t.d[t.rnk].n = p->prob.vecsz.d[i].n; t.d[t.rnk].is = p->prob.vecsz.d[i].os;
                t.d[t.rnk].os = saved.vecsz.d[i].os; t.rnk++;
            }
            rc = emit_copy(p, prob->prec, mkref(BUF_SCRATCH2, 0), mkref(BUF_OUT0, 0), &t, 1);
            if (!rc && prob->kind == B2_C2C)
                rc = emit_copy(p, prob->prec, mkref(BUF_SCRATCH2, 1), mkref(BUF_OUT1, 0), &t, 1);
            p->prob = saved;
            p->inplace = 1;
        }
    }
    if (rc) { b2_plan_destroy(p); return NULL; }
    for (i = 0; i < B2_NSCRATCH; ++i) {
        if (p->scratch_bytes[i]) {
            p->scratch[i] = b2d_malloc(p->scratch_bytes[i]);
            if (!p->scratch[i]) { b2_plan_destroy(p); return NULL; }
        }
    }
    if (b2d_sync()) { b2_plan_destroy(p); return NULL; }
    return p;
}

static void plan_destroy_locked(b2_plan *p)
{
    int i;
    if (!p) return;
    b2_plan_lock_destroy(p);
    for (i = 0; i < p->ntables; ++i) b2_table_release(p->tables[i]);
    free(p->tables);
    for (i = 0; i < B2_NSCRATCH; ++i) b2d_free(p->scratch[i]);
    for (i = 0; i < 4; ++i) b2d_free(p->stage_dev[i]);
    for (i = 0; i < 3; ++i) plan_destroy_locked(p->pipe[i]);
    free(p->steps);
    free(p);
}

void b2_plan_print(const b2_plan *p, FILE *f)
{
    int i, j;
    if (p->is_nop) { fprintf(f, "(b200-nop)"); return; }
    fprintf(f, "(b200-plan");
    if (p->pipe_chunks > 1)
        fprintf(f, " [host arrays: %d chunks pipelined through 3 streams]", p->pipe_chunks);
    if (p->opt.l2_block_bytes && p->prob.kind == B2_C2C && p->prob.sz.rnk >= 2)
        fprintf(f, " [L2-resident pass pairs: %zu KiB groups, %d lanes]", p->opt.l2_block_bytes >> 10, p->opt.l2_lanes);
    for (i = 0; i < p->nsteps; ++i) {
        const b2_step *s = &p->steps[i];
        if (i >= 6 && i < p->nsteps - 3) {
            if (i == 6) fprintf(f, "\n  ... %d more passes ...", p->nsteps - 9);
            continue;
        }
        if (s->kind == STEP_FFT) {
            const b2d_fft_pass *q = &s->u.fft;
            fprintf(f, "\n  (fft-pass \"%s\" n=%d radix=", s->note, q->n);
            for (j = 0; j < q->nstages; ++j) fprintf(f, "%s%d", j ? "x" : "", q->radix[j]);
            fprintf(f, " batch=%lldx%lldx%lld ", (long long)q->bn[0], (long long)q->bn[1], (long long)q->bn[2]);
            if (q->kernel >= 3100) fprintf(f, "32x32 one-exchange tile=%d", q->kernel - 3100);
            else if (q->kernel >= 3000) fprintf(f, "warp-per-transform 32x32");
            else if (q->kernel >= 2100) fprintf(f, "codelet-tile=%d/c2r-merge", q->kernel % 100);
            else if (q->kernel >= 2000) fprintf(f, "codelet-tile=%d/r2c-split", q->kernel % 100);
            else if (q->kernel) fprintf(f, "codelet-tile=%d/f%d", q->kernel % 100, (q->kernel / 100) % 10);
            else fprintf(f, "generic tpb=%d tpx=%d", q->tpb, q->tpx);
            fprintf(f, " %s->%s%s)", q->load_col ? "col" : "row", q->store_col ? "col" : "row",
                    q->bluestein == 2 ? " rader" : (q->bluestein ? " bluestein" : (q->r2r_pair ? " paired-lines" : "")));
        } else if (s->kind == STEP_SPLIT) {
            fprintf(f, "\n  (%s pencils=%lldx%lld lane=%d)", s->note, (long long)s->u.split.nc, (long long)s->u.split.nb, s->lane);
        } else if (s->kind == STEP_COPY) {
            fprintf(f, "\n  (copy %lldx%lldx%lldx%lld)", (long long)s->u.copy.n[0], (long long)s->u.copy.n[1],
                    (long long)s->u.copy.n[2], (long long)s->u.copy.n[3]);
        } else {
            fprintf(f, "\n  (%s n=%d m=%d)", s->note, s->u.rop.n, s->u.rop.m);
        }
    }
    fprintf(f, ")");
}
