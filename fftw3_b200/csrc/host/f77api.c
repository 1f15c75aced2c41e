/* f77api.c -- the legacy Fortran 77 interface (dfftw_* / sfftw_* subroutines), compiled twice like api.c
 * (double: dfftw_, -DB2_SINGLE: sfftw_).
 *
 * Reference behaviour restated (api/f77api.c:33-160, api/f77funcs.h, threads/f77funcs.h, doc/legacy-fortran.texi):
 *   - every argument is passed by reference, the plan comes back through the first argument;
 *   - Fortran arrays are column-major, so the basic / advanced planners receive their dimensions (n, inembed,
 *     onembed and the r2r kinds) in REVERSED order; plan_dft_2d(nx, ny) plans the C transform ny x nx;
 *   - the guru planners take parallel arrays n / is / os (no reversal of the dimensions; the reference does
 *     reverse the r2r kinds there too, api/f77funcs.h:442-457, and so do we);
 *   - wisdom import / export go through Fortran character callbacks.
 * Two symbol spellings are exported for every subroutine, name_ and name__ (gfortran / ifort, and g77's
 * extra underscore for names that contain one: api/f77api.c:113-126).
 */
#include <stdio.h>
#include <stdlib.h>
#include "../../../include/fftw3.h"

#ifdef B2_SINGLE
#define X(name) fftwf_##name
#define F(name) sfftw_##name
typedef float R;
typedef fftwf_complex C;
#else
#define X(name) fftw_##name
#define F(name) dfftw_##name
typedef double R;
typedef fftw_complex C;
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define STR_(a) #a
#define STR(a) STR_(a)
/* define name_ and make name__ an alias of it */
#define SUB(name) \
    void CAT(F(name), __)() __attribute__((alias(STR(CAT(F(name), _))), visibility("default"))); \
    __attribute__((visibility("default"))) void CAT(F(name), _)

#define MAXR 32

static void rev(int rnk, const int *a, int *out)
{
    int i;
    for (i = 0; i < rnk; ++i) out[rnk - 1 - i] = a[i];
}

static void dims_of(int rnk, const int *n, const int *is, const int *os, X(iodim) *d)
{
    int i;
    for (i = 0; i < rnk; ++i) { d[i].n = n[i]; d[i].is = is[i]; d[i].os = os[i]; }
}

static void kinds_of(int rnk, const int *ik, X(r2r_kind) *k)
{
    int i;
    for (i = 0; i < rnk; ++i) k[i] = (X(r2r_kind))ik[rnk - 1 - i];      /* Fortran -> C order */
}

#define RANK_OK(r) ((r) >= 0 && (r) <= MAXR)

/* ---- lifecycle, wisdom, introspection (api/f77funcs.h:26-99) ---- */
SUB(execute)(X(plan) *const p) { X(execute)(*p); }
SUB(destroy_plan)(X(plan) *p) { X(destroy_plan)(*p); }
SUB(copy_plan)(X(plan) *pcopy, X(plan) *p) { *pcopy = X(copy_plan)(*p); }
SUB(cleanup)(void) { X(cleanup)(); }
SUB(forget_wisdom)(void) { X(forget_wisdom)(); }

typedef struct { void (*wr)(char *, void *); void *data; } wr_ctx;
static void wr_char(char c, void *d) { wr_ctx *w = (wr_ctx *)d; w->wr(&c, w->data); }
SUB(export_wisdom)(void (*f77_write_char)(char *, void *), void *data)
{
    wr_ctx w;
    w.wr = f77_write_char; w.data = data;
    X(export_wisdom)(wr_char, &w);
}

typedef struct { void (*rd)(int *, void *); void *data; } rd_ctx;
static int rd_char(void *d)
{
    rd_ctx *r = (rd_ctx *)d;
    int c;
    r->rd(&c, r->data);
    return c < 0 ? EOF : c;
}
SUB(import_wisdom)(int *isuccess, void (*f77_read_char)(int *, void *), void *data)
{
    rd_ctx r;
    r.rd = f77_read_char; r.data = data;
    *isuccess = X(import_wisdom)(rd_char, &r);
}
SUB(import_system_wisdom)(int *isuccess) { *isuccess = X(import_system_wisdom)(); }
SUB(print_plan)(X(plan) *const p) { X(print_plan)(*p); fflush(stdout); }
SUB(flops)(X(plan) *p, double *add, double *mul, double *fma) { X(flops)(*p, add, mul, fma); }
SUB(estimate_cost)(double *cost, X(plan) *const p) { *cost = X(estimate_cost)(*p); }
SUB(cost)(double *cost, X(plan) *const p) { *cost = X(cost)(*p); }
SUB(set_timelimit)(double *t) { X(set_timelimit)(*t); }

/* ---- threads (threads/f77funcs.h:26-44) ---- */
SUB(plan_with_nthreads)(int *nthreads) { X(plan_with_nthreads)(*nthreads); }
SUB(planner_nthreads)(int *nthreads) { *nthreads = X(planner_nthreads)(); }
SUB(init_threads)(int *okay) { *okay = X(init_threads)(); }
SUB(cleanup_threads)(void) { X(cleanup_threads)(); }

/* ---- complex DFT (api/f77funcs.h:104-195) ---- */
SUB(plan_dft)(X(plan) *p, int *rank, const int *n, C *in, C *out, int *sign, int *flags)
{
    int nr[MAXR];
    *p = NULL;
    if (!RANK_OK(*rank)) return;
    rev(*rank, n, nr);
    *p = X(plan_dft)(*rank, nr, in, out, *sign, (unsigned)*flags);
}
SUB(plan_dft_1d)(X(plan) *p, int *n, C *in, C *out, int *sign, int *flags)
{
    *p = X(plan_dft_1d)(*n, in, out, *sign, (unsigned)*flags);
}
SUB(plan_dft_2d)(X(plan) *p, int *nx, int *ny, C *in, C *out, int *sign, int *flags)
{
    *p = X(plan_dft_2d)(*ny, *nx, in, out, *sign, (unsigned)*flags);
}
SUB(plan_dft_3d)(X(plan) *p, int *nx, int *ny, int *nz, C *in, C *out, int *sign, int *flags)
{
    *p = X(plan_dft_3d)(*nz, *ny, *nx, in, out, *sign, (unsigned)*flags);
}
SUB(plan_many_dft)(X(plan) *p, int *rank, const int *n, int *howmany, C *in, const int *inembed, int *istride,
                   int *idist, C *out, const int *onembed, int *ostride, int *odist, int *sign, int *flags)
{
    int nr[MAXR], ie[MAXR], oe[MAXR];
    *p = NULL;
    if (!RANK_OK(*rank)) return;
    rev(*rank, n, nr); rev(*rank, inembed, ie); rev(*rank, onembed, oe);
    *p = X(plan_many_dft)(*rank, nr, *howmany, in, ie, *istride, *idist, out, oe, *ostride, *odist, *sign, (unsigned)*flags);
}
SUB(plan_guru_dft)(X(plan) *p, int *rank, const int *n, const int *is, const int *os, int *howmany_rank,
                   const int *h_n, const int *h_is, const int *h_os, C *in, C *out, int *sign, int *flags)
{
    X(iodim) d[MAXR], h[MAXR];
    *p = NULL;
    if (!RANK_OK(*rank) || !RANK_OK(*howmany_rank)) return;
    dims_of(*rank, n, is, os, d); dims_of(*howmany_rank, h_n, h_is, h_os, h);
    *p = X(plan_guru_dft)(*rank, d, *howmany_rank, h, in, out, *sign, (unsigned)*flags);
}
SUB(plan_guru_split_dft)(X(plan) *p, int *rank, const int *n, const int *is, const int *os, int *howmany_rank,
                         const int *h_n, const int *h_is, const int *h_os, R *ri, R *ii, R *ro, R *io, int *flags)
{
    X(iodim) d[MAXR], h[MAXR];
    *p = NULL;
    if (!RANK_OK(*rank) || !RANK_OK(*howmany_rank)) return;
    dims_of(*rank, n, is, os, d); dims_of(*howmany_rank, h_n, h_is, h_os, h);
    *p = X(plan_guru_split_dft)(*rank, d, *howmany_rank, h, ri, ii, ro, io, (unsigned)*flags);
}
SUB(execute_dft)(X(plan) *const p, C *in, C *out) { X(execute_dft)(*p, in, out); }
SUB(execute_split_dft)(X(plan) *const p, R *ri, R *ii, R *ro, R *io) { X(execute_split_dft)(*p, ri, ii, ro, io); }

/* ---- real-input and real-output DFTs (api/f77funcs.h:197-383) ---- */
#define REAL_DFT(NAME, TI, TO)                                                                                        \
    SUB(plan_dft_##NAME)(X(plan) *p, int *rank, const int *n, TI *in, TO *out, int *flags)                            \
    {                                                                                                                 \
        int nr[MAXR];                                                                                                 \
        *p = NULL;                                                                                                    \
        if (!RANK_OK(*rank)) return;                                                                                  \
        rev(*rank, n, nr);                                                                                            \
        *p = X(plan_dft_##NAME)(*rank, nr, in, out, (unsigned)*flags);                                                \
    }                                                                                                                 \
    SUB(plan_dft_##NAME##_1d)(X(plan) *p, int *n, TI *in, TO *out, int *flags)                                        \
    { *p = X(plan_dft_##NAME##_1d)(*n, in, out, (unsigned)*flags); }                                                  \
    SUB(plan_dft_##NAME##_2d)(X(plan) *p, int *nx, int *ny, TI *in, TO *out, int *flags)                              \
    { *p = X(plan_dft_##NAME##_2d)(*ny, *nx, in, out, (unsigned)*flags); }                                            \
    SUB(plan_dft_##NAME##_3d)(X(plan) *p, int *nx, int *ny, int *nz, TI *in, TO *out, int *flags)                     \
    { *p = X(plan_dft_##NAME##_3d)(*nz, *ny, *nx, in, out, (unsigned)*flags); }                                       \
    SUB(plan_many_dft_##NAME)(X(plan) *p, int *rank, const int *n, int *howmany, TI *in, const int *inembed,          \
                              int *istride, int *idist, TO *out, const int *onembed, int *ostride, int *odist,       \
                              int *flags)                                                                             \
    {                                                                                                                 \
        int nr[MAXR], ie[MAXR], oe[MAXR];                                                                             \
        *p = NULL;                                                                                                    \
        if (!RANK_OK(*rank)) return;                                                                                  \
        rev(*rank, n, nr); rev(*rank, inembed, ie); rev(*rank, onembed, oe);                                          \
        *p = X(plan_many_dft_##NAME)(*rank, nr, *howmany, in, ie, *istride, *idist, out, oe, *ostride, *odist,        \
                                     (unsigned)*flags);                                                               \
    }                                                                                                                 \
    SUB(plan_guru_dft_##NAME)(X(plan) *p, int *rank, const int *n, const int *is, const int *os, int *howmany_rank,   \
                              const int *h_n, const int *h_is, const int *h_os, TI *in, TO *out, int *flags)          \
    {                                                                                                                 \
        X(iodim) d[MAXR], h[MAXR];                                                                                    \
        *p = NULL;                                                                                                    \
        if (!RANK_OK(*rank) || !RANK_OK(*howmany_rank)) return;                                                       \
        dims_of(*rank, n, is, os, d); dims_of(*howmany_rank, h_n, h_is, h_os, h);                                     \
        *p = X(plan_guru_dft_##NAME)(*rank, d, *howmany_rank, h, in, out, (unsigned)*flags);                          \
    }                                                                                                                 \
    SUB(execute_dft_##NAME)(X(plan) *const p, TI *in, TO *out) { X(execute_dft_##NAME)(*p, in, out); }

REAL_DFT(r2c, R, C)
REAL_DFT(c2r, C, R)

SUB(plan_guru_split_dft_r2c)(X(plan) *p, int *rank, const int *n, const int *is, const int *os, int *howmany_rank,
                             const int *h_n, const int *h_is, const int *h_os, R *in, R *ro, R *io, int *flags)
{
    X(iodim) d[MAXR], h[MAXR];
    *p = NULL;
    if (!RANK_OK(*rank) || !RANK_OK(*howmany_rank)) return;
    dims_of(*rank, n, is, os, d); dims_of(*howmany_rank, h_n, h_is, h_os, h);
    *p = X(plan_guru_split_dft_r2c)(*rank, d, *howmany_rank, h, in, ro, io, (unsigned)*flags);
}
SUB(plan_guru_split_dft_c2r)(X(plan) *p, int *rank, const int *n, const int *is, const int *os, int *howmany_rank,
                             const int *h_n, const int *h_is, const int *h_os, R *ri, R *ii, R *out, int *flags)
{
    X(iodim) d[MAXR], h[MAXR];
    *p = NULL;
    if (!RANK_OK(*rank) || !RANK_OK(*howmany_rank)) return;
    dims_of(*rank, n, is, os, d); dims_of(*howmany_rank, h_n, h_is, h_os, h);
    *p = X(plan_guru_split_dft_c2r)(*rank, d, *howmany_rank, h, ri, ii, out, (unsigned)*flags);
}
SUB(execute_split_dft_r2c)(X(plan) *const p, R *in, R *ro, R *io) { X(execute_split_dft_r2c)(*p, in, ro, io); }
SUB(execute_split_dft_c2r)(X(plan) *const p, R *ri, R *ii, R *out) { X(execute_split_dft_c2r)(*p, ri, ii, out); }

/* ---- real-to-real (api/f77funcs.h:385-465) ---- */
SUB(plan_r2r)(X(plan) *p, int *rank, const int *n, R *in, R *out, int *kind, int *flags)
{
    int nr[MAXR];
    X(r2r_kind) k[MAXR];
    *p = NULL;
    if (!RANK_OK(*rank)) return;
    rev(*rank, n, nr); kinds_of(*rank, kind, k);
    *p = X(plan_r2r)(*rank, nr, in, out, k, (unsigned)*flags);
}
SUB(plan_r2r_1d)(X(plan) *p, int *n, R *in, R *out, int *kind, int *flags)
{
    *p = X(plan_r2r_1d)(*n, in, out, (X(r2r_kind))*kind, (unsigned)*flags);
}
SUB(plan_r2r_2d)(X(plan) *p, int *nx, int *ny, R *in, R *out, int *kindx, int *kindy, int *flags)
{
    *p = X(plan_r2r_2d)(*ny, *nx, in, out, (X(r2r_kind))*kindy, (X(r2r_kind))*kindx, (unsigned)*flags);
}
SUB(plan_r2r_3d)(X(plan) *p, int *nx, int *ny, int *nz, R *in, R *out, int *kindx, int *kindy, int *kindz, int *flags)
{
    *p = X(plan_r2r_3d)(*nz, *ny, *nx, in, out, (X(r2r_kind))*kindz, (X(r2r_kind))*kindy, (X(r2r_kind))*kindx, (unsigned)*flags);
}
SUB(plan_many_r2r)(X(plan) *p, int *rank, const int *n, int *howmany, R *in, const int *inembed, int *istride,
                   int *idist, R *out, const int *onembed, int *ostride, int *odist, int *kind, int *flags)
{
    int nr[MAXR], ie[MAXR], oe[MAXR];
    X(r2r_kind) k[MAXR];
    *p = NULL;
    if (!RANK_OK(*rank)) return;
    rev(*rank, n, nr); rev(*rank, inembed, ie); rev(*rank, onembed, oe); kinds_of(*rank, kind, k);
    *p = X(plan_many_r2r)(*rank, nr, *howmany, in, ie, *istride, *idist, out, oe, *ostride, *odist, k, (unsigned)*flags);
}
SUB(plan_guru_r2r)(X(plan) *p, int *rank, const int *n, const int *is, const int *os, int *howmany_rank,
                   const int *h_n, const int *h_is, const int *h_os, R *in, R *out, int *kind, int *flags)
{
    X(iodim) d[MAXR], h[MAXR];
    X(r2r_kind) k[MAXR];
    *p = NULL;
    if (!RANK_OK(*rank) || !RANK_OK(*howmany_rank)) return;
    dims_of(*rank, n, is, os, d); dims_of(*howmany_rank, h_n, h_is, h_os, h); kinds_of(*rank, kind, k);
    *p = X(plan_guru_r2r)(*rank, d, *howmany_rank, h, in, out, k, (unsigned)*flags);
}
SUB(execute_r2r)(X(plan) *const p, R *in, R *out) { X(execute_r2r)(*p, in, out); }
