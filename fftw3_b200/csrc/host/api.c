/* api.c -- the FFTW 3 public API for one precision (compiled twice: default =
 * double / fftw_, -DB2_SINGLE = float / fftwf_).
 *
 * Each function mirrors the reference function of the same name; the reference
 * implementation it replaces is cited next to it (paths relative to
 * /root/reference).  Validation rules, stride conventions, padding and the
 * NULL-on-failure error convention follow the reference; what happens after
 * the canonical problem is built is new (planner.c).
 */
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "b2_internal.h"

#ifdef B2_SINGLE
typedef float R;
#define X(name) fftwf_##name
#define PREC B2D_F32
#else
typedef double R;
#define X(name) fftw_##name
#define PREC B2D_F64
#endif
typedef R C[2];

struct X(plan_s) {
    b2_plan *pln;
    int sign;
};
typedef struct X(plan_s) *X(plan);

typedef struct { int n, is, os; } iodim32;
typedef struct { ptrdiff_t n, is, os; } iodim64;
typedef int r2r_kind_t;
typedef void (*write_char_func)(char c, void *);
typedef int (*read_char_func)(void *);

#define FFTW_FORWARD_ (-1)

/* ------------------------------------------------------------ tensor builders */
/* row-major dims with embedding arrays; api/mktensor-rowmajor.c:23-43 */
static int rowmajor(b2_tensor *t, int rank, const int *n, const int *niphys, const int *nophys,
                    int64_t is, int64_t os)
{
    int i;
    if (rank < 0 || rank > B2_MAXRANK) return -1;
    for (i = 0; i < rank; ++i) if (n[i] <= 0) return -1;      /* api/mktensor-rowmajor.c:45-61 */
    b2_tensor_init(t, rank);
    if (rank > 0) {
        t->d[rank - 1].n = n[rank - 1]; t->d[rank - 1].is = is; t->d[rank - 1].os = os;
        for (i = rank - 1; i > 0; --i) {
            t->d[i - 1].n = n[i - 1];
            t->d[i - 1].is = t->d[i].is * niphys[i];
            t->d[i - 1].os = t->d[i].os * nophys[i];
        }
    }
    return 0;
}

static int howmany_tensor(b2_tensor *t, int howmany, int64_t idist, int64_t odist)
{
    if (howmany < 0) return -1;
    b2_tensor_init(t, 1);
    t->d[0].n = howmany; t->d[0].is = idist; t->d[0].os = odist;
    return 0;
}

/* guru dims; api/mktensor-iodims.h:23-62.  rank INT_MAX is FFTW's RNK_MINFTY. */
#define DEFINE_IODIMS(NAME, TYPE)                                                              \
    static int NAME(b2_tensor *t, int rank, const TYPE *dims, int64_t is, int64_t os, int minfty) \
    {                                                                                          \
        int i;                                                                                 \
        if (rank == INT_MAX) {                                                                 \
            if (!minfty) return -1;                                                            \
            b2_tensor_init(t, 0); t->rnk = B2_RNK_MINFTY; return 0;                            \
        }                                                                                      \
        if (rank < 0 || rank > B2_MAXRANK) return -1;                                          \
        b2_tensor_init(t, rank);                                                               \
        for (i = 0; i < rank; ++i) {                                                           \
            if (minfty ? dims[i].n < 0 : dims[i].n <= 0) return -1;                            \
            t->d[i].n = dims[i].n; t->d[i].is = (int64_t)dims[i].is * is;                      \
            t->d[i].os = (int64_t)dims[i].os * os;                                             \
        }                                                                                      \
        return 0;                                                                              \
    }
DEFINE_IODIMS(iodims32, iodim32)
DEFINE_IODIMS(iodims64, iodim64)

/* ------------------------------------------------------------------- mkplan */
static X(plan) finish(b2_problem *q, int sign)
{
    X(plan) p;
    b2_plan *pln;
    q->prec = PREC;
    pln = b2_mkplan(q);
    if (!pln) return NULL;                       /* api/apiplan.c:139-165: NULL on failure */
    p = (X(plan))malloc(sizeof *p);
    if (!p) { b2_plan_destroy(pln); return NULL; }
    p->pln = pln;
    p->sign = sign;
    return p;
}

static void set_c2c_ptrs(b2_problem *q, C *in, C *out, int sign)
{
    /* kernel/extract-reim.c:27-36: backward == forward with re/im exchanged */
    R *i0 = (R *)in, *o0 = (R *)out;
    if (sign == FFTW_FORWARD_) { q->in0 = i0; q->in1 = i0 + 1; q->out0 = o0; q->out1 = o0 + 1; }
    else { q->in0 = i0 + 1; q->in1 = i0; q->out0 = o0 + 1; q->out1 = o0; }
}

/* ------------------------------------------------------------- complex DFT */
/* api/plan-many-dft.c:26-51 */
X(plan) X(plan_many_dft)(int rank, const int *n, int howmany, C *in, const int *inembed, int istride,
                         int idist, C *out, const int *onembed, int ostride, int odist, int sign,
                         unsigned flags)
{
    b2_problem q;
    memset(&q, 0, sizeof q);
    if (sign != -1 && sign != 1) return NULL;
    if (rowmajor(&q.sz, rank, n, inembed ? inembed : n, onembed ? onembed : n, 2 * (int64_t)istride,
                 2 * (int64_t)ostride)) return NULL;
    if (howmany_tensor(&q.vecsz, howmany, 2 * (int64_t)idist, 2 * (int64_t)odist)) return NULL;
    q.kind = B2_C2C; q.flags = flags;
    set_c2c_ptrs(&q, in, out, sign);
    return finish(&q, sign);
}

/* api/plan-dft.c:23-29 */
X(plan) X(plan_dft)(int rank, const int *n, C *in, C *out, int sign, unsigned flags)
{
    return X(plan_many_dft)(rank, n, 1, in, 0, 1, 1, out, 0, 1, 1, sign, flags);
}
/* api/plan-dft-1d.c:24, plan-dft-2d.c, plan-dft-3d.c */
X(plan) X(plan_dft_1d)(int n, C *in, C *out, int sign, unsigned flags)
{
    return X(plan_dft)(1, &n, in, out, sign, flags);
}
X(plan) X(plan_dft_2d)(int n0, int n1, C *in, C *out, int sign, unsigned flags)
{
    int n[2]; n[0] = n0; n[1] = n1;
    return X(plan_dft)(2, n, in, out, sign, flags);
}
X(plan) X(plan_dft_3d)(int n0, int n1, int n2, C *in, C *out, int sign, unsigned flags)
{
    int n[3]; n[0] = n0; n[1] = n1; n[2] = n2;
    return X(plan_dft)(3, n, in, out, sign, flags);
}

/* api/plan-guru-dft.h:24-44 (guru and guru64 share one body there too) */
#define GURU_DFT(NAME, TYPE, MK)                                                                   \
    X(plan) X(NAME)(int rank, const TYPE *dims, int howmany_rank, const TYPE *howmany_dims, C *in, \
                    C *out, int sign, unsigned flags)                                              \
    {                                                                                              \
        b2_problem q;                                                                              \
        memset(&q, 0, sizeof q);                                                                   \
        if (sign != -1 && sign != 1) return NULL;                                                  \
        if (MK(&q.sz, rank, dims, 2, 2, 0) || MK(&q.vecsz, howmany_rank, howmany_dims, 2, 2, 1))   \
            return NULL;                                                                           \
        q.kind = B2_C2C; q.flags = flags;                                                          \
        set_c2c_ptrs(&q, in, out, sign);                                                           \
        return finish(&q, sign);                                                                   \
    }
GURU_DFT(plan_guru_dft, iodim32, iodims32)
GURU_DFT(plan_guru64_dft, iodim64, iodims64)

/* api/plan-guru-split-dft.h:24-39: no sign argument, always FORWARD on (ri,ii) */
#define GURU_SPLIT_DFT(NAME, TYPE, MK)                                                             \
    X(plan) X(NAME)(int rank, const TYPE *dims, int howmany_rank, const TYPE *howmany_dims, R *ri,  \
                    R *ii, R *ro, R *io, unsigned flags)                                           \
    {                                                                                              \
        b2_problem q;                                                                              \
        memset(&q, 0, sizeof q);                                                                   \
        if (MK(&q.sz, rank, dims, 1, 1, 0) || MK(&q.vecsz, howmany_rank, howmany_dims, 1, 1, 1))   \
            return NULL;                                                                           \
        q.kind = B2_C2C; q.flags = flags;                                                          \
        q.in0 = ri; q.in1 = ii; q.out0 = ro; q.out1 = io;                                          \
        return finish(&q, FFTW_FORWARD_);                                                          \
    }
GURU_SPLIT_DFT(plan_guru_split_dft, iodim32, iodims32)
GURU_SPLIT_DFT(plan_guru64_split_dft, iodim64, iodims64)

/* ------------------------------------------------------------------ r2c / c2r */
/* default embeddings: api/rdft2-pad.c:24-39 */
static void rdft2_pad(int rank, const int *n, const int *nembed, int inplace, int cmplx, int *out)
{
    int i;
    if (nembed) { for (i = 0; i < rank; ++i) out[i] = nembed[i]; return; }
    for (i = 0; i < rank; ++i) out[i] = n[i];
    if (rank > 0 && (cmplx || inplace))
        out[rank - 1] = (n[rank - 1] / 2 + 1) * (cmplx ? 1 : 2);
}

/* api/plan-many-dft-r2c.c:24-57 */
X(plan) X(plan_many_dft_r2c)(int rank, const int *n, int howmany, R *in, const int *inembed, int istride,
                             int idist, C *out, const int *onembed, int ostride, int odist, unsigned flags)
{
    b2_problem q;
    int ni[B2_MAXRANK], no[B2_MAXRANK], i;
    int inplace = ((void *)in == (void *)out);
    memset(&q, 0, sizeof q);
    if (rank < 0 || rank > B2_MAXRANK || howmany < 0) return NULL;
    for (i = 0; i < rank; ++i) if (n[i] <= 0) return NULL;
    rdft2_pad(rank, n, inembed, inplace, 0, ni);
    rdft2_pad(rank, n, onembed, inplace, 1, no);
    if (rowmajor(&q.sz, rank, n, ni, no, istride, 2 * (int64_t)ostride)) return NULL;
    if (howmany_tensor(&q.vecsz, howmany, idist, 2 * (int64_t)odist)) return NULL;
    q.kind = B2_R2C; q.flags = flags;
    q.in0 = in; q.out0 = (R *)out; q.out1 = (R *)out + 1;
    return finish(&q, FFTW_FORWARD_);
}

X(plan) X(plan_dft_r2c)(int rank, const int *n, R *in, C *out, unsigned flags)
{
    return X(plan_many_dft_r2c)(rank, n, 1, in, 0, 1, 1, out, 0, 1, 1, flags);
}
X(plan) X(plan_dft_r2c_1d)(int n, R *in, C *out, unsigned flags) { return X(plan_dft_r2c)(1, &n, in, out, flags); }
X(plan) X(plan_dft_r2c_2d)(int n0, int n1, R *in, C *out, unsigned flags)
{
    int n[2]; n[0] = n0; n[1] = n1;
    return X(plan_dft_r2c)(2, n, in, out, flags);
}
X(plan) X(plan_dft_r2c_3d)(int n0, int n1, int n2, R *in, C *out, unsigned flags)
{
    int n[3]; n[0] = n0; n[1] = n1; n[2] = n2;
    return X(plan_dft_r2c)(3, n, in, out, flags);
}

static int c2r_flags_ok(int rank_gt1, int inplace, unsigned flags)
{
    /* multi-dimensional out-of-place c2r cannot preserve its input
       (api/plan-many-dft-c2r.c:41-42, doc/reference.texi:511-527) */
    if (rank_gt1 && !inplace && (flags & B2F_PRESERVE_INPUT)) return 0;
    return 1;
}

/* api/plan-many-dft-c2r.c:24-59 */
X(plan) X(plan_many_dft_c2r)(int rank, const int *n, int howmany, C *in, const int *inembed, int istride,
                             int idist, R *out, const int *onembed, int ostride, int odist, unsigned flags)
{
    b2_problem q;
    int ni[B2_MAXRANK], no[B2_MAXRANK], i;
    int inplace = ((void *)in == (void *)out);
    memset(&q, 0, sizeof q);
    if (rank < 0 || rank > B2_MAXRANK || howmany < 0) return NULL;
    for (i = 0; i < rank; ++i) if (n[i] <= 0) return NULL;
    if (!c2r_flags_ok(rank > 1, inplace, flags)) return NULL;
    rdft2_pad(rank, n, inembed, inplace, 1, ni);
    rdft2_pad(rank, n, onembed, inplace, 0, no);
    if (rowmajor(&q.sz, rank, n, ni, no, 2 * (int64_t)istride, ostride)) return NULL;
    if (howmany_tensor(&q.vecsz, howmany, 2 * (int64_t)idist, odist)) return NULL;
    q.kind = B2_C2R; q.flags = flags;
    q.in0 = (R *)in; q.in1 = (R *)in + 1; q.out0 = out;
    return finish(&q, 1);
}

X(plan) X(plan_dft_c2r)(int rank, const int *n, C *in, R *out, unsigned flags)
{
    return X(plan_many_dft_c2r)(rank, n, 1, in, 0, 1, 1, out, 0, 1, 1, flags);
}
X(plan) X(plan_dft_c2r_1d)(int n, C *in, R *out, unsigned flags) { return X(plan_dft_c2r)(1, &n, in, out, flags); }
X(plan) X(plan_dft_c2r_2d)(int n0, int n1, C *in, R *out, unsigned flags)
{
    int n[2]; n[0] = n0; n[1] = n1;
    return X(plan_dft_c2r)(2, n, in, out, flags);
}
X(plan) X(plan_dft_c2r_3d)(int n0, int n1, int n2, C *in, R *out, unsigned flags)
{
    int n[3]; n[0] = n0; n[1] = n1; n[2] = n2;
    return X(plan_dft_c2r)(3, n, in, out, flags);
}

/* api/plan-guru-dft-r2c.h, plan-guru-dft-c2r.h, plan-guru-split-dft-r2c.h, ...-c2r.h */
#define GURU_R2C(NAME, TYPE, MK)                                                                   \
    X(plan) X(NAME)(int rank, const TYPE *dims, int howmany_rank, const TYPE *howmany_dims, R *in,  \
                    C *out, unsigned flags)                                                        \
    {                                                                                              \
        b2_problem q;                                                                              \
        memset(&q, 0, sizeof q);                                                                   \
        if (rank < 0) return NULL;   /* rank 0 is a copy (rdft/rank0-rdft2.c) */                                                                 \
        if (MK(&q.sz, rank, dims, 1, 2, 0) || MK(&q.vecsz, howmany_rank, howmany_dims, 1, 2, 1))   \
            return NULL;                                                                           \
        q.kind = B2_R2C; q.flags = flags;                                                          \
        q.in0 = in; q.out0 = (R *)out; q.out1 = (R *)out + 1;                                      \
        return finish(&q, FFTW_FORWARD_);                                                          \
    }
GURU_R2C(plan_guru_dft_r2c, iodim32, iodims32)
GURU_R2C(plan_guru64_dft_r2c, iodim64, iodims64)

#define GURU_SPLIT_R2C(NAME, TYPE, MK)                                                             \
    X(plan) X(NAME)(int rank, const TYPE *dims, int howmany_rank, const TYPE *howmany_dims, R *in,  \
                    R *ro, R *io, unsigned flags)                                                  \
    {                                                                                              \
        b2_problem q;                                                                              \
        memset(&q, 0, sizeof q);                                                                   \
        if (rank < 0) return NULL;   /* rank 0 is a copy (rdft/rank0-rdft2.c) */                                                                 \
        if ((void *)in == (void *)ro && rank > 1) return NULL; /* doc/reference.texi:1452-1459 */  \
        if (MK(&q.sz, rank, dims, 1, 1, 0) || MK(&q.vecsz, howmany_rank, howmany_dims, 1, 1, 1))   \
            return NULL;                                                                           \
        q.kind = B2_R2C; q.flags = flags;                                                          \
        q.in0 = in; q.out0 = ro; q.out1 = io;                                                      \
        return finish(&q, FFTW_FORWARD_);                                                          \
    }
GURU_SPLIT_R2C(plan_guru_split_dft_r2c, iodim32, iodims32)
GURU_SPLIT_R2C(plan_guru64_split_dft_r2c, iodim64, iodims64)

#define GURU_C2R(NAME, TYPE, MK)                                                                   \
    X(plan) X(NAME)(int rank, const TYPE *dims, int howmany_rank, const TYPE *howmany_dims, C *in,  \
                    R *out, unsigned flags)                                                        \
    {                                                                                              \
        b2_problem q;                                                                              \
        memset(&q, 0, sizeof q);                                                                   \
        if (rank < 0) return NULL;   /* rank 0 is a copy (rdft/rank0-rdft2.c) */                                                                 \
        if (!c2r_flags_ok(rank > 1, (void *)in == (void *)out, flags)) return NULL;                \
        if (MK(&q.sz, rank, dims, 2, 1, 0) || MK(&q.vecsz, howmany_rank, howmany_dims, 2, 1, 1))   \
            return NULL;                                                                           \
        q.kind = B2_C2R; q.flags = flags;                                                          \
        q.in0 = (R *)in; q.in1 = (R *)in + 1; q.out0 = out;                                        \
        return finish(&q, 1);                                                                      \
    }
GURU_C2R(plan_guru_dft_c2r, iodim32, iodims32)
GURU_C2R(plan_guru64_dft_c2r, iodim64, iodims64)

#define GURU_SPLIT_C2R(NAME, TYPE, MK)                                                             \
    X(plan) X(NAME)(int rank, const TYPE *dims, int howmany_rank, const TYPE *howmany_dims, R *ri,  \
                    R *ii, R *out, unsigned flags)                                                 \
    {                                                                                              \
        b2_problem q;                                                                              \
        memset(&q, 0, sizeof q);                                                                   \
        if (rank < 0) return NULL;   /* rank 0 is a copy (rdft/rank0-rdft2.c) */                                                                 \
        if ((void *)ri == (void *)out && rank > 1) return NULL;                                    \
        if (!c2r_flags_ok(rank > 1, (void *)ri == (void *)out, flags)) return NULL;                \
        if (MK(&q.sz, rank, dims, 1, 1, 0) || MK(&q.vecsz, howmany_rank, howmany_dims, 1, 1, 1))   \
            return NULL;                                                                           \
        q.kind = B2_C2R; q.flags = flags;                                                          \
        q.in0 = ri; q.in1 = ii; q.out0 = out;                                                      \
        return finish(&q, 1);                                                                      \
    }
GURU_SPLIT_C2R(plan_guru_split_dft_c2r, iodim32, iodims32)
GURU_SPLIT_C2R(plan_guru64_split_dft_c2r, iodim64, iodims64)

/* ----------------------------------------------------------------------- r2r */
/* api/plan-many-r2r.c:26-50 ; kinds are passed through (api/map-r2r-kind.c:24-50) */
X(plan) X(plan_many_r2r)(int rank, const int *n, int howmany, R *in, const int *inembed, int istride,
                         int idist, R *out, const int *onembed, int ostride, int odist,
                         const r2r_kind_t *kind, unsigned flags)
{
    b2_problem q;
    int i;
    memset(&q, 0, sizeof q);
    if (rowmajor(&q.sz, rank, n, inembed ? inembed : n, onembed ? onembed : n, istride, ostride)) return NULL;
    if (howmany_tensor(&q.vecsz, howmany, idist, odist)) return NULL;
    for (i = 0; i < rank; ++i) {
        if (kind[i] < 0 || kind[i] > 10) return NULL;
        q.r2r_kind[i] = kind[i];
    }
    q.kind = B2_R2R; q.flags = flags;
    q.in0 = in; q.out0 = out;
    return finish(&q, FFTW_FORWARD_);
}
X(plan) X(plan_r2r)(int rank, const int *n, R *in, R *out, const r2r_kind_t *kind, unsigned flags)
{
    return X(plan_many_r2r)(rank, n, 1, in, 0, 1, 1, out, 0, 1, 1, kind, flags);
}
X(plan) X(plan_r2r_1d)(int n, R *in, R *out, r2r_kind_t kind, unsigned flags)
{
    return X(plan_r2r)(1, &n, in, out, &kind, flags);
}
X(plan) X(plan_r2r_2d)(int n0, int n1, R *in, R *out, r2r_kind_t k0, r2r_kind_t k1, unsigned flags)
{
    int n[2]; r2r_kind_t k[2];
    n[0] = n0; n[1] = n1; k[0] = k0; k[1] = k1;
    return X(plan_r2r)(2, n, in, out, k, flags);
}
X(plan) X(plan_r2r_3d)(int n0, int n1, int n2, R *in, R *out, r2r_kind_t k0, r2r_kind_t k1, r2r_kind_t k2,
                       unsigned flags)
{
    int n[3]; r2r_kind_t k[3];
    n[0] = n0; n[1] = n1; n[2] = n2; k[0] = k0; k[1] = k1; k[2] = k2;
    return X(plan_r2r)(3, n, in, out, k, flags);
}
#define GURU_R2R(NAME, TYPE, MK)                                                                   \
    X(plan) X(NAME)(int rank, const TYPE *dims, int howmany_rank, const TYPE *howmany_dims, R *in,  \
                    R *out, const r2r_kind_t *kind, unsigned flags)                                \
    {                                                                                              \
        b2_problem q;                                                                              \
        int i;                                                                                     \
        memset(&q, 0, sizeof q);                                                                   \
        if (MK(&q.sz, rank, dims, 1, 1, 0) || MK(&q.vecsz, howmany_rank, howmany_dims, 1, 1, 1))   \
            return NULL;                                                                           \
        for (i = 0; i < rank; ++i) {                                                               \
            if (kind[i] < 0 || kind[i] > 10) return NULL;                                          \
            q.r2r_kind[i] = kind[i];                                                               \
        }                                                                                          \
        q.kind = B2_R2R; q.flags = flags;                                                          \
        q.in0 = in; q.out0 = out;                                                                  \
        return finish(&q, FFTW_FORWARD_);                                                          \
    }
GURU_R2R(plan_guru_r2r, iodim32, iodims32)
GURU_R2R(plan_guru64_r2r, iodim64, iodims64)

/* ------------------------------------------------------------------ execute */
/* api/execute.c:23-27 */
void X(execute)(const X(plan) p)
{
    const b2_problem *q = &p->pln->prob;
    b2_execute(p->pln, q->in0, q->in1, q->out0, q->out1);
}
/* api/execute-dft.c:25-32 */
void X(execute_dft)(const X(plan) p, C *in, C *out)
{
    R *i0 = (R *)in, *o0 = (R *)out;
    if (p->sign == FFTW_FORWARD_) b2_execute(p->pln, i0, i0 + 1, o0, o0 + 1);
    else b2_execute(p->pln, i0 + 1, i0, o0 + 1, o0);
}
/* api/execute-split-dft.c:25-29 */
void X(execute_split_dft)(const X(plan) p, R *ri, R *ii, R *ro, R *io) { b2_execute(p->pln, ri, ii, ro, io); }
/* api/execute-dft-r2c.c:25-30, execute-split-dft-r2c.c */
void X(execute_dft_r2c)(const X(plan) p, R *in, C *out) { b2_execute(p->pln, in, NULL, (R *)out, (R *)out + 1); }
void X(execute_split_dft_r2c)(const X(plan) p, R *in, R *ro, R *io) { b2_execute(p->pln, in, NULL, ro, io); }
/* api/execute-dft-c2r.c:25-30, execute-split-dft-c2r.c */
void X(execute_dft_c2r)(const X(plan) p, C *in, R *out) { b2_execute(p->pln, (R *)in, (R *)in + 1, out, NULL); }
void X(execute_split_dft_c2r)(const X(plan) p, R *ri, R *ii, R *out) { b2_execute(p->pln, ri, ii, out, NULL); }
/* api/execute-r2r.c:25-29 */
void X(execute_r2r)(const X(plan) p, R *in, R *out) { b2_execute(p->pln, in, NULL, out, NULL); }

/* ----------------------------------------------------------------- lifetime */
/* api/apiplan.c:180-210 */
X(plan) X(copy_plan)(X(plan) p)
{
    if (p) __atomic_add_fetch(&p->pln->refcnt, 1, __ATOMIC_SEQ_CST);
    return p;
}
void X(destroy_plan)(X(plan) p)
{
    if (!p) return;
    if (__atomic_sub_fetch(&p->pln->refcnt, 1, __ATOMIC_SEQ_CST) == 0) {
        b2_plan_destroy(p->pln);
        free(p);
    }
}
/* api/the-planner.c:36-49 */
void X(cleanup)(void) { b2_wisdom_forget(); b2_tables_cleanup(); }
void X(forget_wisdom)(void) { b2_wisdom_forget(); }
void X(set_timelimit)(double t) { b2_timelimit = t; }

/* threads/api.c: the symbols exist so that threaded callers link; the GPU grid
   is the parallelism, so they are cheap shims */
static int g_nthreads = 1;
int X(init_threads)(void) { return 1; }
void X(plan_with_nthreads)(int n) { g_nthreads = n > 0 ? n : 1; }
int X(planner_nthreads)(void) { return g_nthreads; }
void X(cleanup_threads)(void) { X(cleanup)(); }
/* planner entry points always serialise on b2_planner_lock(), so there is nothing to switch on */
void X(make_planner_thread_safe)(void) {}
void X(threads_set_callback)(void (*parallel_loop)(void *(*work)(char *), char *jobdata, size_t elsize,
                                                   int njobs, void *data), void *data)
{
    (void)parallel_loop; (void)data;
}

/* ------------------------------------------------------------------- wisdom */
/* api/export-wisdom.c, export-wisdom-to-file.c, export-wisdom-to-string.c */
void X(export_wisdom)(write_char_func w, void *data) { b2_wisdom_export(w, data, PREC); }

static void file_emit(char c, void *d) { fputc(c, (FILE *)d); }
void X(export_wisdom_to_file)(FILE *f) { b2_wisdom_export(file_emit, f, PREC); }
int X(export_wisdom_to_filename)(const char *filename)
{
    FILE *f = fopen(filename, "w");
    int ok;
    if (!f) return 0;
    X(export_wisdom_to_file)(f);
    ok = !ferror(f);
    if (fclose(f)) ok = 0;
    return ok;
}
typedef struct { char *s; size_t n, cap; } strbuf;
static void str_emit(char c, void *d)
{
    strbuf *b = (strbuf *)d;
    if (b->n + 2 > b->cap) {
        size_t nc = b->cap ? 2 * b->cap : 256;
        char *ns = (char *)realloc(b->s, nc);
        if (!ns) return;
        b->s = ns; b->cap = nc;
    }
    b->s[b->n++] = c;
    b->s[b->n] = 0;
}
char *X(export_wisdom_to_string)(void)
{
    strbuf b; b.s = NULL; b.n = b.cap = 0;
    b2_wisdom_export(str_emit, &b, PREC);
    return b.s;                                  /* caller frees with free() */
}
/* api/import-wisdom.c, import-wisdom-from-file.c, import-wisdom-from-string.c, import-system-wisdom.c */
int X(import_wisdom)(read_char_func r, void *data) { return b2_wisdom_import(r, data, PREC); }
static int file_next(void *d) { return fgetc((FILE *)d); }
int X(import_wisdom_from_file)(FILE *f) { return b2_wisdom_import(file_next, f, PREC); }
int X(import_wisdom_from_filename)(const char *filename)
{
    FILE *f = fopen(filename, "r");
    int ok;
    if (!f) return 0;
    ok = X(import_wisdom_from_file)(f);
    if (fclose(f)) ok = 0;
    return ok;
}
static int str_next(void *d)
{
    const char **s = (const char **)d;
    if (!**s) return EOF;
    return (unsigned char)*(*s)++;
}
int X(import_wisdom_from_string)(const char *s) { return b2_wisdom_import(str_next, &s, PREC); }
int X(import_system_wisdom)(void)
{
#ifdef B2_SINGLE
    return X(import_wisdom_from_filename)("/etc/fftw/wisdomf_b200");
#else
    return X(import_wisdom_from_filename)("/etc/fftw/wisdom_b200");
#endif
}

/* ------------------------------------------------------------ introspection */
/* api/print-plan.c:23-53 */
void X(fprint_plan)(const X(plan) p, FILE *f) { b2_plan_print(p->pln, f); }
void X(print_plan)(const X(plan) p) { b2_plan_print(p->pln, stdout); }
char *X(sprint_plan)(const X(plan) p)
{
    char *buf = NULL;
    size_t len = 0;
    FILE *f = open_memstream(&buf, &len);
    if (!f) return NULL;
    b2_plan_print(p->pln, f);
    fclose(f);
    return buf;
}
/* api/flops.c:23-43 */
void X(flops)(const X(plan) p, double *add, double *mul, double *fmas)
{
    *add = p->pln->est_flops_add; *mul = p->pln->est_flops_mul; *fmas = p->pln->est_flops_fma;
}
double X(estimate_cost)(const X(plan) p)
{
    return p->pln->est_flops_add + p->pln->est_flops_mul + 2 * p->pln->est_flops_fma;
}
double X(cost)(const X(plan) p) { return p->pln->cost; }

/* ------------------------------------------------------------------- memory */
/* api/malloc.c: pinned host memory so staged executes run at full PCIe rate;
   falls back to ordinary aligned memory when no device is usable */
#define MAXPINNED 4096
static void *g_pinned[MAXPINNED];
static pthread_mutex_t g_pin_lock = PTHREAD_MUTEX_INITIALIZER;

void *X(malloc)(size_t n)
{
    void *p = NULL;
    int i;
    if (n >= 65536 && b2d_device_count() > 0) {
        p = b2d_malloc_host(n);
        if (p) {
            pthread_mutex_lock(&g_pin_lock);
            for (i = 0; i < MAXPINNED; ++i) if (!g_pinned[i]) { g_pinned[i] = p; break; }
            pthread_mutex_unlock(&g_pin_lock);
            if (i < MAXPINNED) return p;
            b2d_free_host(p);
            p = NULL;
        }
    }
    if (posix_memalign(&p, 64, n ? n : 1)) return NULL;
    return p;
}
R *X(alloc_real)(size_t n) { return (R *)X(malloc)(n * sizeof(R)); }
C *X(alloc_complex)(size_t n) { return (C *)X(malloc)(n * sizeof(C)); }
void X(free)(void *p)
{
    int i;
    if (!p) return;
    pthread_mutex_lock(&g_pin_lock);
    for (i = 0; i < MAXPINNED; ++i) if (g_pinned[i] == p) { g_pinned[i] = NULL; break; }
    pthread_mutex_unlock(&g_pin_lock);
    if (i < MAXPINNED) b2d_free_host(p); else free(p);
}
/* kernel/align.c:33-41 (SIMD builds: address modulo 16) */
int X(alignment_of)(R *p) { return (int)(((uintptr_t)p) % 16); }

const char X(version)[] = "fftw3_b200-1.0 (FFTW 3.3.11 API)"
#ifdef B2_SINGLE
    "-float"
#endif
    ;
const char X(cc)[] = "nvcc -gencode arch=compute_100a,code=sm_100a + gcc";
const char X(codelet_optim)[] = "";
