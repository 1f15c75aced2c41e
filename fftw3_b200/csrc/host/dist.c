/* dist.c -- slab-decomposed 3-D c2c over several GPUs (one process per GPU),
 * built as a composition of ordinary single-GPU plans.
 *
 * Algorithm = the reference's mpi/dft-rank-geq2-transposed.c:47-70:
 *   local transforms over the non-distributed dims, global transpose n0 <-> n1,
 *   transforms along n0 (now local); natural-order output adds the transpose
 *   back (mpi/dft-rank-geq2.c:40-59 via dft-rank1-bigvec.c:45-65).
 * What is different: the reference's transpose = local transpose + MPI_Alltoall
 * + local transpose (mpi/transpose-alltoall.c:49-100).  Here both local
 * transposes are folded into FFT passes:
 *   stage 0  Y: FFT along n1, in place (strided pass)
 *            X: FFT along n2 (contiguous rows), stored straight into the block
 *               layout of the exchange -- [dest][i0][k1'][k2] -- i.e. into the
 *               peers' buffers over NVLink, or into a local send buffer
 *   stage 1  Z: FFT along n0 reading the received blocks, which already form
 *               [n0][local_n1][n2]; written in place (natural order follows) or
 *               as [local_n1][n0][n2] into `local` (TRANSPOSED_OUT)
 *   stage 2     gather the blocks back into [local_n0][n1][n2]
 */
#include <stdlib.h>
#include <string.h>
#include "b2_internal.h"

typedef double C[2];

struct fftw_b200_dist_plan_s {
    int nranks, rank, nstages;
    b2_plan *y;              /* stage 0 */
    b2_plan **x;             /* stage 0, one per destination */
    b2_plan *z;              /* stage 1 */
    b2_plan **g;             /* stage 2, one per source */
};
typedef struct fftw_b200_dist_plan_s *dplan;

static int64_t blk(int64_t n, int p) { return (n + p - 1) / p; }
static int64_t share(int64_t n, int p, int r)
{
    int64_t b = blk(n, p), lo = b * r;
    if (lo >= n) return 0;
    return (n - lo < b) ? n - lo : b;
}

ptrdiff_t fftw_b200_dist_local_size_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                                       ptrdiff_t *local_n0, ptrdiff_t *local_0_start,
                                       ptrdiff_t *local_n1, ptrdiff_t *local_1_start)
{
    int64_t b0 = blk(n0, nranks), b1 = blk(n1, nranks);
    int64_t a = b0 * n1 * n2, b = b1 * n0 * n2;
    if (local_n0) *local_n0 = (ptrdiff_t)share(n0, nranks, rank);
    if (local_0_start) *local_0_start = (ptrdiff_t)(b0 * rank < n0 ? b0 * rank : n0);
    if (local_n1) *local_n1 = (ptrdiff_t)share(n1, nranks, rank);
    if (local_1_start) *local_1_start = (ptrdiff_t)(b1 * rank < n1 ? b1 * rank : n1);
    return (ptrdiff_t)(a > b ? a : b);
}

static void set_ptrs(b2_problem *q, double *in, double *out, int sign)
{
    if (sign < 0) { q->in0 = in; q->in1 = in + 1; q->out0 = out; q->out1 = out + 1; }
    else { q->in0 = in + 1; q->in1 = in; q->out0 = out + 1; q->out1 = out; }
}

static void dim(b2_tensor *t, int64_t n, int64_t is, int64_t os)
{
    t->d[t->rnk].n = n; t->d[t->rnk].is = is; t->d[t->rnk].os = os; t->rnk++;
}

void fftw_b200_dist_destroy_plan(dplan p)
{
    int i;
    if (!p) return;
    b2_plan_destroy(p->y);
    b2_plan_destroy(p->z);
    for (i = 0; i < p->nranks; ++i) {
        if (p->x) b2_plan_destroy(p->x[i]);
        if (p->g) b2_plan_destroy(p->g[i]);
    }
    free(p->x); free(p->g);
    free(p);
}

dplan fftw_b200_dist_plan_dft_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                                 C *local, C *zbuf, void *const *push_targets, void *const *pull_sources,
                                 int sign, unsigned flags)
{
    dplan p;
    b2_problem q;
    int d;
    int64_t b1 = blk(n1, nranks);
    int64_t ln0 = share(n0, nranks, rank), ln1 = share(n1, nranks, rank);
    if (n0 <= 0 || n1 <= 0 || n2 <= 0 || nranks < 1 || rank < 0 || rank >= nranks) return NULL;
    if (sign != -1 && sign != 1) return NULL;
    p = (dplan)calloc(1, sizeof *p);
    if (!p) return NULL;
    p->nranks = nranks; p->rank = rank;
    p->nstages = pull_sources ? 3 : 2;
    p->x = (b2_plan **)calloc((size_t)nranks, sizeof(b2_plan *));
    p->g = (b2_plan **)calloc((size_t)nranks, sizeof(b2_plan *));
    if (!p->x || !p->g) goto fail;

    /* Y: FFT along n1 in place on [ln0][n1][n2] */
    memset(&q, 0, sizeof q);
    q.prec = B2D_F64; q.kind = B2_C2C; q.flags = flags;
    b2_tensor_init(&q.sz, 0); b2_tensor_init(&q.vecsz, 0);
    dim(&q.sz, n1, 2 * n2, 2 * n2);
    dim(&q.vecsz, ln0, 2 * n1 * n2, 2 * n1 * n2);
    dim(&q.vecsz, n2, 2, 2);
    set_ptrs(&q, (double *)local, (double *)local, sign);
    p->y = b2_mkplan(&q);
    if (!p->y) goto fail;

    /* X: FFT along n2, rows (i0, k1 in block d) -> push_targets[d] as [i0][k1'][k2] */
    for (d = 0; d < nranks; ++d) {
        int64_t l1 = share(n1, nranks, d);
        memset(&q, 0, sizeof q);
        q.prec = B2D_F64; q.kind = B2_C2C; q.flags = flags;
        b2_tensor_init(&q.sz, 0); b2_tensor_init(&q.vecsz, 0);
        dim(&q.sz, n2, 2, 2);
        dim(&q.vecsz, ln0, 2 * n1 * n2, 2 * l1 * n2);
        dim(&q.vecsz, l1, 2 * n2, 2 * n2);
        set_ptrs(&q, (double *)local + 2 * d * b1 * n2, (double *)push_targets[d], sign);
        p->x[d] = b2_mkplan(&q);
        if (!p->x[d]) goto fail;
    }

    /* Z: FFT along n0 on zbuf = [n0][ln1][n2] */
    memset(&q, 0, sizeof q);
    q.prec = B2D_F64; q.kind = B2_C2C; q.flags = flags;
    b2_tensor_init(&q.sz, 0); b2_tensor_init(&q.vecsz, 0);
    if (pull_sources) {
        dim(&q.sz, n0, 2 * ln1 * n2, 2 * ln1 * n2);
        dim(&q.vecsz, ln1 * n2, 2, 2);
        set_ptrs(&q, (double *)zbuf, (double *)zbuf, sign);
    } else {
        /* TRANSPOSED_OUT: [n0][ln1][n2] -> local as [ln1][n0][n2] */
        dim(&q.sz, n0, 2 * ln1 * n2, 2 * n2);
        dim(&q.vecsz, ln1, 2 * n2, 2 * n0 * n2);
        dim(&q.vecsz, n2, 2, 2);
        set_ptrs(&q, (double *)zbuf, (double *)local, sign);
    }
    p->z = b2_mkplan(&q);
    if (!p->z) goto fail;

    /* gather back: block from rank s = [ln0][l1(s)][n2] -> local[i0][s*b1 + k1'][k2] */
    if (pull_sources) {
        for (d = 0; d < nranks; ++d) {
            int64_t l1 = share(n1, nranks, d);
            memset(&q, 0, sizeof q);
            q.prec = B2D_F64; q.kind = B2_C2C; q.flags = flags | B2F_ESTIMATE;
            b2_tensor_init(&q.sz, 0); b2_tensor_init(&q.vecsz, 0);
            dim(&q.vecsz, ln0, 2 * l1 * n2, 2 * n1 * n2);
            dim(&q.vecsz, l1 * n2, 2, 2);
            set_ptrs(&q, (double *)pull_sources[d], (double *)local + 2 * d * b1 * n2, -1);
            p->g[d] = b2_mkplan(&q);
            if (!p->g[d]) goto fail;
        }
    }
    return p;
fail:
    fftw_b200_dist_destroy_plan(p);
    return NULL;
}

int fftw_b200_dist_num_stages(const dplan p) { return p->nstages; }

static void run(b2_plan *pl)
{
    if (pl) b2_execute(pl, pl->prob.in0, pl->prob.in1, pl->prob.out0, pl->prob.out1);
}

void fftw_b200_dist_execute_stage(const dplan p, int stage)
{
    int d;
    if (stage == 0) {
        run(p->y);
        /* start with the block for the next rank so that the ranks do not all
           hammer the same destination at once */
        for (d = 0; d < p->nranks; ++d) run(p->x[(p->rank + 1 + d) % p->nranks]);
    } else if (stage == 1) {
        run(p->z);
    } else if (stage == 2) {
        for (d = 0; d < p->nranks; ++d) run(p->g[(p->rank + 1 + d) % p->nranks]);
    }
}

void *fftw_b200_device_malloc(size_t bytes) { return b2d_malloc(bytes); }
void fftw_b200_device_free(void *p) { b2d_free(p); }
int fftw_b200_ipc_export(void *devptr, unsigned char handle[64]) { return b2d_ipc_export(devptr, handle); }
void *fftw_b200_ipc_import(const unsigned char handle[64]) { return b2d_ipc_import(handle); }
void fftw_b200_ipc_close(void *devptr) { b2d_ipc_close(devptr); }
