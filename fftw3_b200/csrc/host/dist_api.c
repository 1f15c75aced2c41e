/* dist_api.c -- the fftw_mpi_* shaped interface for several GPUs, one process per GPU
 * (include/fftw3_b200_dist.h, "communicator interface").
 *
 * Reference: mpi/api.c:248-352 (local_size*), :560-648 (plan_many_dft, plan_dft, plan_dft_2d/3d),
 * mpi/fftw3-mpi.h:58-215, execution through the ordinary fftw_execute (mpi/api.c:889-907).
 *
 * The reference takes an MPI_Comm.  Here the launcher-specific part is reduced to ONE collective the caller
 * supplies in a fftw_b200_comm: a blocking all-gather of a few hundred bytes of host memory, used at plan
 * creation only (MPI_Allgather, torch.distributed.all_gather, ...).  Everything else is done by the library:
 * it allocates the exchange buffer, exports / imports the CUDA-IPC mappings of every rank's exchange buffer
 * and slab, builds the passes, and execution owns its synchronisation -- the ranks meet in a device-side
 * barrier (flags in peer-mapped memory, b2d_peer_barrier), no host round trip and no NCCL call.
 *
 * Algorithm = mpi/dft-rank-geq2-transposed.c:47-70 / dft-rank-geq2.c:40-59:
 *   local transform over dims 1..rnk-1  ->  global transpose n0 <-> n1  ->  transform along n0
 *   (-> transpose back unless FFTW_MPI_TRANSPOSED_OUT).
 * 3-D double transforms with howmany = 1 use the fused plans of dist.c (both exchanges ride on FFT pass stores);
 * every other shape (2-D, rank > 3, howmany > 1, single precision) uses the general path below: the local
 * transform is an ordinary plan, the transposes are strided copies straight into peer memory (the reference's
 * mpi/transpose-alltoall.c:49-100 with the all-to-all replaced by stores over NVLink).
 */
#include <stdlib.h>
#include <string.h>
#include "b2_internal.h"
#include "../../../include/fftw3_b200_dist.h"

#define MAXP B2D_MAX_PEERS
#define FLAG_BYTES 256            /* nranks x 8 bytes of barrier flags live in front of the exchange buffer */

struct fftw_b200_mpi_plan_s {
    int prec, rank, nranks, rnk, transposed_out, sign;
    int64_t n0, n1, R, ln0, s0, ln1, s1, b0, b1;
    void *in, *out;
    char *zalloc;                 /* owned allocation: [flags][exchange buffer] */
    char *zbuf;
    void *peer_z[MAXP], *peer_out[MAXP], *flags[MAXP];
    void *opened[2 * MAXP];
    int nopened;
    fftw_b200_dist_plan fused;    /* 3-D double howmany 1: the fused plan of dist.c */
    b2_plan *local, *scatter[MAXP], *z, *back[MAXP];
    unsigned long long epoch;
};

static int64_t blk(int64_t n, int p) { return (n + p - 1) / p; }
static int64_t share(int64_t n, int p, int r)
{
    int64_t b = blk(n, p), lo = b * r;
    if (lo >= n) return 0;
    return (n - lo < b) ? n - lo : b;
}
static size_t csize(int prec) { return prec == B2D_F32 ? 8 : 16; }

/* ------------------------------------------------------------------ local sizes (mpi/api.c:248-352) */
ptrdiff_t fftw_b200_mpi_local_size_many_transposed(int rnk, const ptrdiff_t *n, ptrdiff_t howmany,
                                                   ptrdiff_t block0, ptrdiff_t block1, const fftw_b200_comm *comm,
                                                   ptrdiff_t *local_n0, ptrdiff_t *local_0_start,
                                                   ptrdiff_t *local_n1, ptrdiff_t *local_1_start)
{
    int64_t rest = howmany, a, b, b0, b1, n1;
    int i, P, r;
    if (!comm || rnk < 1 || howmany < 0) return 0;
    P = comm->nranks; r = comm->rank;
    for (i = 0; i < rnk; ++i) if (n[i] <= 0) return 0;
    n1 = rnk > 1 ? n[1] : 1;
    b0 = block0 > 0 ? block0 : blk(n[0], P);       /* FFTW_MPI_DEFAULT_BLOCK = 0 (mpi/block.c:37-50) */
    b1 = block1 > 0 ? block1 : blk(n1, P);
    for (i = 2; i < rnk; ++i) rest *= n[i];
    if (local_n0) { int64_t lo = b0 * r; *local_n0 = (ptrdiff_t)(lo >= n[0] ? 0 : (n[0] - lo < b0 ? n[0] - lo : b0)); }
    if (local_0_start) *local_0_start = (ptrdiff_t)(b0 * r < n[0] ? b0 * r : n[0]);
    if (local_n1) { int64_t lo = b1 * r; *local_n1 = (ptrdiff_t)(lo >= n1 ? 0 : (n1 - lo < b1 ? n1 - lo : b1)); }
    if (local_1_start) *local_1_start = (ptrdiff_t)(b1 * r < n1 ? b1 * r : n1);
    a = b0 * n1 * rest; b = b1 * n[0] * rest;
    return (ptrdiff_t)(a > b ? a : b);
}

ptrdiff_t fftw_b200_mpi_local_size_many(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t block0,
                                        const fftw_b200_comm *comm, ptrdiff_t *local_n0, ptrdiff_t *local_0_start)
{
    return fftw_b200_mpi_local_size_many_transposed(rnk, n, howmany, block0, 0, comm, local_n0, local_0_start, NULL, NULL);
}

ptrdiff_t fftw_b200_mpi_local_size(int rnk, const ptrdiff_t *n, const fftw_b200_comm *comm,
                                   ptrdiff_t *local_n0, ptrdiff_t *local_0_start)
{
    return fftw_b200_mpi_local_size_many(rnk, n, 1, 0, comm, local_n0, local_0_start);
}

ptrdiff_t fftw_b200_mpi_local_size_2d(ptrdiff_t n0, ptrdiff_t n1, const fftw_b200_comm *comm,
                                      ptrdiff_t *local_n0, ptrdiff_t *local_0_start)
{
    ptrdiff_t n[2]; n[0] = n0; n[1] = n1;
    return fftw_b200_mpi_local_size(2, n, comm, local_n0, local_0_start);
}

ptrdiff_t fftw_b200_mpi_local_size_2d_transposed(ptrdiff_t n0, ptrdiff_t n1, const fftw_b200_comm *comm,
                                                 ptrdiff_t *local_n0, ptrdiff_t *local_0_start,
                                                 ptrdiff_t *local_n1, ptrdiff_t *local_1_start)
{
    ptrdiff_t n[2]; n[0] = n0; n[1] = n1;
    return fftw_b200_mpi_local_size_many_transposed(2, n, 1, 0, 0, comm, local_n0, local_0_start, local_n1, local_1_start);
}

ptrdiff_t fftw_b200_mpi_local_size_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, const fftw_b200_comm *comm,
                                      ptrdiff_t *local_n0, ptrdiff_t *local_0_start)
{
    ptrdiff_t n[3]; n[0] = n0; n[1] = n1; n[2] = n2;
    return fftw_b200_mpi_local_size(3, n, comm, local_n0, local_0_start);
}

ptrdiff_t fftw_b200_mpi_local_size_3d_transposed(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, const fftw_b200_comm *comm,
                                                 ptrdiff_t *local_n0, ptrdiff_t *local_0_start,
                                                 ptrdiff_t *local_n1, ptrdiff_t *local_1_start)
{
    ptrdiff_t n[3]; n[0] = n0; n[1] = n1; n[2] = n2;
    return fftw_b200_mpi_local_size_many_transposed(3, n, 1, 0, 0, comm, local_n0, local_0_start, local_n1, local_1_start);
}

/* ------------------------------------------------------------------ helpers */
static void dim(b2_tensor *t, int64_t n, int64_t is, int64_t os)
{
    t->d[t->rnk].n = n; t->d[t->rnk].is = is; t->d[t->rnk].os = os; t->rnk++;
}

static void problem(b2_problem *q, int prec, unsigned flags, void *in, void *out, int sign)
{
    size_t rs = prec == B2D_F32 ? 4 : 8;
    memset(q, 0, sizeof *q);
    q->prec = prec; q->kind = B2_C2C; q->flags = flags;
    b2_tensor_init(&q->sz, 0); b2_tensor_init(&q->vecsz, 0);
    if (sign < 0) { q->in0 = in; q->in1 = (char *)in + rs; q->out0 = out; q->out1 = (char *)out + rs; }
    else { q->in0 = (char *)in + rs; q->in1 = in; q->out0 = (char *)out + rs; q->out1 = out; }
}

typedef struct { unsigned char hz[64], ho[64]; int64_t oz, oo; int ok; int pad; } exch;

void fftw_b200_mpi_destroy_plan(fftw_b200_mpi_plan p)
{
    int i;
    if (!p) return;
    if (p->fused) fftw_b200_dist_destroy_plan(p->fused);
    b2_plan_destroy(p->local);
    b2_plan_destroy(p->z);
    for (i = 0; i < MAXP; ++i) { b2_plan_destroy(p->scatter[i]); b2_plan_destroy(p->back[i]); }
    b2d_sync();
    for (i = 0; i < p->nopened; ++i) b2d_ipc_close(p->opened[i]);
    b2d_free(p->zalloc);
    free(p);
}

/* ------------------------------------------------------------------ planning (mpi/api.c:560-648) */
static fftw_b200_mpi_plan mkplan(int prec, int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t block, ptrdiff_t tblock,
                                 void *in, void *out, const fftw_b200_comm *comm, int sign, unsigned flags)
{
    fftw_b200_mpi_plan p;
    exch mine, *all = NULL;
    int i, d, P, r, ok = 1;
    int64_t R = howmany, alloc;
    size_t cs = csize(prec);
    unsigned pflags = flags & ~(FFTW_MPI_TRANSPOSED_OUT | FFTW_MPI_TRANSPOSED_IN | FFTW_MPI_SCRAMBLED_IN | FFTW_MPI_SCRAMBLED_OUT);
    if (!comm || !comm->allgather || rnk < 2 || rnk > 8 || howmany < 1 || !in || !out) return NULL;
    if (block || tblock) return NULL;                         /* default block sizes only */
    if (flags & (FFTW_MPI_TRANSPOSED_IN | FFTW_MPI_SCRAMBLED_IN | FFTW_MPI_SCRAMBLED_OUT)) return NULL;
    if (sign != -1 && sign != 1) return NULL;
    P = comm->nranks; r = comm->rank;
    if (P < 1 || P > MAXP || r < 0 || r >= P) return NULL;
    for (i = 0; i < rnk; ++i) if (n[i] <= 0) return NULL;
    if (b2d_pointer_is_device(in) != 1 || b2d_pointer_is_device(out) != 1) return NULL;
    p = (fftw_b200_mpi_plan)calloc(1, sizeof *p);
    if (!p) return NULL;
    for (i = 2; i < rnk; ++i) R *= n[i];
    p->prec = prec; p->rank = r; p->nranks = P; p->rnk = rnk; p->sign = sign;
    p->transposed_out = (flags & FFTW_MPI_TRANSPOSED_OUT) != 0;
    p->n0 = n[0]; p->n1 = n[1]; p->R = R;
    p->b0 = blk(n[0], P); p->b1 = blk(n[1], P);
    p->ln0 = share(n[0], P, r); p->ln1 = share(n[1], P, r);
    p->s0 = p->b0 * r < n[0] ? p->b0 * r : n[0];
    p->s1 = p->b1 * r < n[1] ? p->b1 * r : n[1];
    p->in = in; p->out = out;
    alloc = p->b0 * n[1] * R;
    if (p->b1 * n[0] * R > alloc) alloc = p->b1 * n[0] * R;
    p->zalloc = (char *)b2d_malloc(FLAG_BYTES + (size_t)(alloc > 0 ? alloc : 1) * cs);
    memset(&mine, 0, sizeof mine);
    if (p->zalloc) {
        p->zbuf = p->zalloc + FLAG_BYTES;
        b2d_memset(p->zalloc, 0, FLAG_BYTES);
        b2d_sync();
        mine.oo = b2d_alloc_offset(out);
        mine.ok = mine.oo >= 0 && !b2d_ipc_export(p->zalloc, mine.hz) && !b2d_ipc_export((char *)out - mine.oo, mine.ho);
    }
    all = (exch *)calloc((size_t)P, sizeof *all);
    if (!all || comm->allgather(comm->ctx, &mine, all, sizeof mine)) { free(all); fftw_b200_mpi_destroy_plan(p); return NULL; }
    for (d = 0; d < P; ++d) if (!all[d].ok) ok = 0;            /* all ranks see the same verdict */
    if (ok) {
        for (d = 0; d < P && ok; ++d) {
            if (d == r) { p->flags[d] = p->zalloc; p->peer_z[d] = p->zbuf; p->peer_out[d] = out; continue; }
            {
                char *z = (char *)b2d_ipc_import(all[d].hz), *o;
                if (!z) { ok = 0; break; }
                p->opened[p->nopened++] = z;
                o = (char *)b2d_ipc_import(all[d].ho);
                if (!o) { ok = 0; break; }
                p->opened[p->nopened++] = o;
                p->flags[d] = z; p->peer_z[d] = z + FLAG_BYTES; p->peer_out[d] = o + all[d].oo;
            }
        }
    }
    free(all);
    if (!ok) goto fail_collective;

    /* fused plans of dist.c: 3-D, double, one transform, in place, same pointer semantics */
    if (rnk == 3 && howmany == 1 && prec == B2D_F64 && in == out && !getenv("FFTW3_B200_MPI_GENERAL")) {
        void *push[MAXP];
        for (d = 0; d < P; ++d) push[d] = (char *)p->peer_z[d] + cs * (size_t)(p->s0 * share(n[1], P, d) * n[2]);
        if (!p->transposed_out)
            p->fused = fftw_b200_dist_plan_dft_3d_push(n[0], n[1], n[2], r, P, (fftw_complex *)out, (fftw_complex *)p->zbuf,
                                                       push, p->peer_out, sign, pflags);
        else
            p->fused = fftw_b200_dist_plan_dft_3d(n[0], n[1], n[2], r, P, (fftw_complex *)out, (fftw_complex *)p->zbuf,
                                                  push, NULL, sign, pflags);
    }
    {
        /* every rank must take the same path: agree on it */
        int have = p->fused != NULL, *every = (int *)calloc((size_t)P, sizeof(int)), same = 1;
        if (!every || comm->allgather(comm->ctx, &have, every, sizeof have)) { free(every); goto fail; }
        for (d = 0; d < P; ++d) if (!every[d]) same = 0;
        free(every);
        if (!same && p->fused) { fftw_b200_dist_destroy_plan(p->fused); p->fused = NULL; }
    }
    if (!p->fused) {
        b2_problem q;
        int64_t inner = 2 * R;           /* reals per (i0, i1) row */
        /* local transform over dims 1 .. rnk-1, vector of `howmany` interleaved transforms */
        if (p->ln0 > 0) {
            int64_t st = 2 * howmany;
            problem(&q, prec, pflags, in, out, sign);
            for (i = rnk - 1; i >= 1; --i) { q.sz.d[i - 1].n = n[i]; q.sz.d[i - 1].is = q.sz.d[i - 1].os = st; st *= n[i]; }
            q.sz.rnk = rnk - 1;
            dim(&q.vecsz, p->ln0, n[1] * inner, n[1] * inner);
            if (howmany > 1) dim(&q.vecsz, howmany, 2, 2);
            p->local = b2_mkplan(&q);
            if (!p->local) ok = 0;
            /* scatter: my rows of column block d -> rank d's exchange buffer [n0][ln1(d)][R] at plane s0 */
            for (d = 0; d < P && ok; ++d) {
                int64_t l1 = share(n[1], P, d);
                if (!l1) continue;
                problem(&q, prec, pflags | B2F_ESTIMATE, (char *)out + cs / 2 * (size_t)(p->b1 * d * inner),
                        (char *)p->peer_z[d] + cs / 2 * (size_t)(p->s0 * l1 * inner), -1);
                dim(&q.vecsz, p->ln0, n[1] * inner, l1 * inner);
                dim(&q.vecsz, l1 * R, 2, 2);
                p->scatter[d] = b2_mkplan(&q);
                if (!p->scatter[d]) ok = 0;
            }
        }
        if (p->ln1 > 0 && ok) {
            problem(&q, prec, pflags, p->zbuf, p->zbuf, sign);
            dim(&q.sz, n[0], p->ln1 * inner, p->ln1 * inner);
            dim(&q.vecsz, p->ln1 * R, 2, 2);
            p->z = b2_mkplan(&q);
            if (!p->z) ok = 0;
            if (p->transposed_out && ok) {
                /* [n0][ln1][R] -> out as [ln1][n0][R] */
                problem(&q, prec, pflags | B2F_ESTIMATE, p->zbuf, out, -1);
                dim(&q.vecsz, p->ln1, inner, n[0] * inner);
                dim(&q.vecsz, n[0], p->ln1 * inner, inner);
                dim(&q.vecsz, R, 2, 2);
                p->back[0] = b2_mkplan(&q);
                if (!p->back[0]) ok = 0;
            } else for (d = 0; d < P && ok; ++d) {
                /* rows of owner d -> its slab [ln0(d)][n1][R] at column s1 */
                int64_t l0 = share(n[0], P, d);
                if (!l0) continue;
                problem(&q, prec, pflags | B2F_ESTIMATE, p->zbuf + cs / 2 * (size_t)(p->b0 * d * p->ln1 * inner),
                        (char *)p->peer_out[d] + cs / 2 * (size_t)(p->s1 * inner), -1);
                dim(&q.vecsz, l0, p->ln1 * inner, n[1] * inner);
                dim(&q.vecsz, p->ln1 * R, 2, 2);
                p->back[d] = b2_mkplan(&q);
                if (!p->back[d]) ok = 0;
            }
        }
    }
    {
        /* collective verdict; doubles as the barrier after which peers may write our flags and buffers */
        int *every = (int *)calloc((size_t)P, sizeof(int));
        if (!every || comm->allgather(comm->ctx, &ok, every, sizeof ok)) { free(every); goto fail; }
        for (d = 0; d < P; ++d) if (!every[d]) ok = 0;
        free(every);
    }
    if (!ok) goto fail;
    return p;
fail_collective:
    {
        /* keep the collective sequence of the successful path so that no rank blocks */
        int zero = 0, *every = (int *)calloc((size_t)P, sizeof(int));
        if (every) { comm->allgather(comm->ctx, &zero, every, sizeof zero); comm->allgather(comm->ctx, &zero, every, sizeof zero); }
        free(every);
    }
fail:
    fftw_b200_mpi_destroy_plan(p);
    return NULL;
}

#define DEFINE_API(PFX, PREC, CT)                                                                                       \
    fftw_b200_mpi_plan PFX##plan_many_dft(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t block,              \
                                          ptrdiff_t tblock, CT *in, CT *out, const fftw_b200_comm *comm, int sign,       \
                                          unsigned flags)                                                               \
    { return mkplan(PREC, rnk, n, howmany, block, tblock, in, out, comm, sign, flags); }                                \
    fftw_b200_mpi_plan PFX##plan_dft(int rnk, const ptrdiff_t *n, CT *in, CT *out, const fftw_b200_comm *comm,           \
                                     int sign, unsigned flags)                                                          \
    { return mkplan(PREC, rnk, n, 1, 0, 0, in, out, comm, sign, flags); }                                               \
    fftw_b200_mpi_plan PFX##plan_dft_2d(ptrdiff_t n0, ptrdiff_t n1, CT *in, CT *out, const fftw_b200_comm *comm,         \
                                        int sign, unsigned flags)                                                       \
    { ptrdiff_t n[2]; n[0] = n0; n[1] = n1; return mkplan(PREC, 2, n, 1, 0, 0, in, out, comm, sign, flags); }           \
    fftw_b200_mpi_plan PFX##plan_dft_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, CT *in, CT *out,                       \
                                        const fftw_b200_comm *comm, int sign, unsigned flags)                           \
    { ptrdiff_t n[3]; n[0] = n0; n[1] = n1; n[2] = n2; return mkplan(PREC, 3, n, 1, 0, 0, in, out, comm, sign, flags); }

DEFINE_API(fftw_b200_mpi_, B2D_F64, fftw_complex)
DEFINE_API(fftwf_b200_mpi_, B2D_F32, fftwf_complex)

/* ------------------------------------------------------------------ execution */
static void run(b2_plan *pl)
{
    if (pl) b2_execute_ex(pl, pl->prob.in0, pl->prob.in1, pl->prob.out0, pl->prob.out1, 1);
}

static void barrier(fftw_b200_mpi_plan p)
{
    b2d_peer_barrier(p->flags, p->rank, p->nranks, ++p->epoch);
}

/* One distributed transform; returns when the local result is complete (or, in async mode, once everything is
   enqueued on the launch stream).  Collective: every rank must call it. */
void fftw_b200_mpi_execute(fftw_b200_mpi_plan p)
{
    int d;
    if (!p) return;
    /* the peers may still be reading their exchange buffers / writing our slab from the previous call */
    if (p->fused) {
        fftw_b200_dist_execute_stage(p->fused, 0);
        barrier(p);                                     /* every block has landed in my exchange buffer */
        fftw_b200_dist_execute_stage(p->fused, 1);
        barrier(p);                                     /* every rank's rows have landed in my slab (or: peers are done) */
    } else {
        run(p->local);
        for (d = 0; d < p->nranks; ++d) run(p->scatter[(p->rank + 1 + d) % p->nranks]);
        barrier(p);
        run(p->z);
        for (d = 0; d < p->nranks; ++d) run(p->back[(p->rank + 1 + d) % p->nranks]);
        barrier(p);
    }
    if (!b2_async_mode) b2d_sync();
}
