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
    void *opened[3 * MAXP];
    int nopened;
    fftw_b200_dist_plan fused;    /* 3-D double howmany 1: the fused plan of dist.c */
    b2_plan *local, *scatter[MAXP], *z, *back[MAXP];
    /* kind 1 (six-step 1-D) and 2 (transpose) */
    int kind;
    char *z2alloc;                /* second owned buffer (six-step: rows [lr][m]) */
    void *peer_z2[MAXP];
    b2_plan *t3[MAXP];
    unsigned long long epoch;
};

static int64_t blk(int64_t n, int p) { return (n + p - 1) / p; }
static int64_t share(int64_t n, int p, int r)
{
    int64_t b = blk(n, p), lo = b * r;
    if (lo >= n) return 0;
    return (n - lo < b) ? n - lo : b;
}
/* share of rank r under an explicit block size (fftw_mpi's block / tblock arguments, mpi/block.c:52-70) */
static int64_t shareb(int64_t n, int64_t b, int r)
{
    int64_t lo = b * r;
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

typedef struct { unsigned char hz[64], ho[64], hz2[64]; int64_t oz, oo; int ok; int pad; } exch;

void fftw_b200_mpi_destroy_plan(fftw_b200_mpi_plan p)
{
    int i;
    if (!p) return;
    if (p->fused) fftw_b200_dist_destroy_plan(p->fused);
    b2_plan_destroy(p->local);
    b2_plan_destroy(p->z);
    for (i = 0; i < MAXP; ++i) { b2_plan_destroy(p->scatter[i]); b2_plan_destroy(p->back[i]); b2_plan_destroy(p->t3[i]); }
    b2d_sync();
    for (i = 0; i < p->nopened; ++i) b2d_ipc_close(p->opened[i]);
    b2d_free(p->zalloc);
    b2d_free(p->z2alloc);
    free(p);
}

/* Allocate the exchange buffer(s), exchange CUDA-IPC handles of them and of the allocation `out` lives in, and map
   every peer's.  Collective (one all-gather); returns 1 when every rank succeeded, 0 otherwise (same verdict on
   every rank).  zbytes / z2bytes exclude the flag area. */
static int setup_peers(fftw_b200_mpi_plan p, const fftw_b200_comm *comm, void *out, size_t zbytes, size_t z2bytes)
{
    exch mine, *all;
    int d, P = p->nranks, r = p->rank, ok = 1;
    memset(&mine, 0, sizeof mine);
    p->zalloc = (char *)b2d_malloc(FLAG_BYTES + (zbytes ? zbytes : 16));
    if (z2bytes) p->z2alloc = (char *)b2d_malloc(z2bytes);
    if (p->zalloc && (!z2bytes || p->z2alloc)) {
        p->zbuf = p->zalloc + FLAG_BYTES;
        b2d_memset(p->zalloc, 0, FLAG_BYTES);
        b2d_sync();
        mine.oo = b2d_alloc_offset(out);
        mine.ok = mine.oo >= 0 && !b2d_ipc_export(p->zalloc, mine.hz) && !b2d_ipc_export((char *)out - mine.oo, mine.ho) &&
                  (!z2bytes || !b2d_ipc_export(p->z2alloc, mine.hz2));
    }
    all = (exch *)calloc((size_t)P, sizeof *all);
    if (!all || comm->allgather(comm->ctx, &mine, all, sizeof mine)) { free(all); return 0; }
    for (d = 0; d < P; ++d) if (!all[d].ok) ok = 0;
    for (d = 0; d < P && ok; ++d) {
        char *z, *o;
        if (d == r) { p->flags[d] = p->zalloc; p->peer_z[d] = p->zbuf; p->peer_out[d] = out; p->peer_z2[d] = p->z2alloc; continue; }
        z = (char *)b2d_ipc_import(all[d].hz);
        if (!z) { ok = 0; break; }
        p->opened[p->nopened++] = z;
        o = (char *)b2d_ipc_import(all[d].ho);
        if (!o) { ok = 0; break; }
        p->opened[p->nopened++] = o;
        p->flags[d] = z; p->peer_z[d] = z + FLAG_BYTES; p->peer_out[d] = o + all[d].oo;
        if (z2bytes) {
            char *z2 = (char *)b2d_ipc_import(all[d].hz2);
            if (!z2) { ok = 0; break; }
            p->opened[p->nopened++] = z2;
            p->peer_z2[d] = z2;
        }
    }
    free(all);
    return ok;
}

/* collective verdict on `ok`; doubles as the barrier after which peers may write our flags and buffers */
static int agree(const fftw_b200_comm *comm, int ok)
{
    int d, *every = (int *)calloc((size_t)comm->nranks, sizeof(int));
    if (!every || comm->allgather(comm->ctx, &ok, every, sizeof ok)) { free(every); return 0; }
    for (d = 0; d < comm->nranks; ++d) if (!every[d]) ok = 0;
    free(every);
    return ok;
}

/* ------------------------------------------------------------------ planning (mpi/api.c:560-648) */
static fftw_b200_mpi_plan mkplan(int prec, int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t block, ptrdiff_t tblock,
                                 void *in, void *out, const fftw_b200_comm *comm, int sign, unsigned flags)
{
    fftw_b200_mpi_plan p;
    int i, d, P, r, ok = 1;
    int64_t R = howmany, alloc;
    size_t cs = csize(prec);
    unsigned pflags = flags & ~(FFTW_MPI_TRANSPOSED_OUT | FFTW_MPI_TRANSPOSED_IN | FFTW_MPI_SCRAMBLED_IN | FFTW_MPI_SCRAMBLED_OUT);
    ptrdiff_t nswap[8];
    int64_t B0, B1;
    if (!comm || !comm->allgather || rnk < 2 || rnk > 8 || howmany < 1 || !in || !out) return NULL;
    if (block < 0 || tblock < 0) return NULL;
    if (flags & (FFTW_MPI_SCRAMBLED_IN | FFTW_MPI_SCRAMBLED_OUT)) return NULL;
    if (flags & FFTW_MPI_TRANSPOSED_IN) {
        /* input laid out [local_n1][n0][...] (mpi/fftw3-mpi.h:212-215, mpi/dft-rank-geq2-transposed.c): the
           multi-dimensional DFT does not care which of its dimensions is called the first, so this is the same plan
           with the first two dimensions swapped -- natural-order output of the user's problem is the
           TRANSPOSED_OUT layout of the swapped one and vice versa */
        for (i = 0; i < rnk; ++i) nswap[i] = n[i];
        nswap[0] = n[1]; nswap[1] = n[0];
        n = nswap;
        flags = (flags & ~FFTW_MPI_TRANSPOSED_IN) ^ FFTW_MPI_TRANSPOSED_OUT;
        { ptrdiff_t t = block; block = tblock; tblock = t; }
    }
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
    /* block sizes: FFTW_MPI_DEFAULT_BLOCK (0) = ceil(n / P); a caller's own must still cover the dimension */
    B0 = block > 0 ? block : blk(n[0], P);
    B1 = tblock > 0 ? tblock : blk(n[1], P);
    if (B0 * P < n[0] || B1 * P < n[1]) { free(p); return NULL; }
    p->b0 = B0; p->b1 = B1;
    p->ln0 = shareb(n[0], B0, r); p->ln1 = shareb(n[1], B1, r);
    p->s0 = p->b0 * r < n[0] ? p->b0 * r : n[0];
    p->s1 = p->b1 * r < n[1] ? p->b1 * r : n[1];
    p->in = in; p->out = out;
    alloc = p->b0 * n[1] * R;
    if (p->b1 * n[0] * R > alloc) alloc = p->b1 * n[0] * R;
    ok = setup_peers(p, comm, out, (size_t)(alloc > 0 ? alloc : 1) * cs, 0);
    if (!ok) goto fail_collective;

    /* fused plans of dist.c: 3-D, double, one transform, in place, same pointer semantics */
    if (rnk == 3 && howmany == 1 && prec == B2D_F64 && in == out && B0 == blk(n[0], P) && B1 == blk(n[1], P) &&
        !getenv("FFTW3_B200_MPI_GENERAL")) {
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
                int64_t l1 = shareb(n[1], p->b1, d);
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
                int64_t l0 = shareb(n[0], p->b0, d);
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
    if (!agree(comm, ok)) goto fail;
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

/* ------------------------------------------------------------------ distributed transposes (mpi/api.c:521-556)
   fftw_mpi_plan_many_transpose: an n0 x n1 matrix of `howmany`-tuples of REAL numbers, rows block-distributed,
   becomes the n1 x n0 matrix, rows block-distributed (mpi/transpose-alltoall.c:49-100).  One strided copy per
   destination writes the transposed block straight into the peer's output (in place: through the exchange
   buffer and a local copy back), then a device-side barrier. */
static void rproblem(b2_problem *q, int prec, unsigned flags, void *in, void *out)
{
    memset(q, 0, sizeof *q);
    q->prec = prec; q->kind = B2_R2R; q->flags = flags | B2F_ESTIMATE;
    b2_tensor_init(&q->sz, 0); b2_tensor_init(&q->vecsz, 0);
    q->in0 = in; q->out0 = out;
}

static fftw_b200_mpi_plan mktranspose(int prec, ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t howmany, ptrdiff_t block0,
                                      ptrdiff_t block1, void *in, void *out, const fftw_b200_comm *comm, unsigned flags)
{
    fftw_b200_mpi_plan p;
    b2_problem q;
    int d, P, r, ok = 1, inplace = (in == out);
    size_t rs = prec == B2D_F32 ? 4 : 8;
    int64_t hm = howmany, alloc;
    if (!comm || !comm->allgather || n0 <= 0 || n1 <= 0 || howmany < 1 || !in || !out || block0 < 0 || block1 < 0) return NULL;
    P = comm->nranks; r = comm->rank;
    if (P < 1 || P > MAXP || r < 0 || r >= P) return NULL;
    if (b2d_pointer_is_device(in) != 1 || b2d_pointer_is_device(out) != 1) return NULL;
    p = (fftw_b200_mpi_plan)calloc(1, sizeof *p);
    if (!p) return NULL;
    p->kind = 2; p->prec = prec; p->rank = r; p->nranks = P;
    p->n0 = n0; p->n1 = n1; p->R = hm;
    p->b0 = block0 > 0 ? block0 : blk(n0, P); p->b1 = block1 > 0 ? block1 : blk(n1, P);
    if (p->b0 * P < n0 || p->b1 * P < n1) { free(p); return NULL; }
    p->ln0 = shareb(n0, p->b0, r); p->ln1 = shareb(n1, p->b1, r);
    p->s0 = p->b0 * r < n0 ? p->b0 * r : n0; p->s1 = p->b1 * r < n1 ? p->b1 * r : n1;
    p->in = in; p->out = out;
    alloc = p->b1 * n0 * hm;
    ok = setup_peers(p, comm, out, (size_t)(alloc > 0 ? alloc : 1) * rs, 0);
    if (ok && p->ln0 > 0) {
        for (d = 0; d < P && ok; ++d) {
            /* my rows, column block d -> rank d's [ln1(d)][n0][hm] at column s0 */
            int64_t l1 = shareb(n1, p->b1, d);
            char *dst = (char *)(inplace ? p->peer_z[d] : p->peer_out[d]) + rs * (size_t)(p->s0 * hm);
            if (!l1) continue;
            rproblem(&q, prec, flags, (char *)in + rs * (size_t)(p->b1 * d * hm), dst);
            dim(&q.vecsz, p->ln0, n1 * hm, hm);
            dim(&q.vecsz, l1, hm, n0 * hm);
            if (hm > 1) dim(&q.vecsz, hm, 1, 1);
            p->scatter[d] = b2_mkplan(&q);
            if (!p->scatter[d]) ok = 0;
        }
    }
    if (ok && inplace && p->ln1 > 0) {
        rproblem(&q, prec, flags, p->zbuf, out);
        dim(&q.vecsz, p->ln1 * n0 * hm, 1, 1);
        p->back[0] = b2_mkplan(&q);
        if (!p->back[0]) ok = 0;
    }
    if (!agree(comm, ok)) { fftw_b200_mpi_destroy_plan(p); return NULL; }
    return p;
}

fftw_b200_mpi_plan fftw_b200_mpi_plan_many_transpose(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t howmany, ptrdiff_t block0,
                                                     ptrdiff_t block1, double *in, double *out,
                                                     const fftw_b200_comm *comm, unsigned flags)
{ return mktranspose(B2D_F64, n0, n1, howmany, block0, block1, in, out, comm, flags); }
fftw_b200_mpi_plan fftw_b200_mpi_plan_transpose(ptrdiff_t n0, ptrdiff_t n1, double *in, double *out,
                                                const fftw_b200_comm *comm, unsigned flags)
{ return mktranspose(B2D_F64, n0, n1, 1, 0, 0, in, out, comm, flags); }
fftw_b200_mpi_plan fftwf_b200_mpi_plan_many_transpose(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t howmany, ptrdiff_t block0,
                                                      ptrdiff_t block1, float *in, float *out,
                                                      const fftw_b200_comm *comm, unsigned flags)
{ return mktranspose(B2D_F32, n0, n1, howmany, block0, block1, in, out, comm, flags); }
fftw_b200_mpi_plan fftwf_b200_mpi_plan_transpose(ptrdiff_t n0, ptrdiff_t n1, float *in, float *out,
                                                 const fftw_b200_comm *comm, unsigned flags)
{ return mktranspose(B2D_F32, n0, n1, 1, 0, 0, in, out, comm, flags); }

/* ------------------------------------------------------------------ distributed 1-D transform
   fftw_mpi_plan_dft_1d (mpi/dft-rank1.c:81-148,224-340, mpi/choose-radix.c:50): n = r * m, the "six-step"
   algorithm with three global transposes.  x is viewed as the r x m matrix [j1][j2] (j = j1 m + j2), rows
   block-distributed:
     T1  every rank sends the columns of block d to rank d: Z1 = [j1 (all r)][j2 in my block]   (plain row segments)
     A   FFT_r along j1 (strided pass over Z1, in place) with the twiddle exp(-2 pi i k1 j2 / n) in its store
     T2  rows k1 of block d go to rank d: Z2 = [k1 in my block][j2 (all m)]
     B   FFT_m along the contiguous rows of Z2: element (k1, k2) is X[k1 + r k2]
     T3  transposing copies: rank d receives [k2 in its block][k1 (all r)] = natural order, rows block-distributed
         over m (FFTW_MPI_SCRAMBLED_OUT: skipped, the output stays [k1][k2] over this rank's k1 block)
   Each transpose is a set of strided copies into peer memory followed by a device-side barrier. */
static int64_t choose_r(int64_t n, int prec)
{
    int64_t d, best = -1;
    int radix[64];
    for (d = 2; d * d <= n; ++d) {
        if (n % d) continue;
        if (d <= b2_max_single_pass(prec) && b2_factorize(d, prec, 0, radix) > 0) best = d;        /* closest to sqrt(n) from below */
    }
    return best;
}

ptrdiff_t fftw_b200_mpi_local_size_1d(ptrdiff_t n0, const fftw_b200_comm *comm, int sign, unsigned flags,
                                      ptrdiff_t *local_ni, ptrdiff_t *local_i_start,
                                      ptrdiff_t *local_no, ptrdiff_t *local_o_start)
{
    int64_t r, m, br, bm, lr, lm, sr, sm, a, b;
    int P, k;
    (void)sign;
    if (!comm || n0 < 4) return 0;
    r = choose_r(n0, B2D_F64);
    if (r < 0) return 0;                       /* n0 must be composite (as in the reference, mpi/dft-rank1.c:285) */
    m = n0 / r; P = comm->nranks; k = comm->rank;
    br = blk(r, P); bm = blk(m, P);
    lr = share(r, P, k); lm = share(m, P, k);
    sr = br * k < r ? br * k : r; sm = bm * k < m ? bm * k : m;
    if (local_ni) *local_ni = (ptrdiff_t)(lr * m);
    if (local_i_start) *local_i_start = (ptrdiff_t)(sr * m);
    if (flags & (FFTW_MPI_SCRAMBLED_OUT | FFTW_MPI_SCRAMBLED_IN)) {
        /* scrambled output stays [k1 in my block][k2]; a scrambled-input transform ends with the rows j1 of the
           r x m view on their owners: the same distribution as the input */
        if (local_no) *local_no = (ptrdiff_t)(lr * m);
        if (local_o_start) *local_o_start = (ptrdiff_t)(sr * m);
    } else {
        if (local_no) *local_no = (ptrdiff_t)(lm * r);
        if (local_o_start) *local_o_start = (ptrdiff_t)(sm * r);
    }
    a = br * m; b = bm * r;
    return (ptrdiff_t)(a > b ? a : b);
}

static fftw_b200_mpi_plan mkplan1d(int prec, ptrdiff_t n0, void *in, void *out, const fftw_b200_comm *comm, int sign, unsigned flags)
{
    fftw_b200_mpi_plan p;
    b2_problem q;
    int d, P, k, ok = 1, scrambled = (flags & FFTW_MPI_SCRAMBLED_OUT) != 0, scr_in = (flags & FFTW_MPI_SCRAMBLED_IN) != 0;
    size_t cs = csize(prec);
    unsigned pflags = flags & ~(FFTW_MPI_TRANSPOSED_OUT | FFTW_MPI_TRANSPOSED_IN | FFTW_MPI_SCRAMBLED_IN | FFTW_MPI_SCRAMBLED_OUT);
    int64_t r, m, br, bm, lr, lm, sr, sm;
    if (!comm || !comm->allgather || n0 < 4 || !in || !out || (sign != -1 && sign != 1)) return NULL;
    if (flags & (FFTW_MPI_TRANSPOSED_IN | FFTW_MPI_TRANSPOSED_OUT)) return NULL;
    if (scr_in && scrambled) return NULL;
    P = comm->nranks; k = comm->rank;
    if (P < 1 || P > MAXP || k < 0 || k >= P) return NULL;
    if (b2d_pointer_is_device(in) != 1 || b2d_pointer_is_device(out) != 1) return NULL;
    r = choose_r(n0, prec);
    if (r < 0) return NULL;
    m = n0 / r;
    br = blk(r, P); bm = blk(m, P);
    lr = share(r, P, k); lm = share(m, P, k);
    sr = br * k < r ? br * k : r; sm = bm * k < m ? bm * k : m;
    p = (fftw_b200_mpi_plan)calloc(1, sizeof *p);
    if (!p) return NULL;
    p->kind = 1; p->prec = prec; p->rank = k; p->nranks = P; p->sign = sign;
    p->n0 = r; p->n1 = m; p->R = 1; p->ln0 = lr; p->ln1 = lm; p->s0 = sr; p->s1 = sm; p->b0 = br; p->b1 = bm;
    p->in = in; p->out = out;
    ok = setup_peers(p, comm, out, (size_t)(r * bm > 0 ? r * bm : 1) * cs, (size_t)(br * m > 0 ? br * m : 1) * cs);
    if (ok && scr_in) {
        /* FFTW_MPI_SCRAMBLED_IN: the input is what a SCRAMBLED_OUT transform leaves -- element X[k1 + r k2] at
           [k1 in my block][k2].  With j = j1 m + j2:  w^(jk) = w_r^(j1 k1) w_n^(j2 k1) w_m^(j2 k2), so
             B'  FFT_m along my rows (over k2 -> j2) with the twiddle exp(-+2 pi i j2 k1 / n) in its store
             T1  columns of block d -> rank d's Z1 = [k1 (all r)][j2 in its block]
             A'  FFT_r down the columns of Z1 (over k1 -> j1)
             T2  rows j1 of block d -> rank d's OUTPUT [j1 in its block][j2 (all m)]: natural order
           two transposes instead of three. */
        p->kind = 3;
        if (lr > 0) {
            problem(&q, prec, pflags, in, p->z2alloc, sign);
            dim(&q.sz, m, 2, 2);
            dim(&q.vecsz, lr, 2 * m, 2 * m);
            q.tw_big_n = n0; q.tw_off = sr;
            p->z = b2_mkplan(&q);
            if (!p->z) ok = 0;
            for (d = 0; d < P && ok; ++d) {
                int64_t lmd = share(m, P, d), smd = bm * d < m ? bm * d : m;
                if (!lmd) continue;
                problem(&q, prec, pflags | B2F_ESTIMATE, p->z2alloc + cs * (size_t)smd, (char *)p->peer_z[d] + cs * (size_t)(sr * lmd), -1);
                dim(&q.vecsz, lr, 2 * m, 2 * lmd);
                dim(&q.vecsz, lmd, 2, 2);
                p->scatter[d] = b2_mkplan(&q);
                if (!p->scatter[d]) ok = 0;
            }
        }
        if (ok && lm > 0) {
            problem(&q, prec, pflags, p->zbuf, p->zbuf, sign);
            dim(&q.sz, r, 2 * lm, 2 * lm);
            dim(&q.vecsz, lm, 2, 2);
            p->local = b2_mkplan(&q);
            if (!p->local) ok = 0;
            for (d = 0; d < P && ok; ++d) {
                int64_t lrd = share(r, P, d), srd = br * d < r ? br * d : r;
                if (!lrd) continue;
                problem(&q, prec, pflags | B2F_ESTIMATE, p->zbuf + cs * (size_t)(srd * lm), (char *)p->peer_out[d] + cs * (size_t)sm, -1);
                dim(&q.vecsz, lrd, 2 * lm, 2 * m);
                dim(&q.vecsz, lm, 2, 2);
                p->back[d] = b2_mkplan(&q);
                if (!p->back[d]) ok = 0;
            }
        }
    } else
    if (ok) {
        for (d = 0; d < P && ok && lr > 0; ++d) {
            /* T1: my rows, columns of block d -> rank d's Z1 [r][lm(d)] at row sr */
            int64_t lmd = share(m, P, d), smd = bm * d < m ? bm * d : m;
            if (!lmd) continue;
            problem(&q, prec, pflags | B2F_ESTIMATE, (char *)in + cs * (size_t)smd, (char *)p->peer_z[d] + cs * (size_t)(sr * lmd), -1);
            dim(&q.vecsz, lr, 2 * m, 2 * lmd);
            dim(&q.vecsz, lmd, 2, 2);
            p->scatter[d] = b2_mkplan(&q);
            if (!p->scatter[d]) ok = 0;
        }
        if (ok && lm > 0) {
            /* A: FFT_r down the columns of Z1 [r][lm], in place, twiddle exp(-2 pi i k1 (sm + j2) / n) in the store */
            problem(&q, prec, pflags, p->zbuf, p->zbuf, sign);
            dim(&q.sz, r, 2 * lm, 2 * lm);
            dim(&q.vecsz, lm, 2, 2);
            q.tw_big_n = n0; q.tw_off = sm;
            p->local = b2_mkplan(&q);
            if (!p->local) ok = 0;
            for (d = 0; d < P && ok; ++d) {
                /* T2: rows k1 of block d -> rank d's Z2 [lr(d)][m] at column sm */
                int64_t lrd = share(r, P, d), srd = br * d < r ? br * d : r;
                if (!lrd) continue;
                problem(&q, prec, pflags | B2F_ESTIMATE, p->zbuf + cs * (size_t)(srd * lm), (char *)p->peer_z2[d] + cs * (size_t)sm, -1);
                dim(&q.vecsz, lrd, 2 * lm, 2 * m);
                dim(&q.vecsz, lm, 2, 2);
                p->back[d] = b2_mkplan(&q);
                if (!p->back[d]) ok = 0;
            }
        }
        if (ok && lr > 0) {
            /* B: FFT_m along the rows of Z2 [lr][m] (scrambled output: straight into `out`) */
            problem(&q, prec, pflags, p->z2alloc, scrambled ? out : (void *)p->z2alloc, sign);
            dim(&q.sz, m, 2, 2);
            dim(&q.vecsz, lr, 2 * m, 2 * m);
            p->z = b2_mkplan(&q);
            if (!p->z) ok = 0;
            for (d = 0; d < P && ok && !scrambled; ++d) {
                /* T3: columns k2 of block d, transposed -> rank d's out [lm(d)][r] at column sr */
                int64_t lmd = share(m, P, d), smd = bm * d < m ? bm * d : m;
                if (!lmd) continue;
                problem(&q, prec, pflags | B2F_ESTIMATE, p->z2alloc + cs * (size_t)smd, (char *)p->peer_out[d] + cs * (size_t)sr, -1);
                dim(&q.vecsz, lr, 2 * m, 2);
                dim(&q.vecsz, lmd, 2, 2 * r);
                p->t3[d] = b2_mkplan(&q);
                if (!p->t3[d]) ok = 0;
            }
        }
    }
    if (!agree(comm, ok)) { fftw_b200_mpi_destroy_plan(p); return NULL; }
    return p;
}

fftw_b200_mpi_plan fftw_b200_mpi_plan_dft_1d(ptrdiff_t n0, fftw_complex *in, fftw_complex *out, const fftw_b200_comm *comm,
                                             int sign, unsigned flags)
{ return mkplan1d(B2D_F64, n0, in, out, comm, sign, flags); }
fftw_b200_mpi_plan fftwf_b200_mpi_plan_dft_1d(ptrdiff_t n0, fftwf_complex *in, fftwf_complex *out, const fftw_b200_comm *comm,
                                              int sign, unsigned flags)
{ return mkplan1d(B2D_F32, n0, in, out, comm, sign, flags); }

/* ------------------------------------------------------------------ real data and r2r, 3-D
   fftw_mpi_plan_dft_r2c_3d / _c2r_3d / fftw_mpi_plan_r2r_3d (mpi/api.c:650-760, 770-886) behind the communicator
   interface: the plans of dist.c (double precision, default blocks), this file adds the buffer exchange through the
   callback and the device-side barriers between their stages.  Layout as fftw_mpi: real slab
   [local_n0][n1][2 (n2/2+1)] (padded rows, may alias the complex slab [local_n0][n1][n2/2+1]); local sizes come from
   fftw_b200_mpi_local_size_3d(n0, n1, n2/2+1) in complex elements. */
static fftw_b200_mpi_plan mkreal3d(int what, ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, void *in, void *out,
                                   const fftw_b200_comm *comm, const int *kinds, unsigned flags)
{
    fftw_b200_mpi_plan p;
    int d, P, r, ok;
    int64_t h = n2 / 2 + 1, b0, b1;
    size_t zbytes;
    void *slab, *push[MAXP];
    unsigned pflags = flags & ~(FFTW_MPI_TRANSPOSED_OUT | FFTW_MPI_TRANSPOSED_IN | FFTW_MPI_SCRAMBLED_IN | FFTW_MPI_SCRAMBLED_OUT);
    if (!comm || !comm->allgather || n0 <= 0 || n1 <= 0 || n2 <= 0 || !in || !out || pflags != flags) return NULL;
    P = comm->nranks; r = comm->rank;
    if (P < 1 || P > MAXP || r < 0 || r >= P) return NULL;
    if (b2d_pointer_is_device(in) != 1 || b2d_pointer_is_device(out) != 1) return NULL;
    p = (fftw_b200_mpi_plan)calloc(1, sizeof *p);
    if (!p) return NULL;
    b0 = blk(n0, P); b1 = blk(n1, P);
    p->kind = 4 + what; p->prec = B2D_F64; p->rank = r; p->nranks = P; p->rnk = 3;
    p->n0 = n0; p->n1 = n1; p->R = what == 2 ? n2 : h; p->b0 = b0; p->b1 = b1;
    p->ln0 = share(n0, P, r); p->ln1 = share(n1, P, r);
    p->s0 = b0 * r < n0 ? b0 * r : n0; p->s1 = b1 * r < n1 ? b1 * r : n1;
    p->in = in; p->out = out;
    /* the array the peers write into: the complex slab (r2c: the output, c2r: the input), r2r: the slab itself */
    slab = what == 1 ? in : out;
    zbytes = what == 2 ? sizeof(double) * (size_t)(n0 * b1 * n2) : 16 * (size_t)(P * b0 * b1 * h);
    ok = setup_peers(p, comm, slab, zbytes ? zbytes : 16, 0);
    if (ok) {
        if (what == 2) {
            if (in != out) {
                /* out of place: copy the slab, then transform the copy in place */
                b2_problem q;
                rproblem(&q, B2D_F64, pflags, in, out);
                if (p->ln0 > 0) { dim(&q.vecsz, p->ln0 * n1 * n2, 1, 1); p->local = b2_mkplan(&q); if (!p->local) ok = 0; }
            }
            if (ok) p->fused = fftw_b200_dist_plan_r2r_3d(n0, n1, n2, r, P, (double *)out, (double *)p->zbuf, p->peer_out,
                                                          p->peer_z, kinds, pflags);
        } else {
            for (d = 0; d < P; ++d) push[d] = (char *)p->peer_z[d] + 16 * (size_t)(r * b0 * b1 * h);
            p->fused = what == 0
                ? fftw_b200_dist_plan_dft_r2c_3d(n0, n1, n2, r, P, (double *)in, (fftw_complex *)out, (fftw_complex *)p->zbuf,
                                                 push, p->peer_out, pflags)
                : fftw_b200_dist_plan_dft_c2r_3d(n0, n1, n2, r, P, (fftw_complex *)in, (double *)out, (fftw_complex *)p->zbuf,
                                                 push, p->peer_out, pflags);
        }
        if (!p->fused) ok = 0;
    }
    if (!agree(comm, ok)) { fftw_b200_mpi_destroy_plan(p); return NULL; }
    return p;
}

/* 2-D real data (fftw_mpi_plan_dft_r2c_2d / _c2r_2d): real slab [local_n0][2 (n1/2+1)] (padded rows), complex slab
   [local_n0][h], h = n1/2 + 1.  r2c = local r2c of the rows, then the distributed c2c along n0 over the n0 x h
   complex matrix (scatter column blocks | FFT_n0 | push rows back: the general path above with h playing n1 and no
   local pass); c2r = the same backward, then the local c2r of the rows.  Natural layouts only (the TRANSPOSED
   flags return NULL).  The complex slab is overwritten by c2r, as in fftw_mpi. */
static fftw_b200_mpi_plan mkreal2d(int prec, int c2r, ptrdiff_t n0, ptrdiff_t n1, void *in, void *out, const fftw_b200_comm *comm,
                                   unsigned flags)
{
    fftw_b200_mpi_plan p;
    b2_problem q;
    int d, P, r, ok;
    int64_t h = n1 / 2 + 1, alloc;
    const int sign = c2r ? 1 : -1;
    const size_t cs = csize(prec);
    char *real = (char *)(c2r ? out : in);
    char *cplx = (char *)(c2r ? in : out);
    unsigned pflags = flags & ~(FFTW_MPI_TRANSPOSED_OUT | FFTW_MPI_TRANSPOSED_IN | FFTW_MPI_SCRAMBLED_IN | FFTW_MPI_SCRAMBLED_OUT);
    if (!comm || !comm->allgather || n0 <= 0 || n1 <= 0 || !in || !out) return NULL;
    if (pflags != flags) return NULL;
    P = comm->nranks; r = comm->rank;
    if (P < 1 || P > MAXP || r < 0 || r >= P) return NULL;
    if (b2d_pointer_is_device(in) != 1 || b2d_pointer_is_device(out) != 1) return NULL;
    p = (fftw_b200_mpi_plan)calloc(1, sizeof *p);
    if (!p) return NULL;
    p->kind = c2r ? 8 : 7; p->prec = prec; p->rank = r; p->nranks = P; p->rnk = 2; p->sign = sign;
    p->n0 = n0; p->n1 = h; p->R = 1;
    p->b0 = blk(n0, P); p->b1 = blk(h, P);
    p->ln0 = share(n0, P, r); p->ln1 = share(h, P, r);
    p->s0 = p->b0 * r < n0 ? p->b0 * r : n0;
    p->s1 = p->b1 * r < h ? p->b1 * r : h;
    p->in = in; p->out = out;
    alloc = p->b0 * h;
    if (p->b1 * n0 > alloc) alloc = p->b1 * n0;
    ok = setup_peers(p, comm, cplx, (size_t)(alloc > 0 ? alloc : 1) * cs, 0);
    if (ok && p->ln0 > 0) {
        /* rows: padded real rows <-> this rank's complex rows [ln0][h] */
        memset(&q, 0, sizeof q);
        q.prec = prec; q.kind = c2r ? B2_C2R : B2_R2C; q.flags = pflags;
        b2_tensor_init(&q.sz, 0); b2_tensor_init(&q.vecsz, 0);
        dim(&q.sz, n1, c2r ? 2 : 1, c2r ? 1 : 2);
        dim(&q.vecsz, p->ln0, 2 * h, 2 * h);
        if (c2r) { q.in0 = cplx; q.in1 = cplx + cs / 2; q.out0 = real; }
        else { q.in0 = real; q.out0 = cplx; q.out1 = cplx + cs / 2; }
        p->local = b2_mkplan(&q);
        if (!p->local) ok = 0;
    }
    if (ok) {
        /* scatter: my rows of column block d -> rank d's exchange buffer [n0][lh(d)] at row s0 */
        for (d = 0; d < P && ok && p->ln0 > 0; ++d) {
            int64_t l1 = share(h, P, d);
            if (!l1) continue;
            problem(&q, prec, pflags | B2F_ESTIMATE, cplx + cs * (size_t)(p->b1 * d), (char *)p->peer_z[d] + cs * (size_t)(p->s0 * l1), -1);
            dim(&q.vecsz, p->ln0, 2 * h, 2 * l1);
            dim(&q.vecsz, l1, 2, 2);
            p->scatter[d] = b2_mkplan(&q);
            if (!p->scatter[d]) ok = 0;
        }
        if (ok && p->ln1 > 0) {
            problem(&q, prec, pflags, p->zbuf, p->zbuf, sign);
            dim(&q.sz, n0, 2 * p->ln1, 2 * p->ln1);
            dim(&q.vecsz, p->ln1, 2, 2);
            p->z = b2_mkplan(&q);
            if (!p->z) ok = 0;
            for (d = 0; d < P && ok; ++d) {
                /* rows of owner d -> its complex slab [ln0(d)][h] at column s1 */
                int64_t l0 = share(n0, P, d);
                if (!l0) continue;
                problem(&q, prec, pflags | B2F_ESTIMATE, p->zbuf + cs * (size_t)(p->b0 * d * p->ln1),
                        (char *)p->peer_out[d] + cs * (size_t)p->s1, -1);
                dim(&q.vecsz, l0, 2 * p->ln1, 2 * h);
                dim(&q.vecsz, p->ln1, 2, 2);
                p->back[d] = b2_mkplan(&q);
                if (!p->back[d]) ok = 0;
            }
        }
    }
    if (!agree(comm, ok)) { fftw_b200_mpi_destroy_plan(p); return NULL; }
    return p;
}

fftw_b200_mpi_plan fftw_b200_mpi_plan_dft_r2c_2d(ptrdiff_t n0, ptrdiff_t n1, double *in, fftw_complex *out,
                                                 const fftw_b200_comm *comm, unsigned flags)
{ return mkreal2d(B2D_F64, 0, n0, n1, in, out, comm, flags); }
fftw_b200_mpi_plan fftw_b200_mpi_plan_dft_c2r_2d(ptrdiff_t n0, ptrdiff_t n1, fftw_complex *in, double *out,
                                                 const fftw_b200_comm *comm, unsigned flags)
{ return mkreal2d(B2D_F64, 1, n0, n1, in, out, comm, flags); }

fftw_b200_mpi_plan fftw_b200_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, double *in, fftw_complex *out,
                                                 const fftw_b200_comm *comm, unsigned flags)
{ return mkreal3d(0, n0, n1, n2, in, out, comm, NULL, flags); }
fftw_b200_mpi_plan fftw_b200_mpi_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex *in, double *out,
                                                 const fftw_b200_comm *comm, unsigned flags)
{ return mkreal3d(1, n0, n1, n2, in, out, comm, NULL, flags); }
fftw_b200_mpi_plan fftw_b200_mpi_plan_r2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, double *in, double *out,
                                             const fftw_b200_comm *comm, fftw_r2r_kind kind0, fftw_r2r_kind kind1,
                                             fftw_r2r_kind kind2, unsigned flags)
{
    int kinds[3];
    kinds[0] = (int)kind0; kinds[1] = (int)kind1; kinds[2] = (int)kind2;
    return mkreal3d(2, n0, n1, n2, in, out, comm, kinds, flags);
}

/* ------------------------------------------------------------------ general r2r and real-data plans
   fftw_mpi_plan_many_r2r (mpi/api.c:770-886) for any rank >= 2 and any howmany, both precisions: local r2r over
   dimensions 1 .. rnk-1, column blocks pushed into the owners' exchange buffers, r2r along n0 there, rows pushed back
   -- the sequence of the complex general path above (kind 0), on reals.  kinds[i] along dimension i. */
static fftw_b200_mpi_plan mkr2r(int prec, int rnk, const ptrdiff_t *n, ptrdiff_t howmany, void *in, void *out,
                                const fftw_b200_comm *comm, const int *kinds, unsigned flags)
{
    fftw_b200_mpi_plan p;
    b2_problem q;
    int i, d, P, r, ok = 1;
    int64_t R = howmany, alloc, st;
    size_t rs = prec == B2D_F32 ? 4 : 8;
    if (!comm || !comm->allgather || rnk < 2 || rnk > 8 || howmany < 1 || !in || !out || !kinds) return NULL;
    if (flags & (FFTW_MPI_TRANSPOSED_OUT | FFTW_MPI_TRANSPOSED_IN | FFTW_MPI_SCRAMBLED_IN | FFTW_MPI_SCRAMBLED_OUT)) return NULL;
    P = comm->nranks; r = comm->rank;
    if (P < 1 || P > MAXP || r < 0 || r >= P) return NULL;
    for (i = 0; i < rnk; ++i) if (n[i] <= 0 || kinds[i] < 0 || kinds[i] > 10) return NULL;
    if (b2d_pointer_is_device(in) != 1 || b2d_pointer_is_device(out) != 1) return NULL;
    p = (fftw_b200_mpi_plan)calloc(1, sizeof *p);
    if (!p) return NULL;
    for (i = 2; i < rnk; ++i) R *= n[i];
    p->prec = prec; p->rank = r; p->nranks = P; p->rnk = rnk;
    p->n0 = n[0]; p->n1 = n[1]; p->R = R;
    p->b0 = blk(n[0], P); p->b1 = blk(n[1], P);
    p->ln0 = share(n[0], P, r); p->ln1 = share(n[1], P, r);
    p->s0 = p->b0 * r < n[0] ? p->b0 * r : n[0];
    p->s1 = p->b1 * r < n[1] ? p->b1 * r : n[1];
    p->in = in; p->out = out;
    alloc = p->b0 * n[1] * R;
    if (p->b1 * n[0] * R > alloc) alloc = p->b1 * n[0] * R;
    ok = setup_peers(p, comm, out, (size_t)(alloc > 0 ? alloc : 1) * rs, 0);
    if (ok && p->ln0 > 0) {
        /* local r2r over dims 1 .. rnk-1 of [ln0][n1]...[howmany], in -> out */
        rproblem(&q, prec, flags, in, out);
        q.flags = flags;
        st = howmany;
        for (i = rnk - 1; i >= 1; --i) { q.sz.d[i - 1].n = n[i]; q.sz.d[i - 1].is = q.sz.d[i - 1].os = st; st *= n[i]; q.r2r_kind[i - 1] = kinds[i]; }
        q.sz.rnk = rnk - 1;
        dim(&q.vecsz, p->ln0, n[1] * R, n[1] * R);
        if (howmany > 1) dim(&q.vecsz, howmany, 1, 1);
        p->local = b2_mkplan(&q);
        if (!p->local) ok = 0;
        for (d = 0; d < P && ok; ++d) {
            int64_t l1 = share(n[1], P, d);
            if (!l1) continue;
            rproblem(&q, prec, flags, (char *)out + rs * (size_t)(p->b1 * d * R), (char *)p->peer_z[d] + rs * (size_t)(p->s0 * l1 * R));
            dim(&q.vecsz, p->ln0, n[1] * R, l1 * R);
            dim(&q.vecsz, l1 * R, 1, 1);
            p->scatter[d] = b2_mkplan(&q);
            if (!p->scatter[d]) ok = 0;
        }
    }
    if (ok && p->ln1 > 0) {
        rproblem(&q, prec, flags, p->zbuf, p->zbuf);
        q.flags = flags;
        dim(&q.sz, n[0], p->ln1 * R, p->ln1 * R);
        q.r2r_kind[0] = kinds[0];
        dim(&q.vecsz, p->ln1 * R, 1, 1);
        p->z = b2_mkplan(&q);
        if (!p->z) ok = 0;
        for (d = 0; d < P && ok; ++d) {
            int64_t l0 = share(n[0], P, d);
            if (!l0) continue;
            rproblem(&q, prec, flags, p->zbuf + rs * (size_t)(p->b0 * d * p->ln1 * R), (char *)p->peer_out[d] + rs * (size_t)(p->s1 * R));
            dim(&q.vecsz, l0, p->ln1 * R, n[1] * R);
            dim(&q.vecsz, p->ln1 * R, 1, 1);
            p->back[d] = b2_mkplan(&q);
            if (!p->back[d]) ok = 0;
        }
    }
    if (!agree(comm, ok)) { fftw_b200_mpi_destroy_plan(p); return NULL; }
    return p;
}

fftw_b200_mpi_plan fftw_b200_mpi_plan_many_r2r(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                               double *in, double *out, const fftw_b200_comm *comm,
                                               const fftw_r2r_kind *kind, unsigned flags)
{
    int k[8], i;
    if (iblock || oblock || rnk < 2 || rnk > 8 || !kind) return NULL;
    for (i = 0; i < rnk; ++i) k[i] = (int)kind[i];
    return mkr2r(B2D_F64, rnk, n, howmany, in, out, comm, k, flags);
}
fftw_b200_mpi_plan fftwf_b200_mpi_plan_many_r2r(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                                float *in, float *out, const fftw_b200_comm *comm,
                                                const fftwf_r2r_kind *kind, unsigned flags)
{
    int k[8], i;
    if (iblock || oblock || rnk < 2 || rnk > 8 || !kind) return NULL;
    for (i = 0; i < rnk; ++i) k[i] = (int)kind[i];
    return mkr2r(B2D_F32, rnk, n, howmany, in, out, comm, k, flags);
}
fftw_b200_mpi_plan fftw_b200_mpi_plan_r2r_2d(ptrdiff_t n0, ptrdiff_t n1, double *in, double *out, const fftw_b200_comm *comm,
                                             fftw_r2r_kind kind0, fftw_r2r_kind kind1, unsigned flags)
{
    ptrdiff_t n[2]; int k[2];
    n[0] = n0; n[1] = n1; k[0] = (int)kind0; k[1] = (int)kind1;
    return mkr2r(B2D_F64, 2, n, 1, in, out, comm, k, flags);
}

/* fftw_mpi_plan_many_dft_r2c / _c2r (mpi/api.c:650-760) for rank >= 3, any howmany, both precisions: r2c = local r2c
   over dimensions 1 .. rnk-1 (last one halved: h = n_last/2 + 1), then the distributed c2c along n0 exactly as in the
   complex general path, with R = n2 ... n_(rnk-2) h howmany complex numbers per (i0, i1); c2r = the same backward, then
   the local c2r.  Real slab [local_n0][n1]...[2 h][howmany] (padded), complex slab [local_n0][n1]...[h][howmany]; the
   complex slab is overwritten by c2r.  Rank 2 (the halved dimension itself is exchanged): mkreal2d above. */
static fftw_b200_mpi_plan mkrealnd(int prec, int c2r, int rnk, const ptrdiff_t *n, ptrdiff_t howmany, void *in, void *out,
                                   const fftw_b200_comm *comm, unsigned flags)
{
    fftw_b200_mpi_plan p;
    b2_problem q;
    int i, d, P, r, ok = 1;
    const int sign = c2r ? 1 : -1;
    int64_t h, R = howmany, alloc, inner, str, stc;
    size_t cs = csize(prec), rs = cs / 2;
    char *real = (char *)(c2r ? out : in), *cplx = (char *)(c2r ? in : out);
    if (!comm || !comm->allgather || rnk < 3 || rnk > 8 || howmany < 1 || !in || !out) return NULL;
    if (flags & (FFTW_MPI_TRANSPOSED_OUT | FFTW_MPI_TRANSPOSED_IN | FFTW_MPI_SCRAMBLED_IN | FFTW_MPI_SCRAMBLED_OUT)) return NULL;
    P = comm->nranks; r = comm->rank;
    if (P < 1 || P > MAXP || r < 0 || r >= P) return NULL;
    for (i = 0; i < rnk; ++i) if (n[i] <= 0) return NULL;
    if (b2d_pointer_is_device(in) != 1 || b2d_pointer_is_device(out) != 1) return NULL;
    p = (fftw_b200_mpi_plan)calloc(1, sizeof *p);
    if (!p) return NULL;
    h = n[rnk - 1] / 2 + 1;
    for (i = 2; i < rnk - 1; ++i) R *= n[i];
    R *= h;
    inner = 2 * R;                       /* reals per (i0, i1) row, both in the complex and in the padded real slab */
    p->kind = c2r ? 8 : 0; p->prec = prec; p->rank = r; p->nranks = P; p->rnk = rnk; p->sign = sign;
    p->n0 = n[0]; p->n1 = n[1]; p->R = R;
    p->b0 = blk(n[0], P); p->b1 = blk(n[1], P);
    p->ln0 = share(n[0], P, r); p->ln1 = share(n[1], P, r);
    p->s0 = p->b0 * r < n[0] ? p->b0 * r : n[0];
    p->s1 = p->b1 * r < n[1] ? p->b1 * r : n[1];
    p->in = in; p->out = out;
    alloc = p->b0 * n[1] * R;
    if (p->b1 * n[0] * R > alloc) alloc = p->b1 * n[0] * R;
    ok = setup_peers(p, comm, cplx, (size_t)(alloc > 0 ? alloc : 1) * cs, 0);
    if (ok && p->ln0 > 0) {
        /* local r2c / c2r over dims 1 .. rnk-1; strides in reals: real side (padded last dim 2h), complex side */
        memset(&q, 0, sizeof q);
        q.prec = prec; q.kind = c2r ? B2_C2R : B2_R2C; q.flags = flags;
        b2_tensor_init(&q.sz, 0); b2_tensor_init(&q.vecsz, 0);
        str = howmany; stc = 2 * howmany;
        for (i = rnk - 1; i >= 1; --i) {
            q.sz.d[i - 1].n = n[i];
            q.sz.d[i - 1].is = c2r ? stc : str;
            q.sz.d[i - 1].os = c2r ? str : stc;
            if (i == rnk - 1) { str *= 2 * h; stc *= h; } else { str *= n[i]; stc *= n[i]; }
        }
        q.sz.rnk = rnk - 1;
        dim(&q.vecsz, p->ln0, n[1] * inner, n[1] * inner);
        if (howmany > 1) dim(&q.vecsz, howmany, c2r ? 2 : 1, c2r ? 1 : 2);
        if (c2r) { q.in0 = cplx; q.in1 = cplx + rs; q.out0 = real; }
        else { q.in0 = real; q.out0 = cplx; q.out1 = cplx + rs; }
        p->local = b2_mkplan(&q);
        if (!p->local) ok = 0;
        for (d = 0; d < P && ok; ++d) {
            int64_t l1 = share(n[1], P, d);
            if (!l1) continue;
            problem(&q, prec, flags | B2F_ESTIMATE, cplx + rs * (size_t)(p->b1 * d * inner),
                    (char *)p->peer_z[d] + rs * (size_t)(p->s0 * l1 * inner), -1);
            dim(&q.vecsz, p->ln0, n[1] * inner, l1 * inner);
            dim(&q.vecsz, l1 * R, 2, 2);
            p->scatter[d] = b2_mkplan(&q);
            if (!p->scatter[d]) ok = 0;
        }
    }
    if (ok && p->ln1 > 0) {
        problem(&q, prec, flags, p->zbuf, p->zbuf, sign);
        dim(&q.sz, n[0], p->ln1 * inner, p->ln1 * inner);
        dim(&q.vecsz, p->ln1 * R, 2, 2);
        p->z = b2_mkplan(&q);
        if (!p->z) ok = 0;
        for (d = 0; d < P && ok; ++d) {
            int64_t l0 = share(n[0], P, d);
            if (!l0) continue;
            problem(&q, prec, flags | B2F_ESTIMATE, p->zbuf + rs * (size_t)(p->b0 * d * p->ln1 * inner),
                    (char *)p->peer_out[d] + rs * (size_t)(p->s1 * inner), -1);
            dim(&q.vecsz, l0, p->ln1 * inner, n[1] * inner);
            dim(&q.vecsz, p->ln1 * R, 2, 2);
            p->back[d] = b2_mkplan(&q);
            if (!p->back[d]) ok = 0;
        }
    }
    if (!agree(comm, ok)) { fftw_b200_mpi_destroy_plan(p); return NULL; }
    return p;
}

fftw_b200_mpi_plan fftw_b200_mpi_plan_many_dft_r2c(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                                   double *in, fftw_complex *out, const fftw_b200_comm *comm, unsigned flags)
{
    if (iblock || oblock) return NULL;
    if (rnk == 2 && howmany == 1) return mkreal2d(B2D_F64, 0, n[0], n[1], in, out, comm, flags);
    return mkrealnd(B2D_F64, 0, rnk, n, howmany, in, out, comm, flags);
}
fftw_b200_mpi_plan fftw_b200_mpi_plan_many_dft_c2r(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                                   fftw_complex *in, double *out, const fftw_b200_comm *comm, unsigned flags)
{
    if (iblock || oblock) return NULL;
    if (rnk == 2 && howmany == 1) return mkreal2d(B2D_F64, 1, n[0], n[1], in, out, comm, flags);
    return mkrealnd(B2D_F64, 1, rnk, n, howmany, in, out, comm, flags);
}
fftw_b200_mpi_plan fftwf_b200_mpi_plan_many_dft_r2c(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                                    float *in, fftwf_complex *out, const fftw_b200_comm *comm, unsigned flags)
{
    if (iblock || oblock) return NULL;
    if (rnk == 2 && howmany == 1) return mkreal2d(B2D_F32, 0, n[0], n[1], in, out, comm, flags);
    return mkrealnd(B2D_F32, 0, rnk, n, howmany, in, out, comm, flags);
}
fftw_b200_mpi_plan fftwf_b200_mpi_plan_many_dft_c2r(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                                    fftwf_complex *in, float *out, const fftw_b200_comm *comm, unsigned flags)
{
    if (iblock || oblock) return NULL;
    if (rnk == 2 && howmany == 1) return mkreal2d(B2D_F32, 1, n[0], n[1], in, out, comm, flags);
    return mkrealnd(B2D_F32, 1, rnk, n, howmany, in, out, comm, flags);
}

/* single-precision basic forms: through the general plans (the fused 3-D plans of dist.c are double precision) */
fftw_b200_mpi_plan fftwf_b200_mpi_plan_dft_r2c_2d(ptrdiff_t n0, ptrdiff_t n1, float *in, fftwf_complex *out,
                                                  const fftw_b200_comm *comm, unsigned flags)
{ return mkreal2d(B2D_F32, 0, n0, n1, in, out, comm, flags); }
fftw_b200_mpi_plan fftwf_b200_mpi_plan_dft_c2r_2d(ptrdiff_t n0, ptrdiff_t n1, fftwf_complex *in, float *out,
                                                  const fftw_b200_comm *comm, unsigned flags)
{ return mkreal2d(B2D_F32, 1, n0, n1, in, out, comm, flags); }
fftw_b200_mpi_plan fftwf_b200_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, float *in, fftwf_complex *out,
                                                  const fftw_b200_comm *comm, unsigned flags)
{ ptrdiff_t n[3]; n[0] = n0; n[1] = n1; n[2] = n2; return mkrealnd(B2D_F32, 0, 3, n, 1, in, out, comm, flags); }
fftw_b200_mpi_plan fftwf_b200_mpi_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftwf_complex *in, float *out,
                                                  const fftw_b200_comm *comm, unsigned flags)
{ ptrdiff_t n[3]; n[0] = n0; n[1] = n1; n[2] = n2; return mkrealnd(B2D_F32, 1, 3, n, 1, in, out, comm, flags); }
fftw_b200_mpi_plan fftwf_b200_mpi_plan_r2r_2d(ptrdiff_t n0, ptrdiff_t n1, float *in, float *out, const fftw_b200_comm *comm,
                                              fftwf_r2r_kind kind0, fftwf_r2r_kind kind1, unsigned flags)
{
    ptrdiff_t n[2]; int k[2];
    n[0] = n0; n[1] = n1; k[0] = (int)kind0; k[1] = (int)kind1;
    return mkr2r(B2D_F32, 2, n, 1, in, out, comm, k, flags);
}
fftw_b200_mpi_plan fftwf_b200_mpi_plan_r2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, float *in, float *out,
                                              const fftw_b200_comm *comm, fftwf_r2r_kind kind0, fftwf_r2r_kind kind1,
                                              fftwf_r2r_kind kind2, unsigned flags)
{
    ptrdiff_t n[3]; int k[3];
    n[0] = n0; n[1] = n1; n[2] = n2; k[0] = (int)kind0; k[1] = (int)kind1; k[2] = (int)kind2;
    return mkr2r(B2D_F32, 3, n, 1, in, out, comm, k, flags);
}

/* ------------------------------------------------------------------ wisdom across ranks (mpi/wisdom-api.c:24-103)
   fftw_mpi_gather_wisdom: rank 0 ends up with the union of every rank's wisdom; fftw_mpi_broadcast_wisdom: every
   rank imports rank 0's.  One length all-gather + one padded-text all-gather through the communicator callback. */
static void exchange_wisdom(const fftw_b200_comm *comm, int gather, char *(*export_str)(void), int (*import_str)(const char *))
{
    char *mine, *all = NULL, *padded = NULL;
    unsigned long long len, *lens = NULL, maxlen = 0;
    int d, P;
    if (!comm || !comm->allgather || comm->nranks < 2) return;
    P = comm->nranks;
    mine = export_str();
    len = mine ? (unsigned long long)strlen(mine) + 1 : 1;
    lens = (unsigned long long *)calloc((size_t)P, sizeof *lens);
    if (lens && !comm->allgather(comm->ctx, &len, lens, sizeof len)) {
        for (d = 0; d < P; ++d) if (lens[d] > maxlen) maxlen = lens[d];
        padded = (char *)calloc((size_t)maxlen, 1);
        all = (char *)malloc((size_t)maxlen * (size_t)P);
        if (padded && all) {
            if (mine) memcpy(padded, mine, (size_t)len);
            if (!comm->allgather(comm->ctx, padded, all, (size_t)maxlen)) {
                if (gather) { if (comm->rank == 0) for (d = 1; d < P; ++d) import_str(all + (size_t)d * maxlen); }
                else if (comm->rank != 0) import_str(all);
            }
        }
    }
    free(lens); free(padded); free(all); free(mine);
}

void fftw_b200_mpi_gather_wisdom(const fftw_b200_comm *comm)
{ exchange_wisdom(comm, 1, fftw_export_wisdom_to_string, fftw_import_wisdom_from_string); }
void fftw_b200_mpi_broadcast_wisdom(const fftw_b200_comm *comm)
{ exchange_wisdom(comm, 0, fftw_export_wisdom_to_string, fftw_import_wisdom_from_string); }
void fftwf_b200_mpi_gather_wisdom(const fftw_b200_comm *comm)
{ exchange_wisdom(comm, 1, fftwf_export_wisdom_to_string, fftwf_import_wisdom_from_string); }
void fftwf_b200_mpi_broadcast_wisdom(const fftw_b200_comm *comm)
{ exchange_wisdom(comm, 0, fftwf_export_wisdom_to_string, fftwf_import_wisdom_from_string); }

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
    if (p->kind == 2) {
        for (d = 0; d < p->nranks; ++d) run(p->scatter[(p->rank + 1 + d) % p->nranks]);
        barrier(p);
        run(p->back[0]);
        barrier(p);                                     /* nobody overwrites an exchange buffer still being copied back */
        if (!b2_async_mode) b2d_sync();
        return;
    }
    if (p->kind == 7 || p->kind == 8) {
        /* 2-D real data: rows locally, columns through the exchange (c2r: the other way round) */
        if (p->kind == 7) run(p->local);
        for (d = 0; d < p->nranks; ++d) run(p->scatter[(p->rank + 1 + d) % p->nranks]);
        barrier(p);
        run(p->z);
        for (d = 0; d < p->nranks; ++d) run(p->back[(p->rank + 1 + d) % p->nranks]);
        barrier(p);
        if (p->kind == 8) run(p->local);
        if (!b2_async_mode) b2d_sync();
        return;
    }
    if (p->kind >= 4) {
        if (p->kind == 6) {
            /* r2r: a barrier before every stage (peers read each other's slabs and exchange buffers) */
            int st;
            run(p->local);
            for (st = 0; st < 3; ++st) { barrier(p); fftw_b200_dist_execute_stage(p->fused, st); }
            barrier(p);                                 /* the peers have finished reading my exchange buffer: the
                                                           caller may destroy the plan (found with AddressSanitizer) */
        } else {
            fftw_b200_dist_execute_stage(p->fused, 0);
            barrier(p);                                 /* every row block has landed in my exchange buffer */
            fftw_b200_dist_execute_stage(p->fused, 1);
            barrier(p);                                 /* every rank's rows have landed in my complex slab */
            if (p->kind == 5) fftw_b200_dist_execute_stage(p->fused, 2);
        }
        if (!b2_async_mode) b2d_sync();
        return;
    }
    if (p->kind == 3) {
        run(p->z);                                                                             /* B' (+ twiddle) */
        for (d = 0; d < p->nranks; ++d) run(p->scatter[(p->rank + 1 + d) % p->nranks]);      /* T1 */
        barrier(p);
        run(p->local);                                                                         /* A' */
        for (d = 0; d < p->nranks; ++d) run(p->back[(p->rank + 1 + d) % p->nranks]);          /* T2 -> outputs */
        barrier(p);
        if (!b2_async_mode) b2d_sync();
        return;
    }
    if (p->kind == 1) {
        for (d = 0; d < p->nranks; ++d) run(p->scatter[(p->rank + 1 + d) % p->nranks]);      /* T1 */
        barrier(p);
        run(p->local);                                                                         /* A */
        for (d = 0; d < p->nranks; ++d) run(p->back[(p->rank + 1 + d) % p->nranks]);          /* T2 */
        barrier(p);
        run(p->z);                                                                             /* B */
        for (d = 0; d < p->nranks; ++d) run(p->t3[(p->rank + 1 + d) % p->nranks]);            /* T3 */
        barrier(p);
        if (!b2_async_mode) b2d_sync();
        return;
    }
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
