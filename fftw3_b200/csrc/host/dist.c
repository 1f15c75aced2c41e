/* dist.c -- slab-decomposed 3-D c2c over several GPUs (one process per GPU),
 * built as a composition of ordinary single-GPU plans.
 *
 * Algorithm = the reference's mpi/dft-rank-geq2-transposed.c:47-70:
 *   local transforms over the non-distributed dims, global transpose n0 <-> n1,
 *   transforms along n0 (now local); natural-order output adds the transpose
 *   back (mpi/dft-rank-geq2.c:40-59 via dft-rank1-bigvec.c:45-65).
 * What is different: the reference's transpose = local transpose + MPI_Alltoall
 * + local transpose (mpi/transpose-alltoall.c:49-100).  Here both local
 * transposes are folded into FFT passes and the exchange into their stores:
 *   stage 0  Y: FFT along n1, in place (strided pass)
 *            X: FFT along n2 (contiguous rows).  The slab is cut into chunks of planes and the
 *               block layout of the exchange -- [dest][i0][k1'][k2] -- is reached either
 *               (default, device-resident slabs) by the copy engines: X of chunk c runs in place
 *               at full speed (the rows this rank keeps go straight into its own exchange
 *               buffer), one strided copy per peer then moves the chunk's blocks over NVLink on
 *               side streams while the SMs are already on Y/X of chunk c+1, so no SM ever waits
 *               for a link; or (FFTW3_B200_DIST_EXCHANGE=stores) by X's own stores into the
 *               peers' buffers, X of chunk c on a side stream next to Y of chunk c+1.
 *   stage 1  Z: FFT along n0 reading the received blocks, which already form
 *               [n0][local_n1][n2]; written in place (natural order follows) or
 *               as [local_n1][n0][n2] into `local` (TRANSPOSED_OUT)
 *            or (push plans) with its output ROWS stored straight into the owners' slabs:
 *               row k0 belongs to rank k0 / block, so the second exchange is fused into
 *               the stores of this pass as well and there is no stage 2
 *   stage 2     gather the blocks back into [local_n0][n1][n2] (peer loads).
 *               Stages 1 and 2 are cut into chunks of columns so that the
 *               gather of chunk c overlaps Z of chunk c+1; the caller puts a
 *               barrier between Z_c and gather_c.
 */
#include <stdlib.h>
#include <string.h>
#include "b2_internal.h"

typedef double C[2];

struct fftw_b200_dist_plan_s {
    int nranks, rank, nstages;
    int c0, c1;              /* chunks of stage 0 (planes) and of stages 1/2 (columns) */
    int x_fused[64];         /* per chunk: one X launch scatters to every destination */
    b2_plan **y;             /* [c0] */
    b2_plan **x;             /* [c0] fused or [c0 * nranks] */
    b2_plan **z;             /* [c1] */
    b2_plan **g;             /* [c1 * nranks] */
    int real_gather;         /* r2r plan: x[] = stage-1 gathers (before z), g[] = stage-2 gathers */
    int zcopy;               /* real-data plans whose dim-0 pass cannot split its stores by row (Bluestein / Rader /
                                multi-pass n0): z[] transforms in place, g[s] then copies the rows to their owner s */
    b2_plan *pre, *post;     /* real-data plans: local r2c rows before stage 0 / local c2r rows as the last stage */
    /* stage-0 exchange by the copy engines (see mkdist): X runs in place at full speed, the blocks then
       travel as 2-D copies on side streams while the SMs are already on the next chunk */
    int part_sms;            /* > 0: stage 0 runs X (NVLink-bound) on that many SMs of their own and Y on the rest */
    int ce;
    int64_t ce_n1, ce_n2, ce_ln0;
    double *ce_local;
    void *ce_targets[B2D_MAX_PEERS];
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

static void init_problem(b2_problem *q, unsigned flags)
{
    memset(q, 0, sizeof *q);
    q->prec = B2D_F64; q->kind = B2_C2C; q->flags = flags;
    b2_tensor_init(&q->sz, 0); b2_tensor_init(&q->vecsz, 0);
}

void fftw_b200_dist_destroy_plan(dplan p)
{
    int i;
    if (!p) return;
    for (i = 0; p->y && i < p->c0; ++i) b2_plan_destroy(p->y[i]);
    for (i = 0; p->x && i < p->c0 * p->nranks; ++i) b2_plan_destroy(p->x[i]);
    for (i = 0; p->z && i < p->c1; ++i) b2_plan_destroy(p->z[i]);
    for (i = 0; p->g && i < p->c1 * p->nranks; ++i) b2_plan_destroy(p->g[i]);
    b2_plan_destroy(p->pre);
    b2_plan_destroy(p->post);
    free(p->y); free(p->x); free(p->z); free(p->g);
    free(p);
}

/* CTAs granted to an NVLink-bound pass (scatter, gather) that runs next to an HBM-bound one */
static int comm_ctas(void)
{
    const char *e = getenv("FFTW3_B200_DIST_COMM_CTAS");
    /* An NVLink-bound pass needs few CTAs to keep the links busy; launched with a full grid its CTAs would
       sit on every SM waiting for remote stores and starve the HBM-bound pass that runs next to it.  Both
       the copy kernels (grid-stride) and the specialised FFT kernels (persistent twins, fft_fast.cuh)
       honour the limit.  Default: two CTAs per SM. */
    return e ? atoi(e) : 2 * b2d_sm_count();
}

static void limit_grid(b2_plan *pl, int limit)
{
    int i;
    if (!pl || limit <= 0) return;
    for (i = 0; i < pl->nsteps; ++i) {
        if (pl->steps[i].kind == STEP_COPY) pl->steps[i].u.copy.grid_limit = limit;
        else if (pl->steps[i].kind == STEP_FFT) pl->steps[i].u.fft.grid_limit = limit;
    }
}

/* How the first exchange travels.  "stores": fused into the stores of the X pass (one kernel computes and
   scatters; its CTAs wait on NVLink and hold their SMs while they do).  "copy": X stores locally and the copy
   engines move the blocks, so no SM ever waits for a link; costs no extra HBM traffic for the block a rank
   keeps (X writes that one straight into its own exchange buffer) and one extra read of the rest. */
static int exchange_by_copy(int nranks)
{
    const char *e = getenv("FFTW3_B200_DIST_EXCHANGE");
    (void)nranks;
    if (e && !strcmp(e, "copy")) return 1;
    /* Measured on B200s behind NVSwitch (profiles/r02_dist_exchange_modes.log, r02_dist_partition.log): with one
       peer the copy engines keep up with the links (stage 0 = 8.27 ms at 1024^3 against 8.94 for fused stores sharing
       the SMs, 8.09 for fused stores on SMs of their own); with 3 or 7 peers their strided copies reach only
       ~450 GB/s against ~700 for the pass's own stores.  The fused stores on partitioned SMs win everywhere. */
    return 0;
}

/* SMs set aside for the NVLink-bound scatter pass of stage 0 (CUDA green contexts), the HBM-bound Y pass of the next
   chunk getting the others.  Sharing every SM between the two does not overlap them: the scatter's CTAs hold their
   stores behind the links and the Y CTAs next to them queue behind those (stage 0 = the sum of the two passes,
   profiles/r02_dist_overlap_sweep*_p2.log).  The scatter needs enough SMs to transform its rows at the link rate.
   Measured at 1024^3 on 148 SMs (profiles/r02_dist_partition.log): P = 2: 56 / 64 / 72 / 80 / 88 SMs -> stage 0 =
   9.39 / 8.47 / 8.09 / 8.70 / 9.55 ms (8.94 shared); P = 4: 40 / 56 / 72 -> 6.31 / 5.06 / 5.13 ms (5.81 shared). */
static int partition_sms(int nranks)
{
    const char *e = getenv("FFTW3_B200_DIST_PARTITION");
    int sms = b2d_sm_count(), k;
    if (e) return atoi(e);
    if (sms < 64) return 0;
    k = nranks == 2 ? (sms * 49) / 100 : (sms * 38) / 100;       /* 148 SMs: 72 / 56 */
    return (k / 8) * 8;
}

static int copy_pieces(int nranks)
{
    const char *e = getenv("FFTW3_B200_DIST_COPY_SPLIT");
    int k = e ? atoi(e) : 1;             /* more streams per block bought nothing (same log) */
    return k < 1 ? 1 : k > 6 ? 6 : k;
}

static int chunks_for(int64_t n)
{
    const char *e = getenv("FFTW3_B200_DIST_CHUNKS");
    int c = e ? atoi(e) : 8;      /* stage 0: the HBM-bound Y pass of chunk c + 1 overlaps the NVLink-bound X pass of chunk c */
    if (c < 1) c = 1;
    if (c > 64) c = 64;
    while (c > 1 && n / c < 2) c /= 2;
    return c;
}

/* Redirect the stores of a planned single-pass strided (COL) transform so that output row k
   goes to targets[k / rows] at row k % rows (row stride `os` reals): the exchange rides on the
   pass.  0 on success. */
static int rowsplit(b2_plan *pl, int nranks, int64_t rows, void *const *targets, int64_t target_off, int64_t os)
{
    b2d_fft_pass *f;
    int k, tile;
    if (pl->nsteps != 1 || pl->steps[0].kind != STEP_FFT) return -1;
    f = &pl->steps[0].u.fft;
    if (f->pre_op || f->post_op || f->bluestein || f->bn[2] != 1 || f->bos[0] != 2 || !f->store_col) return -1;
    f->npeer = nranks;
    f->peer_rows = (int)rows;
    for (k = 0; k < nranks; ++k) f->peer_out[k] = (double *)targets[k] + target_off;
    f->os = os;
    f->tw4_shift = -1;
    if ((rows & (rows - 1)) == 0) { int sh = 0; while (((int64_t)1 << sh) < rows) ++sh; f->tw4_shift = sh; }
    tile = f->kernel >= 1000 && f->kernel < 5000 ? f->kernel % 100 : 0;
    f->kernel = (tile && b2d_fast_available(f, 1800 + tile)) ? 1800 + tile : 0;
    return 0;
}

static dplan mkdist(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                    C *local, C *zbuf, void *const *push_targets, void *const *pull_sources,
                    void *const *out_targets, int sign, unsigned flags)
{
    dplan p;
    b2_problem q;
    int d, c;
    int64_t b1 = blk(n1, nranks);
    int64_t ln0 = share(n0, nranks, rank), ln1 = share(n1, nranks, rank);
    /* the single-launch scatter needs equal blocks and device-resident arrays (host arrays
       are staged per plan from the plan's own tensors, which describe one destination) */
    int even1 = (n1 % nranks == 0) && nranks <= B2D_MAX_PEERS && b2d_pointer_is_device(local) == 1;
    if (n0 <= 0 || n1 <= 0 || n2 <= 0 || nranks < 1 || rank < 0 || rank >= nranks) return NULL;
    if (sign != -1 && sign != 1) return NULL;
    p = (dplan)calloc(1, sizeof *p);
    if (!p) return NULL;
    p->nranks = nranks; p->rank = rank;
    if (out_targets && (pull_sources || nranks > B2D_MAX_PEERS || b2d_pointer_is_device(zbuf) != 1)) { free(p); return NULL; }
    p->nstages = pull_sources ? 3 : 2;
    p->ce = nranks > 1 && nranks <= B2D_MAX_PEERS && exchange_by_copy(nranks) && b2d_pointer_is_device(local) == 1
            && b2d_pointer_is_device(push_targets[rank]) == 1;
    if (p->ce) {
        p->ce_n1 = n1; p->ce_n2 = n2; p->ce_ln0 = ln0; p->ce_local = (double *)local;
        for (d = 0; d < nranks; ++d) p->ce_targets[d] = push_targets[d];
    }
    p->c0 = ln0 > 0 ? chunks_for(ln0) : 1;
    p->c1 = pull_sources ? chunks_for(b1) : 1;     /* from the block size: identical on every rank */
    if (!p->ce && nranks > 1 && p->c0 > 1 && b2d_pointer_is_device(local) == 1) {
        void *a, *b;
        int k = partition_sms(nranks);
        if (k >= 8 && !b2d_partition_streams(k, &a, &b)) p->part_sms = k;
    }
    p->y = (b2_plan **)calloc((size_t)p->c0, sizeof(b2_plan *));
    p->x = (b2_plan **)calloc((size_t)p->c0 * nranks, sizeof(b2_plan *));
    p->z = (b2_plan **)calloc((size_t)p->c1, sizeof(b2_plan *));
    p->g = (b2_plan **)calloc((size_t)p->c1 * nranks, sizeof(b2_plan *));
    if (!p->y || !p->x || !p->z || !p->g) goto fail;

    for (c = 0; c < p->c0; ++c) {
        int64_t lo = ln0 * c / p->c0, cnt = ln0 * (c + 1) / p->c0 - lo;   /* planes of this chunk */
        double *lin = (double *)local + 2 * lo * n1 * n2;
        /* Y: FFT along n1 in place on [cnt][n1][n2] */
        init_problem(&q, flags);
        dim(&q.sz, n1, 2 * n2, 2 * n2);
        dim(&q.vecsz, cnt, 2 * n1 * n2, 2 * n1 * n2);
        dim(&q.vecsz, n2, 2, 2);
        set_ptrs(&q, lin, lin, sign);
        p->y[c] = b2_mkplan(&q);
        if (!p->y[c]) goto fail;
        if (p->ce) {
            /* X by pieces: the rows of this rank's own block go straight into its exchange buffer, the rows of
               the blocks before / after it are transformed in place and handed to the copy engines */
            int64_t l1 = share(n1, nranks, rank), first[3], count[3];
            int k, slot = 0;
            first[0] = rank * b1; count[0] = l1;
            first[1] = 0; count[1] = rank * b1 < n1 ? rank * b1 : n1;
            first[2] = rank * b1 + l1; count[2] = n1 - first[2];
            for (k = 0; k < 3; ++k) {
                b2_plan *xp;
                if (count[k] <= 0 || cnt <= 0) continue;
                init_problem(&q, flags);
                dim(&q.sz, n2, 2, 2);
                dim(&q.vecsz, cnt, 2 * n1 * n2, k == 0 ? 2 * l1 * n2 : 2 * n1 * n2);
                dim(&q.vecsz, count[k], 2 * n2, 2 * n2);
                set_ptrs(&q, lin + 2 * first[k] * n2,
                         k == 0 ? (double *)push_targets[rank] + 2 * lo * l1 * n2 : lin + 2 * first[k] * n2, sign);
                xp = b2_mkplan(&q);
                if (!xp) goto fail;
                p->x[c * nranks + slot++] = xp;      /* nranks >= 2 slots; at most 2 pieces are non-empty when nranks == 2 */
            }
            continue;
        }
        /* X: FFT along n2, rows (i0, k1 in block d) -> push_targets[d] as [i0][k1'][k2] */
        for (d = 0; d < nranks; ++d) {
            int64_t l1 = share(n1, nranks, d);
            b2_plan *xp;
            init_problem(&q, flags);
            dim(&q.sz, n2, 2, 2);
            dim(&q.vecsz, cnt, 2 * n1 * n2, 2 * l1 * n2);
            dim(&q.vecsz, l1, 2 * n2, 2 * n2);
            set_ptrs(&q, lin + 2 * d * b1 * n2, (double *)push_targets[d] + 2 * lo * l1 * n2, sign);
            xp = b2_mkplan(&q);
            if (!xp) goto fail;
            p->x[c * nranks + d] = xp;
            if (nranks > 1 && p->c0 > 1 && !p->part_sms) limit_grid(xp, comm_ctas());
            if (d == 0 && even1 && nranks > 1 && cnt > 1 && l1 > 1 && xp->nsteps == 1 && xp->steps[0].kind == STEP_FFT
                && xp->steps[0].u.fft.bn[2] == 1 && xp->steps[0].u.fft.bn[0] == l1) {
                /* all destinations get equal blocks: let batch dim 2 walk the destinations and
                   select the peer buffer, so ONE launch scatters the whole chunk */
                b2d_fft_pass *f = &xp->steps[0].u.fft;
                int k;
                f->bn[2] = nranks; f->bis[2] = 2 * b1 * n2; f->bos[2] = 0;
                f->npeer = nranks;
                f->peer_rot = rank + 1;
                for (k = 0; k < nranks; ++k) f->peer_out[k] = (double *)push_targets[k] + 2 * lo * l1 * n2;
                /* the pass was planned as an ordinary row pass; the kernel chosen for it must also know how to
                   scatter over peers (the warp-per-transform kernel does not): pick a block-cooperative one */
                if (f->kernel >= 3000 || (f->kernel && !b2d_fast_available(f, f->kernel))) {
                    static const int codes[] = { 2, 4, 1, 102, 104, 101 };
                    size_t ci;
                    f->kernel = 0;
                    for (ci = 0; ci < sizeof codes / sizeof codes[0]; ++ci)
                        if (b2d_fast_available(f, codes[ci])) { f->kernel = codes[ci]; break; }
                }
                p->x_fused[c] = 1;
                break;
            }
        }
    }

    for (c = 0; c < p->c1; ++c) {
        /* columns (k1', k2) of this chunk: k1' in [lo, lo + cnt) */
        int64_t lo = ln1 * c / p->c1, cnt = ln1 * (c + 1) / p->c1 - lo;
        init_problem(&q, flags);
        if (pull_sources || out_targets) {
            /* Z in place on zbuf = [n0][ln1][n2] (push plans: planned like this, stores redirected below) */
            dim(&q.sz, n0, 2 * ln1 * n2, 2 * ln1 * n2);
            dim(&q.vecsz, cnt * n2, 2, 2);
            set_ptrs(&q, (double *)zbuf + 2 * lo * n2, (double *)zbuf + 2 * lo * n2, sign);
        } else {
            /* TRANSPOSED_OUT: [n0][ln1][n2] -> local as [ln1][n0][n2] */
            dim(&q.sz, n0, 2 * ln1 * n2, 2 * n2);
            dim(&q.vecsz, ln1, 2 * n2, 2 * n0 * n2);
            dim(&q.vecsz, n2, 2, 2);
            set_ptrs(&q, (double *)zbuf, (double *)local, sign);
        }
        p->z[c] = b2_mkplan(&q);
        if (!p->z[c]) goto fail;
        if (out_targets && !p->z[c]->is_nop) {
            /* push: row k0 of every column goes to its owner rank k0 / b0, into that rank's
               slab [ln0(owner)][n1][n2] at (k0 % b0, my first column + k1', k2) */
            int64_t b0 = blk(n0, nranks), s1 = b1 * rank;
            if (p->z[c]->nsteps != 1 || p->z[c]->steps[0].u.fft.bn[1] != 1) goto fail;   /* multi-pass n0: use the gather plan */
            if (rowsplit(p->z[c], nranks, b0, out_targets, 2 * (s1 + lo) * n2, 2 * n1 * n2)) goto fail;
        }
    }
    if (pull_sources) {
        /* gather back: block from rank s = [ln0][l1(s)][n2] -> local[i0][s*b1 + k1'][k2];
           chunk c takes the k1' range that rank s transformed in ITS chunk c */
        for (d = 0; d < nranks; ++d) {
            int64_t l1 = share(n1, nranks, d);
            for (c = 0; c < p->c1; ++c) {
                /* the same split every rank applies to its own columns in stage 1 */
                int64_t lo = l1 * c / p->c1, cnt = l1 * (c + 1) / p->c1 - lo;
                init_problem(&q, flags | B2F_ESTIMATE);
                dim(&q.vecsz, ln0, 2 * l1 * n2, 2 * n1 * n2);
                dim(&q.vecsz, cnt * n2, 2, 2);
                set_ptrs(&q, (double *)pull_sources[d] + 2 * lo * n2, (double *)local + 2 * (d * b1 + lo) * n2, -1);
                p->g[c * nranks + d] = b2_mkplan(&q);
                if (!p->g[c * nranks + d]) goto fail;
                if (nranks > 1 && p->c1 > 1) limit_grid(p->g[c * nranks + d], 2 * comm_ctas());
            }
        }
    }
    return p;
fail:
    fftw_b200_dist_destroy_plan(p);
    return NULL;
}

dplan fftw_b200_dist_plan_dft_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                                 C *local, C *zbuf, void *const *push_targets, void *const *pull_sources,
                                 int sign, unsigned flags)
{
    return mkdist(n0, n1, n2, rank, nranks, local, zbuf, push_targets, pull_sources, NULL, sign, flags);
}

/* natural-order output with BOTH exchanges fused into pass stores (no gather stage):
   out_targets[d] = rank d's slab `local` (peer-mapped).  NULL when the dim-0 transform
   needs more than one pass; the caller then uses the gather plan. */
dplan fftw_b200_dist_plan_dft_3d_push(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                                      C *local, C *zbuf, void *const *push_targets, void *const *out_targets,
                                      int sign, unsigned flags)
{
    if (!out_targets) return NULL;
    return mkdist(n0, n1, n2, rank, nranks, local, zbuf, push_targets, NULL, out_targets, sign, flags);
}

/* ---- real data: r2c / c2r of an n0 x n1 x n2 real array (mpi/rdft2-rank-geq2(-transposed).c,
   mpi/api.c:650-760).  Layout as fftw_mpi: the real slab is [local_n0][n1][2*(n2/2+1)] (padded
   rows, so it may alias the complex slab [local_n0][n1][n2/2+1]).
     r2c  stage 0: local r2c of the rows (n2), then c2c along n1 whose output ROWS are stored
                   straight into the peers' exchange buffers [n0][n1/P][h] (row-split stores)
          stage 1: c2c along n0, rows stored straight into the owners' complex slabs
     c2r  stages 0/1 the same with backward c2c passes, stage 2: local c2r of the rows.
   Uneven blocks: every exchange buffer uses the row pitch b1 = ceil(n1 / nranks). */
static dplan mkdist_real(int c2r, ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                         double *real, C *cplx, C *zbuf, void *const *push_targets, void *const *out_targets,
                         unsigned flags)
{
    dplan p;
    b2_problem q;
    int64_t h = n2 / 2 + 1, b0 = blk(n0, nranks), b1 = blk(n1, nranks);
    int64_t ln0 = share(n0, nranks, rank), ln1 = share(n1, nranks, rank);
    int sign = c2r ? +1 : -1;
    if (n0 <= 0 || n1 <= 0 || n2 <= 0 || nranks < 1 || rank < 0 || rank >= nranks || nranks > B2D_MAX_PEERS) return NULL;
    if (!push_targets || !out_targets) return NULL;
    if (b2d_pointer_is_device(zbuf) != 1 || b2d_pointer_is_device(cplx) != 1) return NULL;
    p = (dplan)calloc(1, sizeof *p);
    if (!p) return NULL;
    p->nranks = nranks; p->rank = rank; p->c0 = p->c1 = 1;
    p->nstages = c2r ? 3 : 2;
    p->y = (b2_plan **)calloc(1, sizeof(b2_plan *));
    p->x = (b2_plan **)calloc((size_t)nranks, sizeof(b2_plan *));
    p->z = (b2_plan **)calloc(1, sizeof(b2_plan *));
    p->g = (b2_plan **)calloc((size_t)nranks, sizeof(b2_plan *));
    if (!p->y || !p->x || !p->z || !p->g) goto fail;
    if (ln0 > 0) {
        /* the local real pass over the rows */
        init_problem(&q, flags | (c2r ? B2F_DESTROY_INPUT : 0));
        q.kind = c2r ? B2_C2R : B2_R2C;
        dim(&q.sz, n2, c2r ? 2 : 1, c2r ? 1 : 2);
        dim(&q.vecsz, ln0 * n1, 2 * h, 2 * h);
        if (c2r) { q.in0 = (double *)cplx; q.in1 = (double *)cplx + 1; q.out0 = real; }
        else { q.in0 = real; q.out0 = (double *)cplx; q.out1 = (double *)cplx + 1; }
        if (c2r) p->post = b2_mkplan(&q); else p->pre = b2_mkplan(&q);
        if (!(c2r ? p->post : p->pre)) goto fail;
        /* c2c along n1 on [ln0][n1][h], rows k1 -> peer k1 / b1: [n0][b1][h] at (my first plane + i0, k1 % b1, k2) */
        init_problem(&q, flags);
        dim(&q.sz, n1, 2 * h, 2 * h);
        dim(&q.vecsz, ln0, 2 * n1 * h, 2 * b1 * h);
        dim(&q.vecsz, h, 2, 2);
        set_ptrs(&q, (double *)cplx, (double *)zbuf, sign);       /* zbuf only stands in while planning */
        p->x[0] = b2_mkplan(&q);
        if (!p->x[0]) goto fail;
        if (p->x[0]->nsteps == 1 && !rowsplit(p->x[0], nranks, b1, push_targets, 0, 2 * h)) p->x_fused[0] = 1;
        else {
            /* n1 whose transform cannot split its stores by row (non-smooth: Bluestein / Rader, or multi-pass):
               transform in place on the local slab (runs as the "Y" of stage 0), then one strided copy per
               destination pushes its rows -- the reference's transpose as plain copies into peer memory
               (mpi/transpose-alltoall.c:49-100 with the all-to-all replaced by peer stores) */
            int d;
            b2_plan_destroy(p->x[0]); p->x[0] = NULL;
            init_problem(&q, flags);
            dim(&q.sz, n1, 2 * h, 2 * h);
            dim(&q.vecsz, ln0, 2 * n1 * h, 2 * n1 * h);
            dim(&q.vecsz, h, 2, 2);
            set_ptrs(&q, (double *)cplx, (double *)cplx, sign);
            p->y[0] = b2_mkplan(&q);
            if (!p->y[0]) goto fail;
            for (d = 0; d < nranks; ++d) {
                int64_t l1 = share(n1, nranks, d);
                if (!l1) continue;
                init_problem(&q, flags | B2F_ESTIMATE);
                dim(&q.vecsz, ln0, 2 * n1 * h, 2 * b1 * h);
                dim(&q.vecsz, l1 * h, 2, 2);
                set_ptrs(&q, (double *)cplx + 2 * d * b1 * h, (double *)push_targets[d], -1);
                p->x[d] = b2_mkplan(&q);
                if (!p->x[d]) goto fail;
            }
        }
    }
    if (ln1 > 0) {
        /* c2c along n0 on zbuf = [n0][b1][h] (row pitch b1 = the block size on every rank; this rank fills
           ln1 <= b1 of them), rows k0 -> owner k0 / b0: [ln0][n1][h] at (k0 % b0, my first column + k1', k2) */
        init_problem(&q, flags);
        dim(&q.sz, n0, 2 * b1 * h, 2 * b1 * h);
        dim(&q.vecsz, ln1 * h, 2, 2);
        set_ptrs(&q, (double *)zbuf, (double *)zbuf, sign);
        p->z[0] = b2_mkplan(&q);
        if (!p->z[0]) goto fail;
        if (p->z[0]->nsteps != 1 || p->z[0]->steps[0].kind != STEP_FFT || p->z[0]->steps[0].u.fft.bn[1] != 1 ||
            rowsplit(p->z[0], nranks, b0, out_targets, 2 * (b1 * rank) * h, 2 * n1 * h)) {
            /* same fallback for n0: in place on zbuf (the plan above, untouched by a failed rowsplit? no --
               replan it), then copy every owner's rows into its slab */
            int s2;
            b2_plan_destroy(p->z[0]);
            init_problem(&q, flags);
            dim(&q.sz, n0, 2 * b1 * h, 2 * b1 * h);
            dim(&q.vecsz, ln1 * h, 2, 2);
            set_ptrs(&q, (double *)zbuf, (double *)zbuf, sign);
            p->z[0] = b2_mkplan(&q);
            if (!p->z[0]) goto fail;
            p->zcopy = 1;
            for (s2 = 0; s2 < nranks; ++s2) {
                int64_t l0 = share(n0, nranks, s2);
                if (!l0) continue;
                init_problem(&q, flags | B2F_ESTIMATE);
                dim(&q.vecsz, l0, 2 * b1 * h, 2 * n1 * h);
                dim(&q.vecsz, ln1 * h, 2, 2);
                set_ptrs(&q, (double *)zbuf + 2 * (b0 * s2) * b1 * h, (double *)out_targets[s2] + 2 * (b1 * rank) * h, -1);
                p->g[s2] = b2_mkplan(&q);
                if (!p->g[s2]) goto fail;
            }
        }
    }
    return p;
fail:
    fftw_b200_dist_destroy_plan(p);
    return NULL;
}

dplan fftw_b200_dist_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                                     double *real_in, C *cplx_out, C *zbuf, void *const *push_targets,
                                     void *const *out_targets, unsigned flags)
{
    return mkdist_real(0, n0, n1, n2, rank, nranks, real_in, cplx_out, zbuf, push_targets, out_targets, flags);
}

dplan fftw_b200_dist_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                                     C *cplx_in, double *real_out, C *zbuf, void *const *push_targets,
                                     void *const *out_targets, unsigned flags)
{
    return mkdist_real(1, n0, n1, n2, rank, nranks, real_out, cplx_in, zbuf, push_targets, out_targets, flags);
}

/* ---- r2r (fftw_mpi_plan_r2r_3d, mpi/rdft-rank-geq2(-transposed).c, mpi/api.c:770-886): a real
   n0 x n1 x n2 array, kind[i] along dimension i.  Built from validated pieces only:
     stage 0  local r2r over (n1, n2) in place on [ln0][n1][n2]
     stage 1  pull my column block from every rank's slab into zbuf = [n0][ln1][n2] (peer loads),
              then r2r along n0 in place on zbuf
     stage 2  pull my planes of every column block back from the peers' zbufs into the slab
   (the global transposes of mpi/transpose-alltoall.c as gather copies; real rows are not yet
   scattered by the passes themselves).  The caller puts a barrier before every stage. */
dplan fftw_b200_dist_plan_r2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                                 double *local, double *zbuf, void *const *peer_locals, void *const *peer_zbufs,
                                 const int *kinds, unsigned flags)
{
    dplan p;
    b2_problem q;
    int s;
    int64_t b0 = blk(n0, nranks), b1 = blk(n1, nranks);
    int64_t ln0 = share(n0, nranks, rank), ln1 = share(n1, nranks, rank);
    if (n0 <= 0 || n1 <= 0 || n2 <= 0 || nranks < 1 || rank < 0 || rank >= nranks || !kinds) return NULL;
    if (!peer_locals || !peer_zbufs) return NULL;
    p = (dplan)calloc(1, sizeof *p);
    if (!p) return NULL;
    p->nranks = nranks; p->rank = rank; p->c0 = p->c1 = 1;
    p->nstages = 3;
    p->y = (b2_plan **)calloc(1, sizeof(b2_plan *));
    p->x = (b2_plan **)calloc((size_t)nranks, sizeof(b2_plan *));      /* stage 1 gathers, one per source */
    p->z = (b2_plan **)calloc(1, sizeof(b2_plan *));
    p->g = (b2_plan **)calloc((size_t)nranks, sizeof(b2_plan *));      /* stage 2 gathers */
    if (!p->y || !p->x || !p->z || !p->g) goto fail;
    p->real_gather = 1;
    if (ln0 > 0) {
        init_problem(&q, flags);
        q.kind = B2_R2R;
        dim(&q.sz, n1, n2, n2); dim(&q.sz, n2, 1, 1);
        q.r2r_kind[0] = kinds[1]; q.r2r_kind[1] = kinds[2];
        dim(&q.vecsz, ln0, n1 * n2, n1 * n2);
        q.in0 = local; q.out0 = local;
        p->y[0] = b2_mkplan(&q);
        if (!p->y[0]) goto fail;
    }
    if (ln1 > 0) {
        for (s = 0; s < nranks; ++s) {
            /* block of rank s: [ln0(s)][ln1][n2] at column b1*rank of its slab -> zbuf planes b0*s ... */
            int64_t l0 = share(n0, nranks, s);
            if (!l0) continue;
            init_problem(&q, flags | B2F_ESTIMATE);
            q.kind = B2_R2R;                                           /* rank 0: a copy of reals */
            dim(&q.vecsz, l0, n1 * n2, ln1 * n2);
            dim(&q.vecsz, ln1 * n2, 1, 1);
            q.in0 = (double *)peer_locals[s] + b1 * rank * n2;
            q.out0 = zbuf + b0 * s * ln1 * n2;
            p->x[s] = b2_mkplan(&q);
            if (!p->x[s]) goto fail;
        }
        init_problem(&q, flags);
        q.kind = B2_R2R;
        dim(&q.sz, n0, ln1 * n2, ln1 * n2);
        q.r2r_kind[0] = kinds[0];
        dim(&q.vecsz, ln1 * n2, 1, 1);
        q.in0 = zbuf; q.out0 = zbuf;
        p->z[0] = b2_mkplan(&q);
        if (!p->z[0]) goto fail;
    }
    if (ln0 > 0) {
        for (s = 0; s < nranks; ++s) {
            /* my planes of column block s: rank s's zbuf [n0][l1(s)][n2] planes b0*rank ... -> slab column b1*s */
            int64_t l1 = share(n1, nranks, s);
            if (!l1) continue;
            init_problem(&q, flags | B2F_ESTIMATE);
            q.kind = B2_R2R;
            dim(&q.vecsz, ln0, l1 * n2, n1 * n2);
            dim(&q.vecsz, l1 * n2, 1, 1);
            q.in0 = (double *)peer_zbufs[s] + b0 * rank * l1 * n2;
            q.out0 = local + b1 * s * n2;
            p->g[s] = b2_mkplan(&q);
            if (!p->g[s]) goto fail;
        }
    }
    return p;
fail:
    fftw_b200_dist_destroy_plan(p);
    return NULL;
}

ptrdiff_t fftw_b200_ipc_offset(void *devptr) { return (ptrdiff_t)b2d_alloc_offset(devptr); }

int fftw_b200_dist_num_stages(const dplan p) { return p->nstages; }
int fftw_b200_dist_exchange_by_copy(const dplan p) { return p->ce; }
int fftw_b200_dist_partition_sms(const dplan p) { return p->part_sms; }
int fftw_b200_dist_num_chunks(const dplan p, int stage) { return stage == 0 ? p->c0 : p->c1; }

static void run(b2_plan *pl)
{
    if (pl) b2_execute_ex(pl, pl->prob.in0, pl->prob.in1, pl->prob.out0, pl->prob.out1, 1);   /* enqueue only */
}

/* one chunk of a stage.  Stage 0: Y_c on the caller's stream, X_c on side stream 0.
   Stage 1: Z_c on the caller's stream.  Stage 2: gather_c on side stream 1, after
   everything the caller's stream holds so far (the caller's barrier included). */
void fftw_b200_dist_execute_chunk(const dplan p, int stage, int c)
{
    int d;
    void *mainst = b2d_get_stream(), *prev = NULL;
    if (p->real_gather) {
        if (stage == 0) run(p->y[0]);
        else if (stage == 1) {
            for (d = 0; d < p->nranks; ++d) run(p->x[(p->rank + d) % p->nranks]);
            run(p->z[0]);
        } else if (stage == 2) {
            for (d = 0; d < p->nranks; ++d) run(p->g[(p->rank + d) % p->nranks]);
        }
    } else if (stage == 2 && p->post) {
        if (c == 0) run(p->post);
    } else if (stage == 0 && c < p->c0) {
        void *aux = b2d_aux_stream(0);
        if (c == 0) run(p->pre);
        if (p->part_sms) {
            /* partitioned SMs: Y of every chunk on the compute partition's stream, X (fused scatter) of the chunk on
               the communication partition's stream right behind its Y; the caller's stream joins both at the end */
            void *cs = NULL, *ys = NULL;
            if (!b2d_partition_streams(p->part_sms, &cs, &ys)) {
                if (c == 0) b2d_stream_wait_stream(ys, mainst);
                prev = b2d_push_stream(ys);
                run(p->y[c]);
                b2d_pop_stream(prev);
                b2d_stream_wait_stream(cs, ys);
                prev = b2d_push_stream(cs);
                if (p->x_fused[c]) run(p->x[c * p->nranks]);
                else for (d = 0; d < p->nranks; ++d) run(p->x[c * p->nranks + (p->rank + 1 + d) % p->nranks]);
                b2d_pop_stream(prev);
                return;
            }
        }
        run(p->y[c]);
        if (p->ce) {
            int64_t n1 = p->ce_n1, n2 = p->ce_n2, b1 = blk(n1, p->nranks);
            int64_t lo = p->ce_ln0 * c / p->c0, cnt = p->ce_ln0 * (c + 1) / p->c0 - lo;
            for (d = 0; d < p->nranks && d < 3; ++d) run(p->x[c * p->nranks + d]);
            for (d = 1; d < p->nranks && cnt > 0; ++d) {
                int t = (p->rank + d) % p->nranks, k;
                int64_t l1 = share(n1, p->nranks, t);
                /* one copy engine does not fill the links: cut the planes of a block over several streams */
                int pieces = copy_pieces(p->nranks);
                if (l1 <= 0) continue;
                if (pieces > cnt) pieces = (int)cnt;
                for (k = 0; k < pieces; ++k) {
                    int64_t a = cnt * k / pieces, b = cnt * (k + 1) / pieces;
                    void *cs = b2d_aux_stream(2 + ((d - 1) * pieces + k) % 6);
                    if (cs) b2d_stream_wait_stream(cs, mainst);
                    b2d_memcpy2d_async((double *)p->ce_targets[t] + 2 * (lo + a) * l1 * n2, (size_t)(l1 * n2) * sizeof(C),
                                       p->ce_local + 2 * ((lo + a) * n1 + t * b1) * n2, (size_t)(n1 * n2) * sizeof(C),
                                       (size_t)(l1 * n2) * sizeof(C), (size_t)(b - a), cs ? cs : mainst);
                }
            }
            return;
        }
        if (aux) { b2d_stream_wait_stream(aux, mainst); prev = b2d_push_stream(aux); }
        if (p->x_fused[c]) run(p->x[c * p->nranks]);
        else for (d = 0; d < p->nranks; ++d) run(p->x[c * p->nranks + (p->rank + 1 + d) % p->nranks]);
        if (aux) b2d_pop_stream(prev);
    } else if (stage == 1 && c < p->c1) {
        run(p->z[c]);
        if (p->zcopy) for (d = 0; d < p->nranks; ++d) run(p->g[(p->rank + 1 + d) % p->nranks]);
    } else if (stage == 2 && c < p->c1) {
        void *aux = b2d_aux_stream(1);
        if (aux) { b2d_stream_wait_stream(aux, mainst); prev = b2d_push_stream(aux); }
        for (d = 0; d < p->nranks; ++d) run(p->g[c * p->nranks + (p->rank + 1 + d) % p->nranks]);
        if (aux) b2d_pop_stream(prev);
    }
}

/* make the caller's stream wait for the side streams */
void fftw_b200_dist_join(const dplan p)
{
    void *mainst = b2d_get_stream();
    int i;
    for (i = 0; i < 8; ++i) {
        void *aux = b2d_aux_stream(i);
        if (aux) b2d_stream_wait_stream(mainst, aux);
    }
    if (p->part_sms) {
        void *cs = NULL, *ys = NULL;
        if (!b2d_partition_streams(p->part_sms, &cs, &ys)) {
            b2d_stream_wait_stream(mainst, cs);
            b2d_stream_wait_stream(mainst, ys);
        }
    }
    if (!b2_async_mode) b2d_sync();
}

void fftw_b200_dist_execute_stage(const dplan p, int stage)
{
    int c, n = fftw_b200_dist_num_chunks(p, stage);
    for (c = 0; c < n; ++c) fftw_b200_dist_execute_chunk(p, stage, c);
    fftw_b200_dist_join(p);
}

void *fftw_b200_device_malloc(size_t bytes) { return b2d_malloc(bytes); }
void fftw_b200_device_free(void *p) { b2d_free(p); }
int fftw_b200_ipc_export(void *devptr, unsigned char handle[64]) { return b2d_ipc_export(devptr, handle); }
void *fftw_b200_ipc_import(const unsigned char handle[64]) { return b2d_ipc_import(handle); }
void fftw_b200_ipc_close(void *devptr) { b2d_ipc_close(devptr); }
