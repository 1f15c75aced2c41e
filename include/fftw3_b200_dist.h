/* fftw3_b200_dist.h -- slab-decomposed multi-GPU transforms, one process per GPU.
 *
 * This is the B200 counterpart of the reference's MPI interface
 * (mpi/fftw3-mpi.h:58-215): the data distribution, the local sizes and the
 * TRANSPOSED_OUT option are the same; the communicator is replaced by
 * (rank, nranks) plus plain device pointers, because the exchange runs over
 * NVLink either as direct peer stores/loads into CUDA-IPC-mapped buffers or as an
 * NCCL all-to-all issued by the caller between stages.
 *
 *   reference                                  here
 *   fftw_mpi_local_size_3d_transposed   ->     fftw_b200_dist_local_size_3d      (mpi/api.c:248-352)
 *   fftw_mpi_plan_dft_3d                ->     fftw_b200_dist_plan_dft_3d        (mpi/api.c:560-648,
 *                                               mpi/dft-rank-geq2-transposed.c:47-70,113-214)
 *   fftw_execute (of an MPI plan)       ->     fftw_b200_dist_execute_stage x N with a barrier / all-to-all
 *                                               between stages (mpi/transpose-alltoall.c:49-100)
 *
 * Distribution (mpi/block.c:37-50): block = ceil(n / nranks); rank r owns the
 * index range [r*block, min(n, (r+1)*block)) of the distributed dimension --
 * dimension 0 for the input, dimension 1 for the transposed intermediate.
 */
#ifndef FFTW3_B200_DIST_H
#define FFTW3_B200_DIST_H

#include <stddef.h>
#include "fftw3.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fftw_b200_dist_plan_s *fftw_b200_dist_plan;

/* Elements (complex) each rank must allocate for its local array and for each
 * exchange buffer; also returns the rank's share of dim 0 and of dim 1. */
ptrdiff_t fftw_b200_dist_local_size_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                                       ptrdiff_t *local_n0, ptrdiff_t *local_0_start,
                                       ptrdiff_t *local_n1, ptrdiff_t *local_1_start);

/* Plan a forward/backward c2c double transform of an n0 x n1 x n2 array.
 *   local         this rank's slab  [local_n0][n1][n2], transformed in place
 *   zbuf          this rank's exchange buffer, receives [n0][local_n1][n2]
 *   push_targets  nranks pointers: where the block destined to rank d must be
 *                 written -- the peer's zbuf (+ this rank's row offset) for the
 *                 direct NVLink path, or a slice of a local send buffer when the
 *                 caller runs an all-to-all between the stages
 *   pull_sources  nranks pointers to the blocks coming back from rank s for the
 *                 natural-order output, or NULL for FFTW_MPI_TRANSPOSED_OUT
 *                 semantics (result left as [local_n1][n0][n2] in `local`)
 * Stages: 0 = local 2-D transforms fused with the scatter of the exchange,
 *         1 = transforms along dim 0, 2 = gather back (natural order only).
 * The caller synchronises the ranks between stages. */
fftw_b200_dist_plan fftw_b200_dist_plan_dft_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                                               fftw_complex *local, fftw_complex *zbuf,
                                               void *const *push_targets, void *const *pull_sources,
                                               int sign, unsigned flags);
/* Natural-order output with BOTH exchanges fused into pass stores: the dim-0 pass of
 * stage 1 writes output row k0 straight into the slab of its owner (rank k0 / block),
 * so there is no gather stage (2 stages; the caller ends with a barrier: a rank's slab
 * is complete once EVERY rank has finished stage 1).
 *   out_targets   nranks pointers: rank d's `local` slab (peer-mapped, e.g. CUDA IPC)
 * Returns NULL when the dim-0 transform needs more than one pass (use the plan above). */
fftw_b200_dist_plan fftw_b200_dist_plan_dft_3d_push(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                                                    fftw_complex *local, fftw_complex *zbuf,
                                                    void *const *push_targets, void *const *out_targets,
                                                    int sign, unsigned flags);
/* Real data (fftw_mpi_plan_dft_r2c_3d / _c2r_3d, mpi/api.c:650-760): the real slab is
 * [local_n0][n1][2*(n2/2+1)] (padded rows; it may alias the complex slab), the complex slab
 * [local_n0][n1][n2/2+1]; zbuf holds [n0][b1][n2/2+1] with b1 = ceil(n1/nranks) on every rank.
 * Both exchanges ride on pass stores:
 *   push_targets[d] = rank d's zbuf + 2*(first plane of this rank)*b1*(n2/2+1) doubles,
 *   out_targets[d]  = rank d's complex slab.
 * r2c: 2 stages (local r2c rows + dim-1 pass scattering rows | dim-0 pass pushing rows), then a
 * barrier.  c2r: the same two stages backward, a barrier, then stage 2 = local c2r of the rows.
 * Requires single-pass n0 and n1; NULL otherwise. */
fftw_b200_dist_plan fftw_b200_dist_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                                                   double *real_in, fftw_complex *cplx_out, fftw_complex *zbuf,
                                                   void *const *push_targets, void *const *out_targets,
                                                   unsigned flags);
fftw_b200_dist_plan fftw_b200_dist_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                                                   fftw_complex *cplx_in, double *real_out, fftw_complex *zbuf,
                                                   void *const *push_targets, void *const *out_targets,
                                                   unsigned flags);
/* r2r (fftw_mpi_plan_r2r_3d, mpi/api.c:770-886): real [local_n0][n1][n2] slab transformed in place, kinds[i]
 * (fftw_r2r_kind values) along dimension i; zbuf holds [n0][local_n1][n2] reals.  Three stages (local
 * 2-d r2r | gather columns + r2r along dim 0 | gather back) with a barrier BEFORE each of them;
 *   peer_locals[s] = rank s's slab, peer_zbufs[s] = rank s's zbuf (peer-mapped). */
fftw_b200_dist_plan fftw_b200_dist_plan_r2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int rank, int nranks,
                                               double *local, double *zbuf, void *const *peer_locals,
                                               void *const *peer_zbufs, const int *kinds, unsigned flags);
int  fftw_b200_dist_num_stages(const fftw_b200_dist_plan p);
void fftw_b200_dist_execute_stage(const fftw_b200_dist_plan p, int stage);
/* Finer control for overlapping the exchange with compute: every stage is cut
 * into chunks (planes for stage 0, columns for stages 1 and 2; the chunk count of
 * stages 1/2 is the same on every rank).  Stage-0 chunks pipeline internally
 * (scatter of chunk c on a side stream under the transforms of chunk c+1).
 * For natural-order output the caller runs, per chunk c:
 *     execute_chunk(p, 1, c); <barrier on the current stream>; execute_chunk(p, 2, c);
 * and finally fftw_b200_dist_join(p): the gather of chunk c (side stream, NVLink
 * bound) then overlaps the dim-0 transforms of chunk c+1. */
int  fftw_b200_dist_num_chunks(const fftw_b200_dist_plan p, int stage);
/* how the first exchange of this plan travels: 1 = the copy engines move the blocks of a finished chunk while the SMs
 * transform the next one (default for device-resident slabs), 0 = fused into the stores of the row pass
 * (FFTW3_B200_DIST_EXCHANGE=stores, and every plan whose arrays live on the host) */
int  fftw_b200_dist_exchange_by_copy(const fftw_b200_dist_plan p);
/* > 0: stage 0 of this plan runs its NVLink-bound scatter pass on that many SMs of their own (a CUDA green context)
 * and the HBM-bound pass of the next chunk on the remaining SMs (FFTW3_B200_DIST_PARTITION=0 turns it off) */
int  fftw_b200_dist_partition_sms(const fftw_b200_dist_plan p);
void fftw_b200_dist_execute_chunk(const fftw_b200_dist_plan p, int stage, int chunk);
void fftw_b200_dist_join(const fftw_b200_dist_plan p);
void fftw_b200_dist_destroy_plan(fftw_b200_dist_plan p);

/* ======================================================================================================
 * Communicator interface: the fftw_mpi_* shapes (mpi/fftw3-mpi.h:58-215) with the MPI_Comm replaced by the
 * ONE collective the library needs from the launcher -- a blocking all-gather of a few hundred bytes of
 * host memory, used while planning only.  The library then allocates its exchange buffer, maps every
 * rank's exchange buffer and slab into every other rank (CUDA IPC over NVLink), and execution owns its
 * synchronisation: the ranks meet in device-side barriers (flags in peer memory), with no host round trip.
 *
 *   reference (mpi/api.c)                          here
 *   fftw_mpi_local_size(_many)(_transposed)  ->    fftw_b200_mpi_local_size(_many)(_transposed)   :248-352
 *   fftw_mpi_local_size_2d/_3d(_transposed)  ->    fftw_b200_mpi_local_size_2d/_3d(_transposed)
 *   fftw_mpi_plan_many_dft / plan_dft        ->    fftw_b200_mpi_plan_many_dft / plan_dft          :560-648
 *   fftw_mpi_plan_dft_2d / _3d               ->    fftw_b200_mpi_plan_dft_2d / _3d
 *   fftw_execute(mpi plan)                   ->    fftw_b200_mpi_execute                           :889-907
 *   fftwf_mpi_*                              ->    fftwf_b200_mpi_*  (single precision)
 * Arrays are DEVICE memory obtained from cudaMalloc (or fftw_b200_device_malloc): `in` / `out` hold this
 * rank's slab [local_n0][n1]...[howmany] with room for the number of elements local_size returns;
 * FFTW_MPI_TRANSPOSED_OUT leaves [local_n1][n0]..., FFTW_MPI_TRANSPOSED_IN expects the input that way (the two
 * combine); in == out for an in-place transform.  block / tblock:
 * FFTW_MPI_DEFAULT_BLOCK (0) = ceil(n / P), or the caller's own block sizes (block * P >= n0, tblock * P >= n1; the
 * plan then takes the general path).  Planning and execution are collective.
 * ====================================================================================================== */
#define FFTW_MPI_DEFAULT_BLOCK (0)
#define FFTW_MPI_SCRAMBLED_IN (1U << 27)
#define FFTW_MPI_SCRAMBLED_OUT (1U << 28)
#define FFTW_MPI_TRANSPOSED_IN (1U << 29)
#define FFTW_MPI_TRANSPOSED_OUT (1U << 30)

typedef struct fftw_b200_comm {
    int rank, nranks;
    /* gather `bytes` bytes from every rank into recv (rank order, nranks * bytes); 0 on success.  With MPI:
       MPI_Allgather(send, bytes, MPI_BYTE, recv, bytes, MPI_BYTE, *(MPI_Comm *)ctx) */
    int (*allgather)(void *ctx, const void *send, void *recv, size_t bytes);
    void *ctx;
} fftw_b200_comm;

typedef struct fftw_b200_mpi_plan_s *fftw_b200_mpi_plan;

ptrdiff_t fftw_b200_mpi_local_size_many_transposed(int rnk, const ptrdiff_t *n, ptrdiff_t howmany,
                                                   ptrdiff_t block0, ptrdiff_t block1, const fftw_b200_comm *comm,
                                                   ptrdiff_t *local_n0, ptrdiff_t *local_0_start,
                                                   ptrdiff_t *local_n1, ptrdiff_t *local_1_start);
ptrdiff_t fftw_b200_mpi_local_size_many(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t block0,
                                        const fftw_b200_comm *comm, ptrdiff_t *local_n0, ptrdiff_t *local_0_start);
ptrdiff_t fftw_b200_mpi_local_size(int rnk, const ptrdiff_t *n, const fftw_b200_comm *comm,
                                   ptrdiff_t *local_n0, ptrdiff_t *local_0_start);
ptrdiff_t fftw_b200_mpi_local_size_2d(ptrdiff_t n0, ptrdiff_t n1, const fftw_b200_comm *comm,
                                      ptrdiff_t *local_n0, ptrdiff_t *local_0_start);
ptrdiff_t fftw_b200_mpi_local_size_2d_transposed(ptrdiff_t n0, ptrdiff_t n1, const fftw_b200_comm *comm,
                                                 ptrdiff_t *local_n0, ptrdiff_t *local_0_start,
                                                 ptrdiff_t *local_n1, ptrdiff_t *local_1_start);
ptrdiff_t fftw_b200_mpi_local_size_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, const fftw_b200_comm *comm,
                                      ptrdiff_t *local_n0, ptrdiff_t *local_0_start);
ptrdiff_t fftw_b200_mpi_local_size_3d_transposed(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, const fftw_b200_comm *comm,
                                                 ptrdiff_t *local_n0, ptrdiff_t *local_0_start,
                                                 ptrdiff_t *local_n1, ptrdiff_t *local_1_start);

fftw_b200_mpi_plan fftw_b200_mpi_plan_many_dft(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t block,
                                               ptrdiff_t tblock, fftw_complex *in, fftw_complex *out,
                                               const fftw_b200_comm *comm, int sign, unsigned flags);
fftw_b200_mpi_plan fftw_b200_mpi_plan_dft(int rnk, const ptrdiff_t *n, fftw_complex *in, fftw_complex *out,
                                          const fftw_b200_comm *comm, int sign, unsigned flags);
fftw_b200_mpi_plan fftw_b200_mpi_plan_dft_2d(ptrdiff_t n0, ptrdiff_t n1, fftw_complex *in, fftw_complex *out,
                                             const fftw_b200_comm *comm, int sign, unsigned flags);
fftw_b200_mpi_plan fftw_b200_mpi_plan_dft_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex *in, fftw_complex *out,
                                             const fftw_b200_comm *comm, int sign, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_many_dft(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t block,
                                                ptrdiff_t tblock, fftwf_complex *in, fftwf_complex *out,
                                                const fftw_b200_comm *comm, int sign, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_dft(int rnk, const ptrdiff_t *n, fftwf_complex *in, fftwf_complex *out,
                                           const fftw_b200_comm *comm, int sign, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_dft_2d(ptrdiff_t n0, ptrdiff_t n1, fftwf_complex *in, fftwf_complex *out,
                                              const fftw_b200_comm *comm, int sign, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_dft_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftwf_complex *in,
                                              fftwf_complex *out, const fftw_b200_comm *comm, int sign, unsigned flags);
/* Real data of any rank >= 3 and any howmany, both precisions (fftw_mpi_plan_many_dft_r2c / _c2r, mpi/api.c:650-760;
 * rank 2 with howmany 1 takes the 2-D plan above): real slab [local_n0][n1]...[2 (n_last/2+1)][howmany] (padded),
 * complex slab [local_n0][n1]...[n_last/2+1][howmany]; default blocks, natural layouts; c2r overwrites its input. */
fftw_b200_mpi_plan fftw_b200_mpi_plan_many_dft_r2c(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                                   double *in, fftw_complex *out, const fftw_b200_comm *comm, unsigned flags);
fftw_b200_mpi_plan fftw_b200_mpi_plan_many_dft_c2r(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                                   fftw_complex *in, double *out, const fftw_b200_comm *comm, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_many_dft_r2c(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                                    float *in, fftwf_complex *out, const fftw_b200_comm *comm, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_many_dft_c2r(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                                    fftwf_complex *in, float *out, const fftw_b200_comm *comm, unsigned flags);
/* r2r of any rank >= 2 and any howmany (fftw_mpi_plan_many_r2r / fftw_mpi_plan_r2r_2d, mpi/api.c:770-886): slab
 * [local_n0][n1]...[howmany] reals, kind[i] along dimension i; default blocks; in == out or out of place. */
fftw_b200_mpi_plan fftw_b200_mpi_plan_many_r2r(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                               double *in, double *out, const fftw_b200_comm *comm,
                                               const fftw_r2r_kind *kind, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_many_r2r(int rnk, const ptrdiff_t *n, ptrdiff_t howmany, ptrdiff_t iblock, ptrdiff_t oblock,
                                                float *in, float *out, const fftw_b200_comm *comm,
                                                const fftwf_r2r_kind *kind, unsigned flags);
fftw_b200_mpi_plan fftw_b200_mpi_plan_r2r_2d(ptrdiff_t n0, ptrdiff_t n1, double *in, double *out, const fftw_b200_comm *comm,
                                             fftw_r2r_kind kind0, fftw_r2r_kind kind1, unsigned flags);
/* single precision: the basic real-data and r2r forms (through the general plans) */
fftw_b200_mpi_plan fftwf_b200_mpi_plan_dft_r2c_2d(ptrdiff_t n0, ptrdiff_t n1, float *in, fftwf_complex *out,
                                                  const fftw_b200_comm *comm, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_dft_c2r_2d(ptrdiff_t n0, ptrdiff_t n1, fftwf_complex *in, float *out,
                                                  const fftw_b200_comm *comm, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, float *in, fftwf_complex *out,
                                                  const fftw_b200_comm *comm, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftwf_complex *in, float *out,
                                                  const fftw_b200_comm *comm, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_r2r_2d(ptrdiff_t n0, ptrdiff_t n1, float *in, float *out, const fftw_b200_comm *comm,
                                              fftwf_r2r_kind kind0, fftwf_r2r_kind kind1, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_r2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, float *in, float *out,
                                              const fftw_b200_comm *comm, fftwf_r2r_kind kind0, fftwf_r2r_kind kind1,
                                              fftwf_r2r_kind kind2, unsigned flags);
/* Wisdom across ranks (fftw_mpi_gather_wisdom / fftw_mpi_broadcast_wisdom, mpi/wisdom-api.c): after gather rank 0
 * holds the union of every rank's wisdom; after broadcast every rank has imported rank 0's.  Collective. */
void fftw_b200_mpi_gather_wisdom(const fftw_b200_comm *comm);
void fftw_b200_mpi_broadcast_wisdom(const fftw_b200_comm *comm);
void fftwf_b200_mpi_gather_wisdom(const fftw_b200_comm *comm);
void fftwf_b200_mpi_broadcast_wisdom(const fftw_b200_comm *comm);
/* Real data and r2r in 3-D (fftw_mpi_plan_dft_r2c_3d / _c2r_3d / fftw_mpi_plan_r2r_3d, mpi/api.c:650-760, 770-886),
 * double precision, default blocks.  Real slab [local_n0][n1][2 (n2/2+1)] doubles (padded rows; it may alias the
 * complex slab [local_n0][n1][n2/2+1] for an in-place transform); allocate with
 * fftw_b200_mpi_local_size_3d(n0, n1, n2/2+1, ...) complex elements.  c2r overwrites its input (as fftw_mpi's does).
 * r2r: [local_n0][n1][n2] doubles, kindK along dimension K; out != in copies first. */
fftw_b200_mpi_plan fftw_b200_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, double *in, fftw_complex *out,
                                                 const fftw_b200_comm *comm, unsigned flags);
/* 2-D real data (fftw_mpi_plan_dft_r2c_2d / _c2r_2d): real slab [local_n0][2 (n1/2+1)], complex slab
 * [local_n0][n1/2+1]; the halved dimension is the one exchanged, as in fftw_mpi.  Allocate with
 * fftw_b200_mpi_local_size_2d(n0, n1/2+1, ...) complex elements.  Natural layouts only. */
fftw_b200_mpi_plan fftw_b200_mpi_plan_dft_r2c_2d(ptrdiff_t n0, ptrdiff_t n1, double *in, fftw_complex *out,
                                                 const fftw_b200_comm *comm, unsigned flags);
fftw_b200_mpi_plan fftw_b200_mpi_plan_dft_c2r_2d(ptrdiff_t n0, ptrdiff_t n1, fftw_complex *in, double *out,
                                                 const fftw_b200_comm *comm, unsigned flags);
fftw_b200_mpi_plan fftw_b200_mpi_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex *in, double *out,
                                                 const fftw_b200_comm *comm, unsigned flags);
fftw_b200_mpi_plan fftw_b200_mpi_plan_r2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, double *in, double *out,
                                             const fftw_b200_comm *comm, fftw_r2r_kind kind0, fftw_r2r_kind kind1,
                                             fftw_r2r_kind kind2, unsigned flags);
/* Distributed 1-D transform of n0 = r * m points (six-step with three global transposes; mpi/dft-rank1.c:81-148,
 * mpi/api.c:248-352 local_size_1d).  Input: this rank's local_ni consecutive points starting at local_i_start;
 * output: local_no points starting at local_o_start (the two distributions differ: rows of the r x m view on
 * input, rows of the m x r view on output).  FFTW_MPI_SCRAMBLED_OUT skips the last transpose (output element
 * X[k1 + r k2] stays at [k1][k2] in this rank's k1 block); FFTW_MPI_SCRAMBLED_IN takes its input in that layout and
 * leaves natural order distributed like the input (two transposes; n0 / r must fit one pass).  n0 must be
 * composite with a smooth factor <= the one-pass limit, else 0 / NULL (as the reference, n0 must be composite). */
ptrdiff_t fftw_b200_mpi_local_size_1d(ptrdiff_t n0, const fftw_b200_comm *comm, int sign, unsigned flags,
                                      ptrdiff_t *local_ni, ptrdiff_t *local_i_start,
                                      ptrdiff_t *local_no, ptrdiff_t *local_o_start);
fftw_b200_mpi_plan fftw_b200_mpi_plan_dft_1d(ptrdiff_t n0, fftw_complex *in, fftw_complex *out, const fftw_b200_comm *comm,
                                             int sign, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_dft_1d(ptrdiff_t n0, fftwf_complex *in, fftwf_complex *out, const fftw_b200_comm *comm,
                                              int sign, unsigned flags);
/* Distributed transpose of an n0 x n1 matrix of howmany-tuples of reals (mpi/api.c:521-556): rows block-distributed
 * on input ([local_n0][n1][howmany]) and on output ([local_n1][n0][howmany]); in == out allowed. */
fftw_b200_mpi_plan fftw_b200_mpi_plan_many_transpose(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t howmany, ptrdiff_t block0,
                                                     ptrdiff_t block1, double *in, double *out,
                                                     const fftw_b200_comm *comm, unsigned flags);
fftw_b200_mpi_plan fftw_b200_mpi_plan_transpose(ptrdiff_t n0, ptrdiff_t n1, double *in, double *out,
                                                const fftw_b200_comm *comm, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_many_transpose(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t howmany, ptrdiff_t block0,
                                                      ptrdiff_t block1, float *in, float *out,
                                                      const fftw_b200_comm *comm, unsigned flags);
fftw_b200_mpi_plan fftwf_b200_mpi_plan_transpose(ptrdiff_t n0, ptrdiff_t n1, float *in, float *out,
                                                 const fftw_b200_comm *comm, unsigned flags);
/* One distributed transform, collective; returns when this rank's result is complete (or, with
 * fftw_b200_set_async(1), once everything is enqueued on the launch stream). */
void fftw_b200_mpi_execute(fftw_b200_mpi_plan p);
void fftw_b200_mpi_destroy_plan(fftw_b200_mpi_plan p);

/* device memory that can be shared with the other ranks of the job (CUDA IPC) */
void *fftw_b200_device_malloc(size_t bytes);
void  fftw_b200_device_free(void *p);
int   fftw_b200_ipc_export(void *devptr, unsigned char handle[64]);
/* The handle names the whole allocation devptr lives in; this is devptr's byte offset inside
 * it (add it to what fftw_b200_ipc_import returns on the other side), or -1 on failure. */
ptrdiff_t fftw_b200_ipc_offset(void *devptr);
void *fftw_b200_ipc_import(const unsigned char handle[64]);
void  fftw_b200_ipc_close(void *devptr);

#ifdef __cplusplus
}
#endif
#endif
