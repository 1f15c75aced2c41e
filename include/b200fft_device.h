/* b200fft_device.h -- the thin C-ABI shim between the C host layer (api +
 * plan builder, fftw3_b200/csrc/host) and the sm_100a CUDA kernels
 * (fftw3_b200/csrc/device).  Plain pointers and sizes only; no CUDA or torch
 * types in any signature, so the host layer is compiled by a C compiler.
 *
 * Reference counterpart: this seam sits where the reference's plan->apply
 * closures call codelets (dft/direct.c:92-97, dft/dftw-direct.c:46-56,
 * rdft/ct-hc2c-direct.c:45-60) and kernel copies (kernel/cpy2d.c, rdft/rank0.c).
 * A "pass" below is one kernel launch that streams the array through HBM once.
 */
#ifndef B200FFT_DEVICE_H
#define B200FFT_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2D_MAX_STAGES 12
#define B2D_MAX_BATCH_DIMS 3
#define B2D_MAX_PEERS 16

enum { B2D_F64 = 0, B2D_F32 = 1 };

/* element-wise operations fused into the load side of an FFT pass (bit mask) */
enum {
    B2D_LOAD_REAL = 1,      /* input is real (imag := 0), in_im unused               */
    B2D_LOAD_HERMCONJ = 2,  /* loads conj(H), H = Hermitian sequence of logical length
                               n_in given by its non-redundant half (c2r): k > n_in/2
                               reads in[n_in-k]; imag of self-mirrored bins ignored   */
    B2D_LOAD_PAD = 4,       /* k >= n_in reads as zero                               */
    B2D_LOAD_CHIRP = 8,     /* multiply by aux0[k] (Bluestein chirp)                 */
    B2D_LOAD_RADER = 32,    /* element k comes from input index perm[k] (aux0: int32 perm_in[n], perm_out[n]);
                               Rader's prime-size algorithm, bluestein == 2            */
    B2D_LOAD_R2R = 16,      /* real line of n_in elements -> complex work sequence of length n by
                               the PRE map of r2r_kind (aux0 = quarter-wave table)     */
    B2D_LOAD_C2R_MERGE = 64 /* first pass of the half-size transform of an even-size c2r (rdft/ct-hc2c.c:146-273 run
                               backwards): the input is the Hermitian half X[0..m], m = n_in, and logical element j
                               loads as swap(Z_j), Z_j = (X_j + conj X_{m-j}) + i conj(w)^j (X_j - conj X_{m-j}),
                               w = exp(-2 pi i / 2m) from aux2 -- the merge rides on the load, no pass of its own */
};

/* element-wise operations fused into the store side of an FFT pass (bit mask) */
enum {
    B2D_STORE_REALPART = 1,    /* store Re only (c2r)                                 */
    B2D_STORE_TRUNC = 2,       /* store only k < n_out                                */
    B2D_STORE_CHIRP_SCALE = 4, /* Bluestein: z * aux0[k] * scale                      */
    B2D_STORE_TWIDDLE4 = 8,    /* four-step: z *= W_big^(k * b0) (two-level tables)   */
    B2D_STORE_RADER = 32,      /* element k goes to output index perm_out[k], times scale */
    B2D_STORE_R2R = 16,        /* output k scattered into the real line of n_out elements by the
                                  POST map of r2r_kind                                 */
    B2D_STORE_R2C_SPLIT = 64   /* second pass of a four-step half-size transform of an even-size r2c: the split
                                  X_k = 1/2[(Z_k + conj Z_{m-k}) - i w^k (Z_k - conj Z_{m-k})] rides on the store
                                  (aux0 = exp(-2 pi i q / n), q <= n/2; specialised kernels only)      */
};

/* One batched strided 1-D complex FFT pass.  All strides/offsets are in units
 * of the REAL scalar type (like the reference's internal tensors, where a
 * complex interleaved array has stride 2: api/plan-many-dft.c:43-46). */
typedef struct b2d_fft_pass {
    int prec;                 /* B2D_F64 / B2D_F32                                  */
    int n;                    /* transform length computed in shared memory          */
    int nstages;
    int radix[B2D_MAX_STAGES];
    int tpb;                  /* transforms per CTA (tile along batch dim 0)         */
    int tpx;                  /* threads per transform                               */
    int load_col, store_col;  /* 0: consecutive lanes walk the transform (ROW);
                                 1: consecutive lanes walk batch dim 0 (COL)         */
    int pre_op, post_op;
    int bluestein;            /* 1: forward stages, x aux1[k], inverse stages        */
    int cache;                /* bit 0: input is expected in L2 (plain loads), bit 1: keep
                                 the output in L2 (plain stores); else streaming hints  */
    int kernel;               /* 0: generic runtime-radix kernel; else code of a
                                 specialised kernel: tile width + 1000 for COL      */
    int n_in, n_out;          /* valid input / stored output length (pad, truncate)  */
    /* idx_mul != 0: this pass is one half of a four-step transform of a longer line, and the LOGICAL
       index of its element k in the line is k * idx_mul + b0 (b0 = batch dim 0 index, which walks the
       line with stride bis[0] / bos[0]).  LOAD_HERMCONJ mirrors and STORE_TRUNC cuts on that index,
       n_in / n_out then being lengths of the whole line.                                             */
    int64_t idx_mul;
    /* LOAD_R2R / STORE_R2R, kinds whose PRE sequence is real: r2r_pair != 0 packs TWO lines into one
       complex transform (line A -> real part, line B = A + pair_is -> imaginary part; the spectra are
       separated as (Z_k +- conj Z_{n-k}) / 2 before the POST map).  Batch dim 0 then counts pairs. */
    int r2r_pair;
    int64_t pair_is, pair_os;
    int r2r_kind;             /* LOAD_R2R / STORE_R2R: 0..10 = R2HC HC2R DHT REDFT00 REDFT01 REDFT10 REDFT11
                                 RODFT00 RODFT01 RODFT10 RODFT11 (the fftw_r2r_kind values)          */
    int64_t is, os;           /* element stride along the transform                  */
    int64_t bn[B2D_MAX_BATCH_DIMS], bis[B2D_MAX_BATCH_DIMS], bos[B2D_MAX_BATCH_DIMS];
    /* buffers: re/im pointers (im == re +- 1 means interleaved; sign swap = -1)    */
    const void *in_re, *in_im;
    void *out_re, *out_im;
    const void *tw;           /* n complex: exp(-2 pi i k / n)                       */
    const void *aux0, *aux1;  /* op tables                                           */
    const void *aux2;         /* LOAD_C2R_MERGE: exp(-2 pi i q / 2m), q < m           */
    int64_t aux_split;        /* TWIDDLE4: lo-table length L (e = hi*L + lo)          */
    int64_t big_n;            /* TWIDDLE4: N of the enclosing transform               */
    int tw4_shift;            /* log2(aux_split) when big_n and aux_split are powers of two, else -1 */
    int64_t tw4_off;          /* TWIDDLE4: batch dim 0 index b0 stands for global column b0 + tw4_off (a rank's block
                                 of the columns of a distributed six-step 1-D transform) */
    double scale;
    /* peer scatter (multi-GPU exchange fused into the pass): when npeer > 0 batch
       dim 2 does not stride the output but selects peer_out[b2], an interleaved
       complex buffer that may live on another GPU (CUDA IPC / NVLink peer mapping) */
    int npeer;
    void *peer_out[B2D_MAX_PEERS];
    /* peer split by output row (second exchange fused into the last pass): when npeer > 0
       and peer_rows > 0, output row k of every transform goes to peer_out[k / peer_rows] at
       row k % peer_rows, and batch dim 2 strides the output as usual */
    int peer_rows;
    /* peer scatter (peer_rows == 0): consecutive CTAs walk the destinations first, starting
       at peer (peer_rot % npeer), so that every GPU feeds all its peers at once and no
       destination is hit by all senders at the same time */
    int peer_rot;
    /* 0: one CTA per tile.  > 0: launch at most this many CTAs, which loop over the
       tiles -- used to keep an NVLink-bound pass from occupying every SM */
    int grid_limit;
    int tw_smem;              /* generic kernel: stage the n-entry twiddle table in shared memory */
} b2d_fft_pass;

/* One half of a strided transform of length ra * rb done as two register-only sub-passes through an
 * L2-resident work buffer (device/fft_split.cuh).  Pencils are contiguous interleaved complex numbers:
 * `nc` of them per batch item, `nb` batch items `bs` reals apart; consecutive transform indices are
 * `row_stride` reals apart.  work = [ka][r0][b][c] interleaved complex.
 *   phase 0: user rows rb*j + r0 (j < ra) -> FFT_ra -> times W^(r0 ka) -> work[ka][r0]
 *   phase 1: work[ka][r0] (r0 < rb) -> FFT_rb -> user rows ka + ra*kb                              */
typedef struct b2d_split_pass {
    int prec;
    int ra, rb;
    int phase;
    int64_t row_stride;
    int64_t nc, nb, bs;
    void *user_re, *user_im;  /* phase 0 reads them, phase 1 writes them; im = re +- 1 (interleaved) */
    void *work;
    const void *tw;           /* ra * rb entries exp(-2 pi i k / (ra rb)) */
} b2d_split_pass;

/* strided N-d copy / rank-0 transform (kernel/cpy2d.c, rdft/rank0.c analogue);
 * element = `elem_reals` consecutive reals (1: real scalar, 2: interleaved complex) */
typedef struct b2d_copy {
    int prec;
    int elem_reals;
    int rank;                       /* <= 4 */
    int64_t n[4], is[4], os[4];     /* dim 0 fastest                                */
    const void *in;
    void *out;
    /* peer gather: when npeer > 0 dim 3 selects the SOURCE buffer peer_in[i3] */
    int npeer;
    const void *peer_in[B2D_MAX_PEERS];
    int grid_limit;                 /* as in b2d_fft_pass */
} b2d_copy;

/* r2c / c2r even-length split (rdft/ct-hc2c-direct.c:45-60 analogue) and r2r
 * pre/post element maps: see device/real_ops.cuh */
typedef struct b2d_realop {
    int prec;
    int op;                    /* B2D_ROP_* */
    int n;                     /* logical 1-D size                                   */
    int m;                     /* complex work length                                */
    int64_t xs;                /* stride of the user-side line (reals)               */
    int64_t bn[B2D_MAX_BATCH_DIMS], bxs[B2D_MAX_BATCH_DIMS];  /* user-side batch     */
    int64_t wdist;             /* work-buffer distance between lines (complex units) */
    const void *x_re, *x_im;   /* user side (x_im used by split complex ops)          */
    void *y_re, *y_im;
    void *work;                /* interleaved complex work buffer                     */
    const void *tw;            /* op-specific table                                   */
    const void *aux;           /* BLUE_MID: FFT of the Bluestein filter (m complex)   */
    int flags;                 /* BLUE_PRE: B2D_LOAD_* ; BLUE_POST: B2D_STORE_*       */
    int n_lim;                 /* BLUE_PRE: valid inputs ; BLUE_POST: outputs stored   */
    double scale;              /* BLUE_POST: 1 / m                                    */
} b2d_realop;

enum {
    B2D_ROP_R2C_POST = 1,   /* work[0..n/2) = FFT(z), write X[0..n/2] to user (cr,ci)  */
    B2D_ROP_C2R_PRE = 2,    /* user X[0..n/2] -> work z-spectrum for backward FFT     */
    /* Bluestein for sizes whose padded length m does not fit one CTA (dft/bluestein.c:82-128):
       PRE  work[i] = i < n ? load(x_i) * chirp_i : 0          (i < m)
       MID  work[i] = conj(work[i] * B_i)                      (between the two FFT_m)
       POST y_k = conj(work[k]) * chirp_k * scale              (k < n_lim)                  */
    B2D_ROP_BLUE_PRE = 3, B2D_ROP_BLUE_MID = 4, B2D_ROP_BLUE_POST = 5,
    B2D_ROP_R2R_PRE = 16,   /* + kind: user real line -> complex work sequence        */
    B2D_ROP_R2R_POST = 32   /* + kind: complex work spectrum -> user real line        */
};

/* ---- runtime ---- */
int  b2d_device_count(void);            /* 0 if no usable CUDA device                  */
const char *b2d_device_name(void);
int  b2d_sm_count(void);
const char *b2d_last_error(void);
int  b2d_current_device(void);            /* ordinal of the calling thread's current CUDA device, -1 without one */
int  b2d_pointer_is_device(const void *p);   /* 1 device/managed, 0 host, -1 error     */
void *b2d_malloc(size_t bytes);
void b2d_free(void *p);
void *b2d_malloc_host(size_t bytes);     /* pinned */
void b2d_free_host(void *p);
int  b2d_memcpy_h2d(void *dst, const void *src, size_t bytes);
int  b2d_memcpy_d2h(void *dst, const void *src, size_t bytes);
int  b2d_memcpy_d2h_async(void *dst, const void *src, size_t bytes);   /* enqueue only: the caller synchronises the stream */
int  b2d_memcpy_d2d(void *dst, const void *src, size_t bytes);
/* `height` rows of `width` bytes between device (or peer-mapped) arrays, enqueued on `stream` for the copy engines */
int  b2d_memcpy2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, void *stream);
int  b2d_memset(void *dst, int byte, size_t bytes);
int  b2d_sync(void);
void b2d_set_stream(void *cuda_stream);  /* NULL = legacy default stream              */
void *b2d_get_stream(void);             /* the stream launches of THIS thread go to          */
/* per-thread override of the launch stream (nested): returns the previous value for b2d_pop_stream */
void *b2d_push_stream(void *cuda_stream);
void b2d_pop_stream(void *prev);
/* side streams for overlapping NVLink-bound kernels with HBM-bound ones */
void *b2d_aux_stream(int idx);                           /* lazily created, non-blocking  */
void *b2d_pipe_stream(int idx);                          /* 3 streams of the host-array pipeline (exec.c); NULL = none */
/* Two streams bound to disjoint sets of SMs of the current device (CUDA green contexts): `comm_sms` SMs (a multiple of
 * 8) for an NVLink-bound kernel and the remaining SMs for the HBM-bound kernel that runs next to it, so that neither
 * waits behind the other inside an SM.  Created once per (device, comm_sms); 0 on success, -1 when the driver cannot
 * partition (the caller then shares the SMs as before). */
int  b2d_partition_streams(int comm_sms, void **comm_stream, void **compute_stream);
int  b2d_stream_wait_stream(void *waiter, void *signaler); /* event edge signaler -> waiter */
size_t b2d_max_smem_per_block(void);

/* CUDA IPC: share a b2d_malloc'ed buffer with the other single-GPU processes of a job */
int  b2d_ipc_export(void *devptr, unsigned char handle[64]);
void *b2d_ipc_import(const unsigned char handle[64]);
void b2d_ipc_close(void *devptr);
int64_t b2d_alloc_offset(const void *devptr);   /* byte offset inside its allocation, -1 on failure */

/* Barrier among the single-GPU processes of a job, on the launch stream, without the host: flags[d] is rank
   d's flag array (nranks x 8 bytes, zeroed, peer-mapped); the kernel stores `epoch` into slot `rank` of every
   peer's array (system-scope release) and spins until all slots of its own array have reached `epoch`.
   Everything enqueued before it on this stream -- remote stores of pass kernels included -- is complete and
   visible to the peers when they leave the barrier. */
int  b2d_peer_barrier(void *const *flags, int rank, int nranks, unsigned long long epoch);

/* timing on the launch stream (planner measurements) */
int  b2d_timer_start(void);
int  b2d_timer_stop(float *ms);

/* ---- kernels ---- */
int  b2d_launch_fft_pass(const b2d_fft_pass *p);
size_t b2d_fft_pass_smem(const b2d_fft_pass *p);  /* dynamic smem bytes it needs   */
int  b2d_fast_available(const b2d_fft_pass *p, int code); /* specialised kernel exists for this shape? */
int  b2d_launch_split_pass(const b2d_split_pass *p);
int  b2d_split_supported(int prec, int ra, int rb);      /* register-only sub-pass kernels exist for this split? */
int  b2d_launch_copy(const b2d_copy *c);
int  b2d_launch_realop(const b2d_realop *r);

/* number of kernels launched since the library was loaded (bench `gpu_launches`) */
uint64_t b2d_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200FFT_DEVICE_H */
