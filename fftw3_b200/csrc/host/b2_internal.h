/* b2_internal.h -- host-side (plain C) data model of the B200 FFT engine.
 *
 * problem  : canonical description of what the user asked for (the reference's
 *            problem_dft / problem_rdft2 / problem_rdft: dft/dft.h:34-38,
 *            rdft/rdft.h:33-44,97-124).  Strides are in units of the real
 *            scalar type, complex pointers are (re, im) pairs; a BACKWARD
 *            complex transform is the forward one with re/im swapped
 *            (kernel/extract-reim.c:27-36).
 * plan     : a flat list of device passes ("steps"), each one kernel launch,
 *            with symbolic buffer bindings so the same plan can run on new
 *            arrays (api/execute-dft.c:25-32).  This replaces the reference's
 *            tree of solver closures (kernel/ifftw.h:587-591).
 */
#ifndef B2_INTERNAL_H
#define B2_INTERNAL_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include "../../../include/b200fft_device.h"

#define B2_MAXRANK 16
#define B2_RNK_MINFTY (-1)          /* "rank minus infinity": empty tensor (howmany 0) */

typedef struct { int64_t n, is, os; } b2_dim;
typedef struct { int rnk; b2_dim d[B2_MAXRANK]; } b2_tensor;

typedef enum { B2_C2C = 0, B2_R2C = 1, B2_C2R = 2, B2_R2R = 3 } b2_kind;

/* public flag bits we interpret (values fixed by the fftw3.h ABI) */
#define B2F_MEASURE 0u
#define B2F_DESTROY_INPUT (1u << 0)
#define B2F_UNALIGNED (1u << 1)
#define B2F_CONSERVE_MEMORY (1u << 2)
#define B2F_EXHAUSTIVE (1u << 3)
#define B2F_PRESERVE_INPUT (1u << 4)
#define B2F_PATIENT (1u << 5)
#define B2F_ESTIMATE (1u << 6)
#define B2F_WISDOM_ONLY (1u << 21)

typedef struct {
    int prec;                 /* B2D_F64 / B2D_F32 */
    b2_kind kind;
    b2_tensor sz;             /* transform dims (row-major order as given)          */
    b2_tensor vecsz;          /* batch dims                                          */
    /* user pointers at plan time.
       c2c: in0=ri in1=ii out0=ro out1=io
       r2c: in0=r           out0=cr out1=ci   (sz.is: real strides, sz.os: complex)
       c2r: in0=cr in1=ci   out0=r
       r2r: in0=in          out0=out                                                  */
    void *in0, *in1, *out0, *out1;
    int r2r_kind[B2_MAXRANK]; /* public fftw_r2r_kind values                          */
    unsigned flags;
    /* internal (distributed six-step, dist_api.c): rank-1 c2c whose output k of the batch column b0 (the
       contiguous batch dimension) is multiplied by exp(-2 pi i k (b0 + tw_off) / tw_big_n); 0 = none */
    int64_t tw_big_n, tw_off;
} b2_problem;

/* ---- plan ---- */
enum { BUF_NONE = 0, BUF_IN0, BUF_IN1, BUF_OUT0, BUF_OUT1, BUF_SCRATCH0, BUF_SCRATCH1, BUF_SCRATCH2, BUF_SCRATCH3, BUF_SCRATCH4, BUF_SCRATCH5, BUF_TABLE, BUF_COUNT };
#define B2_NSCRATCH 6

typedef struct { int buf; int64_t off; /* in reals (scratch/table: in bytes) */ } b2_ref;

typedef enum { STEP_FFT = 1, STEP_COPY, STEP_REALOP, STEP_SPLIT } b2_step_kind;

typedef struct {
    b2_step_kind kind;
    union { b2d_fft_pass fft; b2d_copy copy; b2d_realop rop; b2d_split_pass split; } u;
    b2_ref r[6];              /* fft: in_re,in_im,out_re,out_im ; copy: in,out ;
                                 realop: x_re,x_im,y_re,y_im,work ; split: user_re,user_im,-,-,work */
    char note[48];            /* for print_plan                                       */
    int lane;                 /* 0: caller's stream; k > 0: side stream k - 1 (exec.c: run_steps) */
    int fence;                /* all lanes are joined before this step: it starts a pass that reads what
                                 the previous pass wrote through other lanes                            */
} b2_step;

typedef struct b2_table {     /* device-resident constant table, refcounted & shared */
    struct b2_table *next;
    int prec, kind;
    int64_t n, aux;
    int device;               /* CUDA device the table lives on: a process that switches GPUs gets one per device */
    void *dev;
    size_t bytes;
    int refs;
} b2_table;

/* decomposition choices above the single pass (what the reference's planner explores as alternative
   solver trees, kernel/planner.c:518-615): measured as whole plans by FFTW_MEASURE and remembered in
   wisdom under the problem's signature.  Environment variables (FFTW3_B200_L2_*, FFTW3_B200_SPLIT*)
   pin them for experiments and tests. */
typedef struct {
    size_t l2_block_bytes;    /* L2-resident pass pairs: group size, 0 = off                    */
    int l2_lanes, l2_keep;    /* lanes the groups rotate over; cache hint of the first pass     */
    int l2_pair_outer;        /* pair the last dim with dim 0 instead of the next-to-last       */
    int split_mode;           /* strided pass as two register-only sub-passes: 0 never, 1 when rows are >= 1 MiB apart, 2 whenever applicable */
    size_t split_bytes;       /* its group size                                                  */
    int split_lanes;
    int real_unfused;         /* bit 0: even-size r2c keeps its split as a pass of its own, bit 1: c2r its merge   */
    int r2r_transposes;       /* dense 2-d r2r with long columns: transposes around the column pass instead of two
                                 line passes with transposed stores                                                */
    int prime_mode;           /* prime sizes: 0 = rule (register-resident Bluestein kernel if any, else Rader),
                                 1 = Rader, 2 = Bluestein                                                          */
} b2_plan_opts;

typedef struct b2_plan {
    int refcnt;
    b2_problem prob;
    int nsteps, cap;
    b2_step *steps;
    size_t scratch_bytes[B2_NSCRATCH];
    void *scratch[B2_NSCRATCH];
    int ntables, tcap;
    b2_table **tables;
    double est_flops_add, est_flops_mul, est_flops_fma;
    double cost;              /* measured ms (or estimate) */
    int is_nop;
    b2_plan_opts opt;
    int alt;                  /* which whole-plan alternative this is (print_plan)               */
    int no_fence;             /* planner state: passes emitted now are independent of the previous one (L2 groups) */
    int inplace;
    int destroys_input;
    /* host staging state (execute on host pointers) */
    void *stage_dev[4];
    size_t stage_cap[4];
    /* host arrays, batched problem: the outermost batch dimension is cut into pipe_chunks chunks that run through
       three copies of a chunk-sized plan on three streams, so that the upload of chunk c + 1, the passes of chunk c
       and the download of chunk c - 1 overlap (PCIe is full duplex); offsets per chunk in reals */
    struct b2_plan *pipe[3];
    int pipe_chunks;
    int64_t pipe_in_off, pipe_out_off;
    void *lock;               /* pthread_mutex_t* */
} b2_plan;

/* exec.c: first / last real reached through user pointer `which` (0, 1 = in; 2, 3 = out) */
void b2_problem_span(const b2_problem *q, int which, int64_t *lo, int64_t *hi);

/* tensor.c */
void b2_tensor_init(b2_tensor *t, int rnk);
int64_t b2_tensor_count(const b2_tensor *t);          /* product of n (1 for rank 0, 0 for minfty) */
void b2_tensor_drop_unit(b2_tensor *t);               /* remove n == 1 dims                          */
void b2_tensor_append(b2_tensor *t, const b2_tensor *a);
void b2_tensor_sort_merge(b2_tensor *t);              /* ascending |os|, merge dims contiguous in both is and os */
void b2_tensor_span(const b2_tensor *t, int use_os, int64_t *lo, int64_t *hi); /* min/max offset reached */
int  b2_tensor_inplace_ok(const b2_tensor *t);        /* all is == os */

/* tables.c : accurate constant tables built in long double on the host */
enum { TAB_TWIDDLE = 1,      /* n entries exp(-2 pi i k / n)                         */
       TAB_CHIRP,            /* n entries exp(-pi i k^2 / n)                         */
       TAB_BLUE_B,           /* aux = M: FFT_M of the Bluestein filter               */
       TAB_R2C,              /* n/2+1 entries exp(-2 pi i k / n)  (half-size split)  */
       TAB_TW4_LO, TAB_TW4_HI, /* two-level tables for exp(-2 pi i e / n), aux = L  */
       TAB_QUARTER,          /* 2n entries: exp(-pi i k/(2n)) then exp(-pi i (2k+1)/(4n)) */
       TAB_RADER_PERM,       /* prime n: int32 perm_in[q] = g^q mod n, then perm_out[m] = g^-m mod n (n-1 each) */
       TAB_RADER_B,          /* prime n: FFT_{n-1} of b_q = exp(-2 pi i g^-q / n)        */
       TAB_COUNT };
b2_table *b2_table_get(int prec, int kind, int64_t n, int64_t aux);
void b2_table_release(b2_table *t);
void b2_tables_cleanup(void);
void b2_unit_root_ld(int64_t m, int64_t n, long double *c, long double *s); /* exp(-2 pi i m/n) */
int  b2_is_prime(int64_t n);
int64_t b2_primitive_root(int64_t p);     /* smallest generator of (Z/p)^*, p prime (kernel/primes.c:81-122 role) */

/* planner.c */
/* One process-wide recursive lock around everything that touches planner state (plan creation and
   destruction, the shared table list, wisdom): the planner entry points may then be called from any
   thread, which is what fftw_make_planner_thread_safe() asks for (doc/threads.texi:225-270,
   api/apiplan.c:23-29).  Execution never takes it. */
void b2_planner_lock(void);
void b2_planner_unlock(void);
b2_plan *b2_mkplan(const b2_problem *prob);
void b2_plan_destroy(b2_plan *p);
void b2_plan_print(const b2_plan *p, FILE *f);
extern double b2_timelimit;

/* choose radix factorisation of n into supported radices; returns #stages or 0 */
int b2_factorize(int64_t n, int prec, int variant, int *radix);
int64_t b2_max_single_pass(int prec);

/* wisdom.c */
typedef struct { uint64_t h[2]; } b2_sig;
b2_sig b2_sig_of_pass(const b2d_fft_pass *p, int inplace);
b2_sig b2_sig_of_problem(const b2_problem *q, int inplace);
int  b2_wisdom_lookup(b2_sig s, unsigned patience, int *variant);
void b2_wisdom_store(b2_sig s, unsigned patience, int variant);
void b2_wisdom_forget(void);
void b2_wisdom_export(void (*emit)(char c, void *), void *data, int prec);
int  b2_wisdom_import(int (*next)(void *), void *data, int prec);

/* exec.c */
void b2_execute(b2_plan *p, void *in0, void *in1, void *out0, void *out1);
void b2_execute_ex(b2_plan *p, void *in0, void *in1, void *out0, void *out1, int nosync);
void b2_plan_lock_init(b2_plan *p);
void b2_plan_lock_destroy(b2_plan *p);
void b2_wisdom_set_prec(int prec);
extern int b2_async_mode;

#endif
