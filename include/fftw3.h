/* fftw3.h -- public C API of fftw3_b200, a B200-native (sm_100a) FFT engine
 * that is source- and ABI-compatible with FFTW 3's transform-execution API.
 *
 * This header is written for this project; it declares the same types,
 * constants and functions as the reference's api/fftw3.h (FFTW 3.3.11) so that
 * code written against FFTW compiles and links unchanged:
 *   - double precision  : fftw_*   (reference: libfftw3)
 *   - single precision  : fftwf_*  (reference: libfftw3f)
 * long double / quad precision are not offered (no such arithmetic on the GPU).
 *
 * Layout of the declarations: include/fftw3_api.inc holds one precision's API
 * and is included once per precision below.
 */
#ifndef FFTW3_H
#define FFTW3_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Complex numbers: C99 `R _Complex` when <complex.h> was included first,
 * otherwise the bit-compatible R[2] (reference: api/fftw3.h:57-63). */
#if !defined(FFTW_NO_Complex) && defined(_Complex_I) && defined(complex) && defined(I)
#define FFTW3_COMPLEX_TYPEDEF(R, C) typedef R _Complex C
#else
#define FFTW3_COMPLEX_TYPEDEF(R, C) typedef R C[2]
#endif

/* kinds of real-to-real transforms; numeric values are ABI (api/fftw3.h:96-100) */
enum fftw_r2r_kind_do_not_use_me {
    FFTW_R2HC = 0,
    FFTW_HC2R = 1,
    FFTW_DHT = 2,
    FFTW_REDFT00 = 3,
    FFTW_REDFT01 = 4,
    FFTW_REDFT10 = 5,
    FFTW_REDFT11 = 6,
    FFTW_RODFT00 = 7,
    FFTW_RODFT01 = 8,
    FFTW_RODFT10 = 9,
    FFTW_RODFT11 = 10
};

/* guru dimension descriptors (api/fftw3.h:102-113) */
struct fftw_iodim_do_not_use_me { int n, is, os; };
struct fftw_iodim64_do_not_use_me { ptrdiff_t n, is, os; };

typedef void (*fftw_write_char_func_do_not_use_me)(char c, void *);
typedef int (*fftw_read_char_func_do_not_use_me)(void *);

/* double precision */
#define FFTW3_NS(name) fftw_##name
#define FFTW3_REAL double
#define FFTW3_CPLX fftw_complex
#include "fftw3_api.inc"
#undef FFTW3_NS
#undef FFTW3_REAL
#undef FFTW3_CPLX

/* single precision */
#define FFTW3_NS(name) fftwf_##name
#define FFTW3_REAL float
#define FFTW3_CPLX fftwf_complex
#include "fftw3_api.inc"
#undef FFTW3_NS
#undef FFTW3_REAL
#undef FFTW3_CPLX

/* transform direction */
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)

#define FFTW_NO_TIMELIMIT (-1.0)

/* planner flags (bit values are ABI: api/fftw3.h:495-519) */
#define FFTW_MEASURE (0U)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_UNALIGNED (1U << 1)
#define FFTW_CONSERVE_MEMORY (1U << 2)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PRESERVE_INPUT (1U << 4)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_WISDOM_ONLY (1U << 21)
/* accepted and ignored (they steer CPU solver families that do not exist here) */
#define FFTW_ESTIMATE_PATIENT (1U << 7)
#define FFTW_BELIEVE_PCOST (1U << 8)
#define FFTW_NO_DFT_R2HC (1U << 9)
#define FFTW_NO_NONTHREADED (1U << 10)
#define FFTW_NO_BUFFERING (1U << 11)
#define FFTW_NO_INDIRECT_OP (1U << 12)
#define FFTW_ALLOW_LARGE_GENERIC (1U << 13)
#define FFTW_NO_RANK_SPLITS (1U << 14)
#define FFTW_NO_VRANK_SPLITS (1U << 15)
#define FFTW_NO_VRECURSE (1U << 16)
#define FFTW_NO_SIMD (1U << 17)
#define FFTW_NO_SLOW (1U << 18)
#define FFTW_NO_FIXED_RADIX_LARGE_N (1U << 19)
#define FFTW_ALLOW_PRUNING (1U << 20)

/* ---- B200 extensions (not in FFTW) -------------------------------------- */
/* Run device-pointer executes on this CUDA stream (NULL = legacy default). */
void fftw_b200_set_stream(void *cuda_stream);
/* 1: fftw_execute* on device pointers returns after enqueueing (no host sync);
 * 0 (default): FFTW's synchronous contract. Host-pointer executes always sync. */
void fftw_b200_set_async(int enabled);
void fftw_b200_synchronize(void);
/* kernels launched by this library so far; name of the device in use */
unsigned long long fftw_b200_launch_count(void);
const char *fftw_b200_device_name(void);

#ifdef __cplusplus
}
#endif
#endif /* FFTW3_H */
