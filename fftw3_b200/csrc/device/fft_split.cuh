// fft_split.cuh -- strided transforms as two register-only sub-passes through an L2-resident buffer.
//
// A transform of length N = RA * RB along a strided dimension (rows `row_stride` apart, pencils
// contiguous across the rows) is split Cooley-Tukey style (dft/ct.c:34-58, decimation in time):
//
//   phase A (r0 = 0..RB-1):  rows RB*j + r0, j < RA  --FFT_RA over j-->  Y[ka][r0], times W_N^(r0 ka)
//   phase B (ka = 0..RA-1):  Y[ka][r0], r0 < RB      --FFT_RB over r0--> rows ka + RA*kb
//
// Y lives in a plan-owned buffer laid out [ka][r0][pencil] that only holds one GROUP of pencils at a
// time (a few MiB): phase B of a group runs right after its phase A, so Y never leaves the 126 MB L2
// and the array still crosses HBM once per dimension.  What is gained over the one-kernel pass
// (fft_fast.cuh, COL): a thread owns one pencil and does a whole radix-RA / radix-RB butterfly in
// registers -- no shared-memory exchange, no barrier -- and a CTA touches only RA (RB) rows, each
// with kilobytes of contiguous data, instead of N rows with 64-128 bytes each.  For a dimension
// whose rows sit in different 2 MiB pages (stride 16 MiB in 1024^3) that removes the address-
// translation bottleneck of the one-kernel pass.
//
// Reference counterpart: the twiddle codelets t1_* driven by dft/dftw-direct.c:46-56 (phase A's
// butterflies + twiddles) and the no-twiddle codelets n1_* of dft/direct.c:92-97 (phase B).
//
// The per-thread body is __host__ __device__ so that tests/emu runs the same index algebra on the CPU.
#pragma once
#include "fft_generic.cuh"

namespace b2split {
using b2::cplx;
using b2::cmul;

#define B2_SPLIT_THREADS 128      /* one pencil per thread: a CTA row visit is 2 KiB (f64) of contiguous data */

template <typename T> B2_HD cplx<T> ld_user(const cplx<T> *q)
{
#ifdef __CUDA_ARCH__
    if (sizeof(T) == 8) { double2 v = __ldcs(reinterpret_cast<const double2 *>(q)); cplx<T> r; r.x = (T)v.x; r.y = (T)v.y; return r; }
    else { float2 v = __ldcs(reinterpret_cast<const float2 *>(q)); cplx<T> r; r.x = (T)v.x; r.y = (T)v.y; return r; }
#else
    return *q;
#endif
}
template <typename T> B2_HD void st_user(cplx<T> *q, cplx<T> v)
{
#ifdef __CUDA_ARCH__
    if (sizeof(T) == 8) __stcs(reinterpret_cast<double2 *>(q), make_double2((double)v.x, (double)v.y));
    else __stcs(reinterpret_cast<float2 *>(q), make_float2((float)v.x, (float)v.y));
#else
    *q = v;
#endif
}

B2_HD int64_t split_blocks(const b2d_split_pass &p)
{
    const int64_t tiles_c = (p.nc + B2_SPLIT_THREADS - 1) / B2_SPLIT_THREADS;
    return tiles_c * (p.phase == 0 ? p.rb : p.ra) * p.nb;
}

// one thread = one pencil of one (tile, r0 | ka, batch item); tws = W^(o * k), k < R (phase 0 only)
template <typename T, int R, int PHASE>
B2_HD void split_thread(const b2d_split_pass &p, int swap, int64_t blk, int tid, const cplx<T> *tws)
{
    const int64_t tiles_c = (p.nc + B2_SPLIT_THREADS - 1) / B2_SPLIT_THREADS;
    const int64_t tc = blk % tiles_c; blk /= tiles_c;
    const int other_n = PHASE == 0 ? p.rb : p.ra;
    const int o = (int)(blk % other_n);            // phase A: r0, phase B: ka
    const int64_t b = blk / other_n;
    const int64_t c = tc * B2_SPLIT_THREADS + tid;
    if (c >= p.nc) return;
    const int64_t rs2 = p.row_stride / 2;          // complex units
    cplx<T> *user = reinterpret_cast<cplx<T> *>(swap ? p.user_im : p.user_re) + (b * p.bs) / 2 + c;
    cplx<T> *work = reinterpret_cast<cplx<T> *>(p.work) + b * p.nc + c;
    const int64_t wstride = p.nb * p.nc;           // between consecutive (ka, r0) slots
    T re[R], im[R];
    if (PHASE == 0) {
#pragma unroll
        for (int j = 0; j < R; ++j) {
            cplx<T> v = ld_user<T>(user + ((int64_t)p.rb * j + o) * rs2);
            re[j] = swap ? v.y : v.x;
            im[j] = swap ? v.x : v.y;
        }
        Butterfly<R, T>::run(re, im);
#pragma unroll
        for (int k = 0; k < R; ++k) {
            cplx<T> v; v.x = re[k]; v.y = im[k];
            if (k > 0) v = cmul(v, tws[k]);
            work[((int64_t)k * p.rb + o) * wstride] = v;
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cplx<T> v = work[((int64_t)o * p.rb + r) * wstride];
            re[r] = v.x; im[r] = v.y;
        }
        Butterfly<R, T>::run(re, im);
#pragma unroll
        for (int k = 0; k < R; ++k) {
            cplx<T> v;
            v.x = swap ? im[k] : re[k];
            v.y = swap ? re[k] : im[k];
            st_user<T>(user + ((int64_t)o + (int64_t)p.ra * k) * rs2, v);
        }
    }
}

#ifdef __CUDACC__
template <typename T, int R, int PHASE>
__global__ void __launch_bounds__(B2_SPLIT_THREADS, 2)
split_kernel(const __grid_constant__ b2d_split_pass p, int swap)
{
    __shared__ cplx<T> tws[R];
    if (PHASE == 0) {
        // twiddles W_N^(r0 * k), the same for every thread of the CTA
        const int64_t tiles_c = (p.nc + B2_SPLIT_THREADS - 1) / B2_SPLIT_THREADS;
        const int o = (int)(((int64_t)blockIdx.x / tiles_c) % p.rb);
        if (threadIdx.x < R) tws[threadIdx.x] = reinterpret_cast<const cplx<T> *>(p.tw)[(int64_t)o * threadIdx.x];
        __syncthreads();
    }
    split_thread<T, R, PHASE>(p, swap, (int64_t)blockIdx.x, (int)threadIdx.x, tws);
}
#endif

}  // namespace b2split
