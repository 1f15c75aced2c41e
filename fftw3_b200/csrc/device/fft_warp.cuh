// fft_warp.cuh -- one warp per transform: contiguous 1024-point transforms as two radix-32 stages.
//
// Lane j loads x[j + 32 r] (r < 32: every load instruction of the warp is 32 consecutive complex numbers),
// does a whole radix-32 butterfly in registers, the warp exchanges through its PRIVATE 16.5 KB of shared
// memory (one __syncwarp, no block barrier), lane j multiplies its 32 inputs of the second stage by
// W_1024^(r j) (table laid out [r][j] in shared memory, loaded once per CTA) and the second radix-32
// butterfly's outputs X[j + 32 r] go straight to HBM, again 32 consecutive numbers per instruction.
//
// Compared with the block-cooperative kernel (fft_fast.cuh: 16 x 16 x 4, two exchanges, three block
// barriers, twiddle derivations): one exchange instead of two (-33 % shared-memory wavefronts), ~10 % fewer
// FP64 instructions, and warps that never wait for each other, so loads, butterflies and stores of
// different transforms overlap freely inside an SM.  CTAs are persistent (a warp loops over transforms).
//
// Reference counterpart: n1_32 + t1_32 codelets for N = 1024 = 32 x 32 (dft/ct.c:34-58,
// dft/dftw-direct.c:46-56), which is exactly how the reference decomposes config C1.
#pragma once
#include <cuda_runtime.h>
#include "fft_fast.cuh"

namespace b2warp {
using b2::cplx;
using b2::cmul;

constexpr int WARPS = 4;
constexpr int N = 1024, R = 32;
constexpr int PITCH = N + N / R;            // one pad element per 32: positions 33 j + r

template <typename T>
__global__ void __launch_bounds__(WARPS * 32, 2)
warp1024_kernel(const __grid_constant__ b2d_fft_pass p, int swap_in, int swap_out, long long ntrans)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T> *tw2 = reinterpret_cast<cplx<T> *>(smem_raw);                    // [r][j], 1024 entries
    cplx<T> *buf = tw2 + N + (threadIdx.x / 32) * PITCH;                     // this warp's exchange area
    const int lane = threadIdx.x & 31;
    const cplx<T> *tw = reinterpret_cast<const cplx<T> *>(p.tw);
    for (int i = threadIdx.x; i < N; i += WARPS * 32) tw2[i] = b2fast::ldg_c(&tw[(i / R) * (i % R)]);
    __syncthreads();
    const cplx<T> *gin_base = reinterpret_cast<const cplx<T> *>(swap_in ? p.in_im : p.in_re);
    cplx<T> *gout_base = reinterpret_cast<cplx<T> *>(swap_out ? p.out_im : p.out_re);
    const unsigned long long keep_pol = (p.cache & 4) ? b2fast::policy_evict_last() : 0ull;
    const long long wstep = (long long)gridDim.x * WARPS;
    for (long long tr = (long long)blockIdx.x * WARPS + threadIdx.x / 32; tr < ntrans; tr += wstep) {
        const int64_t b0 = tr % p.bn[0];
        const int64_t rest = tr / p.bn[0];
        const int64_t b1 = rest % p.bn[1], b2i = rest / p.bn[1];
        const cplx<T> *gin = gin_base + (b0 * p.bis[0] + b1 * p.bis[1] + b2i * p.bis[2]) / 2 + lane;
        cplx<T> *gout = gout_base + (b0 * p.bos[0] + b1 * p.bos[1] + b2i * p.bos[2]) / 2 + lane;
        T re[R], im[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const cplx<T> v = b2fast::ld_stream(gin + 32 * r);
            re[r] = swap_in ? v.y : v.x;
            im[r] = swap_in ? v.x : v.y;
        }
        Butterfly<R, T>::run(re, im);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cplx<T> v; v.x = re[r]; v.y = im[r];
            buf[33 * lane + r] = v;
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cplx<T> v = buf[33 * r + lane];
            if (r > 0) v = cmul(v, tw2[r * R + lane]);
            re[r] = v.x; im[r] = v.y;
        }
        __syncwarp();                              // the exchange area is free for the next transform
        Butterfly<R, T>::run(re, im);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cplx<T> o;
            o.x = swap_out ? im[r] : re[r];
            o.y = swap_out ? re[r] : im[r];
            b2fast::st_out<T>(gout + 32 * r, o, p.cache, keep_pol);
        }
    }
}

template <typename T> constexpr size_t smem_bytes() { return (size_t)(N + WARPS * PITCH) * sizeof(cplx<T>); }

void init(int max_smem, int sms);
int applicable(const b2d_fft_pass &p);             // structural: contiguous 1024-point lines, no fused ops
int launch(const b2d_fft_pass &p, cudaStream_t st);   // 0 launched, 1 not applicable at run time, -1 error

}  // namespace b2warp
