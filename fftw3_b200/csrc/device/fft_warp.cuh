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

// Strided (COL) counterpart: a CTA of 32 * TP threads owns TP adjacent pencils; thread (p, j) = (tid % TP, tid / TP)
// holds rows j + 32 r of pencil p, so every global access of a warp is 32 / TP row segments of TP * 16 bytes (as
// in the block-cooperative COL kernel), but the transform is 32 x 32: ONE exchange through shared memory and one
// block barrier per tile instead of two exchanges and three barriers, and 32 independent loads in flight per thread.
constexpr int CPITCH = PITCH + 2;           // pencil pitch: (CPITCH * 16 B / 4) mod 32 = 8 banks apart

template <typename T, int TP>
__global__ void __launch_bounds__(32 * TP, TP == 4 ? 2 : 1)
col1024_kernel(const __grid_constant__ b2d_fft_pass p, int swap_in, int swap_out, long long ntiles)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T> *tw2 = reinterpret_cast<cplx<T> *>(smem_raw);                    // [r][j]
    const int pp = threadIdx.x % TP, j = threadIdx.x / TP;
    cplx<T> *buf = tw2 + N + pp * CPITCH;
    const cplx<T> *tw = reinterpret_cast<const cplx<T> *>(p.tw);
    for (int i = threadIdx.x; i < N; i += 32 * TP) tw2[i] = b2fast::ldg_c(&tw[(i / R) * (i % R)]);
    __syncthreads();
    const cplx<T> *gin_base = reinterpret_cast<const cplx<T> *>(swap_in ? p.in_im : p.in_re);
    cplx<T> *gout_base = reinterpret_cast<cplx<T> *>(swap_out ? p.out_im : p.out_re);
    const int64_t is2 = p.is / 2, os2 = p.os / 2;
    const unsigned long long keep_pol = (p.cache & 4) ? b2fast::policy_evict_last() : 0ull;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const b2::TileCtx c = b2::decode_block(p, (int64_t)tile);
        const int64_t b0 = c.tile0 * TP + pp;
        const bool valid = b0 < p.bn[0];
        const cplx<T> *gin = gin_base + (b0 * p.bis[0] + c.b1 * p.bis[1] + c.b2 * p.bis[2]) / 2 + (int64_t)j * is2;
        cplx<T> *gout = gout_base + (b0 * p.bos[0] + c.b1 * p.bos[1] + c.b2 * p.bos[2]) / 2 + (int64_t)j * os2;
        T re[R], im[R];
        const int64_t istep = 32 * is2, ostep = 32 * os2;
        const bool from_l2 = (p.cache & 1) != 0;
#pragma unroll
        for (int r = 0; r < R; ++r, gin += istep) {
            cplx<T> v; v.x = T(0); v.y = T(0);
            if (valid) {
                if (from_l2) v = b2fast::ld_plain(gin);
                else if (TP * sizeof(cplx<T>) < 128) v = b2fast::ld_l2pf<256>(gin);
                else v = b2fast::ld_stream(gin);
            }
            re[r] = swap_in ? v.y : v.x;
            im[r] = swap_in ? v.x : v.y;
        }
        Butterfly<R, T>::run(re, im);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cplx<T> v; v.x = re[r]; v.y = im[r];
            buf[33 * j + r] = v;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cplx<T> v = buf[33 * r + j];
            if (r > 0) v = cmul(v, tw2[r * R + j]);
            re[r] = v.x; im[r] = v.y;
        }
        __syncthreads();                           // the exchange area is free for the next tile
        Butterfly<R, T>::run(re, im);
        if (valid) {
#pragma unroll
            for (int r = 0; r < R; ++r, gout += ostep) {
                cplx<T> o;
                o.x = swap_out ? im[r] : re[r];
                o.y = swap_out ? re[r] : im[r];
                b2fast::st_out<T>(gout, o, p.cache, keep_pol);
            }
        }
    }
}

template <typename T> constexpr size_t smem_bytes() { return (size_t)(N + WARPS * PITCH) * sizeof(cplx<T>); }
template <typename T, int TP> constexpr size_t col_smem_bytes() { return (size_t)(N + TP * CPITCH) * sizeof(cplx<T>); }

void init(int max_smem, int sms);
// kernel codes: 3001 = warp-per-transform rows; 3104 / 3108 = strided, 4 / 8 pencils per CTA
int applicable(const b2d_fft_pass &p, int code);   // structural: 1024-point lines, no fused ops
size_t smem_for(const b2d_fft_pass &p, int code);
int launch(const b2d_fft_pass &p, cudaStream_t st);   // 0 launched, 1 not applicable at run time, -1 error

}  // namespace b2warp
