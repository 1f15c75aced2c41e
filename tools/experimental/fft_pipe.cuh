// fft_pipe.cuh -- persistent, software-pipelined Stockham kernel for strided
// (COL) passes: the tile of TPB adjacent pencils for the NEXT work item is
// brought into shared memory with asynchronous copies (cp.async, LDGSTS) while
// the CTA computes and stores the current one.
//
// Why: a thread of the register-resident kernel (fft_fast.cuh) holds 16 complex
// doubles, so the register file bounds an SM to ~8192 resident points; with one
// wide tile per SM nothing overlaps the HBM latency of the next tile, and with
// two narrow tiles the 64-byte accesses waste DRAM bursts
// (profiles/r01_ncu_full_fast_kernels_summary.txt: the strided passes sit at
// 0.59-0.67 of the measured copy bandwidth).  Here the loads of tile i+1 are in
// flight -- in shared memory, not in registers -- during all of tile i:
//   two shared buffers A/B of N*TPB points;  per tile:
//     wait own cp.async -> registers | barrier | issue cp.async for the next tile
//     into the other buffer | stage 1 | exchange through the current buffer |
//     stage 2 | exchange | stage 3 | streaming stores from registers
// Every thread copies exactly the elements it will read itself, so the async
// copies need no barrier of their own.
//
// Reference counterpart: the buffered strided solver dft/buffered.c:41-69 (copy
// a batch of pencils into a contiguous buffer, transform, copy back), with the
// copy engine of the GPU doing the buffering concurrently.
#pragma once
#include <cuda_runtime.h>
#include "fft_fast.cuh"

namespace b2pipe {
using b2::cplx;
using b2::cmul;
using b2fast::ld_stream;
using b2fast::st_stream;
using b2fast::load_twiddles;
using b2fast::unit_root;

template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <typename T, int N, int E, int R1, int R2, int TPB>
struct PipeCfg {
    static constexpr int TPX = N / E;
    static constexpr int THREADS = TPX * TPB;
    static constexpr size_t SMEM_BYTES = (size_t)2 * N * TPB * sizeof(cplx<T>);
    static_assert(E * R1 * R2 == N && R2 > 1, "three-stage sizes only");
    static_assert(E % R1 == 0 && E % R2 == 0 && TPX % E == 0, "radix layout");
};

template <typename T, int N, int E, int R1, int R2, int TPB>
__global__ void __launch_bounds__(PipeCfg<T, N, E, R1, R2, TPB>::THREADS, 1)
pipe_kernel(const __grid_constant__ b2d_fft_pass p, int swap_in, int swap_out)
{
    using Cfg = PipeCfg<T, N, E, R1, R2, TPB>;
    constexpr int TPX = Cfg::TPX;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T> *bufs = reinterpret_cast<cplx<T> *>(smem_raw);

    const int tid = threadIdx.x;
    const int t = tid % TPB;
    const int j = tid / TPB;
    const int64_t ntiles = b2::grid_blocks(p);
    const cplx<T> *gin_base = reinterpret_cast<const cplx<T> *>(swap_in ? p.in_im : p.in_re);
    cplx<T> *gout_base = reinterpret_cast<cplx<T> *>(swap_out ? p.out_im : p.out_re);
    const int64_t is2 = p.is / 2, os2 = p.os / 2;
    const cplx<T> *tw = reinterpret_cast<const cplx<T> *>(p.tw);

    auto prefetch = [&](int64_t tile, cplx<T> *buf) {
        const b2::TileCtx c = b2::decode_block(p, tile);
        const int64_t b0 = c.tile0 * TPB + t;
        if (b0 < p.bn[0]) {
            const cplx<T> *g = gin_base + (b0 * p.bis[0] + c.b1 * p.bis[1] + c.b2 * p.bis[2]) / 2;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int k = j + r * TPX;
                cp_async<(int)sizeof(cplx<T>)>(&buf[k * TPB + t], g + (int64_t)k * is2);
            }
        }
        cp_async_commit();
    };

    int cur = 0;
    if ((int64_t)blockIdx.x < ntiles) prefetch(blockIdx.x, bufs);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, cur ^= 1) {
        cplx<T> *sm = bufs + (size_t)cur * N * TPB;
        auto sidx = [&](int k) -> int { return k * TPB + t; };
        const b2::TileCtx c = b2::decode_block(p, tile);
        const int64_t b0 = c.tile0 * TPB + t;
        const bool valid = b0 < p.bn[0];
        cplx<T> *gout = gout_base + (b0 * p.bos[0] + c.b1 * p.bos[1] + c.b2 * p.bos[2]) / 2;

        T re[E], im[E];
        cp_async_wait_all();                   // this thread's own elements of `tile` have landed
#pragma unroll
        for (int r = 0; r < E; ++r) {
            cplx<T> v = sm[sidx(j + r * TPX)];
            re[r] = swap_in ? v.y : v.x;
            im[r] = swap_in ? v.x : v.y;
        }
        __syncthreads();                       // inputs consumed everywhere; previous tile fully done
        if (tile + gridDim.x < ntiles) prefetch(tile + gridDim.x, bufs + (size_t)(cur ^ 1) * N * TPB);

        // ---- stage 1 (Ns = 1)
        Butterfly<E, T>::run(re, im);
#pragma unroll
        for (int r = 0; r < E; ++r) {
            cplx<T> v; v.x = re[r]; v.y = im[r];
            sm[sidx(j * E + r)] = v;
        }
        __syncthreads();

        // ---- stage 2: radix R1, Ns = E
        {
            constexpr int NB = N / R1, PER = E / R1, TSTEP = N / (E * R1);
            cplx<T> w[R1];
            load_twiddles<R1, T>(tw, TSTEP * (j % E), w);
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int b = j + i * TPX;
#pragma unroll
                for (int r = 0; r < R1; ++r) {
                    cplx<T> v = sm[sidx(b + r * NB)];
                    if (r > 0) v = cmul(v, w[r]);
                    re[i * R1 + r] = v.x; im[i * R1 + r] = v.y;
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                T xr[R1], xi[R1];
#pragma unroll
                for (int r = 0; r < R1; ++r) { xr[r] = re[i * R1 + r]; xi[r] = im[i * R1 + r]; }
                Butterfly<R1, T>::run(xr, xi);
                const int b = j + i * TPX;
                const int k = b % E;
                const int j0 = (b - k) * R1 + k;
#pragma unroll
                for (int r = 0; r < R1; ++r) {
                    cplx<T> v; v.x = xr[r]; v.y = xi[r];
                    sm[sidx(j0 + r * E)] = v;
                }
            }
            __syncthreads();
        }

        // ---- stage 3: radix R2, Ns = E * R1, outputs stream to HBM from registers
        {
            constexpr int NS = E * R1, PER = E / R2;
            cplx<T> w[R2];
            load_twiddles<R2, T>(tw, j, w);
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int b = j + i * TPX;
                T xr[R2], xi[R2];
#pragma unroll
                for (int r = 0; r < R2; ++r) {
                    cplx<T> v = sm[sidx(b + r * NS)];
                    if (r > 0) {
                        cplx<T> wr = w[r];
                        if (i > 0) wr = cmul(wr, unit_root<E, T>((r * i) % E));
                        v = cmul(v, wr);
                    }
                    xr[r] = v.x; xi[r] = v.y;
                }
                Butterfly<R2, T>::run(xr, xi);
#pragma unroll
                for (int r = 0; r < R2; ++r) {
                    cplx<T> o;
                    o.x = swap_out ? xi[r] : xr[r];
                    o.y = swap_out ? xr[r] : xi[r];
                    if (valid) st_stream(gout + (int64_t)(b + r * NS) * os2, o);
                }
            }
        }
        // no barrier here: the next iteration's first barrier (after its register loads, which
        // touch only the other buffer) orders these shared-memory reads before that buffer is refilled
    }
}

struct PipeEntry {
    int prec, n, tpb, code;
    size_t smem;
    int threads;
    void (*launch)(const b2d_fft_pass &, int, int, unsigned, cudaStream_t);
    const void *func;
};

template <typename T, int N, int E, int R1, int R2, int TPB>
void launch_pipe(const b2d_fft_pass &p, int swap_in, int swap_out, unsigned blocks, cudaStream_t st)
{
    using Cfg = PipeCfg<T, N, E, R1, R2, TPB>;
    pipe_kernel<T, N, E, R1, R2, TPB><<<blocks, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(p, swap_in, swap_out);
}

#define B2_PIPE_ENTRY(PREC, T, N, E, R1, R2, TPB)                                              \
    { PREC, N, TPB, 5000 + TPB, PipeCfg<T, N, E, R1, R2, TPB>::SMEM_BYTES,                     \
      PipeCfg<T, N, E, R1, R2, TPB>::THREADS, &launch_pipe<T, N, E, R1, R2, TPB>,              \
      (const void *)&pipe_kernel<T, N, E, R1, R2, TPB> }

}  // namespace b2pipe
