// fft_tma.cuh -- persistent strided (COL) Stockham kernel fed by the Tensor Memory
// Accelerator.
//
// One CTA per SM.  The CTA owns NBUF shared-memory tile buffers (a tile = TPB adjacent
// pencils x N points, laid out [k][t] exactly as the TMA box arrives) and GROUPS
// independent compute groups of TPX*TPB threads.  Work items (tiles) of the CTA are
// numbered s = 0, 1, 2, ...; item s lands in buffer s % NBUF and is transformed by group
// s % GROUPS:
//
//   TMA:     load(s) is issued as soon as the previous user of buffer s % NBUF has done its
//            last shared-memory read (by one elected thread of that group; 4 bulk-tensor
//            copies of 256 x TPB points each, completion counted on an mbarrier)
//   group:   wait full[s % NBUF] -> registers | stage 1 | exchange in the same buffer |
//            stage 2 | exchange | stage-3 reads | release buffer + issue load(s + NBUF) |
//            stage-3 butterflies, streaming stores from registers
//
// So while the two groups compute, NBUF - GROUPS tiles are in flight into shared memory
// with no register or LSU cost, which the register-resident kernel (fft_fast.cuh: loads
// land in registers, <= 8192 points per SM, nothing in flight during compute) and the
// LDGSTS pipeline (fft_pipe.cuh: one compute group, too few warps) could not do.
//
// Reference counterpart: the buffered strided solver dft/buffered.c:41-69 (copy a batch
// of pencils to a contiguous buffer, transform there, copy back) -- the copy-in is what
// the TMA engine does here, concurrently.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "fft_fast.cuh"

namespace b2tma {
using b2::cplx;
using b2::cmul;
using b2fast::st_stream;
using b2fast::smem_twiddles;
using b2fast::tw_nmult;
using b2fast::tw_mult;
using b2fast::ldg_c;
using b2fast::unit_root;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int THREADS>
__device__ __forceinline__ void group_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(THREADS) : "memory"); }

template <typename T, int N, int E, int R1, int R2, int TPB, int GROUPS, int NBUF>
struct TmaCfg {
    static constexpr int TPX = N / E;
    static constexpr int TG = TPX * TPB;                       // threads of one compute group
    static constexpr int THREADS = TG * GROUPS;
    static constexpr int BOXK = N < 256 ? N : 256;             // TMA box: BOXK points x TPB pencils
    static constexpr size_t TILE_BYTES = (size_t)N * TPB * sizeof(cplx<T>);          // what the TMA delivers
    static constexpr int BUF_ELEMS = (N + N / 16) * TPB;          // buffer with room for the padded exchange layout
    static constexpr int TW2_ELEMS = tw_nmult(R1) * E;           // stage-2 twiddles, index k < E
    static constexpr int TW3_ELEMS = tw_nmult(R2) * TPX;         // stage-3 twiddles, index j < TPX
    static constexpr size_t SMEM_BYTES = (size_t)(NBUF * BUF_ELEMS + TW2_ELEMS + TW3_ELEMS) * sizeof(cplx<T>) + NBUF * sizeof(uint64_t) + 128;
    static_assert((BUF_ELEMS * sizeof(cplx<T>)) % 128 == 0, "TMA destination alignment");
    static_assert(E * R1 * R2 == N && R2 > 1, "three-stage sizes only");
    static_assert(E % R1 == 0 && E % R2 == 0 && TPX % E == 0, "radix layout");
    static_assert(TG % 32 == 0 && THREADS <= 1024, "whole warps per group");
    static_assert(NBUF > GROUPS, "at least one tile in flight");
};

template <typename T, int N, int E, int R1, int R2, int TPB, int GROUPS, int NBUF>
__global__ void __launch_bounds__(TmaCfg<T, N, E, R1, R2, TPB, GROUPS, NBUF>::THREADS, 1)
tma_kernel(const __grid_constant__ b2d_fft_pass p, const __grid_constant__ CUtensorMap tmap, int swap_in, int swap_out, int pair)
{
    using Cfg = TmaCfg<T, N, E, R1, R2, TPB, GROUPS, NBUF>;
    constexpr int TPX = Cfg::TPX, TG = Cfg::TG;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127);
    cplx<T> *bufs = reinterpret_cast<cplx<T> *>(base);
    cplx<T> *tws2 = bufs + (size_t)NBUF * Cfg::BUF_ELEMS, *tws3 = tws2 + Cfg::TW2_ELEMS;
    uint64_t *full = reinterpret_cast<uint64_t *>(tws3 + Cfg::TW3_ELEMS);

    const int tid = threadIdx.x;
    const int g = tid / TG, tg = tid % TG;
    const int t = tg % TPB, j = tg / TPB;
    const int64_t ntiles = b2::grid_blocks(p);
    // work item s of this CTA -> tile: `pair` consecutive items are adjacent tiles (their rows share
    // 128-byte lines and TLB entries); the host guarantees ntiles % pair == 0
    const int64_t nsets = ntiles / pair;
    const int64_t nseq = (int64_t)blockIdx.x < nsets ? (nsets - blockIdx.x + gridDim.x - 1) / gridDim.x * pair : 0;
    auto tile_of = [&](int64_t s) -> int64_t {
        return pair == 1 ? (int64_t)blockIdx.x + s * gridDim.x
                         : ((int64_t)blockIdx.x + (s >> 1) * gridDim.x) * 2 + (s & 1);
    };
    cplx<T> *gout_base = reinterpret_cast<cplx<T> *>(swap_out ? p.out_im : p.out_re);
    const unsigned os2 = (unsigned)(p.os / 2);          // try_launch checks that N * os / 2 fits 32 bits
    const cplx<T> *tw = reinterpret_cast<const cplx<T> *>(p.tw);

    auto issue = [&](int64_t s) {                       // one thread
        const b2::TileCtx c = b2::decode_block(p, tile_of(s));
        const int b = (int)(s % NBUF);
        mbar_expect_tx(&full[b], (unsigned)Cfg::TILE_BYTES);
        cplx<T> *dst = bufs + (size_t)b * Cfg::BUF_ELEMS;
#pragma unroll
        for (int kc = 0; kc < N / Cfg::BOXK; ++kc)
            tma_load_4d(dst + (size_t)kc * Cfg::BOXK * TPB, &tmap, (int)(c.tile0 * TPB * 2), kc * Cfg::BOXK,
                        (int)c.b1, (int)c.b2, &full[b]);
    };

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NBUF; ++b) mbar_init(&full[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // stage twiddles, once per CTA (fft_fast.cuh layout [multiplier][index])
        constexpr int TSTEP2 = N / (E * R1);
        for (int idx = tid; idx < Cfg::TW2_ELEMS; idx += Cfg::THREADS)
            tws2[idx] = ldg_c(&tw[TSTEP2 * tw_mult(idx / E) * (idx % E)]);
        for (int idx = tid; idx < Cfg::TW3_ELEMS; idx += Cfg::THREADS)
            tws3[idx] = ldg_c(&tw[tw_mult(idx / TPX) * (idx % TPX)]);
    }
    __syncthreads();
    if (tid == 0) {
        // ptxas schedules the mbarrier-init fence above ahead of the (uniform-datapath) init stores;
        // fence again, after them, so the TMA engine's complete_tx never sees an uninitialised barrier
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        for (int64_t s = 0; s < NBUF && s < nseq; ++s) issue(s);
    }

    for (int64_t s = g; s < nseq; s += GROUPS) {
        const int bi = (int)(s % NBUF);
        cplx<T> *sm = bufs + (size_t)bi * Cfg::BUF_ELEMS;
        // exchanges use a padded layout (one extra row of TPB points every 16): stage-1 writes of
        // neighbouring j are 16 rows apart and would otherwise hit the same banks
        auto sidx = [&](int k) -> int { return (k + (k >> 4)) * TPB + t; };
        // keep the per-tile address arithmetic inside the loop: hoisted out of the persistent loop it
        // costs 32 registers of loop-invariant 64-bit offsets, which spill
        unsigned osu = os2;
        asm volatile("" : "+r"(osu));
        const b2::TileCtx c = b2::decode_block(p, tile_of(s));
        const int64_t b0 = c.tile0 * TPB + t;
        const bool valid = b0 < p.bn[0];
        cplx<T> *gout = gout_base + (b0 * p.bos[0] + c.b1 * p.bos[1] + c.b2 * p.bos[2]) / 2;

        T re[E], im[E];
        mbar_wait(&full[bi], (unsigned)((s / NBUF) & 1));
#pragma unroll
        for (int r = 0; r < E; ++r) {
            cplx<T> v = sm[(j + r * TPX) * TPB + t];          // dense, as delivered
            re[r] = swap_in ? v.y : v.x;
            im[r] = swap_in ? v.x : v.y;
        }
        group_bar<TG>(1 + g);

        // ---- stage 1 (Ns = 1)
        Butterfly<E, T>::run(re, im);
#pragma unroll
        for (int r = 0; r < E; ++r) {
            cplx<T> v; v.x = re[r]; v.y = im[r];
            sm[sidx(j * E + r)] = v;
        }
        group_bar<TG>(1 + g);

        // ---- stage 2: radix R1, Ns = E
        {
            constexpr int NB = N / R1, PER = E / R1;
            cplx<T> w[R1];
            smem_twiddles<R1, T>(tws2, E, j % E, w);
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
            group_bar<TG>(1 + g);
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
            group_bar<TG>(1 + g);
        }

        // ---- stage 3: radix R2, Ns = E * R1; all shared-memory reads first so the buffer can be refilled
        {
            constexpr int NS = E * R1, PER = E / R2;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int b = j + i * TPX;
#pragma unroll
                for (int r = 0; r < R2; ++r) {
                    cplx<T> v = sm[sidx(b + r * NS)];
                    re[i * R2 + r] = v.x; im[i * R2 + r] = v.y;
                }
            }
            // every thread orders its own generic-proxy accesses of this buffer before the
            // async-proxy refill, then the group meets, then one thread re-arms and issues
            fence_proxy_async();
            group_bar<TG>(1 + g);
            if (tg == 0 && s + NBUF < nseq) issue(s + NBUF);
            cplx<T> w[R2];
            smem_twiddles<R2, T>(tws3, TPX, j, w);
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int b = j + i * TPX;
                T xr[R2], xi[R2];
#pragma unroll
                for (int r = 0; r < R2; ++r) {
                    cplx<T> v; v.x = re[i * R2 + r]; v.y = im[i * R2 + r];
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
                    if (valid) st_stream(gout + (size_t)((unsigned)(b + r * NS) * osu), o);
                }
            }
        }
    }
}

struct TmaEntry {
    int prec, n, tpb, boxk;
    size_t smem;
    int threads;
    void (*launch)(const b2d_fft_pass &, const CUtensorMap &, int, int, int, unsigned, cudaStream_t);
    const void *func;
};

template <typename T, int N, int E, int R1, int R2, int TPB, int GROUPS, int NBUF>
void launch_tma(const b2d_fft_pass &p, const CUtensorMap &m, int swap_in, int swap_out, int pair, unsigned blocks, cudaStream_t st)
{
    using Cfg = TmaCfg<T, N, E, R1, R2, TPB, GROUPS, NBUF>;
    tma_kernel<T, N, E, R1, R2, TPB, GROUPS, NBUF><<<blocks, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(p, m, swap_in, swap_out, pair);
}

#define B2_TMA_ENTRY(PREC, T, N, E, R1, R2, TPB, G, NB)                                        \
    { PREC, N, TPB, TmaCfg<T, N, E, R1, R2, TPB, G, NB>::BOXK,                                 \
      TmaCfg<T, N, E, R1, R2, TPB, G, NB>::SMEM_BYTES, TmaCfg<T, N, E, R1, R2, TPB, G, NB>::THREADS, \
      &launch_tma<T, N, E, R1, R2, TPB, G, NB>, (const void *)&tma_kernel<T, N, E, R1, R2, TPB, G, NB> }

// fft_tma_table.cu
const TmaEntry *entry_for(const b2d_fft_pass &p);
void init(int max_smem);
int try_launch(const b2d_fft_pass &p, cudaStream_t st);       // 0 launched, 1 not applicable, -1 error

}  // namespace b2tma
