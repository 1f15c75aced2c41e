// fft_fast.cuh -- compile-time specialised Stockham kernels for the hot
// power-of-two sizes ("codelets" of this engine; the generic kernel in
// fft_generic.cuh is the any-size solver).
//
// Differences from the generic pass:
//   * the first radix stage reads HBM straight into registers and the last
//     stage stores from registers straight to HBM: shared memory is used only
//     for the exchanges between stages (one buffer, N complex per transform);
//   * everything (N, radices, threads) is a template parameter, so the stage
//     loops are fully unrolled straight-line code around the generated
//     butterflies; interleaved complex data moves as 128-bit (f64) / 64-bit
//     (f32) vectors with streaming (evict-first) cache hints;
//   * twiddles: a thread loads at most 6 table entries per stage and derives the
//     others with one complex multiply each (two-level w^(4a+c) = w^(4a) w^c, and
//     compile-time roots of unity for the butterflies it owns beyond the first)
//     -- the L1 load pipe, not HBM, was the limiter with one load per twiddle
//     (profiles/r01_ncu_full_fast_kernels_summary.txt);
//   * COL variant: a CTA owns TPB adjacent pencils of a strided dimension; lanes
//     walk the pencils, so each HBM access is TPB*16 contiguous bytes and the
//     [k][t] shared layout is bank-conflict free.  ROW variant: lanes walk one
//     contiguous transform.
//
// Reference counterpart: the n1/t1 codelets and their drivers
// (dft/direct.c:92-97, dft/dftw-direct.c:46-56; twiddle policies of
// genfft/twiddle.ml, "-twiddle-log3") plus the buffered strided access of
// dft/buffered.c:41-69 / dft/indirect-transpose.c.
#pragma once
#include <cuda_runtime.h>
#include "fft_generic.cuh"

namespace b2fast {
using b2::cplx;
using b2::cmul;

__device__ __forceinline__ cplx<double> ldg_c(const cplx<double> *p)
{
    double2 v = __ldg(reinterpret_cast<const double2 *>(p));
    cplx<double> r; r.x = v.x; r.y = v.y; return r;
}
__device__ __forceinline__ cplx<float> ldg_c(const cplx<float> *p)
{
    float2 v = __ldg(reinterpret_cast<const float2 *>(p));
    cplx<float> r; r.x = v.x; r.y = v.y; return r;
}
// streaming loads / stores: the array is touched once per pass
__device__ __forceinline__ cplx<double> ld_stream(const cplx<double> *p)
{
    double2 v = __ldcs(reinterpret_cast<const double2 *>(p));
    cplx<double> r; r.x = v.x; r.y = v.y; return r;
}
__device__ __forceinline__ cplx<float> ld_stream(const cplx<float> *p)
{
    float2 v = __ldcs(reinterpret_cast<const float2 *>(p));
    cplx<float> r; r.x = v.x; r.y = v.y; return r;
}
__device__ __forceinline__ void st_stream(cplx<double> *p, cplx<double> v)
{
    __stcs(reinterpret_cast<double2 *>(p), make_double2(v.x, v.y));
}
__device__ __forceinline__ void st_stream(cplx<float> *p, cplx<float> v)
{
    __stcs(reinterpret_cast<float2 *>(p), make_float2(v.x, v.y));
}

// L2-resident pass pairs (b2d_fft_pass.cache): the first pass of a pair leaves its output in L2 with
// ordinary write-back stores (bit 1) or pins it there with an evict_last policy (bit 2); the second
// pass reads it back with ordinary loads (bit 0) and streams its own output out.
__device__ __forceinline__ cplx<double> ld_plain(const cplx<double> *p)
{
    double2 v = *reinterpret_cast<const double2 *>(p);
    cplx<double> r; r.x = v.x; r.y = v.y; return r;
}
__device__ __forceinline__ cplx<float> ld_plain(const cplx<float> *p)
{
    float2 v = *reinterpret_cast<const float2 *>(p);
    cplx<float> r; r.x = v.x; r.y = v.y; return r;
}
__device__ __forceinline__ void st_plain(cplx<double> *p, cplx<double> v)
{
    *reinterpret_cast<double2 *>(p) = make_double2(v.x, v.y);
}
__device__ __forceinline__ void st_plain(cplx<float> *p, cplx<float> v)
{
    *reinterpret_cast<float2 *>(p) = make_float2(v.x, v.y);
}
__device__ __forceinline__ unsigned long long policy_evict_last()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_keep(cplx<double> *p, cplx<double> v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_keep(cplx<float> *p, cplx<float> v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
// store of one output element under the pass's cache hints
template <typename T>
__device__ __forceinline__ void st_out(cplx<T> *p, cplx<T> v, int cache, unsigned long long pol)
{
    if (cache & 4) st_keep(p, v, pol);
    else if (cache & 2) st_plain(p, v);
    else st_stream(p, v);
}

// Loads that ask L2 to fetch a whole 128 / 256-byte block from DRAM: a narrow COL tile
// (64 B per row) then costs DRAM one long burst per block instead of several short ones; the
// neighbouring tiles (other CTAs, scheduled next to this one) find their part in L2.
template <int BYTES>
__device__ __forceinline__ cplx<double> ld_l2pf(const cplx<double> *p)
{
    cplx<double> r;
    if (BYTES == 256) asm volatile("ld.global.L2::256B.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    else asm volatile("ld.global.L2::128B.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
template <int BYTES>
__device__ __forceinline__ cplx<float> ld_l2pf(const cplx<float> *p)
{
    cplx<float> r;
    if (BYTES == 256) asm volatile("ld.global.L2::256B.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    else asm volatile("ld.global.L2::128B.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

// compile-time root of unity W_E^m (m is a constant after unrolling)
template <int E, typename T>
__device__ __forceinline__ cplx<T> unit_root(int m)
{
    cplx<T> c;
    if (E == 4) unit_root4<T>(m, c.x, c.y);
    else if (E == 8) unit_root8<T>(m, c.x, c.y);
    else if (E == 10) unit_root10<T>(m, c.x, c.y);
    else if (E == 16) unit_root16<T>(m, c.x, c.y);
    else unit_root32<T>(m, c.x, c.y);
    return c;
}

// w[r] = table[step * r] for r = 1..R-1 (w[0] unused): direct loads for R <= 4,
// two-level for larger R (3 + R/4 - 1 loads, one complex multiply for the rest)
template <int R, typename T>
__device__ __forceinline__ void load_twiddles(const cplx<T> *tw, int step, cplx<T> (&w)[R])
{
    if (R <= 4) {
#pragma unroll
        for (int r = 1; r < R; ++r) w[r] = ldg_c(&tw[step * r]);
    } else {
#pragma unroll
        for (int c = 1; c < 4; ++c) w[c] = ldg_c(&tw[step * c]);
#pragma unroll
        for (int a = 1; a < (R + 3) / 4; ++a) {
            w[4 * a] = ldg_c(&tw[step * 4 * a]);
#pragma unroll
            for (int c = 1; c < 4; ++c) if (4 * a + c < R) w[4 * a + c] = cmul(w[4 * a], w[c]);
        }
    }
}

// Stage twiddles staged in shared memory once per CTA: the table entries a thread needs are
// W^(m*k) for the multipliers m in {1,2,3} (and {4,8,..} for radices > 4); reading them with LDS
// (~30 cycles) instead of LDG through L1/L2 (up to ~300 cycles) removes most long-scoreboard stalls.
__host__ __device__ constexpr int tw_nmult(int R) { return R <= 1 ? 0 : (R <= 4 ? R - 1 : 3 + ((R + 3) / 4 - 1)); }
__host__ __device__ constexpr int tw_mult(int mi) { return mi < 3 ? mi + 1 : 4 * (mi - 2); }

// w[r] for r = 1..R-1 from a shared table laid out [mi][count] (count entries per multiplier)
template <int R, typename T>
__device__ __forceinline__ void smem_twiddles(const cplx<T> *tab, int count, int k, cplx<T> (&w)[R])
{
    if (R <= 4) {
#pragma unroll
        for (int r = 1; r < R; ++r) w[r] = tab[(r - 1) * count + k];
    } else {
#pragma unroll
        for (int c = 1; c < 4; ++c) w[c] = tab[(c - 1) * count + k];
#pragma unroll
        for (int a = 1; a < (R + 3) / 4; ++a) {
            w[4 * a] = tab[(2 + a) * count + k];
#pragma unroll
            for (int c = 1; c < 4; ++c) if (4 * a + c < R) w[4 * a + c] = cmul(w[4 * a], w[c]);
        }
    }
}

// padded row pitch for the ROW layout (same padding rule as the generic kernel)
__host__ __device__ constexpr int padk_c(int k) { return k + (k >> 4); }
__host__ __device__ constexpr int pitch_c(int n) { return ((padk_c(n - 1) + 1 + 14) / 16) * 16 + 1; }

template <typename T, int N, int E, int R1, int R2, int TPB, bool COL, int FLAVOR>
struct FastCfg {
    static constexpr int TPX = N / E;
    static constexpr int THREADS = TPX * TPB;
    // COL layout [k][t]: rows of TPB * sizeof(cplx) bytes.  With rows shorter than 128 bytes the stage-1 exchange
    // writes (row index j * E + r: consecutive lanes are E rows = a multiple of 1 KiB apart) land in the same banks
    // twice per quarter-warp (ncu, round 2: L1 wavefronts 64 % of peak, mio_throttle on the 64-byte-tile kernel):
    // one spare row per 16 skews them apart.
    static constexpr bool PADCOL = COL && (TPB * sizeof(cplx<T>) < 128);
    static constexpr int SMEM_ELEMS = COL ? (PADCOL ? (N + N / 16) * TPB : N * TPB) : pitch_c(N) * TPB;
    static constexpr int TW2_ELEMS = tw_nmult(R1) * E;            // stage-2 twiddles, index k < E
    static constexpr int TW3_ELEMS = tw_nmult(R2) * (N / E);      // stage-3 twiddles, index j < TPX
    static constexpr size_t SMEM_BYTES = (size_t)(SMEM_ELEMS + TW2_ELEMS + TW3_ELEMS) * sizeof(cplx<T>);
    // resident CTAs per SM we ask ptxas to make room for (register cap)
    static constexpr int BY_SMEM = (int)(232448 / (SMEM_BYTES + 1024)) > 0 ? (int)(232448 / (SMEM_BYTES + 1024)) : 1;
    static constexpr int BY_REGS = 65536 / (THREADS * (sizeof(T) == 8 ? (E >= 16 ? 84 : 56) : (E >= 32 ? 84 : (E >= 16 ? 56 : 40))));
    static constexpr int BY_THREADS = 2048 / THREADS;
    static constexpr int MINB_ = BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS;
    static constexpr int MINB__ = (MINB_ < BY_THREADS ? MINB_ : BY_THREADS) > 0 ? (MINB_ < BY_THREADS ? MINB_ : BY_THREADS) : 1;
    // flavor 1: cap registers for as many resident CTAs as shared memory allows; strided kernels of 256 threads: keep
    // the two CTAs per SM that overlap each other's load and compute phases (cap 128 registers)
    // r2r line kernels (flavor 9, ROW) of 256 threads: one CTA per SM leaves every phase (staged loads, three
    // exchanges, POST map) exposed -- 150 registers uncapped, 248 us for 4096 x 4096 REDFT10 rows against 167 with
    // two lines pairs per CTA (profiles/r02_c5b_pieces.txt); capped at 128 two CTAs overlap
    // (one-CTA Bluestein, flavor 7, stays uncapped: at 168 registers a third 128-thread CTA fits but the spills cost
    // more than it hides -- 618 vs 573 us for 1009 x 16384 double, profiles/r02_c5a_experiment.log)
    static constexpr int MINB = FLAVOR == 1 ? MINB__ : (((COL ? FLAVOR != 9 : FLAVOR == 9) && THREADS <= 256 && BY_SMEM >= 2 && sizeof(T) == 8) ? 2 : 1);
    static_assert(E * R1 * R2 == N, "radices must multiply to N");
    static_assert(E % R1 == 0 && E % R2 == 0, "later radices must divide the per-thread element count");
    static_assert(TPX % E == 0 || R2 == 1, "stage-2 twiddle index must be thread-constant");
};

// smem-staged real line (FLAVOR 9 ROW: the tile's lines are brought in with coalesced loads first)
template <typename T>
struct SmemLineIn {
    const T *p;
    __device__ __forceinline__ T operator()(int j) const { return p[j]; }
};

// KIND >= 0: the r2r kind is a compile-time constant (the PRE / POST switches fold away: the runtime-kind
// kernel is 13.9k SASS instructions and instruction-cache bound, profiles/r01_ncu_full_r2r4096_summary.txt).
// one tile (CTA-sized unit of work) of the pass; `block` is its index in the grid of tiles
template <typename T, int N, int E, int R1, int R2, int TPB, bool COL, int FLAVOR, int KIND>
__device__ __forceinline__ void fast_tile(const b2d_fft_pass &p, int swap_in, int swap_out, int64_t block,
                                          unsigned char *smem_raw)
{
    using Cfg = FastCfg<T, N, E, R1, R2, TPB, COL, FLAVOR>;
    const int r2r_kind = KIND >= 0 ? KIND : p.r2r_kind;
    constexpr int TPX = Cfg::TPX;
    cplx<T> *sm = reinterpret_cast<cplx<T> *>(smem_raw);

    const int tid = threadIdx.x;
    const int t = COL ? (tid % TPB) : (tid / TPX);
    const int j = COL ? (tid / TPB) : (tid % TPX);
    auto sidx = [&](int k) -> int { return COL ? ((k + (Cfg::PADCOL ? (k >> 4) : 0)) * TPB + t) : (t * pitch_c(N) + padk_c(k)); };

    const b2::TileCtx c = b2::decode_block(p, block);
    // FLAVOR 10 (last pass of an even-size r2c, rdft/ct-hc2c.c:146-273 / ct-hc2c-direct.c:45-60 in the reference):
    // the CTA transforms HALF = TPB / 2 rows k1 of the four-step's second pass together with their mirror rows
    // n1 - k1, so that every pair (k, m - k) of the half-size spectrum Z meets in shared memory and the split
    // X_k = 1/2 [(Z_k + conj Z_{m-k}) - i w^k (Z_k - conj Z_{m-k})] rides on the store: no separate pass over HBM.
    // Tiles: rows [1 + HALF*tile, 1 + HALF*(tile+1)) and their mirrors; the last tile holds row 0 (its own mirror).
    constexpr int HALF = TPB / 2 > 0 ? TPB / 2 : 1;
    const int64_t f10_n1 = p.aux_split;
    const bool f10_row0 = FLAVOR == 10 && c.tile0 * (2 * HALF) >= f10_n1;
    int64_t b0_ = c.tile0 * TPB + t;
    bool valid_ = b0_ < p.bn[0];
    if (FLAVOR == 10) {
        if (f10_row0) { b0_ = 0; valid_ = (t == 0); }
        else {
            const int64_t k1 = 1 + (int64_t)HALF * c.tile0 + (t % HALF);
            b0_ = t < HALF ? k1 : f10_n1 - k1;
            valid_ = true;
        }
    }
    const int64_t b0 = b0_;
    const bool valid = valid_;
    const int64_t boff_in = b0 * p.bis[0] + c.b1 * p.bis[1] + c.b2 * p.bis[2];
    const int64_t boff_out = b0 * p.bos[0] + c.b1 * p.bos[1] + ((p.npeer && FLAVOR != 8) ? 0 : c.b2 * p.bos[2]);
    // interleaved data: vector pointer at the lower of (re, im)
    const cplx<T> *gin = reinterpret_cast<const cplx<T> *>(swap_in ? p.in_im : p.in_re) + boff_in / 2;
    // peer scatter: batch dim 2 selects the destination buffer (a peer GPU's exchange
    // buffer mapped over NVLink): the transpose is fused with its collective
    cplx<T> *gout = (p.npeer ? reinterpret_cast<cplx<T> *>(p.peer_out[c.b2])
                             : reinterpret_cast<cplx<T> *>(swap_out ? p.out_im : p.out_re)) + boff_out / 2;
    const int64_t is2 = p.is / 2, os2 = p.os / 2;          // strides in complex units
    const cplx<T> *tw = reinterpret_cast<const cplx<T> *>(p.tw);
    cplx<T> *tws2 = sm + Cfg::SMEM_ELEMS, *tws3 = tws2 + Cfg::TW2_ELEMS;

    // final-stage output of element kout of this thread's transform
    //   FLAVOR 2: four-step twiddle W_big^(kout * b0) fused into the store (two-level table)
    //   FLAVOR 3: ROW tile stored in COL order: park the result in shared memory (same
    //             position the thread just read), the CTA streams it out below
    //   FLAVOR 8: COL kernel whose output ROWS are split over peer GPUs (row k -> peer k / peer_rows):
    //             the second exchange of a distributed transform fused into its last pass
    //   FLAVOR 9: real-to-real kinds: the PRE map gathers the first stage's inputs from the real line,
    //             the POST map scatters the last stage's outputs into it (r2r_maps.cuh)
    //   FLAVOR 7: Bluestein in one CTA (dft/bluestein.c:82-128): the stages run twice; the first
    //             run's outputs are multiplied by B = FFT(filter), conjugated and kept in registers
    //             -- output b + r*Ns of the last stage IS input j + q*TPX of the next first stage,
    //             q = i + PER*r -- the second run's outputs get conj * chirp * 1/M and are cut to n_out
    T bre[FLAVOR == 7 ? E : 1], bim[FLAVOR == 7 ? E : 1];
    int rep = 0;
    const int cache = p.cache;
    const unsigned long long keep_pol = (cache & 4) ? policy_evict_last() : 0ull;
    auto emit = [&](int kout, int q, T vr, T vi) {
        if (FLAVOR == 9) {
            if (p.r2r_pair) {       // two real lines per transform: park, the spectra are separated in flush_col()
                cplx<T> o; o.x = vr; o.y = vi;
                sm[sidx(kout)] = o;
                return;
            }
            if (valid) {
                b2::RealLineOut<T> y = { reinterpret_cast<T *>(p.out_re) + boff_out, p.os };
                cplx<T> v; v.x = vr; v.y = vi;
                b2::r2r_post_scatter<T>(r2r_kind, p.n_out, kout, v, reinterpret_cast<const cplx<T> *>(p.aux0), y);
            }
            return;
        }
        if (FLAVOR == 7) {
            cplx<T> v; v.x = vr; v.y = vi;
            if (rep == 0) {
                v = cmul(v, ldg_c(reinterpret_cast<const cplx<T> *>(p.aux1) + kout));
                bre[FLAVOR == 7 ? q : 0] = v.x; bim[FLAVOR == 7 ? q : 0] = -v.y;
            } else if (valid && kout < p.n_out) {
                v.y = -v.y;
                v = cmul(v, ldg_c(reinterpret_cast<const cplx<T> *>(p.aux0) + kout));
                cplx<T> o;
                o.x = (swap_out ? v.y : v.x) * (T)p.scale;
                o.y = (swap_out ? v.x : v.y) * (T)p.scale;
                st_stream(gout + (int64_t)kout * os2, o);
            }
            return;
        }
        if (FLAVOR == 2 || FLAVOR == 11) {
            int64_t e = (int64_t)kout * (b0 + p.tw4_off);
            int64_t eh, el;
            if (p.tw4_shift >= 0) { e &= (p.big_n - 1); eh = e >> p.tw4_shift; el = e & (p.aux_split - 1); }
            else { e %= p.big_n; eh = e / p.aux_split; el = e - eh * p.aux_split; }
            cplx<T> w = cmul(ldg_c(reinterpret_cast<const cplx<T> *>(p.aux1) + eh),
                             ldg_c(reinterpret_cast<const cplx<T> *>(p.aux0) + el));
            cplx<T> v; v.x = vr; v.y = vi;
            v = cmul(v, w);
            vr = v.x; vi = v.y;
        }
        cplx<T> o;
        if (FLAVOR == 3 || FLAVOR == 10) {
            o.x = vr; o.y = vi;
            sm[sidx(kout)] = o;
        } else if (FLAVOR == 8) {
            o.x = swap_out ? vi : vr;
            o.y = swap_out ? vr : vi;
            const int peer = p.tw4_shift >= 0 ? (kout >> p.tw4_shift) : kout / p.peer_rows;
            const int krow = kout - peer * p.peer_rows;
            if (valid) st_stream(reinterpret_cast<cplx<T> *>(p.peer_out[peer]) + boff_out / 2 + (int64_t)krow * os2, o);
        } else {
            o.x = swap_out ? vi : vr;
            o.y = swap_out ? vr : vi;
            if (valid) {
                if (FLAVOR == 4) *(gout + (int64_t)kout * os2) = o;     // let L2 merge the halves of a line
                else st_out<T>(gout + (int64_t)kout * os2, o, cache, keep_pol);
            }
        }
    };
    auto flush_col = [&]() {
        if (FLAVOR == 9) {
            if (!p.r2r_pair) return;
            __syncthreads();
            const cplx<T> *qt = reinterpret_cast<const cplx<T> *>(p.aux0);
            if (!COL && p.store_col && p.pair_os == 1 && p.bos[0] == 2 && !((p.os | p.bos[1] | p.bos[2]) & 1) &&
                !(reinterpret_cast<uintptr_t>(p.out_re) % (2 * sizeof(T)))) {
                // lines stored transposed: element k of the tile's 2 * TPB neighbouring lines is one contiguous
                // piece of the output.  Every thread takes whole pieces (all lines of a k) and writes each pair
                // of lines as one vector, piece by piece -- full sectors instead of one scalar per sector.
                T *ob = reinterpret_cast<T *>(p.out_re) + (c.tile0 * TPB) * p.bos[0] + c.b1 * p.bos[1] + c.b2 * p.bos[2];
                for (int k = tid; k < N; k += Cfg::THREADS) {
#pragma unroll
                    for (int tt = 0; tt < TPB; ++tt) {
                        if (c.tile0 * TPB + tt >= p.bn[0]) break;
                        cplx<T> u, v;
                        b2::r2r_unpack_pair<T>(sm[tt * pitch_c(N) + padk_c(k)], sm[tt * pitch_c(N) + padk_c(k ? N - k : 0)], u, v);
                        b2::StashOut<T> sa, sb;
                        sa.cnt = sb.cnt = 0;
                        b2::r2r_post_scatter<T>(r2r_kind, p.n_out, k, u, qt, sa);
                        b2::r2r_post_scatter<T>(r2r_kind, p.n_out, k, v, qt, sb);
                        for (int w = 0; w < sa.cnt; ++w) {
                            cplx<T> o; o.x = sa.val[w]; o.y = sb.val[w];
                            st_plain(reinterpret_cast<cplx<T> *>(ob + (int64_t)sa.idx[w] * p.os + 2 * tt), o);
                        }
                    }
                }
                return;
            }
            if (!valid) return;
            b2::RealLineOut<T> ya = { reinterpret_cast<T *>(p.out_re) + boff_out, p.os };
            b2::RealLineOut<T> yb = { reinterpret_cast<T *>(p.out_re) + boff_out + p.pair_os, p.os };
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int k = j + r * TPX;
                cplx<T> u, v;
                b2::r2r_unpack_pair<T>(sm[sidx(k)], sm[sidx(k ? N - k : 0)], u, v);
                b2::r2r_post_scatter<T>(r2r_kind, p.n_out, k, u, qt, ya);
                b2::r2r_post_scatter<T>(r2r_kind, p.n_out, k, v, qt, yb);
            }
            return;
        }
        if (FLAVOR == 10) {
            __syncthreads();
            const cplx<T> *w = reinterpret_cast<const cplx<T> *>(p.aux0);      // exp(-2 pi i q / n), q <= m = n / 2
            cplx<T> *ob = reinterpret_cast<cplx<T> *>(p.out_re) + (c.b1 * p.bos[1] + c.b2 * p.bos[2]) / 2;
            const int64_t s1 = p.bos[0] / 2, s2 = p.os / 2;                    // complex strides of k1 and of k2
            auto pair = [&](cplx<T> a, cplx<T> cz, int64_t q, cplx<T> &xq, cplx<T> &xp) {
                const T sr = a.x + cz.x, si = a.y - cz.y, dr = a.x - cz.x, di = a.y + cz.y;
                const cplx<T> wq = ldg_c(w + q);
                const T tr = wq.x * di + wq.y * dr, ti = -(wq.x * dr - wq.y * di);
                xq.x = T(0.5) * (sr + tr); xq.y = T(0.5) * (si + ti);
                xp.x = T(0.5) * (sr - tr); xp.y = T(0.5) * (-si + ti);
            };
            if (!f10_row0) {
                for (int idx = tid; idx < N * HALF; idx += Cfg::THREADS) {
                    const int ta = idx % HALF, k2 = idx / HALF;
                    const int64_t k1 = 1 + (int64_t)HALF * c.tile0 + ta;
                    cplx<T> xq, xp;
                    pair(sm[ta * pitch_c(N) + padk_c(k2)], sm[(HALF + ta) * pitch_c(N) + padk_c(N - 1 - k2)],
                         k1 + f10_n1 * (int64_t)k2, xq, xp);
                    st_stream(ob + k1 * s1 + (int64_t)k2 * s2, xq);
                    st_stream(ob + (f10_n1 - k1) * s1 + (int64_t)(N - 1 - k2) * s2, xp);
                }
            } else {
                for (int k2 = tid; k2 <= N / 2; k2 += Cfg::THREADS) {
                    if (k2 == 0) {
                        const cplx<T> z0 = sm[0];
                        cplx<T> x0, xm;
                        x0.x = z0.x + z0.y; x0.y = T(0); xm.x = z0.x - z0.y; xm.y = T(0);
                        st_stream(ob, x0);
                        st_stream(ob + (int64_t)N * s2, xm);                    // Nyquist bin, index m = n1 * N
                    } else {
                        cplx<T> xq, xp;
                        pair(sm[padk_c(k2)], sm[padk_c(N - k2)], f10_n1 * (int64_t)k2, xq, xp);
                        st_stream(ob + (int64_t)k2 * s2, xq);
                        if (2 * k2 != N) st_stream(ob + (int64_t)(N - k2) * s2, xp);
                    }
                }
            }
            return;
        }
        if (FLAVOR != 3) return;
        __syncthreads();
        cplx<T> *gbase = reinterpret_cast<cplx<T> *>(swap_out ? p.out_im : p.out_re);
        for (int idx = tid; idx < N * TPB; idx += Cfg::THREADS) {
            const int tt = idx % TPB, k = idx / TPB;
            const int64_t bb = c.tile0 * TPB + tt;
            if (bb >= p.bn[0]) continue;
            cplx<T> v = sm[tt * pitch_c(N) + padk_c(k)];
            cplx<T> o;
            o.x = swap_out ? v.y : v.x;
            o.y = swap_out ? v.x : v.y;
            st_out<T>(gbase + (bb * p.bos[0] + c.b1 * p.bos[1] + c.b2 * p.bos[2]) / 2 + (int64_t)k * os2, o, cache, keep_pol);
        }
    };

    T re[E], im[E];
    // FLAVOR 9, ROW: the PRE maps gather with strides 2 / -2 (types 2, 3) or mirrored (types 1, HC2R): bring
    // the tile's real lines into shared memory with coalesced loads first (the raw lines alias the exchange
    // buffer: n_in reals per line <= N complex per transform), then gather from there
    T *raw = reinterpret_cast<T *>(sm);
    if (FLAVOR == 9 && !COL) {
        // eight independent loads in flight per thread and round (a plain strided loop would wait for each load
        // before issuing the next: 32 serial HBM round trips per tile, profiles/r02_c5b_pieces.txt)
        const int nl = p.r2r_pair ? 2 : 1, nin = p.n_in;
        const int total = TPB * nl * nin;
        const T *gbase = reinterpret_cast<const T *>(p.in_re) + c.b1 * p.bis[1] + c.b2 * p.bis[2];
        // lines of an even number of reals at even offsets: whole vectors (two reals), sixteen in flight -- the
        // tile of a 4096-point double pass arrives in ONE round trip instead of four
        const bool vec = !((nin | p.bis[0] | p.bis[1] | p.bis[2] | (p.r2r_pair ? p.pair_is : 0)) & 1) &&
                         !(reinterpret_cast<uintptr_t>(p.in_re) % (2 * sizeof(T)));
        if (vec) {
            const int nin2 = nin / 2, total2 = total / 2;
            const cplx<T> *g2 = reinterpret_cast<const cplx<T> *>(gbase);
            cplx<T> *raw2 = reinterpret_cast<cplx<T> *>(raw);
            for (int base = 0; base < total2; base += 16 * Cfg::THREADS) {
                cplx<T> v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int idx = base + u * Cfg::THREADS + tid;
                    v[u].x = T(0); v[u].y = T(0);
                    if (idx < total2) {
                        const int tt = idx / (nl * nin2), rem = idx - tt * (nl * nin2);
                        const int ln = rem / nin2, e = rem - ln * nin2;
                        const int64_t bb = c.tile0 * TPB + tt;
                        if (bb < p.bn[0]) v[u] = ld_stream(g2 + (bb * p.bis[0] + (ln ? p.pair_is : 0)) / 2 + e);
                    }
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int idx = base + u * Cfg::THREADS + tid;
                    if (idx < total2) raw2[idx] = v[u];
                }
            }
        } else
        for (int base = 0; base < total; base += 8 * Cfg::THREADS) {
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * Cfg::THREADS + tid;
                v[u] = T(0);
                if (idx < total) {
                    const int tt = idx / (nl * nin), rem = idx - tt * (nl * nin);
                    const int ln = rem / nin, e = rem - ln * nin;
                    const int64_t bb = c.tile0 * TPB + tt;
                    if (bb < p.bn[0]) v[u] = __ldcs(gbase + bb * p.bis[0] + (ln ? p.pair_is : 0) + e);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * Cfg::THREADS + tid;
                if (idx < total) raw[idx] = v[u];
            }
        }
        __syncthreads();
    }
    // ---- stage 1: radix E straight from HBM (butterfly index b = j, Ns = 1)
#pragma unroll
    for (int r = 0; r < E; ++r) {
        cplx<T> v; v.x = T(0); v.y = T(0);
        if (FLAVOR == 9 && !COL) {
            if (valid) {
                const int nl = p.r2r_pair ? 2 : 1;
                SmemLineIn<T> x = { raw + (t * nl) * p.n_in };
                v = b2::r2r_pre_value<T>(r2r_kind, p.n_in, j + r * TPX, reinterpret_cast<const cplx<T> *>(p.aux0), x);
                if (p.r2r_pair) {
                    SmemLineIn<T> x2 = { raw + (t * nl + 1) * p.n_in };
                    v.y = b2::r2r_pre_value<T>(r2r_kind, p.n_in, j + r * TPX, reinterpret_cast<const cplx<T> *>(p.aux0), x2).x;
                }
            }
        } else if (FLAVOR == 9) {
            if (valid) {
                b2::RealLineIn<T> x = { reinterpret_cast<const T *>(p.in_re) + boff_in, p.is };
                v = b2::r2r_pre_value<T>(r2r_kind, p.n_in, j + r * TPX, reinterpret_cast<const cplx<T> *>(p.aux0), x);
                if (p.r2r_pair) {
                    b2::RealLineIn<T> x2 = { reinterpret_cast<const T *>(p.in_re) + boff_in + p.pair_is, p.is };
                    v.y = b2::r2r_pre_value<T>(r2r_kind, p.n_in, j + r * TPX, reinterpret_cast<const cplx<T> *>(p.aux0), x2).x;
                }
            }
        } else if (FLAVOR == 7) {
            const int k = j + r * TPX;
            if (valid && k < p.n_in) v = ld_stream(gin + (int64_t)k * is2);
        } else if (FLAVOR == 11) {
            // first pass of an even-size c2r (B2D_LOAD_C2R_MERGE): element k of column b0 is logical index
            // jl = k * idx_mul + b0 of the Hermitian half X[0..m]; its mirror X[m - jl] lies (m - 2 jl) complex
            // elements on (columns are adjacent).  Consecutive lanes read consecutive X and, backwards, consecutive
            // mirrors; the mirror tile is some other CTA's own tile, so one of the two reads hits L2.
            if (valid) {
                const int64_t m = p.n_in;
                const int64_t jl = (int64_t)(j + r * TPX) * p.idx_mul + b0;
                const cplx<T> *pa = gin + (int64_t)(j + r * TPX) * is2;
                cplx<T> a = ld_plain(pa), cm = ld_plain(pa + (m - 2 * jl));      // no evict-first: the other read of each is still to come
                cplx<T> w = ldg_c(reinterpret_cast<const cplx<T> *>(p.aux2) + jl);
                if (jl == 0) { a.y = T(0); cm.y = T(0); }
                const T sr = a.x + cm.x, si = a.y - cm.y, dr = a.x - cm.x, di = a.y + cm.y;
                w.y = -w.y;
                v.x = si + (w.x * dr - w.y * di);        // swapped (Im Z, Re Z): the forward stages then run the
                v.y = sr - (w.x * di + w.y * dr);        // backward transform
            }
        } else if (valid) {
            // flavors 4-6 (narrow COL tiles): 4 = L2::256B loads + write-back stores,
            // 5 = L2::128B loads + streaming stores, 6 = L2::256B loads + streaming stores
            if (cache & 1) v = ld_plain(gin + (int64_t)(j + r * TPX) * is2);          // expected in L2
            else if (FLAVOR == 4 || FLAVOR == 6) v = ld_l2pf<256>(gin + (int64_t)(j + r * TPX) * is2);
            else if (FLAVOR == 5) v = ld_l2pf<128>(gin + (int64_t)(j + r * TPX) * is2);
            else v = ld_stream(gin + (int64_t)(j + r * TPX) * is2);
        }
        re[r] = swap_in ? v.y : v.x;
        im[r] = swap_in ? v.x : v.y;
        if (FLAVOR == 7) {          // chirp on the logical (re, im) value; padding stays zero
            cplx<T> z; z.x = re[r]; z.y = im[r];
            const int k = j + r * TPX;
            if (k < p.n_in) z = cmul(z, ldg_c(reinterpret_cast<const cplx<T> *>(p.aux0) + k));
            re[r] = z.x; im[r] = z.y;
        }
    }
    // per-CTA twiddle tables: requested after the data loads so both are in flight together
    {
        constexpr int TSTEP2 = N / (E * R1);
        for (int idx = tid; idx < Cfg::TW2_ELEMS; idx += Cfg::THREADS)
            tws2[idx] = ldg_c(&tw[TSTEP2 * tw_mult(idx / E) * (idx % E)]);
        for (int idx = tid; idx < Cfg::TW3_ELEMS; idx += Cfg::THREADS)
            tws3[idx] = ldg_c(&tw[tw_mult(idx / TPX) * (idx % TPX)]);
        // visible after the barrier that follows the stage-1 exchange writes
    }
    for (rep = 0; rep < (FLAVOR == 7 ? 2 : 1); ++rep) {
    if (FLAVOR == 7 && rep == 1) {
        __syncthreads();            // first run's last-stage reads are complete
#pragma unroll
        for (int r = 0; r < E; ++r) { re[r] = bre[FLAVOR == 7 ? r : 0]; im[r] = bim[FLAVOR == 7 ? r : 0]; }
    }
    Butterfly<E, T>::run(re, im);
    if (FLAVOR == 9 && !COL) __syncthreads();      // every thread has gathered its inputs from the raw lines
#pragma unroll
    for (int r = 0; r < E; ++r) {
        cplx<T> v; v.x = re[r]; v.y = im[r];
        sm[sidx(j * E + r)] = v;
    }
    __syncthreads();

    // ---- stage 2: radix R1, Ns = E.  Butterflies b = j + i*TPX; k = b % E.
    {
        constexpr int NB = N / R1;            // butterflies per transform
        constexpr int PER = E / R1;           // butterflies per thread
        cplx<T> w[R1];
        // TPX % E == 0 for three-stage sizes: k = j % E for every butterfly of this thread
        if (R2 > 1 || PER == 1) smem_twiddles<R1, T>(tws2, E, j % E, w);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int b = j + i * TPX;
            if (R2 == 1 && PER > 1) smem_twiddles<R1, T>(tws2, E, b % E, w);
#pragma unroll
            for (int r = 0; r < R1; ++r) {
                cplx<T> v = sm[sidx(b + r * NB)];
                if (r > 0) v = cmul(v, w[r]);
                re[i * R1 + r] = v.x; im[i * R1 + r] = v.y;
            }
        }
        if (R2 > 1) __syncthreads();          // all reads done before the buffer is overwritten
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            T xr[R1], xi[R1];
#pragma unroll
            for (int r = 0; r < R1; ++r) { xr[r] = re[i * R1 + r]; xi[r] = im[i * R1 + r]; }
            Butterfly<R1, T>::run(xr, xi);
#pragma unroll
            for (int r = 0; r < R1; ++r) { re[i * R1 + r] = xr[r]; im[i * R1 + r] = xi[r]; }
        }
        if (R2 > 1) {
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int b = j + i * TPX;
                const int k = b % E;
                const int j0 = (b - k) * R1 + k;
#pragma unroll
                for (int r = 0; r < R1; ++r) {
                    cplx<T> v; v.x = re[i * R1 + r]; v.y = im[i * R1 + r];
                    sm[sidx(j0 + r * E)] = v;
                }
            }
            __syncthreads();
        } else {
            // two-stage transform: this was the last stage (Ns = E = N / R1, so k = b)
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int b = j + i * TPX;
#pragma unroll
                for (int r = 0; r < R1; ++r) emit(b + r * E, i + PER * r, re[i * R1 + r], im[i * R1 + r]);
            }
            if (FLAVOR == 7 && rep == 0) continue;
            flush_col();
            return;
        }
    }

    // ---- stage 3: radix R2, Ns = E * R1 (last: k = b, outputs at b + r * Ns).
    // twiddle W_N^(r b) with b = j + i*TPX:  W_N^(r j) * W_E^(r i)  (TPX = N / E)
    if (R2 > 1) {
        constexpr int NS = E * R1;
        constexpr int PER = E / R2;
        constexpr int R2_ = R2 > 1 ? R2 : 2;
        cplx<T> w[R2_];
        smem_twiddles<R2_, T>(tws3, TPX, j, w);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int b = j + i * TPX;
            T xr[R2_], xi[R2_];
#pragma unroll
            for (int r = 0; r < R2; ++r) {
                cplx<T> v = sm[sidx(b + r * NS)];
                if (r > 0) {
                    cplx<T> wr = w[r];
                    if (i > 0) {
                        wr = cmul(wr, unit_root<E, T>((r * i) % E));
                    }
                    v = cmul(v, wr);
                }
                xr[r] = v.x; xi[r] = v.y;
            }
            Butterfly<R2_, T>::run(xr, xi);
#pragma unroll
            for (int r = 0; r < R2; ++r) emit(b + r * NS, i + PER * r, xr[r], xi[r]);
        }
        if (!(FLAVOR == 7 && rep == 0)) flush_col();
    }
    }   // rep
}

// KIND >= 0: the r2r kind is a compile-time constant.  grid_limit > 0 (multi-GPU passes that are bound by
// NVLink, not by the SMs): the grid is smaller than the number of tiles and every CTA loops over tiles,
// so that the pass leaves SMs free for an HBM-bound pass running next to it on another stream.
// (PERSIST is its own instantiation: the tile loop costs the one-tile kernels 25 registers, i.e. a resident CTA.)
template <typename T, int N, int E, int R1, int R2, int TPB, bool COL, int FLAVOR, int KIND = -1, bool PERSIST = false>
__global__ void __launch_bounds__(FastCfg<T, N, E, R1, R2, TPB, COL, FLAVOR>::THREADS, FastCfg<T, N, E, R1, R2, TPB, COL, FLAVOR>::MINB)
fast_kernel(const __grid_constant__ b2d_fft_pass p, int swap_in, int swap_out, long long ntiles)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (!PERSIST) {
        fast_tile<T, N, E, R1, R2, TPB, COL, FLAVOR, KIND>(p, swap_in, swap_out, (int64_t)blockIdx.x, smem_raw);
        return;
    }
    long long tile = blockIdx.x;
    for (;;) {
        fast_tile<T, N, E, R1, R2, TPB, COL, FLAVOR, KIND>(p, swap_in, swap_out, (int64_t)tile, smem_raw);
        tile += gridDim.x;
        if (tile >= ntiles) break;
        __syncthreads();          // the next tile reuses the exchange buffer
    }
}

// ------------------------------------------------------------------ registry
struct FastEntry {
    int prec, n, col, tpb, code;
    int r2r_kind;            // flavour 9: the kind this instantiation is specialised for, -1 = any (runtime switch)
    int persist;             // 1: CTAs loop over the tiles (launched with a grid smaller than the tile count)
    size_t smem;
    int threads;
    void (*launch)(const b2d_fft_pass &, int, int, unsigned, long long, cudaStream_t);
    const void *func;
};

template <typename T, int N, int E, int R1, int R2, int TPB, bool COL, int FLAVOR, int KIND, bool PERSIST>
void launch_one(const b2d_fft_pass &p, int swap_in, int swap_out, unsigned blocks, long long ntiles, cudaStream_t st)
{
    using Cfg = FastCfg<T, N, E, R1, R2, TPB, COL, FLAVOR>;
    fast_kernel<T, N, E, R1, R2, TPB, COL, FLAVOR, KIND, PERSIST><<<blocks, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(p, swap_in, swap_out, ntiles);
}

#define B2_FAST_ENTRY_KP(PREC, T, N, E, R1, R2, TPB, COL, FLAVOR, CODE, KIND, PERSIST)                      \
    { PREC, N, COL, TPB, CODE, KIND, PERSIST, FastCfg<T, N, E, R1, R2, TPB, COL, FLAVOR>::SMEM_BYTES,       \
      FastCfg<T, N, E, R1, R2, TPB, COL, FLAVOR>::THREADS,                                                  \
      &launch_one<T, N, E, R1, R2, TPB, COL, FLAVOR, KIND, (PERSIST != 0)>,                                 \
      (const void *)&fast_kernel<T, N, E, R1, R2, TPB, COL, FLAVOR, KIND, (PERSIST != 0)> }
#define B2_FAST_ENTRY_K(PREC, T, N, E, R1, R2, TPB, COL, FLAVOR, CODE, KIND)                                \
    B2_FAST_ENTRY_KP(PREC, T, N, E, R1, R2, TPB, COL, FLAVOR, CODE, KIND, 0)
#define B2_FAST_ENTRY(PREC, T, N, E, R1, R2, TPB, COL, FLAVOR, CODE)                                        \
    B2_FAST_ENTRY_KP(PREC, T, N, E, R1, R2, TPB, COL, FLAVOR, CODE, -1, 0)

const FastEntry *table(int *count);   // defined in fft_fast_table.cu

void init(int max_smem);
size_t smem_bytes(const b2d_fft_pass &p);
int available(const b2d_fft_pass &p, int code);
int try_launch(const b2d_fft_pass &p, cudaStream_t st);

}  // namespace b2fast
