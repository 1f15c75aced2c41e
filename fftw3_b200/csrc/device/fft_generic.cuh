// fft_generic.cuh -- the any-size, any-layout batched 1-D Stockham pass.
//
// One CTA stages `tpb` transforms of length n in shared memory, runs all radix
// stages there (runtime radix list, straight-line register butterflies from
// butterflies_gen.cuh) and streams the result back: the array crosses HBM once
// per pass.  This is the GPU counterpart of the reference's Cooley-Tukey solver
// plus codelet drivers (dft/ct.c:34-58, dft/dftw-direct.c:46-56,
// dft/direct.c:92-97) and of its buffered/strided helpers (dft/buffered.c:41-69,
// dft/vrank-geq1.c:54-65): the batch loop is the grid, the "buffer" is shared
// memory, the codelets are the generated butterflies.
//
// The kernel body is written as per-thread *phase* functions separated by
// barriers.  Each phase is __host__ __device__ so that tests/emu can run the very
// same index algebra on the CPU (threads as a loop) -- a unit-test double for
// the host-side planner, never part of the product library.
#pragma once
#include <stdint.h>
#include "../../../include/b200fft_device.h"
#include "butterflies_gen.cuh"

#ifdef __CUDACC__
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b2 {

template <typename T> struct alignas(2 * sizeof(T)) cplx { T x, y; };

template <typename T> B2_HD cplx<T> cmul(cplx<T> a, cplx<T> b)
{
    cplx<T> r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}

}  // namespace b2
#include "r2r_maps.cuh"
namespace b2 {

// padded position of element k inside one transform's shared-memory row
B2_HD int padk(int k) { return k + (k >> 4); }
// row pitch (complex elements): padded length forced to 1 mod 8 so that COL-mode
// accesses (lanes walk transforms) fall in distinct 16-byte bank groups
B2_HD int row_pitch(int n)
{
    int p = padk(n - 1) + 1;
    while ((p & 7) != 1) ++p;
    return p;
}

struct TileCtx {            // per-CTA derived values
    int64_t tile0, b1, b2;
};

B2_HD TileCtx decode_block(const b2d_fft_pass &p, int64_t block)
{
    TileCtx c;
    int64_t tiles0 = (p.bn[0] + p.tpb - 1) / p.tpb;
    if (p.npeer && !p.peer_rows) {
        // peer scatter: the destination (batch dim 2) varies fastest and starts at this rank's
        // successor -- all NVLink ports busy, no receiver hot-spot
        c.b2 = (block % p.bn[2] + p.peer_rot) % p.bn[2];
        int64_t rest = block / p.bn[2];
        c.tile0 = rest % tiles0;
        c.b1 = rest / tiles0;
        return c;
    }
    c.tile0 = block % tiles0;
    int64_t rest = block / tiles0;
    c.b1 = rest % p.bn[1];
    c.b2 = rest / p.bn[1];
    return c;
}

B2_HD int64_t grid_blocks(const b2d_fft_pass &p)
{
    int64_t tiles0 = (p.bn[0] + p.tpb - 1) / p.tpb;
    return tiles0 * p.bn[1] * p.bn[2];
}

// ---------------------------------------------------------------- load element
template <typename T>
B2_HD cplx<T> load_elem(const b2d_fft_pass &p, int64_t boff, int64_t b0, int k)
{
    const T *re = (const T *)p.in_re;
    const T *im = (const T *)p.in_im;
    const int op = p.pre_op;
    if (op & B2D_LOAD_R2R) {        // real line -> work sequence of the r2r kind (r2r_maps.cuh)
        RealLineIn<T> x = { (const T *)p.in_re + boff, p.is };
        cplx<T> u = r2r_pre_value<T>(p.r2r_kind, p.n_in, k, (const cplx<T> *)p.aux0, x);
        if (p.r2r_pair) {           // second line of the pair rides in the imaginary part
            RealLineIn<T> x2 = { (const T *)p.in_re + boff + p.pair_is, p.is };
            u.y = r2r_pre_value<T>(p.r2r_kind, p.n_in, k, (const cplx<T> *)p.aux0, x2).x;
        }
        return u;
    }
    cplx<T> z;
    if (op & B2D_LOAD_C2R_MERGE) {
        // logical index j of this element in the half-size line; its mirror m - j sits (m - 2j) logical steps on
        const int64_t m = p.n_in;
        const int64_t j = p.idx_mul ? (int64_t)k * p.idx_mul + b0 : (int64_t)k;
        const int64_t o = boff + (int64_t)k * p.is;
        const int64_t om = o + (m - 2 * j) * (p.idx_mul ? p.bis[0] : p.is);
        cplx<T> a, c;
        a.x = re[o]; a.y = im[o]; c.x = re[om]; c.y = im[om];
        if (j == 0) { a.y = T(0); c.y = T(0); }
        const T sr = a.x + c.x, si = a.y - c.y, dr = a.x - c.x, di = a.y + c.y;
        cplx<T> w = ((const cplx<T> *)p.aux2)[j]; w.y = -w.y;        // conj(w^j)
        z.x = si + (w.x * dr - w.y * di);                            // swapped: (Im Z_j, Re Z_j)
        z.y = sr - (w.x * di + w.y * dr);
        return z;
    }
    if (op & B2D_LOAD_RADER) k = ((const int *)p.aux0)[k];          // a_q = x[g^q mod n]
    if ((op & B2D_LOAD_PAD) && k >= p.n_in) { z.x = T(0); z.y = T(0); return z; }
    if (op & B2D_LOAD_REAL) {
        z.x = re[boff + (int64_t)k * p.is]; z.y = T(0);
    } else if (op & B2D_LOAD_HERMCONJ) {
        // conj(H) of the Hermitian sequence of logical length n_in whose
        // non-redundant half is stored: Re(forward(conj H)) == backward(H)
        const int64_t n = p.n_in;
        if (p.idx_mul) {
            // half of a four-step line: logical index j = k * idx_mul + b0, the line advances by bis[0] per index
            const int64_t j = (int64_t)k * p.idx_mul + b0;
            const int64_t o = boff + (int64_t)k * p.is;
            if (2 * j <= n) { z.x = re[o]; z.y = (j == 0 || 2 * j == n) ? T(0) : -im[o]; }
            else { const int64_t om = o + (n - 2 * j) * p.bis[0]; z.x = re[om]; z.y = im[om]; }
        } else if (2 * k <= n) {
            int64_t o = boff + (int64_t)k * p.is;
            z.x = re[o];
            z.y = (k == 0 || 2 * k == n) ? T(0) : -im[o];
        } else {
            int64_t o = boff + (int64_t)(n - k) * p.is;
            z.x = re[o]; z.y = im[o];
        }
    } else {
        int64_t o = boff + (int64_t)k * p.is;
        z.x = re[o]; z.y = im[o];
    }
    if (op & B2D_LOAD_CHIRP) z = cmul(z, ((const cplx<T> *)p.aux0)[k]);
    return z;
}

// --------------------------------------------------------------- store element
template <typename T>
B2_HD void store_elem(const b2d_fft_pass &p, int64_t boff, int64_t b0, int64_t peer, int k, cplx<T> z)
{
    T *re = (T *)p.out_re;
    T *im = (T *)p.out_im;
    if (p.npeer) {            // batch dim 2 (or the output row) selects the destination buffer (peer GPU), interleaved
        if (p.peer_rows) { peer = k / p.peer_rows; k -= (int)(peer * p.peer_rows); }
        T *base = (T *)p.peer_out[peer];
        if ((const char *)p.out_im < (const char *)p.out_re) { im = base; re = base + 1; }   // backward: swapped
        else { re = base; im = base + 1; }
    }
    const int op = p.post_op;
    if (op & B2D_STORE_R2R) {       // FFT output k scattered to the real output line
        RealLineOut<T> y = { (T *)p.out_re + boff, p.os };
        r2r_post_scatter<T>(p.r2r_kind, p.n_out, k, z, (const cplx<T> *)p.aux0, y);
        return;
    }
    if ((op & B2D_STORE_TRUNC) && (p.idx_mul ? (int64_t)k * p.idx_mul + b0 : (int64_t)k) >= p.n_out) return;
    if (op & B2D_STORE_TWIDDLE4) {
        // exponent e = k * b0 < big_n ; W^e = hi[e / L] * lo[e % L]
        int64_t e = ((int64_t)k * (b0 + p.tw4_off)) % p.big_n;
        int64_t eh = e / p.aux_split, el = e - eh * p.aux_split;
        cplx<T> w = cmul(((const cplx<T> *)p.aux1)[eh], ((const cplx<T> *)p.aux0)[el]);
        z = cmul(z, w);
    }
    if (op & B2D_STORE_RADER) {     // X[g^-m mod n] = x_0 + conv_m (x_0 was folded into bin 0 before the inverse)
        k = ((const int *)p.aux0)[p.n + k];
        z.x *= (T)p.scale; z.y *= (T)p.scale;
    }
    if (op & B2D_STORE_CHIRP_SCALE) {
        z = cmul(z, ((const cplx<T> *)p.aux0)[k]);
        z.x *= (T)p.scale; z.y *= (T)p.scale;
    }
    int64_t o = boff + (int64_t)k * p.os;
    re[o] = z.x;
    if (!(op & B2D_STORE_REALPART)) im[o] = z.y;
}

// shared-memory carve-up: [boff_in tpb][boff_out tpb][b0 tpb] int64, then 2 complex buffers
template <typename T> struct Smem {
    int64_t *boff_in, *boff_out, *b0;
    cplx<T> *a, *b;
    cplx<T> *tw;              // twiddle table copied into shared memory (NULL: read it from global)
    int pitch;
};

template <typename T>
B2_HD Smem<T> carve(const b2d_fft_pass &p, unsigned char *raw)
{
    Smem<T> s;
    s.boff_in = (int64_t *)raw;
    s.boff_out = s.boff_in + p.tpb;
    s.b0 = s.boff_out + p.tpb;
    size_t off = (size_t)3 * p.tpb * sizeof(int64_t);
    off = (off + 15) & ~(size_t)15;
    s.pitch = row_pitch(p.n);
    s.a = (cplx<T> *)(raw + off);
    s.b = s.a + (size_t)p.tpb * s.pitch;
    s.tw = p.tw_smem ? s.b + (size_t)p.tpb * s.pitch : (cplx<T> *)0;
    return s;
}

template <typename T>
inline size_t smem_bytes(const b2d_fft_pass &p)
{
    size_t off = (size_t)3 * p.tpb * sizeof(int64_t);
    off = (off + 15) & ~(size_t)15;
    return off + ((size_t)2 * p.tpb * row_pitch(p.n) + (p.tw_smem ? (size_t)p.n : 0)) * sizeof(cplx<T>);
}

// phase: copy the twiddle table into shared memory (when the planner found room for it)
template <typename T>
B2_HD void phase_twiddles(const b2d_fft_pass &p, const Smem<T> &s, int tid, int nthreads)
{
    if (!s.tw) return;
    const cplx<T> *g = (const cplx<T> *)p.tw;
    for (int i = tid; i < p.n; i += nthreads) s.tw[i] = g[i];
}

// phase 0: per-transform base offsets (threads 0..tpb-1)
template <typename T>
B2_HD void phase_offsets(const b2d_fft_pass &p, const Smem<T> &s, const TileCtx &c, int tid)
{
    if (tid < p.tpb) {
        int64_t b0 = c.tile0 * p.tpb + tid;
        s.b0[tid] = (b0 < p.bn[0]) ? b0 : -1;
        s.boff_in[tid] = b0 * p.bis[0] + c.b1 * p.bis[1] + c.b2 * p.bis[2];
        s.boff_out[tid] = b0 * p.bos[0] + c.b1 * p.bos[1] + ((p.npeer && !p.peer_rows) ? 0 : c.b2 * p.bos[2]);
    }
}

// Plain c2c fast path of the generic kernel: interleaved complex on both sides, no fused ops.
// Decided per launch (pointer deltas and alignment are run-time facts).
B2_HD bool plain_ok(const b2d_fft_pass &p, int *swap_in, int *swap_out)
{
    const int64_t rs = p.prec == B2D_F32 ? 4 : 8;
    if (p.pre_op || p.post_op || p.bluestein || p.npeer) return false;
    const int64_t din = (const char *)p.in_im - (const char *)p.in_re;
    const int64_t dout = (const char *)p.out_im - (const char *)p.out_re;
    if ((din != rs && din != -rs) || (dout != rs && dout != -rs)) return false;
    *swap_in = din < 0; *swap_out = dout < 0;
    if (((uintptr_t)(din < 0 ? p.in_im : p.in_re) % (2 * rs)) || ((uintptr_t)(dout < 0 ? p.out_im : p.out_re) % (2 * rs))) return false;
    if ((p.is & 1) || (p.os & 1)) return false;
    for (int i = 0; i < B2D_MAX_BATCH_DIMS; ++i) if ((p.bis[i] & 1) || (p.bos[i] & 1)) return false;
    return true;
}

// vector load / store phases of the plain path: one 128-bit (f64) access per element and no
// per-element division (lanes walk k in ROW mode, t in COL mode)
template <typename T>
B2_HD void phase_load_plain(const b2d_fft_pass &p, const Smem<T> &s, int tid, int nthreads, int swap)
{
    const int n = p.n, tpb = p.tpb;
    const cplx<T> *g = (const cplx<T> *)(swap ? p.in_im : p.in_re);
    const int64_t is2 = p.is / 2;
    if (!p.load_col || (nthreads % tpb) != 0) {
        for (int t = 0; t < tpb; ++t) {
            if (s.b0[t] < 0) continue;
            const cplx<T> *gl = g + s.boff_in[t] / 2;
            cplx<T> *row = s.a + (size_t)t * s.pitch;
            for (int k = tid; k < n; k += nthreads) {
                cplx<T> v = gl[(int64_t)k * is2];
                if (swap) { T u = v.x; v.x = v.y; v.y = u; }
                row[padk(k)] = v;
            }
        }
    } else {
        const int t = tid % tpb, step = nthreads / tpb;
        if (s.b0[t] < 0) return;
        const cplx<T> *gl = g + s.boff_in[t] / 2;
        cplx<T> *row = s.a + (size_t)t * s.pitch;
        for (int k = tid / tpb; k < n; k += step) {
            cplx<T> v = gl[(int64_t)k * is2];
            if (swap) { T u = v.x; v.x = v.y; v.y = u; }
            row[padk(k)] = v;
        }
    }
}

template <typename T>
B2_HD void phase_store_plain(const b2d_fft_pass &p, const Smem<T> &s, const cplx<T> *src, int tid, int nthreads, int swap)
{
    const int n = p.n, tpb = p.tpb;
    cplx<T> *g = (cplx<T> *)(swap ? p.out_im : p.out_re);
    const int64_t os2 = p.os / 2;
    if (!p.store_col || (nthreads % tpb) != 0) {
        for (int t = 0; t < tpb; ++t) {
            if (s.b0[t] < 0) continue;
            cplx<T> *gl = g + s.boff_out[t] / 2;
            const cplx<T> *row = src + (size_t)t * s.pitch;
            for (int k = tid; k < n; k += nthreads) {
                cplx<T> v = row[padk(k)];
                if (swap) { T u = v.x; v.x = v.y; v.y = u; }
                gl[(int64_t)k * os2] = v;
            }
        }
    } else {
        const int t = tid % tpb, step = nthreads / tpb;
        if (s.b0[t] < 0) return;
        cplx<T> *gl = g + s.boff_out[t] / 2;
        const cplx<T> *row = src + (size_t)t * s.pitch;
        for (int k = tid / tpb; k < n; k += step) {
            cplx<T> v = row[padk(k)];
            if (swap) { T u = v.x; v.x = v.y; v.y = u; }
            gl[(int64_t)k * os2] = v;
        }
    }
}

// phase 1: global -> shared (buffer a)
template <typename T>
B2_HD void phase_load(const b2d_fft_pass &p, const Smem<T> &s, int tid, int nthreads)
{
    const int n = p.n, tpb = p.tpb;
    const int total = n * tpb;
    for (int idx = tid; idx < total; idx += nthreads) {
        int t, k;
        if (p.load_col) { k = idx / tpb; t = idx - k * tpb; }
        else            { t = idx / n;   k = idx - t * n; }
        if (s.b0[t] < 0) continue;
        s.a[(size_t)t * s.pitch + padk(k)] = load_elem<T>(p, s.boff_in[t], s.b0[t], k);
    }
}

// one Stockham stage, radix R, src -> dst
template <int R, typename T>
B2_HD void stage_radix(const b2d_fft_pass &p, const cplx<T> *src, cplx<T> *dst, int pitch,
                       int ns, int tid, int nthreads, const cplx<T> *tw)
{
    const int n = p.n;
    const int nb = n / R;                 // butterflies per transform
    const int tstep = n / (ns * R);       // W_n^(tstep * r * k) = exp(-2 pi i r k / (ns R))
    const bool ns_pow2 = (ns & (ns - 1)) == 0;
    // thread -> (transform t, first butterfly jj): one division per stage, not per butterfly
    const int tpx = nthreads / p.tpb > 0 ? nthreads / p.tpb : 1;
    const int t = tid / tpx;
    const int jj = tid - t * tpx;
    if (t >= p.tpb) return;
    const cplx<T> *x = src + (size_t)t * pitch;
    cplx<T> *y = dst + (size_t)t * pitch;
    for (int j = jj; j < nb; j += tpx) {
        const int k = ns_pow2 ? (j & (ns - 1)) : (j % ns);
        T re[R], im[R];
        cplx<T> w[R];
        if (ns > 1) {
            // two-level twiddles: W^(rk) = W^(4a k) W^(c k), r = 4a + c  (R/4 + 2 table reads, not R - 1)
            const int base = tstep * k;
            if (R <= 4) {
#pragma unroll
                for (int r = 1; r < R; ++r) w[r] = tw[base * r];
            } else {
#pragma unroll
                for (int c = 1; c < 4 && c < R; ++c) w[c] = tw[base * c];
#pragma unroll
                for (int a = 1; a <= (R - 1) / 4; ++a) {
                    w[4 * a] = tw[base * 4 * a];
#pragma unroll
                    for (int c = 1; c < 4; ++c)
                        if (4 * a + c < R) w[4 * a + c] = cmul(w[4 * a], w[c]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cplx<T> v = x[padk(j + r * nb)];
            if (r > 0 && ns > 1) v = cmul(v, w[r]);
            re[r] = v.x; im[r] = v.y;
        }
        Butterfly<R, T>::run(re, im);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            cplx<T> v; v.x = re[r]; v.y = im[r];
            y[padk(j0 + r * ns)] = v;
        }
    }
}

template <typename T>
B2_HD void phase_stage(const b2d_fft_pass &p, int stage, int ns, const cplx<T> *src, cplx<T> *dst,
                       int pitch, int tid, int nthreads, const cplx<T> *tw)
{
    switch (p.radix[stage]) {
    case 2:  stage_radix<2, T>(p, src, dst, pitch, ns, tid, nthreads, tw); break;
    case 3:  stage_radix<3, T>(p, src, dst, pitch, ns, tid, nthreads, tw); break;
    case 4:  stage_radix<4, T>(p, src, dst, pitch, ns, tid, nthreads, tw); break;
    case 5:  stage_radix<5, T>(p, src, dst, pitch, ns, tid, nthreads, tw); break;
    case 6:  stage_radix<6, T>(p, src, dst, pitch, ns, tid, nthreads, tw); break;
    case 7:  stage_radix<7, T>(p, src, dst, pitch, ns, tid, nthreads, tw); break;
    case 8:  stage_radix<8, T>(p, src, dst, pitch, ns, tid, nthreads, tw); break;
    case 9:  stage_radix<9, T>(p, src, dst, pitch, ns, tid, nthreads, tw); break;
    case 10: stage_radix<10, T>(p, src, dst, pitch, ns, tid, nthreads, tw); break;
    case 11: stage_radix<11, T>(p, src, dst, pitch, ns, tid, nthreads, tw); break;
    case 12: stage_radix<12, T>(p, src, dst, pitch, ns, tid, nthreads, tw); break;
    case 13: stage_radix<13, T>(p, src, dst, pitch, ns, tid, nthreads, tw); break;
    case 16: stage_radix<16, T>(p, src, dst, pitch, ns, tid, nthreads, tw); break;
    default: break;
    }
}

// Bluestein mid-phase: z[k] = conj(z[k] * B[k])  (the conj turns the second
// forward run into an inverse transform; undone in phase_store)
template <typename T>
B2_HD void phase_pointwise(const b2d_fft_pass &p, const Smem<T> &s, cplx<T> *buf, int pitch, int tid, int nthreads)
{
    const int n = p.n;
    const int total = n * p.tpb;
    const cplx<T> *B = (const cplx<T> *)p.aux1;
    for (int idx = tid; idx < total; idx += nthreads) {
        int t = idx / n, k = idx - t * n;
        if (s.b0[t] < 0) continue;
        cplx<T> *q = buf + (size_t)t * pitch + padk(k);
        cplx<T> a = *q;
        cplx<T> v = cmul(a, B[k]);
        if (p.bluestein == 2 && k == 0) {
            // Rader (dft/rader.c:95-165): bin 0 of the permuted input's transform is the sum of x_1..x_{n-1}:
            // X_0 = x_0 + A_0 goes straight out; x_0 is added to every convolution output by adding
            // x_0 * M to bin 0 before the inverse (the store scales by 1/M)
            const T *re = (const T *)p.in_re, *im = (const T *)p.in_im;
            T *ore = (T *)p.out_re, *oim = (T *)p.out_im;
            cplx<T> x0; x0.x = re[s.boff_in[t]]; x0.y = im[s.boff_in[t]];
            ore[s.boff_out[t]] = x0.x + a.x; oim[s.boff_out[t]] = x0.y + a.y;
            v.x += x0.x * (T)n; v.y += x0.y * (T)n;
        }
        v.y = -v.y;
        *q = v;
    }
}

// final phase: shared -> global
template <typename T>
B2_HD void phase_store(const b2d_fft_pass &p, const Smem<T> &s, const TileCtx &c, const cplx<T> *src, int tid,
                       int nthreads)
{
    const int n = p.n, tpb = p.tpb;
    const int total = n * tpb;
    for (int idx = tid; idx < total; idx += nthreads) {
        int t, k;
        if (p.store_col) { k = idx / tpb; t = idx - k * tpb; }
        else             { t = idx / n;   k = idx - t * n; }
        if (s.b0[t] < 0) continue;
        cplx<T> z = src[(size_t)t * s.pitch + padk(k)];
        if (p.bluestein) z.y = -z.y;
        if (p.r2r_pair) {           // two real lines per transform: separate their spectra, POST each
            cplx<T> zm = src[(size_t)t * s.pitch + padk(k ? n - k : 0)], u, v;
            r2r_unpack_pair<T>(z, zm, u, v);
            RealLineOut<T> ya = { (T *)p.out_re + s.boff_out[t], p.os };
            RealLineOut<T> yb = { (T *)p.out_re + s.boff_out[t] + p.pair_os, p.os };
            r2r_post_scatter<T>(p.r2r_kind, p.n_out, k, u, (const cplx<T> *)p.aux0, ya);
            r2r_post_scatter<T>(p.r2r_kind, p.n_out, k, v, (const cplx<T> *)p.aux0, yb);
            continue;
        }
        store_elem<T>(p, s.boff_out[t], s.b0[t], c.b2, k, z);
    }
}

#ifdef __CUDACC__
template <typename T, bool PLAIN>
__global__ void fft_generic_kernel(const __grid_constant__ b2d_fft_pass p, int swap_in, int swap_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const Smem<T> s = carve<T>(p, smem_raw);
    const TileCtx c = decode_block(p, (int64_t)blockIdx.x);
    phase_offsets<T>(p, s, c, tid);
    phase_twiddles<T>(p, s, tid, nthreads);
    const cplx<T> *twp = s.tw ? s.tw : (const cplx<T> *)p.tw;
    __syncthreads();
    if (PLAIN) phase_load_plain<T>(p, s, tid, nthreads, swap_in);
    else phase_load<T>(p, s, tid, nthreads);
    __syncthreads();
    cplx<T> *src = s.a, *dst = s.b;
    const int reps = p.bluestein ? 2 : 1;
    for (int rep = 0; rep < reps; ++rep) {
        int ns = 1;
        for (int st = 0; st < p.nstages; ++st) {
            phase_stage<T>(p, st, ns, src, dst, s.pitch, tid, nthreads, twp);
            __syncthreads();
            ns *= p.radix[st];
            cplx<T> *tmp = src; src = dst; dst = tmp;
        }
        if (p.bluestein && rep == 0) {
            phase_pointwise<T>(p, s, src, s.pitch, tid, nthreads);
            __syncthreads();
        }
    }
    if (PLAIN) phase_store_plain<T>(p, s, src, tid, nthreads, swap_out);
    else phase_store<T>(p, s, c, src, tid, nthreads);
}
#endif

}  // namespace b2
