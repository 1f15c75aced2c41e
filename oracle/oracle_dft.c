/* oracle/oracle_dft.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement, in 80-bit long double, of the transforms the reference
 * library (FFTW 3.3.11, /root/reference) defines for its transform-execution
 * path.  It is the checker for the CUDA product: only tests/, bench.py's
 * cpu_baseline/reference legs and __graft_entry__.smoke() may load it.  The
 * product library never links or calls anything in oracle/.
 *
 * What is restated, with the reference text each part follows:
 *   - complex DFT, unnormalised, sign -1 forward / +1 backward
 *         doc/reference.texi:1876-1905 ("The 1d Discrete Fourier Transform")
 *   - accurate twiddles by octant reduction (so errors are O(eps), not O(n eps))
 *         kernel/trig.c:57-80
 *   - mixed-radix decimation in time, n = r*m       dft/ct.c:34-45
 *   - O(n^2) small-prime DFT                         dft/generic.c:34-91
 *   - Bluestein chirp-z for large primes             dft/bluestein.c:82-128
 *   - separable multi-dimensional transforms         dft/rank-geq2.c:42-52
 *   - r2c / c2r "non-redundant half" layout          doc/reference.texi:1941-1999,
 *                                                    api/rdft2-pad.c:24-39
 *   - halfcomplex layout r0..r(n/2), i((n+1)/2-1)..i1   doc/reference.texi:916-927
 *   - the eleven r2r kinds                           doc/reference.texi:2066-2353,
 *                                                    libbench2/verify-r2r.c:109-172
 *
 * PINNING: tests/test_oracle.py checks this file against (a) the reference
 * itself, compiled codelet-less into oracle/_ref/libfftw3_ref.so and the
 * long-double libfftw3l_ref.so, on seeded inputs; (b) the golden vectors in
 * tests/golden/ that were generated from that reference build by
 * tests/golden/make_golden.py; (c) direct O(n^2) cos/sin definitions.
 *
 * All arrays here are CONTIGUOUS row-major; strided API layouts are mapped to
 * logical arrays by the Python side of the tests.  Complex data is interleaved
 * (re, im) long double.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef long double LD;
typedef struct { LD re, im; } cx;

static const LD K2PI = 6.2831853071795864769252867665590057683943388L;

/* exp(sign * 2 pi i * m / n), argument reduced to the first octant so the
 * libm call always sees |x| <= pi/4 (restates kernel/trig.c:57-80). */
static cx unit_root(long long m, long long n, int sign)
{
    cx w;
    LD c, s, t;
    int neg_s = 0, neg_c = 0, swap = 0;
    m %= n;
    if (m < 0) m += n;
    /* angle = 2 pi m / n, kept as an exact integer fraction throughout */
    if (2 * m > n) { m = n - m; neg_s = 1; }                 /* a -> 2pi - a : [0, pi]   */
    if (4 * m > n) { m = n - 2 * m; n = 2 * n; neg_c = 1; }  /* a -> pi - a  : [0, pi/2] */
    if (8 * m > n) { m = n - 4 * m; n = 4 * n; swap = 1; }   /* a -> pi/2 - a: [0, pi/4] */
    t = K2PI * (LD)m / (LD)n;
    c = cosl(t); s = sinl(t);
    if (swap) { LD u = c; c = s; s = u; }
    if (neg_c) c = -c;
    if (neg_s) s = -s;
    w.re = c;
    w.im = (sign < 0) ? -s : s;
    return w;
}

static inline cx cmul(cx a, cx b)
{
    cx r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r;
}

static long long smallest_factor(long long n)
{
    long long p;
    if (n % 2 == 0) return 2;
    for (p = 3; p * p <= n; p += 2) if (n % p == 0) return p;
    return n;
}

/* ---- power-of-two FFT used inside Bluestein (iterative, in place) ---- */
static void fft_pow2(long long n, cx *a, int sign)
{
    long long i, j, len;
    for (i = 1, j = 0; i < n; ++i) {
        long long bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cx t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (len = 2; len <= n; len <<= 1) {
        long long half = len >> 1, k;
        for (k = 0; k < half; ++k) {
            cx w = unit_root(k, len, sign);
            for (i = k; i < n; i += len) {
                cx u = a[i], v = cmul(a[i + half], w);
                a[i].re = u.re + v.re; a[i].im = u.im + v.im;
                a[i + half].re = u.re - v.re; a[i + half].im = u.im - v.im;
            }
        }
    }
}

/* Bluestein: X_k = conj(c_k) * sum_j (x_j conj(c_j)) c_{k-j}, c_j = exp(sign*pi*i*j^2/n)
 * restated with c_j = exp(-sign... ) so that the product equals exp(sign 2 pi i jk/n)
 * (dft/bluestein.c:82-128 uses the same identity jk = (j^2 + k^2 - (k-j)^2)/2). */
static void dft_bluestein(long long n, const cx *in, long long is, cx *out, int sign)
{
    long long nb = 1, j;
    cx *a, *b, *chirp;
    while (nb < 2 * n - 1) nb <<= 1;
    a = (cx *)calloc((size_t)nb, sizeof(cx));
    b = (cx *)calloc((size_t)nb, sizeof(cx));
    chirp = (cx *)malloc((size_t)n * sizeof(cx));
    for (j = 0; j < n; ++j) {
        /* exp(sign * pi i j^2 / n) = unit_root(j^2 mod 2n, 2n) */
        long long q = (long long)(((__int128)j * j) % (2 * n));
        chirp[j] = unit_root(q, 2 * n, sign);
    }
    for (j = 0; j < n; ++j) a[j] = cmul(in[j * is], chirp[j]);
    b[0].re = chirp[0].re; b[0].im = -chirp[0].im;
    for (j = 1; j < n; ++j) {
        b[j].re = chirp[j].re; b[j].im = -chirp[j].im;
        b[nb - j] = b[j];
    }
    fft_pow2(nb, a, -1);
    fft_pow2(nb, b, -1);
    for (j = 0; j < nb; ++j) a[j] = cmul(a[j], b[j]);
    fft_pow2(nb, a, +1);
    for (j = 0; j < n; ++j) {
        cx t = cmul(a[j], chirp[j]);
        out[j].re = t.re / (LD)nb; out[j].im = t.im / (LD)nb;
    }
    free(a); free(b); free(chirp);
}

/* recursive mixed radix DIT: out (contiguous, length n) = DFT of in[0], in[is], ... */
static void dft_rec(long long n, const cx *in, long long is, cx *out, int sign,
                    const cx *tw, long long tws /* tw[k*tws] = w_n^k */)
{
    long long r, m, j, k, q;
    if (n == 1) { out[0] = in[0]; return; }
    r = smallest_factor(n);
    if (r == n) {
        if (n <= 128) {   /* direct definition, dft/generic.c */
            for (k = 0; k < n; ++k) {
                cx acc = {0, 0};
                for (j = 0; j < n; ++j) {
                    cx t = cmul(in[j * is], tw[((j * k) % n) * tws]);
                    acc.re += t.re; acc.im += t.im;
                }
                out[k] = acc;
            }
        } else {
            dft_bluestein(n, in, is, out, sign);
        }
        return;
    }
    m = n / r;
    /* r sub-transforms of length m over the decimated inputs */
    for (q = 0; q < r; ++q)
        dft_rec(m, in + q * is, is * r, out + q * m, sign, tw, tws * r);
    /* radix-r butterflies with twiddles w_n^{q k} */
    {
        cx *tmp = (cx *)malloc((size_t)r * sizeof(cx));
        for (k = 0; k < m; ++k) {
            for (q = 0; q < r; ++q)
                tmp[q] = cmul(out[q * m + k], tw[((q * k) % n) * tws]);
            for (j = 0; j < r; ++j) {
                cx acc = {0, 0};
                for (q = 0; q < r; ++q) {
                    /* w_r^{jq} = w_n^{m j q} */
                    cx t = cmul(tmp[q], tw[(((j * q) % r) * m) * tws]);
                    acc.re += t.re; acc.im += t.im;
                }
                out[j * m + k] = acc;
            }
        }
        free(tmp);
    }
}

/* one strided 1-D transform, any n, result written back strided (os) */
static void dft_line(long long n, cx *base, long long stride, int sign, const cx *tw,
                     cx *scratch_in, cx *scratch_out)
{
    long long j;
    for (j = 0; j < n; ++j) scratch_in[j] = base[j * stride];
    dft_rec(n, scratch_in, 1, scratch_out, sign, tw, 1);
    for (j = 0; j < n; ++j) base[j * stride] = scratch_out[j];
}

static cx *make_twiddles(long long n, int sign)
{
    long long k;
    cx *tw = (cx *)malloc((size_t)n * sizeof(cx));
    for (k = 0; k < n; ++k) tw[k] = unit_root(k, n, sign);
    return tw;
}

/* In-place separable rank-d complex DFT on `howmany` contiguous row-major
 * arrays of shape n[0..rank-1].  data: interleaved long double (re,im). */
int oracle_dft(int rank, const long long *n, long long howmany, LD *data, int sign)
{
    long long total = 1, d;
    cx *x = (cx *)data;
    int dim;
    for (dim = 0; dim < rank; ++dim) { if (n[dim] <= 0) return -1; total *= n[dim]; }
    for (dim = rank - 1; dim >= 0; --dim) {
        long long len = n[dim], stride = 1, nlines;
        cx *tw;
        for (d = dim + 1; d < rank; ++d) stride *= n[d];
        nlines = howmany * (total / len);
        tw = make_twiddles(len, sign);
#pragma omp parallel
        {
            cx *si = (cx *)malloc((size_t)len * sizeof(cx));
            cx *so = (cx *)malloc((size_t)len * sizeof(cx));
            long long l;
#pragma omp for schedule(static)
            for (l = 0; l < nlines; ++l) {
                /* line l: outer index l / stride, inner index l % stride */
                long long outer = l / stride, inner = l % stride;
                dft_line(len, x + outer * stride * len + inner, stride, sign, tw, si, so);
            }
            free(si); free(so);
        }
        free(tw);
    }
    return 0;
}

/* r2c: real n[0..rank-1] -> complex n[0] x ... x (n[rank-1]/2+1)  (forward, sign -1) */
int oracle_r2c(int rank, const long long *n, long long howmany, const LD *in, LD *out)
{
    long long total = 1, nl, nh, rows, b, i, k;
    LD *full;
    int dim;
    if (rank < 1) return -1;
    for (dim = 0; dim < rank; ++dim) total *= n[dim];
    nl = n[rank - 1]; nh = nl / 2 + 1; rows = total / nl;
    full = (LD *)malloc((size_t)(2 * total * howmany) * sizeof(LD));
    for (i = 0; i < total * howmany; ++i) { full[2 * i] = in[i]; full[2 * i + 1] = 0; }
    oracle_dft(rank, n, howmany, full, -1);
    for (b = 0; b < howmany; ++b)
        for (i = 0; i < rows; ++i)
            for (k = 0; k < nh; ++k) {
                long long src = (b * rows + i) * nl + k, dst = (b * rows + i) * nh + k;
                out[2 * dst] = full[2 * src]; out[2 * dst + 1] = full[2 * src + 1];
            }
    free(full);
    return 0;
}

/* c2r: complex half array -> real, backward (sign +1), unnormalised.  The full
 * Hermitian array is rebuilt from the stored half: X[-k] = conj X[k] taken over
 * all dimensions; entries of the stored half that are their own mirror have
 * their imaginary part ignored (api doc: doc/reference.texi:1968-1999;
 * rdft/rdft2-rdft.c:61-74 never reads them). */
int oracle_c2r(int rank, const long long *n, long long howmany, const LD *in, LD *out)
{
    long long total = 1, nl, nh, rows, b, i, k, d;
    LD *full;
    int dim;
    long long idx[16], midx[16];
    if (rank < 1 || rank > 16) return -1;
    for (dim = 0; dim < rank; ++dim) total *= n[dim];
    nl = n[rank - 1]; nh = nl / 2 + 1; rows = total / nl;
    full = (LD *)malloc((size_t)(2 * total * howmany) * sizeof(LD));
    for (b = 0; b < howmany; ++b)
        for (i = 0; i < rows; ++i) {
            long long rem = i, mrow = 0;
            for (dim = rank - 2; dim >= 0; --dim) { idx[dim] = rem % n[dim]; rem /= n[dim]; }
            for (dim = 0; dim < rank - 1; ++dim) {
                midx[dim] = (n[dim] - idx[dim]) % n[dim];
                mrow = mrow * n[dim] + midx[dim];
            }
            for (k = 0; k < nl; ++k) {
                long long dst = (b * rows + i) * nl + k;
                if (k < nh) {
                    long long src = (b * rows + i) * nh + k;
                    full[2 * dst] = in[2 * src]; full[2 * dst + 1] = in[2 * src + 1];
                    /* self-mirrored entries are real by definition */
                    if (mrow == i && (k == 0 || 2 * k == nl)) full[2 * dst + 1] = 0;
                } else {
                    long long src = (b * rows + mrow) * nh + (nl - k);
                    full[2 * dst] = in[2 * src]; full[2 * dst + 1] = -in[2 * src + 1];
                }
            }
        }
    (void)d;
    oracle_dft(rank, n, howmany, full, +1);
    for (i = 0; i < total * howmany; ++i) out[i] = full[2 * i];
    free(full);
    return 0;
}

/* ---- r2r kinds (public numbering of api/fftw3.h:96-100) ---- */
enum { K_R2HC = 0, K_HC2R, K_DHT, K_REDFT00, K_REDFT01, K_REDFT10, K_REDFT11,
       K_RODFT00, K_RODFT01, K_RODFT10, K_RODFT11 };

/* one 1-D r2r transform of a strided line, through an embedding into a complex
 * DFT of length N (the "logical size" or a multiple of it). */
static int r2r_line(int kind, long long n, LD *base, long long stride)
{
    long long N, j, k;
    cx *y, *Y, *tw;
    switch (kind) {
    case K_R2HC: case K_HC2R: case K_DHT: N = n; break;
    case K_REDFT00: if (n < 2) return -1; N = 2 * (n - 1); break;
    case K_RODFT00: N = 2 * (n + 1); break;
    case K_REDFT01: case K_REDFT10: case K_RODFT01: case K_RODFT10: N = 4 * n; break;
    case K_REDFT11: case K_RODFT11: N = 8 * n; break;
    default: return -1;
    }
    y = (cx *)calloc((size_t)N, sizeof(cx));
    Y = (cx *)malloc((size_t)N * sizeof(cx));
#define X(j) base[(j) * stride]
    switch (kind) {
    case K_R2HC: case K_DHT:
        for (j = 0; j < n; ++j) y[j].re = X(j);
        break;
    case K_HC2R:
        y[0].re = X(0);
        for (k = 1; 2 * k < n; ++k) {
            y[k].re = X(k); y[k].im = X(n - k);
            y[n - k].re = X(k); y[n - k].im = -X(n - k);
        }
        if (n % 2 == 0) y[n / 2].re = X(n / 2);
        break;
    case K_REDFT00:
        y[0].re = X(0); y[n - 1].re = X(n - 1);
        for (j = 1; j < n - 1; ++j) { y[j].re = X(j); y[N - j].re = X(j); }
        break;
    case K_REDFT10: case K_REDFT11:
        for (j = 0; j < n; ++j) { y[2 * j + 1].re = X(j); y[N - (2 * j + 1)].re = X(j); }
        break;
    case K_REDFT01:
        y[0].re = X(0);
        for (j = 1; j < n; ++j) { y[j].re = X(j); y[N - j].re = X(j); }
        break;
    case K_RODFT00:
        for (j = 0; j < n; ++j) { y[j + 1].re = X(j); y[N - (j + 1)].re = -X(j); }
        break;
    case K_RODFT10: case K_RODFT11:
        for (j = 0; j < n; ++j) { y[2 * j + 1].re = X(j); y[N - (2 * j + 1)].re = -X(j); }
        break;
    case K_RODFT01:
        for (j = 0; j < n - 1; ++j) { y[j + 1].re = X(j); y[N - (j + 1)].re = -X(j); }
        y[n].re = X(n - 1) / 2; y[3 * n].re = -X(n - 1) / 2;
        break;
    }
    tw = make_twiddles(N, kind == K_HC2R ? +1 : -1);
    dft_rec(N, y, 1, Y, kind == K_HC2R ? +1 : -1, tw, 1);
    free(tw);
    switch (kind) {
    case K_R2HC:
        for (k = 0; 2 * k <= n; ++k) X(k) = Y[k].re;
        for (k = 1; 2 * k < n; ++k) X(n - k) = Y[k].im;
        break;
    case K_HC2R:
        for (j = 0; j < n; ++j) X(j) = Y[j].re;
        break;
    case K_DHT:
        for (k = 0; k < n; ++k) X(k) = Y[k].re - Y[k].im;
        break;
    case K_REDFT00: case K_REDFT10:
        for (k = 0; k < n; ++k) X(k) = Y[k].re;
        break;
    case K_REDFT01: case K_REDFT11:
        for (k = 0; k < n; ++k) X(k) = Y[2 * k + 1].re;
        break;
    case K_RODFT00: case K_RODFT10:
        for (k = 0; k < n; ++k) X(k) = -Y[k + 1].im;
        break;
    case K_RODFT01: case K_RODFT11:
        for (k = 0; k < n; ++k) X(k) = -Y[2 * k + 1].im;
        break;
    }
#undef X
    free(y); free(Y);
    return 0;
}

/* In-place separable r2r on `howmany` contiguous row-major real arrays. */
int oracle_r2r(int rank, const long long *n, const int *kind, long long howmany, LD *data)
{
    long long total = 1, d;
    int dim, err = 0;
    for (dim = 0; dim < rank; ++dim) { if (n[dim] <= 0) return -1; total *= n[dim]; }
    for (dim = rank - 1; dim >= 0; --dim) {
        long long len = n[dim], stride = 1, nlines, l;
        for (d = dim + 1; d < rank; ++d) stride *= n[d];
        nlines = howmany * (total / len);
#pragma omp parallel for schedule(static)
        for (l = 0; l < nlines; ++l) {
            long long outer = l / stride, inner = l % stride;
            if (r2r_line(kind[dim], len, data + outer * stride * len + inner, stride))
                err = -1;
        }
    }
    return err;
}

/* Direct O(n^2) definitions of the 1-D r2r kinds exactly as the reference's
 * manual and verifier state them (doc/reference.texi:2066-2353;
 * libbench2/verify-r2r.c:109-172).  Used to cross-check the embeddings above. */
int oracle_r2r_direct_1d(int kind, long long n, const LD *x, LD *y)
{
    long long j, k;
    const LD PI = K2PI / 2;
    for (k = 0; k < n; ++k) {
        LD acc = 0;
        switch (kind) {
        case K_DHT:
            for (j = 0; j < n; ++j) {
                cx w = unit_root((j * k) % n, n, +1);
                acc += x[j] * (w.re + w.im);
            }
            break;
        case K_REDFT00:
            acc = x[0] + ((k & 1) ? -x[n - 1] : x[n - 1]);
            for (j = 1; j < n - 1; ++j) acc += 2 * x[j] * cosl(PI * j * k / (LD)(n - 1));
            break;
        case K_REDFT10:
            for (j = 0; j < n; ++j) acc += 2 * x[j] * cosl(PI * (j + 0.5L) * k / (LD)n);
            break;
        case K_REDFT01:
            acc = x[0];
            for (j = 1; j < n; ++j) acc += 2 * x[j] * cosl(PI * j * (k + 0.5L) / (LD)n);
            break;
        case K_REDFT11:
            for (j = 0; j < n; ++j) acc += 2 * x[j] * cosl(PI * (j + 0.5L) * (k + 0.5L) / (LD)n);
            break;
        case K_RODFT00:
            for (j = 0; j < n; ++j) acc += 2 * x[j] * sinl(PI * (j + 1) * (k + 1) / (LD)(n + 1));
            break;
        case K_RODFT10:
            for (j = 0; j < n; ++j) acc += 2 * x[j] * sinl(PI * (j + 0.5L) * (k + 1) / (LD)n);
            break;
        case K_RODFT01:
            acc = (k & 1) ? -x[n - 1] : x[n - 1];
            for (j = 0; j < n - 1; ++j) acc += 2 * x[j] * sinl(PI * (j + 1) * (k + 0.5L) / (LD)n);
            break;
        case K_RODFT11:
            for (j = 0; j < n; ++j) acc += 2 * x[j] * sinl(PI * (j + 0.5L) * (k + 0.5L) / (LD)n);
            break;
        default:
            return -1;
        }
        y[k] = acc;
    }
    return 0;
}

/* Direct O(n^2) complex DFT from the definition (doc/reference.texi:1884-1888) */
int oracle_dft_direct_1d(long long n, const LD *in, LD *out, int sign)
{
    long long j, k;
    const cx *x = (const cx *)in;
    cx *y = (cx *)out;
    for (k = 0; k < n; ++k) {
        cx acc = {0, 0};
        for (j = 0; j < n; ++j) {
            cx t = cmul(x[j], unit_root((long long)(((__int128)j * k) % n), n, sign));
            acc.re += t.re; acc.im += t.im;
        }
        y[k] = acc;
    }
    return 0;
}
