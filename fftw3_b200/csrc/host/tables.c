/* tables.c -- device-resident constant tables (twiddles, chirps, split
 * factors), computed on the host in long double with exact-fraction octant
 * reduction and rounded once to the working precision, then cached and
 * refcounted.  Role of the reference's kernel/twiddle.c:124-220 (cache keyed on
 * (n, r, m, instr)) and kernel/trig.c:57-80 (accurate cexp); the tables live in
 * HBM/L2 instead of the CPU cache. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "b2_internal.h"

static b2_table *g_tables = NULL;

int b2_run_contig_fft(int prec, int64_t n, void *dev_inout);   /* planner.c */

void b2_unit_root_ld(int64_t m, int64_t n, long double *c, long double *s)
{
    /* exp(-2 pi i m / n): reduce the exact fraction m/n to the first octant */
    static const long double TWO_PI = 6.2831853071795864769252867665590057683943388L;
    int neg_s = 0, neg_c = 0, swap = 0;
    long double t, cc, ss;
    m %= n;
    if (m < 0) m += n;
    if (2 * m > n) { m = n - m; neg_s = 1; }
    if (4 * m > n) { m = n - 2 * m; n = 2 * n; neg_c = 1; }
    if (8 * m > n) { m = n - 4 * m; n = 4 * n; swap = 1; }
    t = TWO_PI * (long double)m / (long double)n;
    cc = cosl(t); ss = sinl(t);
    if (swap) { long double u = cc; cc = ss; ss = u; }
    if (neg_c) cc = -cc;
    if (neg_s) ss = -ss;
    *c = cc;
    *s = -ss;
}

int b2_is_prime(int64_t n)
{
    int64_t d;
    if (n < 2) return 0;
    for (d = 2; d * d <= n; ++d) if (n % d == 0) return 0;
    return 1;
}

static int64_t powmod(int64_t b, int64_t e, int64_t m)
{
    __int128 r = 1, x = b % m;
    while (e > 0) { if (e & 1) r = r * x % m; x = x * x % m; e >>= 1; }
    return (int64_t)r;
}

int64_t b2_primitive_root(int64_t p)
{
    int64_t fac[64], nf = 0, m = p - 1, d, g;
    for (d = 2; d * d <= m; ++d)
        if (m % d == 0) { fac[nf++] = d; while (m % d == 0) m /= d; }
    if (m > 1) fac[nf++] = m;
    for (g = 2; g < p; ++g) {
        int ok = 1;
        int64_t i;
        for (i = 0; i < nf && ok; ++i) if (powmod(g, (p - 1) / fac[i], p) == 1) ok = 0;
        if (ok) return g;
    }
    return p == 2 ? 1 : 0;
}

static void put(void *host, int prec, int64_t i, long double re, long double im)
{
    if (prec == B2D_F32) { ((float *)host)[2 * i] = (float)re; ((float *)host)[2 * i + 1] = (float)im; }
    else { ((double *)host)[2 * i] = (double)re; ((double *)host)[2 * i + 1] = (double)im; }
}

b2_table *b2_table_get(int prec, int kind, int64_t n, int64_t aux)
{
    b2_table *t;
    int64_t count = 0, i;
    size_t esz = (prec == B2D_F32) ? 8 : 16;
    void *host;
    long double c, s;
    const int device = b2d_current_device();

    for (t = g_tables; t; t = t->next)
        if (t->prec == prec && t->kind == kind && t->n == n && t->aux == aux && t->device == device) { t->refs++; return t; }

    switch (kind) {
    case TAB_TWIDDLE: count = n; break;
    case TAB_CHIRP: count = n; break;
    case TAB_BLUE_B: count = aux; break;
    case TAB_R2C: count = n / 2 + 1; break;
    case TAB_TW4_LO: count = aux; break;
    case TAB_TW4_HI: count = (n + aux - 1) / aux; break;
    case TAB_QUARTER: count = 2 * n; break;
    case TAB_RADER_PERM: count = (int64_t)((2 * (size_t)(n - 1) * sizeof(int) + esz - 1) / esz); break;
    case TAB_RADER_B: count = n - 1; break;
    default: return NULL;
    }
    host = malloc((size_t)(count ? count : 1) * esz);
    if (!host) return NULL;
    switch (kind) {
    case TAB_TWIDDLE: case TAB_R2C: case TAB_TW4_LO:
        for (i = 0; i < count; ++i) { b2_unit_root_ld(i, n, &c, &s); put(host, prec, i, c, s); }
        break;
    case TAB_TW4_HI:
        for (i = 0; i < count; ++i) { b2_unit_root_ld(i * aux, n, &c, &s); put(host, prec, i, c, s); }
        break;
    case TAB_CHIRP:   /* exp(-pi i k^2 / n) = exp(-2 pi i (k^2 mod 2n) / 2n) */
        for (i = 0; i < count; ++i) {
            int64_t q = (int64_t)(((__int128)i * i) % (2 * n));
            b2_unit_root_ld(q, 2 * n, &c, &s); put(host, prec, i, c, s);
        }
        break;
    case TAB_BLUE_B:  /* filter b_j = conj(chirp_j) wrapped around length aux = M; FFT'd on the device below */
        memset(host, 0, (size_t)count * esz);
        for (i = 0; i < n; ++i) {
            int64_t q = (int64_t)(((__int128)i * i) % (2 * n));
            b2_unit_root_ld(q, 2 * n, &c, &s);
            put(host, prec, i, c, -s);
            if (i > 0) put(host, prec, aux - i, c, -s);
        }
        break;
    case TAB_RADER_PERM: case TAB_RADER_B: {
        /* generator powers: perm_in[q] = g^q, perm_out[m] = g^-m = perm_in[(M - m) % M]  (dft/rader.c:95-165) */
        int64_t M = n - 1, g = b2_primitive_root(n), v = 1;
        int *pin = (int *)malloc((size_t)(2 * M) * sizeof(int)), *pout = pin + M;
        if (!pin || g <= 0) { free(pin); free(host); return NULL; }
        for (i = 0; i < M; ++i) { pin[i] = (int)v; v = (int64_t)((__int128)v * g % n); }
        for (i = 0; i < M; ++i) pout[i] = pin[(M - i) % M];
        if (kind == TAB_RADER_PERM) memcpy(host, pin, (size_t)(2 * M) * sizeof(int));
        else for (i = 0; i < M; ++i) { b2_unit_root_ld(pout[i], n, &c, &s); put(host, prec, i, c, s); }
        free(pin);
        break;
    }
    case TAB_QUARTER:
        for (i = 0; i < n; ++i) {
            b2_unit_root_ld(i, 4 * n, &c, &s); put(host, prec, i, c, s);              /* exp(-pi i k/(2n)) */
            b2_unit_root_ld(2 * i + 1, 8 * n, &c, &s); put(host, prec, n + i, c, s);  /* exp(-pi i (2k+1)/(4n)) */
        }
        break;
    }
    t = (b2_table *)calloc(1, sizeof *t);
    if (!t) { free(host); return NULL; }
    t->prec = prec; t->kind = kind; t->n = n; t->aux = aux; t->device = device;
    t->bytes = (size_t)(count ? count : 1) * esz;
    t->dev = b2d_malloc(t->bytes);
    if (!t->dev || b2d_memcpy_h2d(t->dev, host, (size_t)count * esz) || b2d_sync()) {
        b2d_free(t->dev); free(t); free(host);
        return NULL;
    }
    free(host);
    if (kind == TAB_RADER_B && b2_run_contig_fft(prec, n - 1, t->dev)) {
        b2d_free(t->dev); free(t);
        return NULL;
    }
    if (kind == TAB_BLUE_B && b2_run_contig_fft(prec, aux, t->dev)) {
        b2d_free(t->dev); free(t);
        return NULL;
    }
    t->refs = 1;
    t->next = g_tables;
    g_tables = t;
    return t;
}

void b2_table_release(b2_table *t)
{
    b2_table **pp;
    if (!t || --t->refs > 0) return;
    for (pp = &g_tables; *pp; pp = &(*pp)->next)
        if (*pp == t) { *pp = t->next; break; }
    b2d_free(t->dev);
    free(t);
}

void b2_tables_cleanup(void)
{
    /* tables still referenced by live plans stay valid (plans own a reference) */
}
