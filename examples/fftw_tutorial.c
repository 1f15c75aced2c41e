/* A program written against plain FFTW 3 -- nothing here knows about the GPU engine.  It follows the
 * shape of the examples in the reference manual (doc/tutorial.texi: "Complex One-Dimensional DFTs",
 * "One-Dimensional DFTs of Real Data", "More DFTs of Real Data") and checks each result against a
 * direct O(n^2) evaluation of the definition (doc/reference.texi:1872-1905, 2060-2100).
 *
 *   gcc examples/fftw_tutorial.c -Iinclude -Lfftw3_b200/lib -lfftw3_b200 -lm -o tutorial
 * or link it against the real FFTW (-lfftw3): the source does not change.
 * Prints one line per check and exits non-zero if any fails (or if planning fails, e.g. no GPU). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <fftw3.h>

static double frand(void) { return (double)rand() / RAND_MAX - 0.5; }

static int check(const char *what, double err, double tol)
{
    printf("%-34s max error %.3e %s\n", what, err, err <= tol ? "ok" : "FAILED");
    return err <= tol ? 0 : 1;
}

int main(void)
{
    const int n = 96, n0 = 12, n1 = 10;
    const double pi = 3.14159265358979323846;
    int bad = 0, i, j, k;
    srand(7);

    /* complex 1-d, forward then backward */
    {
        fftw_complex *in = fftw_alloc_complex(n), *out = fftw_alloc_complex(n), *back = fftw_alloc_complex(n);
        fftw_plan p = fftw_plan_dft_1d(n, in, out, FFTW_FORWARD, FFTW_ESTIMATE);
        fftw_plan q = fftw_plan_dft_1d(n, out, back, FFTW_BACKWARD, FFTW_ESTIMATE);
        double err = 0, err2 = 0;
        if (!p || !q) { fprintf(stderr, "planning failed (no CUDA device?)\n"); return 2; }
        for (i = 0; i < n; ++i) { in[i][0] = frand(); in[i][1] = frand(); }
        fftw_execute(p);
        fftw_execute(q);
        for (k = 0; k < n; ++k) {
            double re = 0, im = 0;
            for (j = 0; j < n; ++j) {
                double a = -2 * pi * (double)((j * k) % n) / n;
                re += in[j][0] * cos(a) - in[j][1] * sin(a);
                im += in[j][0] * sin(a) + in[j][1] * cos(a);
            }
            err = fmax(err, fmax(fabs(out[k][0] - re), fabs(out[k][1] - im)));
            err2 = fmax(err2, fmax(fabs(back[k][0] / n - in[k][0]), fabs(back[k][1] / n - in[k][1])));
        }
        bad += check("fftw_plan_dft_1d forward", err, 1e-12);
        bad += check("backward(forward(x)) / n", err2, 1e-12);
        fftw_destroy_plan(p); fftw_destroy_plan(q);
        fftw_free(in); fftw_free(out); fftw_free(back);
    }

    /* real 2-d r2c / c2r, in place with the padded layout */
    {
        const int h = n1 / 2 + 1;
        double *a = fftw_alloc_real((size_t)n0 * 2 * h), *ref = (double *)malloc(sizeof(double) * n0 * n1);
        fftw_complex *c = (fftw_complex *)a;
        fftw_plan p = fftw_plan_dft_r2c_2d(n0, n1, a, c, FFTW_ESTIMATE);
        fftw_plan q = fftw_plan_dft_c2r_2d(n0, n1, c, a, FFTW_ESTIMATE);
        double err = 0, err2 = 0;
        if (!p || !q) { fprintf(stderr, "planning failed\n"); return 2; }
        for (i = 0; i < n0; ++i) for (j = 0; j < n1; ++j) ref[i * n1 + j] = a[i * 2 * h + j] = frand();
        fftw_execute(p);
        for (i = 0; i < n0; ++i) for (k = 0; k < h; ++k) {
            double re = 0, im = 0;
            int u, v;
            for (u = 0; u < n0; ++u) for (v = 0; v < n1; ++v) {
                double ang = -2 * pi * ((double)((u * i) % n0) / n0 + (double)((v * k) % n1) / n1);
                re += ref[u * n1 + v] * cos(ang); im += ref[u * n1 + v] * sin(ang);
            }
            err = fmax(err, fmax(fabs(c[i * h + k][0] - re), fabs(c[i * h + k][1] - im)));
        }
        fftw_execute(q);
        for (i = 0; i < n0; ++i) for (j = 0; j < n1; ++j)
            err2 = fmax(err2, fabs(a[i * 2 * h + j] / (n0 * n1) - ref[i * n1 + j]));
        bad += check("fftw_plan_dft_r2c_2d (in place)", err, 1e-12);
        bad += check("c2r(r2c(x)) / (n0 n1)", err2, 1e-12);
        fftw_destroy_plan(p); fftw_destroy_plan(q);
        fftw_free(a); free(ref);
    }

    /* DCT-II (REDFT10) and its inverse DCT-III (REDFT01) */
    {
        double *x = fftw_alloc_real(n), *y = fftw_alloc_real(n), *z = fftw_alloc_real(n);
        fftw_plan p = fftw_plan_r2r_1d(n, x, y, FFTW_REDFT10, FFTW_ESTIMATE);
        fftw_plan q = fftw_plan_r2r_1d(n, y, z, FFTW_REDFT01, FFTW_ESTIMATE);
        double err = 0, err2 = 0;
        if (!p || !q) { fprintf(stderr, "planning failed\n"); return 2; }
        for (i = 0; i < n; ++i) x[i] = frand();
        fftw_execute(p);
        fftw_execute(q);
        for (k = 0; k < n; ++k) {
            double s = 0;
            for (j = 0; j < n; ++j) s += 2 * x[j] * cos(pi * (j + 0.5) * k / n);
            err = fmax(err, fabs(y[k] - s));
            err2 = fmax(err2, fabs(z[k] / (2.0 * n) - x[k]));
        }
        bad += check("fftw_plan_r2r_1d REDFT10", err, 1e-12);
        bad += check("REDFT01(REDFT10(x)) / 2n", err2, 1e-12);
        fftw_destroy_plan(p); fftw_destroy_plan(q);
        fftw_free(x); fftw_free(y); fftw_free(z);
    }

    /* wisdom survives a forget/import round trip */
    {
        char *w = fftw_export_wisdom_to_string();
        int ok = w != NULL;
        fftw_forget_wisdom();
        ok = ok && fftw_import_wisdom_from_string(w) == 1;
        printf("%-34s %s\n", "wisdom export / import", ok ? "ok" : "FAILED");
        bad += !ok;
        free(w);
    }
    fftw_cleanup();
    return bad ? 1 : 0;
}
