/* fftw-wisdom for the B200 engine: plan the given transforms and write the accumulated
 * wisdom (measured kernel-variant choices per pass) to stdout or a file, so deployments
 * can pre-plan.  Same command line as the reference tool (tools/fftw-wisdom.c:73-131:
 * options; :122-131 size syntax; :133-144 canonical sizes), written from scratch against
 * the public fftw3.h API only.  Compiled twice: fftw-wisdom (double) and fftwf-wisdom
 * (-DB2_SINGLE).
 *
 *   size syntax   <type><inplace><direction><geometry>[v<howmany>]
 *                 type c|r|k, inplace i|o, direction f|b (not for k), geometry n1[xn2...],
 *                 for k each dimension is followed by f|b|h|e00|e01|e10|e11|o00|o01|o10|o11
 *                 (prefix letters may come in any order, as in the reference's bench parser)
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../include/fftw3.h"

#ifdef B2_SINGLE
#define X(name) fftwf_##name
typedef float R;
#define TOOL "fftwf-wisdom"
#else
#define X(name) fftw_##name
typedef double R;
#define TOOL "fftw-wisdom"
#endif

#define MAXRANK 8

typedef struct {
    char text[96];
    int type;                 /* 'c', 'r', 'k' */
    int inplace, backward;
    int rank, n[MAXRANK];
    X(r2r_kind) kind[MAXRANK];
    int howmany;
    double points;            /* for the size ordering */
} problem;

static int verbose = 0;

static int parse_kind(const char **s, X(r2r_kind) *k)
{
    const char *p = *s;
    if (*p == 'f') { *k = FFTW_R2HC; *s = p + 1; return 1; }
    if (*p == 'b') { *k = FFTW_HC2R; *s = p + 1; return 1; }
    if (*p == 'h') { *k = FFTW_DHT; *s = p + 1; return 1; }
    if ((*p == 'e' || *p == 'o') && (p[1] == '0' || p[1] == '1') && (p[2] == '0' || p[2] == '1')) {
        static const X(r2r_kind) e[4] = { FFTW_REDFT00, FFTW_REDFT01, FFTW_REDFT10, FFTW_REDFT11 };
        static const X(r2r_kind) o[4] = { FFTW_RODFT00, FFTW_RODFT01, FFTW_RODFT10, FFTW_RODFT11 };
        int idx = 2 * (p[1] - '0') + (p[2] - '0');
        *k = *p == 'e' ? e[idx] : o[idx];
        *s = p + 3;
        return 1;
    }
    return 0;
}

/* 0 on success */
static int parse_problem(const char *str, problem *q)
{
    const char *s = str;
    memset(q, 0, sizeof *q);
    snprintf(q->text, sizeof q->text, "%s", str);
    q->type = 'c'; q->inplace = 0; q->backward = 0; q->howmany = 1;
    while (*s && !isdigit((unsigned char)*s)) {
        switch (*s) {
        case 'c': case 'r': case 'k': q->type = *s; break;
        case 'i': q->inplace = 1; break;
        case 'o': q->inplace = 0; break;
        case 'f': q->backward = 0; break;
        case 'b': q->backward = 1; break;
        default: return -1;
        }
        ++s;
    }
    if (!*s) return -1;
    q->points = 1;
    for (;;) {
        long v = 0;
        if (!isdigit((unsigned char)*s) || q->rank >= MAXRANK) return -1;
        while (isdigit((unsigned char)*s)) { v = v * 10 + (*s - '0'); if (v > 2000000000L) return -1; ++s; }
        if (v <= 0) return -1;
        q->n[q->rank] = (int)v;
        q->kind[q->rank] = FFTW_R2HC;
        if (q->type == 'k' && !parse_kind(&s, &q->kind[q->rank])) return -1;
        q->points *= (double)v;
        q->rank++;
        if (*s == 'x') { ++s; continue; }
        break;
    }
    while (*s == 'v' || *s == '*') {
        long v = 0;
        ++s;
        if (!isdigit((unsigned char)*s)) return -1;
        while (isdigit((unsigned char)*s)) { v = v * 10 + (*s - '0'); if (v > 2000000000L) return -1; ++s; }
        if (v <= 0 || (double)q->howmany * (double)v > 2e9) return -1;
        q->howmany *= (int)v;
    }
    q->points *= q->howmany;
    return *s ? -1 : 0;
}

/* plan it (that is what deposits wisdom); 0 on success */
static int do_problem(const problem *q, unsigned flags)
{
    X(plan) p = NULL;
    size_t total = 1, ctotal = 1, alloc;
    int i;
    void *in, *out;
    for (i = 0; i < q->rank; ++i) {
        total *= (size_t)q->n[i];
        ctotal *= (size_t)(i == q->rank - 1 ? q->n[i] / 2 + 1 : q->n[i]);
    }
    if (verbose) printf("Planning transform: %s\n", q->text);
    /* bytes per transform: complex array, or the padded real array of an r2c/c2r pair */
    alloc = (q->type == 'c' ? 2 * total : (q->type == 'r' ? 2 * ctotal : total)) * sizeof(R) * (size_t)q->howmany;
    in = X(malloc)(alloc);
    out = q->inplace ? in : X(malloc)(alloc);
    if (!in || !out) { fprintf(stderr, TOOL ": out of memory for %s\n", q->text); return -1; }
    memset(in, 0, alloc);
    if (!q->inplace) memset(out, 0, alloc);
    if (q->type == 'c') {
        p = X(plan_many_dft)(q->rank, q->n, q->howmany, (X(complex) *)in, NULL, 1, (int)total,
                             (X(complex) *)out, NULL, 1, (int)total, q->backward ? FFTW_BACKWARD : FFTW_FORWARD, flags);
    } else if (q->type == 'r') {
        /* in place: rows padded to 2*(n/2+1) reals (api/rdft2-pad.c); out of place: dense */
        int rdist = q->inplace ? (int)(2 * ctotal) : (int)total;
        if (!q->backward)
            p = X(plan_many_dft_r2c)(q->rank, q->n, q->howmany, (R *)in, NULL, 1, rdist,
                                     (X(complex) *)out, NULL, 1, (int)ctotal, flags);
        else
            p = X(plan_many_dft_c2r)(q->rank, q->n, q->howmany, (X(complex) *)in, NULL, 1, (int)ctotal,
                                     (R *)out, NULL, 1, rdist, flags);
    } else {
        p = X(plan_many_r2r)(q->rank, q->n, q->howmany, (R *)in, NULL, 1, (int)total,
                             (R *)out, NULL, 1, (int)total, q->kind, flags);
    }
    if (p) X(destroy_plan)(p);
    else fprintf(stderr, TOOL ": could not plan %s (no CUDA device, or unsupported problem)\n", q->text);
    if (!q->inplace) X(free)(out);
    X(free)(in);
    return p ? 0 : -1;
}

static int by_size(const void *a, const void *b)
{
    double x = ((const problem *)a)->points, y = ((const problem *)b)->points;
    return x < y ? -1 : (x > y ? 1 : 0);
}

static void help(FILE *f)
{
    fprintf(f,
            "Usage: " TOOL " [options] [sizes]\n"
            "    Create wisdom (pre-planned/optimized transforms) for specified sizes,\n"
            "    writing wisdom to stdout (or to a file, using -o).\n"
            "\nOptions:\n"
            "                   -h, --help: print this help\n"
            "                -V, --version: print version info\n"
            "                -v, --verbose: verbose output\n"
            "              -c, --canonical: plan/optimize canonical set of sizes\n"
            "     -t <h>, --time-limit=<h>: time limit in hours (default: 0, no limit)\n"
            "  -o FILE, --output-file=FILE: output to FILE instead of stdout\n"
            "                -m, --measure: plan in MEASURE mode (PATIENT is default)\n"
            "               -e, --estimate: plan in ESTIMATE mode (not recommended)\n"
            "             -x, --exhaustive: plan in EXHAUSTIVE mode (may be slow)\n"
            "       -n, --no-system-wisdom: don't read /etc/fftw/ system wisdom file\n"
            "  -w FILE, --wisdom-file=FILE: read wisdom from FILE (stdin if -)\n"
            "            -T N, --threads=N: accepted for compatibility (the GPU engine has no CPU threads)\n"
            "\nSize syntax: <type><inplace><direction><geometry>\n"
            "      <type> = c/r/k for complex/real(r2c,c2r)/r2r\n"
            "   <inplace> = i/o for in/out-of place\n"
            " <direction> = f/b for forward/backward, omitted for k transforms\n"
            "  <geometry> = <n1>[x<n2>[x...]], e.g. 10x12x14\n"
            "               -- for k transforms, after each dimension is a <kind>:\n"
            "                     <kind> = f/b/h/e00/e01/e10/e11/o00/o01/o10/o11\n"
            "                              for R2HC/HC2R/DHT/REDFT00/.../RODFT11\n"
            "               -- an optional v<howmany> plans a batch of contiguous transforms\n");
}

static const char *canonical_sizes[] = {
    "1", "2", "4", "8", "16", "32", "64", "128", "256", "512", "1024", "2048", "4096", "8192", "16384", "32768",
    "65536", "131072", "262144", "524288", "1048576", "10", "100", "1000", "10000", "100000", "1000000",
    "2x2", "4x4", "8x8", "10x10", "16x16", "32x32", "64x64", "100x100", "128x128", "256x256", "512x512",
    "1000x1000", "1024x1024", "2x2x2", "4x4x4", "8x8x8", "10x10x10", "16x16x16", "32x32x32", "64x64x64",
    "100x100x100"
};

static problem *probs = NULL;
static int nprobs = 0, cap = 0;

static int add_problem(const char *s)
{
    if (nprobs == cap) {
        cap = cap ? 2 * cap : 64;
        probs = (problem *)realloc(probs, (size_t)cap * sizeof(problem));
        if (!probs) { fprintf(stderr, TOOL ": out of memory\n"); exit(EXIT_FAILURE); }
    }
    if (parse_problem(s, &probs[nprobs])) {
        fprintf(stderr, TOOL ": cannot parse size \"%s\"\n", s);
        return -1;
    }
    ++nprobs;
    return 0;
}

/* value of an option given as "-o FILE", "-oFILE" or "--long=FILE" */
static const char *optval(int argc, char **argv, int *i, const char *shortname, const char *longname)
{
    const char *a = argv[*i];
    size_t ll = strlen(longname);
    if (!strncmp(a, longname, ll) && a[ll] == '=') return a + ll + 1;
    if (!strcmp(a, longname) || !strcmp(a, shortname)) {
        if (*i + 1 >= argc) { fprintf(stderr, TOOL ": option %s needs an argument\n", a); exit(EXIT_FAILURE); }
        return argv[++*i];
    }
    if (!strncmp(a, shortname, 2) && a[2]) return a + 2;
    return NULL;
}

int main(int argc, char **argv)
{
    unsigned flags = 0;
    int impatient = 0, system_wisdom = 1, canonical = 0, i, failed = 0;
    double hours = 0;
    const char *outname = NULL, *v;
    FILE *out;
    time_t begin;

    for (i = 1; i < argc; ++i) {
        const char *a = argv[i];
        if (a[0] != '-' || !strcmp(a, "-")) break;
        if (!strcmp(a, "-h") || !strcmp(a, "--help")) { help(stdout); return EXIT_SUCCESS; }
        else if (!strcmp(a, "-V") || !strcmp(a, "--version")) {
            printf(TOOL " tool for %s\n", X(version));
            return EXIT_SUCCESS;
        }
        else if (!strcmp(a, "-v") || !strcmp(a, "--verbose")) verbose = 1;
        else if (!strcmp(a, "-c") || !strcmp(a, "--canonical")) canonical = 1;
        else if (!strcmp(a, "-m") || !strcmp(a, "--measure") || !strcmp(a, "-i") || !strcmp(a, "--impatient")) impatient = 1;
        else if (!strcmp(a, "-e") || !strcmp(a, "--estimate")) flags |= FFTW_ESTIMATE;
        else if (!strcmp(a, "-x") || !strcmp(a, "--exhaustive")) flags |= FFTW_EXHAUSTIVE;
        else if (!strcmp(a, "-n") || !strcmp(a, "--no-system-wisdom")) system_wisdom = 0;
        else if ((v = optval(argc, argv, &i, "-t", "--time-limit"))) hours = atof(v);
        else if ((v = optval(argc, argv, &i, "-o", "--output-file"))) outname = strcmp(v, "-") ? v : NULL;
        else if ((v = optval(argc, argv, &i, "-T", "--threads"))) { X(init_threads)(); X(plan_with_nthreads)(atoi(v) > 0 ? atoi(v) : 1); }
        else if ((v = optval(argc, argv, &i, "-w", "--wisdom-file"))) {
            FILE *w = stdin;
            if (strcmp(v, "-") && !(w = fopen(v, "r"))) {
                fprintf(stderr, TOOL ": error opening \"%s\": ", v);
                perror("");
                return EXIT_FAILURE;
            }
            if (!X(import_wisdom_from_file)(w)) {
                fprintf(stderr, TOOL ": error reading wisdom from \"%s\"\n", v);
                return EXIT_FAILURE;
            }
            if (w != stdin) fclose(w);
        }
        else { fprintf(stderr, TOOL ": unknown option %s\n", a); help(stderr); return EXIT_FAILURE; }
    }
    if (!impatient) flags |= FFTW_PATIENT;
    if (system_wisdom && !X(import_system_wisdom)() && verbose)
        fprintf(stderr, TOOL ": system-wisdom import failed\n");

    if (canonical) {
        static const char *types[] = { "cof", "cob", "cif", "cib", "rof", "rob", "rif", "rib" };
        unsigned s, t;
        for (s = 0; s < sizeof canonical_sizes / sizeof canonical_sizes[0]; ++s)
            for (t = 0; t < 8; ++t) {
                char ps[64];
                /* multi-dimensional sizes: in-place only, as the reference does */
                if (strchr(canonical_sizes[s], 'x') && strchr(types[t], 'o')) continue;
                snprintf(ps, sizeof ps, "%s%s", types[t], canonical_sizes[s]);
                add_problem(ps);
            }
    }
    for (; i < argc; ++i) {
        if (!strcmp(argv[i], "-")) {
            char s[1025];
            while (1 == fscanf(stdin, "%1024s", s)) failed |= add_problem(s) != 0;
        } else failed |= add_problem(argv[i]) != 0;
    }
    if (failed) return EXIT_FAILURE;
    qsort(probs, (size_t)nprobs, sizeof(problem), by_size);

    if (!outname) out = stdout;
    else if (!(out = fopen(outname, "w"))) {
        fprintf(stderr, TOOL ": error creating \"%s\"", outname);
        perror("");
        return EXIT_FAILURE;
    }
    begin = time(NULL);
    for (i = 0; i < nprobs; ++i)
        if (hours <= 0 || hours > (double)(time(NULL) - begin) / 3600.0)
            failed |= do_problem(&probs[i], flags) != 0;
    X(export_wisdom_to_file)(out);
    if (out != stdout) fclose(out);
    free(probs);
    X(cleanup)();
    return failed ? EXIT_FAILURE : EXIT_SUCCESS;
}
