/* wisdom.c -- accumulated planning results, exportable as text.
 *
 * Same role and text family as the reference's wisdom (kernel/planner.c:786-906):
 *   (fftw3_b200-<ver> fftw_wisdom #x<registry signature>
 *     (b200_fft_pass <variant> #x<patience> #x<prec> #x<sig hi> #x<sig lo>)
 *     ...
 *   )
 * An entry maps the signature of one device pass (size, strides, batch shape,
 * fused ops, in-placeness) to the kernel variant the measuring planner picked
 * (radix factorisation x CTA tile class).  A file written for a different
 * precision or a different kernel registry is rejected wholesale; malformed
 * input rolls back (kernel/planner.c:847-905).
 */
#include <stdlib.h>
#include <string.h>
#include "b2_internal.h"

#define B2_WISDOM_VERSION "fftw3_b200-1.0"
/* changes whenever the meaning of `variant` changes (kernel registry version) */
#define B2_REGISTRY_VERSION 0x0b2000020001ULL

typedef struct went {
    struct went *next;
    b2_sig sig;
    unsigned patience;
    int variant;
    int prec;
} went;

#define NBUCKET 1024
static went *g_tab[NBUCKET];

static uint64_t fnv(uint64_t h, const void *data, size_t n)
{
    const unsigned char *p = (const unsigned char *)data;
    size_t i;
    for (i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ULL; }
    return h;
}

/* Wisdom is only valid for the solver configuration that produced it (the reference refuses wisdom written
   by a different set of solvers, kernel/planner.c:847-852).  Here the "configuration" is the kernel registry
   AND the device the timings were taken on: registry version, device name and SM count are hashed into the
   signature in the header line; a file from another GPU model is rejected wholesale. */
static unsigned long long registry_sig(void)
{
    const char *name = b2d_device_name();
    int sms = b2d_sm_count();
    uint64_t h = fnv(0xcbf29ce484222325ULL ^ B2_REGISTRY_VERSION, name ? name : "", name ? strlen(name) : 0);
    h = fnv(h, &sms, sizeof sms);
    return (unsigned long long)h;
}

b2_sig b2_sig_of_pass(const b2d_fft_pass *p, int inplace)
{
    b2_sig s;
    int64_t v[24];
    int k = 0, i;
    v[k++] = p->prec; v[k++] = p->n; v[k++] = p->pre_op; v[k++] = p->post_op;
    v[k++] = p->bluestein; v[k++] = p->n_in; v[k++] = p->n_out;
    v[k++] = (p->pre_op & B2D_LOAD_R2R) ? p->r2r_kind + 16 * p->r2r_pair : -1;
    v[k++] = p->is; v[k++] = p->os; v[k++] = inplace;
    v[k++] = p->load_col; v[k++] = p->store_col; v[k++] = p->idx_mul;   /* tw4_off does not change the work */
    for (i = 0; i < B2D_MAX_BATCH_DIMS; ++i) { v[k++] = p->bn[i]; v[k++] = p->bis[i]; v[k++] = p->bos[i]; }
    s.h[0] = fnv(0xcbf29ce484222325ULL, v, (size_t)k * sizeof(int64_t));
    s.h[1] = fnv(0x84222325cbf29ce4ULL ^ s.h[0], v, (size_t)k * sizeof(int64_t));
    return s;
}

/* signature of a whole problem (sizes, strides, kind, in-placeness): keys the measured choice between
   whole-plan alternatives */
b2_sig b2_sig_of_problem(const b2_problem *q, int inplace)
{
    b2_sig s;
    int64_t v[8 + 6 * B2_MAXRANK];
    int k = 0, i;
    v[k++] = 0x706c616e; v[k++] = q->prec; v[k++] = q->kind; v[k++] = inplace;
    v[k++] = q->sz.rnk; v[k++] = q->vecsz.rnk;
    for (i = 0; i < q->sz.rnk; ++i) { v[k++] = q->sz.d[i].n; v[k++] = q->sz.d[i].is; v[k++] = q->sz.d[i].os; }
    for (i = 0; i < q->vecsz.rnk; ++i) { v[k++] = q->vecsz.d[i].n; v[k++] = q->vecsz.d[i].is; v[k++] = q->vecsz.d[i].os; }
    s.h[0] = fnv(0xcbf29ce484222325ULL, v, (size_t)k * sizeof(int64_t));
    s.h[1] = fnv(0x84222325cbf29ce4ULL ^ s.h[0], v, (size_t)k * sizeof(int64_t));
    return s;
}

static went *find(b2_sig s)
{
    went *e;
    for (e = g_tab[s.h[0] % NBUCKET]; e; e = e->next)
        if (e->sig.h[0] == s.h[0] && e->sig.h[1] == s.h[1]) return e;
    return NULL;
}

int b2_wisdom_lookup(b2_sig s, unsigned patience, int *variant)
{
    went *e = find(s);
    if (!e || e->patience < patience) return 0;
    *variant = e->variant;
    return 1;
}

static void store_prec(b2_sig s, unsigned patience, int variant, int prec)
{
    went *e = find(s);
    if (e) {
        if (patience >= e->patience) { e->patience = patience; e->variant = variant; }
        return;
    }
    e = (went *)malloc(sizeof *e);
    if (!e) return;
    e->sig = s; e->patience = patience; e->variant = variant; e->prec = prec;
    e->next = g_tab[s.h[0] % NBUCKET];
    g_tab[s.h[0] % NBUCKET] = e;
}

static int g_store_prec = 0;
void b2_wisdom_set_prec(int prec) { g_store_prec = prec; }

void b2_wisdom_store(b2_sig s, unsigned patience, int variant)
{
    store_prec(s, patience, variant, g_store_prec);
}

static void wisdom_forget_locked(void)
{
    int i;
    for (i = 0; i < NBUCKET; ++i) {
        went *e = g_tab[i];
        while (e) { went *n = e->next; free(e); e = n; }
        g_tab[i] = NULL;
    }
}

static void emit_str(void (*emit)(char, void *), void *d, const char *s)
{
    while (*s) emit(*s++, d);
}

static void wisdom_export_locked(void (*emit)(char c, void *), void *data, int prec)
{
    char buf[160];
    int i;
    snprintf(buf, sizeof buf, "(%s %s #x%llx\n", B2_WISDOM_VERSION,
             prec == B2D_F32 ? "fftwf_wisdom" : "fftw_wisdom", registry_sig());
    emit_str(emit, data, buf);
    for (i = 0; i < NBUCKET; ++i) {
        went *e;
        for (e = g_tab[i]; e; e = e->next) {
            if (e->prec != prec) continue;
            snprintf(buf, sizeof buf, "  (b200_fft_pass %d #x%x #x%x #x%llx #x%llx)\n", e->variant, e->patience,
                     (unsigned)e->prec, (unsigned long long)e->sig.h[0], (unsigned long long)e->sig.h[1]);
            emit_str(emit, data, buf);
        }
    }
    emit_str(emit, data, ")\n");
}

/* ---- import: tiny recursive-descent scanner over a char source ---- */
typedef struct { int (*next)(void *); void *data; int peeked, have; } src;

static int getc_(src *s) { if (s->have) { s->have = 0; return s->peeked; } return s->next(s->data); }
static void ungetc_(src *s, int c) { s->peeked = c; s->have = 1; }
static void skipws(src *s) { int c; while ((c = getc_(s)) == ' ' || c == '\n' || c == '\t' || c == '\r') {} ungetc_(s, c); }

static int token(src *s, char *buf, size_t cap)
{
    size_t n = 0;
    int c;
    skipws(s);
    while ((c = getc_(s)) != EOF && c != ' ' && c != '\n' && c != '\t' && c != '\r' && c != '(' && c != ')') {
        if (n + 1 < cap) buf[n++] = (char)c;
    }
    ungetc_(s, c);
    buf[n] = 0;
    return n > 0;
}

static int expect(src *s, int ch) { skipws(s); return getc_(s) == ch; }

static int hexval(const char *t, unsigned long long *v)
{
    char *end;
    if (t[0] != '#' || t[1] != 'x') return 0;
    *v = strtoull(t + 2, &end, 16);
    return *end == 0;
}

static int wisdom_import_locked(int (*next)(void *), void *data, int prec)
{
    src s;
    char tok[128];
    unsigned long long v;
    went *pending = NULL, *e;
    int ok = 0;
    s.next = next; s.data = data; s.have = 0; s.peeked = 0;
    if (!expect(&s, '(')) return 0;
    if (!token(&s, tok, sizeof tok) || strcmp(tok, B2_WISDOM_VERSION)) return 0;
    if (!token(&s, tok, sizeof tok) || strcmp(tok, prec == B2D_F32 ? "fftwf_wisdom" : "fftw_wisdom")) return 0;
    if (!token(&s, tok, sizeof tok) || !hexval(tok, &v) || v != registry_sig()) return 0;
    for (;;) {
        int c;
        unsigned long long pat, pr, h0, h1;
        long variant;
        char *end;
        skipws(&s);
        c = getc_(&s);
        if (c == ')') { ok = 1; break; }
        if (c != '(') break;
        if (!token(&s, tok, sizeof tok) || strcmp(tok, "b200_fft_pass")) break;
        if (!token(&s, tok, sizeof tok)) break;
        variant = strtol(tok, &end, 10);
        if (*end) break;
        if (!token(&s, tok, sizeof tok) || !hexval(tok, &pat)) break;
        if (!token(&s, tok, sizeof tok) || !hexval(tok, &pr)) break;
        if (!token(&s, tok, sizeof tok) || !hexval(tok, &h0)) break;
        if (!token(&s, tok, sizeof tok) || !hexval(tok, &h1)) break;
        if (!expect(&s, ')')) break;
        if ((int)pr != prec) break;
        e = (went *)malloc(sizeof *e);
        if (!e) break;
        e->sig.h[0] = h0; e->sig.h[1] = h1; e->patience = (unsigned)pat; e->variant = (int)variant; e->prec = prec;
        e->next = pending; pending = e;
    }
    /* commit or roll back */
    while (pending) {
        e = pending; pending = e->next;
        if (ok) store_prec(e->sig, e->patience, e->variant, e->prec);
        free(e);
    }
    return ok;
}

/* public entry points: planner state is shared, see b2_planner_lock() */
void b2_wisdom_forget(void)
{
    b2_planner_lock();
    wisdom_forget_locked();
    b2_planner_unlock();
}

void b2_wisdom_export(void (*emit)(char c, void *), void *data, int prec)
{
    b2_planner_lock();
    wisdom_export_locked(emit, data, prec);
    b2_planner_unlock();
}

int b2_wisdom_import(int (*next)(void *), void *data, int prec)
{
    int ok;
    b2_planner_lock();
    ok = wisdom_import_locked(next, data, prec);
    b2_planner_unlock();
    return ok;
}
