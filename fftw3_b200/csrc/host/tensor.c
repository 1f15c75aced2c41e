/* tensor.c -- stride tensors of the host plan layer.
 * Same canonicalisations as the reference's kernel/tensor7.c (sort by stride
 * :32-59, drop unit dims :99, merge contiguous dims :116-168), re-implemented
 * for the fixed-capacity b2_tensor. */
#include <stdlib.h>
#include "b2_internal.h"

void b2_tensor_init(b2_tensor *t, int rnk)
{
    int i;
    t->rnk = rnk;
    for (i = 0; i < B2_MAXRANK; ++i) { t->d[i].n = 1; t->d[i].is = 0; t->d[i].os = 0; }
}

int64_t b2_tensor_count(const b2_tensor *t)
{
    int i;
    int64_t c = 1;
    if (t->rnk == B2_RNK_MINFTY) return 0;
    for (i = 0; i < t->rnk; ++i) c *= t->d[i].n;
    return c;
}

void b2_tensor_drop_unit(b2_tensor *t)
{
    int i, k = 0;
    if (t->rnk == B2_RNK_MINFTY) return;
    for (i = 0; i < t->rnk; ++i)
        if (t->d[i].n != 1) t->d[k++] = t->d[i];
    t->rnk = k;
}

void b2_tensor_append(b2_tensor *t, const b2_tensor *a)
{
    int i;
    if (a->rnk == B2_RNK_MINFTY || t->rnk == B2_RNK_MINFTY) { t->rnk = B2_RNK_MINFTY; return; }
    for (i = 0; i < a->rnk && t->rnk < B2_MAXRANK; ++i) t->d[t->rnk++] = a->d[i];
}

static int64_t iabs64(int64_t x) { return x < 0 ? -x : x; }

static int cmp_dim(const void *pa, const void *pb)
{
    const b2_dim *a = (const b2_dim *)pa, *b = (const b2_dim *)pb;
    int64_t ka = iabs64(a->os), kb = iabs64(b->os);
    if (ka != kb) return ka < kb ? -1 : 1;
    ka = iabs64(a->is); kb = iabs64(b->is);
    if (ka != kb) return ka < kb ? -1 : 1;
    if (a->n != b->n) return a->n < b->n ? -1 : 1;
    return 0;
}

void b2_tensor_sort_merge(b2_tensor *t)
{
    int i, k;
    if (t->rnk <= 1) return;
    qsort(t->d, (size_t)t->rnk, sizeof(b2_dim), cmp_dim);
    /* dims are ascending in |os|; d[k] (inner) and d[i] (outer) merge when the
       outer stride equals inner.n * inner stride on both sides */
    k = 0;
    for (i = 1; i < t->rnk; ++i) {
        b2_dim *in = &t->d[k], *out = &t->d[i];
        if (out->is == in->n * in->is && out->os == in->n * in->os) {
            in->n *= out->n;
        } else {
            t->d[++k] = *out;
        }
    }
    t->rnk = k + 1;
}

void b2_tensor_span(const b2_tensor *t, int use_os, int64_t *lo, int64_t *hi)
{
    int i;
    int64_t mn = 0, mx = 0;
    for (i = 0; i < t->rnk; ++i) {
        int64_t s = use_os ? t->d[i].os : t->d[i].is;
        int64_t e = (t->d[i].n - 1) * s;
        if (e < 0) mn += e; else mx += e;
    }
    *lo = mn; *hi = mx;
}

int b2_tensor_inplace_ok(const b2_tensor *t)
{
    int i;
    for (i = 0; i < t->rnk; ++i)
        if (t->d[i].is != t->d[i].os) return 0;
    return 1;
}
