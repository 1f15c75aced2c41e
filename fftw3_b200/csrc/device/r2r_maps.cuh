// r2r_maps.cuh -- the real-to-real transforms as maps around ONE complex FFT of length M,
// in the form the FFT passes can fuse: PRE as a gather (work element i from the real input
// line), POST as a scatter (FFT output q to the real output line).  With these in the load
// and the store of a pass, every r2r dimension costs one read and one write of the array.
//
// Identities restated (per kind) from the reference: reodft/reodft010e-r2hc.c:84-290 (types
// 2/3 through a same-size transform with an even/odd permutation and a quarter-wave
// twiddle), reodft/redft00e-r2hc-pad.c and rodft00e-r2hc-pad.c (types 1 through the padded
// symmetric extension), reodft/reodft11e-radix2.c (types 4 through a double-length
// transform), rdft/rdft-dht.c (DHT from R2HC) and the halfcomplex layout of
// rdft/rdft2-rdft.c:42-74; definitions doc/reference.texi:2060-2353.
// tests/proto_algorithms.py holds the numpy prototype of each map; real_ops.cuh holds the
// same maps as stand-alone passes for lengths that do not fit one CTA.
#pragma once
#include <stdint.h>

namespace b2 {

enum { K_R2HC = 0, K_HC2R, K_DHT, K_REDFT00, K_REDFT01, K_REDFT10, K_REDFT11,
       K_RODFT00, K_RODFT01, K_RODFT10, K_RODFT11 };

template <typename T>
struct RealLineIn {          // element j of a strided real line
    const T *p; int64_t s;
    B2_HD T operator()(int j) const { return p[(int64_t)j * s]; }
};
template <typename T>
struct RealLineOut {
    T *p; int64_t s;
    B2_HD void operator()(int k, T v) const { p[(int64_t)k * s] = v; }
};

// collects what a POST map writes for one spectrum element (at most two (index, value) pairs), so that the
// values of neighbouring lines can leave in one vector store
template <typename T>
struct StashOut {
    mutable int idx[2]; mutable T val[2]; mutable int cnt;
    B2_HD void operator()(int k, T v) const { idx[cnt] = k; val[cnt] = v; ++cnt; }
};

// complex work length M for a kind of physical size n
B2_HD int r2r_work_len(int kind, int n)
{
    switch (kind) {
    case K_REDFT00: return 2 * (n - 1);
    case K_RODFT00: return 2 * (n + 1);
    case K_REDFT11: case K_RODFT11: return 2 * n;
    default: return n;
    }
}

// kinds whose PRE sequence is purely real: two lines can share one complex transform
B2_HD bool r2r_pairable(int kind)
{
    return kind == K_R2HC || kind == K_DHT || kind == K_REDFT00 || kind == K_RODFT00 || kind == K_REDFT10 ||
           kind == K_RODFT10;
}

// spectra of the two real sequences packed as u + i v, from Z_k and Z_{M-k}
template <typename T>
B2_HD void r2r_unpack_pair(cplx<T> zk, cplx<T> zm, cplx<T> &u, cplx<T> &v)
{
    u.x = T(0.5) * (zk.x + zm.x); u.y = T(0.5) * (zk.y - zm.y);      // (Z_k + conj Z_{M-k}) / 2
    v.x = T(0.5) * (zk.y + zm.y); v.y = T(0.5) * (zm.x - zk.x);      // (Z_k - conj Z_{M-k}) / (2 i)
}

// PRE: work element i (0 <= i < M).  tw = quarter-wave table (TAB_QUARTER of n), used by types 3 and 4.
template <typename T, typename In>
B2_HD cplx<T> r2r_pre_value(int kind, int n, int i, const cplx<T> *tw, const In &x)
{
    cplx<T> z; z.x = T(0); z.y = T(0);
    switch (kind) {
    case K_R2HC: case K_DHT:
        z.x = x(i);
        break;
    case K_HC2R:
        if (i == 0) z.x = x(0);
        else if (2 * i < n) { z.x = x(i); z.y = -x(n - i); }
        else if (2 * i == n) z.x = x(i);
        else { z.x = x(n - i); z.y = x(i); }
        break;
    case K_REDFT00:
        z.x = (i < n) ? x(i) : x(2 * (n - 1) - i);
        break;
    case K_RODFT00:
        if (i >= 1 && i <= n) z.x = x(i - 1);
        else if (i > n + 1) z.x = -x(2 * (n + 1) - i - 1);
        break;
    case K_REDFT10: case K_RODFT10: {
        const int h = (n + 1) / 2;
        const int j = (i < h) ? 2 * i : 2 * (n - 1 - i) + 1;
        T v = x(j);
        if (kind == K_RODFT10 && (j & 1)) v = -v;
        z.x = v;
        break;
    }
    case K_REDFT01: case K_RODFT01: {
        T a, c;   // a = X_i, c = X_{n-i} (X_n = 0)
        if (kind == K_REDFT01) { a = x(i); c = (i == 0) ? T(0) : x(n - i); }
        else { a = x(n - 1 - i); c = (i == 0) ? T(0) : x(i - 1); }
        cplx<T> v; v.x = a; v.y = c;
        z = cmul(tw[i], v);
        break;
    }
    case K_REDFT11: case K_RODFT11:
        if (i < n) { T v = x(i); z.x = tw[i].x * v; z.y = tw[i].y * v; }
        break;
    }
    return z;
}

// POST: FFT output q (0 <= q < M) with value z goes to at most two elements of the output line
template <typename T, typename Out>
B2_HD void r2r_post_scatter(int kind, int n, int q, cplx<T> z, const cplx<T> *tw, const Out &y)
{
    switch (kind) {
    default:
    case K_R2HC:            // r0 r1 ... r(n/2) i((n+1)/2-1) ... i1
        if (2 * q <= n) y(q, z.x);
        if (q >= 1 && 2 * q < n) y(n - q, z.y);
        break;
    case K_HC2R: case K_REDFT00:
        if (q < n) y(q, z.x);
        break;
    case K_DHT:
        y(q, z.x - z.y);
        break;
    case K_RODFT00:
        if (q >= 1 && q <= n) y(q - 1, -z.y);
        break;
    case K_REDFT10: { cplx<T> v = cmul(tw[q], z); y(q, T(2) * v.x); break; }
    case K_RODFT10: { cplx<T> v = cmul(tw[q], z); y(n - 1 - q, T(2) * v.x); break; }
    case K_REDFT01: case K_RODFT01: {
        const bool even = q < (n + 1) / 2;
        const int k = even ? 2 * q : 2 * (n - 1 - q) + 1;
        y(k, (kind == K_RODFT01 && !even) ? -z.x : z.x);
        break;
    }
    case K_REDFT11: if (q < n) { cplx<T> v = cmul(tw[n + q], z); y(q, T(2) * v.x); } break;
    case K_RODFT11: if (q < n) { cplx<T> v = cmul(tw[n + q], z); y(q, T(-2) * v.y); } break;
    }
}

}  // namespace b2
