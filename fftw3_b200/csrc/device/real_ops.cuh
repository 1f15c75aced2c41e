// real_ops.cuh -- element maps around the complex FFT passes for real-data
// transforms, and the strided N-d copy ("rank-0 transform").
//
// r2r kinds are computed as  user line --PRE--> complex work sequence of length
// M --FFT(M)--> --POST--> user line.  The maps restate, per kind, the identities
// the reference uses in reodft/reodft010e-r2hc.c:84-290 (types 2/3 through one
// same-size transform with an even/odd permutation and a quarter-wave twiddle),
// reodft/redft00e-r2hc-pad.c and rodft00e-r2hc-pad.c (types 1 through a padded
// transform of the symmetric extension), rdft/rdft-dht.c (DHT from R2HC) and
// the halfcomplex layout of rdft/rdft2-rdft.c:42-74.  tests/proto_algorithms.py
// holds the numpy prototype of each map.
//
// Copy: kernel/cpy2d.c:36-204 / rdft/rank0.c:117-381 analogue.
#pragma once
#include "fft_generic.cuh"

namespace b2 {

B2_HD int64_t realop_user_offset(const b2d_realop &r, int64_t b)
{
    int64_t b0 = b % r.bn[0];
    int64_t rest = b / r.bn[0];
    int64_t b1 = rest % r.bn[1];
    int64_t b2 = rest / r.bn[1];
    return b0 * r.bxs[0] + b1 * r.bxs[1] + b2 * r.bxs[2];
}

// PRE: one work element i of batch line b
template <typename T>
B2_HD void r2r_pre_elem(const b2d_realop &r, int kind, int64_t b, int i)
{
    const T *x = (const T *)r.x_re + realop_user_offset(r, b);
    const int64_t s = r.xs;
    const int n = r.n;
    const cplx<T> *tw = (const cplx<T> *)r.tw;
    cplx<T> z; z.x = T(0); z.y = T(0);
    switch (kind) {
    case K_R2HC: case K_DHT:
        z.x = x[i * s];
        break;
    case K_HC2R:
        if (i == 0) z.x = x[0];
        else if (2 * i < n) { z.x = x[i * s]; z.y = -x[(int64_t)(n - i) * s]; }
        else if (2 * i == n) z.x = x[i * s];
        else { z.x = x[(int64_t)(n - i) * s]; z.y = x[i * s]; }
        break;
    case K_REDFT00:
        z.x = (i < n) ? x[i * s] : x[(int64_t)(2 * (n - 1) - i) * s];
        break;
    case K_RODFT00:
        if (i >= 1 && i <= n) z.x = x[(int64_t)(i - 1) * s];
        else if (i > n + 1) z.x = -x[(int64_t)(2 * (n + 1) - i - 1) * s];
        break;
    case K_REDFT10: case K_RODFT10: {
        int h = (n + 1) / 2;
        int j = (i < h) ? 2 * i : 2 * (n - 1 - i) + 1;
        T v = x[(int64_t)j * s];
        if (kind == K_RODFT10 && (j & 1)) v = -v;
        z.x = v;
        break;
    }
    case K_REDFT01: case K_RODFT01: {
        T a, c;   // a = X_i, c = X_{n-i} (X_n = 0)
        if (kind == K_REDFT01) {
            a = x[(int64_t)i * s];
            c = (i == 0) ? T(0) : x[(int64_t)(n - i) * s];
        } else {
            a = x[(int64_t)(n - 1 - i) * s];
            c = (i == 0) ? T(0) : x[(int64_t)(i - 1) * s];
        }
        cplx<T> v; v.x = a; v.y = c;
        z = cmul(tw[i], v);
        break;
    }
    case K_REDFT11: case K_RODFT11:
        if (i < n) { T v = x[(int64_t)i * s]; z.x = tw[i].x * v; z.y = tw[i].y * v; }
        break;
    }
    ((cplx<T> *)r.work)[b * r.wdist + i] = z;
}

// POST: one output element k of batch line b
template <typename T>
B2_HD void r2r_post_elem(const b2d_realop &r, int kind, int64_t b, int k)
{
    T *y = (T *)r.y_re + realop_user_offset(r, b);
    const int64_t s = r.xs;
    const int n = r.n;
    const cplx<T> *Z = (const cplx<T> *)r.work + b * r.wdist;
    const cplx<T> *tw = (const cplx<T> *)r.tw;
    T out;
    switch (kind) {
    default:
    case K_R2HC: out = (2 * k <= n) ? Z[k].x : Z[n - k].y; break;
    case K_HC2R: case K_REDFT00: out = Z[k].x; break;
    case K_DHT: out = Z[k].x - Z[k].y; break;
    case K_RODFT00: out = -Z[k + 1].y; break;
    case K_REDFT10: { cplx<T> v = cmul(tw[k], Z[k]); out = T(2) * v.x; break; }
    case K_RODFT10: { int q = n - 1 - k; cplx<T> v = cmul(tw[q], Z[q]); out = T(2) * v.x; break; }
    case K_REDFT01: case K_RODFT01: {
        int q = (k & 1) ? (n - 1 - (k - 1) / 2) : (k / 2);
        out = Z[q].x;
        if (kind == K_RODFT01 && (k & 1)) out = -out;
        break;
    }
    case K_REDFT11: { cplx<T> v = cmul(tw[n + k], Z[k]); out = T(2) * v.x; break; }
    case K_RODFT11: { cplx<T> v = cmul(tw[n + k], Z[k]); out = T(-2) * v.y; break; }
    }
    y[(int64_t)k * s] = out;
}

// r2c of even n: work holds Z = FFT_{n/2}(x_even + i x_odd); pair index q in
// [0, m/2] produces X_q and X_{m-q}  (q = 0 also produces X_m).
// X_k = 1/2 [(Z_k + conj Z_{m-k}) - i w^k (Z_k - conj Z_{m-k})],  w = exp(-2 pi i / n)
template <typename T>
B2_HD void r2c_post_pair(const b2d_realop &r, int64_t b, int q)
{
    const int m = r.m;
    const cplx<T> *Z = (const cplx<T> *)r.work + b * r.wdist;
    const cplx<T> *tw = (const cplx<T> *)r.tw;       // tw[k] = exp(-2 pi i k / n), k <= m/2
    int64_t off = realop_user_offset(r, b);
    T *yr = (T *)r.y_re + off, *yi = (T *)r.y_im + off;
    const int64_t s = r.xs;
    if (q == 0) {
        cplx<T> z0 = Z[0];
        yr[0] = z0.x + z0.y; yi[0] = T(0);
        yr[(int64_t)m * s] = z0.x - z0.y; yi[(int64_t)m * s] = T(0);
        return;
    }
    const int p = m - q;
    cplx<T> a = Z[q], c = Z[p];
    // sum = a + conj(c), dif = a - conj(c)
    T sr = a.x + c.x, si = a.y - c.y, dr = a.x - c.x, di = a.y + c.y;
    cplx<T> w = tw[q];
    // t = -i * w * dif
    T tr = w.x * di + w.y * dr;      // Re(-i w d) = Im(w d) = w.x di + w.y dr
    T ti = -(w.x * dr - w.y * di);   // Im(-i w d) = -Re(w d)
    yr[(int64_t)q * s] = T(0.5) * (sr + tr);
    yi[(int64_t)q * s] = T(0.5) * (si + ti);
    if (p != q) {
        // X_{m-q} = conj( 1/2 [ sum + i w dif ] ) evaluated through the mirror identity
        yr[(int64_t)p * s] = T(0.5) * (sr - tr);
        yi[(int64_t)p * s] = T(0.5) * (-si + ti);
    }
}

// c2r of even n: user X[0..m] -> work = swap(Z) so that a FORWARD pass on the
// work line followed by reading (im, re) is the backward transform.
// Z_k = (X_k + conj X_{m-k}) + i conj(w)^k (X_k - conj X_{m-k})
template <typename T>
B2_HD void c2r_pre_elem(const b2d_realop &r, int64_t b, int k)
{
    const int m = r.m;
    int64_t off = realop_user_offset(r, b);
    const T *xr = (const T *)r.x_re + off, *xi = (const T *)r.x_im + off;
    const int64_t s = r.xs;
    const cplx<T> *tw = (const cplx<T> *)r.tw;   // tw[k] = exp(-2 pi i k / n), k < m
    cplx<T> a, c;
    a.x = xr[(int64_t)k * s]; a.y = xi[(int64_t)k * s];
    c.x = xr[(int64_t)(m - k) * s]; c.y = xi[(int64_t)(m - k) * s];
    if (k == 0) { a.y = T(0); c.y = T(0); }
    T sr = a.x + c.x, si = a.y - c.y, dr = a.x - c.x, di = a.y + c.y;
    cplx<T> w = tw[k]; w.y = -w.y;               // conj(w^k) = exp(+2 pi i k / n)
    // u = i * w * dif
    T ur = -(w.x * di + w.y * dr);
    T ui = w.x * dr - w.y * di;
    cplx<T> z; z.x = sr + ur; z.y = si + ui;
    cplx<T> o; o.x = z.y; o.y = z.x;             // swapped
    ((cplx<T> *)r.work)[b * r.wdist + k] = o;
}

// ---- Bluestein element maps for sizes too large for the one-CTA kernel ----
template <typename T>
B2_HD void blue_elem(const b2d_realop &r, int64_t b, int i)
{
    cplx<T> *W = (cplx<T> *)r.work + b * r.wdist;
    const cplx<T> *chirp = (const cplx<T> *)r.tw;
    const int64_t s = r.xs;
    if (r.op == B2D_ROP_BLUE_PRE) {
        cplx<T> z; z.x = T(0); z.y = T(0);
        if (i < r.n_lim) {
            int64_t off = realop_user_offset(r, b);
            const T *xr = (const T *)r.x_re + off, *xi = (const T *)r.x_im + off;
            const int n = r.n;
            if (r.flags & B2D_LOAD_REAL) { z.x = xr[(int64_t)i * s]; }
            else if (r.flags & B2D_LOAD_HERMCONJ) {
                if (2 * i <= n) { z.x = xr[(int64_t)i * s]; z.y = (i == 0 || 2 * i == n) ? T(0) : -xi[(int64_t)i * s]; }
                else { z.x = xr[(int64_t)(n - i) * s]; z.y = xi[(int64_t)(n - i) * s]; }
            } else { z.x = xr[(int64_t)i * s]; z.y = xi[(int64_t)i * s]; }
            z = cmul(z, chirp[i]);
        }
        W[i] = z;
    } else if (r.op == B2D_ROP_BLUE_MID) {
        cplx<T> v = cmul(W[i], ((const cplx<T> *)r.aux)[i]);
        v.y = -v.y;
        W[i] = v;
    } else {
        if (i >= r.n_lim) return;
        int64_t off = realop_user_offset(r, b);
        T *yr = (T *)r.y_re + off, *yi = (T *)r.y_im + off;
        cplx<T> v = W[i];
        v.y = -v.y;
        v = cmul(v, chirp[i]);
        yr[(int64_t)i * s] = v.x * (T)r.scale;
        if (!(r.flags & B2D_STORE_REALPART)) yi[(int64_t)i * s] = v.y * (T)r.scale;
    }
}

// strided N-d copy, one element (1 or 2 reals) per index
template <typename T>
B2_HD void copy_elem(const b2d_copy &c, int64_t idx)
{
    int64_t i0 = idx % c.n[0]; idx /= c.n[0];
    int64_t i1 = idx % c.n[1]; idx /= c.n[1];
    int64_t i2 = idx % c.n[2]; idx /= c.n[2];
    int64_t i3 = idx;
    int64_t io = i0 * c.is[0] + i1 * c.is[1] + i2 * c.is[2] + i3 * c.is[3];
    int64_t oo = i0 * c.os[0] + i1 * c.os[1] + i2 * c.os[2] + i3 * c.os[3];
    const T *in = (const T *)c.in;
    T *out = (T *)c.out;
    if (c.npeer) { in = (const T *)c.peer_in[i3]; io -= i3 * c.is[3]; }
    if (c.elem_reals == 2 && (((uintptr_t)(in + io) | (uintptr_t)(out + oo)) % (2 * sizeof(T))) == 0) {
        *reinterpret_cast<cplx<T> *>(out + oo) = *reinterpret_cast<const cplx<T> *>(in + io);
        return;
    }
    out[oo] = in[io];
    if (c.elem_reals == 2) out[oo + 1] = in[io + 1];
}

#ifdef __CUDACC__
// Tiled transposing copy (kernel/transpose.c:24-190, kernel/tile2d.c, rdft/vrank3-transpose.c
// analogue): when the dimension that is contiguous on the input side (da) is not the one that is
// contiguous on the output side (db), move 32x32 tiles through shared memory so that both the
// loads and the stores of a warp are contiguous.  V = one element (real, or a whole complex).
template <typename V, int TS>
__global__ void __launch_bounds__(256) transpose_kernel(const __grid_constant__ b2d_copy c, int da, int db, int dc, int dd,
                                                        int64_t tiles_a, int64_t tiles_b)
{
    // TS x TS tile, 256 threads: TS*TS/256 independent loads in flight per thread (16 for the 64 x 64 tile of
    // 4- and 8-byte elements) -- the 32 x 32 tile with 4 loads per thread ran at 2.5 TB/s (profiles/r02_c5b_pieces.txt)
    __shared__ V tile[TS][TS + 1];
    constexpr int ROWS = 256 / TS;                      // tile rows covered by one sweep of the CTA
    constexpr int SWEEPS = TS / ROWS;
    const int es = c.elem_reals;                        // strides are in reals; V spans es reals
    int64_t blk = blockIdx.x;
    const int64_t ta = blk % tiles_a; blk /= tiles_a;
    const int64_t tb = blk % tiles_b; blk /= tiles_b;
    const int64_t ic = blk % c.n[dc], id = blk / c.n[dc];
    const int64_t rest_in = ic * c.is[dc] + id * c.is[dd], rest_out = ic * c.os[dc] + id * c.os[dd];
    const int tx = threadIdx.x % TS, ty = threadIdx.x / TS;
    const V *in = reinterpret_cast<const V *>(c.in);
    V *out = reinterpret_cast<V *>(c.out);
    V v[SWEEPS];
#pragma unroll
    for (int r = 0; r < SWEEPS; ++r) {
        const int64_t a = ta * TS + tx, b = tb * TS + ty + ROWS * r;
        if (a < c.n[da] && b < c.n[db]) v[r] = in[(a * c.is[da] + b * c.is[db] + rest_in) / es];
    }
#pragma unroll
    for (int r = 0; r < SWEEPS; ++r) tile[ty + ROWS * r][tx] = v[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SWEEPS; ++r) {
        const int64_t b = tb * TS + tx, a = ta * TS + ty + ROWS * r;
        if (a < c.n[da] && b < c.n[db]) out[(a * c.os[da] + b * c.os[db] + rest_out) / es] = tile[tx][ty + ROWS * r];
    }
}

template <typename T>
__global__ void realop_kernel(const __grid_constant__ b2d_realop r, int len, int chunks)
{
    int64_t b = blockIdx.x / chunks;
    int i = (int)(blockIdx.x - b * chunks) * blockDim.x + threadIdx.x;
    if (i >= len) return;
    if (r.op == B2D_ROP_R2C_POST) r2c_post_pair<T>(r, b, i);
    else if (r.op == B2D_ROP_C2R_PRE) c2r_pre_elem<T>(r, b, i);
    else if (r.op >= B2D_ROP_BLUE_PRE && r.op <= B2D_ROP_BLUE_POST) blue_elem<T>(r, b, i);
    else if (r.op & B2D_ROP_R2R_POST) r2r_post_elem<T>(r, r.op & 15, b, i);
    else r2r_pre_elem<T>(r, r.op & 15, b, i);
}

template <typename T>
__global__ void copy_kernel(const __grid_constant__ b2d_copy c, int64_t total)
{
    // grid-stride, 4 independent elements in flight per thread (peer loads over NVLink
    // need the memory-level parallelism)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; idx + 3 * stride < total; idx += 4 * stride) {
        copy_elem<T>(c, idx);
        copy_elem<T>(c, idx + stride);
        copy_elem<T>(c, idx + 2 * stride);
        copy_elem<T>(c, idx + 3 * stride);
    }
    for (; idx < total; idx += stride) copy_elem<T>(c, idx);
}
#endif

}  // namespace b2
