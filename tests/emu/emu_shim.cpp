// tests/emu/emu_shim.cpp -- UNIT-TEST DOUBLE for include/b200fft_device.h.
//
// NOT part of the product.  It exists so that the C host layer (argument
// checking, tensor canonicalisation, plan building, pass descriptors, wisdom)
// and the kernels' index algebra can be unit-tested in `-m "not gpu"` runs on a
// box without a GPU: "device memory" is malloc, and a kernel launch runs the
// very same __host__ __device__ phase functions of fft_generic.cuh /
// real_ops.cuh with CTAs and threads as plain loops.  It is built into
// tests/_emu/ by tests/conftest.py and only ever loaded by tests.  The product
// library (fftw3_b200/lib/libfftw3_b200.so) links shim.cu instead and has no
// CPU path at all.
#include <stdio.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include "../../fftw3_b200/csrc/device/fft_generic.cuh"
#include "../../fftw3_b200/csrc/device/real_ops.cuh"
#include "../../fftw3_b200/csrc/device/fft_split.cuh"

static char g_err[256] = "";
static std::atomic<uint64_t> g_launches{0};
static std::map<char *, size_t> g_dev;          // "device" allocations: base -> bytes
static std::mutex g_dev_mu;                     // execute is re-entrant: the registry must be too
static std::chrono::steady_clock::time_point g_t0;

template <typename T>
static void run_fft_pass(const b2d_fft_pass &p)
{
    using namespace b2;
    int64_t blocks = grid_blocks(p);
    int nthreads = p.tpb * p.tpx;
    if (nthreads < 32) nthreads = 32;
    if (nthreads > 1024) nthreads = 1024;
    std::vector<unsigned char> raw(smem_bytes<T>(p) + 64);
    unsigned char *base = raw.data();
    base += (16 - ((uintptr_t)base & 15)) & 15;
    for (int64_t blk = 0; blk < blocks; ++blk) {
        Smem<T> s = carve<T>(p, base);
        TileCtx c = decode_block(p, blk);
        for (int t = 0; t < nthreads; ++t) phase_offsets<T>(p, s, c, t);
        for (int t = 0; t < nthreads; ++t) phase_twiddles<T>(p, s, t, nthreads);
        const cplx<T> *twp = s.tw ? s.tw : (const cplx<T> *)p.tw;
        int swi = 0, swo = 0;
        const bool plain = plain_ok(p, &swi, &swo);     // same per-launch decision as shim.cu
        for (int t = 0; t < nthreads; ++t) {
            if (plain) phase_load_plain<T>(p, s, t, nthreads, swi); else phase_load<T>(p, s, t, nthreads);
        }
        cplx<T> *src = s.a, *dst = s.b;
        int reps = p.bluestein ? 2 : 1;
        for (int rep = 0; rep < reps; ++rep) {
            int ns = 1;
            for (int st = 0; st < p.nstages; ++st) {
                for (int t = 0; t < nthreads; ++t) phase_stage<T>(p, st, ns, src, dst, s.pitch, t, nthreads, twp);
                ns *= p.radix[st];
                cplx<T> *tmp = src; src = dst; dst = tmp;
            }
            if (p.bluestein && rep == 0)
                for (int t = 0; t < nthreads; ++t) phase_pointwise<T>(p, s, src, s.pitch, t, nthreads);
        }
        for (int t = 0; t < nthreads; ++t) {
            if (plain) phase_store_plain<T>(p, s, src, t, nthreads, swo); else phase_store<T>(p, s, c, src, t, nthreads);
        }
    }
}

template <typename T>
static void run_realop(const b2d_realop &r)
{
    using namespace b2;
    int kind = r.op & 15, len;
    if (r.op == B2D_ROP_R2C_POST) len = r.m / 2 + 1;
    else if (r.op == B2D_ROP_C2R_PRE) len = r.m;
    else if (r.op >= B2D_ROP_BLUE_PRE && r.op <= B2D_ROP_BLUE_POST) len = (r.op == B2D_ROP_BLUE_POST) ? r.n_lim : r.m;
    else if (r.op & B2D_ROP_R2R_POST) len = r.n;
    else len = r2r_work_len(kind, r.n);
    int64_t nb = r.bn[0] * r.bn[1] * r.bn[2];
    for (int64_t b = 0; b < nb; ++b)
        for (int i = 0; i < len; ++i) {
            if (r.op == B2D_ROP_R2C_POST) r2c_post_pair<T>(r, b, i);
            else if (r.op == B2D_ROP_C2R_PRE) c2r_pre_elem<T>(r, b, i);
            else if (r.op >= B2D_ROP_BLUE_PRE && r.op <= B2D_ROP_BLUE_POST) blue_elem<T>(r, b, i);
            else if (r.op & B2D_ROP_R2R_POST) r2r_post_elem<T>(r, kind, b, i);
            else r2r_pre_elem<T>(r, kind, b, i);
        }
}

/* the register-only sub-passes: same per-thread body as the CUDA kernel, CTAs and threads as loops */
template <typename T, int R, int PHASE>
static void run_split(const b2d_split_pass &p, int swap)
{
    using namespace b2split;
    const int64_t blocks = split_blocks(p);
    const int64_t tiles_c = (p.nc + B2_SPLIT_THREADS - 1) / B2_SPLIT_THREADS;
    for (int64_t blk = 0; blk < blocks; ++blk) {
        b2::cplx<T> tws[R];
        const int o = (int)((blk / tiles_c) % p.rb);
        for (int k = 0; k < R; ++k) tws[k] = ((const b2::cplx<T> *)p.tw)[(int64_t)o * k];
        for (int t = 0; t < B2_SPLIT_THREADS; ++t) split_thread<T, R, PHASE>(p, swap, blk, t, tws);
    }
}
template <typename T>
static int run_split_t(const b2d_split_pass &p, int swap)
{
    const int r = p.phase == 0 ? p.ra : p.rb;
    if (p.phase == 0) {
        if (r == 8) run_split<T, 8, 0>(p, swap); else if (r == 16) run_split<T, 16, 0>(p, swap);
        else if (r == 32) run_split<T, 32, 0>(p, swap); else return -1;
    } else {
        if (r == 8) run_split<T, 8, 1>(p, swap); else if (r == 16) run_split<T, 16, 1>(p, swap);
        else if (r == 32) run_split<T, 32, 1>(p, swap); else return -1;
    }
    return 0;
}

extern "C" {
int b2d_device_count(void) { return 1; }
const char *b2d_device_name(void) { return "emulated-for-unit-tests"; }
int b2d_sm_count(void) { return 148; }
const char *b2d_last_error(void) { return g_err; }
size_t b2d_max_smem_per_block(void) { return 232448; }
uint64_t b2d_launch_count(void) { return g_launches.load(); }
int b2d_current_device(void) { return 0; }
int b2d_pointer_is_device(const void *p)
{
    // interior pointers count too (cudaPointerGetAttributes resolves them on the real device)
    std::lock_guard<std::mutex> lk(g_dev_mu);
    auto it = g_dev.upper_bound((char *)p);
    if (it == g_dev.begin()) return 0;
    --it;
    return (const char *)p < it->first + it->second ? 1 : 0;
}
void *b2d_malloc(size_t n)
{
    char *p = (char *)malloc(n ? n : 1);
    std::lock_guard<std::mutex> lk(g_dev_mu);
    g_dev[p] = n ? n : 1;
    return p;
}
void b2d_free(void *p)
{
    if (!p) return;
    { std::lock_guard<std::mutex> lk(g_dev_mu); g_dev.erase((char *)p); }
    free(p);
}
void *b2d_malloc_host(size_t n) { void *p = NULL; if (posix_memalign(&p, 64, n ? n : 1)) return NULL; return p; }
void b2d_free_host(void *p) { free(p); }
int b2d_memcpy_h2d(void *d, const void *s, size_t n) { memcpy(d, s, n); return 0; }
int b2d_memcpy_d2h(void *d, const void *s, size_t n) { memcpy(d, s, n); return 0; }
int b2d_memcpy_d2h_async(void *d, const void *s, size_t n) { memcpy(d, s, n); return 0; }
int b2d_memcpy_d2d(void *d, const void *s, size_t n) { memmove(d, s, n); return 0; }
int b2d_memcpy2d_async(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, void *)
{
    for (size_t r = 0; r < height; ++r) memmove((char *)d + r * dpitch, (const char *)s + r * spitch, width);
    return 0;
}
int b2d_memset(void *d, int b, size_t n) { memset(d, b, n); return 0; }
int b2d_sync(void) { return 0; }
void b2d_set_stream(void *) {}
void *b2d_get_stream(void) { return NULL; }
void *b2d_push_stream(void *) { return NULL; }
void b2d_pop_stream(void *) {}
/* ranks of a unit test are THREADS of one process (tests/test_dist_comm_api.py): the same flag protocol as
   the CUDA kernel, on ordinary memory */
int b2d_peer_barrier(void *const *flags, int rank, int nranks, unsigned long long epoch)
{
    for (int d = 0; d < nranks; ++d)
        __atomic_store_n((unsigned long long *)flags[d] + rank, epoch, __ATOMIC_RELEASE);
    for (int d = 0; d < nranks; ++d)
        while (__atomic_load_n((unsigned long long *)flags[rank] + d, __ATOMIC_ACQUIRE) < epoch) sched_yield();
    return 0;
}
int b2d_partition_streams(int, void **, void **) { return -1; }   /* no SMs to partition here */
void *b2d_pipe_stream(int) { return NULL; }                /* chunks then simply run one after the other */
void *b2d_aux_stream(int) { return NULL; }                 /* everything is synchronous here */
int b2d_stream_wait_stream(void *, void *) { return 0; }
/* "IPC" between the threads that play ranks in a unit test: the handle is the pointer itself */
int b2d_ipc_export(void *p, unsigned char *h) { memset(h, 0, 64); memcpy(h, &p, sizeof p); return 0; }
void *b2d_ipc_import(const unsigned char *h) { void *p; memcpy(&p, h, sizeof p); return p; }
void b2d_ipc_close(void *) {}
int64_t b2d_alloc_offset(const void *p)
{
    std::lock_guard<std::mutex> lk(g_dev_mu);
    auto it = g_dev.upper_bound((char *)p);
    if (it == g_dev.begin()) return -1;
    --it;
    return (const char *)p < it->first + it->second ? (int64_t)((const char *)p - it->first) : -1;
}
int b2d_timer_start(void) { g_t0 = std::chrono::steady_clock::now(); return 0; }
int b2d_timer_stop(float *ms)
{
    *ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - g_t0).count();
    return 0;
}
size_t b2d_fft_pass_smem(const b2d_fft_pass *p)
{
    return p->prec == B2D_F32 ? b2::smem_bytes<float>(*p) : b2::smem_bytes<double>(*p);
}
int b2d_fast_available(const b2d_fft_pass *, int) { return 0; }   /* specialised kernels are GPU-only */
int b2d_launch_fft_pass(const b2d_fft_pass *p)
{
    if (b2d_fft_pass_smem(p) > b2d_max_smem_per_block()) {
        snprintf(g_err, sizeof g_err, "pass needs too much smem");
        return -1;
    }
    if (p->prec == B2D_F32) run_fft_pass<float>(*p); else run_fft_pass<double>(*p);
    ++g_launches;
    return 0;
}
int b2d_split_supported(int, int ra, int rb)
{
    return (ra == 8 || ra == 16 || ra == 32) && (rb == 8 || rb == 16 || rb == 32);
}
int b2d_launch_split_pass(const b2d_split_pass *p)
{
    const size_t rs = p->prec == B2D_F32 ? 4 : 8;
    const intptr_t d = (const char *)p->user_im - (const char *)p->user_re;
    if ((d != (intptr_t)rs && d != -(intptr_t)rs) || (p->row_stride & 1) || (p->bs & 1)) {
        snprintf(g_err, sizeof g_err, "split pass: layout not supported");
        return -1;
    }
    int rc = p->prec == B2D_F32 ? run_split_t<float>(*p, d < 0) : run_split_t<double>(*p, d < 0);
    ++g_launches;
    return rc;
}
int b2d_launch_copy(const b2d_copy *c)
{
    int64_t total = c->n[0] * c->n[1] * c->n[2] * c->n[3];
    for (int64_t i = 0; i < total; ++i) {
        if (c->prec == B2D_F32) b2::copy_elem<float>(*c, i); else b2::copy_elem<double>(*c, i);
    }
    ++g_launches;
    return 0;
}
int b2d_launch_realop(const b2d_realop *r)
{
    if (r->prec == B2D_F32) run_realop<float>(*r); else run_realop<double>(*r);
    ++g_launches;
    return 0;
}
}
