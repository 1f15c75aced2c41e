// shim.cu -- implementation of the C-ABI in include/b200fft_device.h on CUDA.
// Everything the C host layer needs from the device goes through here.
// There is NO CPU fallback: without a usable device every entry point fails.
#include <cuda.h>            /* types of the green-context entry points only: no link dependency on libcuda */
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>

#include "fft_generic.cuh"
#include "real_ops.cuh"
#include "fft_fast.cuh"
#include "fft_split.cuh"

namespace b2split {
int launch(const b2d_split_pass &p, cudaStream_t st);
int supported(int prec, int ra, int rb);
}

namespace {
char g_err[512] = "";
cudaStream_t g_user_stream = 0;               // set by fftw_b200_set_stream: process-wide
// per-thread override (side streams of one execute: plan lanes, the distributed stages): another
// thread that executes meanwhile keeps launching on the user's stream
thread_local cudaStream_t t_stream = 0;
thread_local int t_stream_depth = 0;
#define g_stream (t_stream_depth > 0 ? t_stream : g_user_stream)
int g_init = 0, g_ndev = 0, g_sms = 0;
size_t g_max_smem = 0;
char g_name[256] = "";
cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
std::atomic<uint64_t> g_launches{0};
int g_generic_max_threads[2][2] = { { 1024, 1024 }, { 1024, 1024 } };   // [f32?][plain?]: register-limited block size

int fail(cudaError_t e, const char *what)
{
    snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    return -1;
}

int ensure_init()
{
    if (g_init) return g_ndev > 0 ? 0 : -1;
    g_init = 1;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        snprintf(g_err, sizeof g_err, "no CUDA device: %s", cudaGetErrorString(e));
        g_ndev = 0;
        return -1;
    }
    g_ndev = n;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    strncpy(g_name, prop.name, sizeof g_name - 1);
    g_sms = prop.multiProcessorCount;
    g_max_smem = prop.sharedMemPerBlockOptin;
    cudaFuncSetAttribute(b2::fft_generic_kernel<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_max_smem);
    cudaFuncSetAttribute(b2::fft_generic_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_max_smem);
    cudaFuncSetAttribute(b2::fft_generic_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_max_smem);
    cudaFuncSetAttribute(b2::fft_generic_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_max_smem);
    {
        // the generic kernel is register-heavy: a block of tpb * tpx threads may not be launchable
        // (e.g. 648 threads x 104 registers); its loops stride by blockDim, so the launch clamps
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, b2::fft_generic_kernel<double, false>) == cudaSuccess) g_generic_max_threads[0][0] = fa.maxThreadsPerBlock;
        if (cudaFuncGetAttributes(&fa, b2::fft_generic_kernel<double, true>) == cudaSuccess) g_generic_max_threads[0][1] = fa.maxThreadsPerBlock;
        if (cudaFuncGetAttributes(&fa, b2::fft_generic_kernel<float, false>) == cudaSuccess) g_generic_max_threads[1][0] = fa.maxThreadsPerBlock;
        if (cudaFuncGetAttributes(&fa, b2::fft_generic_kernel<float, true>) == cudaSuccess) g_generic_max_threads[1][1] = fa.maxThreadsPerBlock;
        cudaGetLastError();
    }
    b2fast::init((int)g_max_smem);
    return 0;
}
}  // namespace

namespace {
struct BarrierArgs { unsigned long long *flags[B2D_MAX_PEERS]; int rank, nranks; unsigned long long epoch; };

__global__ void peer_barrier_kernel(BarrierArgs a)
{
    const int d = threadIdx.x;
    if (d >= a.nranks) return;
    __threadfence_system();                                  // this stream's earlier (remote) stores first
    unsigned long long *dst = a.flags[d] + a.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(a.epoch) : "memory");
    const unsigned long long *mine = a.flags[a.rank] + d;
    unsigned long long seen;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
    } while (seen < a.epoch);
}
}  // namespace

extern "C" {

int b2d_peer_barrier(void *const *flags, int rank, int nranks, unsigned long long epoch)
{
    if (ensure_init()) return -1;
    if (nranks < 1 || nranks > B2D_MAX_PEERS || rank < 0 || rank >= nranks) return -1;
    if (nranks == 1) return 0;
    BarrierArgs a;
    for (int i = 0; i < nranks; ++i) a.flags[i] = (unsigned long long *)flags[i];
    a.rank = rank; a.nranks = nranks; a.epoch = epoch;
    peer_barrier_kernel<<<1, 32, 0, g_stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(e, "peer_barrier_kernel launch");
    g_launches++;
    return 0;
}

int b2d_device_count(void) { ensure_init(); return g_ndev; }
const char *b2d_device_name(void) { ensure_init(); return g_name; }
int b2d_sm_count(void) { ensure_init(); return g_sms; }
const char *b2d_last_error(void) { return g_err; }
size_t b2d_max_smem_per_block(void) { ensure_init(); return g_max_smem; }
uint64_t b2d_launch_count(void) { return g_launches.load(); }

int b2d_current_device(void)
{
    int dev = -1;
    if (ensure_init() || cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    return dev;
}

int b2d_pointer_is_device(const void *p)
{
    if (ensure_init()) return -1;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? 1 : 0;
}

void *b2d_malloc(size_t bytes)
{
    if (ensure_init()) return nullptr;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { fail(e, "cudaMalloc"); return nullptr; }
    return p;
}
void b2d_free(void *p) { if (p) cudaFree(p); }

void *b2d_malloc_host(size_t bytes)
{
    if (ensure_init()) return nullptr;
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { fail(e, "cudaMallocHost"); cudaGetLastError(); return nullptr; }
    return p;
}
void b2d_free_host(void *p) { if (p) cudaFreeHost(p); }

int b2d_memcpy_h2d(void *d, const void *s, size_t n)
{
    cudaError_t e = cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, g_stream);
    return e == cudaSuccess ? 0 : fail(e, "memcpy h2d");
}
int b2d_memcpy_d2h(void *d, const void *s, size_t n)
{
    cudaError_t e = cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, g_stream);
    if (e != cudaSuccess) return fail(e, "memcpy d2h");
    e = cudaStreamSynchronize(g_stream);
    return e == cudaSuccess ? 0 : fail(e, "memcpy d2h sync");
}
int b2d_memcpy_d2h_async(void *d, const void *s, size_t n)
{
    cudaError_t e = cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, g_stream);
    return e == cudaSuccess ? 0 : fail(e, "memcpy d2h");
}
int b2d_memcpy_d2d(void *d, const void *s, size_t n)
{
    cudaError_t e = cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, g_stream);
    return e == cudaSuccess ? 0 : fail(e, "memcpy d2d");
}
int b2d_memcpy2d_async(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (!width || !height) return 0;
    if (width == dpitch && width == spitch)
        e = cudaMemcpyAsync(d, s, width * height, cudaMemcpyDeviceToDevice, st);
    else if (dpitch < ((size_t)1 << 31) && spitch < ((size_t)1 << 31))
        e = cudaMemcpy2DAsync(d, dpitch, s, spitch, width, height, cudaMemcpyDeviceToDevice, st);
    else
        for (size_t r = 0; r < height && e == cudaSuccess; ++r)
            e = cudaMemcpyAsync((char *)d + r * dpitch, (const char *)s + r * spitch, width, cudaMemcpyDeviceToDevice, st);
    return e == cudaSuccess ? 0 : fail(e, "memcpy 2d");
}
int b2d_memset(void *d, int byte, size_t n)
{
    cudaError_t e = cudaMemsetAsync(d, byte, n, g_stream);
    return e == cudaSuccess ? 0 : fail(e, "memset");
}
int b2d_sync(void)
{
    cudaError_t e = cudaStreamSynchronize(g_stream);
    if (e != cudaSuccess) return fail(e, "sync");
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail(e, "kernel");
}
void b2d_set_stream(void *s) { g_user_stream = (cudaStream_t)s; }
void *b2d_get_stream(void) { return (void *)g_stream; }
void *b2d_push_stream(void *s)
{
    void *prev = (void *)g_stream;
    t_stream = (cudaStream_t)s;
    ++t_stream_depth;
    return prev;
}
void b2d_pop_stream(void *prev)
{
    if (t_stream_depth > 0) --t_stream_depth;
    t_stream = (cudaStream_t)prev;
}

void *b2d_pipe_stream(int idx)
{
    static cudaStream_t ps[3] = { nullptr, nullptr, nullptr };
    if (ensure_init() || idx < 0 || idx >= 3) return nullptr;
    if (!ps[idx] && cudaStreamCreateWithFlags(&ps[idx], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return (void *)ps[idx];
}

void *b2d_aux_stream(int idx)
{
    static cudaStream_t aux[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    if (ensure_init() || idx < 0 || idx >= 8) return nullptr;
    if (!aux[idx]) {
        int lo = 0, hi = 0;                 /* hi = numerically lowest = greatest priority */
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&aux[idx], cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;
    }
    return (void *)aux[idx];
}

static void *driver_entry_raw(const char *name)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return fn;
}
#define driver_entry_as(T, name) ((T)driver_entry_raw(name))

int b2d_partition_streams(int comm_sms, void **comm_stream, void **compute_stream)
{
    struct Part { int dev, sms, ok; cudaStream_t comm, comp; };
    static Part parts[8];
    static int nparts = 0;
    static std::mutex mu;
    if (ensure_init() || comm_sms < 8) return -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    std::lock_guard<std::mutex> lk(mu);
    for (int i = 0; i < nparts; ++i)
        if (parts[i].dev == dev && parts[i].sms == comm_sms) {
            if (!parts[i].ok) return -1;
            *comm_stream = parts[i].comm; *compute_stream = parts[i].comp;
            return 0;
        }
    if (nparts == 8) return -1;
    Part &P = parts[nparts++];
    P.dev = dev; P.sms = comm_sms; P.ok = 0;
    auto getres = driver_entry_as(CUresult (*)(CUdevice, CUdevResource *, CUdevResourceType), "cuDeviceGetDevResource");
    auto split = driver_entry_as(CUresult (*)(CUdevResource *, unsigned *, const CUdevResource *, CUdevResource *, unsigned, unsigned), "cuDevSmResourceSplitByCount");
    auto gendesc = driver_entry_as(CUresult (*)(CUdevResourceDesc *, CUdevResource *, unsigned), "cuDevResourceGenerateDesc");
    auto gcreate = driver_entry_as(CUresult (*)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned), "cuGreenCtxCreate");
    auto gstream = driver_entry_as(CUresult (*)(CUstream *, CUgreenCtx, unsigned, int), "cuGreenCtxStreamCreate");
    auto devget = driver_entry_as(CUresult (*)(CUdevice *, int), "cuDeviceGet");
    if (!getres || !split || !gendesc || !gcreate || !gstream || !devget) return -1;
    CUdevice cudev;
    CUdevResource all, part, rest;
    CUdevResourceDesc d0, d1;
    CUgreenCtx g0, g1;
    CUstream s0, s1;
    unsigned n = 1;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (devget(&cudev, dev) || getres(cudev, &all, CU_DEV_RESOURCE_TYPE_SM)) return -1;
    if ((int)all.sm.smCount < comm_sms + 8) return -1;
    if (split(&part, &n, &all, &rest, 0, (unsigned)comm_sms) || n != 1 || rest.sm.smCount == 0) return -1;
    if (gendesc(&d0, &part, 1) || gendesc(&d1, &rest, 1)) return -1;
    if (gcreate(&g0, d0, cudev, CU_GREEN_CTX_DEFAULT_STREAM) || gcreate(&g1, d1, cudev, CU_GREEN_CTX_DEFAULT_STREAM)) return -1;
    if (gstream(&s0, g0, CU_STREAM_NON_BLOCKING, hi) || gstream(&s1, g1, CU_STREAM_NON_BLOCKING, 0)) return -1;
    P.comm = (cudaStream_t)s0; P.comp = (cudaStream_t)s1; P.ok = 1;
    *comm_stream = P.comm; *compute_stream = P.comp;
    return 0;
}

int b2d_stream_wait_stream(void *waiter, void *signaler)
{
    cudaEvent_t ev;
    cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) return fail(e, "event create");
    e = cudaEventRecord(ev, (cudaStream_t)signaler);
    if (e == cudaSuccess) e = cudaStreamWaitEvent((cudaStream_t)waiter, ev, 0);
    cudaEventDestroy(ev);                  /* released once the recorded work completes */
    return e == cudaSuccess ? 0 : fail(e, "stream wait");
}

int b2d_ipc_export(void *devptr, unsigned char handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, devptr);
    if (e != cudaSuccess) return fail(e, "cudaIpcGetMemHandle");
    memcpy(handle, &h, 64);
    return 0;
}
/* One mapping per exported allocation and process: several plans may map the same peer slab (a forward and a
   backward plan on one array), but a handle can be opened only once -- imports are cached and refcounted. */
namespace {
struct IpcMap { unsigned char h[64]; void *ptr; int refs; };
IpcMap g_ipc[256];
int g_nipc = 0;
std::mutex g_ipc_mu;
}
void *b2d_ipc_import(const unsigned char handle[64])
{
    if (ensure_init()) return nullptr;
    std::lock_guard<std::mutex> lk(g_ipc_mu);
    for (int i = 0; i < g_nipc; ++i)
        if (g_ipc[i].refs > 0 && !memcmp(g_ipc[i].h, handle, 64)) { g_ipc[i].refs++; return g_ipc[i].ptr; }
    cudaIpcMemHandle_t h;
    void *p = nullptr;
    memcpy(&h, handle, 64);
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { fail(e, "cudaIpcOpenMemHandle"); cudaGetLastError(); return nullptr; }
    int slot = -1;
    for (int i = 0; i < g_nipc; ++i) if (g_ipc[i].refs == 0) { slot = i; break; }
    if (slot < 0 && g_nipc < 256) slot = g_nipc++;
    if (slot >= 0) { memcpy(g_ipc[slot].h, handle, 64); g_ipc[slot].ptr = p; g_ipc[slot].refs = 1; }
    return p;
}
void b2d_ipc_close(void *devptr)
{
    if (!devptr) return;
    std::lock_guard<std::mutex> lk(g_ipc_mu);
    for (int i = 0; i < g_nipc; ++i)
        if (g_ipc[i].refs > 0 && g_ipc[i].ptr == devptr) {
            if (--g_ipc[i].refs == 0) cudaIpcCloseMemHandle(devptr);
            return;
        }
    cudaIpcCloseMemHandle(devptr);
}
int64_t b2d_alloc_offset(const void *devptr)
{
    // cuMemGetAddressRange through the runtime's driver entry point (the library does not link libcuda)
    typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
    static range_fn fn = nullptr;
    if (ensure_init()) return -1;
    if (!fn) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return -1; }
        fn = (range_fn)f;
    }
    unsigned long long base = 0;
    size_t size = 0;
    if (fn(&base, &size, (unsigned long long)(uintptr_t)devptr) != 0) return -1;
    return (int64_t)((unsigned long long)(uintptr_t)devptr - base);
}

int b2d_timer_start(void)
{
    if (ensure_init()) return -1;
    if (!g_ev0) { cudaEventCreate(&g_ev0); cudaEventCreate(&g_ev1); }
    cudaError_t e = cudaEventRecord(g_ev0, g_stream);
    return e == cudaSuccess ? 0 : fail(e, "event record");
}
int b2d_timer_stop(float *ms)
{
    cudaError_t e = cudaEventRecord(g_ev1, g_stream);
    if (e != cudaSuccess) return fail(e, "event record");
    e = cudaEventSynchronize(g_ev1);
    if (e != cudaSuccess) return fail(e, "event sync");
    cudaEventElapsedTime(ms, g_ev0, g_ev1);
    return 0;
}

size_t b2d_fft_pass_smem(const b2d_fft_pass *p)
{
    size_t fast = b2fast::smem_bytes(*p);
    if (fast) return fast;
    return p->prec == B2D_F32 ? b2::smem_bytes<float>(*p) : b2::smem_bytes<double>(*p);
}

int b2d_fast_available(const b2d_fft_pass *p, int code)
{
    if (ensure_init()) return 0;
    return b2fast::available(*p, code);
}

int b2d_launch_fft_pass(const b2d_fft_pass *p)
{
    if (ensure_init()) return -1;
    int64_t blocks = b2::grid_blocks(*p);
    if (blocks <= 0) return 0;
    if (blocks > 2147483647LL) { snprintf(g_err, sizeof g_err, "grid too large"); return -1; }
    int rc = b2fast::try_launch(*p, g_stream);
    if (rc == 0) { g_launches++; return 0; }
    if (rc < 0) { snprintf(g_err, sizeof g_err, "fast kernel launch failed"); return -1; }
    size_t smem = p->prec == B2D_F32 ? b2::smem_bytes<float>(*p) : b2::smem_bytes<double>(*p);
    if (smem > g_max_smem) { snprintf(g_err, sizeof g_err, "pass needs %zu B smem", smem); return -1; }
    int threads = p->tpb * p->tpx;
    if (threads < 32) threads = 32;
    if (threads > 1024) threads = 1024;
    int swi = 0, swo = 0;
    const bool plain = b2::plain_ok(*p, &swi, &swo);
    {
        int cap = g_generic_max_threads[p->prec == B2D_F32 ? 1 : 0][plain ? 1 : 0];
        if (threads > cap) {
            // whole threads-per-transform groups: the stage loops map thread -> (transform, butterfly) by blockDim / tpb
            int tpx = cap / (p->tpb > 0 ? p->tpb : 1);
            threads = tpx >= 1 ? tpx * p->tpb : cap;
            if (threads < 32) threads = cap;
        }
    }
    if (p->prec == B2D_F32) {
        if (plain) b2::fft_generic_kernel<float, true><<<(unsigned)blocks, threads, smem, g_stream>>>(*p, swi, swo);
        else b2::fft_generic_kernel<float, false><<<(unsigned)blocks, threads, smem, g_stream>>>(*p, 0, 0);
    } else {
        if (plain) b2::fft_generic_kernel<double, true><<<(unsigned)blocks, threads, smem, g_stream>>>(*p, swi, swo);
        else b2::fft_generic_kernel<double, false><<<(unsigned)blocks, threads, smem, g_stream>>>(*p, 0, 0);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(e, "fft_generic_kernel launch");
    g_launches++;
    return 0;
}

int b2d_split_supported(int prec, int ra, int rb)
{
    if (ensure_init()) return 0;
    return b2split::supported(prec, ra, rb);
}

int b2d_launch_split_pass(const b2d_split_pass *p)
{
    if (ensure_init()) return -1;
    int rc = b2split::launch(*p, g_stream);
    if (rc) { snprintf(g_err, sizeof g_err, rc > 0 ? "split pass: layout not supported" : "split_kernel launch failed"); return -1; }
    g_launches++;
    return 0;
}

int b2d_launch_copy(const b2d_copy *c)
{
    if (ensure_init()) return -1;
    int64_t total = c->n[0] * c->n[1] * c->n[2] * c->n[3];
    if (total <= 0) return 0;
    {
        /* transposing copy?  the input-contiguous and output-contiguous dims differ and all
           offsets keep whole elements aligned -> 32x32 tiles through shared memory */
        int da = -1, db = -1, dc = -1, dd = -1, i;
        const int es = c->elem_reals;
        const size_t vs = (size_t)es * (c->prec == B2D_F32 ? 4 : 8);
        int ok = !c->npeer;
        for (i = 0; i < 4; ++i) {
            if (c->n[i] <= 1) continue;
            if (llabs(c->is[i]) == es && da < 0) da = i;
            if (llabs(c->os[i]) == es && db < 0) db = i;
            if ((c->is[i] % es) || (c->os[i] % es)) ok = 0;
        }
        if (((uintptr_t)c->in % vs) || ((uintptr_t)c->out % vs)) ok = 0;
        if (ok && da >= 0 && db >= 0 && da != db && c->n[da] >= 16 && c->n[db] >= 16 &&
            c->is[da] > 0 && c->os[db] > 0) {
            for (i = 0; i < 4; ++i) if (i != da && i != db) { if (dc < 0) dc = i; else dd = i; }
            const int ts = (vs == 16 || c->n[da] < 64 || c->n[db] < 64) ? 32 : 64;
            int64_t ta = (c->n[da] + ts - 1) / ts, tb = (c->n[db] + ts - 1) / ts;
            int64_t nblk = ta * tb * c->n[dc] * c->n[dd];
            if (nblk <= 2147483647LL) {
                if (vs == 4 && ts == 64) b2::transpose_kernel<float, 64><<<(unsigned)nblk, 256, 0, g_stream>>>(*c, da, db, dc, dd, ta, tb);
                else if (vs == 4) b2::transpose_kernel<float, 32><<<(unsigned)nblk, 256, 0, g_stream>>>(*c, da, db, dc, dd, ta, tb);
                else if (vs == 8 && ts == 64) b2::transpose_kernel<double, 64><<<(unsigned)nblk, 256, 0, g_stream>>>(*c, da, db, dc, dd, ta, tb);
                else if (vs == 8) b2::transpose_kernel<double, 32><<<(unsigned)nblk, 256, 0, g_stream>>>(*c, da, db, dc, dd, ta, tb);
                else b2::transpose_kernel<double2, 32><<<(unsigned)nblk, 256, 0, g_stream>>>(*c, da, db, dc, dd, ta, tb);
                cudaError_t e2 = cudaGetLastError();
                if (e2 != cudaSuccess) return fail(e2, "transpose_kernel launch");
                g_launches++;
                return 0;
            }
        }
    }
    int64_t blocks = (total + 4 * 256 - 1) / (4 * 256);
    if (c->grid_limit > 0 && blocks > c->grid_limit) blocks = c->grid_limit;
    if (blocks > 2147483647LL) { snprintf(g_err, sizeof g_err, "copy grid too large"); return -1; }
    if (c->prec == B2D_F32) b2::copy_kernel<float><<<(unsigned)blocks, 256, 0, g_stream>>>(*c, total);
    else b2::copy_kernel<double><<<(unsigned)blocks, 256, 0, g_stream>>>(*c, total);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(e, "copy_kernel launch");
    g_launches++;
    return 0;
}

int b2d_launch_realop(const b2d_realop *r)
{
    if (ensure_init()) return -1;
    int len;
    int kind = r->op & 15;
    if (r->op == B2D_ROP_R2C_POST) len = r->m / 2 + 1;
    else if (r->op == B2D_ROP_C2R_PRE) len = r->m;
    else if (r->op >= B2D_ROP_BLUE_PRE && r->op <= B2D_ROP_BLUE_POST) len = (r->op == B2D_ROP_BLUE_POST) ? r->n_lim : r->m;
    else if (r->op & B2D_ROP_R2R_POST) len = r->n;
    else len = b2::r2r_work_len(kind, r->n);
    int64_t nb = r->bn[0] * r->bn[1] * r->bn[2];
    if (nb <= 0 || len <= 0) return 0;
    int chunks = (len + 127) / 128;
    int64_t blocks = (int64_t)chunks * nb;
    if (blocks > 2147483647LL) { snprintf(g_err, sizeof g_err, "realop grid too large"); return -1; }
    if (r->prec == B2D_F32) b2::realop_kernel<float><<<(unsigned)blocks, 128, 0, g_stream>>>(*r, len, chunks);
    else b2::realop_kernel<double><<<(unsigned)blocks, 128, 0, g_stream>>>(*r, len, chunks);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(e, "realop_kernel launch");
    g_launches++;
    return 0;
}

}  // extern "C"
