// Probe: can this box split a B200's SMs into two green contexts and run runtime-API kernels on their streams
// concurrently?  (driver entry points fetched through the runtime: no link dependency on libcuda)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o greenctx_probe greenctx_probe.cu && ./greenctx_probe 56
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void spin(unsigned *smid_min, unsigned *smid_max, unsigned *count, long long cycles)
{
    unsigned id;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    if (threadIdx.x == 0) { atomicMin(smid_min, id); atomicMax(smid_max, id); atomicAdd(count, 1u); }
    long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
}

template <typename F> static F entry(const char *name)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) {
        printf("no entry point %s\n", name); exit(2);
    }
    return (F)fn;
}

int main(int argc, char **argv)
{
    int want = argc > 1 ? atoi(argv[1]) : 56;
    cudaFree(0);
    auto getres = entry<CUresult (*)(CUdevice, CUdevResource *, CUdevResourceType)>("cuDeviceGetDevResource");
    auto split = entry<CUresult (*)(CUdevResource *, unsigned *, const CUdevResource *, CUdevResource *, unsigned, unsigned)>("cuDevSmResourceSplitByCount");
    auto gendesc = entry<CUresult (*)(CUdevResourceDesc *, CUdevResource *, unsigned)>("cuDevResourceGenerateDesc");
    auto gcreate = entry<CUresult (*)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned)>("cuGreenCtxCreate");
    auto gstream = entry<CUresult (*)(CUstream *, CUgreenCtx, unsigned, int)>("cuGreenCtxStreamCreate");
    CUdevResource all, part, rest;
    unsigned n = 1;
    CUresult r;
    if ((r = getres(0, &all, CU_DEV_RESOURCE_TYPE_SM))) { printf("getres %d\n", r); return 1; }
    printf("device SMs %u\n", all.sm.smCount);
    if ((r = split(&part, &n, &all, &rest, 0, (unsigned)want))) { printf("split %d\n", r); return 1; }
    printf("split: groups %u, part %u SMs, rest %u SMs\n", n, part.sm.smCount, rest.sm.smCount);
    CUdevResourceDesc d0, d1;
    CUgreenCtx g0, g1;
    CUstream s0, s1;
    if ((r = gendesc(&d0, &part, 1)) || (r = gendesc(&d1, &rest, 1))) { printf("desc %d\n", r); return 1; }
    if ((r = gcreate(&g0, d0, 0, CU_GREEN_CTX_DEFAULT_STREAM)) || (r = gcreate(&g1, d1, 0, CU_GREEN_CTX_DEFAULT_STREAM))) { printf("create %d\n", r); return 1; }
    if ((r = gstream(&s0, g0, CU_STREAM_NON_BLOCKING, 0)) || (r = gstream(&s1, g1, CU_STREAM_NON_BLOCKING, 0))) { printf("stream %d\n", r); return 1; }
    unsigned *st;
    cudaMalloc(&st, 6 * sizeof(unsigned));
    unsigned init[6] = { 1000, 0, 0, 1000, 0, 0 };
    cudaMemcpy(st, init, sizeof init, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    cudaStream_t main_s;
    cudaStreamCreateWithFlags(&main_s, cudaStreamNonBlocking);
    // main stream -> both partitions -> main stream, with runtime events (what dist.c would do)
    cudaEventRecord(e0, main_s);
    cudaStreamWaitEvent((cudaStream_t)s0, e0, 0);
    cudaStreamWaitEvent((cudaStream_t)s1, e0, 0);
    spin<<<2000, 128, 0, (cudaStream_t)s0>>>(st, st + 1, st + 2, 2000000);     // ~1 ms each wave
    spin<<<2000, 128, 0, (cudaStream_t)s1>>>(st + 3, st + 4, st + 5, 2000000);
    cudaEventRecord(e1, (cudaStream_t)s0);
    cudaEventRecord(e2, (cudaStream_t)s1);
    cudaStreamWaitEvent(main_s, e1, 0);
    cudaStreamWaitEvent(main_s, e2, 0);
    cudaError_t ce = cudaStreamSynchronize(main_s);
    printf("sync: %s\n", cudaGetErrorString(ce));
    unsigned out[6];
    cudaMemcpy(out, st, sizeof out, cudaMemcpyDeviceToHost);
    printf("partition 0: smid %u..%u (%u CTAs)   partition 1: smid %u..%u (%u CTAs)\n", out[0], out[1], out[2], out[3], out[4], out[5]);
    return 0;
}
