// fft_warp_table.cu -- instantiation and launch of the warp-per-transform kernels (fft_warp.cuh).
#include <stdint.h>
#include "fft_warp.cuh"

namespace b2warp {

static int g_max_smem = 0, g_sms = 148;

void init(int max_smem, int sms)
{
    g_max_smem = max_smem;
    if (sms > 0) g_sms = sms;
    cudaFuncSetAttribute(warp1024_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<double>());
    cudaFuncSetAttribute(warp1024_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<float>());
}

int applicable(const b2d_fft_pass &p)
{
    if (p.n != N || p.pre_op || p.post_op || p.bluestein || p.npeer || p.load_col || p.store_col) return 0;
    if (p.is != 2 || p.os != 2) return 0;
    for (int i = 0; i < B2D_MAX_BATCH_DIMS; ++i) if ((p.bis[i] & 1) || (p.bos[i] & 1)) return 0;
    if ((int)(p.prec == B2D_F32 ? smem_bytes<float>() : smem_bytes<double>()) > g_max_smem && g_max_smem) return 0;
    return 1;
}

int launch(const b2d_fft_pass &p, cudaStream_t st)
{
    if (!applicable(p)) return 1;
    const size_t rs = p.prec == B2D_F32 ? 4 : 8;
    const intptr_t din = (const char *)p.in_im - (const char *)p.in_re, dout = (char *)p.out_im - (char *)p.out_re;
    if ((din != (intptr_t)rs && din != -(intptr_t)rs) || (dout != (intptr_t)rs && dout != -(intptr_t)rs)) return 1;
    const int swap_in = din < 0, swap_out = dout < 0;
    if (((uintptr_t)(swap_in ? p.in_im : p.in_re) % (2 * rs)) || ((uintptr_t)(swap_out ? p.out_im : p.out_re) % (2 * rs))) return 1;
    const long long ntrans = (long long)p.bn[0] * p.bn[1] * p.bn[2];
    if (ntrans <= 0) return 0;
    long long grid = (ntrans + WARPS - 1) / WARPS;
    const long long resident = (long long)g_sms * 2;                  // two CTAs per SM by registers and shared memory
    // persistent CTAs, a few transforms per warp: amortises the per-CTA twiddle table
    if (grid > resident * 4) grid = resident * 4;
    if (p.grid_limit > 0 && grid > p.grid_limit) grid = p.grid_limit;
    if (p.prec == B2D_F32) warp1024_kernel<float><<<(unsigned)grid, WARPS * 32, smem_bytes<float>(), st>>>(p, swap_in, swap_out, ntrans);
    else warp1024_kernel<double><<<(unsigned)grid, WARPS * 32, smem_bytes<double>(), st>>>(p, swap_in, swap_out, ntrans);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace b2warp
