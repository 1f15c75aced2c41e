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
    cudaFuncSetAttribute(col1024_kernel<double, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)col_smem_bytes<double, 4>());
    cudaFuncSetAttribute(col1024_kernel<double, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)col_smem_bytes<double, 8>());
    cudaFuncSetAttribute(col1024_kernel<float, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)col_smem_bytes<float, 8>());
    cudaFuncSetAttribute(col1024_kernel<float, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)col_smem_bytes<float, 16>());
}

static int col_tp(const b2d_fft_pass &p, int code)
{
    // pencils per CTA: 64 B / 128 B segments in either precision
    const int tp = code - 3100;
    if (p.prec == B2D_F64) return (tp == 4 || tp == 8) ? tp : 0;
    return (tp == 8 || tp == 16) ? tp : 0;
}

size_t smem_for(const b2d_fft_pass &p, int code)
{
    if (!applicable(p, code)) return 0;
    if (code == 3001) return p.prec == B2D_F32 ? smem_bytes<float>() : smem_bytes<double>();
    const int tp = col_tp(p, code);
    if (p.prec == B2D_F64) return tp == 4 ? col_smem_bytes<double, 4>() : col_smem_bytes<double, 8>();
    return tp == 8 ? col_smem_bytes<float, 8>() : col_smem_bytes<float, 16>();
}

int applicable(const b2d_fft_pass &p, int code)
{
    if (p.n != N || p.pre_op || p.post_op || p.bluestein || p.npeer) return 0;
    for (int i = 0; i < B2D_MAX_BATCH_DIMS; ++i) if ((p.bis[i] & 1) || (p.bos[i] & 1)) return 0;
    if (code == 3001) {
        if (p.load_col || p.store_col || p.is != 2 || p.os != 2) return 0;
        if ((int)(p.prec == B2D_F32 ? smem_bytes<float>() : smem_bytes<double>()) > g_max_smem && g_max_smem) return 0;
        return 1;
    }
    if (!col_tp(p, code)) return 0;
    if (!p.load_col || !p.store_col || p.bis[0] != 2 || p.bos[0] != 2 || (p.is & 1) || (p.os & 1)) return 0;
    return 1;
}

template <typename T, int TP>
static int launch_col(const b2d_fft_pass &p, int swap_in, int swap_out, cudaStream_t st)
{
    b2d_fft_pass q = p;
    q.tpb = TP;
    const long long ntiles = (long long)b2::grid_blocks(q);
    if (ntiles <= 0) return 0;
    long long grid = ntiles;
    const long long resident = (long long)g_sms * (TP * sizeof(b2::cplx<T>) < 128 ? 2 : 1);
    if (grid > resident * 8) grid = resident * 8;          // persistent CTAs: the twiddle table is loaded once per CTA
    if (p.grid_limit > 0 && grid > p.grid_limit) grid = p.grid_limit;
    col1024_kernel<T, TP><<<(unsigned)grid, 32 * TP, col_smem_bytes<T, TP>(), st>>>(q, swap_in, swap_out, ntiles);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch(const b2d_fft_pass &p, cudaStream_t st)
{
    if (!applicable(p, p.kernel)) return 1;
    const size_t rs = p.prec == B2D_F32 ? 4 : 8;
    const intptr_t din = (const char *)p.in_im - (const char *)p.in_re, dout = (char *)p.out_im - (char *)p.out_re;
    if ((din != (intptr_t)rs && din != -(intptr_t)rs) || (dout != (intptr_t)rs && dout != -(intptr_t)rs)) return 1;
    const int swap_in = din < 0, swap_out = dout < 0;
    if (((uintptr_t)(swap_in ? p.in_im : p.in_re) % (2 * rs)) || ((uintptr_t)(swap_out ? p.out_im : p.out_re) % (2 * rs))) return 1;
    if (p.kernel != 3001) {
        const int tp = col_tp(p, p.kernel);
        if (p.prec == B2D_F64) return tp == 4 ? launch_col<double, 4>(p, swap_in, swap_out, st) : launch_col<double, 8>(p, swap_in, swap_out, st);
        return tp == 8 ? launch_col<float, 8>(p, swap_in, swap_out, st) : launch_col<float, 16>(p, swap_in, swap_out, st);
    }
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
