// fft_tma_table.cu -- instantiations, tensor-map encoding and launch of the TMA-fed
// strided kernels (fft_tma.cuh).  Kernel code = 6000 + 100 * l2_promotion + tile width
// (l2_promotion 0 none, 1 = 128 B, 2 = 256 B; +3: a CTA takes adjacent tiles in pairs).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "fft_tma.cuh"

namespace b2tma {

static const TmaEntry tma_table[] = {
    //            prec     T       N     E   R1  R2  TPB G  NBUF
    B2_TMA_ENTRY(B2D_F64, double, 1024, 16, 16, 4,  4,  2, 3),
    B2_TMA_ENTRY(B2D_F64, double, 512,  8,  8,  8,  4,  2, 6),
    B2_TMA_ENTRY(B2D_F64, double, 512,  8,  8,  8,  8,  2, 3),
    B2_TMA_ENTRY(B2D_F32, float,  1024, 16, 16, 4,  8,  2, 3),
    B2_TMA_ENTRY(B2D_F32, float,  512,  8,  8,  8,  8,  2, 6),
};
static const int tma_table_count = (int)(sizeof(tma_table) / sizeof(tma_table[0]));

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_fn g_encode = nullptr;
static int g_max_smem = 0, g_sms = 148;

void init(int max_smem)
{
    g_max_smem = max_smem;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) g_sms = sms;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        g_encode = (encode_fn)fn;
    else
        (void)cudaGetLastError();
    if (getenv("FFTW3_B200_VERBOSE"))
        fprintf(stderr, "[b200 shim] tensor-map encoder %s (query result %d)\n", g_encode ? "found" : "missing", (int)qres);
    for (int i = 0; i < tma_table_count; ++i)
        cudaFuncSetAttribute(tma_table[i].func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_table[i].smem);
}

const TmaEntry *entry_for(const b2d_fft_pass &p)
{
    if (p.kernel < 6000 || p.kernel >= 7000 || !g_encode) return nullptr;
    const int tile = p.kernel % 100, flavor = (p.kernel / 100) % 10;
    if (flavor > 5) return nullptr;
    if (p.pre_op || p.post_op || p.bluestein || p.npeer) return nullptr;
    if (!p.load_col || !p.store_col || p.bis[0] != 2 || p.bos[0] != 2) return nullptr;
    const int64_t rs = p.prec == B2D_F32 ? 4 : 8;
    // tensor-map limits: strides multiples of 16 bytes and < 2^40, extents < 2^32, non-negative strides
    if (p.is <= 0 || (p.is * rs) % 16 || p.is * rs >= (1LL << 40) || (p.os & 1)) return nullptr;
    if (p.os <= 0 || (int64_t)p.n * (p.os / 2) >= (1LL << 32)) return nullptr;     // 32-bit store offsets within a pencil
    if (p.bn[0] * 2 >= (1LL << 32) || p.bn[1] >= (1LL << 31) || p.bn[2] >= (1LL << 31)) return nullptr;
    for (int i = 1; i < B2D_MAX_BATCH_DIMS; ++i) {
        if (p.bos[i] & 1) return nullptr;
        if (p.bn[i] > 1 && (p.bis[i] <= 0 || (p.bis[i] * rs) % 16 || p.bis[i] * rs >= (1LL << 40))) return nullptr;
    }
    for (int i = 0; i < tma_table_count; ++i) {
        const TmaEntry &e = tma_table[i];
        if (e.prec == p.prec && e.n == p.n && e.tpb == tile && (!g_max_smem || (int)e.smem <= g_max_smem)) return &e;
    }
    return nullptr;
}

int try_launch(const b2d_fft_pass &p, cudaStream_t st)
{
    const TmaEntry *e = entry_for(p);
    if (!e) return 1;
    const size_t rs = p.prec == B2D_F32 ? 4 : 8;
    const intptr_t din = (const char *)p.in_im - (const char *)p.in_re;
    const intptr_t dout = (char *)p.out_im - (char *)p.out_re;
    if ((din != (intptr_t)rs && din != -(intptr_t)rs) || (dout != (intptr_t)rs && dout != -(intptr_t)rs)) return 1;
    const int swap_in = din < 0, swap_out = dout < 0;
    void *gin = (void *)(swap_in ? p.in_im : p.in_re);
    if (((uintptr_t)gin % 16) || ((uintptr_t)(swap_out ? p.out_im : p.out_re) % (2 * rs))) return 1;

    b2d_fft_pass q = p;
    q.tpb = e->tpb;
    const int64_t tiles = b2::grid_blocks(q);
    if (tiles <= 0) return 0;

    // tensor of reals: [b2][b1][k][2 * pencil], innermost contiguous
    const int promo = ((p.kernel / 100) % 10) % 3, pair = ((p.kernel / 100) % 10) >= 3 ? 2 : 1;
    if (tiles % pair) return 1;
    cuuint64_t dims[4] = { (cuuint64_t)(2 * p.bn[0]), (cuuint64_t)p.n, (cuuint64_t)p.bn[1], (cuuint64_t)p.bn[2] };
    cuuint64_t strides[3] = { (cuuint64_t)p.is * rs, (cuuint64_t)(p.bn[1] > 1 ? p.bis[1] : p.is) * rs,
                              (cuuint64_t)(p.bn[2] > 1 ? p.bis[2] : p.is) * rs };
    cuuint32_t box[4] = { (cuuint32_t)(2 * e->tpb), (cuuint32_t)e->boxk, 1, 1 };
    cuuint32_t estr[4] = { 1, 1, 1, 1 };
    CUtensorMap map;
    CUresult rc = g_encode(&map, p.prec == B2D_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, gin,
                           dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                      : (promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE),
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return 1;                       // shape the engine cannot describe: generic kernel
    const int64_t blocks = tiles / pair < g_sms ? tiles / pair : g_sms;    // one persistent CTA per SM
    e->launch(q, map, swap_in, swap_out, pair, (unsigned)blocks, st);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace b2tma
