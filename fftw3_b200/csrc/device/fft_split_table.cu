// fft_split_table.cu -- instantiations and launch of the register-only sub-pass kernels (fft_split.cuh).
#include <cuda_runtime.h>
#include "fft_split.cuh"

namespace b2split {

template <typename T, int R, int PHASE>
static int launch_one(const b2d_split_pass &p, int swap, unsigned blocks, cudaStream_t st)
{
    split_kernel<T, R, PHASE><<<blocks, B2_SPLIT_THREADS, 0, st>>>(p, swap);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <typename T>
static int launch_t(const b2d_split_pass &p, int swap, unsigned blocks, cudaStream_t st)
{
    const int r = p.phase == 0 ? p.ra : p.rb;
    if (p.phase == 0) {
        if (r == 8) return launch_one<T, 8, 0>(p, swap, blocks, st);
        if (r == 16) return launch_one<T, 16, 0>(p, swap, blocks, st);
        if (r == 32) return launch_one<T, 32, 0>(p, swap, blocks, st);
    } else {
        if (r == 8) return launch_one<T, 8, 1>(p, swap, blocks, st);
        if (r == 16) return launch_one<T, 16, 1>(p, swap, blocks, st);
        if (r == 32) return launch_one<T, 32, 1>(p, swap, blocks, st);
    }
    return -1;
}

int supported(int prec, int ra, int rb)
{
    (void)prec;
    return (ra == 8 || ra == 16 || ra == 32) && (rb == 8 || rb == 16 || rb == 32);
}

// returns 0 launched, -1 error, 1 not applicable (layout: the caller planned it, so this is an error there)
int launch(const b2d_split_pass &p, cudaStream_t st)
{
    if (!supported(p.prec, p.ra, p.rb)) return 1;
    const size_t rs = p.prec == B2D_F32 ? 4 : 8;
    const intptr_t d = (const char *)p.user_im - (const char *)p.user_re;
    if (d != (intptr_t)rs && d != -(intptr_t)rs) return 1;
    const int swap = d < 0;
    if (((uintptr_t)(swap ? p.user_im : p.user_re) % (2 * rs)) || ((uintptr_t)p.work % (2 * rs))) return 1;
    if ((p.row_stride & 1) || (p.bs & 1)) return 1;
    const int64_t blocks = split_blocks(p);
    if (blocks <= 0) return 0;
    if (blocks > 2147483647LL) return -1;
    return p.prec == B2D_F32 ? launch_t<float>(p, swap, (unsigned)blocks, st) : launch_t<double>(p, swap, (unsigned)blocks, st);
}

}  // namespace b2split
