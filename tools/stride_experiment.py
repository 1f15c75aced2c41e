#!/usr/bin/env python
"""Where does the strided pass fall off?  Same COL kernel, pencil stride swept from 1 KiB to 16 MiB."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fftw3_b200 import binding as B
lib = B.load()
n = 1024
total = n ** 3
x = torch.zeros(total, 2, dtype=torch.float64, device="cuda")
for variant in (14, 15, 50):
    os.environ["FFTW3_B200_FORCE_VARIANT"] = str(variant)
    for K in (64, 1024, 4096, 16384, 65536, 131072, 262144, 1048576):
        dims = [(n, K, K)]
        how = [(K, 1, 1), (total // (n * K), n * K, n * K)] if total // (n * K) > 1 else [(K, 1, 1)]
        p = lib.plan_guru_dft("d", dims, how, x.data_ptr(), x.data_ptr(), -1, B.FFTW_ESTIMATE)
        assert p
        lib.lib.fftw_b200_set_async(1)
        for _ in range(2):
            lib.execute("d", p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            lib.execute("d", p)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        print("variant %d stride %8d KiB: %.3f ms %.0f GB/s" % (variant, K * 16 // 1024, ms, 32.0 * total / ms / 1e6), flush=True)
        lib.lib.fftw_b200_set_async(0)
        lib.destroy_plan("d", p)
