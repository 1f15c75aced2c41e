#!/usr/bin/env python
"""C2 (r2c / c2r of 2^20 x 256 float) with the split / merge fused into the four-step passes and as passes of their own."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fftw3_b200 import binding as B

def timed(lib, prec, plan, steps=20):
    for _ in range(3): lib.execute(prec, plan)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): lib.execute(prec, plan)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

lib = B.load(); lib.lib.fftw_b200_set_async(1)
n, hm = 1 << 20, 256
x = torch.rand(hm, n, dtype=torch.float32, device="cuda") - 0.5
y = torch.zeros(hm, n // 2 + 1, 2, dtype=torch.float32, device="cuda")
for env in ({}, {"FFTW3_B200_C2R_UNFUSED": "1", "FFTW3_B200_R2C_UNFUSED": "1"}):
    for k in ("FFTW3_B200_C2R_UNFUSED", "FFTW3_B200_R2C_UNFUSED"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for flags, nm in ((B.FFTW_ESTIMATE, "estimate"), (B.FFTW_MEASURE, "measure")):
        p = lib.plan_many_dft_r2c("f", [n], hm, x.data_ptr(), None, 1, n, y.data_ptr(), None, 1, n // 2 + 1, flags)
        print("r2c %s %s: %.3f ms  %s" % (nm, "unfused" if env else "fused", timed(lib, "f", p), " ".join(lib.sprint_plan("f", p).split())[:330]), flush=True)
        lib.destroy_plan("f", p)
        p = lib.plan_many_dft_c2r("f", [n], hm, y.data_ptr(), None, 1, n // 2 + 1, x.data_ptr(), None, 1, n, flags)
        print("c2r %s %s: %.3f ms  %s" % (nm, "unfused" if env else "fused", timed(lib, "f", p), " ".join(lib.sprint_plan("f", p).split())[:330]), flush=True)
        lib.destroy_plan("f", p)
