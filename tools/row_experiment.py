#!/usr/bin/env python
"""Contiguous 1024-point passes (C1 and the dim-2 pass of 1024^3): block-cooperative kernel variants vs the
warp-per-transform kernel (variant 54), double and single precision."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fftw3_b200 import binding as B
import fftcheck as F

def timed(lib, prec, plan, steps=20):
    for _ in range(3): lib.execute(prec, plan)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): lib.execute(prec, plan)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

lib = B.load()
for prec in ("d", "f"):
    os.environ["FFTW3_B200_FORCE_VARIANT"] = "54"
    print("parity warp kernel", prec, F.c2c(lib, prec, (1024,), howmany=37), F.c2c(lib, prec, (1024,), howmany=5, inplace=True, sign=1),
          F.c2c(lib, prec, (3, 1024), howmany=2), flush=True)
lib.lib.fftw_b200_set_async(1)
for prec, dt in (("d", torch.float64), ("f", torch.float32)):
    for hm in (16384, 1 << 20):
        x = torch.rand(hm, 1024, 2, dtype=dt, device="cuda") - 0.5
        y = torch.empty_like(x)
        for inplace in (0, 1):
            for v in (12, 13, 14, 18, 19, 54):
                os.environ["FFTW3_B200_FORCE_VARIANT"] = str(v)
                out = x if inplace else y
                p = lib.plan_many_dft(prec, [1024], hm, x.data_ptr(), None, 1, 1024, out.data_ptr(), None, 1, 1024, -1, B.FFTW_ESTIMATE)
                ms = timed(lib, prec, p, 50 if hm == 16384 else 5)
                gb = 2 * x.numel() * x.element_size() / 1e9
                print("%s 1024 x %7d %s variant %2d: %8.4f ms  %6.0f GB/s  %s" % (prec, hm, "in-place " if inplace else "out-place", v, ms, gb / ms * 1e3,
                      " ".join(lib.sprint_plan(prec, p).split())[55:110]), flush=True)
                lib.destroy_plan(prec, p)
                if inplace: x.mul_(1e-3)
        del x, y
