#!/usr/bin/env python
"""Throughput of the any-size (generic) kernel on shapes without a specialised kernel."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fftw3_b200 import binding as B
lib = B.load()


def bench(name, make, reps=5, force=None):
    if force is not None:
        os.environ["FFTW3_B200_FORCE_VARIANT"] = str(force)
    else:
        os.environ.pop("FFTW3_B200_FORCE_VARIANT", None)
    p, nbytes, flops, keep = make()
    assert p, name
    lib.lib.fftw_b200_set_async(1)
    for _ in range(2):
        lib.execute("d", p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.execute("d", p)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-44s %8.3f ms  %7.0f GB/s (1 r+w of the array per pass-equivalent)  %7.0f GFLOP/s   %s"
          % (name, ms, nbytes / ms / 1e6, flops / ms / 1e6, " ".join(lib.sprint_plan("d", p).split())[:150]), flush=True)
    lib.lib.fftw_b200_set_async(0)
    lib.destroy_plan("d", p)


def many(n, hm, flags=B.FFTW_ESTIMATE):
    def mk():
        x = torch.zeros(hm, n, 2, dtype=torch.float64, device="cuda")
        y = torch.zeros_like(x)
        p = lib.plan_many_dft("d", [n], hm, x.data_ptr(), None, 1, n, y.data_ptr(), None, 1, n, -1, flags)
        return p, 32.0 * n * hm, 5.0 * n * hm * math.log2(n), (x, y)
    return mk


def cube(n, flags=B.FFTW_MEASURE):
    def mk():
        x = torch.zeros(n, n, n, 2, dtype=torch.float64, device="cuda")
        p = lib.fn("d", "plan_dft_3d")(n, n, n, x.data_ptr(), x.data_ptr(), -1, flags)
        return p, 32.0 * n ** 3, 5.0 * n ** 3 * math.log2(n ** 3), (x,)
    return mk


bench("generic kernel, 1024 x 16384 (variant 0)", many(1024, 16384), force=0)
bench("generic kernel, 1000 x 16384", many(1000, 16384, B.FFTW_MEASURE))
bench("generic kernel, 1080 x 16384", many(1080, 16384, B.FFTW_MEASURE))
bench("generic kernel, 243 x 65536", many(243, 65536, B.FFTW_MEASURE))
bench("3-D 600^3 (2^3 3 5^2), measure", cube(600))
bench("3-D 384^3 (2^7 3), measure", cube(384))
