#!/usr/bin/env python
"""Strided 1024-point passes of 1024^3 (dims 1 and 0): block-cooperative 16x16x4 kernels vs the 32x32 one-exchange
kernels (variants 55 = 64-byte segments, 56 = 128-byte segments); then the whole transform, ESTIMATE and MEASURE."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fftw3_b200 import binding as B
import fftcheck as F

lib = B.load()
for prec in ("d", "f"):
    for v in (55, 56):
        os.environ["FFTW3_B200_FORCE_VARIANT"] = str(v)
        print("parity variant", v, prec, F.c2c(lib, prec, (1024, 70), howmany=3, inplace=True), F.c2c(lib, prec, (1024, 33), howmany=1, sign=1),
              F.c2c(lib, prec, (5, 1024, 20), howmany=1, inplace=True), flush=True)
os.environ.pop("FFTW3_B200_FORCE_VARIANT", None)
lib.lib.fftw_b200_set_async(1)
n = 1024
a = torch.zeros((n, n, n), dtype=torch.complex128, device="cuda")
ptr = a.data_ptr()
gb = 2 * 16 * n ** 3 / 1e9
strides = [n * n, n, 1]

def timed(plan, steps=5):
    for _ in range(2): lib.execute("d", plan)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): lib.execute("d", plan)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

for d in (1, 0):
    for v in (14, 15, 38, 39, 50, 51, 55, 56):
        os.environ["FFTW3_B200_FORCE_VARIANT"] = str(v)
        dims = [(n, strides[d], strides[d])]
        hm = [(n, strides[e], strides[e]) for e in range(3) if e != d]
        p = lib.plan_guru_dft("d", dims, hm, ptr, ptr, -1, B.FFTW_ESTIMATE)
        ms = timed(p)
        print("dim%d variant %2d: %.3f ms  %.0f GB/s  %s" % (d, v, ms, gb / ms * 1e3, " ".join(lib.sprint_plan("d", p).split())[60:130]), flush=True)
        lib.destroy_plan("d", p)
os.environ.pop("FFTW3_B200_FORCE_VARIANT", None)
os.environ["FFTW3_B200_VERBOSE"] = "1"
for flags, nm in ((B.FFTW_ESTIMATE, "estimate"), (B.FFTW_MEASURE, "measure")):
    p = lib.fn("d", "plan_dft_3d")(n, n, n, ptr, ptr, -1, flags)
    a.zero_()
    print("3-D %s: %.3f ms  %s" % (nm, timed(p, 10), " ".join(lib.sprint_plan("d", p).split())[:330]), flush=True)
    lib.destroy_plan("d", p)
