#!/usr/bin/env python
"""C5b (REDFT10 4096^2 double) in pieces: the fused row pass per tile width, the transposes, the whole plan."""
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
n = 4096
x = torch.rand(n, n, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty_like(x)
import ctypes as C
for force in (None, 12, 13, 18, 19):
    if force is None: os.environ.pop("FFTW3_B200_FORCE_VARIANT", None)
    else: os.environ["FFTW3_B200_FORCE_VARIANT"] = str(force)
    p = lib.plan_many_r2r("d", [n], n, x.data_ptr(), None, 1, n, y.data_ptr(), None, 1, n, ["REDFT10"], B.FFTW_ESTIMATE)
    print("rows REDFT10 force=%s: %.1f us  %s" % (force, 1e3 * timed(lib, "d", p), " ".join(lib.sprint_plan("d", p).split())[:150]), flush=True)
    lib.destroy_plan("d", p)
os.environ.pop("FFTW3_B200_FORCE_VARIANT", None)
# the same row pass storing its lines transposed (output stride n, line distance 1): two of these make the 2-D
# transform without the two transposes
for force in (None, 12):
    if force is None: os.environ.pop("FFTW3_B200_FORCE_VARIANT", None)
    else: os.environ["FFTW3_B200_FORCE_VARIANT"] = str(force)
    p = lib.plan_many_r2r("d", [n], n, x.data_ptr(), None, 1, n, y.data_ptr(), None, n, 1, ["REDFT10"], B.FFTW_ESTIMATE)
    print("rows REDFT10 -> transposed stores force=%s: %.1f us  %s" % (force, 1e3 * timed(lib, "d", p), " ".join(lib.sprint_plan("d", p).split())[:150]), flush=True)
    lib.destroy_plan("d", p)
os.environ.pop("FFTW3_B200_FORCE_VARIANT", None)
h = (B.Iodim * 2)(B.Iodim(n, n, 1), B.Iodim(n, 1, n))
p = lib.fn("d", "plan_guru_r2r")(0, None, 2, C.cast(h, C.c_void_p), x.data_ptr(), y.data_ptr(), None, B.FFTW_ESTIMATE)
print("transpose 4096^2 f64: %.1f us" % (1e3 * timed(lib, "d", p)), flush=True)
lib.destroy_plan("d", p)
for flags, nm in ((B.FFTW_ESTIMATE, "estimate"), (B.FFTW_MEASURE, "measure")):
    p = lib.fn("d", "plan_r2r_2d")(n, n, x.data_ptr(), y.data_ptr(), 5, 5, flags)
    print("2-D REDFT10 %s: %.1f us  %s" % (nm, 1e3 * timed(lib, "d", p), " ".join(lib.sprint_plan("d", p).split())[:400]), flush=True)
    lib.destroy_plan("d", p)
