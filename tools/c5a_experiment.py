#!/usr/bin/env python
"""C5a (prime n = 1009, 16384 double transforms): the one-CTA Bluestein kernel per tile width, and Rader."""
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
n, hm = 1009, 16384
x = torch.rand(hm, n, 2, dtype=torch.float64, device="cuda") - 0.5
y = torch.empty_like(x)
for mode in ("bluestein", "rader"):
    os.environ["FFTW3_B200_PRIME"] = mode
    for force in (None, 12, 13, 14):
        if force is None: os.environ.pop("FFTW3_B200_FORCE_VARIANT", None)
        else: os.environ["FFTW3_B200_FORCE_VARIANT"] = str(force)
        p = lib.plan_many_dft("d", [n], hm, x.data_ptr(), None, 1, n, y.data_ptr(), None, 1, n, -1, B.FFTW_ESTIMATE)
        if not p:
            print("%s force=%s: no plan" % (mode, force)); continue
        print("%s force=%s: %.1f us  %s" % (mode, force, 1e3 * timed(lib, "d", p), " ".join(lib.sprint_plan("d", p).split())[:170]), flush=True)
        lib.destroy_plan("d", p)
