#!/usr/bin/env python
"""One 2-d REDFT10 of a 4096 x 4096 double array (BASELINE config C5b), for ncu captures of the fused r2r pass."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
from fftw3_b200 import binding as B
lib = B.load()
n = 4096
x = torch.zeros(n * n, dtype=torch.float64, device="cuda")
kinds = (C.c_int * 2)(B.R2R_KINDS["REDFT10"], B.R2R_KINDS["REDFT10"])
p = lib.fn("d", "plan_r2r")(2, (C.c_int * 2)(n, n), x.data_ptr(), x.data_ptr(), kinds, B.FFTW_ESTIMATE)
print(" ".join(lib.sprint_plan("d", p).split()))
for _ in range(2):
    lib.execute("d", p)
torch.cuda.synchronize()
