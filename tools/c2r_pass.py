#!/usr/bin/env python
"""One c2r of 2^20 x 256 float (BASELINE config C2, inverse direction), for ncu launch lists / captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fftw3_b200 import binding as B
lib = B.load()
n, hm = 1 << 20, 256
x = torch.zeros(hm, n, dtype=torch.float32, device="cuda")
y = torch.zeros(hm, n // 2 + 1, 2, dtype=torch.float32, device="cuda")
p = lib.plan_many_dft_c2r("f", [n], hm, y.data_ptr(), None, 1, n // 2 + 1, x.data_ptr(), None, 1, n, B.FFTW_ESTIMATE)
print(" ".join(lib.sprint_plan("f", p).split()))
for _ in range(2):
    lib.execute("f", p)
torch.cuda.synchronize()
