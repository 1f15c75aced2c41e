#!/usr/bin/env python
"""Run one strided pass of a [nz][1024][1024] double array with a pinned kernel variant (for ncu captures).
usage: one_pass.py <variant> [dim] [nz]     dim 1: transform along y (stride 16 KiB), nz planes
                                           dim 0: transform along z (nz must be 1024), dim 2: rows"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fftw3_b200 import binding as B
variant = int(sys.argv[1]); dim = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nz = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
if variant >= 0:
    os.environ["FFTW3_B200_FORCE_VARIANT"] = str(variant)
lib = B.load()
n = 1024
x = torch.zeros(nz * n * n, 2, dtype=torch.float64, device="cuda")
if dim == 1:
    dims, how = [(n, n, n)], [(n, 1, 1), (nz, n * n, n * n)]
elif dim == 0:
    dims, how = [(nz, n * n, n * n)], [(n * n, 1, 1)]
else:
    dims, how = [(n, 1, 1)], [(nz * n, n, n)]
p = lib.plan_guru_dft("d", dims, how, x.data_ptr(), x.data_ptr(), -1, B.FFTW_ESTIMATE)
print(" ".join(lib.sprint_plan("d", p).split()))
for _ in range(3):
    lib.execute("d", p)
torch.cuda.synchronize()
