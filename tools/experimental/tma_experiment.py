#!/usr/bin/env python
"""Time single strided passes of an n^3 double array with pinned kernel variants and check
them against the default kernel's result.
usage: tma_experiment.py [n] [variants...]   (default n = 1024, variants = 57..62 + default)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fftw3_b200 import binding as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
variants = [int(v) for v in sys.argv[2:]] or [-1, 57, 58, 59, 60, 61, 62]
lib = B.load()
torch.manual_seed(0)
x0 = torch.rand(n ** 3, 2, dtype=torch.float64, device="cuda") - 0.5
ref = torch.empty_like(x0)
x = torch.empty_like(x0)


def geometry(dim):
    if dim == 1:
        return [(n, n, n)], [(n, 1, 1), (n, n * n, n * n)]
    return [(n, n * n, n * n)], [(n * n, 1, 1)]


for dim in (1, 0):
    dims, how = geometry(dim)
    os.environ.pop("FFTW3_B200_FORCE_VARIANT", None)
    p = lib.plan_guru_dft("d", dims, how, ref.data_ptr(), ref.data_ptr(), -1, B.FFTW_ESTIMATE)
    ref.copy_(x0)
    lib.execute("d", p)
    torch.cuda.synchronize()
    lib.destroy_plan("d", p)
    for sign in (-1, 1):
        for v in variants:
            if sign == 1 and v < 0:
                continue
            if v >= 0:
                os.environ["FFTW3_B200_FORCE_VARIANT"] = str(v)
            else:
                os.environ.pop("FFTW3_B200_FORCE_VARIANT", None)
            p = lib.plan_guru_dft("d", dims, how, x.data_ptr(), x.data_ptr(), sign, B.FFTW_ESTIMATE)
            if not p:
                print("dim", dim, "variant", v, "no plan")
                continue
            desc = " ".join(lib.sprint_plan("d", p).split())
            if sign == -1:
                x.copy_(x0)
            else:
                x.copy_(ref)        # backward of the forward result gives n * x0
            lib.execute("d", p)
            torch.cuda.synchronize()
            if sign == -1:
                err = ((x - ref).norm() / ref.norm()).item()
            else:
                err = ((x / n - x0).norm() / x0.norm()).item()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                lib.execute("d", p)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            gb = 2 * 16 * n ** 3 / 1e9
            print("dim %d sign %+d variant %3d: err %.2e  min %.3f ms  med %.3f ms  %.0f GB/s  | %s"
                  % (dim, sign, v, err, ts[0], ts[2], gb / ts[0] * 1e3, desc[-70:]), flush=True)
            lib.destroy_plan("d", p)
