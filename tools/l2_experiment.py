#!/usr/bin/env python
"""L2-resident pass pairs on the bench workload: one process, one 16 GiB array, many plans.

usage: l2_experiment.py [n] [--quick]
Prints one line per configuration: ms per forward 3-D transform of n^3 c2c double in place.
  passes    single passes (guru plans over one dimension each) -- what the three-pass plan is made of
  3pass     the default plan (one HBM pass per dimension)
  l2 ...    FFTW3_B200_L2_BLOCK_MB x LANES x PAIR x KEEP grid
Every L2 configuration is checked: forward + backward / N must return the input (sampled).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fftw3_b200 import binding as B  # noqa: E402


def timed(lib, plan, steps=5, warm=2):
    for _ in range(warm):
        lib.execute("d", plan)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        lib.execute("d", plan)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1024
    quick = "--quick" in sys.argv
    flags = B.FFTW_MEASURE if "--measure" in sys.argv else B.FFTW_ESTIMATE
    lib = B.load()
    lib.lib.fftw_b200_set_async(1)
    dev = torch.device("cuda", 0)
    a = torch.empty((n, n, n), dtype=torch.complex128, device=dev)
    ar = torch.view_as_real(a)

    def fill():
        g = torch.Generator(device=dev).manual_seed(0)
        ar.copy_(torch.rand(ar.shape, dtype=torch.float64, device=dev, generator=g) - 0.5)

    fill()
    ptr = a.data_ptr()
    gb = 2 * 16 * n ** 3 / 1e9
    KNOBS = ("FFTW3_B200_L2_BLOCK_MB", "FFTW3_B200_L2_LANES", "FFTW3_B200_L2_PAIR", "FFTW3_B200_L2_KEEP",
             "FFTW3_B200_SPLIT", "FFTW3_B200_SPLIT_MB", "FFTW3_B200_SPLIT_LANES")
    for k in KNOBS:
        os.environ.pop(k, None)

    # ---- single passes
    strides = [n * n, n, 1]

    def single(d, label):
        dims = [(n, strides[d], strides[d])]
        hm = [(n, strides[e], strides[e]) for e in range(3) if e != d]
        p = lib.plan_guru_dft("d", dims, hm, ptr, ptr, -1, flags)
        assert p
        ms = timed(lib, p)
        print("pass dim%d %-28s %.3f ms  %.0f GB/s   %s" % (d, label, ms, gb / ms * 1e3, " ".join(lib.sprint_plan("d", p).split())[:160]), flush=True)
        lib.destroy_plan("d", p)
        ar.mul_(1e-3)

    os.environ["FFTW3_B200_SPLIT"] = "0"
    for d in range(3):
        single(d, "one kernel")
    os.environ["FFTW3_B200_SPLIT"] = "2"
    for mb in (4, 8, 16, 32, 64):
        for lanes in (1, 2, 3, 4, 6):
            os.environ["FFTW3_B200_SPLIT_MB"] = str(mb)
            os.environ["FFTW3_B200_SPLIT_LANES"] = str(lanes)
            for d in (0, 1):
                single(d, "split mb=%d lanes=%d" % (mb, lanes))
    for k in KNOBS:
        os.environ.pop(k, None)

    def full(label):
        fill()
        p = lib.fn("d", "plan_dft_3d")(n, n, n, ptr, ptr, -1, flags)
        assert p, label
        ms = timed(lib, p)
        # correctness: forward then backward of fresh data returns n^3 * input
        fill()
        ref = a[3, 5, :64].clone()
        lib.execute("d", p)
        pb = lib.fn("d", "plan_dft_3d")(n, n, n, ptr, ptr, +1, flags)
        lib.execute("d", pb)
        torch.cuda.synchronize()
        got = a[3, 5, :64] / float(n) ** 3
        err = float((got - ref).abs().max() / ref.abs().max())
        nst = lib.sprint_plan("d", p).count("fft-pass") if n <= 64 else -1
        print("%-44s %.3f ms   %.0f GFLOP/s   roundtrip err %.1e %s" % (label, ms, 5 * n ** 3 * 3 * (n.bit_length() - 1) / ms / 1e6, err,
                                                                        "OK" if err < 1e-12 else "WRONG"), flush=True)
        lib.destroy_plan("d", p)
        lib.destroy_plan("d", pb)
        return ms

    os.environ["FFTW3_B200_SPLIT"] = "0"
    full("3pass, one kernel per dim")
    for mode in (1, 2):
        for mb in (8, 16, 32):
            for lanes in (2, 3, 4):
                os.environ["FFTW3_B200_SPLIT"] = str(mode)
                os.environ["FFTW3_B200_SPLIT_MB"] = str(mb)
                os.environ["FFTW3_B200_SPLIT_LANES"] = str(lanes)
                full("split mode=%d mb=%d lanes=%d" % (mode, mb, lanes))
    for k in KNOBS:
        os.environ.pop(k, None)
    if "--no-l2" in sys.argv:
        return
    os.environ["FFTW3_B200_SPLIT"] = "0"
    grid = []
    mbs = (32, 64) if quick else (16, 32, 48, 64, 96)
    for pair in ("inner", "outer"):
        for mb in mbs:
            for lanes in ((3,) if quick else (1, 2, 3, 4)):
                for keep in ((2,) if quick else (2, 4)):
                    grid.append((pair, mb, lanes, keep))
    best = None
    for pair, mb, lanes, keep in grid:
        os.environ["FFTW3_B200_L2_BLOCK_MB"] = str(mb)
        os.environ["FFTW3_B200_L2_LANES"] = str(lanes)
        os.environ["FFTW3_B200_L2_PAIR"] = pair
        os.environ["FFTW3_B200_L2_KEEP"] = str(keep)
        ms = full("l2 pair=%s mb=%d lanes=%d keep=%d" % (pair, mb, lanes, keep))
        if best is None or ms < best[0]:
            best = (ms, pair, mb, lanes, keep)
    print("best:", best)


if __name__ == "__main__":
    main()
