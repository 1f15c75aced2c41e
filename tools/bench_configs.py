#!/usr/bin/env python
"""Measure every BASELINE.json config on one GPU (device-resident arrays, CUDA
events, FFTW_MEASURE) and print one JSON line per config:
GFLOP/s in the reference's convention (libbench2/mflops.c:19-27) and the
fraction of the HBM roofline for ONE read + ONE write of the arrays
(BASELINE.md section 3)."""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from fftw3_b200 import binding as B  # noqa: E402


def timeit(lib, prec, plan, reps, scale_fn=None):
    lib.lib.fftw_b200_set_async(1)
    for _ in range(3):
        lib.execute(prec, plan)
        if scale_fn:
            scale_fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.launch_count()
        e0.record()
        for _ in range(reps):
            lib.execute(prec, plan)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
        if scale_fn:
            scale_fn()
    launches = (lib.launch_count() - l0) // reps
    lib.lib.fftw_b200_set_async(0)
    return best, launches


def e2e_host(lib, prec, make_plan, nbytes_in, nbytes_out, reps=3):
    """The same transform through the public call on PINNED HOST arrays (fftw_malloc): upload, passes and download
    inside the timed region, wall clock.  Batched problems go through the chunk pipeline of csrc/host/exec.c (upload of
    chunk c + 1 | passes of chunk c | download of chunk c - 1); FFTW3_B200_PIPELINE=0 gives the serial path."""
    import ctypes as C
    mal, fre = lib.fn(prec, "malloc"), lib.fn(prec, "free")
    mal.restype, mal.argtypes, fre.argtypes = C.c_void_p, [C.c_size_t], [C.c_void_p]
    hin, hout = mal(nbytes_in), mal(nbytes_out)
    C.memset(hin, 0, nbytes_in)
    out = {}
    for label, env in (("pipelined", None), ("serial", "0")):
        if env is None:
            os.environ.pop("FFTW3_B200_PIPELINE", None)
        else:
            os.environ["FFTW3_B200_PIPELINE"] = env
        p = make_plan(hin, hout)
        lib.execute(prec, p)
        t0 = time.perf_counter()
        for _ in range(reps):
            lib.execute(prec, p)
        out[label] = (time.perf_counter() - t0) / reps * 1e3
        out[label + "_plan"] = " ".join(lib.sprint_plan(prec, p).split())[:90]
        lib.destroy_plan(prec, p)
    os.environ.pop("FFTW3_B200_PIPELINE", None)
    fre(hin); fre(hout)
    return {"e2e_host_ms": out["pipelined"], "e2e_host_serial_ms": out["serial"], "e2e_h2d_bytes": nbytes_in,
            "e2e_d2h_bytes": nbytes_out, "e2e_plan": out["pipelined_plan"]}


def run_configs(lib, flags, peak, only="", echo=False, e2e=True):
    """Times the BASELINE.json configs other than the bench workload; returns one dict per config.
    `roofline_frac_1pass` = (one read + one write of the arrays) / time / peak: the fraction of the
    speed of light a plan that touched HBM exactly once would reach."""
    dev = "cuda"
    out = []

    def report(name, flops, bytes_1pass, ms, launches, plan, prec, extra=None):
        line = {"config": name, "ms": ms, "gflops": flops / ms / 1e6, "ideal_gbs_1pass": bytes_1pass / ms / 1e6,
                "roofline_frac_1pass": bytes_1pass / ms / 1e6 / peak, "launches": launches,
                "plan": " ".join(lib.sprint_plan(prec, plan).split())[:600]}
        if extra:
            line.update(extra)
        if echo:
            print(json.dumps(line), flush=True)
        out.append(line)

    def want(k):
        return not only or k in only.split(",")

    if want("C1"):
        n, hm = 1024, 16384
        x = torch.rand(hm, n, 2, dtype=torch.float64, device=dev) - 0.5
        y = torch.empty_like(x)
        p = lib.plan_many_dft("d", [n], hm, x.data_ptr(), None, 1, n, y.data_ptr(), None, 1, n, -1, flags)
        ms, l = timeit(lib, "d", p, 20)
        ex = e2e_host(lib, "d", lambda a, b: lib.plan_many_dft("d", [n], hm, a, None, 1, n, b, None, 1, n, -1, B.FFTW_ESTIMATE),
                      16 * n * hm, 16 * n * hm) if e2e else None
        report("C1 c2c f64 N=1024 x16384 out-of-place", 5 * n * hm * math.log2(n), 2 * 16 * n * hm, ms, l, p, "d", ex)
        lib.destroy_plan("d", p)
        p = lib.plan_many_dft("d", [n], hm, x.data_ptr(), None, 1, n, x.data_ptr(), None, 1, n, -1, flags)
        ms, l = timeit(lib, "d", p, 20, lambda: x.mul_(1 / 32.0))
        report("C1 c2c f64 N=1024 x16384 in-place", 5 * n * hm * math.log2(n), 2 * 16 * n * hm, ms, l, p, "d")
        lib.destroy_plan("d", p)
        del x, y
    if want("C2"):
        n, hm = 1 << 20, 256
        x = torch.rand(hm, n, dtype=torch.float32, device=dev) - 0.5
        y = torch.empty(hm, n // 2 + 1, 2, dtype=torch.float32, device=dev)
        p = lib.plan_many_dft_r2c("f", [n], hm, x.data_ptr(), None, 1, n, y.data_ptr(), None, 1, n // 2 + 1, flags)
        assert p
        ms, l = timeit(lib, "f", p, 5)
        nb = 4 * n * hm + 8 * (n // 2 + 1) * hm
        ex = e2e_host(lib, "f", lambda a, b: lib.plan_many_dft_r2c("f", [n], hm, a, None, 1, n, b, None, 1, n // 2 + 1, B.FFTW_ESTIMATE),
                      4 * n * hm, 8 * (n // 2 + 1) * hm) if e2e else None
        report("C2 r2c f32 N=2^20 x256", 2.5 * n * hm * math.log2(n), nb, ms, l, p, "f", ex)
        lib.destroy_plan("f", p)
        p = lib.plan_many_dft_c2r("f", [n], hm, y.data_ptr(), None, 1, n // 2 + 1, x.data_ptr(), None, 1, n, flags)
        assert p
        ms, l = timeit(lib, "f", p, 5, lambda: y.mul_(1e-3))
        report("C2 c2r f32 N=2^20 x256", 2.5 * n * hm * math.log2(n), nb, ms, l, p, "f")
        lib.destroy_plan("f", p)
        del x, y
    if want("C3"):
        n = 512
        x = torch.rand(n, n, n, 2, dtype=torch.float64, device=dev) - 0.5
        p = lib.fn("d", "plan_dft_3d")(n, n, n, x.data_ptr(), x.data_ptr(), -1, flags)
        ms, l = timeit(lib, "d", p, 10, lambda: x.mul_(1.0 / n ** 1.5))
        report("C3 c2c f64 512^3 in-place (3 passes; frac is for 1 pass-equivalent)", 5 * n ** 3 * math.log2(n ** 3),
               2 * 16 * n ** 3, ms, l, p, "d")
        lib.destroy_plan("d", p)
        del x
    if want("C5a"):
        n, hm = 1009, 16384
        x = torch.rand(hm, n, 2, dtype=torch.float64, device=dev) - 0.5
        y = torch.empty_like(x)
        p = lib.plan_many_dft("d", [n], hm, x.data_ptr(), None, 1, n, y.data_ptr(), None, 1, n, -1, flags)
        ms, l = timeit(lib, "d", p, 10)
        report("C5a c2c f64 prime N=1009 x16384", 5 * n * hm * math.log2(n), 2 * 16 * n * hm, ms, l, p, "d")
        lib.destroy_plan("d", p)
        del x, y
    if want("C5b"):
        n = 4096
        x = torch.rand(n, n, dtype=torch.float64, device=dev) - 0.5
        y = torch.empty_like(x)
        p = lib.fn("d", "plan_r2r_2d")(n, n, x.data_ptr(), y.data_ptr(), 5, 5, flags)
        assert p
        ms, l = timeit(lib, "d", p, 10)
        report("C5b REDFT10 f64 4096^2 (2 dims; frac is for 1 pass-equivalent)", 2.5 * n * n * math.log2(n * n),
               2 * 8 * n * n, ms, l, p, "d")
        lib.destroy_plan("d", p)
        del x, y
    if want("C1f"):
        n, hm = 1024, 16384
        x = torch.rand(hm, n, 2, dtype=torch.float32, device=dev) - 0.5
        y = torch.empty_like(x)
        p = lib.plan_many_dft("f", [n], hm, x.data_ptr(), None, 1, n, y.data_ptr(), None, 1, n, -1, flags)
        ms, l = timeit(lib, "f", p, 20)
        report("(extra) c2c f32 N=1024 x16384", 5 * n * hm * math.log2(n), 2 * 8 * n * hm, ms, l, p, "f")
        lib.destroy_plan("f", p)
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--flags", default="measure")
    a = ap.parse_args()
    lib = B.load()
    peak = 6454.3
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    flags = B.FFTW_MEASURE if a.flags == "measure" else B.FFTW_ESTIMATE
    run_configs(lib, flags, peak, a.only, echo=True)


if __name__ == "__main__":
    main()
