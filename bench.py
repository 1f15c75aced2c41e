#!/usr/bin/env python
"""bench.py -- headline benchmark of fftw3_b200 (contract: see the task brief).

Workload (BASELINE.json `metric`): 3-D complex double c2c 1024^3, in place,
one "step" = one forward transform of the whole 16 GiB array.
  N = 1 : the whole array on one B200 (fftw_plan_dft_3d through the C-ABI).
  N > 1 : slab decomposition over N GPUs, one process per GPU, exchange over
          NVLink (fftw_mpi_plan_dft_3d equivalent, fftw3_b200/dist.py).
Metric: GFLOP/s in FFTW's convention 5*N*log2(N)/t (libbench2/mflops.c:19-23).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 1024]
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = 1 << 30


def flops_c2c(shape):
    n = 1
    for s in shape:
        n *= s
    return 5.0 * n * math.log2(n)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------ CPU reference
def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


REF_BUILD_NOTE = ("reference sources (planner, Cooley-Tukey solvers, OpenMP threads, API) compiled here with scalar n1/t1 "
                  "codelets emitted by this repo's generator in the reference's codelet ABI -- genfft needs OCaml, which "
                  "the image lacks; no SIMD codelets")


def cpu_reference_run(steps, warmup, sample_n=256, measure=False, timelimit=20.0):
    """The reference's own CPU implementation (oracle/_ref = unmodified FFTW planner / solvers / threads
    compiled here, leaves = scalar codelets from oracle/refbuild/gen_codelets.py) with its OpenMP threads:
    one in-place c2c double transform of sample_n^3 per step."""
    import ctypes as C
    import numpy as np
    path = os.path.join(ROOT, "oracle", "_ref", "libfftw3_ref.so")
    kind = "reference"
    if not os.path.exists(path):
        return None
    cores = host_cores()
    # torchrun exports OMP_NUM_THREADS=1 to its workers and the reference's `#pragma omp parallel for`
    # (threads/openmp.c:77) obeys it: set the team size explicitly, before and after libgomp is loaded
    os.environ["OMP_NUM_THREADS"] = str(cores)
    lib = C.CDLL(path)
    try:
        gomp = C.CDLL("libgomp.so.1")
        gomp.omp_set_dynamic(0)
        gomp.omp_set_num_threads(cores)
        gomp.omp_get_max_threads.restype = C.c_int
        cores = int(gomp.omp_get_max_threads())
    except OSError:
        pass
    lib.fftw_init_threads()
    lib.fftw_plan_with_nthreads(C.c_int(cores))
    lib.fftw_plan_dft_3d.restype = C.c_void_p
    lib.fftw_plan_dft_3d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
    lib.fftw_execute.argtypes = [C.c_void_p]
    lib.fftw_destroy_plan.argtypes = [C.c_void_p]
    lib.fftw_set_timelimit.argtypes = [C.c_double]
    n = sample_n
    # a random 64^3 block tiled over the array (filling 2^30 points with an RNG would take longer than the FFTs)
    rng = np.random.default_rng(0)
    b = min(n, 64)
    blk = (rng.uniform(-0.5, 0.5, (b, b, b)) + 1j * rng.uniform(-0.5, 0.5, (b, b, b))).astype(np.complex128)
    a = np.tile(blk, (n // b, n // b, n // b)) if n % b == 0 else np.resize(blk, (n, n, n))
    lib.fftw_set_timelimit(C.c_double(timelimit if measure else -1.0))
    t0 = time.perf_counter()
    p = lib.fftw_plan_dft_3d(n, n, n, a.ctypes.data, a.ctypes.data, -1, 0 if measure else (1 << 6))
    assert p
    plan_s = time.perf_counter() - t0
    for _ in range(warmup):
        lib.fftw_execute(p)
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.fftw_execute(p)
    dt = (time.perf_counter() - t0) / steps
    lib.fftw_destroy_plan(p)
    return {"value": flops_c2c((n, n, n)) / dt / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": kind,
            "sample": "%d^3 c2c double in place, %s, %d OpenMP threads; %s" % (
                n, ("FFTW_MEASURE (time limit %.0f s, planning took %.1f s)" % (timelimit, plan_s)) if measure else "FFTW_ESTIMATE",
                cores, REF_BUILD_NOTE),
            "ms_per_step": dt * 1e3, "sample_n": n}


def cpu_vectorised_run(sample_n=256, repeats=3):
    """Context only (not the reference arm): a vectorised CPU FFT on the same bounded sample --
    torch.fft on the host (MKL, all cores) -- because the reference builds here without its
    generated SIMD codelets and is far slower than a production FFTW."""
    try:
        import torch
        n = sample_n
        a = torch.rand(n, n, n, dtype=torch.complex128)
        torch.fft.fftn(a)
        best = 1e30
        for _ in range(repeats):
            t0 = time.perf_counter()
            torch.fft.fftn(a)
            best = min(best, time.perf_counter() - t0)
        return {"value": flops_c2c((n, n, n)) / best / 1e9, "unit": "GFLOP/s", "cores": torch.get_num_threads(),
                "kind": "torch.fft on CPU (MKL), not the reference", "sample": "%d^3 c2c double out of place" % n}
    except Exception as e:        # context only: never fail the bench for it
        return {"unavailable": repr(e)[:120]}


def reference_sample_size(args):
    """The whole workload (1024^3) when a quick probe says the requested steps fit in about four minutes of CPU
    time, else 512^3 (BASELINE config C3), else 256^3."""
    if args.ref_sample:
        return args.ref_sample
    probe = cpu_reference_run(1, 1, sample_n=min(256, args.size))
    if probe is None:
        return min(256, args.size)
    t256 = probe["ms_per_step"] * 1e-3
    budget = 240.0
    for n in (args.size, 512, 256):
        if n > args.size:
            continue
        scale = (n / 256.0) ** 3 * (math.log2(n) / 8.0) * 1.5        # + memory-bound slow-down at sizes beyond the caches
        if (max(1, args.steps) + 1) * t256 * scale + 30.0 <= budget:
            return n
    return min(256, args.size)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sn = reference_sample_size(args)
    r = cpu_reference_run(max(1, args.steps), max(0, min(args.warmup, 1)), sample_n=sn, measure=True, timelimit=(10.0 if sn >= 1024 else 20.0))
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libfftw3_ref.so is not built"}))
        return
    n = args.size
    line = {
        "impl": "reference", "metric": "GFLOP/s (5N log2 N), 3-D c2c double", "value": r["value"], "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": ("%d^3 c2c double in place, forward, on the host CPU" % n) if sn == n else
                               ("%d^3 c2c double in place; this arm times the bounded sample %d^3 of it (same transform, "
                                "1/%d of the points) on the host CPU" % (n, sn, (n // sn) ** 3)),
                   "sample_size": sn, "same_config": sn == n},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------ self check
def impulse_expected(n, j, ks):
    """Closed form of the forward DFT of a unit impulse at j = (j0, j1, j2), sampled at ks (m x 3):
    exp(-2 pi i (j . k) / n), with the exponent reduced exactly in integers (libbench2/verify-lib.c:293-323
    checks impulses the same way, against a constant)."""
    import numpy as np
    e = (ks.astype(np.int64) * np.asarray(j, dtype=np.int64)[None, :]).sum(axis=1) % n
    ang = -2.0 * np.pi * e.astype(np.float64) / n
    return np.cos(ang) + 1j * np.sin(ang)


def self_check_single(lib, B, plan, a, n, flags):
    """Outside the timed region, on the very array and plan that were timed: (i) a unit impulse at a
    non-trivial index against the closed-form phases at 512 sampled outputs, (ii) forward + backward of
    random data against the input (relative L2, in plane blocks)."""
    import numpy as np
    import torch
    ar = torch.view_as_real(a)
    j = (n // 3 + 1, n // 5 + 2, n // 7 + 3)
    a.zero_()
    a[j] = 1.0
    lib.execute("d", plan)
    torch.cuda.synchronize()
    rng = np.random.default_rng(1234)
    ks = rng.integers(0, n, size=(512, 3))
    kt = torch.from_numpy(ks).to(a.device)
    got = a[kt[:, 0], kt[:, 1], kt[:, 2]].cpu().numpy()
    imp = float(np.abs(got - impulse_expected(n, j, ks)).max())
    # round trip
    g = torch.Generator(device=a.device).manual_seed(7)
    ar.copy_(torch.rand(ar.shape, dtype=torch.float64, device=a.device, generator=g) - 0.5)
    keep = a.clone()
    pb = lib.fn("d", "plan_dft_3d")(n, n, n, a.data_ptr(), a.data_ptr(), B.FFTW_BACKWARD, flags)
    assert pb, "backward plan returned NULL"
    lib.execute("d", plan)
    lib.execute("d", pb)
    torch.cuda.synchronize()
    lib.destroy_plan("d", pb)
    num = den = 0.0
    step = max(1, n // 16)
    inv = 1.0 / float(n) ** 3
    for i in range(0, n, step):
        d = a[i:i + step] * inv - keep[i:i + step]
        num += float((d.real ** 2 + d.imag ** 2).sum())
        den += float((keep[i:i + step].real ** 2 + keep[i:i + step].imag ** 2).sum())
    del keep
    rt = (num / den) ** 0.5
    lg = 3 * math.log2(n)
    ok = bool(imp <= 1e-13 * lg and rt <= 3.0 * 2.0 ** -52 * lg)
    return {"impulse_max_err": imp, "impulse_at": list(j), "impulse_samples": 512, "roundtrip_rel_l2": rt,
            "tolerance": {"impulse": 1e-13 * lg, "roundtrip": 3.0 * 2.0 ** -52 * lg}, "ok": ok}


# --------------------------------------------------------------------- ours
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from fftw3_b200 import binding as B

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = B.load()
    n = args.size
    shape = (n, n, n)
    dev = torch.device("cuda", local_rank)

    if world > 1:
        from fftw3_b200 import dist as fdist
        return fdist.bench_slab_3d(args, lib, n, world, rank, local_rank, ClockSampler, flops_c2c, cpu_reference_run,
                                   impulse_expected)

    # ---- N = 1: whole array on one GPU, in place, device resident ----
    a = torch.empty(shape, dtype=torch.complex128, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    ar = torch.view_as_real(a)
    ar.copy_(torch.rand(ar.shape, dtype=torch.float64, device=dev, generator=g) - 0.5)
    flags = B.FFTW_MEASURE if not args.estimate else B.FFTW_ESTIMATE
    if args.wisdom and os.path.exists(args.wisdom):
        lib.fn("d", "import_wisdom_from_filename")(args.wisdom.encode())
    t0 = time.perf_counter()
    plan = lib.fn("d", "plan_dft_3d")(n, n, n, a.data_ptr(), a.data_ptr(), B.FFTW_FORWARD, flags)
    assert plan, "fftw_plan_dft_3d returned NULL"
    plan_s = time.perf_counter() - t0
    plan_txt = lib.sprint_plan("d", plan)
    if args.wisdom:
        lib.fn("d", "export_wisdom_to_filename")(args.wisdom.encode())
    # FFTW_MEASURE overwrites the arrays while planning (doc/reference.texi:405-416): fresh data for the timed runs
    ar.copy_(torch.rand(ar.shape, dtype=torch.float64, device=dev, generator=g) - 0.5)
    lib.lib.fftw_b200_set_async(1)
    for _ in range(max(3, args.warmup)):
        lib.execute("d", plan)
        ar.mul_(1.0 / n ** 1.5)        # keep magnitudes bounded across repeated in-place transforms
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    l0 = lib.launch_count()
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(args.steps):
        lib.execute("d", plan)          # enqueued on the legacy default stream == torch's current stream
        ev[i + 1].record()
    torch.cuda.synchronize()
    launches = lib.launch_count() - l0
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[-1])
    ms = total_ms / args.steps
    gflops = flops_c2c(shape) / (ms * 1e-3) / 1e9
    lib.lib.fftw_b200_set_async(0)

    # ---- roofline.  Algorithmic bytes of one HBM pass = one read + one write of the array (DESIGN.md
    # section 3: 32 B per point per pass in f64).  The dominant kernel is the pass along dim 0 (stride
    # 16 MiB); it and the two other passes are timed alone, live, through single-dimension guru plans of
    # the same array (same kernels and variants: their pass signatures hit the wisdom just measured).
    array_bytes = 16 * n ** 3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    strides = [n * n, n, 1]
    per_pass = []
    for d in range(3):
        dims = [(n, strides[d], strides[d])]
        hm = [(n, strides[e], strides[e]) for e in range(3) if e != d]
        sp = lib.plan_guru_dft("d", dims, hm, a.data_ptr(), a.data_ptr(), B.FFTW_FORWARD, flags)
        assert sp
        for _ in range(2):
            lib.execute("d", sp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            lib.execute("d", sp)
        e1.record()
        torch.cuda.synchronize()
        pms = e0.elapsed_time(e1) / 5
        txt = " ".join(lib.sprint_plan("d", sp).split())
        lib.destroy_plan("d", sp)
        ar.mul_(1e-6)
        per_pass.append({"dim": d, "ms": pms, "achieved_gbs": 2 * array_bytes / (pms * 1e-3) / 1e9,
                         "frac": 2 * array_bytes / (pms * 1e-3) / 1e9 / peak, "kernel": txt[:200]})
    dom = max(per_pass, key=lambda q: q["ms"])
    l2_paired = "L2-resident pass pairs" in plan_txt
    hbm_passes = 2 if l2_paired else 3          # a pair of passes that meets in L2 costs one read + one write
    step_gbs = hbm_passes * 2 * array_bytes / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                "traffic": None, "kernel": "pass along dim %d (the slowest of the three), timed alone: %.3f ms" % (dom["dim"], dom["ms"]),
                "algorithmic_bytes_per_launch": 2 * array_bytes,
                "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)",
                "per_pass_alone": per_pass,
                "step": {"hbm_pass_equivalents": hbm_passes, "algorithmic_bytes": hbm_passes * 2 * array_bytes,
                         "achieved": step_gbs, "frac": step_gbs / peak,
                         "as_three_passes_frac": 3 * 2 * array_bytes / (ms * 1e-3) / 1e9 / peak,
                         "note": ("dims 2 and 1 run as L2-resident pairs per plane group: 2 HBM pass-equivalents"
                                  if l2_paired else "one HBM pass per dimension")}}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass

    # ---- correctness of what was just timed (outside the timed region)
    check = None if args.no_check else self_check_single(lib, B, plan, a, n, flags)

    # ---- the other BASELINE.json configs (C1, C2, C3, C5a, C5b), device resident, same planner mode
    extra = None
    if not args.no_extra:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_configs
            extra = bench_configs.run_configs(lib, flags, peak)
        except Exception as e:       # informational: never fail the headline for it
            extra = [{"error": repr(e)[:200]}]

    # ---- e2e: the same transform through the C-ABI on HOST buffers ----
    e2e = None
    if not args.no_e2e:
        del a, ar
        torch.cuda.empty_cache()
        e2e_steps = max(1, args.steps)
        nbytes = array_bytes
        hp = lib.fn("d", "malloc")(nbytes)          # pinned when possible
        assert hp
        import ctypes as C
        host = np.ctypeslib.as_array((C.c_double * (2 * n ** 3)).from_address(hp))
        host[:] = 0.25
        hplan = lib.fn("d", "plan_dft_3d")(n, n, n, hp, hp, B.FFTW_FORWARD, flags)   # wisdom hit: no re-measuring
        assert hplan
        lib.execute("d", hplan)                      # warm-up: allocates the staging buffer
        host[:] = 0.25
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            lib.execute("d", hplan)                  # synchronous: H2D + 3 passes + D2H
        dt = (time.perf_counter() - t0) / e2e_steps
        lib.destroy_plan("d", hplan)
        lib.fn("d", "free")(hp)
        e2e = {"value": flops_c2c(shape) / dt / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "api": "fftw_plan_dft_3d + fftw_execute on fftw_malloc'd host memory"}
    lib.destroy_plan("d", plan)

    cpu = None if args.no_cpu else cpu_reference_run(1, 0, sample_n=min(256, n))
    line = {
        "metric": "GFLOP/s (5N log2 N), 3-D c2c double", "value": gflops, "unit": "GFLOP/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%d^3 c2c double in place, forward, device resident" % n,
                   "l2": "array (%.0f GiB) is larger than L2, no flush needed" % (array_bytes / GIB),
                   "planner": "FFTW_ESTIMATE" if args.estimate else "FFTW_MEASURE", "plan_seconds": plan_s,
                   "plan": " ".join(plan_txt.split())},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "check": check,
        "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
        "cpu_vectorised_context": None if args.no_cpu else cpu_vectorised_run(min(256, n)),
        "extra_configs": extra,
    }
    print(json.dumps(line), flush=True)
    if check is not None and not check["ok"]:
        sys.exit("bench.py: self check FAILED: %r" % (check,))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--estimate", action="store_true", help="plan with FFTW_ESTIMATE instead of FFTW_MEASURE")
    ap.add_argument("--wisdom", default=None, help="wisdom file to import before planning and export after")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the impulse / round-trip check of the timed plan")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs (extra_configs)")
    ap.add_argument("--ref-sample", type=int, default=0, help="--impl reference: cube edge of the CPU sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
