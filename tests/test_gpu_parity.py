"""Parity of the CUDA path against the oracle, through the C-ABI, on a real
GPU.  Tolerance: rel. L2 <= 1.5 * eps * log2(N) (x4 for Bluestein sizes), see
tests/fftcheck.py.  Also runs the reference's OWN self-checking harness
(tests/bench.c + libbench2 verifier, prebuilt into oracle/_ref/bench_b200)
against the product library."""
import os
import subprocess

import numpy as np
import pytest

import fftcheck as F
from fftw3_b200 import binding as B
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRECS = ["d", "f"]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 12, 13, 16, 17, 30, 31, 64, 100, 128, 169, 243, 256, 512,
                               1000, 1009, 1024, 2048, 4096])
def test_c2c_1d(gpu_lib, prec, n):
    err, tol = F.c2c(gpu_lib, prec, (n,), howmany=7)
    assert err <= tol


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("flags", [B.FFTW_ESTIMATE, B.FFTW_MEASURE])
def test_config1_c2c_1024_batched(gpu_lib, prec, flags):
    """BASELINE config 1 at a reduced batch the oracle finishes in seconds."""
    err, tol = F.c2c(gpu_lib, prec, (1024,), howmany=256, flags=flags)
    assert err <= tol
    err, tol = F.c2c(gpu_lib, prec, (1024,), howmany=256, inplace=True, sign=1, flags=flags)
    assert err <= tol


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("shape,inplace,sign", [((16, 12), False, -1), ((8, 6, 10), True, -1), ((5, 7), True, 1),
                                                ((64, 64, 64), True, -1), ((32, 48, 20), False, 1),
                                                ((4, 4, 4, 3), False, 1), ((128, 128), True, -1)])
def test_c2c_nd(gpu_lib, prec, shape, inplace, sign):
    err, tol = F.c2c(gpu_lib, prec, shape, howmany=2, inplace=inplace, sign=sign)
    assert err <= tol


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n", [16384, 65536, 30030, 1 << 18])
def test_c2c_four_step(gpu_lib, prec, n):
    err, tol = F.c2c(gpu_lib, prec, (n,), howmany=3)
    assert err <= tol


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("shape", [(16,), (18,), (15,), (9,), (2,), (1,), (4, 6), (3, 5, 8), (6, 7), (1009,), (1024,),
                                   (4096,), (1 << 16,), (32, 32, 32)])
@pytest.mark.parametrize("inplace", [False, True])
def test_r2c_c2r(gpu_lib, prec, shape, inplace):
    err, tol = F.r2c(gpu_lib, prec, shape, howmany=3, inplace=inplace)
    assert err <= tol
    err, tol = F.c2r(gpu_lib, prec, shape, howmany=3, inplace=inplace)
    assert err <= tol


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("unfused", [False, True])
@pytest.mark.parametrize("shape,howmany", [((1 << 17,), 5), ((4 * 3 ** 9,), 2), ((3, 1 << 16), 2), ((1 << 21,), 1)])
def test_long_even_real_lines_split_and_merge_fused(gpu_lib, prec, unfused, shape, howmany, monkeypatch):
    """Even-size real transforms whose half-size transform is a four-step: c2r merge on the first pass's load
    (fft_fast.cuh flavour 11 / generic kernel), r2c split on the last pass's store (flavour 10), against the same
    oracle and bound; also with both kept as passes of their own."""
    if unfused:
        monkeypatch.setenv("FFTW3_B200_C2R_UNFUSED", "1")
        monkeypatch.setenv("FFTW3_B200_R2C_UNFUSED", "1")
    for inplace in (False, True):
        err, tol = F.c2r(gpu_lib, prec, shape, howmany=howmany, inplace=inplace)
        assert err <= tol, ("c2r", prec, shape, inplace, err, tol)
        err, tol = F.r2c(gpu_lib, prec, shape, howmany=howmany, inplace=inplace)
        assert err <= tol, ("r2c", prec, shape, inplace, err, tol)


def test_config2_r2c_c2r_2e20_float(gpu_lib):
    """BASELINE config 2 (N = 2^20 single precision) at batch 2."""
    err, tol = F.r2c(gpu_lib, "f", (1 << 20,), howmany=2)
    assert err <= tol
    err, tol = F.c2r(gpu_lib, "f", (1 << 20,), howmany=2)
    assert err <= tol


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("kind", list(B.R2R_KINDS))
@pytest.mark.parametrize("n", [2, 3, 8, 9, 16, 37, 256, 1000])
def test_r2r_1d(gpu_lib, prec, kind, n):
    err, tol = F.r2r(gpu_lib, prec, (n,), [kind], howmany=5)
    assert err <= tol


@pytest.mark.parametrize("prec", PRECS)
def test_config5b_redft10_2d(gpu_lib, prec):
    err, tol = F.r2r(gpu_lib, prec, (256, 256), ["REDFT10", "REDFT10"], howmany=1)
    assert err <= tol
    err, tol = F.r2r(gpu_lib, prec, (8, 6), ["REDFT10", "RODFT11"], howmany=2, inplace=True)
    assert err <= tol


def test_config5a_prime_1009(gpu_lib):
    err, tol = F.c2c(gpu_lib, "d", (1009,), howmany=64)
    assert err <= tol


def test_device_pointers_zero_copy(gpu_lib):
    """Device-resident arrays are transformed in place with no staging."""
    import torch
    rng = np.random.default_rng(7)
    x = F.rand_complex(rng, (48, 40, 64), "d")
    t = torch.from_numpy(x).cuda()
    p = gpu_lib.fn("d", "plan_dft_3d")(48, 40, 64, t.data_ptr(), t.data_ptr(), -1, B.FFTW_ESTIMATE)
    assert p
    gpu_lib.execute("d", p)
    torch.cuda.synchronize()
    gpu_lib.destroy_plan("d", p)
    assert O.rel_l2(t.cpu().numpy(), O.dft(x)) <= F.tol_for("d", x.shape)


def test_size_independent_properties_512cubed(gpu_lib):
    """BASELINE config 3 at full size (512^3 double, in place, device resident):
    the oracle cannot run this in seconds, so check what must hold at any size:
    impulse -> constant, linearity, and forward/backward round trip == N * x
    (the reference's own verifier uses the same properties,
    libbench2/verify-lib.c:260-356)."""
    import torch
    n = 512
    N = n ** 3
    dev = torch.device("cuda:0")
    a = torch.zeros((n, n, n), dtype=torch.complex128, device=dev)
    pf = gpu_lib.fn("d", "plan_dft_3d")(n, n, n, a.data_ptr(), a.data_ptr(), -1, B.FFTW_ESTIMATE)
    pb = gpu_lib.fn("d", "plan_dft_3d")(n, n, n, a.data_ptr(), a.data_ptr(), +1, B.FFTW_ESTIMATE)
    assert pf and pb
    a[3, 5, 7] = 1.0
    gpu_lib.execute("d", pf)
    torch.cuda.synchronize()
    assert float((a.abs() - 1).abs().max()) < 1e-13          # |DFT(impulse)| == 1 everywhere
    ph = a[1, 2, 3]
    w = np.exp(-2j * np.pi * (3 * 1 + 5 * 2 + 7 * 3) / n)
    assert abs(complex(ph) - w) < 1e-13
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.rand((n, n, n), dtype=torch.float64, device=dev, generator=g) - 0.5
    a.copy_(x.to(torch.complex128))
    a += 1j * (torch.rand((n, n, n), dtype=torch.float64, device=dev, generator=g) - 0.5)
    ref = a.clone()
    gpu_lib.execute("d", pf)
    gpu_lib.execute("d", pb)
    torch.cuda.synchronize()
    err = float((a / N - ref).abs().pow(2).sum().sqrt() / ref.abs().pow(2).sum().sqrt())
    assert err < 2 * 1.5 * 2.0 ** -52 * 27
    gpu_lib.destroy_plan("d", pf)
    gpu_lib.destroy_plan("d", pb)


def _bench(prec):
    return os.path.join(ROOT, "oracle", "_ref", "bench_b200" if prec == "d" else "benchf_b200")


@pytest.mark.parametrize("prec", PRECS)
def test_reference_verifier_accepts_us(gpu_lib, prec):
    """The reference's tests/bench self-checker (linearity, impulse, shift
    theorems; tolerance 1e-10 / 1e-3 hard-wired in libbench2/bench-main.c:70),
    compiled from the reference sources against OUR header and library."""
    exe = _bench(prec)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/bench_b200 not prebuilt")
    probs = ["oc1024", "ic1024*16", "ic64x64x64", "obc37", "ocf1009", "or1024", "ir1024", "orb1024", "irb18",
             "or15x10", "irb6x5x4", "oc16v4", "//oc128", "ok64e10", "ok64e01", "ok33e00", "ok32o00", "ok16e11x8o11",
             "ok64h", "ok64f", "ok64b", "ic13x5v3", "oc30030", "ofr65536", "ok256e10x256e10"]
    args = [exe, "-oestimate"]
    for pr in probs:
        args += ["--verify", pr]
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_reference_accuracy_metric(gpu_lib):
    """`bench --accuracy` compares us with the reference's 160-bit multiprecision
    FFT (libbench2/mp.c); the reference's own codelet-less build scores
    2.07e-16 (n=1024) and ~7.5e-16 (n=1009) forward L2."""
    exe = _bench("d")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/bench_b200 not prebuilt")
    for prob, bound in (("oc1024", 4e-16), ("oc1009", 1.5e-15), ("oc4096", 5e-16)):
        r = subprocess.run([exe, "-oestimate", "--accuracy", prob], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        vals = [float(v) for v in r.stdout.split()]
        assert vals[1] < bound and vals[4] < bound, (prob, vals)


def test_dist_single_rank_gpu(gpu_lib):
    """The slab plan degenerates to a local 3-D transform on one GPU (P = 1):
    checks the stage composition (Y, X, Z, gather) on the real kernels."""
    import torch
    import torch.distributed as dist
    from fftw3_b200 import dist as D
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29611", rank=0, world_size=1)
    rng = np.random.default_rng(11)
    shape = (32, 64, 128)
    x = F.rand_complex(rng, shape, "d")
    for transposed in (False, True):
        local = torch.from_numpy(x.reshape(-1).copy()).cuda()
        pl = D.SlabPlan3D(gpu_lib, *shape, local, flags=B.FFTW_ESTIMATE, transposed_out=transposed,
                          exchange="collective")
        pl.execute()
        torch.cuda.synchronize()
        pl.destroy()
        got = local.cpu().numpy()
        ref = O.dft(x)
        if transposed:
            got = got.reshape(shape[1], shape[0], shape[2]).transpose(1, 0, 2)
        else:
            got = got.reshape(shape)
        assert O.rel_l2(got, ref) <= F.tol_for("d", shape)


import goldencheck as G  # noqa: E402


@pytest.mark.parametrize("name,x,want", G.cases(), ids=[c[0] for c in G.cases()])
def test_reference_golden_vectors(gpu_lib, name, x, want):
    """committed outputs of the reference's own code (tests/golden/make_golden.py)"""
    got = G.via_lib(gpu_lib, name, x)
    assert O.rel_l2(got, want) < 3e-15


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n", [2053, 4100, 10007])
def test_large_bluestein(gpu_lib, prec, n):
    err, tol = F.c2c(gpu_lib, prec, (n,), howmany=3)
    assert err <= tol
    err, tol = F.r2c(gpu_lib, prec, (n,), howmany=2)
    assert err <= tol


def test_inplace_transpose_and_stride_change(gpu_lib):
    rng = np.random.default_rng(8)
    x = F.rand_complex(rng, (96, 96), "d")
    x0 = x.copy()
    p = gpu_lib.plan_guru_dft("d", [(96, 96, 1), (96, 1, 96)], [], x.ctypes.data, x.ctypes.data, -1, B.FFTW_ESTIMATE)
    assert p
    gpu_lib.execute("d", p)
    gpu_lib.destroy_plan("d", p)
    assert O.rel_l2(x.T, O.dft(x0)) <= F.tol_for("d", (96, 96))


_VARIANT_SHAPES = (((1024,), 37, False), ((512, 30), 3, True), ((13, 1024, 20), 1, True),
                   ((1 << 19,), 2, False), ((256, 64, 36), 1, False),
                   # powers of ten: radix-10 specialised kernels, four-step with non-binary twiddle split
                   ((1000,), 37, False), ((100, 30), 3, True), ((13, 1000, 20), 1, True),
                   ((1000000,), 1, False))


@pytest.mark.parametrize("variant", list(range(12, 24)) + list(range(36, 57)))
def test_every_specialised_kernel_variant(gpu_lib, variant, monkeypatch):
    """Pin each specialised-kernel variant of the planner (tile widths x flavours of
    fft_fast.cuh) and check parity on shapes that exercise
    ROW, COL, four-step (fused twiddle store, transposed store) and partial tiles."""
    monkeypatch.setenv("FFTW3_B200_FORCE_VARIANT", str(variant))
    for prec in PRECS:
        for shape, hm, inplace in _VARIANT_SHAPES:
            err, tol = F.c2c(gpu_lib, prec, shape, howmany=hm, inplace=inplace, sign=-1 if variant % 2 else 1)
            assert err <= tol, (variant, prec, shape)


def test_rank0_transposes(gpu_lib):
    """rank-0 guru transforms are copies/transposes (rdft/rank0.c, kernel/transpose.c): exercised
    here out of place (tiled transposing kernel) for real, complex-float and complex-double data."""
    import ctypes as C
    rng = np.random.default_rng(9)
    a = rng.standard_normal((300, 200))
    b = np.zeros((200, 300))
    h = (B.Iodim * 2)(B.Iodim(300, 200, 1), B.Iodim(200, 1, 300))
    p = gpu_lib.fn("d", "plan_guru_r2r")(0, None, 2, C.cast(h, C.c_void_p), a.ctypes.data, b.ctypes.data, None,
                                         B.FFTW_ESTIMATE)
    assert p
    gpu_lib.execute("d", p)
    gpu_lib.destroy_plan("d", p)
    assert np.array_equal(b, a.T)
    for prec in PRECS:
        x = F.rand_complex(rng, (5, 70, 90), prec)
        y = np.zeros((5, 90, 70), dtype=x.dtype)
        p = gpu_lib.plan_guru_dft(prec, [], [(5, 6300, 6300), (70, 90, 1), (90, 1, 70)], x.ctypes.data, y.ctypes.data,
                                  -1, B.FFTW_ESTIMATE)
        assert p
        gpu_lib.execute(prec, p)
        gpu_lib.destroy_plan(prec, p)
        assert np.array_equal(y, x.transpose(0, 2, 1))


def _random_problems(seed, count):
    """Random problem strings in the reference harness' mini-language
    (libbench2/problem.c:229-318): rank 1-3, sizes built from factors <= 13 plus the odd
    prime, interleaved ('v') and external ('*') batches, c2c / r2c / c2r / r2r with random kinds."""
    rng = np.random.default_rng(seed)
    kinds = ["f", "b", "h", "e00", "e01", "e10", "e11", "o00", "o01", "o10", "o11"]
    out = []
    for _ in range(count):
        rank = int(rng.integers(1, 4))
        budget = 20000 ** (1.0 / rank)
        dims = []
        for _ in range(rank):
            n = 1
            while True:
                f = int(rng.choice([2, 2, 2, 3, 3, 4, 5, 7, 8, 11, 13, 16]))
                if n * f > budget:
                    break
                n *= f
                if rng.random() < 0.25:
                    break
            if rng.random() < 0.08:
                n = int(rng.choice([17, 19, 23, 31, 37, 101, 127]))
            dims.append(max(n, 2))
        fam = str(rng.choice(["c", "c", "r", "k"]))
        s = ("i" if rng.random() < 0.5 else "o") + fam
        if fam != "k":
            s += "f" if rng.random() < 0.5 else "b"
            body = "x".join(str(d) for d in dims)
        else:
            body = "x".join("%d%s" % (max(d, 3), rng.choice(kinds)) for d in dims)
        s += body
        r = rng.random()
        if r < 0.3:
            s += "v%d" % int(rng.integers(2, 9))
        elif r < 0.6:
            s += "*%d" % int(rng.integers(2, 9))
        out.append(s)
    return out


@pytest.mark.parametrize("prec", PRECS)
def test_reference_verifier_random_problems(gpu_lib, prec):
    """The reference's self-checker on seeded random problems (what its tests/check.pl -r does),
    against the product library on the real GPU."""
    exe = _bench(prec)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/bench_b200 not prebuilt")
    probs = _random_problems(20261017 if prec == "d" else 7, 120)
    for i in range(0, len(probs), 30):
        args = [exe, "-oestimate"]
        for pr in probs[i:i + 30]:
            args += ["--verify", pr]
        r = subprocess.run(args, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, (probs[i:i + 30], r.stdout[-1500:], r.stderr[-1500:])


@pytest.mark.gpu
def test_wisdom_tool_produces_importable_wisdom(gpu_lib, tmp_path):
    """fftw-wisdom / fftwf-wisdom (tools/fftw_wisdom.c; reference tools/fftw-wisdom.c) plan the
    requested sizes in MEASURE mode; the file they write lets a FFTW_WISDOM_ONLY plan succeed."""
    import subprocess
    libdir = os.path.join(ROOT, "fftw3_b200", "lib")
    for prec, tool in (("d", "fftw-wisdom"), ("f", "fftwf-wisdom")):
        out = tmp_path / (tool + ".txt")
        r = subprocess.run([os.path.join(libdir, tool), "-v", "-m", "-n", "-o", str(out), "cof1024v64", "rif256",
                            "ki10e10x8e01", "cib32x32"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert r.stdout.count("Planning transform") == 4
        text = out.read_text()
        assert text.count("(b200_fft_pass") >= 4, text
        gpu_lib.fn(prec, "forget_wisdom")()
        x = np.zeros((64, 1024), dtype=np.complex128 if prec == "d" else np.complex64)
        y = np.zeros_like(x)
        args = ([1024], 64, x.ctypes.data, None, 1, 1024, y.ctypes.data, None, 1, 1024, B.FFTW_FORWARD)
        assert not gpu_lib.plan_many_dft(prec, *args, B.FFTW_MEASURE | B.FFTW_WISDOM_ONLY)
        assert gpu_lib.fn(prec, "import_wisdom_from_filename")(str(out).encode()) == 1
        p = gpu_lib.plan_many_dft(prec, *args, B.FFTW_MEASURE | B.FFTW_WISDOM_ONLY)
        assert p, "wisdom written by the tool was not usable"
        gpu_lib.destroy_plan(prec, p)
        gpu_lib.fn(prec, "forget_wisdom")()


@pytest.mark.gpu
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("shape,kinds,inplace", [
    ((4096, 64), ("REDFT10", "RODFT01"), False),     # strided 4096-lines: transposed scratch lines
    ((3000, 40), ("REDFT01", "R2HC"), True),
    ((2, 2500, 33), ("DHT", "RODFT11", "HC2R"), False),
    ((2049, 20), ("REDFT00", "REDFT11"), True),
    ((512, 512), ("REDFT10", "REDFT10"), True),      # both dims fused, COL tile kernel on dim 0
    ((256, 100, 64), ("RODFT10", "REDFT01", "DHT"), False),
    ((1024, 1024), ("RODFT11", "RODFT00"), False),
])
@pytest.mark.parametrize("transposes", [False, True])
def test_r2r_fused_maps_and_long_strided_lines(gpu_lib, prec, shape, kinds, inplace, transposes, monkeypatch):
    """The r2r PRE/POST maps ride in the load/store of one FFT pass per dimension
    (device/r2r_maps.cuh; specialised kernels flavour 9 and the generic kernel); strided lines
    too long for a tile: two line passes with transposed stores for a dense 2-d array, transposed
    scratch lines around the pass otherwise (forced by FFTW3_B200_R2R_TRANSPOSES).  Same oracle and
    bound as the other r2r tests."""
    if transposes:
        monkeypatch.setenv("FFTW3_B200_R2R_TRANSPOSES", "1")
    err, tol = F.r2r(gpu_lib, prec, shape, list(kinds), inplace=inplace)
    assert err <= tol, (prec, shape, kinds, err, tol)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("mode", ["rader", "bluestein"])
def test_prime_sizes_rader_and_bluestein(gpu_lib, prec, mode, monkeypatch):
    """Prime sizes through both algorithms (dft/rader.c:95-165, dft/bluestein.c:82-128): forced
    Rader (n - 1 smooth, one CTA), forced Bluestein, 1-d batches, both signs, and as a dimension
    of a 2-d in-place transform."""
    monkeypatch.setenv("FFTW3_B200_PRIME", mode)
    for n in (5, 7, 13, 17, 31, 101, 257, 1009, 2017, 4099):
        for sign in (-1, 1):
            err, tol = F.c2c(gpu_lib, prec, (n,), howmany=33, sign=sign)
            assert err <= 4 * tol, (mode, n, sign, err, tol)
    err, tol = F.c2c(gpu_lib, prec, (40, 1009), howmany=1, inplace=True)
    assert err <= 4 * tol
    err, tol = F.c2c(gpu_lib, prec, (1009, 24), howmany=2, inplace=False)
    assert err <= 4 * tol


def test_plain_fftw_program_on_the_gpu(gpu_lib, tmp_path):
    """The drop-in claim end to end: examples/fftw_tutorial.c (plain fftw3.h code: c2c, in-place
    2-d r2c/c2r, DCT-II/III, wisdom) compiled with gcc, linked against libfftw3_b200.so, run."""
    import subprocess
    libdir = os.path.join(ROOT, "fftw3_b200", "lib")
    exe = str(tmp_path / "tutorial")
    subprocess.run(["gcc", "-O1", "-Wall", os.path.join(ROOT, "examples", "fftw_tutorial.c"),
                    "-I" + os.path.join(ROOT, "include"), "-L" + libdir, "-lfftw3_b200", "-lm", "-Wl,-rpath," + libdir,
                    "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.count(" ok") == 7 and "FAILED" not in r.stdout, r.stdout + r.stderr
