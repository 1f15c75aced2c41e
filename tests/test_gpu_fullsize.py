"""Parity at the sizes BASELINE.json actually names, new-array execution on the device, and the
"every size plans" contract -- all through the C-ABI on the B200 (`-m gpu`).

Full-size arrays are checked on sampled lines / sampled outputs against the oracle (the oracle is a
long-double CPU code: whole 10^8-point arrays would take minutes), plus size-independent properties.
Reference for the method: libbench2/verify-lib.c:260-414 (linearity / impulse / shift at full size).
"""
import ctypes as C
import threading

import numpy as np
import pytest

import fftcheck as F
from fftw3_b200 import binding as B
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    assert torch.cuda.is_available()
    return torch


# ------------------------------------------------------------------ every size plans
def _sizes(seed, count, hi):
    rng = np.random.default_rng(seed)
    out = set()
    while len(out) < count:
        kind = rng.integers(0, 4)
        if kind == 0:          # anything
            n = int(rng.integers(1, hi))
        elif kind == 1:        # odd smooth
            n = 1
            while True:
                f = int(rng.choice([3, 3, 5, 7, 9, 11, 13]))
                if n * f > hi:
                    break
                n *= f
                if n > 2000 and rng.random() < 0.3:
                    break
        elif kind == 2:        # even smooth
            n = 2
            while True:
                f = int(rng.choice([2, 2, 3, 4, 5, 7, 8]))
                if n * f > hi:
                    break
                n *= f
                if n > 2000 and rng.random() < 0.3:
                    break
        else:                  # around powers of two
            n = int((1 << int(rng.integers(3, 22))) + rng.integers(-3, 4))
        if 1 <= n <= hi:
            out.add(n)
    return sorted(out)


def test_no_null_plans_for_200_sizes(gpu_lib):
    """doc/reference.texi:357-360: the basic interface always succeeds.  Seeded list of 200 sizes <= 2^21
    (random, odd smooth, even smooth, near powers of two): r2c, c2r and c2c plans must all exist, in both
    precisions."""
    torch = _torch()
    hi = 1 << 21
    buf = torch.zeros(2 * (hi + 2), dtype=torch.float64, device="cuda")
    out = torch.zeros(2 * (hi + 2), dtype=torch.float64, device="cuda")
    null = []
    for n in _sizes(2024, 200, hi):
        for prec in ("d", "f"):
            for name, args in (("plan_dft_r2c_1d", (n, buf.data_ptr(), out.data_ptr(), B.FFTW_ESTIMATE)),
                               ("plan_dft_c2r_1d", (n, buf.data_ptr(), out.data_ptr(), B.FFTW_ESTIMATE)),
                               ("plan_dft_1d", (n, buf.data_ptr(), out.data_ptr(), -1, B.FFTW_ESTIMATE))):
                p = gpu_lib.fn(prec, name)(*args)
                if not p:
                    null.append((name, prec, n))
                else:
                    gpu_lib.destroy_plan(prec, p)
    assert not null, null[:20]


@pytest.mark.parametrize("prec", ["d", "f"])
@pytest.mark.parametrize("n", [10395, 15625, 19683, 50625, 177147])
def test_large_odd_real_transforms(gpu_lib, prec, n):
    """Odd smooth n beyond one CTA: the four-step carries the real-data ops (rdft/hc2hc.c:116-190,
    rdft/rdft2-rdft.c:42-130 in the reference).  1-d, in and out of place, and as the last dim of 2-d."""
    for inplace in (False, True):
        e, t = F.r2c(gpu_lib, prec, (n,), howmany=2, inplace=inplace)
        assert e <= t, ("r2c", n, inplace, e, t)
        e, t = F.c2r(gpu_lib, prec, (n,), howmany=2, inplace=inplace)
        assert e <= t, ("c2r", n, inplace, e, t)
    if n <= 50625:
        e, t = F.r2c(gpu_lib, prec, (3, n))
        assert e <= t, ("r2c 2d", n, e, t)
        e, t = F.c2r(gpu_lib, prec, (3, n))
        assert e <= t, ("c2r 2d", n, e, t)


def test_nested_four_step_2e25(gpu_lib):
    """Smooth n beyond the square of the one-pass limit (ADVICE: 2^25 double returned NULL): nested four-step."""
    torch = _torch()
    n = 1 << 25
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(n, 2, dtype=torch.float64, device="cuda", generator=g) - 0.5
    y = torch.empty_like(x)
    p = gpu_lib.fn("d", "plan_dft_1d")(n, x.data_ptr(), y.data_ptr(), -1, B.FFTW_ESTIMATE)
    assert p
    gpu_lib.execute("d", p)
    gpu_lib.destroy_plan("d", p)
    # sampled outputs against the definition, evaluated with exact phase reduction
    xc = torch.view_as_complex(x).cpu().numpy()
    ks = np.random.default_rng(0).integers(0, n, 6)
    j = np.arange(n, dtype=np.int64)
    got = torch.view_as_complex(y)[torch.from_numpy(ks).cuda()].cpu().numpy()
    for k, gk in zip(ks, got):
        ang = -2.0 * np.pi * ((j * int(k)) % n).astype(np.float64) / n
        want = np.sum(xc * (np.cos(ang) + 1j * np.sin(ang)))
        assert abs(gk - want) <= 1e-11 * np.sqrt(n), (k, gk, want)
    # and the inverse returns the input
    q = gpu_lib.fn("d", "plan_dft_1d")(n, y.data_ptr(), y.data_ptr(), +1, B.FFTW_ESTIMATE)
    assert q
    gpu_lib.execute("d", q)
    gpu_lib.destroy_plan("d", q)
    err = float(((y / n - x) ** 2).sum().sqrt() / (x ** 2).sum().sqrt())
    assert err <= 1.5 * 2.0 ** -52 * 25 * 2, err


def test_rank0_r2c_c2r(gpu_lib):
    """rdft/rank0-rdft2.c: a rank-0 r2c is out = in + 0i, c2r is out = Re(in)."""
    x = np.arange(1.0, 7.0)
    y = np.full(6, 9 + 9j)
    h = (B.Iodim * 1)(B.Iodim(6, 1, 1))
    p = gpu_lib.fn("d", "plan_guru_dft_r2c")(0, None, 1, C.cast(h, C.c_void_p), x.ctypes.data, y.ctypes.data, B.FFTW_ESTIMATE)
    assert p
    gpu_lib.execute("d", p)
    gpu_lib.destroy_plan("d", p)
    assert np.array_equal(y, x + 0j)
    z = np.zeros(6)
    y[:] = x + 1j * x[::-1]
    p = gpu_lib.fn("d", "plan_guru_dft_c2r")(0, None, 1, C.cast(h, C.c_void_p), y.ctypes.data, z.ctypes.data, B.FFTW_ESTIMATE)
    assert p
    gpu_lib.execute("d", p)
    gpu_lib.destroy_plan("d", p)
    assert np.array_equal(z, x)


# ------------------------------------------------------------------ new-array execution on the device
def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("prec", ["d", "f"])
def test_new_array_execute_on_device_pointers(gpu_lib, prec):
    """api/execute-dft.c:25-32, execute-split-dft.c, execute-dft-r2c.c, execute-dft-c2r.c, execute-r2r.c on
    DEVICE arrays other than the ones planned on, incl. a multi-pass (scratch-using) plan."""
    torch = _torch()
    rng = np.random.default_rng(21)
    cd, rd = F.CDT[prec], F.RDT[prec]
    for n, hm in ((1024, 37), (1 << 17, 3), (360, 50)):
        a = _dev(torch, F.rand_complex(rng, (hm, n), prec)); b = torch.empty_like(a)
        p = gpu_lib.plan_many_dft(prec, [n], hm, a.data_ptr(), None, 1, n, b.data_ptr(), None, 1, n, -1, B.FFTW_ESTIMATE)
        assert p
        x = F.rand_complex(rng, (hm, n), prec)
        xd = _dev(torch, x); yd = torch.zeros_like(xd)
        gpu_lib.fn(prec, "execute_dft")(p, xd.data_ptr(), yd.data_ptr())
        gpu_lib.destroy_plan(prec, p)
        assert O.rel_l2(yd.cpu().numpy(), O.dft(x, rank=1)) <= F.tol_for(prec, (n,))
        assert np.array_equal(xd.cpu().numpy(), x)
    # split arrays
    n, hm = 1024, 9
    ri, ii = rng.uniform(-.5, .5, (hm, n)).astype(rd), rng.uniform(-.5, .5, (hm, n)).astype(rd)
    t = [_dev(torch, v) for v in (ri, ii, np.zeros_like(ri), np.zeros_like(ri))]
    p = gpu_lib.plan_guru_split_dft(prec, [(n, 1, 1)], [(hm, n, n)], *[v.data_ptr() for v in t], B.FFTW_ESTIMATE)
    assert p
    u = [_dev(torch, v) for v in (ii, ri, np.zeros_like(ri), np.zeros_like(ri))]      # swapped roles, new arrays
    gpu_lib.fn(prec, "execute_split_dft")(p, *[v.data_ptr() for v in u])
    gpu_lib.destroy_plan(prec, p)
    got = u[2].cpu().numpy() + 1j * u[3].cpu().numpy()
    assert O.rel_l2(got, O.dft((ii + 1j * ri).astype(cd), rank=1)) <= F.tol_for(prec, (n,))
    # r2c / c2r, even (three-step) and odd (fused) sizes
    for n in (4096, 1 << 18, 999):
        hm, h = 5, n // 2 + 1
        x = F.rand_real(rng, (hm, n), prec)
        a = _dev(torch, np.zeros_like(x)); b = _dev(torch, np.zeros((hm, h), cd))
        p = gpu_lib.plan_many_dft_r2c(prec, [n], hm, a.data_ptr(), None, 1, n, b.data_ptr(), None, 1, h, B.FFTW_ESTIMATE)
        assert p
        xd = _dev(torch, x); yd = _dev(torch, np.zeros((hm, h), cd))
        gpu_lib.fn(prec, "execute_dft_r2c")(p, xd.data_ptr(), yd.data_ptr())
        gpu_lib.destroy_plan(prec, p)
        X = yd.cpu().numpy()
        assert O.rel_l2(X, O.r2c(x, rank=1)) <= F.tol_for(prec, (n,))
        p = gpu_lib.plan_many_dft_c2r(prec, [n], hm, b.data_ptr(), None, 1, h, a.data_ptr(), None, 1, n, B.FFTW_ESTIMATE)
        assert p
        zd = _dev(torch, np.zeros_like(x))
        Xd = _dev(torch, X)
        gpu_lib.fn(prec, "execute_dft_c2r")(p, Xd.data_ptr(), zd.data_ptr())
        gpu_lib.destroy_plan(prec, p)
        assert O.rel_l2(zd.cpu().numpy(), O.c2r(X, n, rank=1)) <= 2 * F.tol_for(prec, (n,))
    # r2r
    n, hm = 1000, 7
    x = F.rand_real(rng, (hm, n), prec)
    a = _dev(torch, np.zeros_like(x)); b = _dev(torch, np.zeros_like(x))
    p = gpu_lib.plan_many_r2r(prec, [n], hm, a.data_ptr(), None, 1, n, b.data_ptr(), None, 1, n, ["REDFT10"], B.FFTW_ESTIMATE)
    assert p
    xd = _dev(torch, x); yd = _dev(torch, np.zeros_like(x))
    gpu_lib.fn(prec, "execute_r2r")(p, xd.data_ptr(), yd.data_ptr())
    gpu_lib.destroy_plan(prec, p)
    assert O.rel_l2(yd.cpu().numpy(), O.r2r(x, ["REDFT10"], rank=1)) <= F.tol(prec, 2 * n, 2.0)


@pytest.mark.parametrize("prec", ["d", "f"])
def test_new_array_execute_misaligned(gpu_lib, prec):
    """doc/reference.texi:1578-1607: new arrays must have the alignment of the planned ones UNLESS the plan
    was made with FFTW_UNALIGNED.  Plan on aligned arrays with FFTW_UNALIGNED, execute on arrays that start
    one real past a vector boundary -- on the device and on the host (staged path)."""
    torch = _torch()
    rng = np.random.default_rng(22)
    rd = F.RDT[prec]
    for shape in ((1024,), (64, 96), (1 << 16,)):
        n = int(np.prod(shape))
        hm = 3
        a = torch.zeros(hm * n * 2, dtype=torch.float64 if prec == "d" else torch.float32, device="cuda")
        b = torch.zeros_like(a)
        p = gpu_lib.plan_many_dft(prec, list(shape), hm, a.data_ptr(), None, 1, n, b.data_ptr(), None, 1, n, -1,
                                  B.FFTW_ESTIMATE | B.FFTW_UNALIGNED)
        assert p
        x = F.rand_complex(rng, (hm,) + shape, prec)
        flat = np.zeros(hm * n * 2 + 1, rd)
        flat[1:] = x.view(rd).reshape(-1)
        xd = torch.from_numpy(flat).cuda()
        yd = torch.zeros_like(xd)
        isz = flat.itemsize
        assert (xd.data_ptr() + isz) % (2 * isz) != 0
        gpu_lib.fn(prec, "execute_dft")(p, xd.data_ptr() + isz, yd.data_ptr() + isz)
        got = yd.cpu().numpy()[1:].view(F.CDT[prec]).reshape(x.shape)
        assert O.rel_l2(got, O.dft(x, rank=len(shape))) <= F.tol_for(prec, shape), shape
        # host arrays, misaligned the same way
        yh = np.zeros_like(flat)
        gpu_lib.fn(prec, "execute_dft")(p, flat.ctypes.data + isz, yh.ctypes.data + isz)
        got = yh[1:].view(F.CDT[prec]).reshape(x.shape)
        assert O.rel_l2(got, O.dft(x, rank=len(shape))) <= F.tol_for(prec, shape), shape
        gpu_lib.destroy_plan(prec, p)


def test_same_plan_from_four_threads_on_device_arrays(gpu_lib):
    """doc/threads.texi:225-270: fftw_execute* is thread-safe, even on one plan.  A plan that keeps its
    intermediate data in plan-owned scratch (r2c n = 2048: half-size FFT + split) is executed from four
    threads on four different device arrays at once; every result must be right."""
    torch = _torch()
    rng = np.random.default_rng(23)
    n, hm, h = 2048, 256, 1025
    a = torch.zeros(hm, n, dtype=torch.float64, device="cuda")
    b = torch.zeros(hm, h, dtype=torch.complex128, device="cuda")
    p = gpu_lib.plan_many_dft_r2c("d", [n], hm, a.data_ptr(), None, 1, n, b.data_ptr(), None, 1, h, B.FFTW_ESTIMATE)
    assert p
    xs = [F.rand_real(rng, (hm, n), "d") for _ in range(4)]
    xd = [_dev(torch, x) for x in xs]
    yd = [torch.zeros(hm, h, dtype=torch.complex128, device="cuda") for _ in range(4)]
    torch.cuda.synchronize()
    fn = gpu_lib.fn("d", "execute_dft_r2c")

    def work(i):
        for _ in range(20):
            fn(p, xd[i].data_ptr(), yd[i].data_ptr())

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    torch.cuda.synchronize()
    gpu_lib.destroy_plan("d", p)
    for i in range(4):
        assert O.rel_l2(yd[i].cpu().numpy(), O.r2c(xs[i], rank=1)) <= F.tol_for("d", (n,)), i


# ------------------------------------------------------------------ BASELINE configs at their real size
def test_config1_full_size_16384x1024(gpu_lib):
    """C1: 1024-point c2c double, batch 16384 (256 MiB in, 256 MiB out), FFTW_MEASURE.  64 sampled lines
    against the oracle + the first and the last line (tile tails); the rest through Parseval per line."""
    torch = _torch()
    n, hm = 1024, 16384
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(hm, n, 2, dtype=torch.float64, device="cuda", generator=g) - 0.5
    y = torch.empty_like(x)
    p = gpu_lib.plan_many_dft("d", [n], hm, x.data_ptr(), None, 1, n, y.data_ptr(), None, 1, n, -1, B.FFTW_MEASURE)
    assert p
    x.copy_(torch.rand(hm, n, 2, dtype=torch.float64, device="cuda", generator=g) - 0.5)    # MEASURE may clobber
    gpu_lib.execute("d", p)
    gpu_lib.destroy_plan("d", p)
    rows = np.unique(np.concatenate([[0, hm - 1], np.random.default_rng(1).integers(0, hm, 64)]))
    xs = torch.view_as_complex(x)[torch.from_numpy(rows).cuda()].cpu().numpy()
    ys = torch.view_as_complex(y)[torch.from_numpy(rows).cuda()].cpu().numpy()
    assert O.rel_l2(ys, O.dft(xs, rank=1)) <= F.tol_for("d", (n,))
    ex = (x ** 2).sum(dim=(1, 2)) * n
    ey = (y ** 2).sum(dim=(1, 2))
    assert float(((ey - ex).abs() / ex).max()) <= 1e-13


def test_config2_full_size_r2c_c2r_256x2e20_float(gpu_lib):
    """C2: r2c / c2r single precision N = 2^20, batch 256 (1 GiB each side).  Four sampled lines against
    the oracle; every line through the c2r(r2c(x)) = N x round trip."""
    torch = _torch()
    n, hm = 1 << 20, 256
    h = n // 2 + 1
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.rand(hm, n, dtype=torch.float32, device="cuda", generator=g) - 0.5
    keep = x.clone()
    y = torch.empty(hm, h, 2, dtype=torch.float32, device="cuda")
    p = gpu_lib.plan_many_dft_r2c("f", [n], hm, x.data_ptr(), None, 1, n, y.data_ptr(), None, 1, h, B.FFTW_ESTIMATE)
    assert p
    gpu_lib.execute("f", p)
    gpu_lib.destroy_plan("f", p)
    assert torch.equal(x, keep), "r2c modified its input"
    rows = [0, 97, 200, hm - 1]
    xs = x[rows].cpu().numpy()
    ys = torch.view_as_complex(y)[rows].cpu().numpy()
    assert O.rel_l2(ys, O.r2c(xs, rank=1)) <= F.tol_for("f", (n,))
    q = gpu_lib.plan_many_dft_c2r("f", [n], hm, y.data_ptr(), None, 1, h, x.data_ptr(), None, 1, n, B.FFTW_ESTIMATE)
    assert q
    gpu_lib.execute("f", q)
    gpu_lib.destroy_plan("f", q)
    err = ((x / n - keep) ** 2).sum(dim=1).sqrt() / (keep ** 2).sum(dim=1).sqrt()
    assert float(err.max()) <= 2 * F.tol_for("f", (n,))


def test_config5b_full_size_redft10_4096sq(gpu_lib):
    """C5b: 2-D REDFT10 4096^2 double.  Sampled outputs against the definition
    Y[k1,k2] = 4 sum x[j1,j2] cos(pi (j1+1/2) k1 / n) cos(pi (j2+1/2) k2 / n) (doc/reference.texi:2090-2098,
    evaluated in long double), and REDFT01(REDFT10(x)) = (2n)^2 x over the whole array."""
    torch = _torch()
    n = 4096
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    y = torch.empty_like(x)
    p = gpu_lib.fn("d", "plan_r2r_2d")(n, n, x.data_ptr(), y.data_ptr(), 5, 5, B.FFTW_ESTIMATE)
    assert p
    gpu_lib.execute("d", p)
    gpu_lib.destroy_plan("d", p)
    xh = x.cpu().numpy().astype(np.longdouble)
    yh = y.cpu().numpy()
    pi_ld = np.longdouble("3.14159265358979323846264338327950288")
    jodd = 2 * np.arange(n, dtype=np.int64) + 1
    rng = np.random.default_rng(2)
    scale = float(np.sqrt((xh ** 2).sum())) * 4
    for k1, k2 in [(0, 0), (n - 1, n - 1), (1, n - 2)] + [tuple(rng.integers(0, n, 2)) for _ in range(9)]:
        # cos(pi (2j+1) k / (2n)) with the argument reduced exactly in integers (mod 4n) before the long-double cosine
        c1 = np.cos(pi_ld * ((jodd * int(k1)) % (4 * n)).astype(np.longdouble) / (2 * n))
        c2 = np.cos(pi_ld * ((jodd * int(k2)) % (4 * n)).astype(np.longdouble) / (2 * n))
        want = 4 * float(c1 @ xh @ c2)
        assert abs(yh[k1, k2] - want) <= 1e-14 * scale, (k1, k2, yh[k1, k2], want)
    q = gpu_lib.fn("d", "plan_r2r_2d")(n, n, y.data_ptr(), y.data_ptr(), 4, 4, B.FFTW_ESTIMATE)     # REDFT01, in place
    assert q
    gpu_lib.execute("d", q)
    gpu_lib.destroy_plan("d", q)
    err = float(((y / (2.0 * n) ** 2 - x) ** 2).sum().sqrt() / (x ** 2).sum().sqrt())
    assert err <= F.tol("d", (2 * n) ** 2, 4.0), err


def test_config3_full_size_512cubed_sampled_outputs(gpu_lib):
    """C3: 512^3 c2c double in place, FFTW_MEASURE.  Eight sampled outputs against the 3-D definition
    (separable contraction in complex128: error ~ 1e-13 of the signal norm, which still pins every index
    and stride at full size), and forward + backward = N x."""
    torch = _torch()
    n = 512
    g = torch.Generator(device="cuda").manual_seed(9)
    a = torch.empty(n, n, n, dtype=torch.complex128, device="cuda")
    ar = torch.view_as_real(a)
    p = gpu_lib.fn("d", "plan_dft_3d")(n, n, n, a.data_ptr(), a.data_ptr(), -1, B.FFTW_MEASURE)
    assert p
    ar.copy_(torch.rand(ar.shape, dtype=torch.float64, device="cuda", generator=g) - 0.5)
    keep = a.clone()
    gpu_lib.execute("d", p)
    j = torch.arange(n, device="cuda", dtype=torch.float64)
    rng = np.random.default_rng(4)
    norm = float((torch.view_as_real(keep) ** 2).sum().sqrt())
    for k in [(0, 0, 0), (n - 1, n - 1, n - 1)] + [tuple(int(v) for v in rng.integers(0, n, 3)) for _ in range(6)]:
        e = [torch.polar(torch.ones_like(j), -2.0 * np.pi * ((j * kk) % n) / n) for kk in k]
        want = torch.einsum("abc,a,b,c->", keep, e[0], e[1], e[2])
        got = a[k]
        assert abs(complex(got) - complex(want)) <= 1e-11 * norm, (k, complex(got), complex(want))
    q = gpu_lib.fn("d", "plan_dft_3d")(n, n, n, a.data_ptr(), a.data_ptr(), +1, B.FFTW_ESTIMATE)
    assert q
    gpu_lib.execute("d", q)
    gpu_lib.destroy_plan("d", q)
    gpu_lib.destroy_plan("d", p)
    err = float((torch.view_as_real(a / float(n) ** 3 - keep) ** 2).sum().sqrt()) / norm
    assert err <= 2 * F.tol_for("d", (n, n, n)), err


# ------------------------------------------------------------------ several GPUs
def test_multi_gpu_parity_under_torchrun():
    """tests/dist_gpu_check.py (c2c push / gather / collective x natural / transposed, r2c / c2r incl.
    uneven non-smooth blocks, r2r) on min(device_count, 8) ranks; skipped on a one-GPU box."""
    import os
    import subprocess
    import sys
    torch = _torch()
    ng = min(torch.cuda.device_count(), 8)
    if ng < 2:
        pytest.skip("needs >= 2 GPUs (the driver's 1-GPU tier covers the P = 1 collective path elsewhere)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.pop("FFTW3_B200_DIST_PUSH", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ng),
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(root, "tests", "dist_gpu_check.py")],
                       capture_output=True, text=True, timeout=1500, env=env)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0 and "FAIL" not in r.stdout and r.stdout.count(" OK") >= 20, tail
