"""Every function include/fftw3_api.inc declares is called at least once here, through raw ctypes
prototypes written from the header (not through fftw3_b200/binding.py), on the emulated device
layer, and its result checked.  Complements tests/test_abi.py (symbols exist) and the reference's
own harness (tests/bench.c exercises the planner entry points on the GPU)."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, I, U, D = C.c_void_p, C.c_int, C.c_uint, C.c_double
ESTIMATE = 1 << 6
FWD, BWD = -1, 1


class Iodim(C.Structure):
    _fields_ = [("n", I), ("is_", I), ("os", I)]


class Iodim64(C.Structure):
    _fields_ = [("n", C.c_ssize_t), ("is_", C.c_ssize_t), ("os", C.c_ssize_t)]


@pytest.fixture(scope="module")
def L(host_lib):
    lib = C.CDLL(host_lib.path if hasattr(host_lib, "path") else os.path.join(ROOT, "tests", "_emu", "libfftw3_b200_emu.so"))
    return lib


def proto(L, name, res, args):
    f = getattr(L, "fftw_" + name)
    f.restype, f.argtypes = res, args
    return f


def ptr(a):
    return a.ctypes.data


def run(L, plan):
    assert plan
    proto(L, "execute", None, [P])(plan)
    proto(L, "destroy_plan", None, [P])(plan)


def test_basic_complex_interfaces(L):
    rng = np.random.default_rng(0)
    for shape, name, extra in (((12,), "plan_dft_1d", None), ((6, 10), "plan_dft_2d", None), ((4, 6, 5), "plan_dft_3d", None),
                               ((3, 4, 5, 2), "plan_dft", None)):
        x = rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)
        y = np.zeros_like(x)
        if name == "plan_dft":
            n = (I * len(shape))(*shape)
            p = proto(L, name, P, [I, C.POINTER(I), P, P, I, U])(len(shape), n, ptr(x), ptr(y), FWD, ESTIMATE)
        else:
            p = proto(L, name, P, [I] * len(shape) + [P, P, I, U])(*shape, ptr(x), ptr(y), FWD, ESTIMATE)
        run(L, p)
        assert np.abs(y - np.fft.fftn(x)).max() < 1e-12


def test_basic_real_interfaces(L):
    rng = np.random.default_rng(1)
    for shape in ((16,), (6, 10), (4, 6, 5), (3, 4, 6)):
        rank = len(shape)
        x = rng.uniform(-1, 1, shape)
        cshape = shape[:-1] + (shape[-1] // 2 + 1,)
        X = np.zeros(cshape, np.complex128)
        if rank <= 3 and shape != (3, 4, 6):
            nm = "plan_dft_r2c_%dd" % rank
            p = proto(L, nm, P, [I] * rank + [P, P, U])(*shape, ptr(x), ptr(X), ESTIMATE)
        else:
            n = (I * rank)(*shape)
            p = proto(L, "plan_dft_r2c", P, [I, C.POINTER(I), P, P, U])(rank, n, ptr(x), ptr(X), ESTIMATE)
        run(L, p)
        assert np.abs(X - np.fft.rfftn(x)).max() < 1e-12
        y = np.zeros(shape)
        Xc = X.copy()
        if rank <= 3 and shape != (3, 4, 6):
            nm = "plan_dft_c2r_%dd" % rank
            p = proto(L, nm, P, [I] * rank + [P, P, U])(*shape, ptr(Xc), ptr(y), ESTIMATE)
        else:
            n = (I * rank)(*shape)
            p = proto(L, "plan_dft_c2r", P, [I, C.POINTER(I), P, P, U])(rank, n, ptr(Xc), ptr(y), ESTIMATE)
        run(L, p)
        assert np.abs(y / np.prod(shape) - x).max() < 1e-12


def test_basic_r2r_interfaces_and_new_array_execute(L):
    rng = np.random.default_rng(2)
    REDFT00, DHT, R2HC = 3, 2, 0
    x = rng.uniform(-1, 1, 9)
    y = np.zeros(9)
    run(L, proto(L, "plan_r2r_1d", P, [I, P, P, I, U])(9, ptr(x), ptr(y), REDFT00, ESTIMATE))
    want = np.array([x[0] + (-1) ** k * x[8] + 2 * sum(x[j] * np.cos(np.pi * j * k / 8) for j in range(1, 8)) for k in range(9)])
    assert np.abs(y - want).max() < 1e-12
    a = rng.uniform(-1, 1, (4, 6))
    b = np.zeros_like(a)
    run(L, proto(L, "plan_r2r_2d", P, [I, I, P, P, I, I, U])(4, 6, ptr(a), ptr(b), DHT, DHT, ESTIMATE))
    F = np.fft.fft2(a)
    # separable DHT (cas x cas), doc/reference.texi:2301-2353
    H0 = np.real(np.fft.fft(a, axis=0)) - np.imag(np.fft.fft(a, axis=0))
    want2 = np.real(np.fft.fft(H0, axis=1)) - np.imag(np.fft.fft(H0, axis=1))
    assert np.abs(b - want2).max() < 1e-12 and F.shape == (4, 6)
    c = rng.uniform(-1, 1, (3, 4, 5))
    d = np.zeros_like(c)
    p3 = proto(L, "plan_r2r_3d", P, [I, I, I, P, P, I, I, I, U])(3, 4, 5, ptr(c), ptr(d), DHT, DHT, DHT, ESTIMATE)
    run(L, p3)
    kinds = (I * 2)(R2HC, R2HC)
    n = (I * 2)(4, 6)
    p = proto(L, "plan_r2r", P, [I, C.POINTER(I), P, P, C.POINTER(I), U])(2, n, ptr(a), ptr(b), kinds, ESTIMATE)
    assert p
    a2 = rng.uniform(-1, 1, (4, 6))
    b2 = np.zeros_like(a2)
    proto(L, "execute_r2r", None, [P, P, P])(p, ptr(a2), ptr(b2))       # new arrays
    proto(L, "destroy_plan", None, [P])(p)
    # R2HC along both dims: check the (0, 0) and the pure-real row-0 entries against the FFT
    F2 = np.fft.fft2(a2)
    assert abs(b2[0, 0] - F2[0, 0].real) < 1e-12 and abs(b2[0, 3] - F2[0, 3].real) < 1e-12


def test_new_array_execute_for_real_data(L):
    rng = np.random.default_rng(3)
    n = 20
    x, X = np.zeros(n), np.zeros(n // 2 + 1, np.complex128)
    p = proto(L, "plan_dft_r2c_1d", P, [I, P, P, U])(n, ptr(x), ptr(X), ESTIMATE)
    x2, X2 = rng.uniform(-1, 1, n), np.zeros(n // 2 + 1, np.complex128)
    proto(L, "execute_dft_r2c", None, [P, P, P])(p, ptr(x2), ptr(X2))
    proto(L, "destroy_plan", None, [P])(p)
    assert np.abs(X2 - np.fft.rfft(x2)).max() < 1e-12
    q = proto(L, "plan_dft_c2r_1d", P, [I, P, P, U])(n, ptr(X), ptr(x), ESTIMATE)
    y2 = np.zeros(n)
    X3 = X2.copy()                                                  # c2r may destroy its input
    proto(L, "execute_dft_c2r", None, [P, P, P])(q, ptr(X3), ptr(y2))
    proto(L, "destroy_plan", None, [P])(q)
    assert np.abs(y2 / n - x2).max() < 1e-12


def test_guru64_interfaces(L):
    rng = np.random.default_rng(4)
    n, hm = 12, 3
    VP = C.c_void_p
    dims = (Iodim64 * 1)(Iodim64(n, 1, 1))
    how = (Iodim64 * 1)(Iodim64(hm, n, n))
    x = rng.uniform(-1, 1, (hm, n)) + 1j * rng.uniform(-1, 1, (hm, n))
    y = np.zeros_like(x)
    run(L, proto(L, "plan_guru64_dft", P, [I, VP, I, VP, P, P, I, U])(1, C.cast(dims, VP), 1, C.cast(how, VP), ptr(x), ptr(y),
                                                                      FWD, ESTIMATE))
    assert np.abs(y - np.fft.fft(x, axis=1)).max() < 1e-12
    re, im = x.real.copy(), x.imag.copy()
    ro, io = np.zeros((hm, n)), np.zeros((hm, n))
    run(L, proto(L, "plan_guru64_split_dft", P, [I, VP, I, VP, P, P, P, P, U])(1, C.cast(dims, VP), 1, C.cast(how, VP), ptr(re),
                                                                               ptr(im), ptr(ro), ptr(io), ESTIMATE))
    assert np.abs(ro + 1j * io - np.fft.fft(x, axis=1)).max() < 1e-12
    # real: r2c / c2r, interleaved and split
    h = n // 2 + 1
    r = rng.uniform(-1, 1, (hm, n))
    dr = (Iodim64 * 1)(Iodim64(n, 1, 1))
    hr = (Iodim64 * 1)(Iodim64(hm, n, h))
    X = np.zeros((hm, h), np.complex128)
    run(L, proto(L, "plan_guru64_dft_r2c", P, [I, VP, I, VP, P, P, U])(1, C.cast(dr, VP), 1, C.cast(hr, VP), ptr(r), ptr(X), ESTIMATE))
    assert np.abs(X - np.fft.rfft(r, axis=1)).max() < 1e-12
    Xr, Xi = np.zeros((hm, h)), np.zeros((hm, h))
    run(L, proto(L, "plan_guru64_split_dft_r2c", P, [I, VP, I, VP, P, P, P, U])(1, C.cast(dr, VP), 1, C.cast(hr, VP), ptr(r),
                                                                                ptr(Xr), ptr(Xi), ESTIMATE))
    assert np.abs(Xr + 1j * Xi - X).max() < 1e-12
    hc = (Iodim64 * 1)(Iodim64(hm, h, n))
    back = np.zeros((hm, n))
    Xc = X.copy()                                                   # c2r may destroy its input
    run(L, proto(L, "plan_guru64_dft_c2r", P, [I, VP, I, VP, P, P, U])(1, C.cast(dr, VP), 1, C.cast(hc, VP), ptr(Xc), ptr(back),
                                                                      ESTIMATE))
    assert np.abs(back / n - r).max() < 1e-12
    back2 = np.zeros((hm, n))
    Xr2, Xi2 = Xr.copy(), Xi.copy()
    run(L, proto(L, "plan_guru64_split_dft_c2r", P, [I, VP, I, VP, P, P, P, U])(1, C.cast(dr, VP), 1, C.cast(hc, VP), ptr(Xr2),
                                                                                ptr(Xi2), ptr(back2), ESTIMATE))
    assert np.abs(back2 / n - r).max() < 1e-12
    k = (I * 1)(5)                                       # REDFT10
    hh = (Iodim64 * 1)(Iodim64(hm, n, n))
    out = np.zeros((hm, n))
    run(L, proto(L, "plan_guru64_r2r", P, [I, VP, I, VP, P, P, C.POINTER(I), U])(1, C.cast(dr, VP), 1, C.cast(hh, VP), ptr(r), ptr(out),
                                                                                k, ESTIMATE))
    want = np.array([[2 * sum(r[b, j] * np.cos(np.pi * (j + 0.5) * kk / n) for j in range(n)) for kk in range(n)] for b in range(hm)])
    assert np.abs(out - want).max() < 1e-12


def test_memory_threads_and_misc(L, tmp_path):
    ar = proto(L, "alloc_real", P, [C.c_size_t])(100)
    ac = proto(L, "alloc_complex", P, [C.c_size_t])(100)
    assert ar and ac and ar % 16 == 0 and ac % 16 == 0
    proto(L, "free", None, [P])(ar)
    proto(L, "free", None, [P])(ac)
    assert proto(L, "init_threads", I, [])() != 0
    proto(L, "plan_with_nthreads", None, [I])(4)
    assert proto(L, "planner_nthreads", I, [])() >= 1
    CB = C.CFUNCTYPE(None, P, P, C.c_size_t, I, P)
    proto(L, "threads_set_callback", None, [P, P])(None, None)     # accepted, no CPU threads to drive
    assert CB is not None
    proto(L, "set_timelimit", None, [D])(0.5)
    proto(L, "set_timelimit", None, [D])(-1.0)                     # FFTW_NO_TIMELIMIT
    x = np.zeros(32, np.complex128)
    p = proto(L, "plan_dft_1d", P, [I, P, P, I, U])(32, ptr(x), ptr(x), FWD, 0)       # FFTW_MEASURE: deposits wisdom
    assert p
    # print_plan / fprint_plan write the same text sprint_plan returns
    libc = C.CDLL(None)
    libc.fopen.restype, libc.fopen.argtypes = P, [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [P]
    path = str(tmp_path / "plan.txt").encode()
    f = libc.fopen(path, b"w")
    proto(L, "fprint_plan", None, [P, P])(p, f)
    libc.fclose(f)
    text = open(path).read()
    sp = proto(L, "sprint_plan", P, [P])(p)
    assert "fft-pass" in text and C.string_at(sp).decode() == text
    libc.free.argtypes = [P]
    libc.free(sp)
    proto(L, "print_plan", None, [P])(p)                           # to stdout
    proto(L, "destroy_plan", None, [P])(p)
    # wisdom through callbacks and through a file name
    chars = []
    WR = C.CFUNCTYPE(None, C.c_char, P)
    w = WR(lambda c, d: chars.append(c))
    proto(L, "export_wisdom", None, [WR, P])(w, None)
    blob = b"".join(chars)
    assert blob.startswith(b"(fftw3_b200-") and b"b200_fft_pass" in blob
    wpath = str(tmp_path / "wisdom.txt").encode()
    assert proto(L, "export_wisdom_to_filename", I, [C.c_char_p])(wpath) == 1
    assert open(wpath, "rb").read() == blob
    proto(L, "forget_wisdom", None, [])()
    it = iter(blob)
    RD = C.CFUNCTYPE(I, P)
    r = RD(lambda d: next(it, -1))
    assert proto(L, "import_wisdom", I, [RD, P])(r, None) == 1
    chars.clear()
    proto(L, "export_wisdom", None, [WR, P])(w, None)
    assert b"".join(chars) == blob
    proto(L, "cleanup_threads", None, [])()
    proto(L, "cleanup", None, [])()
