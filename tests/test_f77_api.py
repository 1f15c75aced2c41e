"""The legacy Fortran 77 interface (dfftw_* / sfftw_* subroutines; reference api/f77api.c:33-160,
api/f77funcs.h, threads/f77funcs.h, doc/legacy-fortran.texi): every argument by reference, plan returned
through the first argument, dimensions (and r2r kinds) in Fortran order = reversed C order.  A Fortran array
A(nx, ny) is the C / numpy array of shape (ny, nx).  Runs on the emulated device layer (CPU) and on the GPU."""
import ctypes as C

import numpy as np
import pytest

from fftw3_b200 import binding as B
from oracle import oracle as O

I = C.c_int
P = C.c_void_p


def ref(v):
    return C.byref(I(v))


def ints(v):
    return (I * len(v))(*v)


@pytest.fixture()
def L(host_lib):
    return C.CDLL(host_lib.path)


def test_complex_2d_3d_dimension_reversal(L):
    rng = np.random.default_rng(0)
    nx, ny, nz = 12, 10, 6
    x = rng.standard_normal((ny, nx)) + 1j * rng.standard_normal((ny, nx))          # Fortran A(nx, ny)
    y = np.zeros_like(x)
    plan = P()
    L.dfftw_plan_dft_2d_(C.byref(plan), ref(nx), ref(ny), P(x.ctypes.data), P(y.ctypes.data), ref(-1), ref(B.FFTW_ESTIMATE))
    assert plan.value
    L.dfftw_execute_(C.byref(plan))
    assert O.rel_l2(y, O.dft(x)) < 1e-14
    # new-array execute, g77 spelling
    x2 = rng.standard_normal((ny, nx)) + 1j * rng.standard_normal((ny, nx))
    y2 = np.zeros_like(x2)
    L.dfftw_execute_dft__(C.byref(plan), P(x2.ctypes.data), P(y2.ctypes.data))
    assert O.rel_l2(y2, O.dft(x2)) < 1e-14
    L.dfftw_destroy_plan_(C.byref(plan))
    x3 = (rng.standard_normal((nz, ny, nx)) + 1j * rng.standard_normal((nz, ny, nx))).astype(np.complex64)
    y3 = np.zeros_like(x3)
    L.sfftw_plan_dft_3d_(C.byref(plan), ref(nx), ref(ny), ref(nz), P(x3.ctypes.data), P(y3.ctypes.data), ref(1), ref(B.FFTW_ESTIMATE))
    assert plan.value
    L.sfftw_execute_(C.byref(plan))
    L.sfftw_destroy_plan_(C.byref(plan))
    assert O.rel_l2(y3, O.dft(x3, sign=1)) < 2e-6
    # general rank through the array form: n = (nx, ny, nz) in Fortran order
    yd = np.zeros((nz, ny, nx), np.complex128)
    xd = x3.astype(np.complex128)
    L.dfftw_plan_dft_(C.byref(plan), ref(3), ints([nx, ny, nz]), P(xd.ctypes.data), P(yd.ctypes.data), ref(-1), ref(B.FFTW_ESTIMATE))
    assert plan.value
    L.dfftw_execute_(C.byref(plan))
    L.dfftw_destroy_plan_(C.byref(plan))
    assert O.rel_l2(yd, O.dft(xd)) < 1e-14


def test_many_real_and_r2r_with_reversed_arguments(L):
    rng = np.random.default_rng(1)
    nx, ny, hm = 16, 6, 3
    # advanced interface: howmany transforms of A(nx, ny), Fortran-order n / embed
    x = rng.standard_normal((hm, ny, nx)) + 1j * rng.standard_normal((hm, ny, nx))
    y = np.zeros_like(x)
    plan = P()
    L.dfftw_plan_many_dft_(C.byref(plan), ref(2), ints([nx, ny]), ref(hm), P(x.ctypes.data), ints([nx, ny]), ref(1), ref(nx * ny),
                           P(y.ctypes.data), ints([nx, ny]), ref(1), ref(nx * ny), ref(-1), ref(B.FFTW_ESTIMATE))
    assert plan.value
    L.dfftw_execute_(C.byref(plan))
    L.dfftw_destroy_plan_(C.byref(plan))
    assert O.rel_l2(y, O.dft(x, rank=2)) < 1e-14
    # r2c 2-d: real A(nx, ny) -> complex (nx/2+1, ny)
    r = rng.standard_normal((ny, nx))
    c = np.zeros((ny, nx // 2 + 1), np.complex128)
    L.dfftw_plan_dft_r2c_2d_(C.byref(plan), ref(nx), ref(ny), P(r.ctypes.data), P(c.ctypes.data), ref(B.FFTW_ESTIMATE))
    assert plan.value
    L.dfftw_execute_(C.byref(plan))
    L.dfftw_destroy_plan_(C.byref(plan))
    assert O.rel_l2(c, O.r2c(r, rank=2)) < 1e-14
    back = np.zeros_like(r)
    c2 = c.copy()
    L.dfftw_plan_dft_c2r_2d_(C.byref(plan), ref(nx), ref(ny), P(c2.ctypes.data), P(back.ctypes.data), ref(B.FFTW_ESTIMATE))
    assert plan.value
    L.dfftw_execute_dft_c2r_(C.byref(plan), P(c2.ctypes.data), P(back.ctypes.data))
    L.dfftw_destroy_plan_(C.byref(plan))
    assert O.rel_l2(back / (nx * ny), r) < 1e-14
    # r2r: kinds in Fortran order too -- kind(1) belongs to the fastest dimension nx
    out = np.zeros_like(r)
    kinds_f = [B.R2R_KINDS["REDFT10"], B.R2R_KINDS["RODFT00"]]          # (x, y)
    L.dfftw_plan_r2r_(C.byref(plan), ref(2), ints([nx, ny]), P(r.ctypes.data), P(out.ctypes.data), ints(kinds_f), ref(B.FFTW_ESTIMATE))
    assert plan.value
    L.dfftw_execute_r2r_(C.byref(plan), P(r.ctypes.data), P(out.ctypes.data))
    L.dfftw_destroy_plan_(C.byref(plan))
    assert O.rel_l2(out, O.r2r(r, ["RODFT00", "REDFT10"], rank=2)) < 1e-13
    out2 = np.zeros_like(r)
    L.dfftw_plan_r2r_2d_(C.byref(plan), ref(nx), ref(ny), P(r.ctypes.data), P(out2.ctypes.data), ref(kinds_f[0]), ref(kinds_f[1]),
                         ref(B.FFTW_ESTIMATE))
    assert plan.value
    L.dfftw_execute_(C.byref(plan))
    L.dfftw_destroy_plan_(C.byref(plan))
    assert np.allclose(out2, out, rtol=0, atol=1e-12)


def test_guru_wisdom_callbacks_and_introspection(L):
    rng = np.random.default_rng(2)
    n, hm = 64, 5
    x = rng.standard_normal((hm, n)) + 1j * rng.standard_normal((hm, n))
    y = np.zeros_like(x)
    plan = P()
    L.dfftw_plan_guru_dft_(C.byref(plan), ref(1), ints([n]), ints([1]), ints([1]), ref(1), ints([hm]), ints([n]), ints([n]),
                           P(x.ctypes.data), P(y.ctypes.data), ref(-1), ref(B.FFTW_MEASURE))
    assert plan.value
    x[...] = rng.standard_normal((hm, n)) + 1j * rng.standard_normal((hm, n))
    L.dfftw_execute_(C.byref(plan))
    assert O.rel_l2(y, O.dft(x, rank=1)) < 1e-14
    add, mul, fma, cost = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    L.dfftw_flops_(C.byref(plan), C.byref(add), C.byref(mul), C.byref(fma))
    L.dfftw_estimate_cost_(C.byref(cost), C.byref(plan))
    assert add.value > 0 and cost.value > 0
    cp = P()
    L.dfftw_copy_plan_(C.byref(cp), C.byref(plan))
    assert cp.value
    L.dfftw_destroy_plan_(C.byref(cp))
    L.dfftw_destroy_plan_(C.byref(plan))
    # wisdom through the Fortran character callbacks
    chars = []
    WR = C.CFUNCTYPE(None, C.POINTER(C.c_char), P)      # ONE character by reference (not a C string: c_char_p would strlen it)
    wr = WR(lambda c, d: chars.append(c[0]))
    L.dfftw_export_wisdom_(wr, None)
    text = b"".join(chars)
    assert text.startswith(b"(fftw3_b200-") and b"b200_fft_pass" in text
    L.dfftw_forget_wisdom_()
    pos = [0]
    RD = C.CFUNCTYPE(None, C.POINTER(I), P)

    def rd(pc, d):
        pc[0] = text[pos[0]] if pos[0] < len(text) else -1
        pos[0] += 1
    ok = I(0)
    L.dfftw_import_wisdom_(C.byref(ok), RD(rd), None)
    assert ok.value == 1
    nthr, okay = I(0), I(0)
    L.dfftw_init_threads_(C.byref(okay))
    L.dfftw_plan_with_nthreads_(ref(4))
    L.dfftw_planner_nthreads_(C.byref(nthr))
    assert okay.value != 0 and nthr.value >= 1
    t = C.c_double(-1.0)
    L.dfftw_set_timelimit_(C.byref(t))
