"""The FFTW API behaviours a drop-in must reproduce (SURVEY.md section 9 checklist), exercised on
the real host layer with the emulated device shim.  Each test cites the reference text it
restates; items already covered elsewhere: 1, 5, 6, 8 (test_host_emulated.py::test_null_on_invalid,
test_wisdom_roundtrip), 9 (test_strided_advanced_interface, test_embedded_2d), 12
(test_new_array_execute), 16 (test_plan_introspection), 17 (test_wisdom_roundtrip)."""
import ctypes as C
import threading

import numpy as np
import pytest

from fftw3_b200 import binding as B
import fftcheck as F  # noqa: F401  (path set up by conftest)
from oracle import oracle as O

LOGICAL = {"R2HC": lambda n: n, "HC2R": lambda n: n, "DHT": lambda n: n,
           "REDFT00": lambda n: 2 * (n - 1), "RODFT00": lambda n: 2 * (n + 1),
           "REDFT01": lambda n: 2 * n, "REDFT10": lambda n: 2 * n, "REDFT11": lambda n: 2 * n,
           "RODFT01": lambda n: 2 * n, "RODFT10": lambda n: 2 * n, "RODFT11": lambda n: 2 * n}
INVERSE = {"R2HC": "HC2R", "HC2R": "R2HC", "DHT": "DHT", "REDFT00": "REDFT00", "REDFT10": "REDFT01",
           "REDFT01": "REDFT10", "REDFT11": "REDFT11", "RODFT00": "RODFT00", "RODFT10": "RODFT01",
           "RODFT01": "RODFT10", "RODFT11": "RODFT11"}


def test_planning_touches_arrays_only_when_measuring(host_lib):
    """#2 doc/reference.texi:405-416, kernel/timer.c:148-149: FFTW_ESTIMATE never touches the arrays;
    measuring modes may overwrite them (here: planning runs on scratch, arrays stay intact, which
    the contract allows)."""
    x = np.arange(256, dtype=np.float64).view(np.complex128).copy()
    keep = x.copy()
    p = host_lib.fn("d", "plan_dft_1d")(128, x.ctypes.data, x.ctypes.data, -1, B.FFTW_ESTIMATE)
    assert p and np.array_equal(x, keep)
    host_lib.destroy_plan("d", p)


@pytest.mark.parametrize("kind", sorted(LOGICAL))
def test_unnormalised_roundtrip_scales_by_logical_size(host_lib, kind):
    """#3 doc/reference.texi:432-435,904-911,936-990: kind followed by its inverse kind multiplies by
    the logical size (n, 2(n-1), 2(n+1) or 2n)."""
    n = 12
    rng = np.random.default_rng(1)
    x0 = rng.uniform(-0.5, 0.5, n)
    x, y, z = x0.copy(), np.zeros(n), np.zeros(n)
    p = host_lib.plan_many_r2r("d", [n], 1, x.ctypes.data, None, 1, n, y.ctypes.data, None, 1, n, [kind], B.FFTW_ESTIMATE)
    q = host_lib.plan_many_r2r("d", [n], 1, y.ctypes.data, None, 1, n, z.ctypes.data, None, 1, n, [INVERSE[kind]],
                              B.FFTW_ESTIMATE)
    assert p and q
    host_lib.execute("d", p)
    host_lib.execute("d", q)
    host_lib.destroy_plan("d", p)
    host_lib.destroy_plan("d", q)
    assert np.allclose(z, LOGICAL[kind](n) * x0, rtol=0, atol=1e-12)


@pytest.mark.parametrize("n", [8, 9])
def test_halfcomplex_layout(host_lib, n):
    """#4 doc/reference.texi:916-927, rdft/rdft2-rdft.c:42-74: r0 r1 ... r(n/2) i((n+1)/2-1) ... i1"""
    rng = np.random.default_rng(2)
    x = rng.uniform(-0.5, 0.5, n)
    y = np.zeros(n)
    p = host_lib.plan_many_r2r("d", [n], 1, x.ctypes.data, None, 1, n, y.ctypes.data, None, 1, n, ["R2HC"], B.FFTW_ESTIMATE)
    host_lib.execute("d", p)
    host_lib.destroy_plan("d", p)
    X = np.fft.fft(x)
    want = np.concatenate([X.real[:n // 2 + 1], X.imag[1:(n + 1) // 2][::-1]])
    assert np.allclose(y, want, atol=1e-13)


def test_guru_split_is_always_forward_backward_by_pointer_swap(host_lib):
    """#11 api/plan-guru-split-dft.h:30-31, kernel/extract-reim.c:27-36: no sign argument; passing
    (ii, ri, io, ro) computes the backward transform."""
    n = 20
    rng = np.random.default_rng(3)
    re, im = rng.uniform(-0.5, 0.5, n), rng.uniform(-0.5, 0.5, n)
    ro, io = np.zeros(n), np.zeros(n)
    p = host_lib.plan_guru_split_dft("d", [(n, 1, 1)], [], re.ctypes.data, im.ctypes.data, ro.ctypes.data, io.ctypes.data,
                                    B.FFTW_ESTIMATE)
    assert p
    host_lib.execute("d", p)
    host_lib.destroy_plan("d", p)
    assert np.allclose(ro + 1j * io, np.fft.fft(re + 1j * im), atol=1e-13)
    p = host_lib.plan_guru_split_dft("d", [(n, 1, 1)], [], im.ctypes.data, re.ctypes.data, io.ctypes.data, ro.ctypes.data,
                                    B.FFTW_ESTIMATE)
    host_lib.execute("d", p)
    host_lib.destroy_plan("d", p)
    assert np.allclose(ro + 1j * io, np.fft.ifft(re + 1j * im) * n, atol=1e-13)


def test_malloc_alignment_class(host_lib):
    """#13 kernel/align.c:24-41, tests/fftw-bench.c:230-234: fftw_malloc memory has alignment class 0"""
    for prec in ("d", "f"):
        for size in (8, 1000, 1 << 20):
            ptr = host_lib.fn(prec, "malloc")(size)
            assert ptr and host_lib.fn(prec, "alignment_of")(ptr) == 0
            host_lib.fn(prec, "free")(ptr)


def test_copy_plan_is_reference_counted(host_lib):
    """#14 api/apiplan.c:180-210: a copy stays usable after the original is destroyed"""
    x = (np.arange(32) + 0j).astype(np.complex128)
    y = np.zeros_like(x)
    p = host_lib.fn("d", "plan_dft_1d")(32, x.ctypes.data, y.ctypes.data, -1, B.FFTW_ESTIMATE)
    q = host_lib.fn("d", "copy_plan")(p)
    assert q
    host_lib.destroy_plan("d", p)
    host_lib.execute("d", q)
    assert np.allclose(y, np.fft.fft(x), atol=1e-12)
    host_lib.destroy_plan("d", q)


def test_flops_and_cost_are_reported(host_lib):
    """#15 api/flops.c:23-43: (add, mul, fma) of the plan, estimate_cost = add + mul + 2 fma"""
    x = np.zeros(1000, np.complex128)
    p = host_lib.fn("d", "plan_dft_1d")(1000, x.ctypes.data, x.ctypes.data, -1, B.FFTW_ESTIMATE)
    a, m, f = C.c_double(), C.c_double(), C.c_double()
    host_lib.fn("d", "flops")(p, C.byref(a), C.byref(m), C.byref(f))
    total = a.value + m.value + 2 * f.value
    assert 0.2 * 5 * 1000 * np.log2(1000) < total < 3 * 5 * 1000 * np.log2(1000)
    host_lib.fn("d", "estimate_cost").restype = C.c_double
    host_lib.fn("d", "cost").restype = C.c_double
    assert host_lib.fn("d", "estimate_cost")(p) >= total * 0.999
    assert host_lib.fn("d", "cost")(p) >= 0
    host_lib.destroy_plan("d", p)


def test_execute_is_reentrant(host_lib):
    """#18 doc/threads.texi:225-270: fftw_execute* may be called concurrently, also on one plan with
    new arrays."""
    n, hm = 256, 6
    rng = np.random.default_rng(4)
    x0 = rng.uniform(-0.5, 0.5, (hm, n)) + 1j * rng.uniform(-0.5, 0.5, (hm, n))
    y0 = np.zeros_like(x0)
    p = host_lib.plan_many_dft("d", [n], hm, x0.ctypes.data, None, 1, n, y0.ctypes.data, None, 1, n, -1, B.FFTW_ESTIMATE)
    assert p
    ex = host_lib.fn("d", "execute_dft")
    ins = [(rng.uniform(-0.5, 0.5, (hm, n)) + 1j * rng.uniform(-0.5, 0.5, (hm, n))) for _ in range(8)]
    outs = [np.zeros_like(a) for a in ins]
    errs = []

    def work(i):
        try:
            for _ in range(5):
                ex(p, ins[i].ctypes.data, outs[i].ctypes.data)
        except Exception as e:      # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(ins))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    host_lib.destroy_plan("d", p)
    assert not errs
    for a, b in zip(ins, outs):
        assert O.rel_l2(b, np.fft.fft(a, axis=1)) < 1e-14


def test_c2r_ignores_imag_of_dc_and_nyquist(host_lib):
    """#19 rdft/rdft2-rdft.c:61-74: the halfcomplex packer never reads imag(DC) / imag(Nyquist)"""
    n = 16
    rng = np.random.default_rng(5)
    x = rng.uniform(-0.5, 0.5, n)
    X = np.fft.rfft(x)
    Xd = X.copy()
    Xd[0] += 3.25j
    Xd[-1] -= 1.5j
    y = np.zeros(n)
    p = host_lib.plan_many_dft_c2r("d", [n], 1, Xd.ctypes.data, None, 1, n // 2 + 1, y.ctypes.data, None, 1, n, B.FFTW_ESTIMATE)
    assert p
    host_lib.execute("d", p)
    host_lib.destroy_plan("d", p)
    assert np.allclose(y, n * x, atol=1e-12)


def test_multidimensional_r2r_is_separable(host_lib):
    """#20 doc/reference.texi:2353-2400, rdft/rank-geq2.c:135-180: kind[i] applies along dimension i"""
    rng = np.random.default_rng(6)
    x = rng.uniform(-0.5, 0.5, (6, 10))
    y = np.zeros_like(x)
    p = host_lib.plan_many_r2r("d", [6, 10], 1, x.ctypes.data, None, 1, 60, y.ctypes.data, None, 1, 60,
                              ["RODFT10", "REDFT01"], B.FFTW_ESTIMATE)
    host_lib.execute("d", p)
    host_lib.destroy_plan("d", p)
    a = np.stack([O.r2r(x[:, j].copy(), ["RODFT10"], rank=1) for j in range(10)], axis=1)
    b = np.stack([O.r2r(a[i].copy(), ["REDFT01"], rank=1) for i in range(6)], axis=0)
    assert O.rel_l2(y, b) < 1e-14


def test_planner_entry_points_are_thread_safe(host_lib):
    """#18 doc/threads.texi:225-270, api/apiplan.c:23-29: after fftw_make_planner_thread_safe() plan
    creation / destruction / wisdom calls may come from any thread (here they always serialise on one
    process-wide lock; ThreadSanitizer runs of the same scenario are clean)."""
    host_lib.fn("d", "make_planner_thread_safe")()
    errs = []

    def work(i):
        try:
            rng = np.random.default_rng(i)
            for it in range(6):
                n = 48 + 16 * ((i + it) % 5)
                x = (rng.uniform(-0.5, 0.5, n) + 1j * rng.uniform(-0.5, 0.5, n))
                y = np.zeros_like(x)
                flags = B.FFTW_ESTIMATE if it % 2 else B.FFTW_MEASURE
                x0 = x.copy()
                p = host_lib.fn("d", "plan_dft_1d")(n, x.ctypes.data, y.ctypes.data, -1, flags)
                assert p
                x[:] = x0
                host_lib.execute("d", p)
                assert O.rel_l2(y, np.fft.fft(x0)) < 1e-14
                assert host_lib.export_wisdom_to_string("d").startswith("(fftw3_b200-")
                host_lib.destroy_plan("d", p)
        except Exception as e:      # pragma: no cover
            errs.append(repr(e))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs


@pytest.mark.parametrize("shape", [(16,), (9,), (6, 10), (5, 4, 7)])
def test_guru_split_r2c_c2r_and_new_array_execute(host_lib, shape):
    """api/plan-guru-split-dft-r2c.h, plan-guru-split-dft-c2r.h, execute-split-dft-r2c.c, execute-split-dft-c2r.c:
    real <-> split (re, im) half spectrum, planned on one set of arrays and executed on another."""
    rank = len(shape)
    rng = np.random.default_rng(8)
    cshape = shape[:-1] + (shape[-1] // 2 + 1,)

    def rm_strides(sh):
        st, acc = [], 1
        for n in reversed(sh):
            st.append(acc)
            acc *= n
        return list(reversed(st))

    rs, cs = rm_strides(shape), rm_strides(cshape)
    x, ro, io = np.zeros(shape), np.zeros(cshape), np.zeros(cshape)
    dims = (B.Iodim * rank)(*[B.Iodim(shape[i], rs[i], cs[i]) for i in range(rank)])
    none = (B.Iodim * 1)(B.Iodim(1, 0, 0))
    VP = C.c_void_p
    p = host_lib.fn("d", "plan_guru_split_dft_r2c")(rank, C.cast(dims, VP), 0, C.cast(none, VP), x.ctypes.data, ro.ctypes.data,
                                                   io.ctypes.data, B.FFTW_ESTIMATE)
    assert p
    x2 = rng.uniform(-0.5, 0.5, shape)
    ro2, io2 = np.zeros(cshape), np.zeros(cshape)
    host_lib.fn("d", "execute_split_dft_r2c")(p, x2.ctypes.data, ro2.ctypes.data, io2.ctypes.data)
    host_lib.destroy_plan("d", p)
    want = np.fft.rfftn(x2)
    assert np.abs(ro2 + 1j * io2 - want).max() < 1e-12
    # back: split half spectrum -> real (input may be destroyed: hand over copies)
    dims = (B.Iodim * rank)(*[B.Iodim(shape[i], cs[i], rs[i]) for i in range(rank)])
    ri, ii, y = want.real.copy(), want.imag.copy(), np.zeros(shape)
    p = host_lib.fn("d", "plan_guru_split_dft_c2r")(rank, C.cast(dims, VP), 0, C.cast(none, VP), ri.ctypes.data, ii.ctypes.data,
                                                   y.ctypes.data, B.FFTW_ESTIMATE)
    assert p
    ri2, ii2, y2 = want.real.copy(), want.imag.copy(), np.zeros(shape)
    host_lib.fn("d", "execute_split_dft_c2r")(p, ri2.ctypes.data, ii2.ctypes.data, y2.ctypes.data)
    host_lib.destroy_plan("d", p)
    assert np.abs(y2 / np.prod(shape) - x2).max() < 1e-12


def test_execute_split_dft_on_new_arrays(host_lib):
    """api/execute-split-dft.c:25-29"""
    n = 24
    rng = np.random.default_rng(9)
    a = [np.zeros(n) for _ in range(4)]
    p = host_lib.plan_guru_split_dft("d", [(n, 1, 1)], [], a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data,
                                    a[3].ctypes.data, B.FFTW_ESTIMATE)
    assert p
    re, im, ro, io = rng.uniform(-0.5, 0.5, n), rng.uniform(-0.5, 0.5, n), np.zeros(n), np.zeros(n)
    host_lib.fn("d", "execute_split_dft")(p, re.ctypes.data, im.ctypes.data, ro.ctypes.data, io.ctypes.data)
    host_lib.destroy_plan("d", p)
    assert np.abs(ro + 1j * io - np.fft.fft(re + 1j * im)).max() < 1e-13
