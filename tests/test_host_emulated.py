"""Host-layer logic (API validation, tensor canonicalisation, pass-list
construction, wisdom) and the kernels' index algebra, exercised WITHOUT a GPU
through the unit-test double of the device shim (tests/emu).  The same checks
run against the real CUDA kernels in test_gpu_parity.py."""
import ctypes as C

import numpy as np
import pytest

import fftcheck as F
from fftw3_b200 import binding as B
from oracle import oracle as O

PRECS = ["d", "f"]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 12, 13, 16, 17, 25, 30, 31, 64, 100, 121, 128, 169, 243, 256, 1000, 1009, 1024])
def test_c2c_1d(emu_lib, prec, n):
    err, tol = F.c2c(emu_lib, prec, (n,), howmany=3)
    assert err <= tol


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("shape,inplace,sign", [((16, 12), False, -1), ((8, 6, 10), True, -1), ((5, 7), True, 1),
                                                ((4, 4, 4, 3), False, 1), ((1, 8), False, -1), ((8, 1, 3), False, -1)])
def test_c2c_nd(emu_lib, prec, shape, inplace, sign):
    err, tol = F.c2c(emu_lib, prec, shape, howmany=2, inplace=inplace, sign=sign)
    assert err <= tol


@pytest.mark.parametrize("prec", PRECS)
def test_c2c_four_step(emu_lib, prec):
    err, tol = F.c2c(emu_lib, prec, (16384,), howmany=2)
    assert err <= tol
    err, tol = F.c2c(emu_lib, prec, (2 * 3 * 5 * 7 * 11 * 13,), howmany=1)   # 30030 mixed radix
    assert err <= tol


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("shape", [(16,), (18,), (15,), (9,), (2,), (1,), (4, 6), (3, 5, 8), (6, 7), (1009,), (37,)])
@pytest.mark.parametrize("inplace", [False, True])
def test_r2c_c2r(emu_lib, prec, shape, inplace):
    err, tol = F.r2c(emu_lib, prec, shape, howmany=2, inplace=inplace)
    assert err <= tol
    err, tol = F.c2r(emu_lib, prec, shape, howmany=2, inplace=inplace)
    assert err <= tol


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("unfused", [False, True])
@pytest.mark.parametrize("shape,howmany", [((1 << 16,), 3), ((4 * 3 ** 9,), 2), ((3, 1 << 15), 1)])
def test_long_even_real_lines_split_and_merge_fused(emu_lib, prec, unfused, shape, howmany, monkeypatch):
    """Even-size real transforms whose half-size complex transform is a four-step: the c2r merge rides on the
    load of the first pass (B2D_LOAD_C2R_MERGE), the r2c split on the store of the last (specialised kernels
    only, so the emulator keeps that one apart) -- rdft/ct-hc2c.c:146-273 fuses them into the twiddle codelets
    the same way.  FFTW3_B200_C2R_UNFUSED / _R2C_UNFUSED keep them as passes of their own."""
    if unfused:
        monkeypatch.setenv("FFTW3_B200_C2R_UNFUSED", "1")
        monkeypatch.setenv("FFTW3_B200_R2C_UNFUSED", "1")
    err, tol = F.c2r(emu_lib, prec, shape, howmany=howmany)
    assert err <= tol
    err, tol = F.c2r(emu_lib, prec, shape, howmany=howmany, inplace=True)
    assert err <= tol
    err, tol = F.r2c(emu_lib, prec, shape, howmany=howmany)
    assert err <= tol
    n = shape[-1]
    dt, cdt = (np.float64, np.complex128) if prec == "d" else (np.float32, np.complex64)
    x = np.zeros(n // 2 + 1, dtype=cdt)
    y = np.zeros(n, dtype=dt)
    p = emu_lib.plan_many_dft_c2r(prec, [n], 1, x.ctypes.data, None, 1, n // 2 + 1, y.ctypes.data, None, 1, n, B.FFTW_ESTIMATE)
    txt = emu_lib.sprint_plan(prec, p)
    emu_lib.destroy_plan(prec, p)
    assert ("realop" in txt) == unfused, txt
    assert txt.count("fft-pass") == 2, txt


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("kind", list(B.R2R_KINDS))
@pytest.mark.parametrize("n", [2, 3, 8, 9, 16, 37])
def test_r2r_1d(emu_lib, prec, kind, n):
    err, tol = F.r2r(emu_lib, prec, (n,), [kind], howmany=3)
    assert err <= tol


@pytest.mark.parametrize("prec", PRECS)
def test_r2r_nd(emu_lib, prec):
    err, tol = F.r2r(emu_lib, prec, (8, 6), ["REDFT10", "RODFT11"], howmany=2)
    assert err <= tol
    err, tol = F.r2r(emu_lib, prec, (5, 12), ["R2HC", "DHT"], howmany=2, inplace=True)
    assert err <= tol
    err, tol = F.r2r(emu_lib, prec, (4, 3, 5), ["REDFT00", "RODFT00", "REDFT01"], howmany=1)
    assert err <= tol


def test_strided_advanced_interface(emu_lib):
    """plan_many_dft with stride/dist/embed: element (j,k) at j*stride + k*dist
    (doc/reference.texi:1061-1075)."""
    rng = np.random.default_rng(3)
    n, howmany, istride, idist, ostride, odist = 12, 5, 3, 40, 2, 30
    xin = F.rand_complex(rng, (howmany * idist + n * istride,), "d")
    out = np.full(howmany * odist + n * ostride, 7 + 7j, dtype=np.complex128)
    p = emu_lib.plan_many_dft("d", [n], howmany, xin.ctypes.data, None, istride, idist, out.ctypes.data, None,
                              ostride, odist, -1, B.FFTW_ESTIMATE)
    assert p
    emu_lib.execute("d", p)
    emu_lib.destroy_plan("d", p)
    touched = np.zeros(out.shape, bool)
    for k in range(howmany):
        ref = O.dft(xin[k * idist:k * idist + n * istride:istride])
        got = out[k * odist:k * odist + n * ostride:ostride]
        touched[k * odist:k * odist + n * ostride:ostride] = True
        assert O.rel_l2(got, ref) < 1e-15
    assert np.all(out[~touched] == 7 + 7j), "wrote outside the described output locations"


def test_embedded_2d(emu_lib):
    rng = np.random.default_rng(4)
    n = (6, 10)
    inembed, onembed = (8, 16), (6, 12)
    big = F.rand_complex(rng, inembed, "d")
    out = np.zeros(onembed, dtype=np.complex128)
    p = emu_lib.plan_many_dft("d", n, 1, big.ctypes.data, inembed, 1, 0, out.ctypes.data, onembed, 1, 0, -1,
                              B.FFTW_ESTIMATE)
    assert p
    emu_lib.execute("d", p)
    emu_lib.destroy_plan("d", p)
    assert O.rel_l2(out[:6, :10], O.dft(big[:6, :10])) < 1e-15


def test_guru_transposed_and_split(emu_lib):
    """guru interface: arbitrary strides (here a transposed output) and split arrays
    (api/plan-guru-dft.h:24-44, plan-guru-split-dft.h:24-39)."""
    rng = np.random.default_rng(5)
    n0, n1, hm = 8, 6, 3
    x = F.rand_complex(rng, (hm, n0, n1), "d")
    y = np.zeros((hm, n1, n0), dtype=np.complex128)        # transposed output
    dims = [(n0, n1, 1), (n1, 1, n0)]
    how = [(hm, n0 * n1, n0 * n1)]
    p = emu_lib.plan_guru_dft("d", dims, how, x.ctypes.data, y.ctypes.data, -1, B.FFTW_ESTIMATE)
    assert p
    emu_lib.execute("d", p)
    emu_lib.destroy_plan("d", p)
    ref = O.dft(x, rank=2)
    assert O.rel_l2(y.transpose(0, 2, 1), ref) < 1e-15
    # guru64 + split arrays, backward obtained by swapping re/im arguments
    ri, ii = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    ro, io = np.zeros_like(ri), np.zeros_like(ii)
    dims = [(n0, n1, n1), (n1, 1, 1)]
    p = emu_lib.plan_guru_split_dft("d", dims, how, ii.ctypes.data, ri.ctypes.data, io.ctypes.data,
                                    ro.ctypes.data, B.FFTW_ESTIMATE)
    assert p
    emu_lib.execute("d", p)
    emu_lib.destroy_plan("d", p)
    assert O.rel_l2(ro + 1j * io, O.dft(x, sign=+1, rank=2)) < 1e-15


def test_new_array_execute(emu_lib):
    rng = np.random.default_rng(6)
    n = 48
    a, b = F.rand_complex(rng, (n,), "d"), np.zeros(n, np.complex128)
    p = emu_lib.fn("d", "plan_dft_1d")(n, a.ctypes.data, b.ctypes.data, 1, B.FFTW_ESTIMATE)
    assert p
    c, d = F.rand_complex(rng, (n,), "d"), np.zeros(n, np.complex128)
    emu_lib.fn("d", "execute_dft")(p, c.ctypes.data, d.ctypes.data)
    assert np.all(b == 0)
    assert O.rel_l2(d, O.dft(c, sign=+1)) < 1e-15
    p2 = emu_lib.fn("d", "copy_plan")(p)
    emu_lib.destroy_plan("d", p)
    emu_lib.execute("d", p2)                      # still alive through the copy
    assert O.rel_l2(b, O.dft(a, sign=+1)) < 1e-15
    emu_lib.destroy_plan("d", p2)


def test_null_on_invalid(emu_lib):
    x = np.zeros(64, np.complex128)
    f = emu_lib.fn("d", "plan_dft_1d")
    assert not f(0, x.ctypes.data, x.ctypes.data, -1, B.FFTW_ESTIMATE)        # n <= 0
    assert not f(-3, x.ctypes.data, x.ctypes.data, -1, B.FFTW_ESTIMATE)
    assert not emu_lib.plan_many_dft("d", [8], -1, x.ctypes.data, None, 1, 8, x.ctypes.data, None, 1, 8, -1,
                                     B.FFTW_ESTIMATE)                          # howmany < 0
    # multi-dimensional out-of-place c2r cannot preserve its input
    y = np.zeros(64, np.float64)
    assert not emu_lib.plan_many_dft_c2r("d", [4, 6], 1, x.ctypes.data, None, 1, 16, y.ctypes.data, None, 1, 24,
                                         B.FFTW_ESTIMATE | B.FFTW_PRESERVE_INPUT)
    # REDFT00 of size 1 has logical size 0
    assert not emu_lib.plan_many_r2r("d", [1], 1, y.ctypes.data, None, 1, 1, y.ctypes.data, None, 1, 1,
                                     ["REDFT00"], B.FFTW_ESTIMATE)
    # howmany == 0 is a valid no-op plan (dft/nop.c:35-52)
    p = emu_lib.plan_many_dft("d", [8], 0, x.ctypes.data, None, 1, 8, x.ctypes.data, None, 1, 8, -1, B.FFTW_ESTIMATE)
    assert p
    emu_lib.execute("d", p)
    emu_lib.destroy_plan("d", p)


def test_wisdom_roundtrip(emu_lib):
    emu_lib.fn("d", "forget_wisdom")()
    x = np.zeros((4, 64), np.complex128)
    mk = lambda flags: emu_lib.plan_many_dft("d", [64], 4, x.ctypes.data, None, 1, 64, x.ctypes.data, None, 1, 64,
                                             -1, flags)
    assert not mk(B.FFTW_WISDOM_ONLY | B.FFTW_MEASURE)          # nothing known yet
    p = mk(B.FFTW_MEASURE)
    assert p
    emu_lib.destroy_plan("d", p)
    s = emu_lib.export_wisdom_to_string("d")
    assert s.startswith("(fftw3_b200-") and "b200_fft_pass" in s
    emu_lib.fn("d", "forget_wisdom")()
    assert not mk(B.FFTW_WISDOM_ONLY | B.FFTW_MEASURE)
    assert emu_lib.fn("d", "import_wisdom_from_string")(s.encode()) == 1
    p = mk(B.FFTW_WISDOM_ONLY | B.FFTW_MEASURE)
    assert p
    emu_lib.destroy_plan("d", p)
    # wrong precision / malformed input is rejected wholesale
    assert emu_lib.fn("f", "import_wisdom_from_string")(s.encode()) == 0
    assert emu_lib.fn("d", "import_wisdom_from_string")(b"(fftw-3.3.11 fftw_wisdom #x0)") == 0
    assert emu_lib.fn("d", "import_wisdom_from_string")(s[:-10].encode()) == 0
    # wisdom measured on another device / kernel registry (different configuration signature in the header
    # line, kernel/planner.c:847-852) is rejected wholesale
    import re
    m = re.match(r"\((\S+) (\S+) #x([0-9a-f]+)", s)
    other = s.replace("#x" + m.group(3), "#x%x" % (int(m.group(3), 16) ^ 0x5a5a), 1)
    emu_lib.fn("d", "forget_wisdom")()
    assert emu_lib.fn("d", "import_wisdom_from_string")(other.encode()) == 0
    assert not mk(B.FFTW_WISDOM_ONLY | B.FFTW_MEASURE)


def test_plan_introspection(emu_lib):
    x = np.zeros(1024, np.complex128)
    p = emu_lib.fn("d", "plan_dft_1d")(1024, x.ctypes.data, x.ctypes.data, -1, B.FFTW_ESTIMATE)
    s = emu_lib.sprint_plan("d", p)
    assert "fft-pass" in s and "n=1024" in s
    a, m, f = C.c_double(), C.c_double(), C.c_double()
    emu_lib.fn("d", "flops")(p, C.byref(a), C.byref(m), C.byref(f))
    assert a.value > 0 and emu_lib.fn("d", "estimate_cost")(p) > 0
    emu_lib.destroy_plan("d", p)
    ptr = emu_lib.fn("d", "malloc")(1000)
    assert ptr and emu_lib.fn("d", "alignment_of")(ptr) == 0
    emu_lib.fn("d", "free")(ptr)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n", [2053, 4100])
def test_large_bluestein(emu_lib, prec, n):
    """prime factors > 13 with a padded length beyond one CTA: five-step Bluestein through scratch"""
    err, tol = F.c2c(emu_lib, prec, (n,), howmany=2)
    assert err <= tol
    if n == 2053:
        err, tol = F.r2c(emu_lib, prec, (n,), howmany=2)
        assert err <= tol
        err, tol = F.c2r(emu_lib, prec, (n,), howmany=2)
        assert err <= tol


def test_inplace_with_different_strides(emu_lib):
    """in place, is != os: solved through scratch like the reference's buffered/indirect solvers
    (dft/indirect.c:55-108); includes the in-place transpose idiom (rank-0 guru transform)."""
    import ctypes as C
    rng = np.random.default_rng(8)
    x = F.rand_complex(rng, (6, 6), "d")
    x0 = x.copy()
    p = emu_lib.plan_guru_dft("d", [(6, 6, 1), (6, 1, 6)], [], x.ctypes.data, x.ctypes.data, -1, B.FFTW_ESTIMATE)
    assert p
    emu_lib.execute("d", p)
    emu_lib.destroy_plan("d", p)
    assert O.rel_l2(x.T, O.dft(x0)) < 1e-15
    a = np.arange(12, dtype=np.float64).reshape(3, 4).copy()
    a0 = a.copy()
    h = (B.Iodim * 2)(B.Iodim(3, 4, 1), B.Iodim(4, 1, 3))
    p = emu_lib.fn("d", "plan_guru_r2r")(0, None, 2, C.cast(h, C.c_void_p), a.ctypes.data, a.ctypes.data, None,
                                         B.FFTW_ESTIMATE)
    assert p
    emu_lib.execute("d", p)
    emu_lib.destroy_plan("d", p)
    assert np.array_equal(a.reshape(-1), a0.T.reshape(-1))
    y = F.rand_complex(rng, (40,), "d")
    y0 = y.copy()
    p = emu_lib.plan_many_dft("d", [8], 2, y.ctypes.data, None, 1, 8, y.ctypes.data, None, 2, 16, -1, B.FFTW_ESTIMATE)
    assert p
    emu_lib.execute("d", p)
    emu_lib.destroy_plan("d", p)
    for k in range(2):
        assert O.rel_l2(y[16 * k:16 * k + 16:2], O.dft(y0[8 * k:8 * k + 8])) < 1e-15


@pytest.mark.parametrize("shape,kinds,inplace", [
    ((4096, 6), ("REDFT10", "RODFT01"), False),
    ((3000, 5), ("REDFT01", "R2HC"), True),
    ((2, 2500, 3), ("DHT", "RODFT11", "HC2R"), False),
    ((2049, 4), ("REDFT00", "REDFT11"), True),
])
@pytest.mark.parametrize("transposes", [False, True])
def test_r2r_long_strided_lines(emu_lib, shape, kinds, inplace, transposes, monkeypatch):
    """r2r dimensions whose strided lines are too long for a tile of them to share a CTA: a dense 2-d array
    runs two line passes that store their lines transposed (through scratch and back); the general case
    (batches, more dimensions, or FFTW3_B200_R2R_TRANSPOSES) brackets the fused pass with transposed scratch
    lines (dft/indirect-transpose.c strategy).  The other dimensions run the PRE/POST maps inside one pass
    (device/r2r_maps.cuh)."""
    if transposes:
        monkeypatch.setenv("FFTW3_B200_R2R_TRANSPOSES", "1")
    err, tol = F.r2r(emu_lib, "d", shape, list(kinds), inplace=inplace)
    assert err <= tol, (shape, kinds, err)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n", [5, 7, 11, 13, 17, 31, 101, 1009, 2017])
def test_rader_primes(emu_lib, prec, n, monkeypatch):
    """Primes whose n - 1 is smooth: Rader's algorithm in one pass (dft/rader.c:95-165) --
    generator-power permutation, length n-1 cyclic convolution, inverse permutation."""
    monkeypatch.setenv("FFTW3_B200_PRIME", "rader")
    for sign in (-1, 1):
        err, tol = F.c2c(emu_lib, prec, (n,), howmany=3, sign=sign)
        assert err <= 4 * tol, (n, sign, err, tol)
    err, tol = F.c2c(emu_lib, prec, (6, n), howmany=1, inplace=True)
    assert err <= 4 * tol
    x = np.zeros(n, dtype=np.complex128 if prec == "d" else np.complex64)
    p = emu_lib.plan_many_dft(prec, [n], 1, x.ctypes.data, None, 1, n, x.ctypes.data, None, 1, n, -1, B.FFTW_ESTIMATE)
    assert ("rader" in emu_lib.sprint_plan(prec, p)) == (n > 13)      # radices up to 13 are direct butterflies
    emu_lib.destroy_plan(prec, p)


def test_host_arrays_batch_pipeline(emu_lib, monkeypatch):
    """Batched problems on host arrays are cut along the outermost batch dimension into chunks that run through three
    chunk plans (upload | passes | download overlap on the GPU; here the chunks run one after the other, which checks
    offsets, regions and staging).  Batches that interleave in memory are not cut."""
    monkeypatch.setenv("FFTW3_B200_PIPE_MIN_KB", "1")
    rng = np.random.default_rng(11)
    # c2c, out of place and in place, 24 transforms -> 8 chunks of 3
    for inplace in (False, True):
        x = F.rand_complex(rng, (24, 96), "d")
        x0 = x.copy()
        y = x if inplace else np.zeros_like(x)
        p = emu_lib.plan_many_dft("d", [96], 24, x.ctypes.data, None, 1, 96, y.ctypes.data, None, 1, 96, -1, B.FFTW_ESTIMATE)
        assert "8 chunks pipelined" in emu_lib.sprint_plan("d", p)
        emu_lib.execute("d", p)
        assert O.rel_l2(y, O.dft(x0, rank=1)) < 1e-15
        # new-array execute on other host arrays goes through the same pipeline
        a = F.rand_complex(rng, (24, 96), "d")
        a0 = a.copy()
        b = a if inplace else np.zeros_like(a)
        emu_lib.fn("d", "execute_dft")(p, a.ctypes.data, b.ctypes.data)
        assert O.rel_l2(b, O.dft(a0, rank=1)) < 1e-15
        emu_lib.destroy_plan("d", p)
    # in-place r2c with padded rows, 2-d transforms, batch of 10 -> 5 chunks
    n0, n1, hm = 6, 10, 10
    buf = np.zeros((hm, n0, 2 * (n1 // 2 + 1)), dtype=np.float32)
    xr = F.rand_real(rng, (hm, n0, n1), "f")
    buf[:, :, :n1] = xr
    p = emu_lib.plan_many_dft_r2c("f", [n0, n1], hm, buf.ctypes.data, None, 1, n0 * 2 * (n1 // 2 + 1),
                                  buf.ctypes.data, None, 1, n0 * (n1 // 2 + 1), B.FFTW_ESTIMATE)
    assert "5 chunks pipelined" in emu_lib.sprint_plan("f", p)
    emu_lib.execute("f", p)
    emu_lib.destroy_plan("f", p)
    got = buf.view(np.complex64).reshape(hm, n0, n1 // 2 + 1)
    assert O.rel_l2(got, np.fft.rfftn(xr.astype(np.float64), axes=(1, 2))) < 1e-6
    # out-of-place r2c / c2r of contiguous lines (one pointer per real array: no re / im slack)
    n, hm = 512, 16
    xr = F.rand_real(rng, (hm, n), "d")
    X = np.zeros((hm, n // 2 + 1), dtype=np.complex128)
    p = emu_lib.plan_many_dft_r2c("d", [n], hm, xr.ctypes.data, None, 1, n, X.ctypes.data, None, 1, n // 2 + 1, B.FFTW_ESTIMATE)
    assert "8 chunks pipelined" in emu_lib.sprint_plan("d", p)
    emu_lib.execute("d", p)
    emu_lib.destroy_plan("d", p)
    assert O.rel_l2(X, np.fft.rfft(xr, axis=1)) < 1e-15
    back = np.zeros_like(xr)
    p = emu_lib.plan_many_dft_c2r("d", [n], hm, X.ctypes.data, None, 1, n // 2 + 1, back.ctypes.data, None, 1, n, B.FFTW_ESTIMATE)
    assert "8 chunks pipelined" in emu_lib.sprint_plan("d", p)
    emu_lib.execute("d", p)
    emu_lib.destroy_plan("d", p)
    assert O.rel_l2(back / n, xr) < 1e-15
    # the batch index is the fastest one in memory: chunks would interleave, so the plan is not cut
    z = F.rand_complex(rng, (64, 16), "d")          # element j of transform b at z[j][b]
    z0 = z.copy()
    p = emu_lib.plan_many_dft("d", [64], 16, z.ctypes.data, None, 16, 1, z.ctypes.data, None, 16, 1, -1, B.FFTW_ESTIMATE)
    assert "pipelined" not in emu_lib.sprint_plan("d", p)
    emu_lib.execute("d", p)
    emu_lib.destroy_plan("d", p)
    assert O.rel_l2(z, np.fft.fft(z0, axis=0)) < 1e-15
    # switched off
    monkeypatch.setenv("FFTW3_B200_PIPELINE", "0")
    x = F.rand_complex(rng, (24, 96), "d")
    p = emu_lib.plan_many_dft("d", [96], 24, x.ctypes.data, None, 1, 96, x.ctypes.data, None, 1, 96, -1, B.FFTW_ESTIMATE)
    assert "pipelined" not in emu_lib.sprint_plan("d", p)
    emu_lib.destroy_plan("d", p)
