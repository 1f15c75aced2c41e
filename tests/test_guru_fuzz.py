"""Seeded fuzz of the guru interface on the emulated device layer: random ranks, sizes, padded and
permuted strides, batches, in/out of place, interleaved and split complex -- compared with
numpy's FFT (itself checked against the oracle elsewhere).  The reference's own test strategy
does the same with random problem strings (tests/check.pl:27-120: random rank, sizes, vector
ranks, in-place flags); here the strides are randomised too, which its bench cannot express."""
import numpy as np
import pytest

from fftw3_b200 import binding as B

SIZES = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 16, 17, 20, 24, 31, 32, 36]


def _layout(rng, dims_n, pad):
    """Random layout of an array with logical shape dims_n: a random axis order with padding between
    the axes; returns (strides in elements, total elements)."""
    order = list(rng.permutation(len(dims_n)))
    strides = [0] * len(dims_n)
    acc = 1
    for ax in order:
        strides[ax] = acc
        acc *= dims_n[ax]
        if pad:
            acc += int(rng.integers(0, 3))
    return strides, max(acc, 1)


def _gather(buf, shape, strides):
    idx = np.zeros(shape, dtype=np.int64)
    for ax, (n, s) in enumerate(zip(shape, strides)):
        sh = [1] * len(shape)
        sh[ax] = n
        idx = idx + (np.arange(n, dtype=np.int64) * s).reshape(sh)
    return buf[idx], idx


@pytest.mark.parametrize("seed", range(120))
def test_random_guru_c2c(host_lib, seed):
    rng = np.random.default_rng(1000 + seed)
    rank = int(rng.integers(0, 4))
    hrank = int(rng.integers(0, 3))
    shape = [int(rng.choice(SIZES)) for _ in range(rank + hrank)]
    while int(np.prod(shape, dtype=np.int64)) > 20000:
        shape[int(np.argmax(shape))] = 2
    inplace = bool(rng.integers(0, 2))
    split = bool(rng.integers(0, 2))
    sign = int(rng.choice([-1, 1]))
    pad = bool(rng.integers(0, 2))
    is_, isz = _layout(rng, shape, pad)
    if inplace:
        os_, osz = is_, isz
    else:
        os_, osz = _layout(rng, shape, pad)
    x = rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)
    dims = [(shape[i], is_[i], os_[i]) for i in range(rank)]
    hows = [(shape[rank + i], is_[rank + i], os_[rank + i]) for i in range(hrank)]
    want = np.fft.fftn(x, axes=tuple(range(rank))) if sign < 0 else np.fft.ifftn(x, axes=tuple(range(rank))) * np.prod(
        shape[:rank] or [1])
    if rank == 0:
        want = x.copy()
    _, iidx = _gather(np.zeros(isz), shape, is_) if shape else (None, np.zeros((), dtype=np.int64))
    _, oidx = _gather(np.zeros(osz), shape, os_) if shape else (None, np.zeros((), dtype=np.int64))
    if split:
        ri, ii = np.full(isz, 7.0), np.full(isz, 7.0)
        ri[iidx], ii[iidx] = x.real, x.imag
        ro, io = (ri, ii) if inplace else (np.full(osz, 9.0), np.full(osz, 9.0))
        if sign < 0:
            p = host_lib.plan_guru_split_dft("d", dims, hows, ri.ctypes.data, ii.ctypes.data, ro.ctypes.data, io.ctypes.data,
                                            B.FFTW_ESTIMATE)
        else:       # backward = forward with re/im exchanged (api/plan-guru-split-dft.h:30-31)
            p = host_lib.plan_guru_split_dft("d", dims, hows, ii.ctypes.data, ri.ctypes.data, io.ctypes.data, ro.ctypes.data,
                                            B.FFTW_ESTIMATE)
        assert p, (shape, dims, hows)
        host_lib.execute("d", p)
        host_lib.destroy_plan("d", p)
        got = ro[oidx] + 1j * io[oidx]
    else:
        a = np.full(isz, 7 + 7j)
        a[iidx] = x
        b = a if inplace else np.full(osz, 9 + 9j)
        p = host_lib.plan_guru_dft("d", dims, hows, a.ctypes.data, b.ctypes.data, sign, B.FFTW_ESTIMATE)
        assert p, (shape, dims, hows)
        host_lib.execute("d", p)
        host_lib.destroy_plan("d", p)
        got = b[oidx]
        if not inplace:         # nothing outside the output tensor may be written, and the input is preserved
            mask = np.ones(osz, dtype=bool)
            mask[oidx.reshape(-1)] = False
            assert np.all(b[mask] == 9 + 9j)
            assert np.array_equal(a[iidx], x) and np.all(np.delete(a, iidx.reshape(-1)) == 7 + 7j)
    scale = max(1.0, float(np.abs(want).max()))
    assert np.abs(got - want).max() <= 1e-12 * scale, (seed, shape, rank, hrank, inplace, split, sign, pad)


@pytest.mark.parametrize("seed", range(60))
def test_random_guru_r2r(host_lib, seed):
    """random r2r kinds / strided batches (api/plan-guru-r2r.h), against the oracle"""
    import ctypes as C
    from oracle import oracle as O
    rng = np.random.default_rng(5000 + seed)
    rank = int(rng.integers(1, 3))
    hrank = int(rng.integers(0, 3))
    shape = [int(rng.choice([2, 3, 4, 5, 6, 8, 9, 12, 16, 17])) for _ in range(rank + hrank)]
    kinds = [str(rng.choice(sorted(B.R2R_KINDS))) for _ in range(rank)]
    for i, k in enumerate(kinds):
        if k == "REDFT00" and shape[i] < 2:
            shape[i] = 2
    inplace = bool(rng.integers(0, 2))
    pad = bool(rng.integers(0, 2))
    is_, isz = _layout(rng, shape, pad)
    os_, osz = (is_, isz) if inplace else _layout(rng, shape, pad)
    x = rng.uniform(-0.5, 0.5, shape)
    _, iidx = _gather(np.zeros(isz), shape, is_)
    _, oidx = _gather(np.zeros(osz), shape, os_)
    a = np.full(isz, 7.0)
    a[iidx] = x
    b = a if inplace else np.full(osz, 9.0)
    dims = (B.Iodim * rank)(*[B.Iodim(shape[i], is_[i], os_[i]) for i in range(rank)])
    hows = (B.Iodim * max(hrank, 1))(*[B.Iodim(shape[rank + i], is_[rank + i], os_[rank + i]) for i in range(hrank)])
    ks = (C.c_int * rank)(*[B.R2R_KINDS[k] for k in kinds])
    p = host_lib.fn("d", "plan_guru_r2r")(rank, C.cast(dims, C.c_void_p), hrank, C.cast(hows, C.c_void_p), a.ctypes.data,
                                         b.ctypes.data, ks, B.FFTW_ESTIMATE)
    assert p, (shape, kinds)
    host_lib.execute("d", p)
    host_lib.destroy_plan("d", p)
    want = x
    for ax in range(rank):      # separable: kind[ax] along axis ax
        moved = np.moveaxis(want, ax, -1)
        flat = np.ascontiguousarray(moved).reshape(-1, shape[ax])
        res = np.stack([O.r2r(row.copy(), [kinds[ax]], rank=1) for row in flat]).reshape(moved.shape)
        want = np.moveaxis(res, -1, ax)
    got = b[oidx]
    scale = max(1.0, float(np.abs(want).max()))
    assert np.abs(got - want).max() <= 1e-11 * scale, (seed, shape, kinds, inplace, pad)
    if not inplace:             # FFTW_PRESERVE_INPUT is the default for r2r, except HC2R (doc/reference.texi:511-527)
        if "HC2R" not in kinds:
            assert np.array_equal(a[iidx], x), (seed, shape, kinds)
        mask = np.ones(osz, dtype=bool)
        mask[oidx.reshape(-1)] = False
        assert np.all(b[mask] == 9.0)


@pytest.mark.parametrize("seed", range(40))
def test_random_guru_r2c_c2r(host_lib, seed):
    """random out-of-place r2c and c2r through the guru interface (api/plan-guru-dft-r2c.h,
    plan-guru-dft-c2r.h): real strides in reals, complex strides in complex elements."""
    import ctypes as C
    rng = np.random.default_rng(9000 + seed)
    rank = int(rng.integers(1, 4))
    hrank = int(rng.integers(0, 3))
    shape = [int(rng.choice([2, 3, 4, 5, 6, 7, 8, 9, 12, 15, 16, 17])) for _ in range(rank + hrank)]
    pad = bool(rng.integers(0, 2))
    cshape = list(shape)
    cshape[rank - 1] = shape[rank - 1] // 2 + 1
    rs, rsz = _layout(rng, shape, pad)
    cs, csz = _layout(rng, cshape, pad)
    x = rng.uniform(-0.5, 0.5, shape)
    _, ridx = _gather(np.zeros(rsz), shape, rs)
    _, cidx = _gather(np.zeros(csz), cshape, cs)
    a = np.full(rsz, 7.0)
    a[ridx] = x
    b = np.full(csz, 9 + 9j)
    dims = (B.Iodim * rank)(*[B.Iodim(shape[i], rs[i], cs[i]) for i in range(rank)])
    hows = (B.Iodim * max(hrank, 1))(*[B.Iodim(shape[rank + i], rs[rank + i], cs[rank + i]) for i in range(hrank)])
    p = host_lib.fn("d", "plan_guru_dft_r2c")(rank, C.cast(dims, C.c_void_p), hrank, C.cast(hows, C.c_void_p),
                                             a.ctypes.data, b.ctypes.data, B.FFTW_ESTIMATE)
    assert p, (shape, rank, hrank)
    host_lib.execute("d", p)
    host_lib.destroy_plan("d", p)
    want = np.fft.rfftn(x, axes=tuple(range(rank)))
    scale = max(1.0, float(np.abs(want).max()))
    assert np.abs(b[cidx] - want).max() <= 1e-12 * scale, (seed, "r2c", shape, rank, hrank)
    assert np.array_equal(a[ridx], x), (seed, "r2c must preserve its input", shape)
    mask = np.ones(csz, dtype=bool)
    mask[cidx.reshape(-1)] = False
    assert np.all(b[mask] == 9 + 9j)
    # c2r of that spectrum (input may be destroyed: work on a copy), unnormalised
    spec = np.full(csz, 3 + 3j)
    spec[cidx] = want
    back = np.full(rsz, 5.0)
    dims = (B.Iodim * rank)(*[B.Iodim(shape[i], cs[i], rs[i]) for i in range(rank)])
    hows = (B.Iodim * max(hrank, 1))(*[B.Iodim(shape[rank + i], cs[rank + i], rs[rank + i]) for i in range(hrank)])
    p = host_lib.fn("d", "plan_guru_dft_c2r")(rank, C.cast(dims, C.c_void_p), hrank, C.cast(hows, C.c_void_p),
                                             spec.ctypes.data, back.ctypes.data, B.FFTW_ESTIMATE)
    assert p, (shape, rank, hrank, "c2r")
    host_lib.execute("d", p)
    host_lib.destroy_plan("d", p)
    n = float(np.prod(shape[:rank]))
    assert np.abs(back[ridx] / n - x).max() <= 1e-12, (seed, "c2r", shape, rank, hrank)


@pytest.mark.parametrize("seed", range(40))
def test_random_guru_c2c_single_precision(host_lib, seed):
    """the same random strided problems in single precision (fftwf_ entry points)"""
    rng = np.random.default_rng(3000 + seed)
    rank = int(rng.integers(1, 4))
    hrank = int(rng.integers(0, 3))
    shape = [int(rng.choice(SIZES)) for _ in range(rank + hrank)]
    while int(np.prod(shape, dtype=np.int64)) > 20000:
        shape[int(np.argmax(shape))] = 2
    inplace = bool(rng.integers(0, 2))
    sign = int(rng.choice([-1, 1]))
    pad = bool(rng.integers(0, 2))
    is_, isz = _layout(rng, shape, pad)
    os_, osz = (is_, isz) if inplace else _layout(rng, shape, pad)
    x = (rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)).astype(np.complex64)
    dims = [(shape[i], is_[i], os_[i]) for i in range(rank)]
    hows = [(shape[rank + i], is_[rank + i], os_[rank + i]) for i in range(hrank)]
    x64 = x.astype(np.complex128)
    want = np.fft.fftn(x64, axes=tuple(range(rank))) if sign < 0 else np.fft.ifftn(x64, axes=tuple(range(rank))) * np.prod(
        shape[:rank])
    _, iidx = _gather(np.zeros(isz), shape, is_)
    _, oidx = _gather(np.zeros(osz), shape, os_)
    a = np.full(isz, 7 + 7j, dtype=np.complex64)
    a[iidx] = x
    b = a if inplace else np.full(osz, 9 + 9j, dtype=np.complex64)
    p = host_lib.plan_guru_dft("f", dims, hows, a.ctypes.data, b.ctypes.data, sign, B.FFTW_ESTIMATE)
    assert p, (shape, dims, hows)
    host_lib.execute("f", p)
    host_lib.destroy_plan("f", p)
    got = b[oidx].astype(np.complex128)
    n = max(2.0, float(np.prod(shape[:rank])))
    assert np.linalg.norm(got - want) <= 4 * 1.2e-7 * np.log2(n) * max(np.linalg.norm(want), 1e-30), (seed, shape, rank, hrank)
    if not inplace:
        mask = np.ones(osz, dtype=bool)
        mask[oidx.reshape(-1)] = False
        assert np.all(b[mask] == np.complex64(9 + 9j))


@pytest.mark.parametrize("seed", range(0, 120, 4))
def test_random_guru_problems_through_the_host_pipeline(emu_lib, seed, monkeypatch):
    """The same seeded guru problems with every batched host-array problem forced through the chunk pipeline of
    csrc/host/exec.c (FFTW3_B200_PIPE_MIN_KB=1): whatever layout the fuzz draws, cutting the outermost batch dimension
    must either be refused (chunks would interleave) or give the same result."""
    monkeypatch.setenv("FFTW3_B200_PIPE_MIN_KB", "1")
    test_random_guru_c2c(emu_lib, seed)
    if seed < 60:
        test_random_guru_r2r(emu_lib, seed)
    if seed < 40:
        test_random_guru_r2c_c2r(emu_lib, seed)
        test_random_guru_c2c_single_precision(emu_lib, seed)
