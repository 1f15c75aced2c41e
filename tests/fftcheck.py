"""Shared parity checks: run a problem through the C API of a loaded library
(host numpy buffers or device buffers) and compare with the oracle.

Tolerance (the `north_star` bound): relative L2 error <= C_TOL * eps * log2(N)
with C_TOL = 1.5 (the reference itself sits near 0.1-0.5 eps log2 N, see
BASELINE.md; Bluestein sizes get a factor 3 since they run two transforms of
2-4x the length).  eps = 2^-52 (double) / 2^-23 (float).
"""
import math

import numpy as np

from fftw3_b200 import binding as B
from oracle import oracle as O

EPS = {"d": 2.0 ** -52, "f": 2.0 ** -23}
RDT = {"d": np.float64, "f": np.float32}
CDT = {"d": np.complex128, "f": np.complex64}
C_TOL = 1.5


def tol(prec, nlogical, factor=1.0):
    return C_TOL * factor * EPS[prec] * max(1.0, math.log2(max(2, nlogical)))


def smooth(n):
    for p in (2, 3, 5, 7, 11, 13):
        while n % p == 0:
            n //= p
    return n == 1


def tol_for(prec, shape):
    n = int(np.prod(shape))
    f = 1.0 if all(smooth(int(s)) for s in shape) else 4.0
    return tol(prec, n, f)


def rand_complex(rng, shape, prec):
    return (rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)).astype(CDT[prec])


def rand_real(rng, shape, prec):
    return rng.uniform(-0.5, 0.5, shape).astype(RDT[prec])


def c2c(lib, prec, shape, howmany=1, sign=-1, inplace=False, flags=B.FFTW_ESTIMATE, seed=0):
    """contiguous batched c2c through plan_many_dft; returns (err, tol)"""
    rng = np.random.default_rng(seed)
    x = rand_complex(rng, (howmany,) + tuple(shape), prec)
    x0 = x.copy()
    y = x if inplace else np.zeros_like(x)
    dist = int(np.prod(shape))
    p = lib.plan_many_dft(prec, shape, howmany, x.ctypes.data, None, 1, dist, y.ctypes.data, None, 1, dist,
                          sign, flags)
    assert p, "plan creation returned NULL for c2c %s x%d" % (shape, howmany)
    if not (flags & B.FFTW_ESTIMATE):
        x[...] = x0            # measuring planners may overwrite the arrays
    lib.execute(prec, p)
    lib.destroy_plan(prec, p)
    if not inplace:
        assert np.array_equal(x, x0), "out-of-place c2c modified its input"
    ref = O.dft(x0, sign=sign, rank=len(shape))
    return O.rel_l2(y, ref), tol_for(prec, shape)


def r2c(lib, prec, shape, howmany=1, inplace=False, flags=B.FFTW_ESTIMATE, seed=0):
    rng = np.random.default_rng(seed)
    shape = tuple(shape)
    nl = shape[-1]
    nh = nl // 2 + 1
    cshape = shape[:-1] + (nh,)
    x0 = rand_real(rng, (howmany,) + shape, prec)
    if inplace:
        buf = np.zeros((howmany,) + cshape, dtype=CDT[prec])
        rview = buf.view(RDT[prec])              # (..., 2*nh) padded rows
        rview[..., :nl] = x0
        p = lib.plan_many_dft_r2c(prec, shape, howmany, buf.ctypes.data, None, 1, 2 * int(np.prod(cshape)),
                                  buf.ctypes.data, None, 1, int(np.prod(cshape)), flags)
        assert p
        rview[..., :nl] = x0
        lib.execute(prec, p)
        y = buf
    else:
        x = x0.copy()
        y = np.zeros((howmany,) + cshape, dtype=CDT[prec])
        p = lib.plan_many_dft_r2c(prec, shape, howmany, x.ctypes.data, None, 1, int(np.prod(shape)),
                                  y.ctypes.data, None, 1, int(np.prod(cshape)), flags)
        assert p, "plan creation returned NULL for r2c %s" % (shape,)
        x[...] = x0
        lib.execute(prec, p)
        assert np.array_equal(x, x0), "r2c modified its input"
    lib.destroy_plan(prec, p)
    ref = O.r2c(x0, rank=len(shape))
    return O.rel_l2(y, ref), tol_for(prec, shape)


def c2r(lib, prec, shape, howmany=1, inplace=False, flags=B.FFTW_ESTIMATE, seed=0):
    rng = np.random.default_rng(seed)
    shape = tuple(shape)
    nl = shape[-1]
    nh = nl // 2 + 1
    cshape = shape[:-1] + (nh,)
    xr = rand_real(rng, (howmany,) + shape, prec)
    X0 = np.fft.rfftn(xr.astype(np.float64), axes=tuple(range(1, 1 + len(shape)))).astype(CDT[prec])
    ref = O.c2r(X0, nl, rank=len(shape))
    if inplace:
        buf = X0.copy()
        p = lib.plan_many_dft_c2r(prec, shape, howmany, buf.ctypes.data, None, 1, int(np.prod(cshape)),
                                  buf.ctypes.data, None, 1, 2 * int(np.prod(cshape)), flags)
        assert p
        buf[...] = X0
        lib.execute(prec, p)
        y = buf.view(RDT[prec])[..., :nl]
    else:
        X = X0.copy()
        y = np.zeros((howmany,) + shape, dtype=RDT[prec])
        p = lib.plan_many_dft_c2r(prec, shape, howmany, X.ctypes.data, None, 1, int(np.prod(cshape)),
                                  y.ctypes.data, None, 1, int(np.prod(shape)), flags)
        assert p, "plan creation returned NULL for c2r %s" % (shape,)
        X[...] = X0
        lib.execute(prec, p)
    lib.destroy_plan(prec, p)
    return O.rel_l2(y, ref), tol_for(prec, shape)


R2R_LOGICAL = {
    "R2HC": lambda n: n, "HC2R": lambda n: n, "DHT": lambda n: n,
    "REDFT00": lambda n: 2 * (n - 1), "RODFT00": lambda n: 2 * (n + 1),
    "REDFT01": lambda n: 2 * n, "REDFT10": lambda n: 2 * n, "REDFT11": lambda n: 2 * n,
    "RODFT01": lambda n: 2 * n, "RODFT10": lambda n: 2 * n, "RODFT11": lambda n: 2 * n,
}


def r2r(lib, prec, shape, kinds, howmany=1, inplace=False, flags=B.FFTW_ESTIMATE, seed=0):
    rng = np.random.default_rng(seed)
    shape = tuple(shape)
    x0 = rand_real(rng, (howmany,) + shape, prec)
    x = x0.copy()
    y = x if inplace else np.zeros_like(x)
    dist = int(np.prod(shape))
    p = lib.plan_many_r2r(prec, shape, howmany, x.ctypes.data, None, 1, dist, y.ctypes.data, None, 1, dist,
                          kinds, flags)
    assert p, "plan creation returned NULL for r2r %s %s" % (shape, kinds)
    x[...] = x0
    lib.execute(prec, p)
    lib.destroy_plan(prec, p)
    ref = O.r2r(x0, kinds, rank=len(shape))
    logical = [R2R_LOGICAL[k](n) for k, n in zip(kinds, shape)]
    # the work transforms are up to 2x the logical size and the quarter-wave
    # twiddles add two roundings
    return O.rel_l2(y, ref), tol(prec, int(np.prod(logical)), 2.0 if all(smooth(m) for m in logical) else 6.0)
