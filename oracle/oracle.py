"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (checker, never the product path).

ctypes front-end to liboracle_dft.so (oracle_dft.c: our long-double CPU
restatement of the transforms the reference defines) and to the reference
itself compiled codelet-less into oracle/_ref/ (see oracle/Makefile).

Only tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

R2R_KINDS = {
    "R2HC": 0, "HC2R": 1, "DHT": 2,
    "REDFT00": 3, "REDFT01": 4, "REDFT10": 5, "REDFT11": 6,
    "RODFT00": 7, "RODFT01": 8, "RODFT10": 9, "RODFT11": 10,
}


def build():
    """Compile the oracle (and, where /root/reference exists, oracle/_ref)."""
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle_dft.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", HERE, "ours"], check=True)
        _LIB = C.CDLL(path)
        LLP = C.POINTER(C.c_longlong)
        VP = C.c_void_p
        _LIB.oracle_dft.argtypes = [C.c_int, LLP, C.c_longlong, VP, C.c_int]
        _LIB.oracle_r2c.argtypes = [C.c_int, LLP, C.c_longlong, VP, VP]
        _LIB.oracle_c2r.argtypes = [C.c_int, LLP, C.c_longlong, VP, VP]
        _LIB.oracle_r2r.argtypes = [C.c_int, LLP, C.POINTER(C.c_int), C.c_longlong, VP]
        _LIB.oracle_r2r_direct_1d.argtypes = [C.c_int, C.c_longlong, VP, VP]
        _LIB.oracle_dft_direct_1d.argtypes = [C.c_longlong, VP, VP, C.c_int]
    return _LIB


def _dims(shape):
    arr = (C.c_longlong * len(shape))(*[int(s) for s in shape])
    return arr


def _cplx_to_ld(x):
    """complex ndarray -> contiguous longdouble array [..., 2]"""
    x = np.asarray(x)
    out = np.empty(x.shape + (2,), dtype=np.longdouble)
    out[..., 0] = x.real
    out[..., 1] = x.imag
    return np.ascontiguousarray(out)


def _ld_to_cplx(a):
    return a[..., 0].astype(np.clongdouble) + 1j * a[..., 1].astype(np.clongdouble)


def dft(x, sign=-1, rank=None):
    """Unnormalised complex DFT over the last `rank` axes (default: all axes);
    leading axes are batch.  Returns clongdouble."""
    x = np.asarray(x)
    rank = x.ndim if rank is None else rank
    shape = x.shape[x.ndim - rank:]
    howmany = int(np.prod(x.shape[:x.ndim - rank], dtype=np.int64)) if x.ndim > rank else 1
    buf = _cplx_to_ld(x)
    if buf.size:
        rc = _lib().oracle_dft(rank, _dims(shape), howmany, buf.ctypes.data, int(sign))
        assert rc == 0
    return _ld_to_cplx(buf)


def dft_direct_1d(x, sign=-1):
    x = np.asarray(x)
    buf = _cplx_to_ld(x)
    out = np.empty_like(buf)
    _lib().oracle_dft_direct_1d(x.shape[0], buf.ctypes.data, out.ctypes.data, int(sign))
    return _ld_to_cplx(out)


def r2c(x, rank=None):
    """Real -> half complex (last transformed axis n -> n//2+1), forward."""
    x = np.ascontiguousarray(np.asarray(x), dtype=np.longdouble)
    rank = x.ndim if rank is None else rank
    shape = x.shape[x.ndim - rank:]
    batch = x.shape[:x.ndim - rank]
    howmany = int(np.prod(batch, dtype=np.int64)) if batch else 1
    out = np.empty(batch + shape[:-1] + (shape[-1] // 2 + 1, 2), dtype=np.longdouble)
    rc = _lib().oracle_r2c(rank, _dims(shape), howmany, x.ctypes.data, out.ctypes.data)
    assert rc == 0
    return _ld_to_cplx(out)


def c2r(X, n_last, rank=None):
    """Half complex -> real of last-axis length n_last, backward, unnormalised."""
    X = np.asarray(X)
    rank = X.ndim if rank is None else rank
    cshape = X.shape[X.ndim - rank:]
    assert cshape[-1] == n_last // 2 + 1
    shape = cshape[:-1] + (n_last,)
    batch = X.shape[:X.ndim - rank]
    howmany = int(np.prod(batch, dtype=np.int64)) if batch else 1
    buf = _cplx_to_ld(X)
    out = np.empty(batch + shape, dtype=np.longdouble)
    rc = _lib().oracle_c2r(rank, _dims(shape), howmany, buf.ctypes.data, out.ctypes.data)
    assert rc == 0
    return out


def r2r(x, kinds, rank=None):
    """Separable r2r of the last `rank` axes; kinds: list of names or ints."""
    x = np.array(x, dtype=np.longdouble, order="C", copy=True)
    rank = x.ndim if rank is None else rank
    shape = x.shape[x.ndim - rank:]
    howmany = int(np.prod(x.shape[:x.ndim - rank], dtype=np.int64)) if x.ndim > rank else 1
    ks = [R2R_KINDS[k] if isinstance(k, str) else int(k) for k in kinds]
    assert len(ks) == rank
    rc = _lib().oracle_r2r(rank, _dims(shape), (C.c_int * rank)(*ks), howmany, x.ctypes.data)
    assert rc == 0
    return x


def r2r_direct_1d(x, kind):
    x = np.ascontiguousarray(np.asarray(x), dtype=np.longdouble)
    y = np.empty_like(x)
    k = R2R_KINDS[kind] if isinstance(kind, str) else int(kind)
    rc = _lib().oracle_r2r_direct_1d(k, x.shape[0], x.ctypes.data, y.ctypes.data)
    assert rc == 0
    return y


def rel_l2(a, b):
    """relative L2 error ||a-b|| / ||b|| evaluated in long double"""
    a = np.asarray(a).astype(np.clongdouble).ravel()
    b = np.asarray(b).astype(np.clongdouble).ravel()
    den = np.sqrt(np.sum(np.abs(b) ** 2))
    num = np.sqrt(np.sum(np.abs(a - b) ** 2))
    if den == 0:
        return float(num)
    return float(num / den)


# --------------------------------------------------------------------------
# The reference itself (codelet-less build under oracle/_ref/), via ctypes.
# --------------------------------------------------------------------------
FFTW_ESTIMATE = 1 << 6
FFTW_ALLOW_LARGE_GENERIC = 1 << 13


class RefFFTW:
    """Minimal ctypes binding to oracle/_ref/libfftw3{,f,l}_ref.so."""

    def __init__(self, prec="d"):
        sfx = {"d": "", "f": "f", "l": "l"}[prec]
        self.pfx = "fftw" + sfx + "_"
        path = os.path.join(HERE, "_ref", "libfftw3%s_ref.so" % sfx)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.rdtype = {"d": np.float64, "f": np.float32, "l": np.longdouble}[prec]
        self.cdtype = {"d": np.complex128, "f": np.complex64, "l": np.clongdouble}[prec]
        f = self._f
        f("plan_many_dft").restype = C.c_void_p
        f("plan_many_dft").argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int,
                                       C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int,
                                       C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int,
                                       C.c_int, C.c_uint]
        f("plan_many_r2r").restype = C.c_void_p
        f("plan_many_r2r").argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int,
                                       C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int,
                                       C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int,
                                       C.POINTER(C.c_int), C.c_uint]
        for name in ("plan_many_dft_r2c", "plan_many_dft_c2r"):
            f(name).restype = C.c_void_p
            f(name).argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int,
                                C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int,
                                C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_uint]
        f("execute").argtypes = [C.c_void_p]
        f("destroy_plan").argtypes = [C.c_void_p]

    def _f(self, name):
        return getattr(self.lib, self.pfx + name)

    def dft(self, x, sign=-1, rank=None, flags=FFTW_ESTIMATE | FFTW_ALLOW_LARGE_GENERIC):
        x = np.ascontiguousarray(x, dtype=self.cdtype)
        rank = x.ndim if rank is None else rank
        shape = x.shape[x.ndim - rank:]
        howmany = int(np.prod(x.shape[:x.ndim - rank], dtype=np.int64)) if x.ndim > rank else 1
        dist = int(np.prod(shape, dtype=np.int64))
        out = np.empty_like(x)
        n = (C.c_int * rank)(*shape)
        p = self._f("plan_many_dft")(rank, n, howmany, x.ctypes.data, None, 1, dist,
                                     out.ctypes.data, None, 1, dist, int(sign), flags)
        if not p:
            return None
        self._f("execute")(p)
        self._f("destroy_plan")(p)
        return out

    def r2c(self, x, rank=None, flags=FFTW_ESTIMATE | FFTW_ALLOW_LARGE_GENERIC):
        x = np.ascontiguousarray(x, dtype=self.rdtype)
        rank = x.ndim if rank is None else rank
        shape = x.shape[x.ndim - rank:]
        batch = x.shape[:x.ndim - rank]
        howmany = int(np.prod(batch, dtype=np.int64)) if batch else 1
        cshape = shape[:-1] + (shape[-1] // 2 + 1,)
        out = np.empty(batch + cshape, dtype=self.cdtype)
        n = (C.c_int * rank)(*shape)
        p = self._f("plan_many_dft_r2c")(rank, n, howmany, x.ctypes.data, None, 1,
                                         int(np.prod(shape)), out.ctypes.data, None, 1,
                                         int(np.prod(cshape)), flags)
        if not p:
            return None
        self._f("execute")(p)
        self._f("destroy_plan")(p)
        return out

    def c2r(self, X, n_last, rank=None, flags=FFTW_ESTIMATE | FFTW_ALLOW_LARGE_GENERIC):
        X = np.array(X, dtype=self.cdtype, order="C", copy=True)   # c2r destroys input
        rank = X.ndim if rank is None else rank
        cshape = X.shape[X.ndim - rank:]
        shape = cshape[:-1] + (n_last,)
        batch = X.shape[:X.ndim - rank]
        howmany = int(np.prod(batch, dtype=np.int64)) if batch else 1
        out = np.empty(batch + shape, dtype=self.rdtype)
        n = (C.c_int * rank)(*shape)
        p = self._f("plan_many_dft_c2r")(rank, n, howmany, X.ctypes.data, None, 1,
                                         int(np.prod(cshape)), out.ctypes.data, None, 1,
                                         int(np.prod(shape)), flags)
        if not p:
            return None
        self._f("execute")(p)
        self._f("destroy_plan")(p)
        return out

    def r2r(self, x, kinds, rank=None, flags=FFTW_ESTIMATE | FFTW_ALLOW_LARGE_GENERIC):
        x = np.ascontiguousarray(x, dtype=self.rdtype)
        rank = x.ndim if rank is None else rank
        shape = x.shape[x.ndim - rank:]
        howmany = int(np.prod(x.shape[:x.ndim - rank], dtype=np.int64)) if x.ndim > rank else 1
        dist = int(np.prod(shape, dtype=np.int64))
        ks = [R2R_KINDS[k] if isinstance(k, str) else int(k) for k in kinds]
        out = np.empty_like(x)
        n = (C.c_int * rank)(*shape)
        p = self._f("plan_many_r2r")(rank, n, howmany, x.ctypes.data, None, 1, dist,
                                     out.ctypes.data, None, 1, dist,
                                     (C.c_int * rank)(*ks), flags)
        if not p:
            return None
        self._f("execute")(p)
        self._f("destroy_plan")(p)
        return out
