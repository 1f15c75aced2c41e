"""Run the committed golden vectors (tests/golden/reference_vectors.npz, produced
by the reference's own code, see tests/golden/make_golden.py) through a library
or through the oracle."""
import os

import numpy as np

from fftw3_b200 import binding as B
from oracle import oracle as O

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz")


def cases():
    z = np.load(PATH)
    names = sorted({k[:-4] for k in z.files if k.endswith("__in")})
    return [(n, z[n + "__in"], z[n + "__out"]) for n in names]


def parse(name):
    """-> (family, kinds/sign info, rank)"""
    parts = name.split("_")
    fam = parts[0]
    if fam == "c2c":
        dims = parts[1].split("x")
        return fam, (-1 if parts[2] == "fwd" else +1), len(dims)
    if fam in ("r2c", "c2r"):
        return fam, None, len(parts[1].split("x"))
    kinds = parts[1].split("x")
    return fam, kinds, len(kinds)


def via_oracle(name, x):
    fam, info, rank = parse(name)
    if fam == "c2c":
        return O.dft(x, sign=info, rank=rank)
    if fam == "r2c":
        return O.r2c(x, rank=rank)
    if fam == "c2r":
        n_last = int(name.split("_")[1].split("x")[-1])
        return O.c2r(x, n_last, rank=rank)
    return O.r2r(x, info, rank=rank)


def via_lib(lib, name, x):
    fam, info, rank = parse(name)
    x = np.ascontiguousarray(x)
    howmany = int(np.prod(x.shape[:x.ndim - rank]))
    if fam == "c2c":
        shape = x.shape[x.ndim - rank:]
        y = np.zeros_like(x)
        d = int(np.prod(shape))
        p = lib.plan_many_dft("d", shape, howmany, x.ctypes.data, None, 1, d, y.ctypes.data, None, 1, d, info,
                              B.FFTW_ESTIMATE)
    elif fam == "r2c":
        shape = x.shape[x.ndim - rank:]
        cshape = shape[:-1] + (shape[-1] // 2 + 1,)
        y = np.zeros(x.shape[:x.ndim - rank] + cshape, np.complex128)
        p = lib.plan_many_dft_r2c("d", shape, howmany, x.ctypes.data, None, 1, int(np.prod(shape)), y.ctypes.data,
                                  None, 1, int(np.prod(cshape)), B.FFTW_ESTIMATE)
    elif fam == "c2r":
        n_last = int(name.split("_")[1].split("x")[-1])
        cshape = x.shape[x.ndim - rank:]
        shape = cshape[:-1] + (n_last,)
        x = x.copy()
        y = np.zeros(x.shape[:x.ndim - rank] + shape, np.float64)
        p = lib.plan_many_dft_c2r("d", shape, howmany, x.ctypes.data, None, 1, int(np.prod(cshape)), y.ctypes.data,
                                  None, 1, int(np.prod(shape)), B.FFTW_ESTIMATE)
    else:
        shape = x.shape[x.ndim - rank:]
        y = np.zeros_like(x)
        d = int(np.prod(shape))
        p = lib.plan_many_r2r("d", shape, howmany, x.ctypes.data, None, 1, d, y.ctypes.data, None, 1, d, info,
                              B.FFTW_ESTIMATE)
    assert p, "NULL plan for golden case " + name
    lib.execute("d", p)
    lib.destroy_plan("d", p)
    return y
