#!/usr/bin/env python3
"""Generate tests/golden/reference_vectors.npz from the REFERENCE ITSELF.

The reference (FFTW 3.3.11 sources under /root/reference) holds no stored golden
vectors, so these are produced by running its own code: the
codelet-less build oracle/_ref/libfftw3_ref.so made by oracle/Makefile from the
unmodified sources.  Inputs are seeded uniform [-0.5, 0.5) like the reference's
verifier (libbench2/verify-lib.c:64-67).  Even-size real transforms cannot be
planned by a codelet-less build (no size-2 real leaf), so the real-data cases
here use odd sizes; even sizes are pinned through the c2c cases plus the
oracle's direct-definition cross checks.

Run in the build container (needs oracle/_ref):  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

CASES_C2C = [  # (name, shape incl. batch, rank, sign)
    ("c2c_16_fwd", (4, 16), 1, -1), ("c2c_30_bwd", (3, 30), 1, +1), ("c2c_64_fwd", (5, 64), 1, -1),
    ("c2c_100_fwd", (2, 100), 1, -1), ("c2c_1024_fwd", (2, 1024), 1, -1), ("c2c_1009_fwd", (2, 1009), 1, -1),
    ("c2c_17_bwd", (3, 17), 1, +1), ("c2c_8x6_fwd", (2, 8, 6), 2, -1), ("c2c_5x7_bwd", (2, 5, 7), 2, +1),
    ("c2c_8x8x8_fwd", (1, 8, 8, 8), 3, -1), ("c2c_4x6x10_bwd", (2, 4, 6, 10), 3, +1),
]
CASES_R2C = [("r2c_9", (3, 9), 1), ("r2c_15", (2, 15), 1), ("r2c_3x5", (2, 3, 5), 2), ("r2c_4x3x7", (1, 4, 3, 7), 3),
             ("r2c_35", (2, 35), 1)]
R2R_ODD = [("R2HC", 9), ("HC2R", 9), ("DHT", 15), ("REDFT00", 9), ("REDFT01", 9), ("REDFT10", 15), ("REDFT11", 9),
           ("RODFT00", 7), ("RODFT01", 9), ("RODFT10", 15), ("RODFT11", 9)]


def main():
    ref = O.RefFFTW("d")
    rng = np.random.default_rng(20261017)
    out = {}
    for name, shape, rank, sign in CASES_C2C:
        x = (rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)).astype(np.complex128)
        y = ref.dft(x, sign=sign, rank=rank)
        assert y is not None, name
        out[name + "__in"], out[name + "__out"] = x, y
    for name, shape, rank in CASES_R2C:
        x = rng.uniform(-0.5, 0.5, shape).astype(np.float64)
        y = ref.r2c(x, rank=rank)
        assert y is not None, name
        out[name + "__in"], out[name + "__out"] = x, y
        z = ref.c2r(y, shape[-1], rank=rank)          # back through the reference's c2r
        assert z is not None, name
        out[name.replace("r2c", "c2r") + "__in"], out[name.replace("r2c", "c2r") + "__out"] = y, z
    for kind, n in R2R_ODD:
        x = rng.uniform(-0.5, 0.5, (3, n)).astype(np.float64)
        y = ref.r2r(x, [kind], rank=1)
        if y is None:
            print("reference cannot plan", kind, n, "- skipped")
            continue
        out["r2r_%s_%d__in" % (kind, n)], out["r2r_%s_%d__out" % (kind, n)] = x, y
    x = rng.uniform(-0.5, 0.5, (2, 5, 7)).astype(np.float64)
    y = ref.r2r(x, ["REDFT10", "RODFT01"], rank=2)
    if y is not None:
        out["r2r_REDFT10xRODFT01_5x7__in"], out["r2r_REDFT10xRODFT01_5x7__out"] = x, y
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "with", len(out) // 2, "cases,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
