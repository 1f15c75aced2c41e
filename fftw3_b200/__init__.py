"""fftw3_b200 -- B200-native FFT engine behind the FFTW 3 C API.

The product is the C-ABI shared library ``fftw3_b200/lib/libfftw3_b200.so``
(host layer in C, kernels in CUDA for sm_100a; exported symbols = the
``fftw_*`` / ``fftwf_*`` API of ``include/fftw3.h``).  This Python package is
only the thin ctypes binding used by tests and ``bench.py``; it mirrors the
C API one-to-one (same names, argument meaning and NULL-on-failure errors).

There is no CPU fallback: loading works anywhere, but plan creation returns
NULL (``None`` here) when no CUDA device is usable.
"""
from .binding import (  # noqa: F401
    FFTW_BACKWARD, FFTW_DESTROY_INPUT, FFTW_ESTIMATE, FFTW_EXHAUSTIVE, FFTW_FORWARD,
    FFTW_MEASURE, FFTW_PATIENT, FFTW_PRESERVE_INPUT, FFTW_UNALIGNED, FFTW_WISDOM_ONLY,
    R2R_KINDS, Lib, build_library, default_library_path, load,
)
