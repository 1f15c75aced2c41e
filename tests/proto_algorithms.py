"""Numpy prototypes of every algorithm the CUDA kernels implement, checked
against the oracle.  Each function mirrors one kernel (or pre/post element map)
in fftw3_b200/csrc/device; keeping them here documents the index algebra and
lets `-m "not gpu"` tests validate it without a GPU."""
import numpy as np


def stockham(x, radices):
    """Stockham autosort forward FFT with the given radix sequence
    (device/fft_generic.cuh stage_phase)."""
    x = np.asarray(x, dtype=np.complex128).copy()
    n = x.shape[0]
    assert int(np.prod(radices)) == n
    ns = 1
    for R in radices:
        y = np.empty_like(x)
        for j in range(n // R):
            k = j % ns
            v = np.array([x[j + r * (n // R)] * np.exp(-2j * np.pi * r * k / (ns * R)) for r in range(R)])
            v = np.fft.fft(v)
            j0 = (j // ns) * ns * R + k
            for r in range(R):
                y[j0 + r * ns] = v[r]
        x = y
        ns *= R
    return x


def r2c_even(x):
    """r2c of even n through a complex FFT of n/2 (real_ops.cuh R2C_POST)."""
    n = x.shape[0]; m = n // 2
    z = x[0::2] + 1j * x[1::2]
    Z = np.fft.fft(z)
    X = np.empty(m + 1, dtype=complex)
    for k in range(m + 1):
        a = Z[k % m]; b = np.conj(Z[(m - k) % m])
        w = np.exp(-2j * np.pi * k / n)
        X[k] = 0.5 * ((a + b) - 1j * w * (a - b))
    return X


def c2r_even(X, n):
    """c2r of even n: build the n/2 complex spectrum, backward FFT, unpack
    (real_ops.cuh C2R_PRE).  Unnormalised: returns n * x."""
    m = n // 2
    Z = np.empty(m, dtype=complex)
    for k in range(m):
        a = X[k]; b = np.conj(X[m - k])
        if k == 0:   # imag of DC / Nyquist ignored
            a = a.real + 0j; b = b.real + 0j
        w = np.exp(+2j * np.pi * k / n)
        Z[k] = (a + b) + 1j * w * (a - b)
    z = np.fft.ifft(Z) * m
    x = np.empty(n)
    x[0::2] = z.real; x[1::2] = z.imag
    return x


def bluestein(x, m=None):
    n = x.shape[0]
    if m is None:
        m = 1
        while m < 2 * n - 1:
            m *= 2
    j = np.arange(n)
    chirp = np.exp(-1j * np.pi * ((j * j) % (2 * n)) / n)     # exp(-pi i j^2 / n)
    a = np.zeros(m, complex); a[:n] = x * chirp
    b = np.zeros(m, complex); b[:n] = np.conj(chirp); b[m - n + 1:] = np.conj(chirp[1:][::-1])
    B = np.fft.fft(b)
    A = np.fft.fft(a) * B
    # inverse via conj(FFT(conj(.)))
    c = np.conj(np.fft.fft(np.conj(A)))
    return c[:n] * chirp / m


def four_step(x, n1, n2):
    """N = n1*n2: pass A = n2 strided FFTs of length n1 + twiddle, pass B = n1
    contiguous FFTs of length n2, transposed store."""
    n = n1 * n2
    a = x.reshape(n1, n2)                      # a[j1, j2] = x[j1*n2 + j2]
    ya = np.fft.fft(a, axis=0)                 # ya[k1, j2]
    k1 = np.arange(n1)[:, None]; j2 = np.arange(n2)[None, :]
    ya = ya * np.exp(-2j * np.pi * k1 * j2 / n)
    yb = np.fft.fft(ya, axis=1)                # yb[k1, k2]
    return yb.T.reshape(n)                     # X[k1 + n1*k2]


# ---- r2r kinds through one complex FFT of length M: (pre, M, post) ----
def r2r_1d(x, kind):
    x = np.asarray(x, dtype=float); n = x.shape[0]
    j = np.arange(n)
    if kind == "R2HC":
        Z = np.fft.fft(x.astype(complex))
        y = np.empty(n); y[: n // 2 + 1] = Z[: n // 2 + 1].real
        for k in range(1, (n + 1) // 2):
            y[n - k] = Z[k].imag
        return y
    if kind == "HC2R":
        z = np.zeros(n, complex); z[0] = x[0]
        for k in range(1, (n + 1) // 2):
            z[k] = x[k] + 1j * x[n - k]; z[n - k] = np.conj(z[k])
        if n % 2 == 0:
            z[n // 2] = x[n // 2]
        return np.fft.fft(np.conj(z)).real
    if kind == "DHT":
        Z = np.fft.fft(x.astype(complex)); return Z.real - Z.imag
    if kind == "REDFT00":
        M = 2 * (n - 1); z = np.zeros(M, complex); z[:n] = x; z[n:] = x[1:n - 1][::-1]
        return np.fft.fft(z).real[:n]
    if kind == "RODFT00":
        M = 2 * (n + 1); z = np.zeros(M, complex); z[1:n + 1] = x; z[n + 2:] = -x[::-1]
        return -np.fft.fft(z).imag[1:n + 1]
    if kind in ("REDFT10", "RODFT10"):
        xx = x if kind == "REDFT10" else x * (-1.0) ** j
        v = np.empty(n, complex)
        h = (n + 1) // 2
        v[:h] = xx[0::2]; v[h:] = xx[1::2][::-1]
        V = np.fft.fft(v)
        y = 2 * (np.exp(-1j * np.pi * j / (2 * n)) * V).real
        return y if kind == "REDFT10" else y[::-1]
    if kind in ("REDFT01", "RODFT01"):
        X = x if kind == "REDFT01" else x[::-1]
        Xn = np.concatenate([X, [0.0]])
        z = np.exp(-1j * np.pi * j / (2 * n)) * (Xn[j] + 1j * Xn[n - j])
        Z = np.fft.fft(z).real
        y = np.empty(n); h = (n + 1) // 2
        y[0::2] = Z[:h]; y[1::2] = Z[h:][::-1]
        return y if kind == "REDFT01" else y * (-1.0) ** j
    if kind in ("REDFT11", "RODFT11"):
        M = 2 * n
        a = np.zeros(M, complex); a[:n] = x * np.exp(-1j * np.pi * j / (2 * n))
        A = np.fft.fft(a)[:n]
        t = np.exp(-1j * np.pi * (2 * j + 1) / (4 * n)) * A
        return 2 * t.real if kind == "REDFT11" else -2 * t.imag
    raise ValueError(kind)
