"""Decompositions above the single pass, on device-resident arrays (where the planner may time them): L2-resident pass
pairs on side streams, register-only sub-pass pairs, and the whole-plan alternatives FFTW_MEASURE chooses among
(csrc/host/planner.c: plan_alternatives -- the role of the reference planner's search over solver trees,
kernel/planner.c:518-615).  Every decomposition is pinned once through its environment knob and checked against the
oracle; then FFTW_MEASURE picks one by itself (whichever wins, the result must be right and the choice must survive a
wisdom round trip).  Runs on the emulated device layer and, with -m gpu, on the B200."""
import ctypes as C

import numpy as np
import pytest

import fftcheck as F
from fftw3_b200 import binding as B
from oracle import oracle as O


class Dev:
    """an array in device memory: emulator = host memory registered as device, GPU = a torch tensor"""

    def __init__(self, lib, arr):
        self.emu = lib.device_name().startswith("emulated")
        self.shape, self.dtype = arr.shape, arr.dtype
        if self.emu:
            lib.lib.fftw_b200_device_malloc.restype = C.c_void_p
            lib.lib.fftw_b200_device_malloc.argtypes = [C.c_size_t]
            lib.lib.fftw_b200_device_free.argtypes = [C.c_void_p]
            self.lib = lib
            self.ptr = lib.lib.fftw_b200_device_malloc(max(arr.nbytes, 16))
            self.view = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_ubyte)), shape=(arr.nbytes,)).view(arr.dtype).reshape(arr.shape)
            self.view[...] = arr
        else:
            import torch
            self.t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
            self.ptr = self.t.data_ptr()

    def set(self, arr):
        if self.emu:
            self.view[...] = arr
        else:
            import torch
            self.t.copy_(torch.from_numpy(np.ascontiguousarray(arr)))

    def get(self):
        if self.emu:
            return self.view.copy()
        import torch
        torch.cuda.synchronize()
        return self.t.cpu().numpy()

    def free(self):
        if self.emu:
            self.lib.lib.fftw_b200_device_free(self.ptr)


def _c2c(lib, prec, shape, flags, inplace=True):
    rng = np.random.default_rng(5)
    x0 = F.rand_complex(rng, shape, prec)
    dx = Dev(lib, x0)
    dy = dx if inplace else Dev(lib, np.zeros_like(x0))
    p = lib.fn(prec, "plan_dft")(len(shape), (C.c_int * len(shape))(*shape), dx.ptr, dy.ptr, -1, flags)
    assert p
    txt = " ".join(lib.sprint_plan(prec, p).split())
    dx.set(x0)
    lib.execute(prec, p)
    y = dy.get()
    lib.destroy_plan(prec, p)
    dx.free()
    if not inplace:
        dy.free()
    return O.rel_l2(y, O.dft(x0, sign=-1, rank=len(shape))), F.tol_for(prec, shape), txt


@pytest.mark.parametrize("prec", ["d", "f"])
@pytest.mark.parametrize("env", [
    {"FFTW3_B200_L2_BLOCK_KB": "64", "FFTW3_B200_L2_LANES": "2"},
    {"FFTW3_B200_L2_BLOCK_KB": "32", "FFTW3_B200_L2_LANES": "4", "FFTW3_B200_L2_PAIR": "outer"},
    {"FFTW3_B200_SPLIT": "1", "FFTW3_B200_SPLIT_KB": "64"},
    {"FFTW3_B200_SPLIT": "2", "FFTW3_B200_SPLIT_KB": "128", "FFTW3_B200_SPLIT_LANES": "2"},
])
@pytest.mark.parametrize("shape", [(32, 64, 64), (16, 1024, 64), (1024, 24)])
def test_pinned_decompositions_against_the_oracle(host_lib, prec, env, shape, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for inplace in (True, False):
        err, tol, txt = _c2c(host_lib, prec, shape, B.FFTW_ESTIMATE, inplace)
        assert err <= tol, (env, shape, inplace, err, tol, txt)


@pytest.mark.parametrize("prec", ["d", "f"])
def test_measure_chooses_among_whole_plan_alternatives(host_lib, prec, monkeypatch):
    """FFTW_MEASURE on device arrays times the alternatives of each problem kind; the chosen index is wisdom under
    the problem's signature, so an ESTIMATE plan of the same problem afterwards is the same plan."""
    monkeypatch.setenv("FFTW3_B200_ALT_MIN_KB", "1")
    host_lib.fn(prec, "forget_wisdom")()
    # 3-d c2c: plain / L2 pairs (2 group sizes)
    err, tol, txt = _c2c(host_lib, prec, (32, 64, 64), B.FFTW_MEASURE)
    assert err <= tol, (err, tol, txt)
    err2, _, txt2 = _c2c(host_lib, prec, (32, 64, 64), B.FFTW_ESTIMATE)
    assert err2 <= tol and txt2 == txt, (txt, txt2)
    # batched prime size: rule / Rader / Bluestein
    err, tol, txt = _c2c(host_lib, prec, (64, 1009), B.FFTW_MEASURE)
    assert err <= 4 * tol, (err, tol, txt)
    assert "rader" in txt or "bluestein" in txt
    # real transforms of long even lines: split / merge fused or apart
    rdt, cdt = (np.float64, np.complex128) if prec == "d" else (np.float32, np.complex64)
    n, hm = 1 << 16, 3
    rng = np.random.default_rng(7)
    xr = F.rand_real(rng, (hm, n), prec)
    dxr, dxc = Dev(host_lib, xr), Dev(host_lib, np.zeros((hm, n // 2 + 1), dtype=cdt))
    p = host_lib.plan_many_dft_r2c(prec, [n], hm, dxr.ptr, None, 1, n, dxc.ptr, None, 1, n // 2 + 1, B.FFTW_MEASURE)
    assert p
    dxr.set(xr)
    host_lib.execute(prec, p)
    X = dxc.get()
    host_lib.destroy_plan(prec, p)
    ref = np.fft.rfft(xr.astype(np.float64), axis=1)
    assert O.rel_l2(X, ref) <= F.tol_for(prec, (n,))
    p = host_lib.plan_many_dft_c2r(prec, [n], hm, dxc.ptr, None, 1, n // 2 + 1, dxr.ptr, None, 1, n, B.FFTW_MEASURE)
    assert p
    dxc.set(ref.astype(cdt))
    host_lib.execute(prec, p)
    back = dxr.get()
    host_lib.destroy_plan(prec, p)
    assert O.rel_l2(back / n, xr.astype(np.float64)) <= 2 * F.tol_for(prec, (n,))
    dxr.free(); dxc.free()
    # dense 2-d r2r with long columns: transposed stores or transposes
    shape = (4096, 8) if prec == "d" else (8192, 8)
    a0 = F.rand_real(rng, shape, prec)
    da, db = Dev(host_lib, a0), Dev(host_lib, np.zeros_like(a0))
    p = host_lib.fn(prec, "plan_r2r_2d")(shape[0], shape[1], da.ptr, db.ptr, B.R2R_KINDS["REDFT10"], B.R2R_KINDS["RODFT01"], B.FFTW_MEASURE)
    assert p
    da.set(a0)
    host_lib.execute(prec, p)
    got = db.get()
    host_lib.destroy_plan(prec, p)
    want = O.r2r(a0.astype(np.float64), ["REDFT10", "RODFT01"])
    assert O.rel_l2(got, want) <= F.tol_for(prec, shape)
    da.free(); db.free()
    host_lib.fn(prec, "forget_wisdom")()
