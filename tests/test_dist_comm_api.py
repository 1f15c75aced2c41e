"""The fftw_mpi_* shaped C interface (include/fftw3_b200_dist.h, fftw_b200_mpi_*): plan_many_dft / plan_dft /
2-D / 3-D, local_size*, execute with the library's own device-side barriers.  Here every rank is a THREAD of
this process on the emulated device layer (whose "IPC" hands pointers through and whose barrier spins on the
same flags as the CUDA kernel); the all-gather the interface asks of its launcher is a threading.Barrier.
The same interface runs across real GPUs in tests/dist_gpu_check.py (pytest -m gpu on a multi-GPU box).
Reference behaviour restated: mpi/api.c:248-352 (local sizes), :560-648 (plans), doc/mpi.texi:440-463."""
import ctypes as C
import threading

import numpy as np
import pytest

from fftw3_b200 import binding as B
from fftw3_b200 import dist as D
from oracle import oracle as O


class ThreadComm:
    """all-gather among threads"""

    def __init__(self, P):
        self.P = P
        self.bar = threading.Barrier(P)
        self.slots = [None] * P

    def comm(self, rank):
        def ag(ctx, send, recv, nbytes):
            self.slots[rank] = C.string_at(send, nbytes)
            self.bar.wait()
            data = b"".join(self.slots)
            C.memmove(recv, data, len(data))
            self.bar.wait()
            return 0
        cb = D.ALLGATHER_FN(ag)
        cs = D.CommStruct(rank, self.P, cb, None)
        cs._keep = cb
        return cs


def _run(lib, n, P, howmany=1, prec="d", sign=-1, transposed=False, inplace=True, transposed_in=False, block=0, tblock=0):
    D._declare(lib)
    L = lib.lib
    cdt = np.complex64 if prec == "f" else np.complex128
    rng = np.random.default_rng(3)
    shape = tuple(n) + ((howmany,) if howmany > 1 else ())
    full = (rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)).astype(cdt)
    ref = O.dft(np.moveaxis(full, -1, 0) if howmany > 1 else full, sign=sign, rank=len(n))
    if howmany > 1:
        ref = np.moveaxis(ref, 0, -1)
    tc = ThreadComm(P)
    got = [None] * P
    errs = []

    def rank_main(r):
        try:
            comm = tc.comm(r)
            # sizes first (no device memory yet)
            nn = (C.c_ssize_t * len(n))(*n)
            v = [C.c_ssize_t() for _ in range(4)]
            D._declare_mpi(lib)
            alloc = int(L.fftw_b200_mpi_local_size_many_transposed(len(n), nn, howmany, block, tblock, C.byref(comm),
                                                                  *[C.byref(x) for x in v]))
            ln0, s0, ln1, s1 = [int(x.value) for x in v]
            isz = np.dtype(cdt).itemsize
            a = L.fftw_b200_device_malloc(max(alloc, 1) * isz)
            b = a if inplace else L.fftw_b200_device_malloc(max(alloc, 1) * isz)
            view = lambda ptr: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)), shape=(max(alloc, 1) * isz,)).view(cdt)
            if transposed_in:
                # this rank holds columns [s1, s1 + ln1) of the global array as [ln1][n0][rest]
                if ln1:
                    piece = np.ascontiguousarray(np.moveaxis(full[:, s1:s1 + ln1], 1, 0))
                    view(a)[:piece.size] = piece.reshape(-1)
            elif ln0:
                view(a)[:full[s0:s0 + ln0].size] = full[s0:s0 + ln0].reshape(-1)
            pl = D.CommPlan(lib, list(n), comm, a, None if inplace else b, howmany=howmany, prec=prec, sign=sign,
                            transposed_out=transposed, transposed_in=transposed_in, block=block, tblock=tblock)
            assert pl.plan, "plan_many_dft returned NULL on rank %d" % r
            pl.execute()
            rest = int(np.prod(shape[2:])) if len(shape) > 2 else 1
            if transposed:
                got[r] = (s1, ln1, view(b)[:ln1 * n[0] * rest].copy().reshape((ln1, n[0]) + shape[2:]) if ln1 else None)
            else:
                got[r] = (s0, ln0, view(b)[:ln0 * n[1] * rest].copy().reshape((ln0, n[1]) + shape[2:]) if ln0 else None)
            pl.destroy()
            L.fftw_b200_device_free(a)
            if not inplace:
                L.fftw_b200_device_free(b)
        except BaseException as e:       # noqa: surface it in the main thread
            errs.append((r, repr(e)))
            tc.bar.abort()

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(P)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not errs, errs
    out = np.zeros(shape, dtype=cdt)
    for start, cnt, arr in got:
        if not cnt:
            continue
        if transposed:
            out[:, start:start + cnt] = np.moveaxis(arr, 0, 1)
        else:
            out[start:start + cnt] = arr
    return O.rel_l2(out, ref)


@pytest.mark.parametrize("n,P,kw", [
    ((8, 6, 10), 2, {}),                                  # fused plans of dist.c behind the communicator interface
    ((8, 6, 10), 2, {"transposed": True}),
    ((12, 10, 7), 3, {"sign": 1}),                        # uneven column blocks
    ((6, 5, 4), 4, {}),                                   # idle ranks
    ((16, 12), 2, {}),                                    # 2-D (general path)
    ((9, 10), 3, {"transposed": True}),
    ((8, 6, 10), 2, {"howmany": 3}),                      # plan_many_dft
    ((4, 6, 5, 3), 2, {}),                                # rank 4
    ((8, 6, 10), 2, {"prec": "f"}),                       # fftwf_b200_mpi_*
    ((8, 6, 10), 2, {"inplace": False}),
    ((17, 19, 3), 2, {}),                                 # non-smooth distributed dims
    ((8, 6, 10), 2, {"transposed_in": True}),             # FFTW_MPI_TRANSPOSED_IN: input [local_n1][n0][n2]
    ((12, 10, 7), 3, {"transposed_in": True, "transposed": True, "sign": 1}),
    ((9, 10), 3, {"transposed_in": True, "inplace": False}),
    ((6, 8, 5), 2, {"transposed_in": True, "howmany": 2}),
    ((8, 6, 10), 2, {"block": 5, "tblock": 4}),           # the caller's own block sizes (mpi/block.c:52-70): 5 + 3 planes, 4 + 2 columns
    ((12, 10), 3, {"block": 6, "tblock": 4, "transposed": True}),      # 6 + 6 + 0 rows
    ((9, 7, 4), 3, {"block": 4, "transposed_in": True}),
])
def test_comm_interface_all_ranks_as_threads(emu_lib, n, P, kw):
    err = _run(emu_lib, n, P, **kw)
    assert err <= (2e-6 if kw.get("prec") == "f" else 1e-14), (n, P, kw, err)


def test_local_size_matches_the_reference_distribution(emu_lib):
    """block = ceil(n / P) (mpi/block.c:37-50); alloc covers both the slab and the transposed slab"""
    D._declare(emu_lib)
    D._declare_mpi(emu_lib)
    tc = ThreadComm(1)
    for P in (1, 2, 3, 5, 8):
        tot0 = tot1 = 0
        for r in range(P):
            cs = tc.comm(0)
            cs.rank, cs.nranks = r, P
            nn = (C.c_ssize_t * 3)(10, 7, 6)
            v = [C.c_ssize_t() for _ in range(4)]
            alloc = emu_lib.lib.fftw_b200_mpi_local_size_many_transposed(3, nn, 2, 0, 0, C.byref(cs), *[C.byref(x) for x in v])
            ln0, s0, ln1, s1 = [int(x.value) for x in v]
            b0, b1 = -(-10 // P), -(-7 // P)
            assert s0 == min(b0 * r, 10) and ln0 == max(0, min(b0, 10 - b0 * r))
            assert s1 == min(b1 * r, 7) and ln1 == max(0, min(b1, 7 - b1 * r))
            assert alloc >= max(ln0 * 7, ln1 * 10) * 6 * 2
            tot0 += ln0
            tot1 += ln1
        assert tot0 == 10 and tot1 == 7


def _threads(P, fn):
    tc = ThreadComm(P)
    errs, res = [], [None] * P

    def main(r):
        try:
            res[r] = fn(r, tc.comm(r))
        except BaseException as e:       # noqa
            errs.append((r, repr(e)))
            tc.bar.abort()
    th = [threading.Thread(target=main, args=(r,)) for r in range(P)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not errs, errs
    return res


@pytest.mark.parametrize("n0,P,kw", [
    (4096, 2, {}), (4096, 2, {"sign": 1}), (1000, 3, {}), (6 * 35, 4, {}), (4096, 2, {"scrambled": True}),
    (2048, 2, {"prec": "f"}), (1024, 2, {"inplace": True}), (34 * 3, 2, {}),
])
def test_distributed_1d_six_step(emu_lib, n0, P, kw):
    """fftw_mpi_plan_dft_1d (mpi/dft-rank1.c:81-148): input block-distributed in natural order, output
    block-distributed in natural order (or [k1][k2]-scrambled), against the oracle."""
    lib = emu_lib
    D._declare(lib)
    L = lib.lib
    prec, sign, scr, inplace = kw.get("prec", "d"), kw.get("sign", -1), kw.get("scrambled", False), kw.get("inplace", False)
    cdt = np.complex64 if prec == "f" else np.complex128
    isz = np.dtype(cdt).itemsize
    rng = np.random.default_rng(5)
    x = (rng.uniform(-0.5, 0.5, n0) + 1j * rng.uniform(-0.5, 0.5, n0)).astype(cdt)
    ref = O.dft(x, sign=sign, rank=1)

    def rank_main(r, comm):
        alloc, lni, si, lno, so = D.local_size_1d(lib, n0, comm, sign, (1 << 28) if scr else 0)
        assert alloc >= max(lni, lno)
        a = L.fftw_b200_device_malloc(max(alloc, 1) * isz)
        b = a if inplace else L.fftw_b200_device_malloc(max(alloc, 1) * isz)
        view = lambda ptr: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)), shape=(max(alloc, 1) * isz,)).view(cdt)
        view(a)[:lni] = x[si:si + lni]
        pl = D.CommPlan1D(lib, n0, comm, a, b, prec=prec, sign=sign, scrambled_out=scr)
        assert pl.plan, "plan_dft_1d returned NULL"
        pl.execute()
        out = (so, view(b)[:lno].copy())
        pl.destroy()
        L.fftw_b200_device_free(a)
        if not inplace:
            L.fftw_b200_device_free(b)
        return out

    res = _threads(P, rank_main)
    got = np.zeros(n0, cdt)
    for so, arr in res:
        got[so:so + len(arr)] = arr
    if scr:
        # scrambled: position k1 * m + k2 holds X[k1 + r k2]; recover r from the first rank's block
        r_ = None
        for cand in range(2, n0):
            if n0 % cand == 0:
                m_ = n0 // cand
                y = got.reshape(cand, m_).T.reshape(-1)
                if O.rel_l2(y, ref) < 1e-5:
                    r_ = cand
                    got = y
                    break
        assert r_ is not None, "no r x m un-scrambling matches"
    assert O.rel_l2(got, ref) <= (3e-6 if prec == "f" else 2e-14), (n0, P, kw)


@pytest.mark.parametrize("n0,P,prec,inplace", [(4096, 2, "d", False), (1000, 3, "d", True), (6 * 35, 4, "d", False), (2048, 2, "f", True)])
def test_distributed_1d_scrambled_out_then_scrambled_in(emu_lib, n0, P, prec, inplace):
    """FFTW_MPI_SCRAMBLED_OUT forward followed by FFTW_MPI_SCRAMBLED_IN backward (mpi/fftw3-mpi.h:209-211,
    mpi/dft-rank1.c:224-340): the pair skips the transposes that would only put the spectrum in natural order; the
    forward result is checked against the oracle through its [k1][k2] layout, the round trip returns n0 * x in
    natural order."""
    lib = emu_lib
    D._declare(lib)
    L = lib.lib
    cdt = np.complex64 if prec == "f" else np.complex128
    isz = np.dtype(cdt).itemsize
    rng = np.random.default_rng(9)
    x = (rng.uniform(-0.5, 0.5, n0) + 1j * rng.uniform(-0.5, 0.5, n0)).astype(cdt)

    def rank_main(r, comm):
        alloc, lni, si, lno, so = D.local_size_1d(lib, n0, comm, -1, 1 << 28)
        alloc2, lni2, si2, lno2, so2 = D.local_size_1d(lib, n0, comm, +1, 1 << 27)
        assert (lni2, si2) == (lno, so)                       # the scrambled layout is the same on both sides
        cnt = max(alloc, alloc2, 1)
        a = L.fftw_b200_device_malloc(cnt * isz)
        b = a if inplace else L.fftw_b200_device_malloc(cnt * isz)
        view = lambda ptr: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)), shape=(cnt * isz,)).view(cdt)
        view(a)[:lni] = x[si:si + lni]
        fwd = D.CommPlan1D(lib, n0, comm, a, b, prec=prec, sign=-1, scrambled_out=True)
        bwd = D.CommPlan1D(lib, n0, comm, b, a, prec=prec, sign=+1, scrambled_in=True)
        assert fwd.plan and bwd.plan
        fwd.execute()
        spec = (so, view(b)[:lno].copy())
        bwd.execute()
        back = (so2, view(a)[:lno2].copy())
        fwd.destroy(); bwd.destroy()
        L.fftw_b200_device_free(a)
        if not inplace:
            L.fftw_b200_device_free(b)
        return spec, back

    res = _threads(P, rank_main)
    back = np.zeros(n0, cdt)
    for _, (so2, arr) in res:
        back[so2:so2 + len(arr)] = arr
    assert O.rel_l2(back / n0, x) <= (3e-6 if prec == "f" else 2e-14), (n0, P)


@pytest.mark.parametrize("blocks", [(0, 0), (1, 1)])
@pytest.mark.parametrize("n0,n1,P,hm,inplace,prec", [(12, 10, 2, 1, False, "d"), (7, 9, 3, 2, False, "d"), (8, 6, 2, 1, True, "d"),
                                                     (5, 16, 4, 3, True, "f"), (64, 48, 2, 1, False, "d")])
def test_distributed_transpose(emu_lib, n0, n1, P, hm, inplace, prec, blocks):
    """fftw_mpi_plan_many_transpose (mpi/api.c:521-556): bit-exact; blocks = (1, 1): the caller's own block sizes,
    one more than the default on both sides (so the last rank gets less, possibly nothing)"""
    lib = emu_lib
    D._declare(lib)
    L = lib.lib
    rdt = np.float32 if prec == "f" else np.float64
    isz = np.dtype(rdt).itemsize
    full = np.arange(n0 * n1 * hm, dtype=rdt).reshape(n0, n1, hm)

    def rank_main(r, comm):
        b0, b1 = -(-n0 // P) + blocks[0], -(-n1 // P) + blocks[1]
        ln0, s0 = max(0, min(b0, n0 - b0 * r)), min(b0 * r, n0)
        ln1, s1 = max(0, min(b1, n1 - b1 * r)), min(b1 * r, n1)
        cnt = max(b0 * n1, b1 * n0) * hm
        a = L.fftw_b200_device_malloc(max(cnt, 1) * isz)
        b = a if inplace else L.fftw_b200_device_malloc(max(cnt, 1) * isz)
        view = lambda ptr: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)), shape=(max(cnt, 1) * isz,)).view(rdt)
        view(a)[:ln0 * n1 * hm] = full[s0:s0 + ln0].reshape(-1)
        pl = D.CommTranspose(lib, n0, n1, comm, a, b, howmany=hm, prec=prec,
                             block0=b0 if blocks[0] else 0, block1=b1 if blocks[1] else 0)
        assert pl.plan
        pl.execute()
        out = (s1, view(b)[:ln1 * n0 * hm].copy().reshape(ln1, n0, hm))
        pl.destroy()
        L.fftw_b200_device_free(a)
        if not inplace:
            L.fftw_b200_device_free(b)
        return out

    res = _threads(P, rank_main)
    got = np.zeros((n1, n0, hm), rdt)
    for s1, arr in res:
        got[s1:s1 + arr.shape[0]] = arr
    assert np.array_equal(got, full.transpose(1, 0, 2))


@pytest.mark.parametrize("shape,P,inplace", [((8, 6, 10), 2, False), ((8, 6, 10), 2, True), ((12, 10, 7), 3, False),
                                              ((6, 5, 4), 4, True), ((17, 19, 6), 2, False)])
def test_real_data_3d_through_the_communicator_interface(emu_lib, shape, P, inplace):
    """fftw_mpi_plan_dft_r2c_3d / _c2r_3d shapes (mpi/api.c:650-760): padded real slabs, r2c against numpy's rfftn
    of the whole array, c2r back to n0 n1 n2 x; uneven, idle and non-smooth blocks."""
    lib = emu_lib
    D._declare(lib)
    L = lib.lib
    n0, n1, n2 = shape
    h = n2 // 2 + 1
    rng = np.random.default_rng(13)
    full = rng.uniform(-0.5, 0.5, shape)
    ref = np.fft.rfftn(full)

    def rank_main(r, comm):
        nn = (C.c_ssize_t * 3)(n0, n1, h)
        v = [C.c_ssize_t() for _ in range(4)]
        D._declare_mpi(lib)
        alloc = int(L.fftw_b200_mpi_local_size_many_transposed(3, nn, 1, 0, 0, C.byref(comm), *[C.byref(x) for x in v]))
        ln0, s0 = int(v[0].value), int(v[1].value)
        cplx = L.fftw_b200_device_malloc(max(alloc, 1) * 16)
        real = cplx if inplace else L.fftw_b200_device_malloc(max(alloc, 1) * 16)
        rview = np.ctypeslib.as_array(C.cast(real, C.POINTER(C.c_double)), shape=(max(alloc, 1) * 2,))
        cview = np.ctypeslib.as_array(C.cast(cplx, C.POINTER(C.c_double)), shape=(max(alloc, 1) * 2,)).view(np.complex128)
        pad = rview[:ln0 * n1 * 2 * h].reshape(ln0, n1, 2 * h)
        pad[:, :, :n2] = full[s0:s0 + ln0]
        fwd = D.CommPlanReal3D(lib, shape, comm, real, cplx, "r2c")
        bwd = D.CommPlanReal3D(lib, shape, comm, cplx, real, "c2r")
        assert fwd.plan and bwd.plan
        pad[:, :, :n2] = full[s0:s0 + ln0]
        fwd.execute()
        spec = cview[:ln0 * n1 * h].copy().reshape(ln0, n1, h)
        bwd.execute()
        back = rview[:ln0 * n1 * 2 * h].reshape(ln0, n1, 2 * h)[:, :, :n2].copy()
        fwd.destroy(); bwd.destroy()
        L.fftw_b200_device_free(cplx)
        if not inplace:
            L.fftw_b200_device_free(real)
        return s0, spec, back

    res = _threads(P, rank_main)
    got = np.zeros((n0, n1, h), np.complex128)
    back = np.zeros(shape)
    for s0, spec, b in res:
        got[s0:s0 + spec.shape[0]] = spec
        back[s0:s0 + b.shape[0]] = b
    assert O.rel_l2(got, ref) <= 2e-14, (shape, P)
    assert O.rel_l2(back / (n0 * n1 * n2), full) <= 2e-14, (shape, P)


@pytest.mark.parametrize("shape,P,kinds,inplace", [((8, 6, 10), 2, ("REDFT10", "RODFT01", "R2HC"), True),
                                                    ((9, 10, 4), 3, ("DHT", "REDFT00", "RODFT11"), False),
                                                    ((6, 5, 4), 4, ("REDFT01", "REDFT10", "HC2R"), True)])
def test_r2r_3d_through_the_communicator_interface(emu_lib, shape, P, kinds, inplace):
    """fftw_mpi_plan_r2r_3d shape (mpi/api.c:770-886) against the oracle's separable r2r"""
    lib = emu_lib
    D._declare(lib)
    L = lib.lib
    n0, n1, n2 = shape
    rng = np.random.default_rng(17)
    full = rng.uniform(-0.5, 0.5, shape)
    ref = O.r2r(full, list(kinds))

    def rank_main(r, comm):
        b0 = -(-n0 // P)
        ln0, s0 = max(0, min(b0, n0 - b0 * r)), min(b0 * r, n0)
        cnt = max(b0 * n1 * n2, 1)
        a = L.fftw_b200_device_malloc(cnt * 8)
        b = a if inplace else L.fftw_b200_device_malloc(cnt * 8)
        view = lambda ptr: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(cnt,))
        view(a)[:ln0 * n1 * n2] = full[s0:s0 + ln0].reshape(-1)
        pl = D.CommPlanReal3D(lib, shape, comm, a, b, "r2r", kinds=kinds)
        assert pl.plan
        pl.execute()
        out = view(b)[:ln0 * n1 * n2].copy().reshape(ln0, n1, n2)
        if not inplace:
            assert np.array_equal(view(a)[:ln0 * n1 * n2], full[s0:s0 + ln0].reshape(-1))
        pl.destroy()
        L.fftw_b200_device_free(a)
        if not inplace:
            L.fftw_b200_device_free(b)
        return s0, out

    res = _threads(P, rank_main)
    got = np.zeros(shape)
    for s0, arr in res:
        got[s0:s0 + arr.shape[0]] = arr
    assert O.rel_l2(got, ref) <= 2e-14, (shape, P, kinds)


@pytest.mark.parametrize("shape,P,inplace", [((8, 10), 2, False), ((8, 10), 2, True), ((12, 9), 3, False), ((6, 4), 4, True),
                                              ((17, 22), 2, False)])
def test_real_data_2d_through_the_communicator_interface(emu_lib, shape, P, inplace):
    """fftw_mpi_plan_dft_r2c_2d / _c2r_2d shapes: the halved dimension (n1/2 + 1 complex columns) is the one that
    is exchanged; r2c against numpy's rfft2 of the whole array, c2r back to n0 n1 x."""
    lib = emu_lib
    D._declare(lib)
    L = lib.lib
    n0, n1 = shape
    h = n1 // 2 + 1
    rng = np.random.default_rng(19)
    full = rng.uniform(-0.5, 0.5, shape)
    ref = np.fft.rfft2(full)

    def rank_main(r, comm):
        nn = (C.c_ssize_t * 2)(n0, h)
        v = [C.c_ssize_t() for _ in range(4)]
        D._declare_mpi(lib)
        alloc = int(L.fftw_b200_mpi_local_size_many_transposed(2, nn, 1, 0, 0, C.byref(comm), *[C.byref(x) for x in v]))
        ln0, s0 = int(v[0].value), int(v[1].value)
        cplx = L.fftw_b200_device_malloc(max(alloc, 1) * 16)
        real = cplx if inplace else L.fftw_b200_device_malloc(max(alloc, 1) * 16)
        rview = np.ctypeslib.as_array(C.cast(real, C.POINTER(C.c_double)), shape=(max(alloc, 1) * 2,))
        cview = np.ctypeslib.as_array(C.cast(cplx, C.POINTER(C.c_double)), shape=(max(alloc, 1) * 2,)).view(np.complex128)
        pad = rview[:ln0 * 2 * h].reshape(ln0, 2 * h)
        fwd = D.CommPlanReal3D(lib, shape, comm, real, cplx, "r2c")
        bwd = D.CommPlanReal3D(lib, shape, comm, cplx, real, "c2r")
        assert fwd.plan and bwd.plan
        pad[:, :n1] = full[s0:s0 + ln0]
        fwd.execute()
        spec = cview[:ln0 * h].copy().reshape(ln0, h)
        bwd.execute()
        back = rview[:ln0 * 2 * h].reshape(ln0, 2 * h)[:, :n1].copy()
        fwd.destroy(); bwd.destroy()
        L.fftw_b200_device_free(cplx)
        if not inplace:
            L.fftw_b200_device_free(real)
        return s0, spec, back

    res = _threads(P, rank_main)
    got = np.zeros((n0, h), np.complex128)
    back = np.zeros(shape)
    for s0, spec, b in res:
        got[s0:s0 + spec.shape[0]] = spec
        back[s0:s0 + b.shape[0]] = b
    assert O.rel_l2(got, ref) <= 2e-14, (shape, P)
    assert O.rel_l2(back / (n0 * n1), full) <= 2e-14, (shape, P)


@pytest.mark.parametrize("shape,P,kinds,kw", [
    ((12, 10), 2, ("REDFT10", "RODFT01"), {}),
    ((9, 8), 3, ("R2HC", "DHT"), {"inplace": False}),
    ((8, 6, 5), 2, ("RODFT00", "REDFT11", "HC2R"), {"howmany": 2}),
    ((4, 6, 5, 3), 2, ("REDFT01", "REDFT10", "DHT", "RODFT10"), {}),
    ((6, 5), 4, ("REDFT00", "RODFT11"), {"prec": "f"}),
])
def test_many_r2r_through_the_communicator_interface(emu_lib, shape, P, kinds, kw):
    """fftw_mpi_plan_many_r2r shape (mpi/api.c:770-886): any rank >= 2, howmany interleaved tuples, both precisions,
    against the oracle's separable r2r."""
    lib = emu_lib
    D._declare(lib)
    L = lib.lib
    howmany, prec, inplace = kw.get("howmany", 1), kw.get("prec", "d"), kw.get("inplace", True)
    rdt = np.float32 if prec == "f" else np.float64
    isz = np.dtype(rdt).itemsize
    n0, n1 = shape[0], shape[1]
    rest = int(np.prod(shape[2:])) * howmany if len(shape) > 2 else howmany
    rng = np.random.default_rng(23)
    full = rng.uniform(-0.5, 0.5, shape + ((howmany,) if howmany > 1 else ())).astype(rdt)
    x64 = full.astype(np.float64)
    ref = O.r2r(np.moveaxis(x64, -1, 0), list(kinds), rank=len(shape)) if howmany > 1 else O.r2r(x64, list(kinds))
    if howmany > 1:
        ref = np.moveaxis(ref, 0, -1)

    def rank_main(r, comm):
        b0, b1 = -(-n0 // P), -(-n1 // P)
        ln0, s0 = max(0, min(b0, n0 - b0 * r)), min(b0 * r, n0)
        cnt = max(b0 * n1, b1 * n0) * rest
        a = L.fftw_b200_device_malloc(max(cnt, 1) * isz)
        b = a if inplace else L.fftw_b200_device_malloc(max(cnt, 1) * isz)
        view = lambda ptr: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)), shape=(max(cnt, 1) * isz,)).view(rdt)
        view(a)[:ln0 * n1 * rest] = full[s0:s0 + ln0].reshape(-1)
        pl = D.CommPlanManyR2R(lib, list(shape), comm, a, b, kinds, howmany=howmany, prec=prec)
        assert pl.plan
        pl.execute()
        out = view(b)[:ln0 * n1 * rest].copy().reshape((ln0,) + full.shape[1:])
        pl.destroy()
        L.fftw_b200_device_free(a)
        if not inplace:
            L.fftw_b200_device_free(b)
        return s0, out

    res = _threads(P, rank_main)
    got = np.zeros(full.shape, np.float64)
    for s0, arr in res:
        got[s0:s0 + arr.shape[0]] = arr
    assert O.rel_l2(got, ref) <= (5e-6 if prec == "f" else 3e-14), (shape, P, kinds, kw)


@pytest.mark.parametrize("shape,P,kw", [
    ((8, 6, 10), 2, {"howmany": 2}),
    ((6, 5, 4, 7), 3, {}),
    ((8, 6, 9), 2, {"prec": "f", "inplace": True}),
    ((5, 4, 6), 4, {"howmany": 3, "inplace": True}),
])
def test_many_real_data_through_the_communicator_interface(emu_lib, shape, P, kw):
    """fftw_mpi_plan_many_dft_r2c / _c2r shapes (mpi/api.c:650-760): rank >= 3, howmany interleaved tuples, both
    precisions; r2c against numpy's rfftn, c2r back to N x."""
    lib = emu_lib
    D._declare(lib)
    L = lib.lib
    howmany, prec, inplace = kw.get("howmany", 1), kw.get("prec", "d"), kw.get("inplace", False)
    rdt, cdt = (np.float32, np.complex64) if prec == "f" else (np.float64, np.complex128)
    n0, n1, nl = shape[0], shape[1], shape[-1]
    h = nl // 2 + 1
    mid = shape[2:-1]
    rng = np.random.default_rng(29)
    full = rng.uniform(-0.5, 0.5, shape + (howmany,)).astype(rdt)
    ref = np.fft.rfftn(full.astype(np.float64), axes=tuple(range(len(shape))))
    Rc = int(np.prod(mid)) * h * howmany                      # complex numbers per (i0, i1)

    def rank_main(r, comm):
        b0, b1 = -(-n0 // P), -(-n1 // P)
        ln0, s0 = max(0, min(b0, n0 - b0 * r)), min(b0 * r, n0)
        cnt = max(b0 * n1, b1 * n0) * Rc                      # complex elements
        csz = np.dtype(cdt).itemsize
        cplx = L.fftw_b200_device_malloc(max(cnt, 1) * csz)
        real = cplx if inplace else L.fftw_b200_device_malloc(max(cnt, 1) * csz)
        rview = np.ctypeslib.as_array(C.cast(real, C.POINTER(C.c_ubyte)), shape=(max(cnt, 1) * csz,)).view(rdt)
        cview = np.ctypeslib.as_array(C.cast(cplx, C.POINTER(C.c_ubyte)), shape=(max(cnt, 1) * csz,)).view(cdt)
        pshape = (ln0, n1) + mid + (2 * h, howmany)
        pad = rview[:int(np.prod(pshape))].reshape(pshape)
        fwd = D.CommPlanManyReal(lib, list(shape), comm, real, cplx, "r2c", howmany=howmany, prec=prec)
        bwd = D.CommPlanManyReal(lib, list(shape), comm, cplx, real, "c2r", howmany=howmany, prec=prec)
        assert fwd.plan and bwd.plan
        pad[..., :nl, :] = full[s0:s0 + ln0]
        fwd.execute()
        cshape = (ln0, n1) + mid + (h, howmany)
        spec = cview[:int(np.prod(cshape))].copy().reshape(cshape)
        bwd.execute()
        back = rview[:int(np.prod(pshape))].reshape(pshape)[..., :nl, :].copy()
        fwd.destroy(); bwd.destroy()
        L.fftw_b200_device_free(cplx)
        if not inplace:
            L.fftw_b200_device_free(real)
        return s0, spec, back

    res = _threads(P, rank_main)
    got = np.zeros(ref.shape, np.complex128)
    back = np.zeros(full.shape)
    for s0, spec, b in res:
        got[s0:s0 + spec.shape[0]] = spec
        back[s0:s0 + b.shape[0]] = b
    tol = 5e-6 if prec == "f" else 3e-14
    assert O.rel_l2(got, ref) <= tol, (shape, P, kw)
    assert O.rel_l2(back / float(np.prod(shape)), full.astype(np.float64)) <= tol, (shape, P, kw)


@pytest.mark.parametrize("shape,P", [((8, 10), 2), ((6, 5, 8), 3)])
def test_single_precision_basic_real_and_r2r_forms(emu_lib, shape, P):
    """fftwf_b200_mpi_plan_dft_r2c_2d/_3d, _c2r_2d/_3d, _r2r_2d/_3d: the single-precision basic forms"""
    lib = emu_lib
    D._declare(lib)
    L = lib.lib
    nl = shape[-1]
    h = nl // 2 + 1
    n0, n1 = shape[0], shape[1]
    mid = shape[2:-1]
    rng = np.random.default_rng(31)
    full = rng.uniform(-0.5, 0.5, shape).astype(np.float32)
    ref = np.fft.rfftn(full.astype(np.float64))
    kinds = ("REDFT10", "RODFT01", "DHT")[:len(shape)]
    ref_r2r = O.r2r(full.astype(np.float64), list(kinds))
    # rank 2: the halved dimension is exchanged (rows of h complex); rank 3: n1 is exchanged
    cols = h if len(shape) == 2 else n1
    Rc = 1 if len(shape) == 2 else int(np.prod(mid)) * h

    def rank_main(r, comm):
        b0, b1 = -(-n0 // P), -(-cols // P)
        ln0, s0 = max(0, min(b0, n0 - b0 * r)), min(b0 * r, n0)
        cnt = max(b0 * cols, b1 * n0) * Rc
        cplx = L.fftw_b200_device_malloc(max(cnt, 1) * 8)
        real = L.fftw_b200_device_malloc(max(cnt, 1) * 8)
        rview = np.ctypeslib.as_array(C.cast(real, C.POINTER(C.c_float)), shape=(max(cnt, 1) * 2,))
        cview = np.ctypeslib.as_array(C.cast(cplx, C.POINTER(C.c_float)), shape=(max(cnt, 1) * 2,)).view(np.complex64)
        pshape = (ln0,) + shape[1:-1] + (2 * h,)
        pad = rview[:int(np.prod(pshape))].reshape(pshape)
        fwd = D.CommPlanReal3D(lib, shape, comm, real, cplx, "r2c", prec="f")
        bwd = D.CommPlanReal3D(lib, shape, comm, cplx, real, "c2r", prec="f")
        assert fwd.plan and bwd.plan
        pad[..., :nl] = full[s0:s0 + ln0]
        fwd.execute()
        cshape = (ln0,) + shape[1:-1] + (h,)
        spec = cview[:int(np.prod(cshape))].copy().reshape(cshape)
        bwd.execute()
        back = rview[:int(np.prod(pshape))].reshape(pshape)[..., :nl].copy()
        fwd.destroy(); bwd.destroy()
        # r2r in place on a slab of the natural shape
        rr = L.fftw_b200_device_malloc(max(-(-n0 // P) * n1, -(-n1 // P) * n0, 1) * int(np.prod(shape[2:])) * 4 if len(shape) > 2
                                       else max(-(-n0 // P) * n1, -(-n1 // P) * n0, 1) * 4)
        nel = ln0 * int(np.prod(shape[1:]))
        v2 = np.ctypeslib.as_array(C.cast(rr, C.POINTER(C.c_float)), shape=(max(nel, 1),))
        v2[:nel] = full[s0:s0 + ln0].reshape(-1)
        pr = D.CommPlanReal3D(lib, shape, comm, rr, rr, "r2r", kinds=kinds, prec="f")
        assert pr.plan
        pr.execute()
        r2r_out = v2[:nel].copy().reshape((ln0,) + shape[1:])
        pr.destroy()
        for ptr in (cplx, real, rr):
            L.fftw_b200_device_free(ptr)
        return s0, spec, back, r2r_out

    res = _threads(P, rank_main)
    got = np.zeros(ref.shape, np.complex128)
    back = np.zeros(shape)
    got_r2r = np.zeros(shape)
    for s0, spec, b, rr in res:
        got[s0:s0 + spec.shape[0]] = spec
        back[s0:s0 + b.shape[0]] = b
        got_r2r[s0:s0 + rr.shape[0]] = rr
    assert O.rel_l2(got, ref) <= 5e-6
    assert O.rel_l2(back / float(np.prod(shape)), full.astype(np.float64)) <= 5e-6
    assert O.rel_l2(got_r2r, ref_r2r) <= 5e-6
