"""Peer-exchange plans of the distributed transform, all ranks simulated in ONE process.

The peer path (exchange fused into pass stores through peer-mapped pointers) needs CUDA IPC
between processes, which the CPU test double cannot provide -- but inside one process every
"rank" can simply be handed the other ranks' buffers.  That exercises dist.c's plan
construction (scatter targets, row-split push stores, gather sources, uneven and idle ranks)
and the kernels' peer addressing on the emulated device layer, stage by stage with the same
barrier structure the real launcher uses (fftw3_b200/dist.py: SlabPlan3D.execute).
Reference behaviour: mpi/dft-rank-geq2.c:40-59, mpi/dft-rank-geq2-transposed.c:47-70."""
import ctypes as C

import numpy as np
import pytest

from fftw3_b200 import binding as B
from fftw3_b200 import dist as D
from oracle import oracle as O


def _run(lib, shape, P, mode, sign):
    D._declare(lib)
    L = lib.lib
    n0, n1, n2 = shape
    rng = np.random.default_rng(11)
    full = rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)
    info = [D.local_size_3d(lib, n0, n1, n2, r, P) for r in range(P)]
    b0 = (n0 + P - 1) // P
    nbytes = [16 * max(i[0], 1) for i in info]
    loc = [L.fftw_b200_device_malloc(nb) for nb in nbytes]
    zb = [L.fftw_b200_device_malloc(nb) for nb in nbytes]
    assert all(loc) and all(zb)

    def view(ptr, count):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(2 * count,)).view(np.complex128)

    for r, (alloc, ln0, s0, ln1, s1) in enumerate(info):
        view(loc[r], max(alloc, 1))[:] = 0
        if ln0:
            view(loc[r], alloc)[:ln0 * n1 * n2] = full[s0:s0 + ln0].reshape(-1)
    plans = []
    VP = C.c_void_p * P
    for r in range(P):
        share1 = [info[d][3] for d in range(P)]
        push = VP(*[zb[d] + 16 * (r * b0) * share1[d] * n2 for d in range(P)])
        if mode == "push":
            p = L.fftw_b200_dist_plan_dft_3d_push(n0, n1, n2, r, P, loc[r], zb[r], push, VP(*loc), sign, B.FFTW_ESTIMATE)
        elif mode == "gather":
            pull = VP(*[zb[s] + 16 * (r * b0) * share1[s] * n2 for s in range(P)])
            p = L.fftw_b200_dist_plan_dft_3d(n0, n1, n2, r, P, loc[r], zb[r], push, pull, sign, B.FFTW_ESTIMATE)
        else:
            p = L.fftw_b200_dist_plan_dft_3d(n0, n1, n2, r, P, loc[r], zb[r], push, None, sign, B.FFTW_ESTIMATE)
        assert p, (mode, r)
        plans.append(p)
    nst = L.fftw_b200_dist_num_stages(plans[0])
    assert nst == (3 if mode == "gather" else 2)
    for st in range(nst):                     # a barrier between stages = finish the stage on every rank
        for r in range(P):
            L.fftw_b200_dist_execute_stage(plans[r], st)
    want = O.dft(full, sign=sign) if sign < 0 else O.dft(full, sign=+1)
    err = 0.0
    for r, (alloc, ln0, s0, ln1, s1) in enumerate(info):
        if mode == "transposed":
            if ln1:
                got = view(loc[r], alloc)[:ln1 * n0 * n2].reshape(ln1, n0, n2)
                err = max(err, O.rel_l2(got, np.transpose(want, (1, 0, 2))[s1:s1 + ln1]))
        elif ln0:
            got = view(loc[r], alloc)[:ln0 * n1 * n2].reshape(ln0, n1, n2)
            err = max(err, O.rel_l2(got, want[s0:s0 + ln0]))
    for p in plans:
        L.fftw_b200_dist_destroy_plan(p)
    for q in loc + zb:
        L.fftw_b200_device_free(q)
    return err


@pytest.mark.parametrize("mode", ["push", "gather", "transposed"])
@pytest.mark.parametrize("shape,P,sign", [
    ((8, 6, 10), 2, -1),
    ((12, 10, 8), 3, +1),       # uneven column blocks (4, 4, 2)
    ((6, 5, 4), 4, -1),         # block 2: the last rank owns no planes
    ((16, 16, 16), 4, -1),      # power-of-two block: shift path of the row split
    ((5, 3, 7), 1, -1),
    ((12, 10, 1), 3, -1),       # a 2-d array as n0 x n1 x 1 (fftw_mpi_plan_dft_2d)
])
def test_peer_plans_all_ranks_in_one_process(emu_lib, mode, shape, P, sign):
    err = _run(emu_lib, shape, P, mode, sign)
    assert err <= 1e-14, (mode, shape, P, err)


@pytest.mark.parametrize("mode", ["push", "transposed"])
@pytest.mark.parametrize("shape,P,sign", [
    ((8, 6, 10), 2, -1),
    ((12, 10, 8), 3, +1),       # uneven column blocks (4, 4, 2): strided copies of different widths
    ((6, 5, 4), 4, -1),         # the last rank owns no planes: nothing to copy from it
    ((16, 16, 16), 4, -1),
])
def test_first_exchange_by_copy_engines(emu_lib, mode, shape, P, sign, monkeypatch):
    """FFTW3_B200_DIST_EXCHANGE=copy: X runs in place (the rows a rank keeps go straight into its own exchange
    buffer) and one strided copy per peer moves each chunk's blocks (csrc/host/dist.c, exchange_by_copy)."""
    monkeypatch.setenv("FFTW3_B200_DIST_EXCHANGE", "copy")
    err = _run(emu_lib, shape, P, mode, sign)
    assert err <= 1e-14, (mode, shape, P, err)


def _run_real(lib, shape, P, inplace):
    """r2c then c2r of a real n0 x n1 x n2 array over P simulated ranks; returns (err_fwd, err_roundtrip)."""
    D._declare(lib)
    L = lib.lib
    n0, n1, n2 = shape
    h = n2 // 2 + 1
    rng = np.random.default_rng(3)
    full = rng.uniform(-0.5, 0.5, shape)
    b0, b1 = (n0 + P - 1) // P, (n1 + P - 1) // P
    ln0 = [max(0, min(b0, n0 - b0 * r)) for r in range(P)]
    slab_c = max(b0 * n1 * h, 1)                      # complex elements of a slab / of zbuf ([n0][b1][h] <= P*b0*b1*h)
    zb_c = max(P * b0 * b1 * h, 1)
    cpl = [L.fftw_b200_device_malloc(16 * slab_c) for _ in range(P)]
    rea = cpl if inplace else [L.fftw_b200_device_malloc(16 * slab_c) for _ in range(P)]
    zb = [L.fftw_b200_device_malloc(16 * zb_c) for _ in range(P)]

    def rview(ptr, count):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(count,))

    for r in range(P):
        rv = rview(rea[r], 2 * slab_c)
        rv[:] = 0
        if ln0[r]:
            pad = np.zeros((ln0[r], n1, 2 * h))
            pad[:, :, :n2] = full[r * b0:r * b0 + ln0[r]]
            rv[:pad.size] = pad.reshape(-1)
    VP = C.c_void_p * P
    fwd, bwd = [], []
    for r in range(P):
        push = VP(*[zb[d] + 16 * (r * b0) * b1 * h for d in range(P)])
        out = VP(*cpl)
        p = L.fftw_b200_dist_plan_dft_r2c_3d(n0, n1, n2, r, P, rea[r], cpl[r], zb[r], push, out, B.FFTW_ESTIMATE)
        assert p, ("r2c", r)
        fwd.append(p)
        p = L.fftw_b200_dist_plan_dft_c2r_3d(n0, n1, n2, r, P, cpl[r], rea[r], zb[r], push, out, B.FFTW_ESTIMATE)
        assert p, ("c2r", r)
        bwd.append(p)
    assert L.fftw_b200_dist_num_stages(fwd[0]) == 2 and L.fftw_b200_dist_num_stages(bwd[0]) == 3
    for st in range(2):
        for r in range(P):
            L.fftw_b200_dist_execute_stage(fwd[r], st)
    want = O.r2c(full, rank=3)
    err_f = 0.0
    for r in range(P):
        if ln0[r]:
            got = rview(cpl[r], 2 * slab_c)[:2 * ln0[r] * n1 * h].view(np.complex128).reshape(ln0[r], n1, h)
            err_f = max(err_f, O.rel_l2(got, want[r * b0:r * b0 + ln0[r]]))
    for st in range(3):
        for r in range(P):
            L.fftw_b200_dist_execute_stage(bwd[r], st)
    err_b = 0.0
    for r in range(P):
        if ln0[r]:
            got = rview(rea[r], 2 * slab_c)[:ln0[r] * n1 * 2 * h].reshape(ln0[r], n1, 2 * h)[:, :, :n2]
            err_b = max(err_b, O.rel_l2(got / (n0 * n1 * n2), full[r * b0:r * b0 + ln0[r]]))
    for p in fwd + bwd:
        L.fftw_b200_dist_destroy_plan(p)
    for q in set(cpl + rea + zb):
        L.fftw_b200_device_free(q)
    return err_f, err_b


@pytest.mark.parametrize("inplace", [False, True])
@pytest.mark.parametrize("shape,P", [
    ((8, 6, 10), 2),
    ((12, 9, 7), 3),        # odd last dimension
    ((12, 10, 7), 3),       # uneven column blocks (4, 4, 2)
    ((6, 5, 4), 4),         # column blocks (2, 2, 1, 0) and the last rank owns no planes
    ((6, 8, 16), 4),        # block 2: the last rank owns no planes
    ((5, 3, 8), 1),
    ((6, 17, 10), 2),       # non-smooth n1 (Rader / Bluestein pass cannot split its stores by row): copy fallback
    ((19, 12, 6), 3),       # non-smooth n0: the same fallback on the second exchange
    ((17, 19, 5), 2),       # both
])
def test_real_data_plans_all_ranks_in_one_process(emu_lib, shape, P, inplace):
    """Distributed r2c / c2r (mpi/api.c:650-760): local real pass over the rows plus two c2c
    passes whose row-split stores carry the exchanges; forward against the oracle, then the
    c2r of that result against the input (unnormalised: scaled by n0*n1*n2)."""
    err_f, err_b = _run_real(emu_lib, shape, P, inplace)
    assert err_f <= 1e-14 and err_b <= 1e-14, (shape, P, inplace, err_f, err_b)


@pytest.mark.parametrize("shape,P,kinds", [
    ((8, 6, 10), 2, ("REDFT10", "RODFT01", "R2HC")),
    ((12, 10, 7), 3, ("DHT", "REDFT00", "RODFT11")),       # uneven column blocks
    ((6, 5, 4), 4, ("RODFT00", "HC2R", "REDFT11")),        # idle last rank
    ((5, 3, 8), 1, ("REDFT01", "REDFT10", "RODFT10")),
])
def test_r2r_plans_all_ranks_in_one_process(emu_lib, shape, P, kinds):
    """Distributed r2r (mpi/api.c:770-886): local 2-d r2r, gather of the column blocks, r2r along
    dim 0, gather back -- every rank simulated in one process, stage by stage."""
    lib = emu_lib
    D._declare(lib)
    L = lib.lib
    n0, n1, n2 = shape
    rng = np.random.default_rng(21)
    full = rng.uniform(-0.5, 0.5, shape)
    b0, b1 = (n0 + P - 1) // P, (n1 + P - 1) // P
    ln0 = [max(0, min(b0, n0 - b0 * r)) for r in range(P)]
    ln1 = [max(0, min(b1, n1 - b1 * r)) for r in range(P)]
    loc = [L.fftw_b200_device_malloc(8 * max(b0 * n1 * n2, 1)) for _ in range(P)]
    zb = [L.fftw_b200_device_malloc(8 * max(n0 * b1 * n2, 1)) for _ in range(P)]

    def rview(ptr, count):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(count,))

    for r in range(P):
        if ln0[r]:
            rview(loc[r], b0 * n1 * n2)[:ln0[r] * n1 * n2] = full[r * b0:r * b0 + ln0[r]].reshape(-1)
    VP = C.c_void_p * P
    ks = (C.c_int * 3)(*[B.R2R_KINDS[k] for k in kinds])
    plans = []
    for r in range(P):
        p = L.fftw_b200_dist_plan_r2r_3d(n0, n1, n2, r, P, loc[r], zb[r], VP(*loc), VP(*zb), ks, B.FFTW_ESTIMATE)
        assert p, r
        plans.append(p)
    for st in range(3):
        for r in range(P):
            L.fftw_b200_dist_execute_stage(plans[r], st)
    want = O.r2r(full, list(kinds), rank=3)
    for r in range(P):
        if ln0[r]:
            got = rview(loc[r], b0 * n1 * n2)[:ln0[r] * n1 * n2].reshape(ln0[r], n1, n2)
            assert O.rel_l2(got, want[r * b0:r * b0 + ln0[r]]) < 1e-13, (shape, P, kinds, r)
    for p in plans:
        L.fftw_b200_dist_destroy_plan(p)
    for q in loc + zb:
        L.fftw_b200_device_free(q)
