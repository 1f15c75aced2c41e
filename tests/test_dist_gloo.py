"""N>1 path on CPU: slab-decomposed 3-D transform over world_size 2 and 3
(uneven blocks) with the gloo backend.  The host layer is the real one; the
device layer is the unit-test double (tests/emu).  Mirrors how the reference
tests its MPI code: several ranks on one machine, gather, compare
(mpi/mpi-bench.c:66-120, mpi/Makefile.am:52-72)."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, shape, transposed, sign, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fftw3_b200 import binding as B
        from fftw3_b200 import dist as D
        lib = B.Lib(os.path.join(ROOT, "tests", "_emu", "libfftw3_b200_emu.so"))
        n0, n1, n2 = shape
        rng = np.random.default_rng(5)
        full = rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)
        alloc, ln0, s0, ln1, s1 = D.local_size_3d(lib, n0, n1, n2, rank, world)
        local = torch.zeros(max(alloc, 1), dtype=torch.complex128)
        if ln0:
            local[:ln0 * n1 * n2] = torch.from_numpy(full[s0:s0 + ln0].reshape(-1))
        plan = D.SlabPlan3D(lib, n0, n1, n2, local, sign=sign, flags=B.FFTW_ESTIMATE,
                            transposed_out=transposed, exchange="collective")
        plan.execute()
        plan.destroy()
        if transposed:
            out = local[:ln1 * n0 * n2].numpy().reshape(ln1, n0, n2).copy()
            q.put((rank, s1, ln1, out))
        else:
            out = local[:ln0 * n1 * n2].numpy().reshape(ln0, n1, n2).copy()
            q.put((rank, s0, ln0, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,transposed,sign", [
    (2, (8, 6, 10), False, -1),
    (2, (8, 6, 10), True, -1),
    (3, (7, 5, 4), False, 1),        # uneven blocks, one short rank
    (3, (4, 9, 6), True, -1),
    (4, (5, 6, 8), False, -1),       # a rank with a short block and possibly an idle one
])
def test_slab_3d_matches_oracle(emu_lib, world, shape, transposed, sign):
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, transposed, sign, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    full = rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)
    ref = O.dft(full, sign=sign)
    n0, n1, n2 = shape
    got = np.zeros(shape, dtype=np.complex128)
    for rank, start, cnt, out in parts:
        if cnt == 0:
            continue
        if transposed:       # [local_n1][n0][n2] holds X[k0][k1][k2] at out[k1 - start][k0][k2]
            got[:, start:start + cnt, :] = out.transpose(1, 0, 2)
        else:
            got[start:start + cnt] = out
    assert O.rel_l2(got, ref) < 1e-14


def _batch_worker(rank, world, port, n, howmany, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fftw3_b200 import binding as B
        from fftw3_b200 import dist as D
        lib = B.Lib(os.path.join(ROOT, "tests", "_emu", "libfftw3_b200_emu.so"))
        rng = np.random.default_rng(9)
        full = rng.uniform(-0.5, 0.5, (howmany, n)) + 1j * rng.uniform(-0.5, 0.5, (howmany, n))
        count, first = D.batch_share(howmany, rank, world)
        local = np.ascontiguousarray(full[first:first + count]) if count else np.zeros((1, n), dtype=np.complex128)
        plan = D.ShardedBatchPlan(lib, n, howmany, local, flags=B.FFTW_ESTIMATE)
        plan.execute()
        plan.destroy()
        q.put((rank, first, count, local[:count].copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,howmany", [(2, 64, 10), (3, 30, 7), (4, 16, 3)])
def test_batched_1d_shards_without_collective(emu_lib, world, n, howmany):
    """Independent transforms are block-distributed over the ranks and need no exchange
    (SURVEY.md section 8e): every rank plans its share; the union equals the oracle."""
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_batch_worker, args=(r, world, port, n, howmany, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(9)
    full = rng.uniform(-0.5, 0.5, (howmany, n)) + 1j * rng.uniform(-0.5, 0.5, (howmany, n))
    got = np.zeros_like(full)
    seen = 0
    for rank, first, count, arr in parts:
        got[first:first + count] = arr
        seen += count
    assert seen == howmany
    assert O.rel_l2(got, O.dft(full, rank=1)) < 1e-14


def _wisdom_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fftw3_b200 import binding as B
        from fftw3_b200 import dist as D
        lib = B.Lib(os.path.join(ROOT, "tests", "_emu", "libfftw3_b200_emu.so"))
        D._declare(lib)
        D._declare_mpi(lib)
        comm = D.torch_comm()

        def plans(n, flags):
            x = np.zeros((4, n), dtype=np.complex128)
            p = lib.plan_many_dft("d", [n], 4, x.ctypes.data, None, 1, n, x.ctypes.data, None, 1, n, -1, flags)
            ok = bool(p)
            if p:
                lib.destroy_plan("d", p)
            return ok

        only = B.FFTW_MEASURE | B.FFTW_WISDOM_ONLY
        res = {}
        # rank 0 measures a size nobody else knows, then broadcasts
        if rank == 0:
            assert plans(1024, B.FFTW_MEASURE)
        res["before_bcast"] = plans(1024, only)
        lib.lib.fftw_b200_mpi_broadcast_wisdom(C.byref(comm))
        res["after_bcast"] = plans(1024, only)
        # the last rank measures another size, then everything is gathered on rank 0
        if rank == world - 1:
            assert plans(512, B.FFTW_MEASURE)
        res["before_gather"] = plans(512, only)
        lib.lib.fftw_b200_mpi_gather_wisdom(C.byref(comm))
        res["after_gather"] = plans(512, only)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_wisdom_broadcast_and_gather_over_the_communicator(emu_lib):
    """fftw_mpi_broadcast_wisdom / fftw_mpi_gather_wisdom (mpi/wisdom-api.c): FFTW_WISDOM_ONLY plans fail on the ranks
    that have not measured a size and succeed once the wisdom has travelled (3 processes, gloo all-gather as the
    communicator callback)."""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_wisdom_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert got[r]["before_bcast"] == (r == 0), got
        assert got[r]["after_bcast"], got
        assert got[r]["before_gather"] == (r == world - 1), got
        assert got[r]["after_gather"] == (r in (0, world - 1)), got
