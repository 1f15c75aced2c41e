"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_gpu_check.py
Every rank transforms its slab with the real CUDA kernels (peer-store and
NCCL all-to-all exchange paths, natural and transposed output); rank 0 gathers
and compares with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from fftw3_b200 import binding as B
    from fftw3_b200 import dist as D
    from oracle import oracle as O
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    lib = B.load()
    ok = True
    for shape in [(64, 48, 32), (30, 14, 25), (128, 128, 128)]:
        n0, n1, n2 = shape
        rng = np.random.default_rng(5)
        full = rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)
        ref = O.dft(full) if rank == 0 else None
        for exchange in ("peer", "peer-gather", "collective"):
            for transposed in (False, True):
                alloc, ln0, s0, ln1, s1 = D.local_size_3d(lib, n0, n1, n2, rank, world)
                local = torch.zeros(max(alloc, 1), dtype=torch.complex128, device="cuda")
                if ln0:
                    local[:ln0 * n1 * n2] = torch.from_numpy(full[s0:s0 + ln0].reshape(-1).copy()).cuda()
                # "peer": both exchanges fused into pass stores (push plan); "peer-gather": the older
                # plan whose second exchange is a gather stage
                os.environ["FFTW3_B200_DIST_PUSH"] = "0" if exchange == "peer-gather" else "1"
                pl = D.SlabPlan3D(lib, n0, n1, n2, local, flags=B.FFTW_ESTIMATE, transposed_out=transposed,
                                  exchange="peer" if exchange == "peer-gather" else exchange)
                pushed = pl.push
                pl.execute()
                pl.execute() if False else None
                torch.cuda.synchronize()
                pl.destroy()
                cnt = (ln1 * n0 * n2) if transposed else (ln0 * n1 * n2)
                mine = local[:cnt].cpu()
                gathered = [None] * world
                dist.all_gather_object(gathered, (rank, s1 if transposed else s0, ln1 if transposed else ln0, mine.numpy()))
                if rank == 0:
                    got = np.zeros(shape, dtype=np.complex128)
                    for r, start, c, arr in gathered:
                        if c == 0:
                            continue
                        if transposed:
                            got[:, start:start + c, :] = arr.reshape(c, n0, n2).transpose(1, 0, 2)
                        else:
                            got[start:start + c] = arr.reshape(c, n1, n2)
                    err = O.rel_l2(got, ref)
                    good = err < 5e-15
                    ok &= good
                    print("dist check %s P=%d %-11s transposed=%d push=%d rel L2 %.2e %s"
                          % (shape, world, exchange, transposed, pushed, err, "OK" if good else "FAIL"), flush=True)
    # real data: r2c then c2r (round trip) of padded slabs, even and uneven column blocks
    # (30, 17, 25) and (19, 34, 20): uneven blocks AND a non-smooth distributed dimension (copy fallback of dist.c)
    for shape in [(64, 16 * world, 50), (24, 8 * world, 33), (30, 7 * world + 1, 25), (30, 17, 25), (19, 34, 20), (128, 128, 128)]:
        n0, n1, n2 = shape
        h = n2 // 2 + 1
        rng = np.random.default_rng(7)
        full = rng.uniform(-0.5, 0.5, shape)
        ref = O.r2c(full, rank=3) if rank == 0 else None
        for inplace in (False, True):
            alloc, ln0, s0, ln1, s1 = D.local_size_3d(lib, n0, n1, n2, rank, world)
            real = torch.zeros(max(ln0, 1) * n1 * 2 * h, dtype=torch.float64, device="cuda")
            if ln0:
                pad = np.zeros((ln0, n1, 2 * h))
                pad[:, :, :n2] = full[s0:s0 + ln0]
                real[:pad.size] = torch.from_numpy(pad.reshape(-1)).cuda()
            cplx = real.view(torch.complex128) if inplace else torch.zeros(max(ln0, 1) * n1 * h, dtype=torch.complex128,
                                                                          device="cuda")
            fw = D.SlabPlanReal3D(lib, n0, n1, n2, real, cplx, "r2c", flags=B.FFTW_ESTIMATE)
            fw.execute()
            torch.cuda.synchronize()
            spec = cplx[:ln0 * n1 * h].cpu().numpy().copy()
            fw.destroy()
            bw = D.SlabPlanReal3D(lib, n0, n1, n2, real, cplx, "c2r", flags=B.FFTW_ESTIMATE)
            bw.execute()
            torch.cuda.synchronize()
            back = real[:ln0 * n1 * 2 * h].cpu().numpy().reshape(ln0, n1, 2 * h)[:, :, :n2] / (n0 * n1 * n2) if ln0 else None
            bw.destroy()
            gathered = [None] * world
            dist.all_gather_object(gathered, (s0, ln0, spec, back))
            if rank == 0:
                got = np.zeros((n0, n1, h), dtype=np.complex128)
                rt = np.zeros(shape)
                for start, c, sp, bk in gathered:
                    if c:
                        got[start:start + c] = sp.reshape(c, n1, h)
                        rt[start:start + c] = bk
                e1, e2 = O.rel_l2(got, ref), O.rel_l2(rt, full)
                good = e1 < 5e-15 and e2 < 5e-15
                ok &= good
                print("dist check %s P=%d r2c/c2r inplace=%d rel L2 fwd %.2e roundtrip %.2e %s"
                      % (shape, world, inplace, e1, e2, "OK" if good else "FAIL"), flush=True)
    # r2r: kinds per dimension, uneven blocks included
    for shape, kinds in [((64, 48, 32), ("REDFT10", "RODFT01", "R2HC")), ((30, 7 * world + 1, 25), ("DHT", "REDFT00", "RODFT11"))]:
        n0, n1, n2 = shape
        rng = np.random.default_rng(11)
        full = rng.uniform(-0.5, 0.5, shape)
        ref = O.r2r(full, list(kinds), rank=3) if rank == 0 else None
        alloc, ln0, s0, ln1, s1 = D.local_size_3d(lib, n0, n1, n2, rank, world)
        local = torch.zeros(max(ln0, 1) * n1 * n2, dtype=torch.float64, device="cuda")
        if ln0:
            local[:ln0 * n1 * n2] = torch.from_numpy(full[s0:s0 + ln0].reshape(-1).copy()).cuda()
        pl = D.SlabPlanR2R3D(lib, n0, n1, n2, local, kinds, flags=B.FFTW_ESTIMATE)
        pl.execute()
        torch.cuda.synchronize()
        pl.destroy()
        gathered = [None] * world
        dist.all_gather_object(gathered, (s0, ln0, local[:ln0 * n1 * n2].cpu().numpy()))
        if rank == 0:
            got = np.zeros(shape)
            for start, c, arr in gathered:
                if c:
                    got[start:start + c] = arr.reshape(c, n1, n2)
            err = O.rel_l2(got, ref)
            good = err < 2e-14
            ok &= good
            print("dist check %s P=%d r2r %s rel L2 %.2e %s" % (shape, world, "/".join(kinds), err, "OK" if good else "FAIL"),
                  flush=True)
    # the C communicator interface (fftw_b200_mpi_*): library-owned exchange buffers, IPC mappings and device-side barriers
    comm = D.torch_comm()
    for n, kw in [((64, 48, 32), {}), ((64, 48, 32), {"transposed": True}), ((30, 14, 25), {"sign": 1}), ((96, 80), {}),
                  ((45, 64), {"transposed": True}), ((32, 24, 20), {"howmany": 2}), ((8, 12, 10, 6), {}),
                  ((64, 48, 32), {"prec": "f"}), ((34, 19, 6), {}), ((128, 128, 128), {})]:
        howmany, prec, sign, transposed = kw.get("howmany", 1), kw.get("prec", "d"), kw.get("sign", -1), kw.get("transposed", False)
        cdt, tdt = (np.complex64, torch.complex64) if prec == "f" else (np.complex128, torch.complex128)
        shape = tuple(n) + ((howmany,) if howmany > 1 else ())
        rng = np.random.default_rng(13)
        full = (rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)).astype(cdt)
        ref = None
        if rank == 0:
            ref = O.dft(np.moveaxis(full, -1, 0) if howmany > 1 else full, sign=sign, rank=len(n))
            if howmany > 1:
                ref = np.moveaxis(ref, 0, -1)
        probe = D.CommPlan.__new__(D.CommPlan)      # sizes before allocation
        D._declare(lib); D._declare_mpi(lib)
        import ctypes as C
        nn = (C.c_ssize_t * len(n))(*n)
        v = [C.c_ssize_t() for _ in range(4)]
        alloc = int(lib.lib.fftw_b200_mpi_local_size_many_transposed(len(n), nn, howmany, 0, 0, C.byref(comm), *[C.byref(x) for x in v]))
        ln0, s0, ln1, s1 = [int(x.value) for x in v]
        local = torch.zeros(max(alloc, 1), dtype=tdt, device="cuda")
        if ln0:
            local[:full[s0:s0 + ln0].size] = torch.from_numpy(full[s0:s0 + ln0].reshape(-1).copy()).cuda()
        pl = D.CommPlan(lib, list(n), comm, local.data_ptr(), howmany=howmany, prec=prec, sign=sign, transposed_out=transposed)
        assert pl.plan, "fftw_b200_mpi_plan_many_dft returned NULL"
        pl.execute()
        pl.execute() if False else None
        torch.cuda.synchronize()
        rest = int(np.prod(shape[2:])) if len(shape) > 2 else 1
        cnt = (ln1 * n[0] * rest) if transposed else (ln0 * n[1] * rest)
        mine = local[:cnt].cpu().numpy()
        pl.destroy()
        gathered = [None] * world
        dist.all_gather_object(gathered, (s1 if transposed else s0, ln1 if transposed else ln0, mine))
        if rank == 0:
            got = np.zeros(shape, dtype=cdt)
            for start, c, arr in gathered:
                if not c:
                    continue
                if transposed:
                    got[:, start:start + c] = np.moveaxis(arr.reshape((c, n[0]) + shape[2:]), 0, 1)
                else:
                    got[start:start + c] = arr.reshape((c, n[1]) + shape[2:])
            err = O.rel_l2(got, ref)
            good = err < (3e-6 if prec == "f" else 5e-15)
            ok &= good
            print("dist check comm-api %s P=%d %s rel L2 %.2e %s" % (n, world, kw, err, "OK" if good else "FAIL"), flush=True)
    # distributed 1-D transform (six-step) and distributed transposes
    for n0, kw in [(1 << 20, {}), (1 << 20, {"sign": 1}), (1000 * 1024, {}), (1 << 16, {"prec": "f"}), (6 * 35 * 11, {}),
                   (1 << 18, {"inplace": True})]:
        prec, sign, inplace = kw.get("prec", "d"), kw.get("sign", -1), kw.get("inplace", False)
        cdt, tdt = (np.complex64, torch.complex64) if prec == "f" else (np.complex128, torch.complex128)
        rng = np.random.default_rng(17)
        x = (rng.uniform(-0.5, 0.5, n0) + 1j * rng.uniform(-0.5, 0.5, n0)).astype(cdt)
        ref = O.dft(x, sign=sign, rank=1) if rank == 0 else None
        alloc, lni, si, lno, so = D.local_size_1d(lib, n0, comm, sign, 0)
        a = torch.zeros(max(alloc, 1), dtype=tdt, device="cuda")
        b = a if inplace else torch.zeros(max(alloc, 1), dtype=tdt, device="cuda")
        a[:lni] = torch.from_numpy(x[si:si + lni].copy()).cuda()
        pl = D.CommPlan1D(lib, n0, comm, a.data_ptr(), b.data_ptr(), prec=prec, sign=sign)
        assert pl.plan, "fftw_b200_mpi_plan_dft_1d returned NULL"
        pl.execute()
        torch.cuda.synchronize()
        mine = b[:lno].cpu().numpy()
        pl.destroy()
        gathered = [None] * world
        dist.all_gather_object(gathered, (so, mine))
        if rank == 0:
            got = np.zeros(n0, cdt)
            for start, arr in gathered:
                got[start:start + len(arr)] = arr
            err = O.rel_l2(got, ref)
            good = err < (4e-6 if prec == "f" else 2e-14)
            ok &= good
            print("dist check 1-d six-step n=%d P=%d %s rel L2 %.2e %s" % (n0, world, kw, err, "OK" if good else "FAIL"), flush=True)
    for n0, n1, hm, inplace in [(96, 80, 1, False), (45, 64, 3, False), (128, 64, 1, True), (1024, 2048, 2, False)]:
        full = np.arange(n0 * n1 * hm, dtype=np.float64).reshape(n0, n1, hm)
        b0, b1 = -(-n0 // world), -(-n1 // world)
        ln0, s0 = max(0, min(b0, n0 - b0 * rank)), min(b0 * rank, n0)
        ln1, s1 = max(0, min(b1, n1 - b1 * rank)), min(b1 * rank, n1)
        cnt = max(b0 * n1, b1 * n0) * hm
        a = torch.zeros(max(cnt, 1), dtype=torch.float64, device="cuda")
        b = a if inplace else torch.zeros(max(cnt, 1), dtype=torch.float64, device="cuda")
        a[:ln0 * n1 * hm] = torch.from_numpy(full[s0:s0 + ln0].reshape(-1).copy()).cuda()
        pl = D.CommTranspose(lib, n0, n1, comm, a.data_ptr(), b.data_ptr(), howmany=hm)
        assert pl.plan, "fftw_b200_mpi_plan_many_transpose returned NULL"
        pl.execute()
        torch.cuda.synchronize()
        mine = b[:ln1 * n0 * hm].cpu().numpy().reshape(ln1, n0, hm)
        pl.destroy()
        gathered = [None] * world
        dist.all_gather_object(gathered, (s1, mine))
        if rank == 0:
            got = np.zeros((n1, n0, hm))
            for start, arr in gathered:
                got[start:start + arr.shape[0]] = arr
            good = np.array_equal(got, full.transpose(1, 0, 2))
            ok &= good
            print("dist check transpose %dx%dx%d P=%d inplace=%d %s" % (n0, n1, hm, world, inplace, "OK" if good else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
