"""fftw3_b200.dist -- slab-decomposed 3-D transforms over several GPUs, one
process per GPU (the fftw_mpi_plan_dft_3d equivalent; see
include/fftw3_b200_dist.h for the C-ABI and the reference files it mirrors).

This module is plumbing only: it allocates the exchange buffers, trades CUDA-IPC
handles or runs the all-to-all through ``torch.distributed``, and puts the
barriers between the stages.  All arithmetic and all data movement inside a GPU
(and, on the peer path, between GPUs) is done by the library's CUDA kernels.

Exchange paths
  * ``peer``        the last local FFT pass stores straight into the peers'
                    exchange buffers (CUDA IPC mappings over NVLink/NVSwitch):
                    transpose fused with its collective; ranks only meet at
                    barriers.  GPU only.
  * ``collective``  the pass writes a local send buffer and an all-to-all moves
                    it (NCCL on GPUs; send/recv pairs on the gloo backend used by
                    the CPU unit tests).
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import binding as B

FFTW_MPI_TRANSPOSED_OUT = 1 << 30          # same bit as mpi/fftw3-mpi.h:214
FFTW_MPI_TRANSPOSED_IN = 1 << 29


def _declare(lib):
    L = lib.lib
    if getattr(L, "_dist_declared", False):
        return
    P, I = C.c_void_p, C.c_int
    SP = C.POINTER(C.c_ssize_t)
    L.fftw_b200_dist_local_size_3d.restype = C.c_ssize_t
    L.fftw_b200_dist_local_size_3d.argtypes = [C.c_ssize_t] * 3 + [I, I, SP, SP, SP, SP]
    L.fftw_b200_dist_plan_dft_3d.restype = P
    L.fftw_b200_dist_plan_dft_3d.argtypes = [C.c_ssize_t] * 3 + [I, I, P, P, P, P, I, C.c_uint]
    L.fftw_b200_dist_plan_dft_3d_push.restype = P
    L.fftw_b200_dist_plan_dft_3d_push.argtypes = [C.c_ssize_t] * 3 + [I, I, P, P, P, P, I, C.c_uint]
    for name in ("fftw_b200_dist_plan_dft_r2c_3d", "fftw_b200_dist_plan_dft_c2r_3d"):
        getattr(L, name).restype = P
        getattr(L, name).argtypes = [C.c_ssize_t] * 3 + [I, I, P, P, P, P, P, C.c_uint]
    L.fftw_b200_dist_plan_r2r_3d.restype = P
    L.fftw_b200_dist_plan_r2r_3d.argtypes = [C.c_ssize_t] * 3 + [I, I, P, P, P, P, P, C.c_uint]
    L.fftw_b200_ipc_offset.restype = C.c_ssize_t
    L.fftw_b200_ipc_offset.argtypes = [P]
    L.fftw_b200_dist_num_stages.argtypes = [P]
    L.fftw_b200_dist_execute_stage.argtypes = [P, I]
    L.fftw_b200_dist_num_chunks.argtypes = [P, I]
    L.fftw_b200_dist_exchange_by_copy.argtypes = [P]
    L.fftw_b200_dist_partition_sms.argtypes = [P]
    L.fftw_b200_dist_execute_chunk.argtypes = [P, I, I]
    L.fftw_b200_dist_join.argtypes = [P]
    L.fftw_b200_dist_destroy_plan.argtypes = [P]
    L.fftw_b200_device_malloc.restype = P
    L.fftw_b200_device_malloc.argtypes = [C.c_size_t]
    L.fftw_b200_device_free.argtypes = [P]
    L.fftw_b200_ipc_export.argtypes = [P, C.c_char_p]
    L.fftw_b200_ipc_import.restype = P
    L.fftw_b200_ipc_import.argtypes = [C.c_char_p]
    L.fftw_b200_ipc_close.argtypes = [P]
    L._dist_declared = True


# --------------------------------------------------------------------------
# communicator interface (include/fftw3_b200_dist.h: fftw_b200_mpi_*)
# --------------------------------------------------------------------------
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)


class CommStruct(C.Structure):
    _fields_ = [("rank", C.c_int), ("nranks", C.c_int), ("allgather", ALLGATHER_FN), ("ctx", C.c_void_p)]


def torch_comm(group=None):
    """fftw_b200_comm whose all-gather is torch.distributed's (the launcher-specific collective the C
    interface asks for; an MPI program would pass MPI_Allgather instead).  Keep the returned object alive as
    long as plans made with it are being created."""
    rank, P = dist.get_rank(group), dist.get_world_size(group)
    on_gpu = dist.get_backend(group) == "nccl"

    def ag(ctx, send, recv, nbytes):
        try:
            buf = torch.frombuffer(bytearray(C.string_at(send, nbytes)), dtype=torch.uint8)
            if on_gpu:
                buf = buf.cuda()
            out = [torch.empty_like(buf) for _ in range(P)]
            dist.all_gather(out, buf, group=group)
            data = torch.cat(out).cpu().numpy().tobytes()
            C.memmove(recv, data, len(data))
            return 0
        except Exception:        # never let an exception cross the C boundary
            return 1

    cb = ALLGATHER_FN(ag)
    cs = CommStruct(rank, P, cb, None)
    cs._keep = cb
    return cs


def _declare_mpi(lib):
    L = lib.lib
    if getattr(L, "_mpi_declared", False):
        return
    P, I, U = C.c_void_p, C.c_int, C.c_uint
    S = C.c_ssize_t
    SP = C.POINTER(C.c_ssize_t)
    CP = C.POINTER(CommStruct)
    L.fftw_b200_mpi_local_size_many_transposed.restype = S
    L.fftw_b200_mpi_local_size_many_transposed.argtypes = [I, SP, S, S, S, CP, SP, SP, SP, SP]
    for pfx in ("fftw_b200_mpi_", "fftwf_b200_mpi_"):
        f = getattr(L, pfx + "plan_many_dft")
        f.restype = P
        f.argtypes = [I, SP, S, S, S, P, P, CP, I, U]
    L.fftw_b200_mpi_local_size_1d.restype = S
    L.fftw_b200_mpi_local_size_1d.argtypes = [S, CP, I, U, SP, SP, SP, SP]
    for pfx in ("fftw_b200_mpi_", "fftwf_b200_mpi_"):
        f = getattr(L, pfx + "plan_dft_1d")
        f.restype = P
        f.argtypes = [S, P, P, CP, I, U]
        f = getattr(L, pfx + "plan_many_transpose")
        f.restype = P
        f.argtypes = [S, S, S, S, S, P, P, CP, U]
    for pfx in ("fftw_b200_mpi_", "fftwf_b200_mpi_"):
        f = getattr(L, pfx + "plan_many_r2r")
        f.restype = P
        f.argtypes = [I, SP, S, S, S, P, P, C.POINTER(CommStruct), C.POINTER(I), U]
    for pfx in ("fftw_b200_mpi_", "fftwf_b200_mpi_"):
        for nm in ("plan_many_dft_r2c", "plan_many_dft_c2r"):
            f = getattr(L, pfx + nm)
            f.restype = P
            f.argtypes = [I, SP, S, S, S, P, P, C.POINTER(CommStruct), U]
    L.fftw_b200_mpi_plan_r2r_2d.restype = P
    L.fftw_b200_mpi_plan_r2r_2d.argtypes = [S, S, P, P, C.POINTER(CommStruct), I, I, U]
    for pfx in ("fftw_b200_mpi_", "fftwf_b200_mpi_"):
        for name in ("plan_dft_r2c_2d", "plan_dft_c2r_2d"):
            f = getattr(L, pfx + name)
            f.restype = P
            f.argtypes = [S, S, P, P, C.POINTER(CommStruct), U]
    for name, extra in (("plan_dft_r2c_3d", []), ("plan_dft_c2r_3d", []), ("plan_r2r_3d", [I, I, I])):
        f = getattr(L, "fftwf_b200_mpi_" + name)
        f.restype = P
        f.argtypes = [S, S, S, P, P, C.POINTER(CommStruct)] + extra + [U]
    L.fftwf_b200_mpi_plan_r2r_2d.restype = P
    L.fftwf_b200_mpi_plan_r2r_2d.argtypes = [S, S, P, P, C.POINTER(CommStruct), I, I, U]
    for name, extra in (("plan_dft_r2c_3d", []), ("plan_dft_c2r_3d", []), ("plan_r2r_3d", [I, I, I])):
        f = getattr(L, "fftw_b200_mpi_" + name)
        f.restype = P
        f.argtypes = [S, S, S, P, P, C.POINTER(CommStruct)] + extra + [U]
    for name in ("fftw_b200_mpi_gather_wisdom", "fftw_b200_mpi_broadcast_wisdom", "fftwf_b200_mpi_gather_wisdom",
                 "fftwf_b200_mpi_broadcast_wisdom"):
        getattr(L, name).restype = None
        getattr(L, name).argtypes = [C.POINTER(CommStruct)]
    L.fftw_b200_mpi_execute.argtypes = [P]
    L.fftw_b200_mpi_destroy_plan.argtypes = [P]
    L._mpi_declared = True


def local_size_1d(lib, n0, comm, sign=B.FFTW_FORWARD, flags=0):
    """(alloc, local_ni, local_i_start, local_no, local_o_start) of fftw_b200_mpi_local_size_1d"""
    _declare_mpi(lib)
    v = [C.c_ssize_t() for _ in range(4)]
    alloc = lib.lib.fftw_b200_mpi_local_size_1d(n0, C.byref(comm), int(sign), int(flags), *[C.byref(x) for x in v])
    return (int(alloc),) + tuple(int(x.value) for x in v)


class CommPlan1D:
    """fftw_mpi_plan_dft_1d through the communicator interface (six-step over the ranks)"""

    def __init__(self, lib, n0, comm, in_ptr, out_ptr, prec="d", sign=B.FFTW_FORWARD, flags=B.FFTW_ESTIMATE, scrambled_out=False,
                 scrambled_in=False):
        _declare(lib)
        _declare_mpi(lib)
        self.L = lib.lib
        fn = getattr(self.L, ("fftwf_" if prec == "f" else "fftw_") + "b200_mpi_plan_dft_1d")
        self.plan = fn(n0, in_ptr, out_ptr, C.byref(comm), int(sign),
                       int(flags) | ((1 << 28) if scrambled_out else 0) | ((1 << 27) if scrambled_in else 0))

    def execute(self):
        self.L.fftw_b200_mpi_execute(self.plan)

    def destroy(self):
        if self.plan:
            self.L.fftw_b200_mpi_destroy_plan(self.plan)
            self.plan = None


class CommTranspose(CommPlan1D):
    """fftw_mpi_plan_many_transpose: n0 x n1 matrix of howmany-tuples of reals, rows block-distributed"""

    def __init__(self, lib, n0, n1, comm, in_ptr, out_ptr, howmany=1, prec="d", flags=B.FFTW_ESTIMATE, block0=0, block1=0):
        _declare(lib)
        _declare_mpi(lib)
        self.L = lib.lib
        fn = getattr(self.L, ("fftwf_" if prec == "f" else "fftw_") + "b200_mpi_plan_many_transpose")
        self.plan = fn(n0, n1, howmany, block0, block1, in_ptr, out_ptr, C.byref(comm), int(flags))


class CommPlanManyR2R(CommPlan1D):
    """fftw_mpi_plan_many_r2r through the communicator interface: any rank >= 2, howmany interleaved tuples"""

    def __init__(self, lib, n, comm, in_ptr, out_ptr, kinds, howmany=1, prec="d", flags=B.FFTW_ESTIMATE):
        _declare(lib)
        _declare_mpi(lib)
        self.L = lib.lib
        nn = (C.c_ssize_t * len(n))(*n)
        ks = (C.c_int * len(n))(*[B.R2R_KINDS[k] if isinstance(k, str) else int(k) for k in kinds])
        fn = getattr(self.L, ("fftwf_" if prec == "f" else "fftw_") + "b200_mpi_plan_many_r2r")
        self.plan = fn(len(n), nn, howmany, 0, 0, in_ptr, out_ptr, C.byref(comm), ks, int(flags))


class CommPlanManyReal(CommPlan1D):
    """fftw_mpi_plan_many_dft_r2c / _c2r through the communicator interface"""

    def __init__(self, lib, n, comm, in_ptr, out_ptr, what="r2c", howmany=1, prec="d", flags=B.FFTW_ESTIMATE):
        _declare(lib)
        _declare_mpi(lib)
        self.L = lib.lib
        nn = (C.c_ssize_t * len(n))(*n)
        fn = getattr(self.L, ("fftwf_" if prec == "f" else "fftw_") + "b200_mpi_plan_many_dft_" + what)
        self.plan = fn(len(n), nn, howmany, 0, 0, in_ptr, out_ptr, C.byref(comm), int(flags))


class CommPlanReal3D(CommPlan1D):
    """fftw_mpi_plan_dft_r2c_3d / _c2r_3d / fftw_mpi_plan_r2r_3d through the communicator interface (double)"""

    def __init__(self, lib, n, comm, in_ptr, out_ptr, what="r2c", kinds=None, flags=B.FFTW_ESTIMATE, prec="d"):
        _declare(lib)
        _declare_mpi(lib)
        self.L = lib.lib
        pfx = ("fftwf_" if prec == "f" else "fftw_") + "b200_mpi_"
        if what == "r2r":
            ks = [B.R2R_KINDS[k] if isinstance(k, str) else int(k) for k in kinds]
            if len(n) == 2:
                self.plan = getattr(self.L, pfx + "plan_r2r_2d")(n[0], n[1], in_ptr, out_ptr, C.byref(comm), ks[0], ks[1], int(flags))
            else:
                self.plan = getattr(self.L, pfx + "plan_r2r_3d")(n[0], n[1], n[2], in_ptr, out_ptr, C.byref(comm), ks[0], ks[1], ks[2], int(flags))
        elif len(n) == 2:
            self.plan = getattr(self.L, pfx + "plan_dft_%s_2d" % what)(n[0], n[1], in_ptr, out_ptr, C.byref(comm), int(flags))
        else:
            self.plan = getattr(self.L, pfx + "plan_dft_%s_3d" % what)(n[0], n[1], n[2], in_ptr, out_ptr, C.byref(comm), int(flags))


class CommPlan:
    """fftw_mpi_plan_many_dft through the C communicator interface: `local` is this rank's slab (device
    memory from cudaMalloc / fftw_b200_device_malloc), transformed in place unless `out` is given."""

    def __init__(self, lib, n, comm, local_ptr, out_ptr=None, howmany=1, prec="d", sign=B.FFTW_FORWARD,
                 flags=B.FFTW_ESTIMATE, transposed_out=False, transposed_in=False, block=0, tblock=0):
        _declare(lib)
        _declare_mpi(lib)
        self.L = lib.lib
        self.comm = comm
        nn = (C.c_ssize_t * len(n))(*n)
        v = [C.c_ssize_t() for _ in range(4)]
        self.alloc = int(self.L.fftw_b200_mpi_local_size_many_transposed(len(n), nn, howmany, block, tblock, C.byref(comm),
                                                                        *[C.byref(x) for x in v]))
        self.ln0, self.s0, self.ln1, self.s1 = [int(x.value) for x in v]
        fn = getattr(self.L, ("fftwf_" if prec == "f" else "fftw_") + "b200_mpi_plan_many_dft")
        fl = int(flags) | (FFTW_MPI_TRANSPOSED_OUT if transposed_out else 0) | (FFTW_MPI_TRANSPOSED_IN if transposed_in else 0)
        self.plan = fn(len(n), nn, howmany, block, tblock, local_ptr, out_ptr if out_ptr is not None else local_ptr,
                       C.byref(comm), int(sign), fl)

    def execute(self):
        self.L.fftw_b200_mpi_execute(self.plan)

    def destroy(self):
        if self.plan:
            self.L.fftw_b200_mpi_destroy_plan(self.plan)
            self.plan = None


def local_size_3d(lib, n0, n1, n2, rank, nranks):
    """(alloc_elements, local_n0, local_0_start, local_n1, local_1_start)"""
    _declare(lib)
    v = [C.c_ssize_t() for _ in range(4)]
    alloc = lib.lib.fftw_b200_dist_local_size_3d(n0, n1, n2, rank, nranks, *[C.byref(x) for x in v])
    return (int(alloc),) + tuple(int(x.value) for x in v)


def _blk(n, p):
    return (n + p - 1) // p


def _share(n, p, r):
    b = _blk(n, p)
    return max(0, min(b, n - b * r))


class SlabPlan3D:
    """Distributed c2c double transform of an n0 x n1 x n2 array.

    ``local`` is this rank's slab: a complex128 tensor (CUDA, or CPU for the
    unit tests with the emulated device layer) of at least ``alloc`` elements
    whose first local_n0*n1*n2 entries are [local_n0][n1][n2].
    """

    def __init__(self, lib, n0, n1, n2, local, group=None, sign=B.FFTW_FORWARD, flags=B.FFTW_MEASURE,
                 transposed_out=False, exchange="auto"):
        _declare(lib)
        self.lib, self.L = lib, lib.lib
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.P = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n = (n0, n1, n2)
        self.local = local
        self.transposed_out = transposed_out
        self.cuda = local.is_cuda
        if exchange == "auto":
            exchange = "peer" if (self.cuda and self.P > 1) else "collective"
        self.exchange = exchange
        alloc, self.ln0, self.s0, self.ln1, self.s1 = local_size_3d(lib, n0, n1, n2, self.rank, self.P)
        self.alloc = alloc
        assert local.numel() >= alloc and local.dtype == torch.complex128 and local.is_contiguous()
        P, r = self.P, self.rank
        b0, b1 = _blk(n0, P), _blk(n1, P)
        self._owned = []
        self._opened = []
        self.push = False
        out_arr = None
        # chunk (src s -> dst d) = [ln0(s)][ln1(d)][n2]
        if exchange == "peer":
            self.zptr = self._dev_alloc(alloc * 16)
            handles = [None] * P
            h = C.create_string_buffer(64)
            assert self.L.fftw_b200_ipc_export(self.zptr, h) == 0, "cudaIpcGetMemHandle failed"
            # natural order without a gather stage: the dim-0 pass stores its rows straight into the
            # owners' slabs, so every rank also maps every other rank's `local` (FFTW3_B200_DIST_PUSH=0
            # keeps the gather plan)
            lh, loff = None, -1
            if not transposed_out and os.environ.get("FFTW3_B200_DIST_PUSH", "1") != "0":
                loff = int(self.L.fftw_b200_ipc_offset(local.data_ptr()))
                if loff >= 0:
                    lb = C.create_string_buffer(64)
                    if self.L.fftw_b200_ipc_export(local.data_ptr() - loff, lb) == 0:
                        lh = bytes(lb.raw)
            dist.all_gather_object(handles, (bytes(h.raw), lh, loff), group=group)
            want_push = not transposed_out and all(x[1] is not None for x in handles)
            self.peer = []
            peer_local = []
            for s in range(P):
                if s == r:
                    self.peer.append(self.zptr)
                    peer_local.append(local.data_ptr())
                else:
                    ptr = self.L.fftw_b200_ipc_import(handles[s][0])
                    assert ptr, "cudaIpcOpenMemHandle failed for rank %d" % s
                    self._opened.append(ptr)
                    self.peer.append(ptr)
                    if want_push:
                        lp = self.L.fftw_b200_ipc_import(handles[s][1])
                        assert lp, "cudaIpcOpenMemHandle (slab) failed for rank %d" % s
                        self._opened.append(lp)
                        peer_local.append(lp + handles[s][2])
            if want_push:
                out_arr = (C.c_void_p * P)(*peer_local)
            # my rows start at r*b0 inside every peer's [n0][ln1(d)][n2]
            push = [self.peer[d] + 16 * (r * b0) * _share(n1, P, d) * n2 for d in range(P)]
            pull = [self.peer[s] + 16 * (r * b0) * _share(n1, P, s) * n2 for s in range(P)]
            zbuf = self.zptr
        else:
            mk = (lambda: torch.empty(alloc, dtype=torch.complex128, device=local.device))
            self.send, self.recv = mk(), mk()
            self.send_counts = [self.ln0 * _share(n1, P, d) * n2 for d in range(P)]
            self.recv_counts = [_share(n0, P, s) * self.ln1 * n2 for s in range(P)]
            so = np.concatenate([[0], np.cumsum(self.send_counts)])[:-1]
            push = [self.send.data_ptr() + 16 * int(so[d]) for d in range(P)]
            zbuf = self.recv.data_ptr()
            if not transposed_out:
                self.recv2 = mk()
                # second exchange: I send rows of zbuf = [n0][ln1][n2] for dest d: contiguous
                self.send2_counts = [_share(n0, P, d) * self.ln1 * n2 for d in range(P)]
                self.recv2_counts = [self.ln0 * _share(n1, P, s) * n2 for s in range(P)]
                ro = np.concatenate([[0], np.cumsum(self.recv2_counts)])[:-1]
                pull = [self.recv2.data_ptr() + 16 * int(ro[s]) for s in range(P)]
        VP = C.c_void_p * P
        push_arr = VP(*push)
        pull_arr = None if transposed_out else VP(*pull)
        self.plan = None
        if out_arr is not None:
            self.plan = self.L.fftw_b200_dist_plan_dft_3d_push(n0, n1, n2, r, P, local.data_ptr(), zbuf, push_arr,
                                                              out_arr, int(sign), int(flags))
            ok = torch.tensor([1 if self.plan else 0])
            if P > 1:                                   # all ranks must agree on the stage structure
                ok = ok.to(local.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0 and self.plan:
                self.L.fftw_b200_dist_destroy_plan(self.plan)
                self.plan = None
            self.push = bool(self.plan)
        if not self.plan:
            self.plan = self.L.fftw_b200_dist_plan_dft_3d(n0, n1, n2, r, P, local.data_ptr(), zbuf, push_arr,
                                                         pull_arr, int(sign), int(flags))
        assert self.plan, "fftw_b200_dist_plan_dft_3d returned NULL"
        self.nstages = self.L.fftw_b200_dist_num_stages(self.plan)
        self._token = torch.zeros(1, device=local.device)

    # ---- helpers ---------------------------------------------------------
    def _dev_alloc(self, nbytes):
        p = self.L.fftw_b200_device_malloc(nbytes)
        assert p, "device allocation of %d bytes failed" % nbytes
        self._owned.append(p)
        return p

    def _barrier(self):
        if self.P > 1:
            # stream-ordered on NCCL: every rank's kernels queued before it are complete when it ends
            dist.all_reduce(self._token, group=self.group)

    def _alltoall(self, out, out_counts, inp, in_counts):
        if self.P == 1:
            out[:in_counts[0]].copy_(inp[:in_counts[0]])
            return
        n_out, n_in = sum(out_counts), sum(in_counts)
        if dist.get_backend(self.group) == "nccl":
            dist.all_to_all_single(torch.view_as_real(out[:n_out]), torch.view_as_real(inp[:n_in]),
                                   [2 * c // 2 for c in out_counts], [2 * c // 2 for c in in_counts], group=self.group)
            return
        # gloo has no all-to-all: pairwise send/recv (mpi/transpose-pairwise.c:49-99 in spirit)
        oo = np.concatenate([[0], np.cumsum(out_counts)])
        io = np.concatenate([[0], np.cumsum(in_counts)])
        ops = []
        bufs = []
        for k in range(self.P):
            d = (self.rank + k) % self.P
            s = (self.rank - k) % self.P
            if d == self.rank:
                out[oo[s]:oo[s + 1]].copy_(inp[io[d]:io[d + 1]])
                continue
            sb = torch.view_as_real(inp[io[d]:io[d + 1]]).contiguous()
            rb = torch.empty((out_counts[s], 2), dtype=torch.float64)
            bufs.append((s, rb))
            if in_counts[d]:
                ops.append(dist.P2POp(dist.isend, sb, d, group=self.group))
            if out_counts[s]:
                ops.append(dist.P2POp(dist.irecv, rb, s, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for s, rb in bufs:
            out[oo[s]:oo[s + 1]].copy_(torch.view_as_complex(rb))

    # ---- execution -------------------------------------------------------
    def execute(self):
        L = self.L
        if self.exchange == "peer" and self.push:
            # both exchanges ride on pass stores; the closing barrier of the previous call already
            # ordered the peers' reads of their buffers before this call's first remote store
            L.fftw_b200_dist_execute_stage(self.plan, 0)
            self._barrier()                         # all blocks have landed
            L.fftw_b200_dist_execute_stage(self.plan, 1)
            self._barrier()                         # every rank's rows have landed in my slab
        elif self.exchange == "peer":
            self._barrier()                         # peers are done reading their zbuf from the last call
            L.fftw_b200_dist_execute_stage(self.plan, 0)
            self._barrier()                         # all blocks have landed
            if self.nstages == 2:
                L.fftw_b200_dist_execute_stage(self.plan, 1)
            else:
                # chunked: the gather of chunk c (peer loads on a side stream) overlaps
                # the dim-0 transforms of chunk c+1
                for c in range(L.fftw_b200_dist_num_chunks(self.plan, 1)):
                    L.fftw_b200_dist_execute_chunk(self.plan, 1, c)
                    self._barrier()                 # every rank finished its chunk c
                    L.fftw_b200_dist_execute_chunk(self.plan, 2, c)
                L.fftw_b200_dist_join(self.plan)
        else:
            L.fftw_b200_dist_execute_stage(self.plan, 0)
            self._alltoall(self.recv, self.recv_counts, self.send, self.send_counts)
            L.fftw_b200_dist_execute_stage(self.plan, 1)
            if self.nstages == 3:
                self._alltoall(self.recv2, self.recv2_counts, self.recv, self.send2_counts)
                L.fftw_b200_dist_execute_stage(self.plan, 2)

    def destroy(self):
        if self.plan:
            self.L.fftw_b200_dist_destroy_plan(self.plan)
            self.plan = None
        if self.P > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        for p in self._opened:
            self.L.fftw_b200_ipc_close(p)
        for p in self._owned:
            self.L.fftw_b200_device_free(p)
        self._opened, self._owned = [], []


def batch_share(howmany, rank, nranks):
    """(count, first) of the transforms rank `rank` owns when a batch of `howmany` independent
    transforms is block-distributed (mpi/block.c:37-50 applied to the batch index)."""
    b = _blk(howmany, nranks)
    first = min(b * rank, howmany)
    return max(0, min(b, howmany - first)), first


class ShardedBatchPlan:
    """Batched 1-D c2c over several GPUs: the batch index is block-distributed and every rank
    transforms its own share with an ordinary single-GPU plan -- independent units, NO collective
    on the data path (the reference's counterpart is the `howmany` loop of
    dft/vrank-geq1.c:54-65 spread over MPI ranks by the caller).

    ``local`` holds this rank's transforms contiguously: [count][n] complex (device or host)."""

    def __init__(self, lib, n, howmany, local_in, local_out=None, prec="d", sign=B.FFTW_FORWARD,
                 flags=B.FFTW_MEASURE, group=None):
        self.lib = lib
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.P = dist.get_world_size(group) if dist.is_initialized() else 1
        self.count, self.first = batch_share(howmany, self.rank, self.P)
        self.prec = prec
        self.plan = None
        if self.count:
            pin = local_in.data_ptr() if hasattr(local_in, "data_ptr") else local_in.ctypes.data
            out = local_in if local_out is None else local_out
            pout = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
            self.plan = lib.plan_many_dft(prec, [n], self.count, pin, None, 1, n, pout, None, 1, n, sign, flags)
            assert self.plan, "plan_many_dft returned NULL"

    def execute(self):
        if self.plan:
            self.lib.execute(self.prec, self.plan)

    def destroy(self):
        if self.plan:
            self.lib.destroy_plan(self.prec, self.plan)
            self.plan = None


def _map_peers(L, group, rank, P, ptr):
    """Every rank exports the allocation `ptr` lives in (CUDA IPC) and maps all the others;
    returns (list of P device pointers addressing each rank's `ptr`, list of mappings to close)."""
    off = int(L.fftw_b200_ipc_offset(ptr))
    assert off >= 0, "cannot locate the allocation of a device pointer"
    hb = C.create_string_buffer(64)
    assert L.fftw_b200_ipc_export(ptr - off, hb) == 0, "cudaIpcGetMemHandle failed"
    got = [None] * P
    dist.all_gather_object(got, (bytes(hb.raw), off), group=group)
    ptrs, opened = [], []
    for s in range(P):
        if s == rank:
            ptrs.append(ptr)
        else:
            base = L.fftw_b200_ipc_import(got[s][0])
            assert base, "cudaIpcOpenMemHandle failed for rank %d" % s
            opened.append(base)
            ptrs.append(base + got[s][1])
    return ptrs, opened


class SlabPlanReal3D:
    """Distributed r2c / c2r of an n0 x n1 x n2 real array (fftw_mpi_plan_dft_r2c_3d / _c2r_3d;
    C-ABI: fftw_b200_dist_plan_dft_r2c_3d / _c2r_3d in include/fftw3_b200_dist.h).

    ``real``  float64 CUDA tensor holding this rank's slab [local_n0][n1][2*(n2//2+1)] (padded rows),
    ``cplx``  complex128 CUDA tensor [local_n0][n1][n2//2+1]; pass ``real.view(torch.complex128)``
              for an in-place transform.  Peer exchange only (both exchanges are stores of FFT
              passes into peer memory).
    """

    def __init__(self, lib, n0, n1, n2, real, cplx, direction="r2c", group=None, flags=B.FFTW_MEASURE):
        _declare(lib)
        self.lib, self.L = lib, lib.lib
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.P = dist.get_world_size(group) if dist.is_initialized() else 1
        self.direction = direction
        P, r = self.P, self.rank
        h = n2 // 2 + 1
        b0, b1 = _blk(n0, P), _blk(n1, P)
        self.ln0, self.s0 = _share(n0, P, r), min(b0 * r, n0)
        assert real.is_cuda and cplx.is_cuda and real.dtype == torch.float64 and cplx.dtype == torch.complex128
        assert real.numel() >= self.ln0 * n1 * 2 * h and cplx.numel() >= self.ln0 * n1 * h
        self.real, self.cplx = real, cplx
        self._owned, self._opened = [], []
        zbytes = 16 * max(P * b0 * b1 * h, 1)
        self.zptr = self.L.fftw_b200_device_malloc(zbytes)
        assert self.zptr, "device allocation of %d bytes failed" % zbytes
        self._owned.append(self.zptr)
        if P > 1:
            zpeers, o1 = _map_peers(self.L, group, r, P, self.zptr)
            cpeers, o2 = _map_peers(self.L, group, r, P, cplx.data_ptr())
            self._opened += o1 + o2
        else:
            zpeers, cpeers = [self.zptr], [cplx.data_ptr()]
        VP = C.c_void_p * P
        push = VP(*[zpeers[d] + 16 * (r * b0) * b1 * h for d in range(P)])
        out = VP(*cpeers)
        fn = self.L.fftw_b200_dist_plan_dft_r2c_3d if direction == "r2c" else self.L.fftw_b200_dist_plan_dft_c2r_3d
        a, b = (real.data_ptr(), cplx.data_ptr()) if direction == "r2c" else (cplx.data_ptr(), real.data_ptr())
        self.plan = fn(n0, n1, n2, r, P, a, b, self.zptr, push, out, int(flags))
        assert self.plan, "distributed %s plan returned NULL" % direction
        self._token = torch.zeros(1, device=real.device)

    def _barrier(self):
        if self.P > 1:
            dist.all_reduce(self._token, group=self.group)

    def execute(self):
        L = self.L
        L.fftw_b200_dist_execute_stage(self.plan, 0)
        self._barrier()                     # exchange buffers complete
        L.fftw_b200_dist_execute_stage(self.plan, 1)
        self._barrier()                     # every rank's rows have landed in my complex slab
        if self.direction == "c2r":
            L.fftw_b200_dist_execute_stage(self.plan, 2)

    def destroy(self):
        if self.plan:
            self.L.fftw_b200_dist_destroy_plan(self.plan)
            self.plan = None
        if self.P > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        for p in self._opened:
            self.L.fftw_b200_ipc_close(p)
        for p in self._owned:
            self.L.fftw_b200_device_free(p)
        self._opened, self._owned = [], []


class SlabPlanR2R3D:
    """Distributed r2r of an n0 x n1 x n2 real array, kinds[i] along dimension i
    (fftw_mpi_plan_r2r_3d; C-ABI fftw_b200_dist_plan_r2r_3d).  ``local`` = float64 CUDA tensor
    [local_n0][n1][n2], transformed in place.  The two global transposes are gather copies with
    peer loads; a barrier precedes every stage."""

    def __init__(self, lib, n0, n1, n2, local, kinds, group=None, flags=B.FFTW_MEASURE):
        _declare(lib)
        self.lib, self.L = lib, lib.lib
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.P = dist.get_world_size(group) if dist.is_initialized() else 1
        P, r = self.P, self.rank
        b0, b1 = _blk(n0, P), _blk(n1, P)
        self.ln0 = _share(n0, P, r)
        assert local.is_cuda and local.dtype == torch.float64 and local.numel() >= self.ln0 * n1 * n2
        self.local = local
        self._owned, self._opened = [], []
        self.zptr = self.L.fftw_b200_device_malloc(8 * max(n0 * b1 * n2, 1))
        assert self.zptr, "device allocation failed"
        self._owned.append(self.zptr)
        if P > 1:
            zpeers, o1 = _map_peers(self.L, group, r, P, self.zptr)
            lpeers, o2 = _map_peers(self.L, group, r, P, local.data_ptr())
            self._opened += o1 + o2
        else:
            zpeers, lpeers = [self.zptr], [local.data_ptr()]
        VP = C.c_void_p * P
        ks = (C.c_int * 3)(*[B.R2R_KINDS[k] if isinstance(k, str) else int(k) for k in kinds])
        self.plan = self.L.fftw_b200_dist_plan_r2r_3d(n0, n1, n2, r, P, local.data_ptr(), self.zptr, VP(*lpeers),
                                                     VP(*zpeers), ks, int(flags))
        assert self.plan, "distributed r2r plan returned NULL"
        self._token = torch.zeros(1, device=local.device)

    def execute(self):
        for st in range(3):
            if self.P > 1:
                dist.all_reduce(self._token, group=self.group)      # barrier before every stage
            self.L.fftw_b200_dist_execute_stage(self.plan, st)

    def destroy(self):
        if self.plan:
            self.L.fftw_b200_dist_destroy_plan(self.plan)
            self.plan = None
        if self.P > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        for p in self._opened:
            self.L.fftw_b200_ipc_close(p)
        for p in self._owned:
            self.L.fftw_b200_device_free(p)
        self._opened, self._owned = [], []


# --------------------------------------------------------------------------
# bench.py leg for N > 1 (one rank per GPU, launched by torchrun)
# --------------------------------------------------------------------------
def _nvlink_tx_kib(index):
    """Sum of the NVLink transmit counters of one GPU (`nvidia-smi nvlink -gt d`), KiB; None if unavailable."""
    import re
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True,
                             timeout=20).stdout
    except (OSError, subprocess.SubprocessError):
        return None
    vals = [int(v) for v in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out)]
    if not vals:
        _nvlink_tx_kib.raw = out[:400]
    return sum(vals) if vals else None


def _self_check_slab(lib, n, world, rank, local, plan, flags, exchange, impulse_expected):
    """Correctness of the distributed plan that was timed, outside the timed region: a unit impulse
    at a non-trivial global index against the closed-form phases (every rank samples its own output
    planes), then forward + backward against the input (relative L2 over all ranks)."""
    import math
    dev = local.device
    alloc, ln0, s0, ln1, s1 = local_size_3d(lib, n, n, n, rank, world)
    nel = ln0 * n * n
    slab = local[:nel].view(ln0, n, n) if ln0 else None
    j = (n // 3 + 1, n // 5 + 2, n // 7 + 3)
    local.zero_()
    if ln0 and s0 <= j[0] < s0 + ln0:
        slab[j[0] - s0, j[1], j[2]] = 1.0
    torch.cuda.synchronize()
    dist.barrier()
    plan.execute()
    torch.cuda.synchronize()
    dist.barrier()
    err = torch.zeros(1, dtype=torch.float64, device=dev)
    if ln0:
        rng = np.random.default_rng(1000 + rank)
        ks = np.stack([rng.integers(s0, s0 + ln0, 256), rng.integers(0, n, 256), rng.integers(0, n, 256)], axis=1)
        kt = torch.from_numpy(ks).to(dev)
        got = slab[kt[:, 0] - s0, kt[:, 1], kt[:, 2]].cpu().numpy()
        err[0] = float(np.abs(got - impulse_expected(n, j, ks)).max())
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    imp = float(err.item())
    # round trip
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    lr = torch.view_as_real(local)
    lr.copy_(torch.rand(lr.shape, dtype=torch.float64, device=dev, generator=g) - 0.5)
    keep = local[:nel].clone()
    back = SlabPlan3D(lib, n, n, n, local, sign=B.FFTW_BACKWARD, flags=flags, transposed_out=False, exchange=exchange)
    torch.cuda.synchronize()
    dist.barrier()
    plan.execute()
    back.execute()
    torch.cuda.synchronize()
    dist.barrier()
    back.destroy()
    acc = torch.zeros(2, dtype=torch.float64, device=dev)
    if ln0:
        d = torch.view_as_real(local[:nel] * (1.0 / float(n) ** 3) - keep)
        acc[0] = (d ** 2).sum()
        acc[1] = (torch.view_as_real(keep) ** 2).sum()
    dist.all_reduce(acc)
    rt = float((acc[0] / acc[1]).sqrt().item())
    lg = 3 * math.log2(n)
    return {"impulse_max_err": imp, "impulse_at": list(j), "impulse_samples": 256 * world, "roundtrip_rel_l2": rt,
            "tolerance": {"impulse": 1e-13 * lg, "roundtrip": 3.0 * 2.0 ** -52 * lg},
            "ok": bool(imp <= 1e-13 * lg and rt <= 3.0 * 2.0 ** -52 * lg)}


def bench_slab_3d(args, lib, n, world, rank, local_rank, ClockSampler, flops_c2c, cpu_reference_run,
                  impulse_expected=None):
    import json
    import os
    import sys
    import time

    dev = torch.device("cuda", local_rank)
    _declare(lib)
    alloc, ln0, s0, ln1, s1 = local_size_3d(lib, n, n, n, rank, world)
    local = torch.empty(alloc, dtype=torch.complex128, device=dev)
    g = torch.Generator(device=dev).manual_seed(rank)
    lr = torch.view_as_real(local)
    lr.copy_(torch.rand(lr.shape, dtype=torch.float64, device=dev, generator=g) - 0.5)
    flags = B.FFTW_ESTIMATE if args.estimate else B.FFTW_MEASURE
    exchange = os.environ.get("FFTW3_B200_EXCHANGE", "peer")
    t0 = time.perf_counter()
    plan = SlabPlan3D(lib, n, n, n, local, flags=flags, transposed_out=False, exchange=exchange)
    plan_s = time.perf_counter() - t0
    lib.lib.fftw_b200_set_async(1)

    def timed(pl, steps, warmup):
        for _ in range(warmup):
            pl.execute()
            lr.mul_(1.0 / n ** 1.5)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.launch_count()
        e0.record()
        for _ in range(steps):
            pl.execute()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # slowest rank defines the step
        return float(ms.item()), lib.launch_count() - l0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    tx0 = _nvlink_tx_kib(local_rank) if rank == 0 else None
    ms, launches = timed(plan, args.steps, max(3, args.warmup))
    tx1 = _nvlink_tx_kib(local_rank) if rank == 0 else None
    clocks = sampler.stop() if rank == 0 else None

    # per-stage times of one transform (diagnostic, outside the timed region): stage 0 = Y + X (first
    # exchange rides on X's stores), stage 1 = Z (second exchange rides on its stores)
    stage_ms = None
    if plan.exchange == "peer" and plan.push:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        acc = [0.0, 0.0]
        for _ in range(3):
            torch.cuda.synchronize()
            dist.barrier()
            evs[0].record()
            lib.lib.fftw_b200_dist_execute_stage(plan.plan, 0)
            plan._barrier()
            evs[1].record()
            lib.lib.fftw_b200_dist_execute_stage(plan.plan, 1)
            plan._barrier()
            evs[2].record()
            torch.cuda.synchronize()
            acc[0] += evs[0].elapsed_time(evs[1]) / 3
            acc[1] += evs[1].elapsed_time(evs[2]) / 3
        st = torch.tensor(acc, device=dev)
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        stage_ms = [float(v) for v in st.tolist()]

    # the same transform with FFTW_MPI_TRANSPOSED_OUT semantics (one exchange instead of two)
    plan_t = SlabPlan3D(lib, n, n, n, local, flags=flags, transposed_out=True, exchange=exchange)
    ms_t, _ = timed(plan_t, max(2, args.steps // 2), 2)

    # e2e: host slabs, H2D + transform + D2H every step, through the public call
    e2e = None
    if not args.no_e2e:
        nel = ln0 * n * n
        host = torch.empty(max(nel, 1), dtype=torch.complex128).pin_memory()
        host.fill_(0.25)
        steps = max(1, min(args.steps, 2))
        for it in range(steps + 1):
            if it == 1:
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
            local[:nel].copy_(host[:nel], non_blocking=True)
            plan.execute()
            host[:nel].copy_(local[:nel], non_blocking=True)
            torch.cuda.synchronize()
        dist.barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / steps], device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": flops_c2c((n, n, n)) / float(dt.item()) / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": 16 * nel * world, "d2h_bytes_per_step": 16 * nel * world,
               "ms_per_step": float(dt.item()) * 1e3, "steps": steps,
               "api": "fftw3_b200.dist.SlabPlan3D.execute on pinned host slabs (one per rank)"}
    lib.lib.fftw_b200_set_async(0)
    pushed = plan.push
    by_copy = bool(lib.lib.fftw_b200_dist_exchange_by_copy(plan.plan))
    part_sms = int(lib.lib.fftw_b200_dist_partition_sms(plan.plan))
    check = None
    if impulse_expected is not None and not getattr(args, "no_check", False):
        check = _self_check_slab(lib, n, world, rank, local, plan, flags, exchange, impulse_expected)
    plan.destroy()
    plan_t.destroy()
    if rank != 0:
        return
    cpu = None if args.no_cpu else cpu_reference_run(1, 0, sample_n=min(256, n))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                            "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    slab_bytes = 16 * ln0 * n * n
    passes = 3 if pushed else 4                 # Y, X(+scatter), Z(+row push)  |  Y, X(+scatter), Z, gather
    achieved = passes * 2 * slab_bytes / (ms * 1e-3) / 1e9
    nv_bytes = 2 * slab_bytes * (world - 1) / world      # two exchanges, sent per GPU
    line = {
        "metric": "GFLOP/s (5N log2 N), 3-D c2c double", "value": flops_c2c((n, n, n)) / (ms * 1e-3) / 1e9,
        "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%d^3 c2c double in place, forward, slab-decomposed over %d GPUs, natural-order "
                               "output (two exchanges)" % (n, world),
                   "exchange": exchange + ((", first exchange = copy engines under the next chunk's transforms, second "
                                            "fused into the dim-0 pass's stores (no gather stage)" if by_copy else
                                            ", both exchanges fused into pass stores (no gather stage)" +
                                            ("; stage 0: scatter pass on %d SMs of its own, Y pass of the next chunk on "
                                             "the rest (green contexts)" % part_sms if part_sms else "")) if pushed
                                           else ", second exchange = gather stage"), "l2": "slabs are larger than L2, no flush needed",
                   "planner": "FFTW_ESTIMATE" if args.estimate else "FFTW_MEASURE", "plan_seconds": plan_s,
                   "transposed_out_ms_per_step": ms_t,
                   "transposed_out_gflops": flops_c2c((n, n, n)) / (ms_t * 1e-3) / 1e9},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "passes": passes, "per": "GPU",
                     "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback (of fallback)",
                     "nvlink": {"sent_bytes_per_gpu_per_step": nv_bytes,
                                "if_serialised_gbs": nv_bytes / (ms * 1e-3) / 1e9, "peak_gbs": 900.0,
                                "counter_tx_bytes_per_step": (None if tx0 is None or tx1 is None else
                                                              (tx1 - tx0) * 1024.0 / (args.steps + max(3, args.warmup))),
                                "counter_raw": getattr(_nvlink_tx_kib, "raw", None),
                                "stage_ms": stage_ms,
                                "stage_exchange_gbs": (None if not stage_ms else
                                                       [nv_bytes / 2 / (t * 1e-3) / 1e9 for t in stage_ms]),
                                "note": "each of the two stages moves sent_bytes/2 over NVLink; "
                                        "stage_exchange_gbs = that / the stage's time (a lower bound of the link rate "
                                        "while the store pass runs)"}},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "check": check,
        "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
    }
    print(json.dumps(line), flush=True)
    if check is not None and not check["ok"]:
        sys.exit("bench.py: self check FAILED: %r" % (check,))
