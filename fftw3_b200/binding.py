"""ctypes binding of the fftw3_b200 C-ABI (include/fftw3.h).

Pointers are passed as plain integers, so callers can hand in numpy host arrays
(``arr.ctypes.data``) or CUDA device memory (``tensor.data_ptr()``) alike --
exactly what a C caller of the FFTW API would do.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))

FFTW_FORWARD, FFTW_BACKWARD = -1, 1
FFTW_MEASURE = 0
FFTW_DESTROY_INPUT = 1 << 0
FFTW_UNALIGNED = 1 << 1
FFTW_EXHAUSTIVE = 1 << 3
FFTW_PRESERVE_INPUT = 1 << 4
FFTW_PATIENT = 1 << 5
FFTW_ESTIMATE = 1 << 6
FFTW_WISDOM_ONLY = 1 << 21

R2R_KINDS = {
    "R2HC": 0, "HC2R": 1, "DHT": 2,
    "REDFT00": 3, "REDFT01": 4, "REDFT10": 5, "REDFT11": 6,
    "RODFT00": 7, "RODFT01": 8, "RODFT10": 9, "RODFT11": 10,
}


class Iodim(C.Structure):
    _fields_ = [("n", C.c_int), ("is_", C.c_int), ("os", C.c_int)]


class Iodim64(C.Structure):
    _fields_ = [("n", C.c_ssize_t), ("is_", C.c_ssize_t), ("os", C.c_ssize_t)]


def default_library_path():
    return os.path.join(HERE, "lib", "libfftw3_b200.so")


def build_library():
    """Compile the product library in-tree (nvcc, sm_100a)."""
    subprocess.run(["make", "-s", "-j8", "-C", os.path.join(HERE, "csrc")], check=True)
    return default_library_path()


def _ints(v):
    if v is None:
        return None
    return (C.c_int * len(v))(*[int(x) for x in v])


class Lib:
    """One loaded copy of the library; ``prec`` selects fftw_ ('d') or fftwf_ ('f')."""

    def __init__(self, path=None):
        self.path = path or default_library_path()
        if not os.path.exists(self.path):
            raise FileNotFoundError(
                "%s is missing: build it with `make -C fftw3_b200/csrc` (needs nvcc); "
                "there is no CPU fallback" % self.path)
        self.lib = C.CDLL(self.path)
        self._declare()

    # ---- declarations -------------------------------------------------
    def _declare(self):
        P, I, U = C.c_void_p, C.c_int, C.c_uint
        IP = C.POINTER(C.c_int)
        for pfx in ("fftw_", "fftwf_"):
            def f(name, res, args, pfx=pfx):
                fn = getattr(self.lib, pfx + name)
                fn.restype = res
                fn.argtypes = args
            f("plan_dft_1d", P, [I, P, P, I, U])
            f("plan_dft_2d", P, [I, I, P, P, I, U])
            f("plan_dft_3d", P, [I, I, I, P, P, I, U])
            f("plan_dft", P, [I, IP, P, P, I, U])
            f("plan_many_dft", P, [I, IP, I, P, IP, I, I, P, IP, I, I, I, U])
            f("plan_guru_dft", P, [I, P, I, P, P, P, I, U])
            f("plan_guru64_dft", P, [I, P, I, P, P, P, I, U])
            f("plan_guru_split_dft", P, [I, P, I, P, P, P, P, P, U])
            f("plan_guru64_split_dft", P, [I, P, I, P, P, P, P, P, U])
            for nm in ("r2c", "c2r"):
                f("plan_dft_%s_1d" % nm, P, [I, P, P, U])
                f("plan_dft_%s_2d" % nm, P, [I, I, P, P, U])
                f("plan_dft_%s_3d" % nm, P, [I, I, I, P, P, U])
                f("plan_dft_%s" % nm, P, [I, IP, P, P, U])
                f("plan_many_dft_%s" % nm, P, [I, IP, I, P, IP, I, I, P, IP, I, I, U])
                f("plan_guru_dft_%s" % nm, P, [I, P, I, P, P, P, U])
                f("plan_guru64_dft_%s" % nm, P, [I, P, I, P, P, P, U])
                f("plan_guru_split_dft_%s" % nm, P, [I, P, I, P, P, P, P, U])
                f("plan_guru64_split_dft_%s" % nm, P, [I, P, I, P, P, P, P, U])
            f("plan_r2r_1d", P, [I, P, P, I, U])
            f("plan_r2r_2d", P, [I, I, P, P, I, I, U])
            f("plan_r2r_3d", P, [I, I, I, P, P, I, I, I, U])
            f("plan_r2r", P, [I, IP, P, P, IP, U])
            f("plan_many_r2r", P, [I, IP, I, P, IP, I, I, P, IP, I, I, IP, U])
            f("plan_guru_r2r", P, [I, P, I, P, P, P, IP, U])
            f("plan_guru64_r2r", P, [I, P, I, P, P, P, IP, U])
            f("execute", None, [P])
            f("execute_dft", None, [P, P, P])
            f("execute_split_dft", None, [P, P, P, P, P])
            f("execute_dft_r2c", None, [P, P, P])
            f("execute_dft_c2r", None, [P, P, P])
            f("execute_split_dft_r2c", None, [P, P, P, P])
            f("execute_split_dft_c2r", None, [P, P, P, P])
            f("execute_r2r", None, [P, P, P])
            f("copy_plan", P, [P])
            f("destroy_plan", None, [P])
            f("cleanup", None, [])
            f("forget_wisdom", None, [])
            f("set_timelimit", None, [C.c_double])
            f("init_threads", I, [])
            f("plan_with_nthreads", None, [I])
            f("planner_nthreads", I, [])
            f("cleanup_threads", None, [])
            f("make_planner_thread_safe", None, [])
            f("export_wisdom_to_filename", I, [C.c_char_p])
            f("export_wisdom_to_string", P, [])
            f("import_wisdom_from_filename", I, [C.c_char_p])
            f("import_wisdom_from_string", I, [C.c_char_p])
            f("import_system_wisdom", I, [])
            f("sprint_plan", P, [P])
            f("print_plan", None, [P])
            f("flops", None, [P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)])
            f("estimate_cost", C.c_double, [P])
            f("cost", C.c_double, [P])
            f("malloc", P, [C.c_size_t])
            f("alloc_real", P, [C.c_size_t])
            f("alloc_complex", P, [C.c_size_t])
            f("free", None, [P])
            f("alignment_of", I, [P])
        self.lib.fftw_b200_set_stream.argtypes = [P]
        self.lib.fftw_b200_set_async.argtypes = [I]
        self.lib.fftw_b200_launch_count.restype = C.c_ulonglong
        self.lib.fftw_b200_device_name.restype = C.c_char_p
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [P]

    def fn(self, prec, name):
        return getattr(self.lib, ("fftwf_" if prec == "f" else "fftw_") + name)

    # ---- conveniences used by the tests / bench -------------------------
    def plan_many_dft(self, prec, n, howmany, inp, inembed, istride, idist, out, onembed, ostride,
                      odist, sign, flags):
        return self.fn(prec, "plan_many_dft")(len(n), _ints(n), howmany, inp, _ints(inembed), istride,
                                              idist, out, _ints(onembed), ostride, odist, sign, flags)

    def plan_many_dft_r2c(self, prec, n, howmany, inp, inembed, istride, idist, out, onembed, ostride,
                          odist, flags):
        return self.fn(prec, "plan_many_dft_r2c")(len(n), _ints(n), howmany, inp, _ints(inembed),
                                                  istride, idist, out, _ints(onembed), ostride, odist, flags)

    def plan_many_dft_c2r(self, prec, n, howmany, inp, inembed, istride, idist, out, onembed, ostride,
                          odist, flags):
        return self.fn(prec, "plan_many_dft_c2r")(len(n), _ints(n), howmany, inp, _ints(inembed),
                                                  istride, idist, out, _ints(onembed), ostride, odist, flags)

    def plan_many_r2r(self, prec, n, howmany, inp, inembed, istride, idist, out, onembed, ostride,
                      odist, kinds, flags):
        ks = [R2R_KINDS[k] if isinstance(k, str) else int(k) for k in kinds]
        return self.fn(prec, "plan_many_r2r")(len(n), _ints(n), howmany, inp, _ints(inembed), istride,
                                              idist, out, _ints(onembed), ostride, odist, _ints(ks), flags)

    def plan_guru_dft(self, prec, dims, howmany_dims, inp, out, sign, flags, wide=False):
        T = Iodim64 if wide else Iodim
        d = (T * max(1, len(dims)))(*[T(*x) for x in dims])
        h = (T * max(1, len(howmany_dims)))(*[T(*x) for x in howmany_dims])
        name = "plan_guru64_dft" if wide else "plan_guru_dft"
        return self.fn(prec, name)(len(dims), C.cast(d, C.c_void_p), len(howmany_dims),
                                   C.cast(h, C.c_void_p), inp, out, sign, flags)

    def plan_guru_split_dft(self, prec, dims, howmany_dims, ri, ii, ro, io, flags):
        d = (Iodim * max(1, len(dims)))(*[Iodim(*x) for x in dims])
        h = (Iodim * max(1, len(howmany_dims)))(*[Iodim(*x) for x in howmany_dims])
        return self.fn(prec, "plan_guru_split_dft")(len(dims), C.cast(d, C.c_void_p), len(howmany_dims),
                                                    C.cast(h, C.c_void_p), ri, ii, ro, io, flags)

    def execute(self, prec, plan):
        self.fn(prec, "execute")(plan)

    def destroy_plan(self, prec, plan):
        self.fn(prec, "destroy_plan")(plan)

    def sprint_plan(self, prec, plan):
        p = self.fn(prec, "sprint_plan")(plan)
        s = C.string_at(p).decode()
        self.libc.free(p)
        return s

    def export_wisdom_to_string(self, prec):
        p = self.fn(prec, "export_wisdom_to_string")()
        s = C.string_at(p).decode()
        self.libc.free(p)
        return s

    def launch_count(self):
        return int(self.lib.fftw_b200_launch_count())

    def device_name(self):
        return self.lib.fftw_b200_device_name().decode()


_DEFAULT = None


def load(path=None):
    """Load (once) and return the product library binding."""
    global _DEFAULT
    if path is not None:
        return Lib(path)
    if _DEFAULT is None:
        _DEFAULT = Lib()
    return _DEFAULT
