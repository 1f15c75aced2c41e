"""The product C-ABI library loads on a GPU-less box and exports every symbol
include/fftw3.h declares; without a device plan creation fails cleanly (NULL),
it never computes on the CPU."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    inc = open(os.path.join(ROOT, "include", "fftw3_api.inc")).read()
    names = set(re.findall(r"FFTW3_NS\((\w+)\)\s*\(", inc))
    names -= {"plan_s"}
    names |= {"version", "cc", "codelet_optim"}
    ext = re.findall(r"\b(fftw_b200_\w+)\s*\(", open(os.path.join(ROOT, "include", "fftw3.h")).read())
    return sorted(names), sorted(set(ext))


def test_exports_every_declared_symbol():
    from fftw3_b200 import binding
    path = binding.default_library_path()
    if not os.path.exists(path):
        binding.build_library()
    lib = C.CDLL(path)
    names, ext = declared_symbols()
    assert len(names) >= 76      # 73 functions + version, cc, codelet_optim
    missing = [p + n for p in ("fftw_", "fftwf_") for n in names if not hasattr(lib, p + n)]
    missing += [e for e in ext if not hasattr(lib, e)]
    assert not missing, missing
    for alias in ("libfftw3.so.3", "libfftw3f.so.3"):
        assert os.path.exists(os.path.join(os.path.dirname(path), alias))


def test_exports_every_symbol_of_the_distributed_header_and_binding():
    """Everything include/fftw3_b200_dist.h declares is exported, and the ctypes prototypes of fftw3_b200/dist.py
    (slab plans and the communicator interface) bind against the product library -- no compute call."""
    from fftw3_b200 import binding, dist
    path = binding.default_library_path()
    lib = C.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "fftw3_b200_dist.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(fftwf?_b200_\w+)\s*\(", hdr)))
    assert len(names) >= 40, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing

    class _L:            # what dist._declare* need of binding.Lib
        pass
    shell = _L()
    shell.lib = lib
    dist._declare(shell)
    dist._declare_mpi(shell)


def test_exports_nothing_but_the_fftw_namespace():
    """A libfftw3.so.3 stand-in must not leak internals (C++ kernel templates, b2_* / b2d_* helpers):
    the linker version script fftw3_b200/csrc/exports.map keeps fftw_* and fftwf_* only."""
    import subprocess
    from fftw3_b200 import binding
    out = subprocess.run(["nm", "-D", "--defined-only", binding.default_library_path()], capture_output=True, text=True,
                         check=True).stdout
    names = [ln.split()[-1] for ln in out.splitlines() if ln.strip()]
    leaked = [n for n in names if not n.startswith(("fftw_", "fftwf_", "dfftw_", "sfftw_"))]
    assert len(names) > 150 and not leaked, leaked[:10]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        return                       # on the GPU box the parity tests cover the real path
    from fftw3_b200 import binding as B
    lib = B.load()
    x = np.zeros(64, np.complex128)
    p = lib.fn("d", "plan_dft_1d")(64, x.ctypes.data, x.ctypes.data, -1, B.FFTW_ESTIMATE)
    assert not p, "a plan was created without a CUDA device: that would be a CPU fallback"


def test_product_library_does_not_reference_oracle():
    """The shipped library must not link or embed anything from oracle/."""
    from fftw3_b200 import binding
    data = open(binding.default_library_path(), "rb").read()
    assert b"oracle_dft" not in data and b"liboracle" not in data and b"_ref.so" not in data


def test_wisdom_tool_command_line():
    """tools/fftw_wisdom.c (the reference's tools/fftw-wisdom.c:73-131 command line): help text,
    size-syntax errors, and -- with no GPU -- a loud failure instead of empty wisdom."""
    import subprocess
    libdir = os.path.join(ROOT, "fftw3_b200", "lib")
    for tool in ("fftw-wisdom", "fftwf-wisdom"):
        exe = os.path.join(libdir, tool)
        assert os.path.exists(exe), "run __graft_entry__.build()"
        r = subprocess.run([exe, "--help"], capture_output=True, text=True)
        assert r.returncode == 0 and "Size syntax" in r.stdout and "--canonical" in r.stdout
        r = subprocess.run([exe, "-V"], capture_output=True, text=True)
        assert r.returncode == 0 and "FFTW 3.3" in r.stdout
        r = subprocess.run([exe, "-e", "cof12q"], capture_output=True, text=True)
        assert r.returncode != 0 and "cannot parse" in r.stderr
        r = subprocess.run([exe, "--bogus"], capture_output=True, text=True)
        assert r.returncode != 0
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([os.path.join(libdir, "fftw-wisdom"), "-e", "cof64", "ki10e10x8e01v3", "rib16x6"],
                           capture_output=True, text=True)
        assert r.returncode != 0 and r.stderr.count("could not plan") == 3


def _build_example(libdir, libname, out):
    import subprocess
    src = os.path.join(ROOT, "examples", "fftw_tutorial.c")
    subprocess.run(["gcc", "-O1", "-Wall", src, "-I" + os.path.join(ROOT, "include"), "-L" + libdir, "-l" + libname, "-lm",
                    "-Wl,-rpath," + libdir, "-o", out], check=True)
    return subprocess.run([out], capture_output=True, text=True, timeout=300)


def test_plain_fftw_program_links_and_runs(emu_lib, tmp_path):
    """examples/fftw_tutorial.c uses nothing but fftw3.h (doc/tutorial.texi style).  Its logic is
    checked here against the host layer on the emulated device; linked against the product
    library on a box without a GPU it must fail loudly at planning (no CPU fallback)."""
    r = _build_example(os.path.join(ROOT, "tests", "_emu"), "fftw3_b200_emu", str(tmp_path / "tut_emu"))
    assert r.returncode == 0 and r.stdout.count(" ok") == 7 and "FAILED" not in r.stdout, r.stdout + r.stderr
    import torch
    if not torch.cuda.is_available():
        r = _build_example(os.path.join(ROOT, "fftw3_b200", "lib"), "fftw3_b200", str(tmp_path / "tut_gpu"))
        assert r.returncode == 2 and "planning failed" in r.stderr
