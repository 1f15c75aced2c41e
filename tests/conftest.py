import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def emu_lib():
    """The REAL host layer linked against the unit-test double of the device shim
    (tests/emu/emu_shim.cpp).  Test infrastructure only; never the product."""
    from fftw3_b200 import binding
    override = os.environ.get("FFTW3_B200_EMU_LIB")       # e.g. a sanitizer build of the same sources
    if override:
        return binding.Lib(override)
    subprocess.run(["sh", os.path.join(ROOT, "tests", "emu", "build_emu.sh")], check=True)
    return binding.Lib(os.path.join(ROOT, "tests", "_emu", "libfftw3_b200_emu.so"))


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on a real GPU.  Fails loudly if the CUDA library is
    missing or no device is usable -- there is no fallback to hide behind."""
    from fftw3_b200 import binding
    path = binding.default_library_path()
    assert os.path.exists(path), "product library not built: run python -c 'import __graft_entry__ as g; g.build()'"
    lib = binding.Lib(path)
    name = lib.device_name()
    assert name, "no CUDA device visible to libfftw3_b200.so"
    return lib


@pytest.fixture(scope="session", params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def host_lib(request):
    """The host-logic suites (API contract checklist, every entry point, guru fuzz) run twice: on the
    emulated device layer here (`-m "not gpu"`) and on the product library on the B200 (`-m gpu`), where
    they also exercise shim.cu's dispatch, the specialised kernels, alignment fallbacks and host staging."""
    return request.getfixturevalue("emu_lib" if request.param == "emu" else "gpu_lib")
