import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def emu_lib():
    """The csrc kernels compiled against the SIMT emulation shim (test infrastructure only)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu

    return build_emu.build()


@pytest.fixture()
def emu(emu_lib):
    """Install the emulator build as the process-wide library for one test."""
    from wavebreaking_b200 import _lib

    prev = _lib._LIB
    lib = _lib.use_library(emu_lib, "cpu")
    yield lib
    _lib._LIB = prev


@pytest.fixture()
def gpu():
    """The real CUDA library; fails loudly if it is missing."""
    from wavebreaking_b200 import _lib

    _lib._LIB = None
    lib = _lib.get()
    assert lib.is_cuda
    return lib
