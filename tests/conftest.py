import os
import subprocess
import sys
from pathlib import Path

import pytest

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "emul: runs the kernel sources on the test-only CPU SIMT emulator")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    oracle_lib.lib()  # builds oracle/libfdeflate_oracle.so with gcc if needed
    return oracle_lib


@pytest.fixture(scope="session")
def emul_lib():
    """The kernel sources compiled for the CPU SIMT emulator (tests/emul). Test infrastructure only."""
    from fdeflate_b200 import NativeLib

    so = HERE / "emul" / "libfdb_emul.so"
    subprocess.run(["make", "-C", str(HERE / "emul")], check=True, capture_output=True)
    return NativeLib(so)


@pytest.fixture(scope="session")
def emul_ctx(emul_lib):
    from fdeflate_b200 import Context

    return Context(0, emul_lib)


@pytest.fixture(scope="session")
def gpu_ctx():
    """The product: libfdeflate_b200.so on cuda:0 through the C ABI. Fails loudly if it is missing."""
    from fdeflate_b200 import Context

    return Context(0)
