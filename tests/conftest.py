import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lf-vio_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # build the product library and the oracle's native helpers if they are not there yet
    import subprocess

    if not os.path.exists(os.path.join(ROOT, "lf-vio_b200", "liblfvio.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "lf-vio_b200"), "-j8"], check=True)
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8"], check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def lfvio_mod():
    """The ctypes view of liblfvio.so.  Built by `make -C lf-vio_b200` / __graft_entry__.build()."""
    import lfvio

    return lfvio


@pytest.fixture(scope="session")
def gpu(lfvio_mod):
    if lfvio_mod.lib.lfvio_device_count() < 1:
        pytest.fail("test is marked gpu but liblfvio sees no CUDA device (no CPU fallback exists)")
    return 0
