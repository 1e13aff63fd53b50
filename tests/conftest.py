"""Shared fixtures. `-m gpu` tests need a B200 and call through the C ABI of libc5gpu.so;
everything else runs on CPU (oracle vs golden vectors, host logic, the ABI surface, and the
device functions' logic compiled for the host in tests/hostsim/)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Native artefacts (prebuilt ones are reused; make only rebuilds what changed)."""
    import __graft_entry__ as entry
    have = all(os.path.exists(os.path.join(ROOT, p)) for p in (
        "course5_b200/libc5gpu.so", "tests/hostsim/libc5hostsim.so", "oracle/libc5oracle.so"))
    if not have:
        entry.build()
    return True


@pytest.fixture(scope="session")
def port(built):
    from oracle import refbind
    return refbind.Port()


@pytest.fixture(scope="session")
def ref(built):
    """The unmodified reference (oracle/_ref). Present wherever it was built; skipped otherwise."""
    from oracle import refbind
    if not os.path.exists(refbind.REF_SO):
        pytest.skip("oracle/_ref/libc5ref.so not built (needs /root/reference)")
    return refbind.Ref()


@pytest.fixture(scope="session")
def hostsim_lib(built):
    from course5_b200 import api
    return api.load_library(os.path.join(ROOT, "tests", "hostsim", "libc5hostsim.so"))


@pytest.fixture(scope="session")
def gpu_lib(built):
    from course5_b200 import api
    return api.load_library()
