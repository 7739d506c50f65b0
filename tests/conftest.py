import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """The GPU gate must be honest: no test of tests/test_gpu_*.py may carry an unconditional `skip` mark after collection (in round 1
    a mark set on an imported function by the emulation suite silently skipped every solve-parity test on the B200).  Conditional
    skips (`skipif` on the GPU count, pytest.skip() at run time) are not affected."""
    leaked = [it.nodeid for it in items if os.path.basename(str(it.fspath)).startswith("test_gpu_") and it.get_closest_marker("skip") is not None]
    if leaked:
        raise pytest.UsageError("unconditional skip marks on GPU tests (leaked from another module?): %s" % leaked[:5])


@pytest.fixture(scope="session")
def pkg():
    return graft.load_package()


@pytest.fixture(scope="session")
def fo():
    from oracle import fea_oracle
    return fea_oracle


@pytest.fixture(scope="session")
def golden_c1():
    return dict(np.load(os.path.join(GOLDEN, "c1_tet_beam.npz")))


@pytest.fixture(scope="session")
def golden_c2():
    return dict(np.load(os.path.join(GOLDEN, "c2_hex_simp.npz")))


@pytest.fixture(scope="session")
def golden_syn():
    return dict(np.load(os.path.join(GOLDEN, "synthetic_tet.npz")))


@pytest.fixture(scope="session")
def have_gpu():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcudart.so")  # noqa: F841
    except OSError:
        pass
    import torch
    return torch.cuda.is_available()
