import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
        return pkg.lib().Nfft4GPB200DeviceCount() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """The product library with a CUDA device selected.  GPU tests FAIL (not skip) when the extension is missing."""
    import preconditioned_additive_gaussian_processes_with_fourier_acceleration_b200 as pkg
    L = pkg.lib()
    assert L.Nfft4GPB200DeviceCount() > 0, "no CUDA device: -m gpu tests must run on the GPU box"
    assert L.Nfft4GPB200Init(0) == 0
    return pkg
