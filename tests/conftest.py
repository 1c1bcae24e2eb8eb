import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import open_whisper_kit_b200 as pkg
        return pkg.load(strict_api=False).whisper_b200_device_count() > 0
    except OSError:
        return False


@pytest.fixture(scope="session")
def lib():
    """The product library.  GPU tests must never silently pass without it."""
    import open_whisper_kit_b200 as pkg
    lib = pkg.load(strict_api=False)
    if lib.whisper_b200_device_count() <= 0:
        pytest.fail("no CUDA device visible: GPU parity tests cannot run (there is no CPU fallback)")
    return lib


@pytest.fixture(scope="session")
def model_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("models"))
