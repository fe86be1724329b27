import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    return binding.Oracle()


@pytest.fixture(scope="session")
def ref_dir():
    return os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="session")
def blobs224(pkg, ref_dir):
    """bundled blobs (when oracle/_ref/Network travelled) + seeded synthetic fill"""
    return pkg.synth.model_blobs(os.path.join(ref_dir, "Network"), 224, seed=0)


@pytest.fixture(scope="session")
def synth_blobs224(pkg):
    """fully synthetic model, reproducible on any box"""
    return pkg.synth.model_blobs(None, 224, seed=7)


@pytest.fixture(scope="session")
def lib(pkg):
    L = pkg.lib()
    if pkg.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests need a B200 (there is no CPU fallback)")
    return L
