import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture
def hostsim_lib():
    """The product sources compiled for the CPU (tests/hostsim) bound as the active library."""
    import harness
    import yart_b200
    lib = harness.hostsim()
    yart_b200.use_library(lib)
    yield lib


@pytest.fixture
def cuda_lib():
    """The product library (CUDA).  Fails loudly if it is missing — GPU tests never fall back."""
    import yart_b200
    from yart_b200 import capi
    lib = capi.load()  # raises if libyart_b200.so is not built
    yart_b200.use_library(lib)
    yield lib


@pytest.fixture
def cuda_samplers_lib():
    """The YB_RNG_SAMPLERS build of the product library (NaiveSampler / StratifiedSampler compiled in)."""
    import yart_b200
    from yart_b200 import capi
    lib = capi.load(capi.SAMPLERS_LIB)
    yart_b200.use_library(lib)
    yield lib
    yart_b200.use_library(capi.load())
