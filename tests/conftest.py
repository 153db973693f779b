import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(autouse=True)
def reference_order_by_default():
    """The parity suites assert BITWISE equality with the reference, which is what YC_TRAVERSAL_REFERENCE_ORDER
    delivers; contexts created without an explicit `traversal` use it.  The wide walk (the library's default for
    scenes without alpha-tested materials) is tested by name in test_wide_bvh.py / test_gpu_parity.py with the
    north-star tolerances (ids exact bar ties, t <= 1e-5, relMSE < 1e-3)."""
    import gc
    import yart_b200
    old = yart_b200.default_traversal
    yart_b200.default_traversal = yart_b200.TRAVERSAL_REFERENCE_ORDER
    yield
    yart_b200.default_traversal = old
    # contexts a test did not close hold 3-4 GB of wavefront storage each on the GPU: release them now rather than
    # whenever the collector gets to them (a full `-m gpu` run in one process would otherwise pile them up)
    gc.collect()


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
