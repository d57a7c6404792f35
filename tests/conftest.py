"""pytest configuration.  `-m "not gpu"` runs here on CPU; `-m gpu` runs on a B200 and goes through the C ABI."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (runs the CUDA library through the C ABI)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build what is buildable here: the product library needs nvcc (cross-compiles without a GPU); the oracle needs g++;
    oracle/_ref additionally needs /root/reference and is otherwise used as shipped."""
    from yune_b200 import build
    try:
        build.build_library()
    except Exception as e:                      # the GPU box has nvcc too; a missing library fails the tests that need it
        print("library build skipped:", e)
    build.build_hostcheck()
    build.build_oracle()


@pytest.fixture(scope="session")
def oracle():
    from tests.refbind import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_kernels():
    from tests.refbind import RefKernels, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return RefKernels()


@pytest.fixture(scope="session")
def ref_host():
    from tests.refbind import RefHost, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return RefHost()


@pytest.fixture(scope="session")
def gpu_manager():
    import yune_b200 as yb
    m = yb.CUDAManager().setup(0)      # raises loudly when there is no B200 / no library
    m.setOption("max_iterations", 200000)
    yield m
    m.close()
