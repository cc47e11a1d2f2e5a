import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        from xmimsim_b200 import abi
        return abi.lib().xmb_cuda_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: no silent fallbacks.
    return


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def srm1155():
    import xmimsim_b200 as x
    return x.read_xmsi(os.path.join(GOLDEN, "srm1155.xmsi"))


@pytest.fixture(scope="session", autouse=True)
def _plugin_provider():
    """The plugin symbols (xmi_solid_angle_calculation_cl, ...) refuse to fall back to the analytic stand-in on their
    own; the image has no xraylib, so the test process registers the stand-in explicitly, as a host would register its
    provider (tests/test_plugin_harness_gpu.py covers the refusal in fresh processes)."""
    from xmimsim_b200 import abi
    L = abi.lib()
    L.xmb_plugin_set_provider(L.xmb_xrl_surrogate())
    yield
