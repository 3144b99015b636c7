import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.build()
    return oracle_lib


@pytest.fixture(scope="session")
def dm():
    import depthmatch
    depthmatch.load()
    return depthmatch


@pytest.fixture(autouse=True)
def _reset_tuning_options(request):
    """Tests that flip a tuning switch on the shared default context leave it in auto."""
    yield
    if "gpu" in request.keywords and _has_gpu():
        import depthmatch
        from depthmatch import api
        for ctx in list(api._default.values()):
            ctx.set_option("ssd_form", "auto")
            ctx.set_option("no_small_tiles", "0")
            ctx.set_option("sweep", "0")
            ctx.set_option("conv", "0")
            ctx.set_option("volume_kernel", "0")
