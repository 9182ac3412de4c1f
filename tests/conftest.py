import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(params=["tc", "simt"])
def gemm_path(request):
    """Run a GPU test on both dense-layer paths: tcgen05 3xTF32 tensor-core GEMMs (default) and the
    strict-fp32 SIMT GEMMs (GCBF_TENSOR_CORES=0)."""
    from gcbfplus_b200 import _lib
    old = _lib.USE_TC
    _lib.USE_TC = request.param == "tc"
    yield request.param
    _lib.USE_TC = old
