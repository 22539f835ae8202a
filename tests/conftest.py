import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if stale) + load the C-ABI library and init it on cuda:0.  Fails loudly when anything is missing."""
    import torch
    from superresolution_aniso_mri_b200 import _lib, build
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    build.build_library()
    return _lib.lib_for_device(0)
