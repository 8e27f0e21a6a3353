import os
import sys

import pytest

# the reference goldens, the CPU oracle and the parity tests are defined on the structural stand-in backbone (timm is not
# in the image): opt in explicitly -- the package's default is the timm-compatible definition (esmstereo_b200/backbone.py)
os.environ.setdefault("ESM_BACKBONE", "standin")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("filterwarnings", "ignore:esmstereo_b200. ESM_BACKBONE=standin")
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
