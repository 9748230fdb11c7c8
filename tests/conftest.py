import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The development switches of the library must not leak into a test run from the caller's environment:
# a parity test that silently ran the torch fallback or another kernel family would prove nothing.  Tests that compare kernel
# families set the switch themselves, in a subprocess.
# (A test that re-runs pytest under a switch marks its child with TPX_TEST_SUBPROCESS=1.)
if os.environ.get("TPX_TEST_SUBPROCESS") != "1":
    for _name in ("TPX_ALLOW_TORCH", "TPX_DISABLE_TC", "TPX_DISABLE_TCW", "TPX_TCW_SPLIT", "TPX_TCP", "TPX_SAVE_GB"):
        os.environ.pop(_name, None)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
