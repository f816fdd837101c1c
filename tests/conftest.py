import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def use_cusim():
    """Point the Python host at tests/cusim/_build/libwgpu_b200_sim.so: the product's host runtime source compiled
    against a software model of CUDA (tests/cusim/cusim_device.h), so that the kernels themselves run on host cores.
    Test infrastructure only -- the product library is untouched and still refuses to run without a device."""
    from tests.cusim import build as cusim_build
    from wgpu_cpu_b200 import api
    lib = cusim_build.build()
    assert api._lib is None or api.LIB_PATH == lib, "the native library was loaded before the model was selected"
    api.LIB_PATH = lib


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "cusim: runs the CUDA kernels on the software model of tests/cusim (no GPU needed)")
    # WGB_CUSIM=1 python -m pytest tests -m gpu : the whole GPU suite on the software model (slow; for development
    # on a machine without a GPU).  The default CPU tier runs the subset in tests/test_cusim_parity.py.
    if os.environ.get("WGB_CUSIM") == "1":
        use_cusim()
