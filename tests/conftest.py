import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The libraries are build artefacts (git-ignored): a checkout that was never built gets them here, once (the same call
    __graft_entry__.build() makes); with both files in place this is two stat() calls."""
    lib = os.path.join(ROOT, "etpnav_b200", "libetpnav_b200.so")
    helper = os.path.join(ROOT, "etpnav_b200", "_gmap_mirror.so")
    if not (os.path.exists(lib) and os.path.exists(helper)):
        from etpnav_b200.build import build
        build(force=False, verbose=False)


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
