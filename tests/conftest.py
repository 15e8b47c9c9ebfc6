import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "unverified: GPU test written after the round's last GPU window; runs only with BKGPU_EXPERIMENTAL=1 "
                                       "(the same switch turns on the library paths it exercises)")


EXPERIMENTAL = os.environ.get("BKGPU_EXPERIMENTAL", "0") not in ("", "0")


def pytest_collection_modifyitems(config, items):
    if EXPERIMENTAL:
        return
    skip = pytest.mark.skip(reason="not yet run on a GPU: set BKGPU_EXPERIMENTAL=1")
    for item in items:
        if "unverified" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the native pieces exist (the driver runs build() first; this keeps a bare
    `pytest` self-contained)."""
    from baikaldb_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH) or not os.path.exists(os.path.join(ROOT, "oracle", "libbk_oracle.so")):
        import __graft_entry__
        __graft_entry__.build()
