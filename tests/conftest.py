import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected explicitly with -m gpu; when no device is present skip them loudly.
    has_gpu = None
    for item in items:
        if "gpu" in item.keywords:
            if has_gpu is None:
                try:
                    from cosmic_profiles_b200 import _lib
                    has_gpu = _lib.device_count() > 0
                except Exception:
                    has_gpu = False
            if not has_gpu:
                item.add_marker(pytest.mark.skip(reason="no CUDA device visible"))
