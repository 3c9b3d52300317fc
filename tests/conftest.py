import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def published_kats():
    """Known answers published by independent implementations (see the file's _source)."""
    with open(os.path.join(ROOT, "tests", "golden", "published_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gpu():
    """Initialises the CUDA library on cuda:0; GPU tests fail loudly (no fallback) if that is impossible."""
    from kzg_b200 import _ffi
    _ffi.init(0)
    return _ffi


@pytest.fixture(scope="session")
def cpu():
    from oracle import cpu as c
    c.build()
    return c
