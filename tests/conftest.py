import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (*.so are git-ignored): build the product library and the checkers
    once, exactly as __graft_entry__.build() does.  Building is not a fallback: without nvcc this fails loudly."""
    need = [os.path.join(ROOT, "mallie_b200", "libmallie_b200.so"), os.path.join(ROOT, "oracle", "liboracle.so"),
            os.path.join(ROOT, "mallie_b200", "host_api_check"), os.path.join(ROOT, "mallie_b200", "mallie_b200_cli")]
    if all(os.path.exists(p) for p in need):
        return
    import __graft_entry__
    __graft_entry__.build()
