import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_b200():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] >= 10
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a B200 skips the GPU tests instead of failing in sf_create (the
    product has no CPU fallback).  `-m gpu` on the GPU box runs them; there nothing is skipped here."""
    if _has_b200():
        return
    skip = pytest.mark.skip(reason="needs a B200 (sm_100): the product has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def sf():
    """The product binding; builds libsf_b200.so in-tree when it is missing (nvcc cross-compiles)."""
    import simplefluid_b200 as mod
    if not os.path.exists(mod.library_path()):
        mod.build_library()
    return mod


@pytest.fixture(scope="session")
def ob():
    """The oracle binding (test infrastructure)."""
    import oracle_binding as mod
    mod.build_oracle()
    return mod
