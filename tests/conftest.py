import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


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
