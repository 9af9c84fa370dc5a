import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def bmo():
    """The product package (beamletoptics.jl_b200 loaded as bmo_b200) with libbmo.so built."""
    import __graft_entry__ as ge
    ge.build_libbmo()
    return ge.load_package()


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    import __graft_entry__ as ge
    ge.build_oracle()
    from oracle import oracle
    return oracle
