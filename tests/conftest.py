import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope='session')
def small_problem():
    from mcmc_spec_b200 import synthetic
    return synthetic.make_problem_small()


@pytest.fixture(scope='session')
def small_oracle(small_problem):
    from oracle_adapter import to_oracle
    return to_oracle(small_problem)


@pytest.fixture(scope='session')
def native_lib():
    """libmcmcspec.so, built in-tree if needed (nvcc cross-compiles without a GPU)."""
    from mcmc_spec_b200 import build, native
    build.build()
    return native.load_library()
