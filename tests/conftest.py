import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# verbatim copies of the reference's own test scripts live under tests/golden/ref_scripts (one is called test_coal.py): they are run
# as subprocesses by tests/test_*_compat.py, never collected as test modules of this suite
collect_ignore_glob = ["golden/*"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def ref():
    """the reference's own serial/OpenMP back-ends, built from /root/reference by oracle/build_ref.py (the oracle)"""
    from tests.support import oracle_library
    return oracle_library()


@pytest.fixture(scope="session")
def b200():
    """the product: B200 back-end through the flat binding; parity runs replay the reference's mt19937 stream"""
    from tests.support import b200_library
    return b200_library()


@pytest.fixture(params=["toms748", "secant"])
def cond_solver(request):
    """runs a test under both root searches of the condensation step: the default (the reference's TOMS 748 with identical
    trial points) and the opt-in safeguarded secant (half the evaluations, same per-step tolerance, different trajectory)"""
    from libcloudphxx_b200 import engine
    engine.set_cond_solver(request.param)
    yield request.param
    engine.set_cond_solver("toms748")
