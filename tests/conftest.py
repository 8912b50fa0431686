import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from oracle.oracle_py import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def oracle_omp(built):
    from oracle.oracle_py import Oracle
    return Oracle(openmp=True)


@pytest.fixture(scope="session")
def tws(built):
    import terrainwatersim_b200 as t
    t._abi.load()
    return t
