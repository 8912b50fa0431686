import os
import sys
from pathlib import Path

# Several strip sims of ONE test process share one GPU and wait for each other with spinning flag kernels; with more
# streams than hardware work queues (default 8) two streams can alias onto one queue and a wait kernel then blocks the very
# push it is waiting for.  32 queues keep every stream of the test configurations on its own queue.  Must be set before
# CUDA initialises.  (One process per GPU — the deployment — never depends on it: a wait only needs work that was
# enqueued before it.)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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
