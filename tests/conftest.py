import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def vectors():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json"), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture()
def oracle():
    from oracle.pyoracle import Oracle
    o = Oracle()
    yield o
    o.close()


@pytest.fixture(scope="session")
def engine_lib():
    """libtgingest must be loadable everywhere (symbols only on CPU boxes)."""
    from distributed_crawler_b200 import engine
    return engine.lib()


@pytest.fixture()
def engine(engine_lib):
    from distributed_crawler_b200.engine import Engine
    e = Engine()
    yield e
    e.close()
