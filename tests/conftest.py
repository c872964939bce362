import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oc():
    """C oracle (oracle/libgl_oracle.so), built on demand.  Checker only."""
    from oracle_c import OracleC
    return OracleC()


@pytest.fixture(scope="session")
def ctx():
    """Product context on cuda:0.  Fails loudly (no fallback) when the CUDA library or device is missing."""
    import plonky25_b200 as g
    g.build()
    c = g.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden():
    import json
    d = os.path.join(ROOT, "tests", "golden")
    return {name[:-5]: json.load(open(os.path.join(d, name))) for name in os.listdir(d) if name.endswith(".json")}
