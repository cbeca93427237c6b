import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "ref_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    import cpu_libs
    return cpu_libs.oracle()


@pytest.fixture(scope="session")
def ref():
    import cpu_libs
    r = cpu_libs.ref()
    if r is None:
        pytest.skip("oracle/_ref/libpbref.so not built (needs /root/reference at build time)")
    return r
