import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ge():
    import __graft_entry__ as ge
    return ge


@pytest.fixture(scope="session")
def pkg(ge):
    return ge.load_package()


@pytest.fixture(scope="session")
def oracle(ge):
    return ge.load_oracle()


@pytest.fixture(scope="session")
def known():
    with open(os.path.join(GOLDEN, "reference_known_answers.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def stl_points():
    return {n: np.load(os.path.join(GOLDEN, f"{n}_face_centres_f32.npy")) for n in ("cavity", "bifurcation")}


@pytest.fixture(scope="session")
def ctx(pkg):
    """A real device context. Fails loudly (no fallback) when the GPU or the library is missing."""
    c = pkg.Context(0)
    yield c
    c.close()
