import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "doctests.json")) as f:
        return {v["name"]: v for v in json.load(f)["vectors"]}


def untuple(x):
    """JSON has lists where Rust has tuples."""
    if isinstance(x, (list, tuple)):
        return tuple(untuple(y) for y in x)
    return x
