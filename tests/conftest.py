import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def read_golden(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def golden():
    """The reference's own fixtures: (plain, gzip-compressed) pairs of tests/decoder.rs:4-15."""
    return [
        (read_golden("10x10y"), read_golden("10x10y.compressed.gz")),
        (read_golden("alice29.txt"), read_golden("alice29.txt.compressed.gz")),
    ]


@pytest.fixture(scope="session")
def alice():
    return read_golden("alice29.txt")
