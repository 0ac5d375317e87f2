import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def pkg():
    """The product package. Its directory name has a hyphen, so it is imported through importlib."""
    return importlib.import_module("long-tail-gan_b200")


@pytest.fixture(scope="session")
def lib(pkg):
    pkg._lib.build()
    return pkg._lib.load()
