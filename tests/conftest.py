import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE_RESOURCES = "/root/reference/resources"  # CPU container only; absent on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(HERE, "golden")


def has_reference_resources():
    return os.path.isdir(REFERENCE_RESOURCES)
