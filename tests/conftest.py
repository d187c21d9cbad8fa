"""pytest configuration: the `gpu` marker, paths, shared fixtures."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def goldens():
    """The reference's own fixture (transeq/testdata/test.fna x data.json, transeq_test.go:14-62)."""
    with open(os.path.join(ROOT, "tests", "golden", "emboss_goldens.json")) as f:
        return json.load(f)
