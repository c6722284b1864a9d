import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture
def record(request):
    """record(name, **values): append a measured parity figure to gpurun_out/parity_measurements.jsonl (when that scratch
    directory exists, i.e. under gpurun) so that tolerances in the tests can be read next to what was actually measured."""
    import json

    def rec(name, **values):
        d = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(d):
            with open(os.path.join(d, "parity_measurements.jsonl"), "a") as f:
                f.write(json.dumps(dict(test=request.node.name, name=name, **values)) + "\n")
    return rec
