import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)      # tests/kernel_handles.py (single-kernel helpers of the unit tests)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def measured():
    """record(name, value): measured parity figures of the GPU tests, appended to gpurun_out/measured_parity.jsonl when
    that directory exists (the bounds in the tests are set from these numbers; profiles/ keeps a copy per round)."""
    import json
    out_dir = os.path.join(ROOT, "gpurun_out")

    def record(name, value):
        print(f"[measured] {name}: {value}")
        if os.path.isdir(out_dir):
            with open(os.path.join(out_dir, "measured_parity.jsonl"), "a") as f:
                f.write(json.dumps({"name": name, "value": value}) + "\n")
    return record
