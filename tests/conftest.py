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
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_sessionfinish(session, exitstatus):
    """After a GPU session: persist how often the BFGS success-flag escape hatch of test_gpu_fit fired."""
    mod = sys.modules.get("test_gpu_fit")
    if mod is None or not getattr(mod, "FLIP_STATS", {}).get("fits_compared"):
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_notes.json"), "w") as f:
            json.dump(dict(mod.FLIP_STATS, exitstatus=int(exitstatus)), f, indent=1)
    except OSError:
        pass
