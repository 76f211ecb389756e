import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionfinish(session, exitstatus):
    """Achieved parity numbers per fixture (golden_util.REPORT): written next to the GPU run's other outputs so that
    the match fractions are on record, not only "passed" (profiles/r02_parity_*.json are copies of these)."""
    import json
    import golden_util
    if not golden_util.REPORT:
        return
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        kind = "gpu" if any(k.startswith("cuda") for k in golden_util.REPORT) else "cpu"
        with open(os.path.join(out, f"parity_report_{kind}.json"), "w") as f:
            json.dump(golden_util.REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass
