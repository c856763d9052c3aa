import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "fullsize: BASELINE.json configs at full size against the oracle (minutes of CPU)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "efe_golden.npz"))
    out = {}
    for k in z.files:
        case, field = k.split("/")
        out.setdefault(case, {})[field] = z[k]
    return out
