import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_aug():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "augment.npz"))


@pytest.fixture(scope="session")
def golden_simple():
    """Step fixtures of the imported reference for the simple / gated / cross-attention encoders (make_golden.py simple)."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden_simple.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_contrastive():
    """Step fixtures of the imported reference's stand-alone InfoNCE / SimCLR Lightning modules (make_golden.py contrastive)."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden_contrastive.json")) as f:
        return json.load(f)
