import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    """Golden vectors produced by running the reference (tests/golden/make_golden.py)."""
    arrays = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    with open(os.path.join(ROOT, "tests", "golden", "golden_meta.json")) as f:
        doc = json.load(f)
    return arrays, doc["meta"], doc["operators"]


def json_to_opdict(j):
    return {tuple((int(i), bool(d)) for i, d in label): float(c) for label, c in zip(j["labels"], j["coeffs"])}


def assert_opdict_close(a, b, tol=1e-13):
    assert set(a.keys()) == set(b.keys()), (sorted(set(a) ^ set(b))[:4], len(a), len(b))
    for k in a:
        assert abs(a[k] - b[k]) <= tol, (k, a[k], b[k])
