"""pytest configuration: the ``gpu`` marker and import paths.

``-m "not gpu"`` (build container, no GPU): oracle vs the committed golden vectors, host planner,
C-ABI library loads and exports every symbol of include/kvc.h.
``-m gpu`` (B200): parity of the CUDA path against the oracle and the golden vectors.
"""

import os
import sys

import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
PKG = os.path.join(ROOT, "cs3602-llm-inference-acceleration_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json

    import numpy as np

    import cases

    data = np.load(cases.GOLDEN_NPZ)
    manifest = json.loads(bytes(data["manifest"]).decode())
    return data, manifest
