import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (its directory name has a hyphen)."""
    return importlib.import_module("a-nice-rag_b200")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("a-nice-rag_b200.synth")


@pytest.fixture(scope="session")
def small_case():
    path = os.path.join(ROOT, "tests", "golden", "small_case.npz")
    with np.load(path, allow_pickle=True) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def config0_golden():
    path = os.path.join(ROOT, "tests", "golden", "config0_outputs.npz")
    with np.load(path, allow_pickle=True) as z:
        return {k: z[k] for k in z.files}
