import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_case_names():
    return sorted(os.path.basename(p)[len("case_"):-len(".npz")]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "case_*.npz")))


def load_case(name):
    case = dict(np.load(os.path.join(GOLDEN_DIR, f"case_{name}.npz")))
    model = str(case["model"])
    state = dict(np.load(os.path.join(GOLDEN_DIR, f"model_{model}.npz")))
    return case, state


@pytest.fixture(scope="session")
def golden_names():
    return golden_case_names()
