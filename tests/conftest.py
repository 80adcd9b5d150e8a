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


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a box without a CUDA device."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu tests run on the B200 box with -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def parity_report(record):
    """Append one parity record (dict) to gpurun_out/parity_report.jsonl when that directory exists
    (the GPU box's scratch output, merged back by gpurun); always echo it to stdout for `pytest -s`."""
    import json
    line = json.dumps(record)
    print("[parity] " + line)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_report.jsonl"), "a") as f:
            f.write(line + "\n")


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
