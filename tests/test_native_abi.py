"""CPU-side checks of the drop-in boundary: the library builds/loads, exports every
symbol include/helen_b200.h declares, and fails loudly (no fallback) without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

from helen_b200 import _native
from helen_b200 import build as hb_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    hb_build.build()
    return _native.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "helen_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    declared = header_symbols()
    assert declared, "no symbols parsed from the header"
    assert sorted(_native.SIGNATURES) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/helen_b200.h but not exported"


def test_abi_version(lib):
    assert lib.hb_abi_version() == _native.HB_ABI_VERSION


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from helen_b200.predictor import WindowPredictor
    from oracle import random_state_dict
    with pytest.raises(RuntimeError) as err:
        WindowPredictor(random_state_dict(10, 0), device=0)
    assert "CUDA" in str(err.value) or "cuda" in str(err.value)


def test_model_rejects_cpu_tensors():
    from helen_b200.models.TransducerModel import TransducerGRU
    model = TransducerGRU(1, 10, 1, 128, 5, 11)
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, 100, 10), torch.zeros(2, 2, 128))


def test_unsupported_configs_are_loud():
    from helen_b200.models.TransducerModel import TransducerGRU
    with pytest.raises(ValueError):
        TransducerGRU(1, 10, 2, 128, 5, 11)
    with pytest.raises(ValueError):
        TransducerGRU(1, 10, 1, 256, 5, 11)


def test_state_dict_surface_matches_reference_keys():
    from helen_b200.models.TransducerModel import TransducerGRU
    from oracle import STATE_DICT_KEYS, state_dict_shapes
    model = TransducerGRU(1, 90, 1, 128, 5, 11)
    sd = model.state_dict()
    assert list(sd) == list(STATE_DICT_KEYS)
    for k, shape in state_dict_shapes(90).items():
        assert tuple(sd[k].shape) == shape


def test_pkl_round_trip_with_and_without_module_prefix(tmp_path):
    from helen_b200.models.ModelHander import ModelHandler
    from helen_b200.models.TransducerModel import TransducerGRU
    model = TransducerGRU(1, 90, 1, 128, 5, 11)
    path = str(tmp_path / "m.pkl")
    ModelHandler.save_model(model, None, 128, 1, 7, path)
    loaded, hidden, layers, epochs = ModelHandler.load_simple_model(path, 1, 90, 1000, 5, 11)
    assert (hidden, layers, epochs) == (128, 1, 7)
    for k, v in model.state_dict().items():
        assert torch.equal(v, loaded.state_dict()[k])
    # DataParallel-style checkpoint
    torch.save({"model_state_dict": {"module." + k: v for k, v in model.state_dict().items()},
                "model_optimizer": {}, "hidden_size": 128, "gru_layers": 1, "epochs": 3}, path)
    loaded, _, _, epochs = ModelHandler.load_simple_model(path, 1, 90, 1000, 5, 11)
    assert epochs == 3
    for k, v in model.state_dict().items():
        assert torch.equal(v, loaded.state_dict()[k])


def test_only_the_checkers_touch_the_oracle():
    """oracle/ is test infrastructure: the package, tools/ and every other script must not import it.  Allowed:
    tests/, __graft_entry__.py (smoke() and build() of the checker) and bench.py (cpu_baseline / --impl reference)."""
    allowed_files = {os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")}
    pattern = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    offenders = []
    for base, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in (".git", "gpurun_out", "__pycache__", "tests", "oracle", "baseline", ".pytest_cache")]
        for name in files:
            path = os.path.join(base, name)
            if name.endswith(".py") and path not in allowed_files and pattern.search(open(path).read()):
                offenders.append(os.path.relpath(path, ROOT))
    assert offenders == []
