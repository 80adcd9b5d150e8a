"""End-to-end driver on the GPU: .pkl checkpoint -> predict() -> prediction file, checked against
the oracle (the HDF5 layer is the in-memory fake; h5py is not installed on the boxes)."""
import numpy as np
import pytest
import torch

import fake_h5
from helen_b200 import hdf5
from oracle import OracleWeights, predict_windows, random_state_dict
from oracle.explicit import top2_margin

pytestmark = pytest.mark.gpu


def test_predict_driver_writes_reference_schema(tmp_path, monkeypatch):
    from helen_b200.models.predict_gpu import predict_gpu
    fake_h5.reset()
    monkeypatch.setattr(hdf5, "open_file", fake_h5.open_file)
    sd = random_state_dict(90, seed=4)
    model_path = str(tmp_path / "model.pkl")
    torch.save({"model_state_dict": {"module." + k: v for k, v in sd.items()}, "model_optimizer": {},
                "hidden_size": 128, "gru_layers": 1, "epochs": 1}, model_path)
    rng = np.random.default_rng(1)
    images = []
    for i in range(5):
        length = 1000 if i != 3 else 640                     # one short image exercises the zero padding
        img = rng.integers(0, 256, (length, 90), dtype=np.uint8)
        pos = np.stack([np.arange(length) + i * 1000, np.zeros(length, int), np.zeros(length, int)], 1)
        fake_h5.add_image("in.h5", f"img{i}", "chrX", i * 1000, i * 1000 + length, i, img, pos)
        padded = np.zeros((1000, 90), np.uint8)
        padded[:length] = img
        images.append(padded)
    prefix = str(tmp_path / "pred")
    predict_gpu([["in.h5"]], prefix, model_path, batch_size=2, total_callers=1, devices=[0], num_workers=0)
    ref = predict_windows(OracleWeights.from_state_dict(sd), np.stack(images))
    out = fake_h5.open_file(prefix + "_0.hdf")
    for i in range(5):
        length = 1000 if i != 3 else 640
        chunk = out[f"predictions/chrX/chrX-{i * 1000}-{i * 1000 + length}/{i}"]
        for name, key, prob in (("bases", "base_label", "base_prob"), ("rles", "rle_label", "rle_prob")):
            got = chunk[name][()]
            diff = got != ref[key][i]
            assert got.dtype == np.uint8 and got.shape == (1000,)
            assert not diff.any() or (top2_margin(ref[prob][i])[diff] < 1e-5).all()
        assert chunk["position"][()].shape == (1000, 3)
    fake_h5.reset()
