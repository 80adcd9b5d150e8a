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


def test_polish_genome_on_real_files(tmp_path, monkeypatch):
    """`helen polish` end to end on files on disk (PolishInterface.py:49-91): MarginPolish-layout images -> call_consensus on
    cuda:0 (native feed, prediction file through the native writer) -> stitch (native listing and region reads) -> FASTA.
    The predictions are checked against the oracle, the FASTA against the stitch of the same prediction file through the
    package's Python reader."""
    import os
    import helen_b200.StitchInterface as iface
    from helen_b200 import DataStore as ds
    from helen_b200.PolishInterface import polish_genome
    monkeypatch.setenv("HELEN_B200_HDF5", "minih5")
    monkeypatch.delenv("HELEN_B200_PACKED_PREDICTIONS", raising=False)
    ds.forget_packed_views()
    features = 90
    sd = random_state_dict(features, seed=9)
    model_path = str(tmp_path / "model.pkl")
    torch.save({"model_state_dict": {"module." + k: v for k, v in sd.items()}, "model_optimizer": {},
                "hidden_size": 128, "gru_layers": 1, "epochs": 1}, model_path)
    image_dir = tmp_path / "images"
    image_dir.mkdir()
    rng = np.random.default_rng(3)
    images, n = [], 24
    with hdf5.open_file(str(image_dir / "imgs_0.h5"), "w") as f:
        for i in range(n):
            region, chunk = i // 3, i % 3                               # three overlapping images per 2000-base region
            image = rng.integers(0, 256, (1000, features), dtype=np.uint8)
            start = region * 1800 + chunk * 400
            position = np.stack([np.arange(1000) + start, np.zeros(1000, np.int64), np.zeros(1000, np.int64)], 1)
            base = "images/img_%03d/" % i
            f[base + "contig"] = np.array([b"chrT"], dtype="S")
            f[base + "contig_start"] = np.array([region * 1800])
            f[base + "contig_end"] = np.array([region * 1800 + 1999])
            f[base + "feature_chunk_idx"] = np.array([chunk])
            f[base + "image"] = image
            f[base + "position"] = position
            images.append(image)
    out_dir = str(tmp_path / "out")
    prediction_dir = polish_genome(str(image_dir), model_path, 8, 0, 2, out_dir, "HELEN_prediction", True, None, 1)
    files = [os.path.join(prediction_dir, p) for p in os.listdir(prediction_dir) if p.endswith("hdf")]
    assert len(files) == 1
    ref = predict_windows(OracleWeights.from_state_dict(sd), np.stack(images))
    with hdf5.open_file(files[0], "r") as f:
        for i in range(n):
            region, chunk = i // 3, i % 3
            got = f["predictions/chrT/chrT-%d-%d/%d/bases" % (region * 1800, region * 1800 + 1999, chunk)][()]
            diff = got != ref["base_label"][i]
            assert not diff.any() or (top2_margin(ref["base_prob"][i])[diff] < 1e-5).all()
    fasta = open(os.path.join(out_dir, "HELEN_prediction.fa")).read()
    assert fasta.startswith(">chrT\n") and len(fasta) > 1000
    monkeypatch.setenv("HELEN_B200_NATIVE_READER", "0")
    ds.forget_packed_views()
    monkeypatch.setattr(iface, "get_file_paths_from_directory", lambda directory: files)
    again = open(iface.perform_stitch(prediction_dir, str(tmp_path / "again"), "p", 2)).read()
    assert again == fasta
    ds.forget_packed_views()
