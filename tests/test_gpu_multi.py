"""Two-GPU tests of the product's one-process-per-GPU drivers (run with `gpurun --gpus 2`; skipped on one GPU).

* predict_gpu(file_chunks, ..., total_callers=2, devices=[0, 1]) - the reference's mp.spawn driver
  (helen/modules/python/models/predict_gpu.py:186-226, CallConsensusInterface.py:135-149): each rank reads its own image
  files and writes its own <prefix>_<rank>.hdf; both are checked against the oracle.
* train_distributed on two GPUs for two epochs: NCCL all-reduce of the flat gradient buffer after every chunk step; the
  loop itself asserts after every epoch that the replicas are bit-identical (grad_sync.assert_replicas_identical).
Image and prediction files are real HDF5 files written / read through helen_b200.hdf5 (minih5 when h5py is absent)."""
import os

import numpy as np
import pytest
import torch

from oracle import OracleWeights, predict_windows, random_state_dict
from oracle.explicit import top2_margin

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]


def _write_images(path, n_images, features, seed, with_labels=False, length=1000):
    from helen_b200 import hdf5
    rng = np.random.default_rng(seed)
    images = []
    with hdf5.open_file(path, "w") as f:
        for i in range(n_images):
            n = length if (i % 4 or with_labels) else length - 300
            image = rng.integers(0, 256, (n, features), dtype=np.uint8)
            base = "images/img_%04d/" % i
            f[base + "contig"] = np.array([b"chr%d" % seed], dtype="S")
            f[base + "contig_start"] = np.array([i * 1000])
            f[base + "contig_end"] = np.array([i * 1000 + n])
            f[base + "feature_chunk_idx"] = np.array([i])
            f[base + "image"] = image
            f[base + "position"] = np.stack([np.arange(n) + i * 1000, np.zeros(n, np.int64), np.zeros(n, np.int64)], 1)
            if with_labels:
                f[base + "label_base"] = (image[:, 0] // 52).clip(0, 4).reshape(-1, 1).astype(np.int64)
                f[base + "label_run_length"] = (image[:, 1] // 24).clip(0, 10).reshape(-1, 1).astype(np.int64)
            padded = np.zeros((length, features), np.uint8)
            padded[:n] = image
            images.append(padded)
    return np.stack(images)


def test_predict_gpu_two_callers(tmp_path):
    from helen_b200 import hdf5
    from helen_b200.models.predict_gpu import predict_gpu
    features = 90
    sd = random_state_dict(features, seed=9)
    model_path = str(tmp_path / "model.pkl")
    torch.save({"model_state_dict": sd, "model_optimizer": {}, "hidden_size": 128, "gru_layers": 1, "epochs": 1}, model_path)
    shards = []
    for rank in range(2):
        files = [str(tmp_path / ("images_%d_%d.h5" % (rank, k))) for k in range(2)]
        shards.append((files, [_write_images(p, 9 + 3 * k + rank, features, seed=10 * rank + k + 1) for k, p in enumerate(files)]))
    prefix = str(tmp_path / "pred")
    predict_gpu([s[0] for s in shards], prefix, model_path, batch_size=8, total_callers=2, devices=[0, 1], num_workers=0)
    weights = OracleWeights.from_state_dict(sd)
    for rank, (files, arrays) in enumerate(shards):
        assert os.path.exists(prefix + "_%d.hdf" % rank)
        with hdf5.open_file(prefix + "_%d.hdf" % rank, "r") as out:
            for k, images in enumerate(arrays):
                ref = predict_windows(weights, images)
                contig = "chr%d" % (10 * rank + k + 1)
                for i in range(images.shape[0]):
                    n = 1000 if i % 4 else 700
                    chunk = out["predictions/%s/%s-%d-%d/%d" % (contig, contig, i * 1000, i * 1000 + n, i)]
                    for name, key, prob in (("bases", "base_label", "base_prob"), ("rles", "rle_label", "rle_prob")):
                        got = np.asarray(chunk[name][()])
                        diff = got != ref[key][i]
                        assert got.dtype == np.uint8 and got.shape == (1000,)
                        assert not diff.any() or (top2_margin(ref[prob][i])[diff] < 1e-5).all(), (rank, k, i, name)


def test_train_distributed_two_gpus(tmp_path):
    from helen_b200.models.ModelHander import ModelHandler
    from helen_b200.models.train_distributed import train_distributed
    data_dir = tmp_path / "images"
    data_dir.mkdir()
    _write_images(str(data_dir / "train.h5"), 12, 10, seed=3, with_labels=True, length=200)
    model_dir, stats_dir = str(tmp_path) + "/models_", str(tmp_path) + "/stats_"
    train_distributed(str(data_dir), str(data_dir), batch_size=3, epochs=2, gpu_mode=True, num_workers=0, retrain_model=False,
                      retrain_model_path=None, gru_layers=1, hidden_size=128, learning_rate=1e-3, weight_decay=0.0,
                      model_dir=model_dir, stats_dir=stats_dir, device_ids=[0, 1], total_callers=2, train_mode=True)
    # (the loop raised if the two replicas had diverged)  rank 0 leaves one reference-format checkpoint per epoch
    model, hidden, layers, epochs = ModelHandler.load_simple_model(model_dir + "HELEN_epoch_2_checkpoint.pkl", 1, 10, 1000, 5, 11)
    assert (hidden, layers, epochs) == (128, 1, 1)
    losses = [float(line.split(",")[1]) for line in open(stats_dir + "test_loss.csv")]
    assert len(losses) == 2 and np.isfinite(losses).all() and losses[1] < losses[0], losses
