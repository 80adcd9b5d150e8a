"""Host-side mirror of the reference surface: dataset padding rules, prediction file schema,
file sharding, CLI flags and argument-error behaviour (CPU only)."""
import os

import numpy as np
import pytest

import fake_h5
from helen_b200 import hdf5
from helen_b200.options import ImageSizeOptions, TrainOptions


@pytest.fixture(autouse=True)
def _fake_hdf5(monkeypatch):
    fake_h5.reset()
    monkeypatch.setattr(hdf5, "open_file", fake_h5.open_file)
    yield
    fake_h5.reset()


def test_options_match_reference_constants():
    assert (ImageSizeOptions.IMAGE_HEIGHT, ImageSizeOptions.SEQ_LENGTH, ImageSizeOptions.SEQ_OVERLAP) == (90, 1000, 200)
    assert (ImageSizeOptions.TOTAL_BASE_LABELS, ImageSizeOptions.TOTAL_RLE_LABELS) == (5, 11)
    assert (TrainOptions.TRAIN_WINDOW, TrainOptions.WINDOW_JUMP, TrainOptions.GRU_LAYERS, TrainOptions.HIDDEN_SIZE) == (100, 50, 1, 128)


def test_sequence_dataset_pads_short_images():
    from helen_b200.models.dataloader_predict import SequenceDataset
    rng = np.random.default_rng(0)
    full = rng.integers(0, 256, (1000, 90), dtype=np.uint8)
    short = rng.integers(0, 256, (412, 90), dtype=np.uint8)
    pos_full = np.stack([np.arange(1000), np.zeros(1000, int), np.zeros(1000, int)], 1)
    fake_h5.add_image("a.h5", "img0", "chr1", 0, 1000, 3, full, pos_full)
    fake_h5.add_image("a.h5", "img1", "chr1", 1000, 1412, 4, short, pos_full[:412])
    fake_h5.open_file("empty.h5", "w")                      # a file without images only warns
    ds = SequenceDataset(None, file_list=["a.h5", "empty.h5"])
    assert len(ds) == 2
    contig, start, end, chunk, image, position, path = ds[0]
    assert (contig, start, end, chunk, path) == ("chr1", 0, 1000, 3, "a.h5")
    assert image.dtype == np.uint8 and image.shape == (1000, 90) and np.array_equal(image, full)
    _, _, _, _, image, position, _ = ds[1]
    assert image.shape == (1000, 90) and np.array_equal(image[:412], short) and not image[412:].any()
    assert position.shape == (1000, 3) and (position[412:] == -1).all()


def test_sequence_dataset_batches_through_dataloader():
    from torch.utils.data import DataLoader
    from helen_b200.models.dataloader_predict import SequenceDataset
    for i in range(5):
        fake_h5.add_image("b.h5", f"img{i}", "ctg", i * 1000, (i + 1) * 1000, i,
                          np.full((1000, 90), i, np.uint8), np.zeros((1000, 3), int))
    loader = DataLoader(SequenceDataset(None, file_list=["b.h5"]), batch_size=4, shuffle=False, num_workers=0)
    contig, start, end, chunk, images, position, path = next(iter(loader))
    assert images.shape == (4, 1000, 90) and images.dtype.is_floating_point is False
    assert list(contig) == ["ctg"] * 4 and start.tolist() == [0, 1000, 2000, 3000]


def test_sequence_dataset_keeps_one_file_open(monkeypatch):
    """Consecutive items of one file reuse its handle; moving to the next file closes it; workers start without one."""
    import pickle
    import helen_b200.hdf5 as hb_hdf5
    from helen_b200.models.dataloader_predict import SequenceDataset
    for name in ("c.h5", "d.h5"):
        for i in range(3):
            fake_h5.add_image(name, f"img{i}", "ctg", i * 1000, (i + 1) * 1000, i, np.zeros((1000, 10), np.uint8), np.zeros((1000, 3), int))
    ds = SequenceDataset(None, file_list=["c.h5", "d.h5"])
    opened, closed = [], []
    real_open = hb_hdf5.open_file

    def counting_open(path, mode='r'):
        handle = real_open(path, mode)
        opened.append(path)
        monkeypatch.setattr(handle, "close", lambda: closed.append(path), raising=False)
        return handle
    monkeypatch.setattr(hb_hdf5, "open_file", counting_open)
    for i in range(len(ds)):
        ds[i]
    assert opened == ["c.h5", "d.h5"] and closed == ["c.h5"]
    clone = pickle.loads(pickle.dumps(ds))
    assert clone._open_file is None and len(clone) == 6
    ds.close()
    assert closed == ["c.h5", "d.h5"]


def test_datastore_schema_and_dedup():
    from helen_b200.DataStore import DataStore
    store = DataStore("out_0.hdf", "w")
    pos = np.full((1000, 3), -1)
    pos[:10, 0] = np.arange(10)
    bases = np.arange(1000) % 5
    rles = np.arange(1000) % 11
    store.write_prediction("chr2", np.int64(100), np.int64(1100), np.int64(7), pos, bases, rles, "f.h5")
    store.write_prediction("chr2", np.int64(100), np.int64(1100), np.int64(7), pos, bases * 0, rles, "f.h5")   # duplicate: ignored
    store.write_prediction("chr2", np.int64(100), np.int64(1100), np.int64(8), pos, bases, rles, "f.h5")
    f = fake_h5.open_file("out_0.hdf")
    region = f["predictions/chr2/chr2-100-1100"]
    assert int(region["contig_start"][()]) == 100 and int(region["contig_end"][()]) == 1100
    chunk = region["7"]
    assert chunk["bases"][()].dtype == np.uint8 and np.array_equal(chunk["bases"][()], bases)
    assert chunk["rles"][()].dtype == np.uint8
    assert chunk["position"][()].dtype == np.uint32 and chunk["position"][()][20, 0] == 4294967295   # -1 wraps, as in the reference
    assert "8" in region


def test_datastore_batched_write_equals_per_record():
    """write_predictions with the collated batch a DataLoader yields (lists of str, torch tensors) leaves the same
    file as write_prediction per record (predict_gpu.py:176-179)."""
    import torch
    from helen_b200.DataStore import DataStore
    rng = np.random.default_rng(0)
    n = 5
    contig = ["chr1", "chr1", "chr2", "chr2", "chr2"]
    start = torch.tensor([0, 0, 500, 500, 500])
    end = torch.tensor([900, 900, 1400, 1400, 1400])
    chunk = torch.tensor([0, 1, 0, 1, 1])                      # the last record repeats a key: ignored in both paths
    pos = torch.from_numpy(rng.integers(-1, 2000, (n, 1000, 3)))
    bases, rles = rng.integers(0, 5, (n, 1000)).astype(np.uint8), rng.integers(0, 11, (n, 1000)).astype(np.uint8)
    one = DataStore("per_record.hdf", "w")
    for i in range(n):
        one.write_prediction(contig[i], start[i], end[i], chunk[i], pos[i], bases[i], rles[i], "f.h5")
    many = DataStore("batched.hdf", "w")
    many.write_predictions(contig, start, end, chunk, pos, bases, rles, ["f.h5"] * n)

    def dump(node, prefix=""):
        if isinstance(node, dict):
            return {k2: v2 for k, v in node.items() for k2, v2 in dump(v, prefix + "/" + k).items()}
        return {prefix: node.value}
    a, b = dump(fake_h5.open_file("per_record.hdf").root), dump(fake_h5.open_file("batched.hdf").root)
    assert sorted(a) == sorted(b) and len(a) == 2 * 2 + 4 * 3
    for key in a:
        assert a[key].dtype == b[key].dtype and np.array_equal(a[key], b[key]), key
    with pytest.raises(ValueError):
        many.write_predictions(contig[:2], start, end, chunk, pos, bases, rles)


def test_round_robin_file_sharding():
    from helen_b200.CallConsensusInterface import shard_files
    files = [f"f{i}.h5" for i in range(7)]
    assert shard_files(files, 3) == [["f0.h5", "f3.h5", "f6.h5"], ["f1.h5", "f4.h5"], ["f2.h5", "f5.h5"]]
    assert shard_files(files[:2], 8) == [["f0.h5"], ["f1.h5"]]          # empty callers are dropped
    assert shard_files([], 4) == []


def test_cli_flags_and_defaults():
    from helen_b200.helen import build_parser
    p = build_parser()
    a = p.parse_args(["polish", "-i", "imgs", "-m", "m.pkl"])
    assert (a.batch_size, a.num_workers, a.threads, a.output_dir, a.output_prefix, a.gpu_mode, a.device_ids, a.callers) == \
        (512, 8, 1, "./output/", "HELEN_prediction", False, None, 8)
    a = p.parse_args(["call_consensus", "-i", "imgs", "-m", "m.pkl", "-b", "256", "-w", "4", "-g", "-d_ids", "0,1", "-t", "32"])
    assert (a.batch_size, a.num_workers, a.gpu_mode, a.device_ids, a.threads) == (256, 4, True, "0,1", 32)
    assert p.parse_args(["call_consensus", "-i", "x", "-m", "y"]).threads == 16


@pytest.mark.parametrize("kwargs", [
    dict(model_path="/nonexistent/model.pkl"),
    dict(image_dir="/nonexistent/dir"),
    dict(batch_size=0),
    dict(num_workers=-1),
    dict(threads=0),
    dict(gpu_mode=False),            # no CPU fallback in this package
])
def test_call_consensus_argument_errors_exit_1(tmp_path, kwargs):
    from helen_b200.CallConsensusInterface import call_consensus
    model = tmp_path / "m.pkl"
    model.write_bytes(b"x")
    args = dict(image_dir=str(tmp_path), model_path=str(model), batch_size=8, num_workers=0, threads=1,
                output_dir=str(tmp_path / "out"), output_prefix="p", gpu_mode=True, device_ids=None, callers=1)
    args.update(kwargs)
    with pytest.raises(SystemExit) as err:
        call_consensus(**args)
    assert err.value.code == 1


def test_helen_train_cli_and_output_directories(tmp_path, capsys):
    """helen_train's flags (helen_train.py:10-137) and the trained_models_<stamp>/stats_<stamp>/ layout
    (FileManager.py:26-49); running a training needs a GPU and is covered by tests/test_gpu_train.py."""
    from helen_b200 import helen_train
    from helen_b200.FileManager import FileManager
    p = helen_train.build_parser()
    a = p.parse_args(["train", "--train_image_dir", "tr", "--test_image_dir", "te", "--gpu_mode", "-d_ids", "2,3"])
    assert (a.batch_size, a.epoch_size, a.output_dir, a.retrain_model, a.retrain_model_path, a.num_workers) == \
        (100, 10, "./model", False, False, 16)
    assert a.gpu_mode is True and a.device_ids == "2,3"
    t = p.parse_args(["test", "--test_image_dir", "te", "--model_path", "m.pkl"])
    assert (t.batch_size, t.gpu_mode, t.print_details, t.output_dir, t.num_workers) == (100, False, False, "./debug_output", 40)
    assert helen_train.main(["version"]) == 0 and "VERSION" in capsys.readouterr().out
    assert helen_train.main([]) == 1
    model_dir, stats_dir = FileManager.handle_train_output_directory(str(tmp_path / "out"))
    assert os.path.isdir(model_dir) and os.path.isdir(stats_dir)
    assert os.path.basename(model_dir.rstrip("/")).startswith("trained_models_") and stats_dir.startswith(model_dir)
    with pytest.raises(SystemExit):                       # no CPU path: loud, like call_consensus without --gpu_mode
        from helen_b200.TrainInterface import train_interface
        train_interface("tr", "te", False, None, 1, 2, 0, str(tmp_path / "out2"), False, False)


def test_confusion_matrix_text(tmp_path):
    from helen_b200.TrainInterface import write_confusion_matrix
    path = str(tmp_path / "cm.txt")
    write_confusion_matrix([[5, 1], [0, 7]], ["-", "A"], path)
    assert open(path).read().splitlines() == ["true\\pred\t-\tA", "-\t5\t1", "A\t0\t7"]


def test_predict_driver_loop_with_stub_predictor(tmp_path, monkeypatch):
    """The driver's batch loop (predict_gpu.py:94-179 in the reference: load, predict, write one record per image)
    without a GPU: the CUDA predictor is replaced by a stub that labels every column with (image index % 5, 1).
    Checks the order of records, the short-image padding, the overlapped writer and the prediction-file schema;
    the real predictor runs through the same loop in tests/test_gpu_driver.py."""
    import torch
    import helen_b200.models.predict_gpu as drv

    class StubPredictor:
        image_features = 90

        def __init__(self, state_dict, device=0):
            self.calls = 0

        def predict(self, images):
            self.calls += 1
            assert images.dtype == torch.uint8 and images.shape[1:] == (1000, 90)
            tag = images[:, 0, 0].to(torch.uint8) % 5                      # first pixel carries the image index
            return tag[:, None].expand(-1, 1000).contiguous(), torch.ones(images.shape[0], 1000, dtype=torch.uint8)

        def close(self):
            pass

    class StubEvent:
        def record(self, stream=None):
            pass

        def synchronize(self):
            pass

    monkeypatch.setattr(drv, "WindowPredictor", StubPredictor)
    monkeypatch.setattr(drv.torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(drv.torch.cuda, "Event", StubEvent)
    monkeypatch.setattr(drv.torch.cuda, "current_stream", lambda d=None: None)
    monkeypatch.setattr(drv, "_cuda_device", lambda device_id: torch.device("cpu"))
    model_path = str(tmp_path / "m.pkl")
    torch.save({"model_state_dict": {}, "model_optimizer": {}, "hidden_size": 128, "gru_layers": 1, "epochs": 1}, model_path)
    for i in range(5):
        length = 1000 if i != 2 else 300
        image = np.full((length, 90), i, np.uint8)
        pos = np.stack([np.arange(length) + 1000 * i, np.zeros(length, int), np.zeros(length, int)], 1)
        fake_h5.add_image("drv.h5", f"img{i}", "chrD", 1000 * i, 1000 * i + length, i, image, pos)
    prefix = str(tmp_path / "pred")
    drv.predict_gpu([["drv.h5"]], prefix, model_path, batch_size=2, total_callers=1, devices=[0], num_workers=0)
    out = fake_h5.open_file(prefix + "_0.hdf")
    for i in range(5):
        length = 1000 if i != 2 else 300
        chunk = out[f"predictions/chrD/chrD-{1000 * i}-{1000 * i + length}/{i}"]
        assert (chunk["bases"][()] == i % 5).all() and (chunk["rles"][()] == 1).all()
        position = chunk["position"][()]
        assert position.dtype == np.uint32 and position[0, 0] == 1000 * i and (position[length:] == 4294967295).all()
