"""Interoperability with libhdf5, wherever h5py is installed (it is not in the build container nor on the GPU boxes, so
these tests skip there): files written by the package's writers must read back through h5py, and files written by h5py
must read through minih5 and through the native feed library.  (Round-1 advice: real h5py semantics - string datasets,
the uint32 position round trip - were only exercised through tests/fake_h5.py.)"""
import numpy as np
import pytest

h5py = pytest.importorskip("h5py")

from helen_b200 import _feed_native, minih5  # noqa: E402


@pytest.mark.parametrize("native_writer", ["1", "0"])
def test_h5py_reads_what_the_package_writes(tmp_path, monkeypatch, native_writer):
    from helen_b200.DataStore import DataStore
    monkeypatch.setenv("HELEN_B200_HDF5", "minih5")
    monkeypatch.setenv("HELEN_B200_NATIVE_WRITER", native_writer)
    path = str(tmp_path / "pred.hdf")
    position = np.stack([np.arange(1000), np.zeros(1000, np.int64), np.zeros(1000, np.int64)], 1)[None].repeat(4, 0)
    position[:, 990:] = -1
    gen = np.random.default_rng(0)
    bases, rles = gen.integers(0, 5, (4, 1000)), gen.integers(0, 11, (4, 1000))
    store = DataStore(path, mode="w", packed=False)
    store.write_predictions(["chrA"] * 4, [0, 0, 2000, 2000], [1999, 1999, 3999, 3999], [0, 1, 0, 1], position, bases, rles)
    store.close()
    with h5py.File(path, "r") as f:
        assert sorted(f["predictions/chrA"].keys()) == ["chrA-0-1999", "chrA-2000-3999"]
        chunk = f["predictions/chrA/chrA-2000-3999/1"]
        assert chunk["position"].dtype == np.uint32 and chunk["position"][995, 0] == 4294967295
        assert np.array_equal(chunk["bases"][()], bases[3]) and np.array_equal(chunk["rles"][()], rles[3])
        assert f["predictions/chrA/chrA-2000-3999/contig_start"][()] == 2000


def test_package_readers_read_what_h5py_writes(tmp_path, monkeypatch):
    from helen_b200.models.bulk_reader import BulkImageBatches
    path = str(tmp_path / "images.h5")
    gen = np.random.default_rng(1)
    images = []
    with h5py.File(path, "w") as f:
        for i in range(9):
            length = 1000 if i % 2 else 800
            image = gen.integers(0, 256, (length, 90), dtype=np.uint8)
            group = f.create_group("images/img_%03d" % i)
            group["contig"] = np.array([b"chr'7'"], dtype="S")
            group["contig_start"] = np.array([i * 1000])
            group["contig_end"] = np.array([i * 1000 + length])
            group["feature_chunk_idx"] = np.array([i])
            group["image"] = image
            group["position"] = np.stack([np.arange(length), np.zeros(length, np.int64), np.zeros(length, np.int64)], 1)
            images.append(image)
    with minih5.File(path) as f:
        assert list(f["images"].keys()) == ["img_%03d" % i for i in range(9)]
        assert np.array_equal(f["images/img_004/image"][()], images[4])
    assert len(_feed_native.ImageFile(path)) == 9
    for native in (True, False):
        monkeypatch.setenv("HELEN_B200_HDF5", "minih5")
        batch = BulkImageBatches(None, file_list=[path], batch_size=16, native=native)[0]
        assert batch[0][0] == "chr7" and batch[4].shape == (9, 1000, 90)
        for i in range(9):
            assert np.array_equal(batch[4][i, :len(images[i])].numpy(), images[i]) and not batch[4][i, len(images[i]):].any()
