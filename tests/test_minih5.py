"""helen_b200.minih5: the pure-NumPy HDF5 subset (no h5py / libhdf5 on the build and GPU boxes).

Round trips through real files for everything the package stores, group B-trees of one, two and three levels, the
chunked / deflate / shuffle read path on a hand-assembled file, and the MarginPolish image layout and the prediction
schema through the package's own reader and writer classes (SequenceDataset, DataStore, the stitch reader)."""
import struct
import zlib

import numpy as np
import pytest

from helen_b200 import hdf5, minih5


@pytest.fixture(autouse=True)
def _force_minih5(monkeypatch):
    monkeypatch.setenv("HELEN_B200_HDF5", "minih5")


def test_round_trip_of_every_stored_type(tmp_path):
    path = str(tmp_path / "a.h5")
    rng = np.random.default_rng(0)
    values = {
        "u8": rng.integers(0, 256, (1000, 90), dtype=np.uint8),
        "u32": rng.integers(0, 2 ** 32, (1000, 3), dtype=np.uint32),
        "i64": rng.integers(-2 ** 40, 2 ** 40, (17,), dtype=np.int64),
        "i16": rng.integers(-300, 300, (4, 5, 6)).astype(np.int16),
        "f32": rng.standard_normal((7, 3)).astype(np.float32),
        "f64": rng.standard_normal(11),
        "names": np.array([b"chr1", b"chrX_random", b""], dtype="S"),
        "empty": np.zeros((0, 3), np.uint32),
    }
    with minih5.File(path, "w") as f:
        for k, v in values.items():
            f["grp/sub/" + k] = v
        f["grp/scalar_int"] = 12345
        f["grp/scalar_np"] = np.int64(-7)
        f["top"] = np.arange(5)
        assert "grp/sub/u8" in f and "grp/nope" not in f
        with pytest.raises(ValueError):
            f["top"] = 1
    with minih5.File(path, "r") as f:
        assert sorted(f.keys()) == ["grp", "top"]
        assert sorted(f["grp"].keys()) == ["scalar_int", "scalar_np", "sub"]
        for k, v in values.items():
            got = f["grp"]["sub"][k][()]
            assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v), k
            assert f["grp/sub/" + k].shape == v.shape
        assert f["grp/scalar_int"][()] == 12345 and f["grp/scalar_np"][()] == -7
        assert np.array_equal(f["top"][1:3], [1, 2])
        assert "grp/sub/u32" in f and "grp/sub/zzz" not in f and "nope/x" not in f
        with pytest.raises(KeyError):
            f["grp/missing"]


@pytest.mark.parametrize("n", [1, 8, 9, 257, 3000])
def test_groups_with_many_members(tmp_path, n):
    """8 members fill one symbol node, 256 one B-tree node: 3000 needs a second B-tree level."""
    path = str(tmp_path / "g.h5")
    names = ["chr1-%d-%d" % (i * 1000, i * 1000 + 999) for i in range(n)]
    with minih5.File(path, "w") as f:
        for i, name in enumerate(names):
            f["predictions/chr1/%s/contig_start" % name] = i
    with minih5.File(path, "r") as f:
        grp = f["predictions/chr1"]
        assert sorted(grp.keys()) == sorted(names)
        for i in (0, n // 2, n - 1):
            assert f["predictions/chr1/%s/contig_start" % names[i]][()] == i


def _assemble_chunked_file(path, array, chunk, deflate, shuffle):
    """A version-0 file with one 2-D dataset 'd' in the root group, stored in chunks behind a version-1 B-tree."""
    w = minih5._Writer(path)
    rank = array.ndim
    grid = [range(0, s, c) for s, c in zip(array.shape, chunk)]
    records = []
    for i in grid[0]:
        for j in grid[1]:
            block = np.zeros(chunk, array.dtype)
            part = array[i:i + chunk[0], j:j + chunk[1]]
            block[:part.shape[0], :part.shape[1]] = part
            raw = block.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, np.uint8).reshape(-1, array.dtype.itemsize).T.tobytes()
            if deflate:
                raw = zlib.compress(raw)
            records.append(((i, j), w._append(raw), len(raw)))
    body = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(records), minih5.UNDEF, minih5.UNDEF)
    for (i, j), address, nbytes in records:
        body += struct.pack("<IIQQQ", nbytes, 0, i, j, 0) + struct.pack("<Q", address)
    body += struct.pack("<IIQQQ", 0, 0, array.shape[0], array.shape[1], 0)
    btree = w._append(body)
    dataspace = struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", d) for d in array.shape)
    layout = struct.pack("<BBBQ", 3, 2, rank + 1, btree) + b"".join(struct.pack("<I", c) for c in chunk) + struct.pack("<I", array.dtype.itemsize)
    filters = []
    if shuffle:
        filters.append(struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<I", array.dtype.itemsize) + b"\0" * 4)
    if deflate:
        filters.append(struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<I", 6) + b"\0" * 4)
    msgs = [minih5._message(0x01, dataspace), minih5._message(0x03, minih5._datatype_message(array.dtype)), minih5._message(0x08, layout)]
    if filters:
        msgs.append(minih5._message(0x0B, struct.pack("<BB6x", 1, len(filters)) + b"".join(filters)))
    header = w._append(minih5._object_header(msgs))
    leaf = minih5._WNode()
    leaf.is_dataset = True
    w.root.children["d"] = leaf
    w._write_dataset = lambda node: header               # the assembled header stands in for the contiguous one
    w.close()


@pytest.mark.parametrize("deflate,shuffle", [(False, False), (True, False), (True, True)])
def test_chunked_and_filtered_datasets_are_read(tmp_path, deflate, shuffle):
    path = str(tmp_path / "c.h5")
    array = np.random.default_rng(1).integers(0, 60000, (1000, 90)).astype(np.uint16)
    _assemble_chunked_file(path, array, (256, 32), deflate, shuffle)
    with minih5.File(path, "r") as f:
        assert np.array_equal(f["d"][()], array)


def _write_marginpolish_like(path, n_images, features=90, with_labels=False):
    rng = np.random.default_rng(7)
    images = []
    with hdf5.open_file(path, "w") as f:
        for i in range(n_images):
            length = 1000 if i % 3 else 700
            image = rng.integers(0, 256, (length, features), dtype=np.uint8)
            position = np.stack([np.arange(length) + i * 1000, np.zeros(length, np.int64), np.zeros(length, np.int64)], 1)
            base = "images/img_%04d/" % i
            f[base + "contig"] = np.array([b"chr20"], dtype="S")
            f[base + "contig_start"] = np.array([i * 1000])
            f[base + "contig_end"] = np.array([i * 1000 + length])
            f[base + "feature_chunk_idx"] = np.array([i])
            f[base + "image"] = image
            f[base + "position"] = position
            if with_labels:
                f[base + "label_base"] = rng.integers(0, 5, (length, 1))
                f[base + "label_run_length"] = rng.integers(0, 11, (length, 1))
            images.append((image, position))
    return images


def test_sequence_dataset_reads_marginpolish_layout(tmp_path):
    from helen_b200.models.dataloader_predict import SequenceDataset
    path = str(tmp_path / "images.h5")
    images = _write_marginpolish_like(path, 7)
    data = SequenceDataset(None, file_list=[path])
    assert len(data) == 7
    for i in range(7):
        contig, start, end, chunk_id, image, position, filename = data[i]
        ref_image, ref_position = images[i]
        assert (contig, start, chunk_id, filename) == ("chr20", i * 1000, i, path) and end == i * 1000 + len(ref_image)
        assert image.shape == (1000, 90) and image.dtype == np.uint8
        assert np.array_equal(image[:len(ref_image)], ref_image) and not image[len(ref_image):].any()
        assert np.array_equal(position[:len(ref_image)], ref_position) and (position[len(ref_image):] == -1).all()


def test_datastore_schema_round_trip(tmp_path):
    from helen_b200.DataStore import DataStore
    path = str(tmp_path / "pred_0.hdf")
    rng = np.random.default_rng(3)
    store = DataStore(path, mode="w", packed=False)
    records = []
    for i in range(20):
        position = np.stack([np.arange(1000) + 1000 * i, np.zeros(1000, np.int64), np.zeros(1000, np.int64)], 1)
        position[990:] = -1                                   # padded columns wrap to 4294967295 in the uint32 file (DataStore.py:126)
        bases, rles = rng.integers(0, 5, 1000), rng.integers(0, 11, 1000)
        store.write_prediction("chr7", 1000 * (i // 2), 1000 * (i // 2) + 1999, i % 2, position, bases, rles)
        records.append((position, bases, rles))
    store.close()
    with hdf5.open_file(path, "r") as f:
        assert list(f["predictions"].keys()) == ["chr7"]
        assert len(f["predictions/chr7"].keys()) == 10
        for i, (position, bases, rles) in enumerate(records):
            region = "chr7-%d-%d" % (1000 * (i // 2), 1000 * (i // 2) + 1999)
            chunk = f["predictions"]["chr7"][region][str(i % 2)]
            assert chunk["position"][()].dtype == np.uint32 and chunk["bases"][()].dtype == np.uint8
            assert np.array_equal(chunk["position"][()], position.astype(np.uint32))
            assert np.array_equal(chunk["bases"][()], bases) and np.array_equal(chunk["rles"][()], rles)
            assert f["predictions"]["chr7"][region]["contig_start"][()] == 1000 * (i // 2)


def test_bulk_batches_equal_the_item_reader(tmp_path):
    """models/bulk_reader.BulkImageBatches (a block of images per item) against SequenceDataset + default collation."""
    import torch
    from torch.utils.data import DataLoader
    from helen_b200.models.bulk_reader import BulkImageBatches
    from helen_b200.models.dataloader_predict import SequenceDataset
    paths = [str(tmp_path / "a.h5"), str(tmp_path / "b.h5")]
    _write_marginpolish_like(paths[0], 11)
    _write_marginpolish_like(paths[1], 5)
    items = list(DataLoader(SequenceDataset(None, file_list=paths), batch_size=1, shuffle=False))
    bulk = BulkImageBatches(None, file_list=paths, batch_size=4)
    assert bulk.total_images == 16 and len(bulk) == 3 + 2          # a batch never spans two files
    k = 0
    for batch in DataLoader(bulk, batch_size=None, shuffle=False, num_workers=2):
        contig, start, end, chunk_id, images, position, filename = batch
        for i in range(images.shape[0]):
            ref = items[k]
            assert contig[i] == ref[0][0] and int(start[i]) == int(ref[1]) and int(end[i]) == int(ref[2]) and int(chunk_id[i]) == int(ref[3])
            assert torch.equal(images[i], ref[4][0]) and torch.equal(position[i], ref[5][0]) and filename[i] == ref[6][0]
            k += 1
    assert k == 16


def test_open_predictions_reuses_the_reader_and_notices_rewrites(tmp_path, monkeypatch):
    """The stitch opens a prediction file once per region (Stitch.py:214-245): with the package's own reader that is one
    shared handle per process - and a file rewritten under the same name is read anew."""
    from helen_b200 import DataStore as ds
    monkeypatch.setenv("HELEN_B200_HDF5", "minih5")
    ds.forget_packed_views()
    path = str(tmp_path / "p.hdf")

    def write(value):
        store = ds.DataStore(path, mode="w", packed=False)
        position = np.zeros((2, 1000, 3), np.int64)
        store.write_predictions(["chr1", "chr1"], [0, 0], [1999, 1999], [0, 1], position, np.full((2, 1000), value), np.zeros((2, 1000)))
        store.close()

    write(1)
    with ds.open_predictions(path) as a:
        first = a
        assert int(a["predictions/chr1/chr1-0-1999/1/bases"][()][5]) == 1
    with ds.open_predictions(path) as b:
        assert b is first                                   # not reopened, not closed by the first block
        assert sorted(b["predictions"]["chr1"]["chr1-0-1999"].keys()) == ["0", "1", "contig_end", "contig_start"]
    import os, time
    time.sleep(0.01)
    write(3)
    os.utime(path, ns=(time.time_ns(), time.time_ns()))
    with ds.open_predictions(path) as c:
        assert c is not first and int(c["predictions/chr1/chr1-0-1999/0/bases"][()][5]) == 3
    ds.forget_packed_views()
