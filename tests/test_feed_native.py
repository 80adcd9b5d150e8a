"""Native input feed (include/helen_feed.h) against the general reader on the same files.

The native library restates, for whole batches and in C++, what SequenceDataset does per image
(reference helen/modules/python/models/dataloader_predict.py:54-88); helen_b200.minih5 is the checker here: both must
list the images in the same order and return identical arrays, paddings and contig names."""
import os
import re

import numpy as np
import pytest
import torch

from helen_b200 import _feed_native, hdf5
from helen_b200.models.bulk_reader import BulkImageBatches

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_images(path, n_images, features=90, position_dtype=np.int64, scalar_dtype=np.int64, contig=b"chr20", seed=7, long_image=None):
    rng = np.random.default_rng(seed)
    with hdf5.open_file(path, "w") as f:
        for i in range(n_images):
            length = 1000 if i % 3 else 700 + (i % 5)
            if long_image == i:
                length = 1001
            image = rng.integers(0, 256, (length, features), dtype=np.uint8)
            position = np.stack([np.arange(length) + i * 1000, rng.integers(0, 4, length), np.zeros(length, np.int64)], 1).astype(position_dtype)
            base = "images/img_%05d/" % ((i * 7919) % 100000)        # names out of creation order: the file order is by name
            f[base + "contig"] = np.array([contig], dtype="S")
            f[base + "contig_start"] = np.array([i * 1000], dtype=scalar_dtype)
            f[base + "contig_end"] = np.array([i * 1000 + length], dtype=scalar_dtype)
            f[base + "feature_chunk_idx"] = np.array([i], dtype=scalar_dtype)
            f[base + "image"] = image
            f[base + "position"] = position


def test_library_exports_what_the_header_declares():
    header = open(os.path.join(ROOT, "include", "helen_feed.h")).read()
    declared = set(re.findall(r"\b(hf_[a-z_]+)\s*\(", header))
    assert declared == set(_feed_native.SIGNATURES), declared ^ set(_feed_native.SIGNATURES)
    lib = _feed_native.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.hf_abi_version() == _feed_native.HF_ABI_VERSION


@pytest.mark.parametrize("n_images,position_dtype,scalar_dtype,features", [(11, np.int64, np.int64, 90), (300, np.int32, np.int32, 10),
                                                                            (29, np.uint32, np.int16, 33)])
def test_native_batches_equal_the_general_reader(tmp_path, monkeypatch, n_images, position_dtype, scalar_dtype, features):
    monkeypatch.setenv("HELEN_B200_HDF5", "minih5")
    paths = [str(tmp_path / "a.h5"), str(tmp_path / "b.h5")]
    write_images(paths[0], n_images, features, position_dtype, scalar_dtype)
    write_images(paths[1], 5, features, position_dtype, scalar_dtype, contig=b"chr'X'_alt", seed=9)
    native = BulkImageBatches(None, file_list=paths, batch_size=64, native=True, threads=3)
    general = BulkImageBatches(None, file_list=paths, batch_size=64, native=False)
    assert native._native_paths == set(paths) and not general._native_paths
    assert native.blocks == general.blocks and native._names == general._names        # same images in the same order
    assert len(native._names[paths[0]]) == n_images
    for index in range(len(native)):
        a, b = native[index], general[index]
        assert a[0] == b[0] and a[6] == b[6]
        for x, y in zip(a[1:6], b[1:6]):
            assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y)
    assert native[len(native) - 1][0][0] == "chrX_alt"                                 # apostrophes removed (dataloader_predict.py:70)
    native.close()
    general.close()


def test_native_feed_through_worker_processes(tmp_path):
    from torch.utils.data import DataLoader
    path = str(tmp_path / "w.h5")
    write_images(path, 40, 10)
    want = list(DataLoader(BulkImageBatches(None, file_list=[path], batch_size=16, native=False), batch_size=None))
    got = list(DataLoader(BulkImageBatches(None, file_list=[path], batch_size=16, native=True), batch_size=None, num_workers=2))
    assert len(got) == len(want) == 3
    for a, b in zip(got, want):
        assert list(a[0]) == list(b[0]) and all(torch.equal(x, y) for x, y in zip(a[1:6], b[1:6]))


def test_errors_and_fallback(tmp_path):
    junk = str(tmp_path / "junk.h5")
    with open(junk, "wb") as f:
        f.write(b"not an hdf5 file at all" * 10)
    with pytest.raises(IOError):
        _feed_native.ImageFile(junk)
    long_path = str(tmp_path / "long.h5")
    write_images(long_path, 4, 10, long_image=2)
    with pytest.raises(ValueError, match="IMAGE SIZE ERROR"):
        BulkImageBatches(None, file_list=[long_path], batch_size=8, native=True)[0]
    with pytest.raises(ValueError, match="IMAGE SIZE ERROR"):
        BulkImageBatches(None, file_list=[long_path], batch_size=8, native=False)[0]
    # a file without /images: zero images, skipped with a warning by both readers
    empty = str(tmp_path / "empty.h5")
    with hdf5.open_file(empty, "w") as f:
        f["other/x"] = np.arange(3)
    assert len(_feed_native.ImageFile(empty)) == 0
    assert len(BulkImageBatches(None, file_list=[empty], native=True)) == 0
    # truncated file: the library refuses to read past the mapping instead of faulting
    whole = open(long_path, "rb").read()
    cut = str(tmp_path / "cut.h5")
    with open(cut, "wb") as f:
        f.write(whole[:len(whole) // 3])
    try:
        handle = _feed_native.ImageFile(cut)
        with pytest.raises((IOError, _feed_native.Unsupported)):
            handle.read_block(0, len(handle), 1000, 2)
    except IOError:
        pass


def test_chunked_file_takes_the_general_reader(tmp_path):
    """A dataset layout outside the subset (chunked + deflate) is refused with `Unsupported`, not misread."""
    from test_minih5 import _assemble_chunked_file
    path = str(tmp_path / "c.h5")
    _assemble_chunked_file(path, np.arange(1000 * 90, dtype=np.uint16).reshape(1000, 90), (256, 32), True, False)
    handle = _feed_native.ImageFile(path)                     # opens: the root group is fine, there is just no /images
    assert len(handle) == 0


def test_ring_buffers_and_thread_prefetch(tmp_path):
    """The driver's feed (models/predict_gpu.py): native reads into a ring of reused buffers, one background thread ahead."""
    from helen_b200.models.prefetch import ThreadPrefetcher
    paths = [str(tmp_path / "a.h5"), str(tmp_path / "b.h5")]
    write_images(paths[0], 50, 10)
    write_images(paths[1], 21, 10, seed=3)
    general = BulkImageBatches(None, file_list=paths, batch_size=8, native=False)
    ringed = BulkImageBatches(None, file_list=paths, batch_size=8, native=True, threads=2, ring=5)
    with pytest.raises(ValueError, match="too small"):
        ThreadPrefetcher(BulkImageBatches(None, file_list=paths, batch_size=8, native=True, ring=4), depth=2)
    assert ringed.native_for_all() and not general.native_for_all()
    seen = 0
    held = []
    for index, batch in enumerate(ThreadPrefetcher(ringed, depth=2)):
        want = general[index]
        assert batch[0] == want[0] and all(torch.equal(x, y) for x, y in zip(batch[1:6], want[1:6]))
        held.append((batch, want))
        if len(held) > 1:                                   # the previous batch is still intact (the driver's in-flight batch)
            old, old_want = held.pop(0)
            assert all(torch.equal(x, y) for x, y in zip(old[1:6], old_want[1:6]))
        seen += batch[4].shape[0]
    assert seen == 71 and len(ringed._ring_sets) == 5
    # a consumer that stops early leaves no thread behind a full queue
    it = iter(ThreadPrefetcher(ringed, depth=1))
    next(it)
    it.close()


def test_prefetcher_passes_errors_on(tmp_path):
    from helen_b200.models.prefetch import ThreadPrefetcher
    path = str(tmp_path / "long.h5")
    write_images(path, 12, 10, long_image=9)
    data = BulkImageBatches(None, file_list=[path], batch_size=4, native=True, ring=5)
    with pytest.raises(ValueError, match="IMAGE SIZE ERROR"):
        list(ThreadPrefetcher(data, depth=2))


def test_corrupted_files_do_not_crash_the_reader(tmp_path):
    """The library parses files it did not write: every offset and length it reads from the file is checked against the
    mapping, so damaged metadata must end in an error status, never in a fault.  300 files with random bytes of the
    structure region overwritten (and some truncated) are read in a child process; it has to exit normally."""
    import subprocess
    import sys
    good = str(tmp_path / "good.h5")
    write_images(good, 12, 10)
    data = bytearray(open(good, "rb").read())
    rng = np.random.default_rng(2024)
    meta_start = max(96, len(data) - 30000)                     # raw data comes first in files of this writer; the structure is the tail
    paths = []
    for k in range(300):
        mutated = bytearray(data)
        region = (0, 96) if k % 10 == 0 else (meta_start, len(data))
        for _ in range(int(rng.integers(1, 6))):
            at = int(rng.integers(region[0], region[1]))
            mutated[at] = int(rng.integers(0, 256))
        if k % 7 == 0:
            mutated = mutated[:int(rng.integers(meta_start, len(data)))]
        path = str(tmp_path / ("m%03d.h5" % k))
        with open(path, "wb") as f:
            f.write(mutated)
        paths.append(path)
    # ... and a prediction file (the stitch's listing and per-region reads search and parse the same structures)
    from helen_b200.DataStore import DataStore
    pred = str(tmp_path / "pred.hdf")
    store = DataStore(pred, mode="w", packed=False)
    ids = np.arange(300)
    store.write_predictions(["chr%d" % (i % 2) for i in ids], ids * 100, ids * 100 + 99, ids % 3, np.zeros((300, 20, 3), np.int64),
                            np.ones((300, 20), np.uint8), np.ones((300, 20), np.uint8))
    store.close()
    data = bytearray(open(pred, "rb").read())
    meta_start = 96 + 300 * 20 * (12 + 2)
    for k in range(300):
        mutated = bytearray(data)
        for _ in range(int(rng.integers(1, 6))):
            mutated[int(rng.integers(meta_start, len(data)))] = int(rng.integers(0, 256))
        if k % 9 == 0:
            mutated = mutated[:int(rng.integers(meta_start, len(data)))]
        path = str(tmp_path / ("p%03d.hdf" % k))
        with open(path, "wb") as f:
            f.write(mutated)
        paths.append(path)
    paths.append(pred)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    child = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "feed_fuzz_child.py")] + paths, capture_output=True, text=True, env=env, timeout=300)
    assert child.returncode == 0, "reader crashed (exit %d)\n%s" % (child.returncode, child.stderr[-2000:])
    assert "'ok'" in child.stdout


def test_prediction_regions_through_the_library_equal_the_python_reader(tmp_path, monkeypatch):
    """hf_read_prediction_region (the stitch's per-region read, Stitch.py:214-245) against the package's Python reader on a
    real file: many regions (several B-tree levels to search), chunk names that sort as strings ("10" before "2"), the
    uint32 wrap of padded positions, a repeated record; then the whole stitch with and without the library."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import stitch_inputs
    import helen_b200.StitchInterface as iface
    from helen_b200 import DataStore as ds
    from helen_b200.Stitch import decode_region
    monkeypatch.setenv("HELEN_B200_HDF5", "minih5")
    ds.forget_packed_views()
    records = []
    for k in range(3):
        records += stitch_inputs.prediction_records(seed=100 + k, regions=12, images_per_region=3 if k % 2 else 11, contig="chr%d" % k)
    records.append(records[5])                                      # a repeated (region, chunk): the first one wins
    path = str(tmp_path / "pred_0.hdf")
    store = ds.DataStore(path, mode="w", packed=False)
    for lo in range(0, len(records), 64):
        batch = records[lo:lo + 64]
        store.write_predictions([r[0] for r in batch], [r[1] for r in batch], [r[2] for r in batch], [r[3] for r in batch],
                                np.stack([r[4] for r in batch]), np.stack([r[5] for r in batch]), np.stack([r[6] for r in batch]))
    store.close()
    native = _feed_native.ImageFile(path)
    checked = 0
    with hdf5.open_file(path, "r") as f:
        for contig in f["predictions"].keys():
            regions = f["predictions"][contig].keys()
            for region_name in regions:
                region = f["predictions"][contig][region_name]
                chunks = sorted(set(region.keys()) - {"contig_start", "contig_end"})
                want_p = np.concatenate([np.asarray(region[c]["position"][()], dtype=np.int64).reshape(-1, 3) for c in chunks])
                want_b = np.concatenate([np.asarray(region[c]["bases"][()]).reshape(-1) for c in chunks])
                want_r = np.concatenate([np.asarray(region[c]["rles"][()]).reshape(-1) for c in chunks])
                got_p, got_b, got_r = native.read_prediction_region(contig, region_name, capacity_rows=1500)   # forces the retry with a larger buffer
                assert np.array_equal(got_p, want_p) and np.array_equal(got_b, want_b) and np.array_equal(got_r, want_r)
                assert decode_region(got_p, got_b, got_r) == decode_region(want_p, want_b, want_r)
                checked += 1
                if len(chunks) > 10:
                    assert chunks.index("10") < chunks.index("2")
    assert checked >= 36
    with hdf5.open_file(path, "r") as f:                              # the stitch's first pass (StitchInterface.py:52-66)
        assert native.list_predictions() == list(f["predictions"].keys())
        for contig in f["predictions"].keys():
            names, starts, ends = native.list_predictions(contig)
            assert names == list(f["predictions"][contig].keys())
            assert starts.tolist() == [int(f["predictions"][contig][r]["contig_start"][()]) for r in names]
            assert ends.tolist() == [int(f["predictions"][contig][r]["contig_end"][()]) for r in names]
    with pytest.raises(_feed_native.Unsupported):
        native.list_predictions("chrNone")
    for missing in (("chrNone", "x"), ("chr0", "chr0-1-2")):
        with pytest.raises(_feed_native.Unsupported):
            native.read_prediction_region(*missing)
    native.close()
    # a contig with thousands of regions: the region is found by searching a B-tree of several levels
    big = str(tmp_path / "big_0.hdf")
    store = ds.DataStore(big, mode="w", packed=False)
    n = 3000
    starts = (np.arange(n) * 7919) % 1000003 * 10
    tiny_position = (np.arange(n * 15).reshape(n, 5, 3) % 4000000000).astype(np.int64)
    tiny_position[:, 4] = -1
    tiny_bases, tiny_rles = (np.arange(n * 5).reshape(n, 5) % 5), (np.arange(n * 5).reshape(n, 5) % 11)
    store.write_predictions(["chrBig"] * n, starts, starts + 9, np.zeros(n, np.int64), tiny_position, tiny_bases, tiny_rles)
    store.close()
    native = _feed_native.ImageFile(big)
    for i in list(range(0, n, 37)) + [n - 1]:
        p_, b_, r_ = native.read_prediction_region("chrBig", "chrBig-%d-%d" % (starts[i], starts[i] + 9))
        assert np.array_equal(p_, tiny_position[i].astype(np.uint32).astype(np.int64)) and np.array_equal(b_, tiny_bases[i]) and np.array_equal(r_, tiny_rles[i])
    with pytest.raises(_feed_native.Unsupported):
        native.read_prediction_region("chrBig", "chrBig-5-14")
    native.close()
    fasta = {}
    for use in ("1", "0"):
        monkeypatch.setenv("HELEN_B200_NATIVE_READER", use)
        ds.forget_packed_views()
        monkeypatch.setattr(iface, "get_file_paths_from_directory", lambda directory: [path])
        fasta[use] = open(iface.perform_stitch(str(tmp_path), str(tmp_path / ("out" + use)), "p", 3)).read()
    assert fasta["1"] == fasta["0"] and fasta["1"].count(">") == 3
    ds.forget_packed_views()
