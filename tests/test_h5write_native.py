"""Native prediction-file writer (include/helen_h5write.h) against the pure-Python writer of helen_b200/minih5.py:
the same calls must give the same FILE, byte for byte, and both readers (minih5, the native feed library) read it."""
import os
import re

import numpy as np
import pytest

from helen_b200 import _h5write_native, minih5

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_what_the_header_declares():
    header = open(os.path.join(ROOT, "include", "helen_h5write.h")).read()
    declared = set(re.findall(r"\b(hw_[a-z_]+)\s*\(", header))
    assert declared == set(_h5write_native.SIGNATURES), declared ^ set(_h5write_native.SIGNATURES)
    lib = _h5write_native.load()
    assert all(hasattr(lib, name) for name in declared) and lib.hw_abi_version() == _h5write_native.HW_ABI_VERSION


def _both(tmp_path, monkeypatch, fill):
    files = {}
    for native in ("1", "0"):
        monkeypatch.setenv("HELEN_B200_NATIVE_WRITER", native)
        path = str(tmp_path / ("w%s.h5" % native))
        with minih5.File(path, "w") as f:
            assert type(f._writer).__name__ == ("_NativeWriter" if native == "1" else "_Writer")
            fill(f)
        files[native] = path
    a, b = open(files["1"], "rb").read(), open(files["0"], "rb").read()
    assert a == b, "native and Python writers differ (%d / %d bytes)" % (len(a), len(b))
    return files["1"]


def test_same_bytes_for_mixed_content(tmp_path, monkeypatch):
    rng = np.random.default_rng(0)
    big = rng.integers(0, 2 ** 31, (300, 1000, 3)).astype(np.uint32)      # a raw block above the writer's buffering threshold

    def fill(f):
        f["a/b/x"] = np.arange(10, dtype=np.int64).reshape(2, 5)
        f["a/b/y"] = np.array([b"chr20"], dtype="S")
        f["a/b/text"] = np.array(["chrM", "chr1"])
        f["a/s"] = 7
        f["a/flag"] = np.array([True, False])
        f["z"] = np.float32(1.5)
        f["d"] = np.linspace(0, 1, 7)
        f["a/c/e"] = np.zeros((0, 3), np.uint8)
        f["be"] = np.arange(5, dtype=">i4")
        f.set_rows(["r/%d" % i for i in range(300)], "pos", big)
        f.set_rows(["r/%d" % i for i in range(300)], "v", np.arange(300, dtype=np.int16))
        f.set_rows([], "nothing", np.zeros((0, 4), np.uint8))
        assert "a/b" in f and "r/299/pos" in f and "a/q" not in f and "a/b/x/deeper" not in f

    path = _both(tmp_path, monkeypatch, fill)
    with minih5.File(path) as f:
        assert np.array_equal(f["a/b/x"][()], np.arange(10).reshape(2, 5)) and f["a/s"][()] == 7
        assert f["a/b/text"][()].tolist() == [b"chrM", b"chr1"] and f["a/flag"][()].tolist() == [1, 0]
        assert f["z"][()] == np.float32(1.5) and np.array_equal(f["be"][()], np.arange(5))
        assert f["a/c/e"][()].shape == (0, 3) and len(f["r"].keys()) == 300
        assert np.array_equal(f["r/123/pos"][()], big[123]) and f["r/299/v"][()] == 299


def test_same_bytes_for_many_members(tmp_path, monkeypatch):
    """Groups with thousands of members: several symbol-table nodes and B-tree levels."""
    def fill(f):
        for i in range(2300):
            f["g/m%05d" % ((i * 7919) % 100000)] = np.int32(i)
        f.set_rows(["h/%d/%d" % (i % 40, i) for i in range(1500)], "x", np.arange(3000, dtype=np.uint8).reshape(1500, 2))

    path = _both(tmp_path, monkeypatch, fill)
    with minih5.File(path) as f:
        names = f["g"].keys()
        assert len(names) == 2300 and names == sorted(names) and len(f["h"].keys()) == 40
        assert f["g/m%05d" % ((5 * 7919) % 100000)][()] == 5 and f["h/7/47/x"][()].tolist() == [94, 95]


def test_errors_match(tmp_path, monkeypatch):
    for native in ("1", "0"):
        monkeypatch.setenv("HELEN_B200_NATIVE_WRITER", native)
        with minih5.File(str(tmp_path / ("e%s.h5" % native)), "w") as f:
            f["a/x"] = np.arange(3)
            with pytest.raises(ValueError, match="name already exists"):
                f["a/x"] = np.arange(3)
            with pytest.raises(ValueError, match="is a dataset"):
                f["a/x/y"] = 1
            with pytest.raises(ValueError, match="name already exists"):
                f.set_rows(["a", "b"], "x", np.zeros((2, 2)))
            assert "b/x" not in f                                   # the failed call wrote nothing
            with pytest.raises(TypeError):
                f["o"] = np.array([object()])
            with pytest.raises(TypeError):
                f["c"] = np.array([1 + 2j])
            with pytest.raises(ValueError):
                f.set_rows(["a"], "rows", np.zeros((2, 2)))


def test_prediction_files_are_identical_and_read_back(tmp_path, monkeypatch):
    """DataStore.write_predictions in the reference's schema (DataStore.py:83-133) through both writers."""
    from helen_b200.DataStore import DataStore
    monkeypatch.setenv("HELEN_B200_HDF5", "minih5")
    rng = np.random.default_rng(3)
    n = 64
    position = np.stack([np.arange(1000), np.zeros(1000, np.int64), np.zeros(1000, np.int64)], 1)[None].repeat(n, 0)
    position[:, 990:] = -1
    paths = {}
    for native in ("1", "0"):
        monkeypatch.setenv("HELEN_B200_NATIVE_WRITER", native)
        paths[native] = str(tmp_path / ("pred%s.hdf" % native))
        store = DataStore(paths[native], mode="w", packed=False)
        gen = np.random.default_rng(5)
        for batch in range(3):
            ids = np.arange(n) + batch * n
            bases, rles = gen.integers(0, 5, (n, 1000)), gen.integers(0, 11, (n, 1000))
            # two chunks per region, and the second batch repeats some records of the first (first one wins)
            starts = (ids // 2) * 2000 if batch != 1 else ((ids - n) // 2) * 2000
            chunk = ids % 2
            store.write_predictions(["chr%d" % (i % 3) for i in (ids if batch != 1 else ids - n)], starts, starts + 1999, chunk, position, bases, rles)
        store.close()
    assert open(paths["1"], "rb").read() == open(paths["0"], "rb").read()
    with minih5.File(paths["1"]) as f:
        assert sorted(f["predictions"].keys()) == ["chr0", "chr1", "chr2"]
        region = f["predictions/chr1"]["chr1-0-1999"]
        assert region["contig_start"][()] == 0 and region["1"]["position"][()].dtype == np.uint32
        assert region["1"]["position"][()][995, 0] == 4294967295


@pytest.mark.parametrize("native", ["1", "0"])
def test_batched_prediction_writes_equal_the_per_record_loop(tmp_path, monkeypatch, native):
    """DataStore.write_predictions' batch path (set_rows) stores what write_prediction stores record by record."""
    from helen_b200.DataStore import DataStore
    monkeypatch.setenv("HELEN_B200_HDF5", "minih5")
    monkeypatch.setenv("HELEN_B200_NATIVE_WRITER", native)
    gen = np.random.default_rng(11)
    n = 48
    position = gen.integers(-1, 5000, (n, 1000, 3))
    bases, rles = gen.integers(0, 5, (n, 1000)), gen.integers(0, 11, (n, 1000))
    contigs = ["chr%d" % (i % 2) for i in range(n)]
    starts = (np.arange(n) // 3) * 3000
    chunk = np.arange(n) % 3
    chunk[7] = chunk[6]                                              # a duplicate (region, chunk): the first record wins
    a, b = str(tmp_path / "batch.hdf"), str(tmp_path / "loop.hdf")
    store = DataStore(a, mode="w", packed=False)
    store.write_predictions(contigs[:30], starts[:30], starts[:30] + 2999, chunk[:30], position[:30], bases[:30], rles[:30])
    store.write_predictions(contigs[20:], starts[20:], starts[20:] + 2999, chunk[20:], position[20:], bases[20:], rles[20:])   # overlaps the first call
    store.close()
    store = DataStore(b, mode="w", packed=False)
    for i in list(range(30)) + list(range(20, n)):
        store.write_prediction(contigs[i], starts[i], starts[i] + 2999, chunk[i], position[i], bases[i], rles[i])
    store.close()

    def content(path):
        out = {}
        with minih5.File(path) as f:
            def walk(group, prefix):
                for key in group.keys():
                    node = group[key]
                    if isinstance(node, minih5.Group):
                        walk(node, prefix + "/" + key)
                    else:
                        out[prefix + "/" + key] = (node.dtype, np.array(node[()]))
            walk(f["predictions"], "")
        return out

    x, y = content(a), content(b)
    assert sorted(x) == sorted(y) and len(x) > 100
    for key in x:
        assert x[key][0] == y[key][0] and np.array_equal(x[key][1], y[key][1]), key
