"""Minimal in-memory stand-in for the h5py.File API used by the reader / writer (h5py is not
installed in the build container nor on the GPU box)."""
import numpy as np

_STORE = {}


class _Dataset:
    def __init__(self, value):
        self.value = np.array(value)                   # a copy, as h5py / minih5 take at assignment: the caller may reuse its buffer

    def __getitem__(self, key):
        return self.value if key == () else self.value[key]


class _Group(dict):
    def keys(self):
        return dict.keys(self)


class FakeFile:
    def __init__(self, path, mode='r'):
        self.path, self.mode = path, mode
        if mode == 'w':
            _STORE[path] = _Group()
        self.root = _STORE.setdefault(path, _Group())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        pass

    def __contains__(self, key):
        return key in self.root

    def __getitem__(self, key):
        node = self.root
        for part in key.strip('/').split('/'):
            node = node[part]
        return node

    def __setitem__(self, key, value):
        parts = key.strip('/').split('/')
        node = self.root
        for part in parts[:-1]:
            node = node.setdefault(part, _Group())
        if parts[-1] in node:
            raise ValueError("dataset exists: " + key)
        node[parts[-1]] = _Dataset(value)


def open_file(path, mode='r'):
    return FakeFile(path, mode)


def add_image(path, name, contig, start, end, chunk_idx, image, position):
    f = FakeFile(path, 'a')
    grp = f.root.setdefault('images', _Group()).setdefault(name, _Group())
    grp['contig'] = _Dataset(np.array([contig.encode()]))
    grp['contig_start'] = _Dataset(np.array([start]))
    grp['contig_end'] = _Dataset(np.array([end]))
    grp['feature_chunk_idx'] = _Dataset(np.array([chunk_idx]))
    grp['image'] = _Dataset(image)
    grp['position'] = _Dataset(position)


def add_labels(path, name, label_base, label_run_length):
    grp = FakeFile(path, 'a').root['images'][name]
    grp['label_base'] = _Dataset(np.asarray(label_base).reshape(-1, 1))          # MarginPolish stores [T, 1]
    grp['label_run_length'] = _Dataset(np.asarray(label_run_length).reshape(-1, 1))


def reset():
    _STORE.clear()
