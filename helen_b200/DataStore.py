"""Prediction file writer with the reference's HDF5 schema (helen/modules/python/DataStore.py:83-133):

predictions/<contig>/<contig>-<start>-<end>/contig_start, contig_end            (scalars)
predictions/<contig>/<contig>-<start>-<end>/<chunk_id>/position  uint32 [1000, 3]
                                                      /bases     uint8  [1000]
                                                      /rles      uint8  [1000]

Packed mode (opt-in: ``DataStore(..., packed=True)`` or HELEN_B200_PACKED_PREDICTIONS=1): the schema above costs three
HDF5 datasets per image, which a writer cannot create at the rate the GPU predicts them.  A packed file holds ONE group
per written batch instead,

predictions_packed/<n>/contig S[B], contig_start i64[B], contig_end i64[B], chunk_id i64[B],
                       position uint32 [B, 1000, 3], bases uint8 [B, 1000], rles uint8 [B, 1000]

and is read back through ``PackedPredictions``, a view with the nested shape of the reference schema, so the stitch
code is the same for both.  Packed files are for helen_b200's own ``stitch``; the reference's cannot read them.
"""
import os

import numpy as np

from . import hdf5


def _item(value):
    return value.item() if hasattr(value, "item") else value


class DataStore(object):
    _prediction_path_ = 'predictions'

    _packed_path_ = 'predictions_packed'

    def __init__(self, filename, mode='r', packed=None):
        self.filename = filename
        self.mode = mode
        self.packed = (os.environ.get("HELEN_B200_PACKED_PREDICTIONS", "0") not in ("", "0")) if packed is None else bool(packed)
        self.file_handler = hdf5.open_file(self.filename, self.mode)
        self._written_regions = set()
        self._written_chunks = set()
        self._packed_batches = 0

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()

    def close(self):
        if self.file_handler is not None:
            self.file_handler.close()
            self.file_handler = None

    def write_prediction(self, contig, contig_start, contig_end, chunk_id, position,
                         predicted_bases, predicted_rles, filename=None):
        if self.packed:                                     # a batch of one
            return self.write_predictions([contig], [_item(contig_start)], [_item(contig_end)], [_item(chunk_id)],
                                          np.asarray(position)[None], np.asarray(predicted_bases)[None],
                                          np.asarray(predicted_rles)[None])
        contig_start, contig_end, chunk_id = _item(contig_start), _item(contig_end), _item(chunk_id)
        chunk_name_prefix = str(contig) + "-" + str(contig_start) + "-" + str(contig_end)
        chunk_name_suffix = str(chunk_id)
        name = str(contig) + chunk_name_prefix + chunk_name_suffix
        base = '{}/{}/{}'.format(self._prediction_path_, contig, chunk_name_prefix)
        if chunk_name_prefix not in self._written_regions:
            self._written_regions.add(chunk_name_prefix)
            self.file_handler[base + '/contig_start'] = contig_start
            self.file_handler[base + '/contig_end'] = contig_end
        if name not in self._written_chunks:
            self._written_chunks.add(name)
            chunk = base + '/' + chunk_name_suffix
            # the reference stores positions as uint32, so the (-1, -1, -1) padding wraps (DataStore.py:126-127)
            self.file_handler[chunk + '/position'] = np.asarray(position).astype(np.uint32)
            self.file_handler[chunk + '/bases'] = np.asarray(predicted_bases).astype(np.uint8)
            self.file_handler[chunk + '/rles'] = np.asarray(predicted_rles).astype(np.uint8)

    def _as_uint32(self, position):
        """position.astype(uint32) (DataStore.py:126 stores positions as uint32) into a buffer kept between calls: a fresh
        6 MB array per batch costs more in page faults than the conversion itself."""
        if position.dtype == np.uint32:
            return position
        buf = getattr(self, "_position_buffer", None)
        if buf is None or buf.shape[1:] != position.shape[1:] or buf.shape[0] < position.shape[0]:
            buf = self._position_buffer = np.empty(position.shape, np.uint32)
        out = buf[:position.shape[0]]
        np.copyto(out, position, casting="unsafe")
        return out

    def write_predictions(self, contig, contig_start, contig_end, chunk_id, position, predicted_bases, predicted_rles,
                          filename=None):
        """One batch of records (the per-batch loop of predict_gpu.py:176-179 in one call): same file content as
        calling write_prediction per record, with the dtype conversions done once per batch."""
        position = self._as_uint32(np.asarray(position))
        predicted_bases = np.asarray(predicted_bases).astype(np.uint8, copy=False)
        predicted_rles = np.asarray(predicted_rles).astype(np.uint8, copy=False)
        contig_start = np.asarray(contig_start).reshape(-1).tolist()
        contig_end = np.asarray(contig_end).reshape(-1).tolist()
        chunk_id = np.asarray(chunk_id).reshape(-1).tolist()
        if not (len(contig) == len(contig_start) == len(contig_end) == len(chunk_id) == len(position)
                == len(predicted_bases) == len(predicted_rles)):
            raise ValueError("write_predictions: all arguments must have one entry per record")
        if self.packed:
            group = '{}/{}'.format(self._packed_path_, self._packed_batches)
            self._packed_batches += 1
            self.file_handler[group + '/contig'] = np.array([str(c).encode() for c in contig], dtype='S')
            self.file_handler[group + '/contig_start'] = np.asarray(contig_start, dtype=np.int64)
            self.file_handler[group + '/contig_end'] = np.asarray(contig_end, dtype=np.int64)
            self.file_handler[group + '/chunk_id'] = np.asarray(chunk_id, dtype=np.int64)
            self.file_handler[group + '/position'] = position
            self.file_handler[group + '/bases'] = predicted_bases
            self.file_handler[group + '/rles'] = predicted_rles
            return
        if hasattr(self.file_handler, "set_rows"):
            # the package's own HDF5 layer: the same datasets as the loop below, with one file write per array and the
            # bookkeeping of write_prediction (first record of a region / chunk wins) done on the batch
            regions, region_rows, chunks, chunk_rows = [], [], [], []
            for i in range(len(contig)):
                prefix = "{}-{}-{}".format(contig[i], contig_start[i], contig_end[i])
                base = '{}/{}/{}'.format(self._prediction_path_, contig[i], prefix)
                if prefix not in self._written_regions:
                    self._written_regions.add(prefix)
                    regions.append(base)
                    region_rows.append(i)
                name = str(contig[i]) + prefix + str(chunk_id[i])
                if name not in self._written_chunks:
                    self._written_chunks.add(name)
                    chunks.append(base + '/' + str(chunk_id[i]))
                    chunk_rows.append(i)
            if regions:
                self.file_handler.set_rows(regions, 'contig_start', np.asarray([contig_start[i] for i in region_rows]))
                self.file_handler.set_rows(regions, 'contig_end', np.asarray([contig_end[i] for i in region_rows]))
            if chunks:
                every = len(chunk_rows) == len(contig)
                self.file_handler.set_rows(chunks, 'position', position if every else position[chunk_rows])
                self.file_handler.set_rows(chunks, 'bases', predicted_bases if every else predicted_bases[chunk_rows])
                self.file_handler.set_rows(chunks, 'rles', predicted_rles if every else predicted_rles[chunk_rows])
            return
        for i in range(len(contig)):
            self.write_prediction(contig[i], contig_start[i], contig_end[i], chunk_id[i], position[i],
                                  predicted_bases[i], predicted_rles[i])


class _Value(object):
    """A dataset of the view: value[()] like an h5py dataset."""

    def __init__(self, value):
        self.value = value

    def __getitem__(self, key):
        return self.value if key == () else self.value[key]


class PackedPredictions(object):
    """Read view of a packed prediction file: view['predictions'][contig][region][chunk]['bases'][()] and the
    region's 'contig_start' / 'contig_end', as in the reference schema.  The first record of a (region, chunk) pair
    wins, like DataStore.write_prediction's duplicate rule."""

    def __init__(self, h5file):
        contigs = {}
        packed = h5file[DataStore._packed_path_]
        for batch in sorted(packed.keys(), key=int):
            group = packed[batch]
            names = [c.decode() if isinstance(c, bytes) else str(c) for c in np.asarray(group['contig'][()]).reshape(-1)]
            starts = np.asarray(group['contig_start'][()]).reshape(-1).tolist()
            ends = np.asarray(group['contig_end'][()]).reshape(-1).tolist()
            chunk_ids = np.asarray(group['chunk_id'][()]).reshape(-1).tolist()
            position, bases, rles = group['position'][()], group['bases'][()], group['rles'][()]
            for i, name in enumerate(names):
                region = contigs.setdefault(name, {}).setdefault(
                    "{}-{}-{}".format(name, starts[i], ends[i]),
                    {'contig_start': _Value(starts[i]), 'contig_end': _Value(ends[i])})
                region.setdefault(str(chunk_ids[i]), {'position': _Value(position[i]), 'bases': _Value(bases[i]),
                                                      'rles': _Value(rles[i])})
        self._root = {DataStore._prediction_path_: contigs}

    def __contains__(self, key):
        return key in self._root

    def __getitem__(self, key):
        return self._root[key]

    def __enter__(self):
        return self

    def __exit__(self, *args):
        return False

    def close(self):
        pass


_packed_views = {}
_shared_readers = {}


class _SharedReader(object):
    """A prediction file opened once per process with the package's own reader and handed out to every `with
    open_predictions(path)`: leaving the block does not close it.  The reader keeps a group's member table once it has
    parsed it, and the stitch opens a file once per REGION (Stitch.py:214-245): reopening parsed the contig's whole table
    again each time - 1.9 ms per region with 1 k regions in the file, 8.5 ms with 8 k, quadratic in the genome size."""

    def __init__(self, handle):
        self._handle = handle

    def __enter__(self):
        return self._handle

    def __exit__(self, *args):
        return False

    def __contains__(self, key):
        return key in self._handle

    def __getitem__(self, key):
        return self._handle[key]

    def keys(self):
        return self._handle.keys()

    def close(self):
        pass


def open_predictions(path):
    """A prediction file for reading, either schema.  Views of packed files, and files read through the package's own
    HDF5 layer, are opened once per process and path (the stitch asks once per region); h5py handles are opened per call
    as in the reference."""
    if path in _packed_views:
        return _packed_views[path]
    try:
        stamp = os.stat(path)
        stamp = (stamp.st_mtime_ns, stamp.st_size)
    except OSError:
        stamp = None
    cached = _shared_readers.get(path)
    if cached is not None:
        if cached[0] == stamp:
            return cached[1]
        cached[1]._handle.close()                      # the file was rewritten since: read it anew
        del _shared_readers[path]
    handle = hdf5.open_file(path, 'r')
    if DataStore._prediction_path_ in handle or DataStore._packed_path_ not in handle:
        if stamp is not None and hdf5.backend() == "minih5" and type(handle).__module__.endswith("minih5"):
            shared = _SharedReader(handle)
            _shared_readers[path] = (stamp, shared)
            return shared
        return handle
    view = PackedPredictions(handle)
    handle.close()
    _packed_views[path] = view
    return view


_region_readers = {}


def region_reader(path):
    """The native reader of a prediction file (include/helen_feed.h: hf_read_prediction_region), opened once per process
    and file, or None: h5py is in use, the library is not built, HELEN_B200_NATIVE_READER=0, or the file is outside the
    library's subset.  The stitch reads a region's rows through it in one call without the interpreter lock; a region it
    cannot serve (a packed file, ...) raises _feed_native.Unsupported and the caller reads it through open_predictions."""
    if os.environ.get("HELEN_B200_NATIVE_READER", "1") == "0" or hdf5.backend() != "minih5":
        return None
    try:
        stamp = os.stat(path)
        stamp = (stamp.st_mtime_ns, stamp.st_size)
    except OSError:
        return None
    cached = _region_readers.get(path)
    if cached is not None and cached[0] == stamp:
        return cached[1]
    reader = None
    try:
        from . import _feed_native
        if os.path.exists(_feed_native.LIB_PATH):
            reader = _feed_native.ImageFile(path)
    except Exception:
        reader = None
    _region_readers[path] = (stamp, reader)
    return reader


def forget_packed_views():
    """Drops the per-process caches of open_predictions (tests rewrite files under the same name)."""
    _packed_views.clear()
    for _, shared in _shared_readers.values():
        shared._handle.close()
    _shared_readers.clear()
    for _, reader in _region_readers.values():
        if reader is not None:
            reader.close()
    _region_readers.clear()
