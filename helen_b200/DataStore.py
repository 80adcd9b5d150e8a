"""Prediction file writer with the reference's HDF5 schema (helen/modules/python/DataStore.py:83-133):

predictions/<contig>/<contig>-<start>-<end>/contig_start, contig_end            (scalars)
predictions/<contig>/<contig>-<start>-<end>/<chunk_id>/position  uint32 [1000, 3]
                                                      /bases     uint8  [1000]
                                                      /rles      uint8  [1000]
"""
import numpy as np

from . import hdf5


def _item(value):
    return value.item() if hasattr(value, "item") else value


class DataStore(object):
    _prediction_path_ = 'predictions'

    def __init__(self, filename, mode='r'):
        self.filename = filename
        self.mode = mode
        self.file_handler = hdf5.open_file(self.filename, self.mode)
        self._written_regions = set()
        self._written_chunks = set()

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

    def write_predictions(self, contig, contig_start, contig_end, chunk_id, position, predicted_bases, predicted_rles,
                          filename=None):
        """One batch of records (the per-batch loop of predict_gpu.py:176-179 in one call): same file content as
        calling write_prediction per record, with the dtype conversions done once per batch."""
        position = np.asarray(position).astype(np.uint32)
        predicted_bases = np.asarray(predicted_bases).astype(np.uint8)
        predicted_rles = np.asarray(predicted_rles).astype(np.uint8)
        contig_start = np.asarray(contig_start).reshape(-1).tolist()
        contig_end = np.asarray(contig_end).reshape(-1).tolist()
        chunk_id = np.asarray(chunk_id).reshape(-1).tolist()
        if not (len(contig) == len(contig_start) == len(contig_end) == len(chunk_id) == len(position)
                == len(predicted_bases) == len(predicted_rles)):
            raise ValueError("write_predictions: all arguments must have one entry per record")
        for i in range(len(contig)):
            self.write_prediction(contig[i], contig_start[i], contig_end[i], chunk_id[i], position[i],
                                  predicted_bases[i], predicted_rles[i])
