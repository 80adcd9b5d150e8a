"""Input feed: MarginPolish image files -> ready uint8 batches, a block of images per request.

The reference's SequenceDataset (helen/modules/python/models/dataloader_predict.py:54-88) opens the HDF5 file for every
image and returns one image per item; the DataLoader then collates `batch_size` items.  At the rate the GPU path
consumes windows (~90 k/s, 8 GB/s of pixels at 90 features) the per-item work has to go: here an item IS a batch -
`batch_size` consecutive images of one file, read in one pass from the file held open (memory-mapped by
helen_b200.minih5, or through h5py), padded and stacked into [n, 1000, F] / [n, 1000, 3] arrays once.  Used with
``DataLoader(batch_size=None, num_workers=N, pin_memory=True)`` the worker processes hand the stacked tensors over
through shared memory and the main process only pins them.

Where the file keeps to what libhdf5 writes by default (contiguous little-endian datasets in symbol-table groups) the
batch is filled by the native feed library (include/helen_feed.h, helen_b200/csrc_host/feed_host.cpp: the file mapped
once, `HELEN_B200_FEED_THREADS` host threads per batch, no Python per image); any other file, or
HELEN_B200_NATIVE_FEED=0, takes the general reader (h5py or minih5) with identical results (tests/test_feed_native.py).

Order and content are those of SequenceDataset: files in list order, images in `images.keys()` order, short images
right-padded with zero columns and (-1, -1, -1) positions; a batch never spans two files (the last batch of a file may be
short), which changes nothing in the prediction files since every image is predicted on its own.
"""
import os
import sys

import numpy as np
import torch
from torch.utils.data import Dataset

from .. import _feed_native, hdf5
from ..FileManager import FileManager
from ..options import ImageSizeOptions
from ..TextColor import TextColor


def _scalar(dataset_value):
    return np.asarray(dataset_value).reshape(-1)[0]


class BulkImageBatches(Dataset):
    """Item i -> (contig list, contig_start i64[n], contig_end i64[n], chunk_id i64[n], images u8[n, 1000, F],
    position i64[n, 1000, 3], path list): the tuple predict() takes from the reference's DataLoader."""

    def __init__(self, image_directory, file_list=None, batch_size=512, native=None, threads=None, ring=0):
        hdf_files = file_list if file_list is not None else FileManager.get_file_paths_from_directory(image_directory)
        self.blocks = []                               # (path, first image, count)
        self._names = {}                               # path -> image names in file order
        self.batch_size = max(int(batch_size), 1)
        self.total_images = 0
        self.native = (os.environ.get("HELEN_B200_NATIVE_FEED", "1") != "0") if native is None else bool(native)
        self.threads = int(os.environ.get("HELEN_B200_FEED_THREADS", "4")) if threads is None else int(threads)
        self._native_paths = set()                     # files the native library reads (the others go through helen_b200.hdf5)
        self._native_file = None
        # ring > 0: the native reader fills `ring` sets of preallocated (page-locked where CUDA is present) tensors in
        # rotation instead of fresh ones - no page faults, no pinning copy.  The consumer must be done with a batch before
        # `ring - 1` further ones have been requested (models/predict_gpu.py keeps at most five alive).
        self.ring = int(ring)
        self._ring_sets, self._ring_next = [], 0
        for path in hdf_files:
            names = self._native_names(path) if self.native else None
            if names is None:
                with hdf5.open_file(path, 'r') as handle:
                    names = list(handle['images'].keys()) if 'images' in handle else []
            if not names:
                sys.stderr.write(TextColor.YELLOW + "WARN: NO IMAGES FOUND IN FILE: " + path + "\n" + TextColor.END)
                continue
            self._names[path] = names
            self.total_images += len(names)
            for first in range(0, len(names), self.batch_size):
                self.blocks.append((path, first, min(self.batch_size, len(names) - first)))
        self._open_path, self._open_file, self._open_pid = None, None, None

    def _native_names(self, path):
        """Image names through the native library, or None if the file is outside its subset."""
        try:
            f = _feed_native.ImageFile(path)
        except (_feed_native.Unsupported, IOError):
            return None
        try:
            names = f.names()
            if names:
                f.features(0)                          # the datasets themselves are inside the subset, too
            self._native_paths.add(path)
            return names
        except _feed_native.Unsupported:
            return None
        finally:
            f.close()

    def __len__(self):
        return len(self.blocks)

    def _file(self, path):
        if self._open_pid != os.getpid():              # a handle inherited through fork is not ours to use or close
            self._open_path, self._open_file, self._open_pid = None, None, os.getpid()
        if self._open_path != path:
            self.close()
            self._open_file, self._open_path = hdf5.open_file(path, 'r'), path
        return self._open_file

    def close(self):
        if self._open_file is not None and self._open_pid == os.getpid():
            self._open_file.close()
        self._open_path, self._open_file = None, None
        if self._native_file is not None and self._native_file[0] == os.getpid():
            self._native_file[2].close()
        self._native_file = None

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_open_path"], state["_open_file"], state["_native_file"] = None, None, None
        state["_ring_sets"], state["_ring_next"] = [], 0
        return state

    def native_for_all(self):
        """True when every listed file is read by the native library (then no worker processes are needed)."""
        return self.native and set(self._names) <= self._native_paths

    def _ring_arrays(self, count, features):
        seq = ImageSizeOptions.SEQ_LENGTH
        if len(self._ring_sets) < self.ring:
            pin = torch.cuda.is_available()
            self._ring_sets.append((features,
                                    torch.empty((self.batch_size, seq, features), dtype=torch.uint8, pin_memory=pin),
                                    torch.empty((self.batch_size, seq, 3), dtype=torch.int64, pin_memory=pin),
                                    torch.empty((3, self.batch_size), dtype=torch.int64)))
        slot = self._ring_next % len(self._ring_sets)
        self._ring_next += 1
        f, images, position, scalars = self._ring_sets[slot]
        if f != features or count > self.batch_size:    # a file with another feature count: plain arrays for this batch
            return (np.empty((count, seq, features), np.uint8), np.empty((count, seq, 3), np.int64),
                    np.empty(count, np.int64), np.empty(count, np.int64), np.empty(count, np.int64))
        return (images[:count].numpy(), position[:count].numpy(), scalars[0, :count].numpy(), scalars[1, :count].numpy(),
                scalars[2, :count].numpy())

    def _native_handle(self, path):
        cur = self._native_file                        # (pid, path, handle): a handle inherited through fork is not ours
        if cur is None or cur[0] != os.getpid() or cur[1] != path:
            if cur is not None and cur[0] == os.getpid():
                cur[2].close()
            self._native_file = (os.getpid(), path, _feed_native.ImageFile(path))
        return self._native_file[2]

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __getitem__(self, index):
        path, first, count = self.blocks[index]
        seq = ImageSizeOptions.SEQ_LENGTH
        if path in self._native_paths:
            try:
                contigs, starts, ends, chunk_ids, images, position = self._native_handle(path).read_block(
                    first, count, seq, self.threads, out=self._ring_arrays if self.ring > 0 else None)
                return (contigs, torch.from_numpy(starts), torch.from_numpy(ends), torch.from_numpy(chunk_ids),
                        torch.from_numpy(images), torch.from_numpy(position), [path] * count)
            except _feed_native.Unsupported:           # a later image of the file steps outside the subset
                self._native_paths.discard(path)
        names = self._names[path][first:first + count]
        groups = self._file(path)['images']
        images = position = None
        contigs, starts, ends, chunk_ids = [], np.empty(count, np.int64), np.empty(count, np.int64), np.empty(count, np.int64)
        for k, name in enumerate(names):
            group = groups[name]
            contig = _scalar(group['contig'][()])
            if isinstance(contig, bytes):
                contig = contig.decode()
            contigs.append(str(contig).replace("'", ''))
            starts[k] = int(_scalar(group['contig_start'][()]))
            ends[k] = int(_scalar(group['contig_end'][()]))
            chunk_ids[k] = int(_scalar(group['feature_chunk_idx'][()]))
            image = np.asarray(group['image'][()])
            pos = np.asarray(group['position'][()])
            if images is None:                         # zero columns / (-1, -1, -1) positions are the padding of short images
                images = np.zeros((count, seq, image.shape[1]), np.uint8)
                position = np.full((count, seq, 3), -1, np.int64)
            if image.shape[0] > seq or image.shape[1] != images.shape[2] or pos.shape[0] != image.shape[0]:
                raise ValueError("IMAGE SIZE ERROR: " + str(path) + " " + str(image.shape))
            images[k, :image.shape[0]] = image
            position[k, :pos.shape[0]] = pos
        return (contigs, torch.from_numpy(starts), torch.from_numpy(ends), torch.from_numpy(chunk_ids),
                torch.from_numpy(images), torch.from_numpy(position), [path] * count)
