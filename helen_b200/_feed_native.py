"""ctypes binding of include/helen_feed.h (host-side input feed).

`ImageFile` lists the images of a MarginPolish HDF5 file and fills the arrays of a whole batch in one native call.
`Unsupported` means "a valid file outside the library's subset of the format": callers fall back to the general reader
(helen_b200.hdf5).  A missing library raises from load(), like the other bindings.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int64, c_void_p

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libhelen_feed.so")

HF_ABI_VERSION = 2
HF_OK, HF_UNSUPPORTED, HF_E_SIZE, HF_E_FORMAT, HF_E_ARGUMENT = 0, 1, 2, 3, 4
CONTIG_STRIDE = 256

# every symbol include/helen_feed.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "hf_abi_version": (c_int, []),
    "hf_open": (c_int, [c_char_p, POINTER(c_void_p), c_char_p, c_int]),
    "hf_close": (None, [c_void_p]),
    "hf_image_count": (c_int64, [c_void_p]),
    "hf_image_names": (c_int64, [c_void_p, c_char_p, c_int64]),
    "hf_image_features": (c_int, [c_void_p, c_int64, POINTER(c_int), c_char_p, c_int]),
    "hf_read_block": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_int, c_int, c_char_p, c_int]),
    "hf_list_predictions": (c_int, [c_void_p, c_char_p, c_char_p, c_int64, c_void_p, c_void_p, c_int64, POINTER(c_int64), POINTER(c_int64), c_char_p, c_int]),
    "hf_read_prediction_region": (c_int, [c_void_p, c_char_p, c_char_p, c_int64, c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_char_p, c_int]),
}

_lib = None


class Unsupported(Exception):
    """The file is valid HDF5 but uses a feature the native reader does not implement."""


def load():
    """Load libhelen_feed.so and bind every declared symbol; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} has not been built (run `python -m helen_b200.build` or __graft_entry__.build())")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = restype, argtypes
    if lib.hf_abi_version() != HF_ABI_VERSION:
        raise RuntimeError(f"{LIB_PATH}: ABI version {lib.hf_abi_version()}, binding expects {HF_ABI_VERSION}")
    _lib = lib
    return lib


def _raise(status, message, path):
    if status == HF_UNSUPPORTED:
        raise Unsupported(message)
    if status == HF_E_SIZE:
        raise ValueError(message)                      # "IMAGE SIZE ERROR: ..." (dataloader_predict.py:85-86)
    if status == HF_E_ARGUMENT:
        raise ValueError(f"helen_feed: {message}")
    raise IOError(f"helen_feed: {message or path}")


class ImageFile(object):
    """One MarginPolish image file, memory-mapped by the native library."""

    def __init__(self, path):
        self._lib = load()
        self.path = path
        self._handle = c_void_p()
        err = ctypes.create_string_buffer(512)
        status = self._lib.hf_open(os.fsencode(path), ctypes.byref(self._handle), err, len(err))
        if status != HF_OK:
            self._handle = None
            _raise(status, err.value.decode(errors="replace"), path)

    def close(self):
        # the handle is forgotten BEFORE the library frees it, and nothing here needs module globals: close() also runs
        # from __del__ at interpreter shutdown, where a failing statement after hf_close would leave a dangling handle
        # for the next close() to free again
        handle, self._handle = getattr(self, "_handle", None), None
        if handle is not None and handle.value:
            self._lib.hf_close(handle)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _open_handle(self):
        if self._handle is None:
            raise ValueError(f"{self.path}: image file is closed")
        return self._handle

    def __len__(self):
        return int(self._lib.hf_image_count(self._open_handle()))

    def names(self):
        need = int(self._lib.hf_image_names(self._open_handle(), None, 0))
        if need == 0:
            return []
        buf = ctypes.create_string_buffer(need)
        self._lib.hf_image_names(self._handle, buf, need)
        return [n.decode() for n in buf.raw[:need].split(b"\0")[:-1]]

    def features(self, index):
        out, err = c_int(0), ctypes.create_string_buffer(512)
        status = self._lib.hf_image_features(self._open_handle(), index, ctypes.byref(out), err, len(err))
        if status != HF_OK:
            _raise(status, err.value.decode(errors="replace"), self.path)
        return out.value

    def read_block(self, first, count, seq_len, threads=1, out=None):
        """-> (contigs list[str], start i64[n], end i64[n], chunk_id i64[n], images u8[n, seq, F], position i64[n, seq, 3])
        `out(count, features)` may supply the five arrays (C-contiguous, e.g. views of page-locked buffers) to fill."""
        features = self.features(first)
        if out is not None:
            images, position, starts, ends, chunk_ids = out(count, features)
            for a, shape, dtype in ((images, (count, seq_len, features), np.uint8), (position, (count, seq_len, 3), np.int64),
                                    (starts, (count,), np.int64), (ends, (count,), np.int64), (chunk_ids, (count,), np.int64)):
                if a.shape != shape or a.dtype != dtype or not a.flags.c_contiguous:
                    raise ValueError("helen_feed: output array of the wrong shape, type or layout")
        else:
            images = np.empty((count, seq_len, features), np.uint8)
            position = np.empty((count, seq_len, 3), np.int64)
            starts, ends, chunk_ids = np.empty(count, np.int64), np.empty(count, np.int64), np.empty(count, np.int64)
        contigs = np.zeros((count, CONTIG_STRIDE), np.uint8)
        err = ctypes.create_string_buffer(512)
        status = self._lib.hf_read_block(self._open_handle(), first, count, seq_len, features, images.ctypes.data, position.ctypes.data,
                                         starts.ctypes.data, ends.ctypes.data, chunk_ids.ctypes.data, contigs.ctypes.data, CONTIG_STRIDE,
                                         int(threads), err, len(err))
        if status != HF_OK:
            _raise(status, err.value.decode(errors="replace"), self.path)
        names = [bytes(row).split(b"\0", 1)[0].decode() for row in contigs]
        return names, starts, ends, chunk_ids, images, position

    def read_prediction_region(self, contig, region, capacity_rows=8000):
        """-> (position i64[rows, 3], bases u8[rows], rles u8[rows]) of predictions/<contig>/<region>, chunks in string order."""
        err, total = ctypes.create_string_buffer(512), c_int64(0)
        while True:
            position = np.empty((capacity_rows, 3), np.int64)
            bases, rles = np.empty(capacity_rows, np.uint8), np.empty(capacity_rows, np.uint8)
            status = self._lib.hf_read_prediction_region(self._open_handle(), str(contig).encode(), str(region).encode(), capacity_rows,
                                                         position.ctypes.data, bases.ctypes.data, rles.ctypes.data, ctypes.byref(total), err, len(err))
            if status != HF_OK:
                _raise(status, err.value.decode(errors="replace"), self.path)
            if total.value <= capacity_rows:
                return position[:total.value], bases[:total.value], rles[:total.value]
            capacity_rows = total.value

    def list_predictions(self, contig=None):
        """contig None -> [contig names]; else ([region names], contig_start i64[n], contig_end i64[n]) in file order."""
        count, need, err = c_int64(0), c_int64(0), ctypes.create_string_buffer(512)
        key = None if contig is None else str(contig).encode()
        status = self._lib.hf_list_predictions(self._open_handle(), key, None, 0, None, None, 0, ctypes.byref(count), ctypes.byref(need), err, len(err))
        if status != HF_OK:
            _raise(status, err.value.decode(errors="replace"), self.path)
        names = ctypes.create_string_buffer(max(need.value, 1))
        starts, ends = np.empty(count.value, np.int64), np.empty(count.value, np.int64)
        status = self._lib.hf_list_predictions(self._open_handle(), key, names, need.value, starts.ctypes.data, ends.ctypes.data, count.value,
                                               ctypes.byref(count), ctypes.byref(need), err, len(err))
        if status != HF_OK:
            _raise(status, err.value.decode(errors="replace"), self.path)
        listed = [n.decode("utf-8") for n in names.raw[:need.value].split(b"\0")[:-1]]
        return listed if contig is None else (listed, starts, ends)
