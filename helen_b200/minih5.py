"""A small pure-NumPy reader and writer for the part of HDF5 this path touches.

Why it exists: the reference reads MarginPolish images and writes prediction files through h5py
(helen/modules/python/models/dataloader_predict.py:54-88, helen/modules/python/DataStore.py:83-133), and h5py / libhdf5
are not installed in the build container nor on the GPU boxes.  ``helen_b200.hdf5.open_file`` uses h5py when it can be
imported and this module otherwise, so ``helen polish`` runs end to end on a box without libhdf5.

What is implemented (HDF5 File Format Specification 1.x / 2.0 names in parentheses):

reading
  * superblock versions 0-3; object headers version 1 and 2 (with continuation blocks)
  * groups stored as symbol tables (B-tree v1 + local heap + SNOD nodes: what libhdf5 writes by default) and groups
    stored as compact link messages; "dense" groups (fractal heap) are refused with a clear message
  * datasets: contiguous, compact and chunked (B-tree v1) layout; filters deflate, shuffle, fletcher32
  * datatypes: fixed-point, IEEE floating point, fixed-length strings, variable-length strings (global heap)
writing (``mode='w'``)
  * a classic file any libhdf5 reads: superblock 0, version-1 object headers, symbol-table groups, contiguous datasets
    of integers / floats / fixed-length byte strings, scalars as rank-0 datasets.  Raw data goes to disk as each dataset
    is assigned; the group structure (a few dozen bytes per object) is kept in memory and written by ``close()``.

The API is the subset of h5py's that the package uses: ``File(path, mode)``, ``name in f``, ``f[name]``,
``group.keys()``, ``dataset[()]`` / ``dataset[slice]``, ``dataset.shape`` / ``.dtype``, ``f[name] = value``, ``close()``
and the context manager.  Not a general HDF5 library; status of validation: files written here are read back here
(tests/test_minih5.py) and follow the specification byte for byte as far as the author could check it without libhdf5 -
no file produced by libhdf5 was available to test the reader against.
"""
import math
import mmap
import struct
import zlib

import numpy as np

import os

SIGNATURE = b"\x89HDF\r\n\x1a\n"


def _native_writer_wanted():
    """HELEN_B200_NATIVE_WRITER=0 keeps the pure-Python writer; so does a build without libhelen_feed.so."""
    if os.environ.get("HELEN_B200_NATIVE_WRITER", "1") == "0":
        return False
    from . import _h5write_native
    return os.path.exists(_h5write_native.LIB_PATH)
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16                    # libhdf5's defaults: 2K entries per symbol node, 2K children per B-tree node


_MSG_HEAD = struct.Struct("<HHB")


class Hdf5FormatError(IOError):
    pass


# =============================================================================================
# reading
# =============================================================================================
class _Reader(object):
    def __init__(self, path):
        self.path = path
        self.fh = open(path, "rb")
        # the whole file is mapped: parsing the (many, small) group structures costs no system calls and contiguous
        # datasets come out as views of the page cache
        try:
            self.map = mmap.mmap(self.fh.fileno(), 0, access=mmap.ACCESS_READ)
        except (ValueError, OSError):
            self.map = None
        self.size_of_offsets = self.size_of_lengths = 8
        self.base = 0
        self.root_header = None
        self.root_symtab = None
        self._read_superblock()

    def close(self):
        if self.map is not None:
            try:
                self.map.close()
            except BufferError:                        # arrays handed out still view the mapping: the OS unmaps it with them
                pass
            self.map = None
        if self.fh is not None:
            self.fh.close()
            self.fh = None

    def read(self, offset, n):
        start = self.base + offset
        if self.map is not None:
            if start + n > len(self.map):
                raise Hdf5FormatError("%s: short read at %d (+%d)" % (self.path, offset, n))
            return self.map[start:start + n]
        self.fh.seek(start)
        data = self.fh.read(n)
        if len(data) != n:
            raise Hdf5FormatError("%s: short read at %d (+%d)" % (self.path, offset, n))
        return data

    def view(self, offset, n):
        """Zero-copy bytes of a contiguous dataset when the file is mapped."""
        start = self.base + offset
        if self.map is not None and start + n <= len(self.map):
            return memoryview(self.map)[start:start + n]
        return self.read(offset, n)

    def _uint(self, data, pos, size):
        return int.from_bytes(data[pos:pos + size], "little")

    def _read_superblock(self):
        pos = 0
        while True:                                   # the superblock may sit at 0, 512, 1024, ...
            self.fh.seek(pos)
            if self.fh.read(8) == SIGNATURE:
                break
            pos = 512 if pos == 0 else pos * 2
            if pos > (1 << 26):
                raise Hdf5FormatError("%s: not an HDF5 file (no signature)" % self.path)
        self.fh.seek(pos)
        head = self.fh.read(128).ljust(128, b"\0")
        version = head[8]
        if version in (0, 1):
            self.size_of_offsets, self.size_of_lengths = head[13], head[14]
            p = 24 + (4 if version == 1 else 0)
            o = self.size_of_offsets
            self.base = self._uint(head, p, o)
            p += 4 * o                                # base, free-space info, end of file, driver info
            # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
            self.root_header = self._uint(head, p + o, o)
            cache_type = self._uint(head, p + 2 * o, 4)
            if cache_type == 1:
                self.root_symtab = (self._uint(head, p + 2 * o + 8, o), self._uint(head, p + 3 * o + 8, o))
        elif version in (2, 3):
            self.size_of_offsets, self.size_of_lengths = head[9], head[10]
            o = self.size_of_offsets
            self.base = self._uint(head, 12, o)
            self.root_header = self._uint(head, 12 + 3 * o, o)
        else:
            raise Hdf5FormatError("%s: superblock version %d is not supported" % (self.path, version))
        if self.size_of_offsets != 8 or self.size_of_lengths != 8:
            raise Hdf5FormatError("%s: only 8-byte offsets / lengths are supported" % self.path)

    # ---- object headers -------------------------------------------------------------------
    def messages(self, address):
        """[(type, flags, data bytes)] of the object header at `address` (versions 1 and 2, continuations followed)."""
        first = self.read(address, 16)
        out = []
        if first[:4] == b"OHDR":
            flags = first[5]
            p = 6
            if flags & 0x20:
                p += 16                               # four timestamps
            if flags & 0x10:
                p += 4                                # attribute phase-change values
            size_bytes = 1 << (flags & 3)
            head = self.read(address, p + size_bytes)
            chunk_size = self._uint(head, p, size_bytes)
            blocks = [(address + p + size_bytes, chunk_size)]
            track_order = bool(flags & 0x04)
            while blocks:
                start, length = blocks.pop(0)
                data = self.read(start, length)
                q = 0
                while q + 4 <= length - 0:            # (a chunk ends with a 4-byte checksum inside `length` for OCHK blocks only)
                    mtype, msize, mflags = data[q], self._uint(data, q + 1, 2), data[q + 3]
                    q += 4 + (2 if track_order else 0)
                    if q + msize > length:
                        break
                    body = data[q:q + msize]
                    q += msize
                    if mtype == 0x10:
                        cont, clen = self._uint(body, 0, 8), self._uint(body, 8, 8)
                        blocks.append((cont + 4, clen - 8))       # skip "OCHK", drop the checksum
                    elif mtype != 0:
                        out.append((mtype, mflags, body))
            return out
        if first[0] != 1:
            raise Hdf5FormatError("%s: object header version %d at %d" % (self.path, first[0], address))
        n_messages = self._uint(first, 2, 2)
        header_size = self._uint(first, 8, 4)
        blocks = [(address + 16, header_size)]
        while blocks and len(out) < n_messages + 64:
            start, length = blocks.pop(0)
            data = self.read(start, length)
            q = 0
            while q + 8 <= length:
                mtype, msize, mflags = _MSG_HEAD.unpack_from(data, q)
                body = data[q + 8:q + 8 + msize]
                q += 8 + msize
                if mtype == 0x10:
                    blocks.append((self._uint(body, 0, 8), self._uint(body, 8, 8)))
                elif mtype != 0:
                    out.append((mtype, mflags, body))
        return out

    # ---- groups ---------------------------------------------------------------------------
    def links(self, address, symtab=None):
        """name -> object header address of the group whose header is at `address`."""
        out = {}
        msgs = self.messages(address)
        for mtype, _, body in msgs:
            if mtype == 0x11:
                symtab = (self._uint(body, 0, 8), self._uint(body, 8, 8))
            elif mtype == 0x06:
                name, target = self._link_message(body)
                if target is not None:
                    out[name] = target
            elif mtype == 0x02:
                # link info: version, flags, [max creation index], fractal heap address, name index B-tree address
                p = 2 + (8 if body[1] & 1 else 0)
                if self._uint(body, p, 8) != UNDEF:
                    raise Hdf5FormatError("%s: group with dense link storage (fractal heap) is not supported by minih5; "
                                          "install h5py to read this file" % self.path)
        if symtab is not None and not out:
            btree, heap = symtab
            heap_data = self._local_heap(heap)
            self._walk_group_btree(btree, heap_data, out)
        return out

    def _link_message(self, body):
        flags = body[1]
        p = 2
        link_type = 0
        if flags & 0x08:
            link_type = body[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        nbytes = 1 << (flags & 3)
        name_len = self._uint(body, p, nbytes)
        p += nbytes
        name = body[p:p + name_len].decode("utf-8")
        p += name_len
        if link_type != 0:
            return name, None                          # soft / external links are not followed
        return name, self._uint(body, p, 8)

    def _local_heap(self, address):
        head = self.read(address, 32)
        if head[:4] != b"HEAP":
            raise Hdf5FormatError("%s: local heap signature missing at %d" % (self.path, address))
        size, data_address = self._uint(head, 8, 8), self._uint(head, 24, 8)
        return self.read(data_address, size)

    def _walk_group_btree(self, address, heap, out):
        head = self.read(address, 24)
        if head[:4] != b"TREE" or head[4] != 0:
            raise Hdf5FormatError("%s: group B-tree node expected at %d" % (self.path, address))
        level, used = head[5], self._uint(head, 6, 2)
        body = self.read(address + 24, (2 * used + 1) * 8)
        for i in range(used):
            child = self._uint(body, (2 * i + 1) * 8, 8)
            if level > 0:
                self._walk_group_btree(child, heap, out)
            else:
                node = self.read(child, 8)
                if node[:4] != b"SNOD":
                    raise Hdf5FormatError("%s: symbol table node expected at %d" % (self.path, child))
                count = self._uint(node, 6, 2)
                entries = self.read(child + 8, count * 40)
                for k in range(count):
                    name_offset = self._uint(entries, k * 40, 8)
                    end = heap.index(b"\0", name_offset)
                    out[heap[name_offset:end].decode("utf-8")] = self._uint(entries, k * 40 + 8, 8)

    # ---- datasets -------------------------------------------------------------------------
    def dataset_info(self, address):
        info = {"shape": None, "dtype": None, "layout": None, "filters": [], "vlen_string": False}
        for mtype, _, body in self.messages(address):
            if mtype == 0x01:
                version, rank, flags = body[0], body[1], body[2]
                p = 8 if version == 1 else 4
                if version == 2 and body[3] == 2:
                    info["shape"] = (0,)
                else:
                    info["shape"] = tuple(self._uint(body, p + 8 * i, 8) for i in range(rank))
            elif mtype == 0x03:
                info["dtype"], info["vlen_string"] = self._datatype(body)
            elif mtype == 0x08:
                info["layout"] = self._layout(body)
            elif mtype == 0x0B:
                info["filters"] = self._filters(body)
        if info["shape"] is None or info["dtype"] is None or info["layout"] is None:
            return None                                  # not a dataset
        return info

    _dtype_cache = {}

    def _datatype(self, body):
        key = bytes(body[:8])
        hit = self._dtype_cache.get(key)
        if hit is not None and (key[0] & 0x0F) in (0, 1, 3):
            return hit
        out = self._datatype_uncached(body)
        if (key[0] & 0x0F) in (0, 1, 3):
            self._dtype_cache[key] = out
        return out

    def _datatype_uncached(self, body):
        cls, version = body[0] & 0x0F, body[0] >> 4
        bits0, bits1 = body[1], body[2]
        size = self._uint(body, 4, 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype("%s%s%d" % (order, "i" if bits0 & 0x08 else "u", size)), False
        if cls == 1:
            return np.dtype("%sf%d" % (order, size)), False
        if cls == 3:
            return np.dtype("S%d" % size), False
        if cls == 9:
            if (bits0 & 0x0F) == 1:                      # variable-length string
                return np.dtype("V16"), True
            raise Hdf5FormatError("%s: variable-length sequences are not supported" % self.path)
        if cls == 4:                                     # bit field: raw unsigned
            return np.dtype("%su%d" % (order, size)), False
        if cls == 8:                                     # enumeration: read as its base type
            base, _ = self._datatype(body[8:])
            return base, False
        raise Hdf5FormatError("%s: datatype class %d (version %d) is not supported" % (self.path, cls, version))

    def _layout(self, body):
        version = body[0]
        if version == 3:
            cls = body[1]
            if cls == 0:
                size = self._uint(body, 2, 2)
                return ("compact", body[4:4 + size])
            if cls == 1:
                return ("contiguous", self._uint(body, 2, 8), self._uint(body, 10, 8))
            if cls == 2:
                ndim = body[2]
                btree = self._uint(body, 3, 8)
                dims = tuple(self._uint(body, 11 + 4 * i, 4) for i in range(ndim))
                return ("chunked", btree, dims[:-1], dims[-1])
        elif version in (1, 2):
            ndim, cls = body[1], body[2]
            p = 8
            address = None
            if cls != 0:
                address = self._uint(body, p, 8)
                p += 8
            dims = tuple(self._uint(body, p + 4 * i, 4) for i in range(ndim))
            p += 4 * ndim
            if cls == 1:
                return ("contiguous", address, None)
            if cls == 2:
                elem = self._uint(body, p, 4)
                return ("chunked", address, dims, elem)
            size = self._uint(body, p, 4)
            return ("compact", body[p + 4:p + 4 + size])
        raise Hdf5FormatError("%s: data layout version %d is not supported" % (self.path, version))

    def _filters(self, body):
        version, count = body[0], body[1]
        p = 8 if version == 1 else 2
        out = []
        for _ in range(count):
            fid = self._uint(body, p, 2)
            p += 2
            name_len = 0
            if version == 1 or fid >= 256:
                name_len = self._uint(body, p, 2)
                p += 2
            p += 2                                       # flags
            n_client = self._uint(body, p, 2)
            p += 2
            if name_len:
                p += (name_len + 7) // 8 * 8 if version == 1 else name_len
            client = [self._uint(body, p + 4 * i, 4) for i in range(n_client)]
            p += 4 * n_client
            if version == 1 and n_client % 2:
                p += 4
            out.append((fid, client))
        return out

    def read_dataset(self, info):
        shape, dtype = info["shape"], info["dtype"]
        count = math.prod(shape) if shape else 1
        layout = info["layout"]
        if layout[0] == "compact":
            raw = bytes(layout[1])
        elif layout[0] == "contiguous":
            raw = b"" if layout[1] == UNDEF or count == 0 else self.view(layout[1], count * dtype.itemsize)
        else:
            raw = self._read_chunked(layout, info, shape, dtype)
        array = np.frombuffer(raw, dtype=dtype, count=count).reshape(shape)
        if info["vlen_string"]:
            flat = [self._vlen_string(bytes(v)) for v in array.reshape(-1)]
            return np.array(flat, dtype=object).reshape(shape)
        # (a contiguous dataset of a mapped file is returned as a read-only view of the page cache: no copy)
        if array.ndim == 0:
            return array.copy()[()]
        return array if layout[0] == "contiguous" and self.map is not None else array.copy()

    def _read_chunked(self, layout, info, shape, dtype):
        _, btree, chunk_dims, _elem = layout
        rank = len(shape)
        out = np.zeros(shape, dtype=dtype)
        if btree == UNDEF:
            return out.tobytes()
        for offsets, address, nbytes, mask in self._chunks(btree, rank):
            raw = self.read(address, nbytes)
            for index in range(len(info["filters"]) - 1, -1, -1):
                fid, client = info["filters"][index]
                if mask & (1 << index):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    width = client[0] if client else dtype.itemsize
                    n = len(raw) // width
                    raw = np.frombuffer(raw[:n * width], np.uint8).reshape(width, n).T.tobytes() + raw[n * width:]
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise Hdf5FormatError("%s: filter %d is not supported" % (self.path, fid))
            chunk = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunk_dims))).reshape(chunk_dims)
            sel_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offsets, chunk_dims, shape))
            sel_in = tuple(slice(0, s.stop - s.start) for s in sel_out)
            out[sel_out] = chunk[sel_in]
        return out.tobytes()

    def _chunks(self, address, rank):
        head = self.read(address, 24)
        if head[:4] != b"TREE" or head[4] != 1:
            raise Hdf5FormatError("%s: chunk B-tree node expected at %d" % (self.path, address))
        level, used = head[5], self._uint(head, 6, 2)
        key_size = 8 + 8 * (rank + 1)
        body = self.read(address + 24, used * (key_size + 8) + key_size)
        for i in range(used):
            k = i * (key_size + 8)
            nbytes, mask = self._uint(body, k, 4), self._uint(body, k + 4, 4)
            offsets = tuple(self._uint(body, k + 8 + 8 * d, 8) for d in range(rank))
            child = self._uint(body, k + key_size, 8)
            if level > 0:
                for item in self._chunks(child, rank):
                    yield item
            else:
                yield offsets, child, nbytes, mask

    def _vlen_string(self, ref):
        length, address, index = self._uint(ref, 0, 4), self._uint(ref, 4, 8), self._uint(ref, 12, 4)
        if address == 0 or address == UNDEF:
            return ""
        head = self.read(address, 16)
        if head[:4] != b"GCOL":
            raise Hdf5FormatError("%s: global heap collection expected at %d" % (self.path, address))
        data = self.read(address, self._uint(head, 8, 8))
        p = 16
        while p + 16 <= len(data):
            obj_index, obj_size = self._uint(data, p, 2), self._uint(data, p + 8, 8)
            if obj_index == index:
                return data[p + 16:p + 16 + length].decode("utf-8", "replace")
            if obj_index == 0:
                break
            p += 16 + (obj_size + 7) // 8 * 8
        raise Hdf5FormatError("%s: global heap object %d not found" % (self.path, index))


class Dataset(object):
    def __init__(self, reader, info, name):
        self._reader, self._info, self.name = reader, info, name
        self.shape = info["shape"]
        self.dtype = np.dtype(object) if info["vlen_string"] else info["dtype"]
        self._value = None

    def __getitem__(self, key):
        if self._value is None:
            self._value = self._reader.read_dataset(self._info)
        return self._value if key == () or key is Ellipsis else self._value[key]

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        value = np.asarray(self[()])
        return value.astype(dtype) if dtype is not None else value


class Group(object):
    def __init__(self, reader, address, name, symtab=None):
        self._reader, self._address, self.name = reader, address, name
        self._links = reader.links(address, symtab)
        self._cache = {}

    def keys(self):
        return list(self._links.keys())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._links)

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, key):
        node = self
        for part in [p for p in str(key).split("/") if p]:
            if not isinstance(node, Group):
                raise KeyError(key)
            node = node._child(part)
        return node

    def _child(self, part):
        if part not in self._cache:
            if part not in self._links:
                raise KeyError(part)
            address = self._links[part]
            info = self._reader.dataset_info(address)
            path = self.name.rstrip("/") + "/" + part
            self._cache[part] = Dataset(self._reader, info, path) if info is not None else Group(self._reader, address, path)
        return self._cache[part]


# =============================================================================================
# writing
# =============================================================================================
def _pad8(data):
    return data + b"\0" * (-len(data) % 8)


def _message(mtype, body):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def _datatype_message(dtype):
    dtype = np.dtype(dtype)
    if dtype.kind in "iu":
        bits = 0x08 if dtype.kind == "i" else 0
        return struct.pack("<BBBBI", 0x10, bits, 0, 0, dtype.itemsize) + struct.pack("<HH", 0, 8 * dtype.itemsize)
    if dtype.kind == "f" and dtype.itemsize in (4, 8):
        if dtype.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            sign = 31
        else:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            sign = 63
        return struct.pack("<BBBBI", 0x11, 0x20, sign, 0, dtype.itemsize) + props
    if dtype.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, max(dtype.itemsize, 1))     # null-padded, ASCII
    raise TypeError("minih5 cannot store dtype %s" % dtype)


def _storable(value):
    """The array as both writers store it: utf-8 bytes for text, uint8 for booleans, little-endian."""
    array = np.asarray(value)
    if array.dtype.kind == "U":
        array = np.char.encode(array, "utf-8")
    if array.dtype.kind == "b":
        array = array.astype(np.uint8)
    if array.dtype.kind == "O":
        raise TypeError("minih5 cannot store object arrays")
    if array.dtype.byteorder == ">":
        array = array.astype(array.dtype.newbyteorder("<"))
    if array.dtype.kind not in "iufS" or (array.dtype.kind == "f" and array.dtype.itemsize not in (4, 8)):
        raise TypeError("minih5 cannot store dtype %s" % array.dtype)     # (at assignment, not when close() writes the header)
    return array


class _NativeWriter(object):
    """The same writer in C++ (include/helen_h5write.h, helen_b200/csrc_host/h5write_host.cpp): byte-identical files, but
    the per-dataset bookkeeping and the group structure written at close() - two thirds of the time of a prediction file
    in the reference's schema - cost no Python."""

    def __init__(self, path):
        import ctypes
        from . import _h5write_native as native
        self._ct, self._native, self._lib = ctypes, native, native.load()
        self.path = path
        self._err = ctypes.create_string_buffer(512)
        self._handle = ctypes.c_void_p()
        status = self._lib.hw_create(os.fsencode(path), ctypes.byref(self._handle), self._err, len(self._err))
        if status != native.HW_OK:
            self._handle = None
            native.raise_for(status, self._err.value.decode(errors="replace"))

    def _check(self, status):
        if status != self._native.HW_OK:
            self._native.raise_for(status, self._err.value.decode(errors="replace"))

    def _described(self, array, shape):
        if array.dtype.kind not in "iufS":
            raise TypeError("minih5 cannot store dtype %s" % array.dtype)
        dims = (self._ct.c_uint64 * max(len(shape), 1))(*shape)
        return array.dtype.kind.encode(), array.dtype.itemsize, len(shape), dims

    def set(self, path, value):
        if not [p for p in str(path).split("/") if p]:
            raise ValueError("empty dataset name")
        array = _storable(value)
        kind, itemsize, rank, dims = self._described(array, array.shape)
        data = np.ascontiguousarray(array)             # (a 0-d array becomes 1-d here: the shape above is the dataset's)
        self._check(self._lib.hw_dataset(self._handle, str(path).encode(), kind, itemsize, rank, dims, data.ctypes.data,
                                         self._err, len(self._err)))

    def set_rows(self, parents, name, value):
        array = np.ascontiguousarray(_storable(value))
        if array.ndim < 1 or array.shape[0] != len(parents):
            raise ValueError("set_rows: one row per parent")
        kind, itemsize, rank, dims = self._described(array, array.shape[1:])
        packed = b"\0".join(str(p).encode() for p in parents) + b"\0"
        self._check(self._lib.hw_rows(self._handle, packed, len(parents), str(name).encode(), kind, itemsize, rank, dims,
                                      array.ctypes.data, self._err, len(self._err)))

    def contains(self, key):
        return bool(self._lib.hw_contains(self._handle, str(key).encode()))

    def close(self):
        handle, self._handle = self._handle, None      # (forgotten first: see _feed_native.ImageFile.close)
        if handle is not None:
            self._check(self._lib.hw_close(handle, self._err, len(self._err)))


class _WNode(object):
    def __init__(self):
        self.children = {}                            # groups
        self.shape = self.dtype = self.data_address = None
        self.nbytes = 0
        self.is_dataset = False


class _Writer(object):
    def __init__(self, path):
        self.path = path
        self.fh = open(path, "wb")
        self.fh.write(b"\0" * 96)                     # the superblock is written last
        self.pos = 96
        self.root = _WNode()

    def _append(self, data):
        pad = -self.pos % 8
        if pad:
            self.fh.write(b"\0" * pad)
            self.pos += pad
        address = self.pos
        self.fh.write(data)
        self.pos += len(data)
        return address

    def set(self, path, value):
        parts = [p for p in str(path).split("/") if p]
        if not parts:
            raise ValueError("empty dataset name")
        node = self.root
        for part in parts[:-1]:
            nxt = node.children.get(part)
            if nxt is None:
                nxt = node.children[part] = _WNode()
            if nxt.is_dataset:
                raise ValueError("%s: %s is a dataset" % (path, part))
            node = nxt
        if parts[-1] in node.children:
            raise ValueError("Unable to create dataset (name already exists): " + str(path))
        array = _storable(value)
        leaf = _WNode()
        leaf.is_dataset, leaf.shape, leaf.dtype = True, array.shape, array.dtype
        # the array's own buffer goes to the file: no tobytes() copy (a fresh multi-megabyte bytes object per dataset is
        # mostly page faults - it was a third of the packed prediction writer's time)
        flat = np.ascontiguousarray(array).reshape(-1)
        leaf.nbytes = flat.nbytes
        leaf.data_address = self._append(flat.view(np.uint8).data if flat.nbytes else b"") if flat.nbytes else UNDEF
        node.children[parts[-1]] = leaf

    def set_rows(self, parents, name, value):
        """len(parents) datasets ``parents[i] + '/' + name`` = row i of `value`; the array goes to the file in one piece."""
        array = np.ascontiguousarray(_storable(value))
        if array.ndim < 1 or array.shape[0] != len(parents):
            raise ValueError("set_rows: one row per parent")
        nodes = []
        for parent in parents:
            node = self.root
            for part in [p for p in str(parent).split("/") if p]:
                nxt = node.children.get(part)
                if nxt is None:
                    nxt = node.children[part] = _WNode()
                if nxt.is_dataset:
                    raise ValueError("%s/%s: %s is a dataset" % (parent, name, part))
                node = nxt
            if name in node.children:
                raise ValueError("Unable to create dataset (name already exists): %s/%s" % (parent, name))
            nodes.append(node)
        row_bytes = array.nbytes // max(len(parents), 1)
        base = self._append(array.reshape(-1).view(np.uint8).data) if array.nbytes else UNDEF
        for i, node in enumerate(nodes):
            if name in node.children:
                raise ValueError("Unable to create dataset (name already exists): %s (twice in one call)" % name)
            leaf = _WNode()
            leaf.is_dataset, leaf.shape, leaf.dtype, leaf.nbytes = True, array.shape[1:], array.dtype, row_bytes
            leaf.data_address = base + i * row_bytes if row_bytes else UNDEF
            node.children[name] = leaf

    def contains(self, key):
        node = self.root
        for part in [p for p in str(key).split("/") if p]:
            if node.is_dataset or part not in node.children:
                return False
            node = node.children[part]
        return True

    # ---- structure, written by close() ----------------------------------------------------
    def _write_dataset(self, node):
        rank = len(node.shape)
        dataspace = struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", d) for d in node.shape)
        fill = struct.pack("<BBBB", 2, 2, 2, 0)       # version 2, late allocation, write at allocation time, undefined
        layout = struct.pack("<BBQQ", 3, 1, node.data_address, node.nbytes)
        return self._append(_object_header([_message(0x01, dataspace), _message(0x03, _datatype_message(node.dtype)),
                                            _message(0x05, fill), _message(0x08, layout)]))

    def _write_group(self, node):
        """Object header address of the group, after writing its members, heap, symbol nodes and B-tree."""
        entries = []
        for name in sorted(node.children, key=lambda s: s.encode("utf-8")):
            child = node.children[name]
            if child.is_dataset:
                entries.append((name, self._write_dataset(child), None))
            else:
                address, btree, heap = self._write_group(child)
                entries.append((name, address, (btree, heap)))
        # local heap: the empty string at offset 0, then the member names, each padded to 8 bytes
        heap_data = bytearray(8)
        offsets = []
        for name, _, _ in entries:
            offsets.append(len(heap_data))
            heap_data += _pad8(name.encode("utf-8") + b"\0")
        free_offset = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)         # one free block (next = 1: last, size 16): libhdf5 wants room to grow
        data_address = self._append(bytes(heap_data))
        heap_address = self._append(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_offset, data_address))
        # symbol table nodes of up to 2 * LEAF_K entries, in name order
        leaves = []                                    # (address, heap offset of the largest name)
        per_leaf = 2 * LEAF_K
        for start in range(0, max(len(entries), 1), per_leaf):
            chunk = entries[start:start + per_leaf]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for k, (name, address, sub) in enumerate(chunk):
                if sub is None:
                    body += struct.pack("<QQII16x", offsets[start + k], address, 0, 0)
                else:
                    body += struct.pack("<QQIIQQ", offsets[start + k], address, 1, 0, sub[0], sub[1])
            body += b"\0" * (40 * (per_leaf - len(chunk)))
            leaves.append((self._append(body), offsets[start + len(chunk) - 1] if chunk else 0))
        # B-tree over the leaves, 2 * INTERNAL_K children per node
        level, nodes = 0, leaves
        while True:
            parents = []
            fan = 2 * INTERNAL_K
            groups = [nodes[i:i + fan] for i in range(0, len(nodes), fan)]
            addresses = []
            for chunk in groups:
                body = b"TREE" + struct.pack("<BBH", 0, level, len(chunk)) + struct.pack("<QQ", UNDEF, UNDEF)
                body += struct.pack("<Q", 0)           # key 0: the empty string sorts before every name
                for address, last_key in chunk:
                    body += struct.pack("<QQ", address, last_key)
                body += b"\0" * (16 * (fan - len(chunk)))
                addresses.append(self._append(body))
                parents.append((addresses[-1], chunk[-1][1]))
            # sibling pointers
            for i, address in enumerate(addresses):
                left = addresses[i - 1] if i > 0 else UNDEF
                right = addresses[i + 1] if i + 1 < len(addresses) else UNDEF
                self.fh.seek(address + 8)
                self.fh.write(struct.pack("<QQ", left, right))
            self.fh.seek(self.pos)
            if len(parents) == 1:
                btree_address = parents[0][0]
                break
            level, nodes = level + 1, parents
        header = self._append(_object_header([_message(0x11, struct.pack("<QQ", btree_address, heap_address))]))
        return header, btree_address, heap_address

    def close(self):
        if self.fh is None:
            return
        header, btree, heap = self._write_group(self.root)
        pad = -self.pos % 8
        if pad:
            self.fh.write(b"\0" * pad)
            self.pos += pad
        superblock = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        superblock += struct.pack("<QQQQ", 0, UNDEF, self.pos, UNDEF)
        superblock += struct.pack("<QQIIQQ", 0, header, 1, 0, btree, heap)
        assert len(superblock) == 96
        self.fh.seek(0)
        self.fh.write(superblock)
        self.fh.close()
        self.fh = None


# =============================================================================================
# the h5py-like front
# =============================================================================================
class File(object):
    """``File(path, 'r')`` reads, ``File(path, 'w')`` creates (the structure reaches the disk on close())."""

    def __init__(self, path, mode="r"):
        self.filename, self.mode = path, mode
        self._writer = self._reader = self._root = None
        if mode == "r":
            self._reader = _Reader(path)
            self._root = Group(self._reader, self._reader.root_header, "/", self._reader.root_symtab)
        elif mode in ("w", "x", "w-"):
            self._writer = _NativeWriter(path) if _native_writer_wanted() else _Writer(path)
        else:
            raise ValueError("minih5 supports modes 'r' and 'w' (got %r); appending to an existing file needs h5py" % (mode,))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        if self._reader is not None:
            self._reader.close()
            self._reader = None
        if self._writer is not None:
            self._writer.close()
            self._writer = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def keys(self):
        if self._root is None:
            if isinstance(self._writer, _NativeWriter):
                raise IOError("minih5: a file opened for writing cannot be listed before it is closed")
            return list(self._writer.root.children.keys())
        return self._root.keys()

    def __contains__(self, key):
        if self._root is not None:
            return key in self._root
        return self._writer.contains(key)

    def set_rows(self, parents, name, value):
        """len(parents) datasets ``parents[i] + '/' + name``, dataset i = value[i]: what a loop of assignments writes, with
        one file write and no per-dataset Python (DataStore.write_predictions)."""
        if self._writer is None:
            raise IOError("minih5: file is not open for writing")
        self._writer.set_rows(parents, name, value)

    def __getitem__(self, key):
        if self._root is None:
            raise IOError("minih5: a file opened for writing cannot be read back before it is closed")
        return self._root[key]

    def __setitem__(self, key, value):
        if self._writer is None:
            raise IOError("minih5: file is not open for writing")
        self._writer.set(key, value)
