"""Minimal HDF5 reader (and fixture writer) for the reference's sample files -- no h5py / libhdf5 needed.

The reference keeps collected samples in ``Levels/<l>/collected_values``: a 1-D, chunked, resizable dataset whose
element type is the array ``(2, M)`` of ``float64`` (``mlmc/tool/hdf5.py:311-320``), written by h5py with its default
("earliest") format bounds, read back through ``dataset[chunk_slice]`` (``mlmc/tool/hdf5.py:365-376``,
``mlmc/sample_storage_hdf.py:169-184``).  h5py / libhdf5 are not part of this image, so this module reads that layout
directly from the file, following the HDF5 File Format Specification (version 0 / 2 superblock, version 1 object
headers with continuation blocks, old-style groups = symbol-table B-tree + local heap, dataspace / datatype / layout /
filter / attribute messages, contiguous and chunked (version 1 B-tree) raw data, optional deflate + shuffle filters).
What the reference's files do not use (version 2 object headers, dense link / attribute storage, extensible-array chunk
indices of ``libver='latest'`` files, variable-length data) raises ``Unsupported`` with the feature named; unsupported
ATTRIBUTE types are skipped (``attrs[name] is None``).

Pinning: there is no h5py-written MLMC file in this environment.  The parser core is checked against a file written by
the real HDF5 library (a MATLAB v7.3 file shipped with scipy's test data, tests/test_hdf5_min.py); the chunked path
round-trips the fixture writer below (same on-disk structures: several chunks, two-level chunk B-tree, partial last
chunk).  ``SampleStorageHDF`` prefers h5py when it is importable.

``write_mlmc_file`` writes the reference's file structure for tests and benchmarks (old-style groups, chunked
``collected_values``, ``level_parameters`` / ``n_ops_estimate`` attributes).
"""
import bisect
import mmap
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


#: optional ``f(dst_address, src_address, n_rows, keep_bytes, src_pitch)``: a multi-threaded host copy for large runs of
#: rows (set by ``mlmc_b200.sample_storage`` to the native helper; this module itself has no dependency but NumPy)
parallel_copy = None


class Unsupported(NotImplementedError):
    pass


# ------------------------------------------------------------------------------------------------ datatype messages
def parse_datatype(buf, pos):
    """-> (numpy dtype or None if not representable, position after the datatype description)."""
    cls_ver, b0, b1, _b2, size = struct.unpack_from("<BBBBI", buf, pos)
    cls, ver = cls_ver & 0x0F, cls_ver >> 4
    pos += 8
    if cls == 0:                                     # fixed point
        order = ">" if b0 & 1 else "<"
        signed = bool(b0 & 8)
        return np.dtype("%s%s%d" % (order, "i" if signed else "u", size)), pos + 4
    if cls == 1:                                     # floating point
        order = ">" if b0 & 1 else "<"
        return np.dtype("%sf%d" % (order, size)), pos + 12
    if cls == 3:                                     # fixed-length string
        return np.dtype("S%d" % size), pos
    if cls == 6:                                     # compound
        n_members = b0 | (b1 << 8)
        names, formats, offsets = [], [], []
        ok = True
        for _ in range(n_members):
            end = buf.find(b"\x00", pos)
            name = bytes(buf[pos:end]).decode("utf-8", "replace")
            if ver < 3:
                pos += ((end - pos) // 8 + 1) * 8       # null-terminated, padded to a multiple of 8
            else:
                pos = end + 1
            if ver == 3:
                n_off = 1
                while (1 << (8 * n_off)) <= size and n_off < 4:
                    n_off += 1
                offset = int.from_bytes(bytes(buf[pos:pos + n_off]), "little")
                pos += n_off
            else:
                offset, = struct.unpack_from("<I", buf, pos)
                pos += 4
            dims = None
            if ver == 1:
                rank = buf[pos]
                pos += 12                               # dimensionality, reserved, permutation, reserved
                dims = struct.unpack_from("<4I", buf, pos)[:rank]
                pos += 16
            member, pos = parse_datatype(buf, pos)
            if member is None:
                ok = False
            elif dims:
                member = np.dtype((member, tuple(int(d) for d in dims)))
            names.append(name)
            formats.append(member)
            offsets.append(offset)
        if not ok:
            return None, pos
        return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), pos
    if cls == 9:                                     # variable length: base type follows; data lives in the global heap
        _base, pos = parse_datatype(buf, pos)
        return None, pos
    if cls == 10:                                    # array
        rank = buf[pos]
        pos += 4 if ver < 3 else 1
        dims = struct.unpack_from("<%dI" % rank, buf, pos)
        pos += 4 * rank
        if ver < 3:
            pos += 4 * rank                             # permutation indices
        base, pos = parse_datatype(buf, pos)
        if base is None:
            return None, pos
        return np.dtype((base, tuple(int(d) for d in dims))), pos
    if cls == 7:                                     # reference
        return None, pos
    raise Unsupported("HDF5 datatype class %d" % cls)


def parse_dataspace(buf, pos):
    """-> (shape, maxshape)"""
    ver, rank, flags = struct.unpack_from("<BBB", buf, pos)
    if ver == 1:
        pos += 8
    elif ver == 2:
        if buf[pos + 3] == 2:                           # null dataspace
            return None, None
        pos += 4
    else:
        raise Unsupported("dataspace message version %d" % ver)
    shape = struct.unpack_from("<%dQ" % rank, buf, pos)
    pos += 8 * rank
    maxshape = shape
    if flags & 1:
        maxshape = tuple(None if m == UNDEF else m for m in struct.unpack_from("<%dQ" % rank, buf, pos))
    return tuple(int(s) for s in shape), maxshape


# ------------------------------------------------------------------------------------------------ objects
class _Object:
    def __init__(self, f, address):
        self._f = f
        self._messages = f._object_messages(address)
        self._attrs = None

    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = {}
            buf = self._f._buf
            for mtype, pos, size in self._messages:
                if mtype == 0x000C:
                    name, value = self._f._parse_attribute(buf, pos, size)
                    self._attrs[name] = value
                elif mtype == 0x0015:
                    info_flags = buf[pos + 1]
                    p = pos + 2 + (2 if info_flags & 1 else 0)
                    if struct.unpack_from("<Q", buf, p)[0] != UNDEF:
                        raise Unsupported("dense attribute storage (fractal heap)")
        return self._attrs


class Group(_Object):
    def __init__(self, f, address):
        super().__init__(f, address)
        self._links = None

    def _load(self):
        if self._links is not None:
            return
        links = {}
        buf = self._f._buf
        for mtype, pos, size in self._messages:
            if mtype == 0x0011:                           # symbol table: B-tree + local heap
                btree, heap = struct.unpack_from("<QQ", buf, pos)
                heap_data = self._f._local_heap(heap)
                self._f._walk_group_btree(btree, heap_data, links)
            elif mtype == 0x0006:                         # link message (new-style compact group)
                name, address = self._f._parse_link(buf, pos)
                if address is not None:
                    links[name] = address
            elif mtype == 0x0002:                         # link info: dense storage?
                flags = buf[pos + 1]
                p = pos + 2 + (8 if flags & 1 else 0)
                if struct.unpack_from("<Q", buf, p)[0] != UNDEF:
                    raise Unsupported("dense link storage (fractal heap)")
        self._links = links

    def keys(self):
        self._load()
        return list(self._links)

    def __contains__(self, name):
        self._load()
        return name in self._links

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            node._load()
            if part not in node._links:
                raise KeyError(path)
            node = node._f._open(node._links[part])
        return node


class Dataset(_Object):
    def __init__(self, f, address):
        super().__init__(f, address)
        buf = f._buf
        self.shape = self.maxshape = self.dtype = None
        self._layout = None
        self._filters = []
        for mtype, pos, size in self._messages:
            if mtype == 0x0001:
                self.shape, self.maxshape = parse_dataspace(buf, pos)
            elif mtype == 0x0003:
                self.dtype, _ = parse_datatype(buf, pos)
            elif mtype == 0x0008:
                self._layout = self._parse_layout(buf, pos)
            elif mtype == 0x000B:
                self._filters = self._parse_filters(buf, pos)
        if self._layout is None or self.shape is None:
            raise Unsupported("dataset without dataspace / layout message")
        if self.dtype is None:
            raise Unsupported("dataset element type (variable-length / reference data)")
        self._chunk_index = None

    def _parse_layout(self, buf, pos):
        ver, cls = buf[pos], buf[pos + 1]
        if ver in (1, 2):                                 # HDF5 <= 1.6 writers: dimensionality first, no explicit sizes
            ndim, cls = buf[pos + 1], buf[pos + 2]
            p = pos + 8
            address = None
            if cls != 0:
                address, = struct.unpack_from("<Q", buf, p)
                p += 8
            dims = struct.unpack_from("<%dI" % ndim, buf, p)
            p += 4 * ndim
            if cls == 1:
                return ("contiguous", address, None)
            if cls == 2:
                return ("chunked", address, tuple(int(d) for d in dims[:-1]), int(dims[-1]))
            size, = struct.unpack_from("<I", buf, p)
            return ("compact", p + 4, size)
        if ver != 3:
            raise Unsupported("data layout message version %d (files written with libver='latest' use version 4 chunk "
                              "indices)" % ver)
        if cls == 1:
            address, size = struct.unpack_from("<QQ", buf, pos + 2)
            return ("contiguous", address, size)
        if cls == 2:
            ndim = buf[pos + 2]
            btree, = struct.unpack_from("<Q", buf, pos + 3)
            dims = struct.unpack_from("<%dI" % ndim, buf, pos + 11)
            return ("chunked", btree, tuple(int(d) for d in dims[:-1]), int(dims[-1]))
        if cls == 0:
            size, = struct.unpack_from("<H", buf, pos + 2)
            return ("compact", pos + 4, size)
        raise Unsupported("layout class %d" % cls)

    @staticmethod
    def _parse_filters(buf, pos):
        ver, n = buf[pos], buf[pos + 1]
        pos += 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid, = struct.unpack_from("<H", buf, pos)
            pos += 2
            name_len = 0
            if ver == 1 or fid >= 256:
                name_len, = struct.unpack_from("<H", buf, pos)
                pos += 2
            _flags, n_vals = struct.unpack_from("<HH", buf, pos)
            pos += 4
            if name_len:
                pos += name_len if ver == 2 else ((name_len + 7) // 8) * 8
            vals = struct.unpack_from("<%dI" % n_vals, buf, pos)
            pos += 4 * n_vals
            if ver == 1 and n_vals % 2:
                pos += 4
            out.append((fid, vals))
        return out

    @property
    def chunks(self):
        return self._layout[2] if self._layout[0] == "chunked" else None

    def __len__(self):
        return self.shape[0]

    def _index(self):
        """Chunk index of a chunked dataset: sorted list of (first row, file address, stored bytes, filter mask)."""
        if self._chunk_index is None:
            entries = []
            if self._layout[1] != UNDEF:
                self._f._walk_chunk_btree(self._layout[1], len(self._layout[2]) + 1, entries)
            entries.sort()
            self._index_firsts = [e[0] for e in entries]
            self._chunk_index = entries
        return self._chunk_index

    def _decode_chunk(self, raw, mask):
        for k, (fid, vals) in reversed(list(enumerate(self._filters))):
            if mask & (1 << k):
                continue
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                width = vals[0] if vals else self.dtype.base.itemsize
                arr = np.frombuffer(raw, dtype=np.uint8)
                n = arr.size // width
                raw = arr[:n * width].reshape(width, n).T.tobytes() + arr[n * width:].tobytes()
            else:
                raise Unsupported("HDF5 filter %d" % fid)
        return raw

    def read_rows(self, lo=0, hi=None, out=None, keep_items=None):
        """Elements ``[lo, hi)`` along the first axis as an array ``[hi - lo, *shape[1:], *element subarray shape]``.
        ``out``: optional C-contiguous array of that shape and base dtype (e.g. a pinned staging buffer) to fill.
        ``keep_items`` (with ``out``; unfiltered chunked or contiguous data): only the first ``keep_items`` base items of
        every element are copied, ``out`` holds ``(hi - lo) * keep_items`` items -- level 0 of an MLMC file without its
        stored zero coarse row, straight from the file mapping."""
        n = self.shape[0]
        hi = n if hi is None else min(hi, n)
        lo = max(0, min(lo, hi))
        base, sub = (self.dtype.base, self.dtype.shape) if self.dtype.subdtype else (self.dtype, ())
        tail = tuple(self.shape[1:]) + tuple(sub)
        row_items = int(np.prod(tail, dtype=np.int64)) if tail else 1
        row_bytes = row_items * base.itemsize
        if out is None:
            if keep_items is not None:
                raise ValueError("keep_items needs an output array")
            out = np.empty((hi - lo,) + tail, dtype=base)
        flat = out.reshape(-1).view(np.uint8) if out.size else np.empty(0, dtype=np.uint8)
        buf = self._f._buf
        kind = self._layout[0]
        if hi == lo:
            return out
        keep_bytes = row_bytes if keep_items is None else int(keep_items) * base.itemsize
        if keep_bytes != row_bytes and (self._filters or not 0 < keep_bytes < row_bytes):
            raise Unsupported("keep_items on filtered chunks / out of range")
        if flat.size != (hi - lo) * keep_bytes:
            raise ValueError("output array of the wrong size")
        rows_out = flat.reshape(hi - lo, keep_bytes)              # one row of kept bytes per element

        def put(a, b, src, first):
            """rows [a, b) from ``src`` = the stored bytes of the elements first, first + 1, ..."""
            if parallel_copy is not None and (b - a) * keep_bytes >= (2 << 20):
                parallel_copy(rows_out.ctypes.data + (a - lo) * keep_bytes, src.ctypes.data + (a - first) * row_bytes,
                              b - a, keep_bytes, row_bytes)
            elif keep_bytes == row_bytes:
                flat[(a - lo) * row_bytes:(b - lo) * row_bytes] = src[(a - first) * row_bytes:(b - first) * row_bytes]
            else:
                rows_out[a - lo:b - lo] = src.reshape(-1, row_bytes)[a - first:b - first, :keep_bytes]

        if kind in ("contiguous", "compact"):
            start = self._layout[1] if kind == "compact" else self._f._base + self._layout[1]
            if kind == "contiguous" and self._layout[1] == UNDEF:
                flat[:] = 0
            else:
                put(lo, hi, np.frombuffer(buf, dtype=np.uint8, count=(hi - lo) * row_bytes, offset=start + lo * row_bytes), lo)
            return out
        chunk_dims = self._layout[2]
        if any(c != s for c, s in zip(chunk_dims[1:], self.shape[1:])):
            raise Unsupported("chunks that split trailing dimensions")
        c_rows = chunk_dims[0]
        index = self._index()
        filled = np.zeros(hi - lo, dtype=bool) if len(index) * c_rows < n else None
        base_off = self._f._base
        # first chunk that can overlap [lo, hi): the index is sorted by first row
        k = max(0, bisect.bisect_right(self._index_firsts, lo) - 1)
        while k < len(index) and index[k][0] < hi:
            first, address, stored, mask = index[k]
            if self._filters:
                a, b = max(first, lo), min(first + c_rows, hi, n)
                if a < b:
                    raw = self._decode_chunk(bytes(buf[base_off + address:base_off + address + stored]), mask)
                    put(a, b, np.frombuffer(raw, dtype=np.uint8), first)
                    if filled is not None:
                        filled[a - lo:b - lo] = True
                k += 1
                continue
            # unfiltered chunks written one after the other (the usual case) are ONE copy: a run of chunks that are
            # consecutive in rows and in the file
            run_rows = c_rows
            k2 = k + 1
            while (k2 < len(index) and index[k2][0] < hi and index[k2][0] == first + run_rows
                   and index[k2][1] == address + run_rows * row_bytes):
                run_rows += c_rows
                k2 += 1
            a, b = max(first, lo), min(first + run_rows, hi, n)
            if a < b:
                put(a, b, np.frombuffer(buf, dtype=np.uint8, count=run_rows * row_bytes, offset=base_off + address), first)
                if filled is not None:
                    filled[a - lo:b - lo] = True
            k = k2
        if filled is not None and not filled.all():                  # unallocated chunks read as the fill value (0)
            rows_out[~filled] = 0
        return out

    def __getitem__(self, key):
        if key == () or key is Ellipsis:
            return self.read_rows()
        if isinstance(key, slice):
            lo, hi, step = key.indices(self.shape[0])
            rows = self.read_rows(lo, hi)
            return rows if step == 1 else rows[::step]
        raise Unsupported("dataset indexing other than [()] / [slice]")

    def iter_chunks(self):
        """Row slices of the stored chunks (h5py's ``Dataset.iter_chunks`` for a 1-D dataset)."""
        if self._layout[0] != "chunked":
            yield (slice(0, self.shape[0], 1),)
            return
        c_rows = self._layout[2][0]
        for first in range(0, self.shape[0], c_rows):
            yield (slice(first, min(first + c_rows, self.shape[0]), 1),)


class File(Group):
    """Read-only HDF5 file (memory-mapped)."""

    def __init__(self, path):
        self._fh = open(path, "rb")
        try:
            self._buf = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:
            self._fh.close()
            raise
        self._cache = {}
        offset = 0
        while True:                                      # the superblock sits at 0, 512, 1024, 2048, ...
            if self._buf[offset:offset + 8] == SIGNATURE:
                break
            offset = 512 if offset == 0 else offset * 2
            if offset + 8 > len(self._buf):
                self.close()
                raise ValueError("%s is not an HDF5 file" % path)
        ver = self._buf[offset + 8]
        if ver in (0, 1):
            size_off, size_len = self._buf[offset + 13], self._buf[offset + 14]
            p = offset + 24 + (4 if ver == 1 else 0)
            self._base, = struct.unpack_from("<Q", self._buf, p)
            root_entry = p + 32
            root_address, = struct.unpack_from("<Q", self._buf, root_entry + 8)
        elif ver in (2, 3):
            size_off, size_len = self._buf[offset + 9], self._buf[offset + 10]
            self._base, _ext, _eof, root_address = struct.unpack_from("<QQQQ", self._buf, offset + 12)
        else:
            self.close()
            raise Unsupported("superblock version %d" % ver)
        if size_off != 8 or size_len != 8:
            self.close()
            raise Unsupported("offset / length sizes other than 8 bytes")
        super().__init__(self, root_address)

    def close(self):
        try:
            self._buf.close()
        except (BufferError, ValueError):
            pass                                          # NumPy views of the map are still alive: the GC closes it
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- low level ----
    def _object_messages(self, address):
        """[(type, position of the message body in the file buffer, size)] of an object header, continuation blocks
        followed."""
        buf = self._buf
        pos = self._base + address
        if buf[pos:pos + 4] == b"OHDR":
            raise Unsupported("version 2 object headers (file written with libver='latest')")
        ver, _res, n_msgs, _refs, hsize = struct.unpack_from("<BBHII", buf, pos)
        if ver != 1:
            raise Unsupported("object header version %d" % ver)
        blocks = [(pos + 16, hsize)]
        out = []
        while blocks and len(out) < n_msgs:
            p, remaining = blocks.pop(0)
            end = p + remaining
            while p + 8 <= end and len(out) < n_msgs:
                mtype, msize, _flags = struct.unpack_from("<HHB", buf, p)
                body = p + 8
                if mtype == 0x0010:
                    c_off, c_len = struct.unpack_from("<QQ", buf, body)
                    blocks.append((self._base + c_off, c_len))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    def _open(self, address):
        hit = self._cache.get(address)
        if hit is None:
            types = {m[0] for m in self._object_messages(address)}
            hit = Dataset(self, address) if 0x0008 in types else Group(self, address)
            self._cache[address] = hit
        return hit

    def _local_heap(self, address):
        pos = self._base + address
        if self._buf[pos:pos + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        size, _free, data = struct.unpack_from("<QQQ", self._buf, pos + 8)
        return (self._base + data, size)

    def _heap_string(self, heap, offset):
        start = heap[0] + offset
        end = self._buf.find(b"\x00", start)
        return bytes(self._buf[start:end]).decode("utf-8", "replace")

    def _walk_group_btree(self, address, heap, links):
        buf = self._buf
        pos = self._base + address
        if buf[pos:pos + 4] != b"TREE":
            raise ValueError("bad B-tree signature")
        node_type, level, used = struct.unpack_from("<BBH", buf, pos + 4)
        if node_type != 0:
            raise ValueError("group B-tree expected")
        p = pos + 24
        for i in range(used):
            child, = struct.unpack_from("<Q", buf, p + 8 + 16 * i)
            if level > 0:
                self._walk_group_btree(child, heap, links)
            else:
                self._read_symbol_node(child, heap, links)

    def _read_symbol_node(self, address, heap, links):
        buf = self._buf
        pos = self._base + address
        if buf[pos:pos + 4] != b"SNOD":
            raise ValueError("bad symbol table node signature")
        n, = struct.unpack_from("<H", buf, pos + 6)
        for i in range(n):
            name_off, header = struct.unpack_from("<QQ", buf, pos + 8 + 40 * i)
            links[self._heap_string(heap, name_off)] = header

    def _parse_link(self, buf, pos):
        ver, flags = buf[pos], buf[pos + 1]
        p = pos + 2
        link_type = 0
        if flags & 0x08:
            link_type = buf[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        len_size = 1 << (flags & 3)
        n = int.from_bytes(bytes(buf[p:p + len_size]), "little")
        p += len_size
        name = bytes(buf[p:p + n]).decode("utf-8", "replace")
        p += n
        if link_type != 0:
            return name, None                             # soft / external links are not followed
        return name, struct.unpack_from("<Q", buf, p)[0]

    def _walk_chunk_btree(self, address, n_key_dims, entries):
        buf = self._buf
        pos = self._base + address
        if buf[pos:pos + 4] != b"TREE":
            raise ValueError("bad B-tree signature")
        node_type, level, used = struct.unpack_from("<BBH", buf, pos + 4)
        if node_type != 1:
            raise ValueError("chunk B-tree expected")
        key_size = 8 + 8 * n_key_dims
        p = pos + 24
        for i in range(used):
            k = p + i * (key_size + 8)
            stored, mask = struct.unpack_from("<II", buf, k)
            first, = struct.unpack_from("<Q", buf, k + 8)
            child, = struct.unpack_from("<Q", buf, k + key_size)
            if level > 0:
                self._walk_chunk_btree(child, n_key_dims, entries)
            else:
                entries.append((int(first), int(child), int(stored), int(mask)))

    def _parse_attribute(self, buf, pos, size):
        ver = buf[pos]
        if ver == 1:
            name_size, dt_size, ds_size = struct.unpack_from("<HHH", buf, pos + 2)
            p = pos + 8
            pad = lambda v: (v + 7) // 8 * 8                                   # noqa: E731
        elif ver in (2, 3):
            name_size, dt_size, ds_size = struct.unpack_from("<HHH", buf, pos + 2)
            p = pos + 8 + (1 if ver == 3 else 0)
            pad = lambda v: v                                                  # noqa: E731
        else:
            raise Unsupported("attribute message version %d" % ver)
        name = bytes(buf[p:p + name_size]).split(b"\x00")[0].decode("utf-8", "replace")
        p += pad(name_size)
        try:
            dtype, _ = parse_datatype(buf, p)
        except Unsupported:
            dtype = None
        p += pad(dt_size)
        shape, _ = parse_dataspace(buf, p)
        p += pad(ds_size)
        if dtype is None or shape is None:
            return name, None
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        base, sub = (dtype.base, dtype.shape) if dtype.subdtype else (dtype, ())
        n_items = count * (int(np.prod(sub, dtype=np.int64)) if sub else 1)
        value = np.frombuffer(buf, dtype=base, count=n_items, offset=p).reshape(tuple(shape) + tuple(sub)).copy()
        return name, value


# ------------------------------------------------------------------------------------------------ fixture writer
class _Writer:
    """Sequential writer of the structures above (version 0 superblock, version 1 object headers, old-style groups)."""

    def __init__(self):
        self.buf = bytearray()

    def alloc(self, n, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        pos = len(self.buf)
        self.buf.extend(b"\x00" * n)
        return pos

    def put(self, pos, data):
        self.buf[pos:pos + len(data)] = data

    @staticmethod
    def _pad8(b):
        return b + b"\x00" * (-len(b) % 8)

    @staticmethod
    def float_type(size=8):
        if size != 8:
            raise ValueError("float64 only")
        # class 1, version 1; bit field: little endian, mantissa normalisation = implied (bits 4-5 = 2), sign bit 63
        return struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)

    @classmethod
    def array_type(cls, dims):
        base = cls.float_type()
        size = 8 * int(np.prod(dims))
        body = struct.pack("<B3x", len(dims)) + struct.pack("<%dI" % len(dims), *dims) + \
            struct.pack("<%dI" % len(dims), *range(len(dims))) + base
        return struct.pack("<BBBBI", 0x2A, 0, 0, 0, size) + body              # class 10, version 2

    @staticmethod
    def dataspace(shape, maxshape=None):
        flags = 1 if maxshape is not None else 0
        out = struct.pack("<BBB5x", 1, len(shape), flags) + struct.pack("<%dQ" % len(shape), *shape)
        if maxshape is not None:
            out += struct.pack("<%dQ" % len(shape), *[UNDEF if m is None else m for m in maxshape])
        return out

    def attribute(self, name, array):
        array = np.ascontiguousarray(array, dtype="<f8")
        nm = name.encode() + b"\x00"
        dt, ds = self.float_type(), self.dataspace(array.shape)
        return struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + self._pad8(nm) + self._pad8(dt) + \
            self._pad8(ds) + array.tobytes()

    def object_header(self, messages):
        body = b""
        for mtype, data in messages:
            data = self._pad8(data)
            body += struct.pack("<HHB3x", mtype, len(data), 0) + data
        pos = self.alloc(16 + len(body))
        self.put(pos, struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body)
        return pos

    def group(self, links, attrs=()):
        """links: {name: object header address} -> object header address of a new old-style group."""
        names = sorted(links)
        heap_data = bytearray(b"\x00" * 8)                 # offset 0: the empty string (first B-tree key)
        name_off = {}
        for name in names:
            name_off[name] = len(heap_data)
            heap_data += self._pad8(name.encode() + b"\x00")
        heap_data += b"\x00" * 16
        data_pos = self.alloc(len(heap_data))
        self.put(data_pos, bytes(heap_data))
        heap_pos = self.alloc(32)
        self.put(heap_pos, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, data_pos))
        # symbol table nodes of at most 8 entries (2 * group leaf K), one B-tree leaf node over them
        per_node = 8
        groups = [names[i:i + per_node] for i in range(0, len(names), per_node)] or [[]]
        if len(groups) > 32:
            raise ValueError("too many links for the fixture writer")
        children, keys = [], [0]
        for chunk in groups:
            node = self.alloc(8 + 40 * per_node)
            raw = b"SNOD" + struct.pack("<BxH", 1, len(chunk))
            for name in chunk:
                raw += struct.pack("<QQI4x16x", name_off[name], links[name], 0)
            self.put(node, raw)
            children.append(node)
            keys.append(name_off[chunk[-1]] if chunk else 0)
        tree = self.alloc(24 + 8 * (2 * 16 + 1) + 8 * (2 * 16))
        raw = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(children), UNDEF, UNDEF)
        for i, child in enumerate(children):
            raw += struct.pack("<QQ", keys[i], child)
        raw += struct.pack("<Q", keys[-1])
        self.put(tree, raw)
        messages = [(0x0011, struct.pack("<QQ", tree, heap_pos))] + [(0x000C, self.attribute(k, v)) for k, v in attrs]
        return self.object_header(messages), tree, heap_pos

    def chunked_rows(self, rows, chunk_rows, node_entries=64):
        """rows float64[N, 2, M] -> object header address of a 1-D chunked dataset of element type (2, M) f8."""
        rows = np.ascontiguousarray(rows, dtype="<f8")
        n = rows.shape[0]
        sub = tuple(int(d) for d in rows.shape[1:])
        row_bytes = 8 * int(np.prod(sub))
        entries = []
        for first in range(0, n, chunk_rows):
            block = np.zeros((chunk_rows,) + sub, dtype="<f8")
            part = rows[first:first + chunk_rows]
            block[:len(part)] = part
            pos = self.alloc(chunk_rows * row_bytes)
            self.put(pos, block.tobytes())
            entries.append((first, pos))

        def node(level, items):
            """items: [(first row, child address)] -> address of a B-tree node"""
            key_size = 8 + 8 * 2
            pos = self.alloc(24 + (2 * 32 + 1) * key_size + 2 * 32 * 8)
            raw = b"TREE" + struct.pack("<BBHQQ", 1, level, len(items), UNDEF, UNDEF)
            for first, child in items:
                raw += struct.pack("<IIQQ", chunk_rows * row_bytes, 0, first, 0) + struct.pack("<Q", child)
            raw += struct.pack("<IIQQ", 0, 0, (items[-1][0] + chunk_rows) if level == 0 else n, 0)
            self.put(pos, raw)
            return pos
        btree = UNDEF
        if entries:
            level, items = 0, entries
            while True:
                groups = [items[i:i + node_entries] for i in range(0, len(items), node_entries)]
                items = [(g[0][0], node(level, g)) for g in groups]
                if len(items) == 1:
                    btree = items[0][1]
                    break
                level += 1
        layout = struct.pack("<BBB", 3, 2, 2) + struct.pack("<Q", btree) + struct.pack("<II", chunk_rows, row_bytes)
        fill = struct.pack("<BBBB", 2, 1, 2, 0)            # fill value message v2: allocate late, write if set, undefined
        messages = [(0x0001, self.dataspace((n,), (None,))), (0x0003, self.array_type(sub)), (0x0005, fill),
                    (0x0008, layout)]
        return self.object_header(messages)


def write_mlmc_file(path, level_rows, level_parameters, n_ops=None, chunk_rows=None):
    """Write sample rows in the reference's HDF5 structure (``mlmc/tool/hdf5.py:14-45``): root attribute
    ``level_parameters``, group ``Levels``, per level a group ``<l>`` with attribute ``n_ops_estimate`` = (time, count)
    and the chunked dataset ``collected_values`` ``[N]`` of element type ``(2, M)`` float64.  Level 0 keeps its zero
    coarse row, as the reference stores it.  ``chunk_rows``: rows per chunk (default: ~64 KiB chunks)."""
    w = _Writer()
    super_pos = w.alloc(96)
    level_links = {}
    for l, rows in enumerate(level_rows):
        rows = np.asarray(rows, dtype=np.float64)
        if rows.ndim != 3:
            raise ValueError("level rows must be [N, 2, M]")
        if rows.shape[1] == 1:
            rows = np.concatenate([rows, np.zeros_like(rows)], axis=1)
        c_rows = chunk_rows or max(1, (64 << 10) // (16 * rows.shape[2]))
        dset = w.chunked_rows(rows, c_rows)
        ops = (float(n_ops[l]), 1.0) if n_ops is not None else (0.0, 0.0)
        group, _t, _h = w.group({"collected_values": dset}, attrs=[("n_ops_estimate", np.array(ops))])
        level_links[str(l)] = group
    levels_group, _t, _h = w.group(level_links)
    root, root_tree, root_heap = w.group({"Levels": levels_group},
                                         attrs=[("level_parameters", np.asarray(level_parameters, dtype=np.float64))])
    eof = len(w.buf)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQI4xQQ", 0, root, 1, root_tree, root_heap)       # root symbol table entry (cached B-tree / heap)
    w.put(super_pos, sb)
    with open(path, "wb") as f:
        f.write(bytes(w.buf))
    return path
