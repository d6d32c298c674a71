"""Minimal HDF5 reader (and fixture writer) for ilastik's exported class / probability images.

The reference's only real input path is ``h5py.File(path)[first key][()]`` on an ilastik export
(tiff_analysis.py:118-120, :639-641; refine_boundaries.py:28-31 reads ``"exported_data"``).  h5py is not
installable in this image, so this module reads the subset of the HDF5 file format those files use, straight
from the published format specification ("HDF5 File Format Specification Version 3.0"):

* superblock versions 0 / 1 (what h5py and ilastik write by default) and 2 / 3;
* groups stored as a symbol table: version-1 B-tree (node type 0) + local heap + symbol-table nodes; version-2
  object headers with link messages are followed as well (compact groups);
* object headers version 1 (and 2), continuation blocks included;
* dataspace (simple, versions 1 and 2), datatype classes 0 (fixed point) and 1 (floating point), either byte order;
* data layout message version 3 (and 1 / 2): compact, contiguous, chunked with a version-1 B-tree chunk index;
* filter pipeline versions 1 / 2 with deflate (id 1) and shuffle (id 2) -- what ``compression="gzip"`` produces.

Not handled (raises ``H5Error``): version-4 layouts (``libver="latest"`` chunk indices), variable-length / compound
types, external storage, SZIP / LZF / other filters, virtual datasets.

Usage mirrors the h5py calls of the reference::

    with h5_io.File(path) as f:
        key = next(iter(f.keys()))
        ds_arr = f[key][()]

``write_dataset`` writes a single-group file in the same subset (contiguous or chunked, optionally shuffle + deflate)
-- used for the tests' fixtures and to hand images to tools that expect ilastik-style input.

STATUS: no HDF5 library exists in this image, so the reader is checked against this module's own writer, against
structures assembled field by field in the tests from the specification, and against the one libhdf5-written file
the image holds (a MATLAB v7.3 file from scipy's test data: user block, version-0 superblock, symbol-table group,
contiguous float64 dataset; tests/test_h5_io.py::test_reads_a_file_written_by_libhdf5).  Chunked / deflated datasets
written by libhdf5 -- ilastik's default export layout -- have not been read yet: treat those as unverified.
"""

import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


class _Buf:
    """Random-access little-endian reader over the whole file (ilastik exports of 2048^2 images are a few MB)."""

    def __init__(self, data):
        self.d = data

    def u(self, off, n):
        if off + n > len(self.d):
            raise H5Error(f"read of {n} bytes at {off} runs past the end of the file ({len(self.d)} bytes)")
        return int.from_bytes(self.d[off : off + n], "little")

    def bytes(self, off, n):
        if off + n > len(self.d):
            raise H5Error(f"read of {n} bytes at {off} runs past the end of the file ({len(self.d)} bytes)")
        return self.d[off : off + n]


class Dataset:
    def __init__(self, f, name, shape, dtype, layout, filters, fill):
        self._f, self.name, self.shape, self.dtype, self._layout, self._filters, self._fill = f, name, tuple(shape), dtype, layout, filters, fill

    @property
    def ndim(self):
        return len(self.shape)

    def __getitem__(self, key):
        a = self._read()
        return a if key == () or key is Ellipsis else a[key]

    def __array__(self, dtype=None, copy=None):
        a = self._read()
        return a if dtype is None else a.astype(dtype)

    def _read(self):
        b, lay = self._f._b, self._layout
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        isz = self.dtype.itemsize
        if lay[0] == "compact":
            return np.frombuffer(lay[1], dtype=self.dtype, count=n).reshape(self.shape).copy()
        if lay[0] == "contiguous":
            addr = lay[1]
            if addr == UNDEF:  # never written: the fill value
                return np.full(self.shape, self._fill, dtype=self.dtype)
            return np.frombuffer(b.bytes(self._f._base + addr, n * isz), dtype=self.dtype, count=n).reshape(self.shape).copy()
        # chunked
        _, btree, cdims = lay
        rank = len(self.shape)
        if len(cdims) != rank + 1 or cdims[-1] != isz:
            raise H5Error(f"chunk dimensions {cdims} do not fit a rank-{rank} dataset of {isz}-byte elements")
        cshape = tuple(cdims[:-1])
        out = np.full(self.shape, self._fill, dtype=self.dtype)
        if btree != UNDEF:
            for size, mask, offs, addr in self._f._chunks(btree, rank):
                raw = b.bytes(self._f._base + addr, size)
                for i, (fid, cd) in reversed(list(enumerate(self._filters))):  # undo the pipeline back to front
                    if mask & (1 << i):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else isz
                        raw = np.frombuffer(raw, dtype=np.uint8).reshape(es, -1).T.tobytes() if es > 1 and len(raw) % es == 0 else raw
                    else:
                        raise H5Error(f"dataset {self.name!r}: filter id {fid} is not supported (deflate = 1 and shuffle = 2 are)")
                chunk = np.frombuffer(raw, dtype=self.dtype, count=int(np.prod(cshape, dtype=np.int64))).reshape(cshape)
                sel_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, self.shape))
                sel_in = tuple(slice(0, sl.stop - sl.start) for sl in sel_out)
                out[sel_out] = chunk[sel_in]
        return out


class File:
    """Read-only view of an HDF5 file's root group: ``keys()``, ``f[name]`` -> ``Dataset``, context manager."""

    def __init__(self, path, mode="r"):
        if mode != "r":
            raise H5Error("this reader is read-only; see write_dataset for fixtures")
        with open(path, "rb") as fh:
            self._b = _Buf(fh.read())
        self._parse_superblock()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self._b = None
        return False

    # ---------------------------------------------------------------- superblock
    def _parse_superblock(self):
        b = self._b
        at = 0
        while True:  # the superblock may sit at 0, 512, 1024, ... (user block)
            if at + 8 > len(b.d):
                raise H5Error("not an HDF5 file: signature not found")
            if b.bytes(at, 8) == SIGNATURE:
                break
            at = 512 if at == 0 else at * 2
        ver = b.u(at + 8, 1)
        self._links = None
        if ver in (0, 1):
            self._so, self._sl = b.u(at + 13, 1), b.u(at + 14, 1)
            if self._so != 8 or self._sl != 8:
                raise H5Error(f"offsets / lengths of {self._so} / {self._sl} bytes are not supported (8 / 8 are)")
            p = at + 24 + (4 if ver == 1 else 0)
            self._base = b.u(p, 8)
            root = p + 32  # base, free-space, end-of-file, driver-info addresses, then the root symbol table entry
            self._root_header = b.u(root + 8, 8)
            cache = b.u(root + 16, 4)
            self._root_btree, self._root_heap = (b.u(root + 24, 8), b.u(root + 32, 8)) if cache == 1 else (None, None)
        elif ver in (2, 3):
            self._so, self._sl = b.u(at + 9, 1), b.u(at + 10, 1)
            if self._so != 8 or self._sl != 8:
                raise H5Error(f"offsets / lengths of {self._so} / {self._sl} bytes are not supported (8 / 8 are)")
            self._base = b.u(at + 12, 8)
            self._root_header = b.u(at + 36, 8)
            self._root_btree = self._root_heap = None
        else:
            raise H5Error(f"superblock version {ver} is not supported")
        if self._base == UNDEF:
            self._base = 0
        if self._root_btree is None:
            msgs = self._messages(self._root_header)
            st = [m for m in msgs if m[0] == 0x11]
            if st:
                self._root_btree, self._root_heap = self._b.u(st[0][1], 8), self._b.u(st[0][1] + 8, 8)
            else:  # compact group: link messages in the object header
                self._links = {}
                for t, off, size in msgs:
                    if t == 0x06:
                        name, addr = self._link(off)
                        if addr is not None:
                            self._links[name] = addr
        self._entries = None

    # ---------------------------------------------------------------- object headers
    def _messages(self, addr):
        """[(type, data offset, size)] of the object header at ``addr`` (continuation blocks followed)."""
        b, at = self._b, self._base + addr
        out = []
        if b.bytes(at, 4) == b"OHDR":  # version 2
            flags = b.u(at + 5, 1)
            p = at + 6
            if flags & 0x20:
                p += 16  # four timestamps
            if flags & 0x10:
                p += 4  # max compact / min dense attribute counts
            szlen = 1 << (flags & 3)
            chunk = b.u(p, szlen)
            p += szlen
            blocks = [(p, chunk)]
            track = bool(flags & 0x04)
            while blocks:
                p, n = blocks.pop(0)
                end = p + n
                while p + 4 + (2 if track else 0) <= end:
                    t, size, _fl = b.u(p, 1), b.u(p + 1, 2), b.u(p + 3, 1)
                    p += 4 + (2 if track else 0)
                    if t == 0x10:
                        blocks.append((self._base + b.u(p, 8) + 4, b.u(p + 8, 8) - 8))  # skip "OCHK", drop the checksum
                    elif t != 0:
                        out.append((t, p, size))
                    p += size
            return out
        ver = b.u(at, 1)
        if ver != 1:
            raise H5Error(f"object header version {ver} at {addr} is not supported")
        nmsg, hsize = b.u(at + 2, 2), b.u(at + 8, 4)
        blocks = [(at + 16, hsize)]
        while blocks and len(out) < nmsg + 64:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end:
                t, size = b.u(p, 2), b.u(p + 2, 2)
                p += 8
                if t == 0x10:
                    blocks.append((self._base + b.u(p, 8), b.u(p + 8, 8)))
                elif t != 0:
                    out.append((t, p, size))
                p += (size + 7) & ~7
        return out

    def _link(self, off):
        b = self._b
        flags = b.u(off + 1, 1)
        p = off + 2
        ltype = 0
        if flags & 0x08:
            ltype = b.u(p, 1)
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        ln = 1 << (flags & 3)
        n = b.u(p, ln)
        p += ln
        name = b.bytes(p, n).decode("utf-8")
        p += n
        return name, (b.u(p, 8) if ltype == 0 else None)

    # ---------------------------------------------------------------- groups
    def _heap_name(self, off):
        b = self._b
        h = self._base + self._root_heap
        if b.bytes(h, 4) != b"HEAP":
            raise H5Error("local heap signature not found")
        data = self._base + b.u(h + 24, 8)
        end = b.d.index(b"\0", data + off)
        return b.d[data + off : end].decode("utf-8")

    def _walk_group(self, node, out):
        b = self._b
        at = self._base + node
        sig = b.bytes(at, 4)
        if sig == b"TREE":
            if b.u(at + 4, 1) != 0:
                raise H5Error("group B-tree expected (node type 0)")
            n = b.u(at + 6, 2)
            p = at + 8 + 16 + 8  # siblings, first key
            for _ in range(n):
                self._walk_group(b.u(p, 8), out)
                p += 16  # child + next key
        elif sig == b"SNOD":
            n = b.u(at + 6, 2)
            p = at + 8
            for _ in range(n):
                out.append((self._heap_name(b.u(p, 8)), b.u(p + 8, 8)))
                p += 40
        else:
            raise H5Error(f"unexpected node signature {sig!r} in a group")

    def _load(self):
        if self._entries is None:
            if self._links is not None:
                self._entries = dict(sorted(self._links.items()))
            else:
                found = []
                self._walk_group(self._root_btree, found)
                self._entries = dict(sorted(found))  # h5py iterates a group in name order
        return self._entries

    def keys(self):
        return self._load().keys()

    def __iter__(self):
        return iter(self._load())

    def __contains__(self, name):
        return name in self._load()

    # ---------------------------------------------------------------- datasets
    def __getitem__(self, name):
        entries = self._load()
        if name not in entries:
            raise KeyError(f"Unable to open object (object '{name}' doesn't exist)")
        b = self._b
        shape = dtype = layout = None
        filters, fill = [], 0
        for t, off, size in self._messages(entries[name]):
            if t == 0x01:
                ver, rank = b.u(off, 1), b.u(off + 1, 1)
                p = off + (8 if ver == 1 else 4)
                shape = [b.u(p + 8 * i, 8) for i in range(rank)]
            elif t == 0x03:
                dtype = self._datatype(off)
            elif t == 0x08:
                layout = self._layout_msg(off)
            elif t == 0x0B:
                filters = self._filters_msg(off)
            elif t == 0x11:
                raise H5Error(f"{name!r} is a group; only datasets of the root group are read")
        if shape is None or dtype is None or layout is None:
            raise H5Error(f"{name!r}: dataspace, datatype or layout message missing")
        return Dataset(self, name, shape, dtype, layout, filters, fill)

    def _datatype(self, off):
        b = self._b
        cv, bits0, bits1, size = b.u(off, 1), b.u(off + 1, 1), b.u(off + 2, 1), b.u(off + 4, 4)
        cls = cv & 0x0F
        order = ">" if (bits0 & 1) else "<"
        if cls == 0:
            kind = "i" if (bits0 & 0x08) else "u"
        elif cls == 1:
            kind = "f"
        else:
            raise H5Error(f"datatype class {cls} is not supported (fixed and floating point are)")
        if size not in (1, 2, 4, 8):
            raise H5Error(f"{size}-byte elements are not supported")
        return np.dtype(f"{'|' if size == 1 else order}{kind}{size}")

    def _layout_msg(self, off):
        b = self._b
        ver = b.u(off, 1)
        if ver == 3:
            cls = b.u(off + 1, 1)
            if cls == 0:
                n = b.u(off + 2, 2)
                return ("compact", b.bytes(off + 4, n))
            if cls == 1:
                return ("contiguous", b.u(off + 2, 8), b.u(off + 10, 8))
            if cls == 2:
                nd = b.u(off + 2, 1)
                return ("chunked", b.u(off + 3, 8), [b.u(off + 11 + 4 * i, 4) for i in range(nd)])
            raise H5Error(f"layout class {cls} is not supported")
        if ver in (1, 2):
            nd, cls = b.u(off + 1, 1), b.u(off + 2, 1)
            p = off + 8
            addr = None
            if cls != 0:
                addr = b.u(p, 8)
                p += 8
            dims = [b.u(p + 4 * i, 4) for i in range(nd)]
            p += 4 * nd
            if cls == 1:
                return ("contiguous", addr, 0)
            if cls == 2:
                return ("chunked", addr, dims + [b.u(p, 4)])
            n = b.u(p, 4)
            return ("compact", b.bytes(p + 4, n))
        raise H5Error(f"data layout message version {ver} is not supported (files written with libver='latest' use version 4)")

    def _filters_msg(self, off):
        b = self._b
        ver, n = b.u(off, 1), b.u(off + 1, 1)
        p = off + (8 if ver == 1 else 2)
        out = []
        for _ in range(n):
            fid = b.u(p, 2)
            p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = b.u(p, 2)
                p += 2
            p += 2  # flags
            ncd = b.u(p, 2)
            p += 2
            p += (nlen + 7) & ~7 if ver == 1 else nlen
            cd = [b.u(p + 4 * i, 4) for i in range(ncd)]
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _chunks(self, node, rank):
        """(size, filter mask, offsets, address) of every chunk under the version-1 B-tree node at ``node``."""
        b = self._b
        at = self._base + node
        if b.bytes(at, 4) != b"TREE" or b.u(at + 4, 1) != 1:
            raise H5Error("chunk B-tree (node type 1) expected")
        level, n = b.u(at + 5, 1), b.u(at + 6, 2)
        ksz = 8 + 8 * (rank + 1)
        p = at + 24
        for _ in range(n):
            size, mask = b.u(p, 4), b.u(p + 4, 4)
            offs = [b.u(p + 8 + 8 * i, 8) for i in range(rank)]
            child = b.u(p + ksz, 8)
            if level == 0:
                yield size, mask, offs, child
            else:
                yield from self._chunks(child, rank)
            p += ksz + 8


def read_first_dataset(path):
    """``h5py.File(path)[next(iter(f.keys()))][()]`` (tiff_analysis.py:118-120, :639-641)."""
    with File(path) as f:
        return f[next(iter(f.keys()))][()]


# ---------------------------------------------------------------------------------------------- fixture writer
def _dtype_msg(dt):
    dt = np.dtype(dt)
    be = 1 if dt.byteorder == ">" else 0
    if dt.kind in "ui":
        bits0 = be | (0x08 if dt.kind == "i" else 0)
        return struct.pack("<BBBBIHH", 0x10 | 0, bits0, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
    if dt.kind == "f":
        if dt.itemsize == 4:
            prop = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits1 = 31
        elif dt.itemsize == 8:
            prop = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits1 = 63
        else:
            raise H5Error("float16 / float128 are not written")
        return struct.pack("<BBBBI", 0x10 | 1, be | 0x20, bits1, 0, dt.itemsize) + prop
    raise H5Error(f"dtype {dt} is not written")


def _msg(mtype, data, flags=0):
    pad = (-len(data)) % 8
    return struct.pack("<HHBBBB", mtype, len(data) + pad, flags, 0, 0, 0) + data + b"\0" * pad


def _object_header(msgs):
    body = b"".join(msgs)
    return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0" * 4 + body


def write_dataset_v2(path, arrays):
    """The same content in the newer structures (what ``libver="latest"`` groups look like): superblock version 2, a root
    group whose version-2 object header ("OHDR") carries one link message per dataset, version-2 dataset headers with a
    version-2 dataspace and a contiguous version-3 layout.  Checksum fields are left zero (the reader does not verify
    them).  Fixture for the reader's second code path."""
    if isinstance(arrays, np.ndarray):
        arrays = {"exported_data": arrays}
    names = sorted(arrays)
    out = bytearray(b"\0" * 48)  # superblock v2: 8 + 4 + 4 x 8 + 4

    def put(blob):
        while len(out) % 8:
            out.append(0)
        at = len(out)
        out.extend(blob)
        return at

    def ohdr(msgs):
        body = b"".join(struct.pack("<BHB", t, len(d), 0) + d for t, d in msgs)
        return b"OHDR" + bytes([2, 0x02]) + struct.pack("<I", len(body)) + body + b"\0" * 4  # flags 0x02: 4-byte chunk size; checksum 0

    headers = {}
    for nm in names:
        a = np.ascontiguousarray(arrays[nm])
        data_at = put(a.tobytes())
        space = struct.pack("<BBBB", 2, a.ndim, 0, 1) + b"".join(struct.pack("<Q", s) for s in a.shape)
        headers[nm] = put(ohdr([(0x01, space), (0x03, _dtype_msg(a.dtype)), (0x08, struct.pack("<BBQQ", 3, 1, data_at, a.nbytes))]))
    links = []
    for nm in names:
        raw = nm.encode("utf-8")
        links.append((0x06, struct.pack("<BBB", 1, 0, len(raw)) + raw + struct.pack("<Q", headers[nm])))  # version 1, flags 0: hard link, 1-byte name length
    root = put(ohdr(links))
    eof = len(out)
    sb = SIGNATURE + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, UNDEF, eof, root) + b"\0" * 4
    out[: len(sb)] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(out))
    return path


def write_dataset(path, arrays, chunks=None, compression=None, shuffle=False):
    """Write ``{name: array}`` (or a single array under ``"exported_data"``) as datasets of the root group of a new
    file: superblock 0, symbol-table group, version-1 object headers, layout version 3 -- contiguous, or chunked
    (``chunks=(..)``) with an optional shuffle + deflate pipeline (``compression="gzip"``)."""
    if isinstance(arrays, np.ndarray):
        arrays = {"exported_data": arrays}
    names = sorted(arrays)
    if len(names) > 8:
        raise H5Error("the fixture writer puts at most 8 datasets into its single symbol-table node")
    out = bytearray(b"\0" * 96)  # superblock filled in at the end

    def put(blob, align=8):
        while len(out) % align:
            out.append(0)
        at = len(out)
        out.extend(blob)
        return at

    # local heap data: "" at offset 0, then the names
    heap = bytearray(b"\0" * 8)
    name_off = {}
    for nm in names:
        name_off[nm] = len(heap)
        heap.extend(nm.encode("utf-8") + b"\0")
        while len(heap) % 8:
            heap.append(0)
    free_off = len(heap)
    heap.extend(struct.pack("<QQ", 1, 16))  # one free block: next = 1 (none), size 16
    heap_data = put(bytes(heap))
    heap_hdr = put(b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(heap), free_off, heap_data))
    headers = {}
    for nm in names:
        a = np.ascontiguousarray(arrays[nm])
        dt = a.dtype
        space = struct.pack("<BBBBI", 1, a.ndim, 0, 0, 0) + b"".join(struct.pack("<Q", s) for s in a.shape)
        msgs = [_msg(0x01, space), _msg(0x03, _dtype_msg(dt), flags=1)]
        if chunks is None:
            data_at = put(a.tobytes())
            msgs.append(_msg(0x08, struct.pack("<BBQQ", 3, 1, data_at, a.nbytes)))
        else:
            cshape = tuple(int(c) for c in chunks)
            if len(cshape) != a.ndim:
                raise H5Error("chunks must have one entry per dimension")
            pipeline = ([(2, [dt.itemsize])] if shuffle else []) + ([(1, [4])] if compression in ("gzip", "deflate") else [])
            entries = []
            grid = [range(0, s, c) for s, c in zip(a.shape, cshape)]
            for offs in np.stack(np.meshgrid(*grid, indexing="ij"), -1).reshape(-1, a.ndim) if a.ndim else []:
                block = np.zeros(cshape, dtype=dt)
                sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, a.shape))
                block[tuple(slice(0, sl.stop - sl.start) for sl in sel)] = a[sel]
                raw = block.tobytes()
                for fid, cd in pipeline:
                    if fid == 2 and dt.itemsize > 1:
                        raw = np.frombuffer(raw, dtype=np.uint8).reshape(-1, dt.itemsize).T.tobytes()
                    elif fid == 1:
                        raw = zlib.compress(raw, cd[0])
                entries.append((len(raw), [int(o) for o in offs], put(raw)))
            # version-1 B-tree of the chunks: leaves of at most 64 entries (libhdf5's default 2K for chunk trees),
            # internal levels above them while one level holds more than one node; a node's key i is the key of the
            # left-most chunk under child i, its last key the upper bound (the dataset shape for the right-most node)
            def key(size, offs):
                return struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)

            level, nodes = 0, [(e[0], e[1], e[2]) for e in entries]  # (size, offsets, address) of chunks, then of nodes
            end_key = key(0, list(a.shape))
            while True:
                groups = [nodes[i : i + 64] for i in range(0, len(nodes), 64)] or [[]]
                bodies = []
                for gi, grp in enumerate(groups):
                    body = bytearray()
                    for size, offs, addr in grp:
                        body += key(size, offs) + struct.pack("<Q", addr)
                    nxt = groups[gi + 1][0] if gi + 1 < len(groups) else None
                    body += key(nxt[0], nxt[1]) if nxt else end_key
                    bodies.append(body)
                # sibling addresses: the nodes of a level are written back to back, so they are known up front
                sizes = [24 + len(bd) for bd in bodies]
                base = put(b"")  # aligned position of the level's first node
                starts = [base + sum(sizes[:i]) for i in range(len(sizes))]
                written = []
                for gi, (grp, bd) in enumerate(zip(groups, bodies)):
                    left = starts[gi - 1] if gi > 0 else UNDEF
                    right = starts[gi + 1] if gi + 1 < len(groups) else UNDEF
                    at = put(b"TREE" + bytes([1, level]) + struct.pack("<HQQ", len(grp), left, right) + bytes(bd))
                    assert at == starts[gi]
                    written.append((grp[0][0] if grp else 0, grp[0][1] if grp else [0] * a.ndim, at))
                if len(written) == 1:
                    btree_at = written[0][2]
                    break
                level, nodes = level + 1, written
            msgs.append(_msg(0x08, struct.pack("<BBBQ", 3, 2, a.ndim + 1, btree_at) + b"".join(struct.pack("<I", c) for c in cshape) + struct.pack("<I", dt.itemsize)))
            if pipeline:
                fl = struct.pack("<BB6x", 1, len(pipeline))
                for fid, cd in pipeline:
                    fl += struct.pack("<HHHH", fid, 0, 1 if fid == 2 else 0, len(cd)) + b"".join(struct.pack("<I", c) for c in cd) + (b"\0" * 4 if len(cd) % 2 else b"")
                msgs.append(_msg(0x0B, fl))
        headers[nm] = put(_object_header(msgs))
    snod = bytearray(b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(names)))
    for nm in names:
        snod += struct.pack("<QQII16x", name_off[nm], headers[nm], 0, 0)
    snod += b"\0" * (40 * (8 - len(names)))
    snod_at = put(bytes(snod))
    tree = b"TREE" + bytes([0, 0]) + struct.pack("<HQQ", 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_at, name_off[names[-1]] if names else 0)
    tree_at = put(tree + b"\0" * (8 * 2 * 16))  # room for the unused keys / children of a 2K = 32 entry node
    root = put(_object_header([_msg(0x11, struct.pack("<QQ", tree_at, heap_hdr))]))
    eof = len(out)
    sb = SIGNATURE + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0) + struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree_at, heap_hdr)
    out[: len(sb)] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(out))
    return path
