"""Uncompressed multi-page TIFF reader / writer for z-stacks.

``split_zstack.py:50-65`` reads a ``(Z, C, Y, X)`` stack with ``tifffile.TiffReader(...).asarray()``
and writes every plane with ``tifffile.TiffWriter(..., bigtiff=False).write(channel)``.  tifffile is
not installable here, and the device pipeline wants the pixels in pinned host memory anyway, so this
module does the two things that path needs and nothing else:

* ``read_stack(path)``        -> numpy array shaped like ``TiffReader.asarray()`` for microscope stacks:
  ``(pages, Y, X)`` folded to ``(Z, C, Y, X)`` / ``(T, Z, C, Y, X)`` when the first page carries an
  ImageJ (``images= channels= slices= frames=``) or tifffile (``{"shape": [...]}``) description;
* ``read_stack_pinned(path)`` -> the same pixels read straight into a pinned ``torch`` tensor
  (``readinto`` per strip run, no intermediate copy), ready for ``segment_zstack_pinned``;
* ``write_plane`` / ``write_stack`` -> classic little-endian TIFF, one strip per page, min-is-black,
  with the ImageJ description a hyperstack needs to be folded back.

Supported: classic TIFF and BigTIFF, both byte orders, uncompressed strips (any rows per strip),
1 sample per pixel, uint8 / uint16 / uint32 / int8 / int16 / int32 / float32 / float64.  Anything else
(compression, tiles, RGB) raises ``TiffError`` -- the pipeline never guesses at pixels.
"""

import json
import re
import struct

import numpy as np

__all__ = ["TiffError", "read_pages", "read_stack", "read_stack_pinned", "write_plane", "write_stack"]


class TiffError(ValueError):
    pass


_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8, 17: 8, 18: 8}
_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 6: "b", 7: "B", 8: "h", 9: "i", 11: "f", 12: "d", 16: "Q", 17: "q", 18: "Q"}
_DTYPES = {(1, 8): "u1", (1, 16): "u2", (1, 32): "u4", (2, 8): "i1", (2, 16): "i2", (2, 32): "i4", (3, 32): "f4", (3, 64): "f8"}
# tags
_W, _H, _BITS, _COMP, _PHOTO, _DESC, _STRIP_OFF, _SPP, _RPS, _STRIP_CNT, _PLANAR, _FMT = 256, 257, 258, 259, 262, 270, 273, 277, 278, 279, 284, 339
_TILE_W = 322


class _Page:
    __slots__ = ("width", "height", "dtype", "runs", "description")


def _read_ifds(f):
    head = f.read(16)
    if len(head) < 8 or head[:2] not in (b"II", b"MM"):
        raise TiffError("not a TIFF file")
    bo = "<" if head[:2] == b"II" else ">"
    magic = struct.unpack(bo + "H", head[2:4])[0]
    if magic == 42:
        big = False
        off = struct.unpack(bo + "I", head[4:8])[0]
    elif magic == 43:
        big = True
        if struct.unpack(bo + "HH", head[4:8]) != (8, 0):
            raise TiffError("malformed BigTIFF header")
        off = struct.unpack(bo + "Q", head[8:16])[0]
    else:
        raise TiffError(f"not a TIFF file (magic {magic})")
    cnt_fmt, ent_size, val_fmt, val_size = ("Q", 20, "Q", 8) if big else ("H", 12, "I", 4)
    pages, seen = [], set()
    while off:
        if off in seen:
            raise TiffError("IFD chain loops")
        seen.add(off)
        f.seek(off)
        (n,) = struct.unpack(bo + cnt_fmt, f.read(struct.calcsize(cnt_fmt)))
        raw = f.read(n * ent_size + val_size)
        if len(raw) < n * ent_size + val_size:
            raise TiffError("truncated IFD")
        tags = {}
        for i in range(n):
            e = raw[i * ent_size : (i + 1) * ent_size]
            tag, typ = struct.unpack(bo + "HH", e[:4])
            (count,) = struct.unpack(bo + val_fmt, e[4 : 4 + val_size])
            field = e[4 + val_size :]
            size = _TYPE_SIZE.get(typ)
            if size is None:
                continue
            nbytes = size * count
            if nbytes <= val_size:
                data = field[:nbytes]
            else:
                (ptr,) = struct.unpack(bo + val_fmt, field)
                here = f.tell()
                f.seek(ptr)
                data = f.read(nbytes)
                f.seek(here)
            if typ == 2:
                tags[tag] = data.split(b"\0")[0].decode("latin-1")
            elif typ in (5, 10):
                tags[tag] = struct.unpack(bo + ("II" if typ == 5 else "ii") * count, data)
            else:
                tags[tag] = struct.unpack(bo + _TYPE_FMT[typ] * count, data)
        (off,) = struct.unpack(bo + val_fmt, raw[n * ent_size :])
        pages.append(_page_from_tags(tags, bo))
    if not pages:
        raise TiffError("no image in the file")
    return pages


def _page_from_tags(tags, bo):
    def one(tag, default=None):
        v = tags.get(tag)
        if v is None:
            if default is None:
                raise TiffError(f"missing TIFF tag {tag}")
            return default
        return v[0] if isinstance(v, tuple) else v

    if one(_COMP, 1) != 1:
        raise TiffError(f"compressed TIFF (compression {one(_COMP)}) is not supported")
    if _TILE_W in tags:
        raise TiffError("tiled TIFF is not supported")
    if one(_SPP, 1) != 1:
        raise TiffError("only single-sample (grayscale) pages are supported")
    key = (one(_FMT, 1), one(_BITS, 1))
    if key not in _DTYPES:
        raise TiffError(f"unsupported sample format {key}")
    p = _Page()
    p.width, p.height = int(one(_W)), int(one(_H))
    p.dtype = np.dtype(bo + _DTYPES[key]) if key[1] > 8 else np.dtype(_DTYPES[key])
    offs, cnts = tags.get(_STRIP_OFF), tags.get(_STRIP_CNT)
    if offs is None:
        raise TiffError("missing strip offsets")
    rps = min(int(one(_RPS, p.height)), p.height)
    row_bytes = p.width * p.dtype.itemsize
    if cnts is None:  # allowed to be absent for a single strip
        cnts = tuple(min(rps, p.height - i * rps) * row_bytes for i in range(len(offs)))
    if sum(cnts) != p.height * row_bytes:
        raise TiffError("strip byte counts do not add up to the image size")
    runs = []  # contiguous (offset, nbytes) pieces, merged
    for o, c in zip(offs, cnts):
        if runs and runs[-1][0] + runs[-1][1] == o:
            runs[-1] = (runs[-1][0], runs[-1][1] + c)
        else:
            runs.append((int(o), int(c)))
    p.runs = runs
    p.description = tags.get(_DESC, "")
    return p


def _fold_shape(n_pages, description):
    """Leading axes of the stack from the first page's description; ``(n_pages,)`` when unknown."""
    if description.startswith("ImageJ="):
        kv = dict(re.findall(r"(\w+)=([^\n]+)", description))
        c, z, t = (int(kv.get(k, 1)) for k in ("channels", "slices", "frames"))
        if c * z * t == n_pages:
            return tuple(v for v in (t, z, c) if v > 1) or (1,)
    elif description.startswith("{"):
        try:
            shape = tuple(int(v) for v in json.loads(description)["shape"])
        except (ValueError, KeyError, TypeError):
            shape = ()
        if len(shape) >= 2 and int(np.prod(shape[:-2], dtype=np.int64)) == n_pages:
            return shape[:-2] or (1,)
    return (n_pages,)


def _check_uniform(pages):
    p0 = pages[0]
    for p in pages[1:]:
        if (p.width, p.height, p.dtype) != (p0.width, p0.height, p0.dtype):
            raise TiffError("pages differ in size or dtype: not a stack")
    return p0


def _fill(f, pages, out_bytes):
    """Read every page's strips into ``out_bytes`` (a writable flat uint8 view), in page order."""
    mv = memoryview(out_bytes)
    pos = 0
    for p in pages:
        for off, n in p.runs:
            f.seek(off)
            got = f.readinto(mv[pos : pos + n])
            if got != n:
                raise TiffError("truncated pixel data")
            pos += n


def read_pages(path):
    """``(pages, Y, X)`` array plus the first page's description."""
    with open(path, "rb") as f:
        pages = _read_ifds(f)
        p0 = _check_uniform(pages)
        out = np.empty((len(pages), p0.height, p0.width), dtype=p0.dtype.newbyteorder("="))
        _fill(f, pages, out.view(np.uint8).reshape(-1))
    if p0.dtype.itemsize > 1 and not p0.dtype.isnative:  # bytes were read as stored: bring them to host order
        out.byteswap(inplace=True)
    return out, p0.description


def read_stack(path):
    """The stack as ``tifffile`` would shape it for a microscope hyperstack: pages folded to the
    leading axes the description names (``(Z, C, Y, X)`` for split_zstack.py:50), a single page
    squeezed to ``(Y, X)``."""
    pages, desc = read_pages(path)
    lead = _fold_shape(pages.shape[0], desc)
    out = pages.reshape(lead + pages.shape[1:])
    return out[0] if out.shape[0] == 1 and len(lead) == 1 else out


def read_stack_pinned(path):
    """Same pixels, read directly into pinned host memory (a ``torch`` tensor), so the H2D copy of
    ``segment_zstack_pinned`` can run at full PCIe rate.  uint16 / uint8 little-endian stacks only
    (what the microscope writes); other files go through ``read_stack``."""
    import torch

    with open(path, "rb") as f:
        pages = _read_ifds(f)
        p0 = _check_uniform(pages)
        if p0.dtype.itemsize > 1 and p0.dtype.byteorder == ">":
            raise TiffError("big-endian stacks go through read_stack")
        tdt = {"u1": torch.uint8, "u2": torch.uint16, "i2": torch.int16, "i4": torch.int32, "f4": torch.float32, "f8": torch.float64}.get(p0.dtype.str[1:])
        if tdt is None:
            raise TiffError(f"no pinned path for dtype {p0.dtype}")
        lead = _fold_shape(len(pages), p0.description)
        t = torch.empty(lead + (p0.height, p0.width), dtype=tdt)
        if torch.cuda.is_available():
            t = t.pin_memory()
        _fill(f, pages, t.view(torch.uint8).reshape(-1).numpy())
    return t


# ------------------------------------------------------------------------------------ writer
def _ifd(entries, next_off):
    """Classic little-endian IFD from ``(tag, type, count, value_bytes_or_int)``; values must fit in 4 bytes."""
    out = struct.pack("<H", len(entries))
    for tag, typ, count, val in sorted(entries):
        out += struct.pack("<HHI", tag, typ, count) + (val if isinstance(val, bytes) else struct.pack("<I", val))
    return out + struct.pack("<I", next_off)


def write_stack(path, arr, description=None):
    """Write ``arr`` (``(..., Y, X)``) as an uncompressed classic TIFF, one page per leading index, one
    strip per page -- what ``TiffWriter(path, bigtiff=False).write(arr)`` produces for grayscale data up
    to 4 GiB.  A ``(Z, C, Y, X)`` array gets the ImageJ hyperstack description so that readers fold the
    pages back."""
    a = np.ascontiguousarray(arr)
    if a.ndim < 2:
        raise TiffError("need at least a 2-D image")
    fmt_bits = {v: k for k, v in _DTYPES.items()}.get(a.dtype.str[1:])
    if fmt_bits is None:
        raise TiffError(f"unsupported dtype {a.dtype}")
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    H, W = a.shape[-2:]
    lead = a.shape[:-2]
    n = int(np.prod(lead, dtype=np.int64)) if lead else 1
    if description is None and len(lead) >= 1 and n > 1:
        names = ["frames", "slices", "channels"][-len(lead):] if len(lead) <= 3 else None
        if names is None:
            raise TiffError("more than three leading axes")
        description = "ImageJ=1.11a\nimages=%d\n" % n + "".join("%s=%d\n" % (k, v) for k, v in zip(names, lead) if v > 1) + "hyperstack=true\n"
    desc = (description or "").encode("latin-1") + b"\0"
    page_bytes = H * W * a.dtype.itemsize
    # layout: header | per page: [description (first page)] IFD | pixel data of all pages
    n_entries = 11
    ifd_size = 2 + 12 * n_entries + 4
    first_extra = (len(desc) + 1) & ~1
    meta = 8 + first_extra + (ifd_size + 12) * n  # +12: one more entry on the first page rounds up safely
    data0 = (meta + 15) & ~15
    if data0 + page_bytes * n >= 1 << 32:
        raise TiffError("stack does not fit a classic TIFF (4 GiB)")
    with open(path, "wb") as f:
        f.write(b"II" + struct.pack("<HI", 42, 8 + first_extra))
        f.write(desc.ljust(first_extra, b"\0"))
        pos = 8 + first_extra
        flat = a.reshape(n, H, W)
        for i in range(n):
            entries = [
                (_W, 4, 1, W), (_H, 4, 1, H), (_BITS, 3, 1, fmt_bits[1]), (_COMP, 3, 1, 1), (_PHOTO, 3, 1, 1),
                (_STRIP_OFF, 4, 1, data0 + i * page_bytes), (_SPP, 3, 1, 1), (_RPS, 4, 1, H), (_STRIP_CNT, 4, 1, page_bytes),
                (_PLANAR, 3, 1, 1), (_FMT, 3, 1, fmt_bits[0]),
            ]
            if i == 0 and len(desc) > 1:
                entries.append((_DESC, 2, len(desc), 8))
            size = 2 + 12 * len(entries) + 4
            nxt = pos + size if i + 1 < n else 0
            f.write(_ifd(entries, nxt))
            pos += size
        f.write(b"\0" * (data0 - pos))
        f.write(flat.tobytes() if n * page_bytes < (1 << 26) else memoryview(flat).cast("B"))


def write_plane(path, plane):
    """One 2-D plane per file, as split_zstack.py:64-65 does for every ``z_slice[channel]``."""
    p = np.asarray(plane)
    if p.ndim != 2:
        raise TiffError(f"expected a 2-D plane, got shape {p.shape}")
    write_stack(path, p)
