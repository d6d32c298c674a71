"""ilastik ``.h5`` input path (tiff_analysis.py:118-120, :639-641; refine_boundaries.py:28-31) without h5py.

h5py / libhdf5 are not installable in this image, so ``h5_io.File`` is checked against (1) files produced by this
repo's own fixture writer in every layout the reader supports, (2) the byte positions the HDF5 File Format
Specification gives for the structures the writer emits and (3) the ONE file in this image that libhdf5 itself
wrote: scipy's MATLAB v7.3 test file (``tests/golden/libhdf5_matlab73_testdouble.mat``, a copy of
``scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat``, BSD-3; MATLAB R2008 = HDF5 1.6/1.8: 512-byte user block,
version-0 superblock, symbol-table root group, contiguous float64 dataset with attributes).  Chunked / filtered
datasets written by libhdf5 remain unverified (no such file exists here)."""
import os
import struct

import numpy as np
import pytest

from particle_col_image_segmentation_b200 import h5_io, synth


@pytest.mark.parametrize("dtype", ["uint8", "uint16", "int32", "float32", "float64", ">u2", ">f4"])
@pytest.mark.parametrize("layout", ["contiguous", "chunked", "gzip", "shuffle+gzip"])
def test_roundtrip(tmp_path, dtype, layout):
    rng = np.random.default_rng(1)
    a = (rng.random((3, 37, 53)) * 200).astype(dtype)
    kw = {}
    if layout != "contiguous":
        kw = dict(chunks=(1, 16, 32), compression="gzip" if "gzip" in layout else None, shuffle="shuffle" in layout)
    p = h5_io.write_dataset(str(tmp_path / "x.h5"), a, **kw)
    with h5_io.File(p) as f:
        assert list(f.keys()) == ["exported_data"]
        ds = f["exported_data"]
        assert ds.shape == a.shape and ds.dtype == a.dtype
        got = ds[()]
        assert got.dtype == a.dtype and np.array_equal(got, a)
        assert np.array_equal(np.array(ds), a) and np.array_equal(ds[1, 5:9], a[1, 5:9])


def test_reference_call_pattern_on_a_class_image(tmp_path):
    """``next(iter(f.keys()))`` + ``[()]`` on an ilastik-style (H, W, 1) uint8 class image, then normalize_ds_arr."""
    img = synth.class_image(256, 320, seed=3)[:, :, None]
    p = h5_io.write_dataset(str(tmp_path / "simple_seg.h5"), {"exported_data": img, "zz_other": np.arange(5)}, chunks=None)
    with h5_io.File(p, "r") as f:
        key = next(iter(f.keys()))
        arr = f[key][()]
    assert key == "exported_data" and arr.shape == (256, 320, 1) and np.array_equal(arr, img)
    assert np.array_equal(h5_io.read_first_dataset(p), img)
    from particle_col_image_segmentation_b200 import tiff_analysis

    assert tiff_analysis.normalize_ds_arr(arr).shape == (256, 320)
    with pytest.raises(KeyError):
        with h5_io.File(p) as f:
            f["missing"]


def test_probability_stack_chunked_like_ilastik(tmp_path):
    """refine_boundaries.py:29-34: ``np.array(f["exported_data"])`` is (C, H, W) float32; channel 3 is the boundary map."""
    _, prob = synth.touching_particles(128, 160, seed=5, pitch=32.0)
    stack = np.stack([prob * 0.1, prob * 0.2, 1 - prob, prob]).astype(np.float32)
    p = h5_io.write_dataset(str(tmp_path / "probabilities.h5"), stack, chunks=(1, 64, 64), compression="gzip", shuffle=True)
    with h5_io.File(p) as f:
        got = np.array(f["exported_data"])
    assert got.dtype == np.float32 and np.array_equal(got, stack) and np.array_equal(got[3], prob)


def test_structures_sit_where_the_specification_puts_them(tmp_path):
    a = np.arange(24, dtype="<u2").reshape(4, 6)
    raw = open(h5_io.write_dataset(str(tmp_path / "s.h5"), a), "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0          # format signature, superblock version 0
    assert raw[13] == 8 and raw[14] == 8                              # size of offsets / lengths
    assert struct.unpack_from("<HH", raw, 16) == (4, 16)             # group leaf / internal node K
    base, free, eof, drv = struct.unpack_from("<QQQQ", raw, 24)
    assert base == 0 and free == h5_io.UNDEF and eof == len(raw) and drv == h5_io.UNDEF
    name_off, root_hdr, cache = struct.unpack_from("<QQI", raw, 56)  # root group symbol table entry
    btree, heap = struct.unpack_from("<QQ", raw, 80)
    assert cache == 1 and raw[btree : btree + 4] == b"TREE" and raw[heap : heap + 4] == b"HEAP"
    assert raw[root_hdr] == 1                                          # version-1 object header ...
    assert struct.unpack_from("<H", raw, root_hdr + 16)[0] == 0x11    # ... whose first message is the symbol table message
    assert struct.unpack_from("<QQ", raw, root_hdr + 24) == (btree, heap)
    snod = struct.unpack_from("<Q", raw, btree + 32)[0]               # child 0 of the root B-tree node
    assert raw[snod : snod + 4] == b"SNOD" and struct.unpack_from("<H", raw, snod + 6)[0] == 1
    link, hdr = struct.unpack_from("<QQ", raw, snod + 8)
    data_seg = struct.unpack_from("<Q", raw, heap + 24)[0]
    assert raw[data_seg + link : data_seg + link + 14] == b"exported_data\0"
    types = []
    p, end = hdr + 16, hdr + 16 + struct.unpack_from("<I", raw, hdr + 8)[0]
    while p < end:
        t, size = struct.unpack_from("<HH", raw, p)
        types.append(t)
        if t == 0x08:                                                  # contiguous layout: address and size of the raw data
            ver, cls, addr, nbytes = struct.unpack_from("<BBQQ", raw, p + 8)
            assert (ver, cls, nbytes) == (3, 1, a.nbytes) and raw[addr : addr + a.nbytes] == a.tobytes()
        p += 8 + size
    assert types == [0x01, 0x03, 0x08]


def test_unsupported_files_fail_loudly(tmp_path):
    p = tmp_path / "not.h5"
    p.write_bytes(b"II*\0" + b"\0" * 600)
    with pytest.raises(h5_io.H5Error):
        h5_io.File(str(p))
    q = h5_io.write_dataset(str(tmp_path / "v.h5"), np.zeros((4, 4), np.uint8))
    raw = bytearray(open(q, "rb").read())
    raw[8] = 7  # an unknown superblock version
    (tmp_path / "v7.h5").write_bytes(bytes(raw))
    with pytest.raises(h5_io.H5Error):
        h5_io.File(str(tmp_path / "v7.h5"))


def test_newer_structures_superblock2_ohdr_links(tmp_path):
    """Superblock version 2, version-2 object headers, link messages instead of a symbol table, version-2 dataspace."""
    a = np.arange(5 * 7, dtype=np.int32).reshape(5, 7) - 9
    b = np.linspace(0, 1, 12).reshape(3, 4)
    p = h5_io.write_dataset_v2(str(tmp_path / "v2.h5"), {"exported_data": a, "other": b})
    raw = open(p, "rb").read()
    assert raw[8] == 2 and raw[9] == 8 and raw[10] == 8                      # superblock version, offset / length sizes
    root = struct.unpack_from("<Q", raw, 36)[0]
    assert raw[root : root + 4] == b"OHDR" and raw[root + 4] == 2            # version-2 root object header
    with h5_io.File(p) as f:
        assert list(f.keys()) == ["exported_data", "other"] and "other" in f
        assert np.array_equal(f["exported_data"][()], a) and f["exported_data"].dtype == np.int32
        assert np.array_equal(np.array(f["other"]), b)


def test_reads_a_file_written_by_libhdf5():
    """MATLAB v7.3 files are HDF5 files behind a 512-byte user block.  The variable of this one, ``testdouble``, is
    0 : pi/4 : 2*pi as a 1x9 MATLAB row = a (9, 1) HDF5 dataset (MATLAB stores column-major); scipy ships the same
    variable as a version-5 MAT file, which scipy.io.loadmat reads independently of any HDF5 code."""
    p = os.path.join(os.path.dirname(__file__), "golden", "libhdf5_matlab73_testdouble.mat")
    raw = open(p, "rb").read()
    assert raw[:6] == b"MATLAB" and raw.find(b"\x89HDF\r\n\x1a\n") == 512  # the superblock is not at offset 0
    with h5_io.File(p) as f:
        assert list(f.keys()) == ["testdouble"]
        ds = f["testdouble"]
        assert ds.shape == (9, 1) and ds.dtype == np.float64
        got = ds[()]
    want = np.arange(0, 9) * (np.pi / 4)
    assert np.array_equal(got[:, 0], want)
    assert np.array_equal(h5_io.read_first_dataset(p)[:, 0], want)
    sio = pytest.importorskip("scipy.io")
    v5 = os.path.join(os.path.dirname(sio.__file__), "matlab", "tests", "data", "testdouble_7.4_GLNX86.mat")
    if os.path.exists(v5):
        assert np.array_equal(sio.loadmat(v5)["testdouble"], got.T)  # bit for bit what MATLAB wrote in the other format


@pytest.mark.parametrize("compression", [None, "gzip"])
def test_multi_level_chunk_btree(tmp_path, compression):
    """A 2048^2 ilastik export in 64x64 chunks has 1024 chunks: libhdf5 keeps at most 64 entries in a chunk B-tree node,
    so real files carry trees of two and more levels.  150 x 70 chunks of 8x8 = 10 500 chunks -> three levels."""
    rng = np.random.default_rng(5)
    a = rng.integers(0, 4, (1, 1200, 560), dtype=np.uint8)
    p = h5_io.write_dataset(str(tmp_path / "deep.h5"), a, chunks=(1, 8, 8), compression=compression)
    raw = open(p, "rb").read()
    levels = {raw[i + 5] for i in range(0, len(raw) - 8, 8) if raw[i : i + 4] == b"TREE" and raw[i + 4] == 1}
    assert levels == {0, 1, 2}
    with h5_io.File(p) as f:
        assert np.array_equal(f["exported_data"][()], a)
    b = rng.random((130, 70)).astype(np.float32)  # two levels, ragged edge chunks
    q = h5_io.write_dataset(str(tmp_path / "two.h5"), b, chunks=(8, 8), compression=compression, shuffle=bool(compression))
    assert np.array_equal(h5_io.read_first_dataset(q), b)
