"""TIFF z-stack reader / writer (split_zstack.py:50-65): pinned against files assembled
by hand from the TIFF 6.0 layout, against Pillow and OpenCV (independent codecs present in the image),
and by round trips."""

import os
import struct

import numpy as np
import pytest

from particle_col_image_segmentation_b200 import split_zstack, tiff_io


def _hand_tiff(pages, bo="<", big=False, rows_per_strip=None, description=None):
    """Minimal TIFF built straight from the specification, independent of tiff_io's writer."""
    e = bo
    out = bytearray()
    out += (b"II" if e == "<" else b"MM")
    if big:
        out += struct.pack(e + "HHHQ", 43, 8, 0, 0)
    else:
        out += struct.pack(e + "HI", 42, 0)
    ifd_ptr_pos = len(out) - (8 if big else 4)
    for pi, page in enumerate(pages):
        h, w = page.shape
        rps = rows_per_strip or h
        data = page.astype(page.dtype.newbyteorder(e)).tobytes()
        rb = w * page.dtype.itemsize
        strips = [data[i * rps * rb : (i + 1) * rps * rb] for i in range((h + rps - 1) // rps)]
        offs = []
        for s in strips:  # strips deliberately separated by padding so that they are not contiguous
            out += b"\xAA" * 6
            offs.append(len(out))
            out += s
        def arr(vals, fmt):
            nonlocal out
            if len(out) % 2:
                out += b"\0"
            pos = len(out)
            out += struct.pack(e + fmt * len(vals), *vals)
            return pos
        lng = "Q" if big else "I"
        ltyp = 16 if big else 4
        off_pos = arr(offs, lng) if len(offs) > 1 else offs[0]
        cnt_pos = arr([len(s) for s in strips], lng) if len(offs) > 1 else len(strips[0])
        kind = {"u": 1, "i": 2, "f": 3}[page.dtype.kind]
        ents = [(256, 3, 1, w), (257, 3, 1, h), (258, 3, 1, page.dtype.itemsize * 8), (259, 3, 1, 1), (262, 3, 1, 1),
                (273, ltyp, len(offs), off_pos), (277, 3, 1, 1), (278, 3, 1, rps), (279, ltyp, len(offs), cnt_pos), (339, 3, 1, kind)]
        if description and pi == 0:
            d = description.encode() + b"\0"
            dpos = len(out)
            out += d
            ents.append((270, 2, len(d), dpos))
        if len(out) % 2:
            out += b"\0"
        ifd = len(out)
        struct.pack_into(e + ("Q" if big else "I"), out, ifd_ptr_pos, ifd)
        out += struct.pack(e + ("Q" if big else "H"), len(ents))
        for tag, typ, cnt, val in sorted(ents):
            out += struct.pack(e + "HH", tag, typ)
            if big:
                out += struct.pack(e + "Q", cnt)
                out += struct.pack(e + ("H6x" if typ == 3 else "Q"), val)
            else:
                out += struct.pack(e + "I", cnt)
                out += struct.pack(e + ("H2x" if typ == 3 else "I"), val)
        ifd_ptr_pos = len(out)
        out += struct.pack(e + ("Q" if big else "I"), 0)
    return bytes(out)


@pytest.mark.parametrize("bo", ["<", ">"])
@pytest.mark.parametrize("big", [False, True])
@pytest.mark.parametrize("dtype,rps", [("u2", None), ("u2", 3), ("u1", 5), ("f4", None), ("i2", 2)])
def test_reader_against_hand_built_files(tmp_path, bo, big, dtype, rps):
    rng = np.random.default_rng(7)
    pages = [(rng.random((11, 13)) * 200).astype(dtype) for _ in range(6)]
    p = tmp_path / "hand.tif"
    p.write_bytes(_hand_tiff(pages, bo, big, rps, description="ImageJ=1.53\nimages=6\nchannels=2\nslices=3\nhyperstack=true\n"))
    got = tiff_io.read_stack(str(p))
    assert got.shape == (3, 2, 11, 13) and got.dtype == np.dtype(dtype)
    assert np.array_equal(got.reshape(6, 11, 13), np.stack(pages))
    flat, desc = tiff_io.read_pages(str(p))
    assert flat.shape == (6, 11, 13) and desc.startswith("ImageJ=")


def test_tifffile_shaped_description_and_plain_pages(tmp_path):
    pages = [np.full((4, 5), i, np.uint16) for i in range(8)]
    p = tmp_path / "shaped.tif"
    p.write_bytes(_hand_tiff(pages, description='{"shape": [2, 4, 4, 5]}'))
    assert tiff_io.read_stack(str(p)).shape == (2, 4, 4, 5)
    p.write_bytes(_hand_tiff(pages))
    assert tiff_io.read_stack(str(p)).shape == (8, 4, 5)
    p.write_bytes(_hand_tiff(pages[:1]))
    assert tiff_io.read_stack(str(p)).shape == (4, 5)


def test_writer_read_by_pillow_and_opencv(tmp_path):
    from PIL import Image, ImageSequence
    import cv2

    rng = np.random.default_rng(3)
    stack = rng.integers(0, 65535, (3, 2, 17, 23), dtype=np.uint16)
    p = str(tmp_path / "w.tif")
    tiff_io.write_stack(p, stack)
    with Image.open(p) as im:
        frames = [np.array(f) for f in ImageSequence.Iterator(im)]
    assert len(frames) == 6 and all(f.dtype == np.uint16 for f in frames)
    assert np.array_equal(np.stack(frames).reshape(stack.shape), stack)
    ok, mats = cv2.imreadmulti(p, flags=cv2.IMREAD_UNCHANGED)
    assert ok and np.array_equal(np.stack(mats).reshape(stack.shape), stack)
    assert np.array_equal(tiff_io.read_stack(p), stack)
    plane = rng.integers(0, 255, (9, 31), dtype=np.uint8)
    tiff_io.write_plane(p, plane)
    assert np.array_equal(np.array(Image.open(p)), plane) and np.array_equal(tiff_io.read_stack(p), plane)


def test_reader_on_pillow_and_opencv_files(tmp_path):
    from PIL import Image
    import cv2

    rng = np.random.default_rng(4)
    frames = [rng.integers(0, 65535, (12, 20), dtype=np.uint16) for _ in range(4)]
    p = str(tmp_path / "pil.tif")
    Image.fromarray(frames[0]).save(p, save_all=True, append_images=[Image.fromarray(f) for f in frames[1:]], compression=None)
    assert np.array_equal(tiff_io.read_stack(p), np.stack(frames))
    p2 = str(tmp_path / "cv.tif")
    assert cv2.imwritemulti(p2, frames, [cv2.IMWRITE_TIFF_COMPRESSION, 1])
    assert np.array_equal(tiff_io.read_stack(p2), np.stack(frames))
    p3 = str(tmp_path / "lzw.tif")
    Image.fromarray(frames[0]).save(p3, compression="tiff_lzw")
    with pytest.raises(tiff_io.TiffError, match="compress"):
        tiff_io.read_stack(p3)
    (tmp_path / "junk.tif").write_bytes(b"not a tiff at all")
    with pytest.raises(tiff_io.TiffError):
        tiff_io.read_stack(str(tmp_path / "junk.tif"))


def test_process_tif_layout(tmp_path, monkeypatch):
    """split_zstack.py:40-65: folder / file names and plane contents."""
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(5)
    os.makedirs("exp")
    z4 = rng.integers(0, 4000, (3, 4, 8, 10), dtype=np.uint16)
    src = "exp/Tp_1_CY5_RFP_GFP_DAPI_1_zstack.tif"
    tiff_io.write_stack(src, z4)
    assert split_zstack.get_clean_file_name(src) == ("_CY5_RFP_GFP_DAPI", "exp/Tp_1_1")
    written = split_zstack.process_tif(src, [1, 2])
    assert not os.path.exists(src) and os.path.exists("exp/Tp_1_1/Tp_1_CY5_RFP_GFP_DAPI_1_zstack.tif")
    assert len(written) == 6
    for z in range(3):
        for ci, name in ((1, "RFP"), (2, "GFP")):
            f = f"exp/Tp_1_1/Tp_1_1_zstack_{name}/Tp_1_1_zstack_z{z}_{name}.tif"
            assert f in written and np.array_equal(tiff_io.read_stack(f), z4[z, ci]), f
    z2 = rng.integers(0, 4000, (2, 2, 8, 10), dtype=np.uint16)
    src2 = "exp/B_RFP_GFP_2_zstack.tif"
    tiff_io.write_stack(src2, z2)
    written2 = split_zstack.process_tif(src2, [1, 2], move=False)  # 2-channel stacks are RFP, GFP whatever was asked
    assert os.path.exists(src2) and len(written2) == 4
    assert np.array_equal(tiff_io.read_stack("exp/B_2/B_2_zstack_GFP/B_2_zstack_z1_GFP.tif"), z2[1, 1])


@pytest.mark.gpu
def test_segment_tif_matches_oracle(tmp_path):
    import torch

    from oracle import pipeline as opipe
    from particle_col_image_segmentation_b200 import synth

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    planes = synth.zstack_u16(3, 96, 128, seed=9)
    stack = np.stack([planes // 2, planes], axis=1)  # (Z, C, Y, X), channel 1 is the one segmented
    p = str(tmp_path / "s_RFP_GFP_1_zstack.tif")
    tiff_io.write_stack(p, stack)
    t = tiff_io.read_stack_pinned(p)
    assert tuple(t.shape) == stack.shape and t.is_pinned() and np.array_equal(t.numpy(), stack)
    out = split_zstack.segment_tif(p, channel=1, chunk=2)
    want = opipe.segment_zstack(planes)
    assert np.array_equal(out["labels"].numpy(), want["labels"]) and np.array_equal(out["edt"].numpy(), want["edt"])
    assert np.array_equal(out["refined"].numpy().astype(bool), want["refined"]) and np.array_equal(out["table"].numpy(), want["table"])


def test_read_stack_pinned_without_a_gpu(tmp_path):
    """The pinned reader degrades to pageable memory when no CUDA device is present; the pixels are the same."""
    import torch

    rng = np.random.default_rng(6)
    stack = rng.integers(0, 65535, (2, 3, 9, 14), dtype=np.uint16)
    p = str(tmp_path / "p.tif")
    tiff_io.write_stack(p, stack)
    t = tiff_io.read_stack_pinned(p)
    assert t.dtype == torch.uint16 and tuple(t.shape) == stack.shape and np.array_equal(t.numpy(), stack)
    assert t.is_pinned() == torch.cuda.is_available()
    big = str(tmp_path / "be.tif")
    open(big, "wb").write(_hand_tiff([stack[0, 0]], bo=">"))
    with pytest.raises(tiff_io.TiffError, match="big-endian"):
        tiff_io.read_stack_pinned(big)
