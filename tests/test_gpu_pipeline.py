"""Full segment pipeline (threshold -> median -> label -> regionprops -> refine -> EDT)
against ``oracle.pipeline`` on the same seeded stacks.  Everything is bit-exact."""

import numpy as np
import pytest
import torch

from oracle import pipeline as opipe
from particle_col_image_segmentation_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def seg():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from particle_col_image_segmentation_b200 import split_zstack

    return split_zstack


def check(got, want):
    assert np.array_equal(got["threshold"], want["threshold"]), (got["threshold"], want["threshold"])
    for k in ("mask", "labels", "refined", "edt", "counts"):
        assert got[k].dtype == want[k].dtype, (k, got[k].dtype, want[k].dtype)
        if not np.array_equal(got[k], want[k]):
            bad = np.argwhere(got[k] != want[k])
            raise AssertionError(f"{k}: {len(bad)} differ, first {bad[0].tolist()} got {got[k][tuple(bad[0])]} want {want[k][tuple(bad[0])]}")
    assert got["table"].shape == want["table"].shape
    assert np.array_equal(got["table"], want["table"]), np.argwhere(got["table"] != want["table"])[:5]


@pytest.mark.parametrize("shape,chunk", [((3, 96, 128), 2), ((5, 130, 257), 16), ((2, 512, 512), 1)])
def test_pipeline_matches_oracle(seg, shape, chunk):
    stack = synth.zstack_u16(*shape, seed=shape[1])
    check(seg.segment_zstack(stack, chunk=chunk), opipe.segment_zstack(stack))


def test_pipeline_noisy_and_params(seg):
    rng = np.random.default_rng(3)
    stack = synth.zstack_u16(2, 200, 200, seed=11)
    noise = rng.random(stack.shape) < 0.02
    stack = np.where(noise, rng.integers(0, 30000, stack.shape), stack).astype(np.uint16)
    for dn, ms in ((5, 20), (3, 1), (0, 50)):
        check(seg.segment_zstack(stack, denoise_size=dn, min_size=ms), opipe.segment_zstack(stack, denoise_size=dn, min_size=ms))


def test_pipeline_degenerate(seg):
    flat = np.full((2, 64, 96), 700, np.uint16)  # single-valued: Otsu returns the value, empty mask
    check(seg.segment_zstack(flat), opipe.segment_zstack(flat))
    single = synth.slice_u16(128, 160, seed=5)
    got = seg.segment_zstack(single)
    want = opipe.segment_slice(single)
    assert int(got["threshold"]) == want["threshold"]
    for k in ("mask", "labels", "refined", "edt"):
        assert np.array_equal(got[k], want[k]), k


def test_config1_slice(seg):
    """BASELINE.json configs[0]: one synthetic 512x512 uint16 slice."""
    img = synth.slice_u16(512, 512, seed=1001)
    got = seg.segment_zstack(img[None])
    check(got, opipe.segment_zstack(img[None]))
    assert got["counts"][0] > 10


def test_split_channels_layout(seg):
    z = np.arange(2 * 4 * 3 * 5, dtype=np.uint16).reshape(2, 4, 3, 5)
    planes = seg.split_channels(z, (1, 2))
    assert list(planes) == ["RFP", "GFP"] and np.array_equal(planes["GFP"], z[:, 2])
    z2 = z[:, :2]
    planes2 = seg.split_channels(z2, (1, 2))
    assert list(planes2) == ["RFP", "GFP"] and np.array_equal(planes2["RFP"], z2[:, 0])
    assert seg.plane_name("img", 3, "GFP") == "img_z3_GFP.tif"


def test_segment_plan_graph_replay_matches_eager(seg):
    """The captured CUDA graph replays the same kernels: outputs must be identical, repeatedly."""
    stack = synth.zstack_u16(6, 160, 288, seed=77)
    d = torch.from_numpy(stack).cuda()
    eager = seg.SegmentPlan(d, chunk=4)()
    want = eager.to_numpy()
    plan = seg.SegmentPlan(d, chunk=4, graph=True)
    for _ in range(3):
        got = plan().to_numpy()
        for k in ("threshold", "mask", "labels", "refined", "edt", "table", "counts"):
            assert np.array_equal(got[k], want[k]), k
    check(want, opipe.segment_zstack(stack))


def test_pinned_host_path(seg):
    stack = synth.zstack_u16(3, 128, 160, seed=31)
    host_in = torch.from_numpy(stack).pin_memory()
    host_out = seg.alloc_host_outputs(*stack.shape)
    for _ in range(2):
        n = seg.segment_zstack_pinned(host_in, host_out, chunk=2)
    want = opipe.segment_zstack(stack)
    assert n == len(want["table"]) and np.array_equal(host_out["table"].numpy(), want["table"])
    assert np.array_equal(host_out["labels"].numpy(), want["labels"]) and np.array_equal(host_out["edt"].numpy(), want["edt"])
    assert np.array_equal(host_out["mask"].numpy().astype(bool), want["mask"]) and np.array_equal(host_out["refined"].numpy().astype(bool), want["refined"])
    # a caller that only needs the table and the labels: nothing else crosses PCIe, the answers are the same
    sel = seg.alloc_host_outputs(*stack.shape, outputs=("labels",))
    assert set(sel) == {"labels", "threshold", "counts"}
    n2 = seg.segment_zstack_pinned(host_in, sel, chunk=2, outputs=("labels",))
    assert n2 == n and np.array_equal(sel["table"].numpy(), want["table"]) and np.array_equal(sel["labels"].numpy(), want["labels"])
    assert np.array_equal(sel["threshold"].numpy(), want["threshold"])
    with pytest.raises(ValueError):
        seg.segment_zstack_pinned(host_in, sel, outputs=("edt",))  # no buffer for it


def test_fill_holes_from_table_matches_scipy(seg):
    """The bounding-box restricted hole filling the pipeline uses == scipy.ndimage.binary_fill_holes."""
    from scipy import ndimage as ndi

    from particle_col_image_segmentation_b200 import _lib, ops

    rng = np.random.default_rng(5)
    for shape, p in (((3, 70, 97), 0.55), ((2, 130, 257), 0.45), ((1, 64, 64), 0.8), ((2, 33, 65), 0.2)):
        m = rng.random(shape) < p
        m[0, 5:25, 5:30] = True
        m[0, 10:20, 10:25] = False  # a hole with islands inside
        m[0, 13:16, 14:18] = True
        B, H, W = shape
        bits = ops.pack(torch.from_numpy(m).cuda())
        labels, counts, offsets = ops.label_bits(bits, W, connectivity=8)
        table = ops.new_table(max(1, int(offsets[-1])), bits.device)
        ops.region_table(labels, offsets, table, fg_bits=bits)
        lib = _lib.load()
        n = lib.pcs_fill_holes_table_workspace_bytes(B, H, W)
        ws = torch.empty(n, dtype=torch.uint8, device=bits.device)
        out = torch.empty_like(bits)
        mask = torch.empty((B, H, W), dtype=torch.uint8, device=bits.device)
        _lib.check(lib.pcs_fill_holes_table_bits(bits.data_ptr(), table.data_ptr(), table.shape[1], offsets.data_ptr(), 1, out.data_ptr(), mask.data_ptr(), B, H, W, ws.data_ptr(), n, ops._stream()))
        got = ops.unpack(out, W, torch.bool).cpu().numpy()
        for i in range(B):
            assert np.array_equal(got[i], ndi.binary_fill_holes(m[i])), (shape, i)
        assert np.array_equal(mask.cpu().numpy().astype(bool), got)


def _spiral(n):
    """Square spiral wall, one pixel wide, lanes one pixel wide: concave everywhere, no hole."""
    m = np.zeros((n, n), bool)
    y = x = 0
    dy, dx = 0, 1
    seg = n - 1
    m[0, 0] = True
    turns = 0
    while seg > 0:
        for _ in range(seg):
            y += dy
            x += dx
            m[y, x] = True
        dy, dx = dx, -dy
        turns += 1
        if turns >= 3 and turns % 2 == 1:
            seg -= 2
    return m


@pytest.mark.parametrize("min_size", [1, 6, 40])
def test_refine_labeled_matches_scipy(seg, min_size):
    """Row-gap candidates + seeds == remove_small_objects then scipy.ndimage.binary_fill_holes,
    on shapes chosen to stress the candidate argument (spirals, C shapes, nested rings, islands in
    holes, holes on the first / last rows and columns, wide gaps across many words)."""
    from scipy import ndimage as ndi

    from oracle.skimage_shim.morphology import remove_small_objects
    from particle_col_image_segmentation_b200 import ops

    rng = np.random.default_rng(17 + min_size)
    cases = []
    for shape, p in (((3, 70, 97), 0.55), ((2, 130, 257), 0.45), ((1, 64, 64), 0.8), ((2, 33, 65), 0.2), ((2, 50, 1100), 0.5), ((1, 200, 40), 0.6)):
        cases.append(rng.random(shape) < p)
    hand = np.zeros((4, 96, 200), bool)
    hand[0, :41, :41] = _spiral(41)
    hand[0, 50:90, 10:190] = True  # ring 150 px wide: the gap spans several words
    hand[0, 55:85, 15:185] = False
    hand[0, 60:80, 60:140] = True  # nested ring inside the hole
    hand[0, 64:76, 64:136] = False
    hand[0, 68:72, 90:100] = True  # island in the inner hole
    hand[0, 69, 92] = False        # one-pixel hole in the island
    hand[1, 0:20, 5:30] = True     # C open to the top border row
    hand[1, 0:15, 10:25] = False
    hand[1, 76:96, 5:30] = True    # closed by the bottom row? no: open to the bottom border
    hand[1, 81:96, 10:25] = False
    hand[1, 30:60, 0:25] = True    # ring touching the left border: still a hole
    hand[1, 35:55, 1:20] = False
    hand[1, 30:60, 170:200] = True  # C open to the right
    hand[1, 35:55, 175:200] = False
    hand[1, 30:60, 90:120] = True   # hole closed only diagonally (8-connected wall): still a hole
    hand[1, 35:55, 95:115] = False
    hand[1, 30, 90] = False
    hand[2] = rng.random((96, 200)) < 0.62
    hand[3, 10:80, 20:180] = True
    hand[3, 20:70, 30:170] = rng.random((50, 140)) < 0.5  # dense debris inside a big hole
    cases.append(hand)
    for m in cases:
        B, H, W = m.shape
        bits = ops.pack(torch.from_numpy(m).cuda())
        labels, counts, offsets = ops.label_bits(bits, W, connectivity=8)
        table = ops.new_table(max(1, int(offsets[-1])), bits.device)
        ops.region_table(labels, offsets, table, fg_bits=bits)
        out, mask = ops.refine_labeled(bits, labels, table, offsets, min_size, W, want_mask=True)
        got = ops.unpack(out, W, torch.bool).cpu().numpy()
        for i in range(B):
            want = ndi.binary_fill_holes(remove_small_objects(m[i], min_size=min_size, connectivity=2))
            assert np.array_equal(got[i], want), (m.shape, i, np.argwhere(got[i] != want)[:5])
        assert np.array_equal(mask.cpu().numpy().astype(bool), got)


def test_first_call_on_two_streams_in_a_fresh_process(seg):
    """The EDT's sqrt table used to be filled lazily on the caller's stream: a first call that spread its chunks over two
    streams (or ran inside a graph capture) could read the table before it was written.  The table is a compile-time
    constant now; this runs exactly that first call in a fresh process and compares every output with the oracle."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np, torch\n"
        "from oracle import pipeline as opipe\n"
        "from particle_col_image_segmentation_b200 import split_zstack, synth\n"
        "stack = synth.zstack_u16(4, 96, 160, seed=19)\n"
        "got = split_zstack.SegmentPlan(torch.from_numpy(stack).cuda(), chunk=1, streams=2, graph=False)().to_numpy()\n"
        "want = opipe.segment_zstack(stack)\n"
        "assert all(np.array_equal(got[k], want[k]) for k in ('threshold', 'mask', 'labels', 'refined', 'edt', 'table')), 'first call differs'\n"
        "print('first call ok')\n"
    ) % root
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "first call ok" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]
