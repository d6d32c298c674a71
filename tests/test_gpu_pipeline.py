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
