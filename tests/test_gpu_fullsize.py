"""BASELINE.json configs at their full sizes.  Where the oracle is too slow for the whole
input, size-independent properties are checked on the full result and the oracle on a sample."""

import numpy as np
import pytest
import torch
from scipy import ndimage as ndi

from oracle import pipeline as opipe
from oracle import refine as orefine
from oracle.skimage_shim import morphology as omorph
from particle_col_image_segmentation_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from particle_col_image_segmentation_b200 import morphology, ndimage, ops, refine_boundaries, split_zstack

    class NS:
        pass

    ns = NS()
    ns.morphology, ns.ndimage, ns.ops, ns.rb, ns.seg = morphology, ndimage, ops, refine_boundaries, split_zstack
    return ns


def test_config2_zstack_2048x2048x64(mods):
    """configs[1]: split_zstack + segment a synthetic 2048x2048x64 uint16 z-stack on one B200."""
    Z, S = 64, 2048
    stack = synth.zstack_u16_device(Z, S, S, seed=1002, device=torch.device("cuda"))
    res = mods.seg.SegmentPlan(stack, chunk=32)()
    table = res.table_device()
    torch.cuda.synchronize()
    counts = res.counts.cpu().numpy()
    # properties that hold for any input
    assert int(table.shape[0]) == int(counts.sum())
    assert torch.equal(res.labels.amax(dim=(1, 2)).cpu(), res.counts.cpu())  # labels are 1..N, consecutive
    assert torch.equal((res.labels != 0), res.mask.bool())                   # labelled pixels == mask
    assert bool(((res.edt > 0) == res.refined.bool()).all())                  # EDT support == refined mask
    assert float(table[:, 2].sum()) == float(res.mask.sum())                  # areas add up to the mask
    # the oracle on three slices of this very stack: everything bit-exact
    for zi in (0, 31, 63):
        want = opipe.segment_slice(stack[zi].cpu().numpy(), z=zi)
        assert int(res.threshold[zi]) == want["threshold"]
        assert np.array_equal(res.mask[zi].cpu().numpy().astype(bool), want["mask"])
        assert np.array_equal(res.labels[zi].cpu().numpy(), want["labels"])
        assert np.array_equal(res.refined[zi].cpu().numpy().astype(bool), want["refined"])
        assert np.array_equal(res.edt[zi].cpu().numpy(), want["edt"])
        assert np.array_equal(table[table[:, 0] == zi].cpu().numpy(), want["table"])
    # the graph replay and a different chunking give the same bytes
    res2 = mods.seg.SegmentPlan(stack, chunk=64, graph=True)()
    torch.cuda.synchronize()
    for k in ("mask", "labels", "refined", "edt", "threshold", "counts"):
        assert torch.equal(getattr(res, k), getattr(res2, k)), k
    assert torch.equal(table, res2.table_device())


def test_config5_shard_2048x2048x256(mods):
    """configs[4] / north_star target: one 2048x2048x256 uint16 stack (a GPU's share of the 64-stack batch),
    processed in chunks of 64 slices.  Size-independent properties on the whole result, the oracle on two
    slices, and the sharding identity: slices [128, 192) run as a separate shard give the same bytes."""
    Z, S = 256, 2048
    stack = synth.zstack_u16_device(Z, S, S, seed=1005, device=torch.device("cuda"))
    res = mods.seg.SegmentPlan(stack, chunk=64, graph=True)()
    table = res.table_device()
    torch.cuda.synchronize()
    assert int(table.shape[0]) == int(res.counts.sum())
    assert torch.equal(res.labels.amax(dim=(1, 2)).cpu(), res.counts.cpu())
    assert torch.equal((res.labels != 0), res.mask.bool())
    assert bool(((res.edt > 0) == res.refined.bool()).all())
    assert float(table[:, 2].sum()) == float(res.mask.sum())
    assert bool((table[1:, 0] >= table[:-1, 0]).all())  # rows ordered by slice, then by label
    for zi in (5, 250):
        want = opipe.segment_slice(stack[zi].cpu().numpy(), z=zi)
        assert np.array_equal(res.labels[zi].cpu().numpy(), want["labels"])
        assert np.array_equal(res.edt[zi].cpu().numpy(), want["edt"])
        assert np.array_equal(table[table[:, 0] == zi].cpu().numpy(), want["table"])
    shard = mods.seg.SegmentPlan(stack[128:192], chunk=16, z0=128)()
    st = shard.table_device()
    torch.cuda.synchronize()
    for k in ("mask", "labels", "refined", "edt", "threshold", "counts"):
        assert torch.equal(getattr(shard, k), getattr(res, k)[128:192]), k
    assert torch.equal(st, table[(table[:, 0] >= 128) & (table[:, 0] < 192)])


def test_config3_touching_particles_4096(mods):
    """configs[2]: refine_boundaries morphology + EDT on 4096x4096 masks with ~10k touching particles."""
    mask, prob = synth.touching_particles(4096, 4096, seed=1003)
    # EDT (both polarities: inside the particles, and distance to them) and disk dilations, full size, vs scipy
    assert np.array_equal(mods.ndimage.distance_transform_edt(mask), ndi.distance_transform_edt(mask))
    assert np.array_equal(mods.ndimage.distance_transform_edt(~mask), ndi.distance_transform_edt(~mask))
    d2 = np.rint(ndi.distance_transform_edt(~mask) ** 2).astype(np.int64)
    assert np.array_equal(mods.morphology.binary_dilation(mask, omorph.disk(20)), d2 <= 400)   # tiff_analysis.py:990
    assert np.array_equal(mods.morphology.binary_dilation(mask, omorph.disk(2)), ndi.binary_dilation(mask, omorph.disk(2)))
    assert np.array_equal(mods.ndimage.binary_fill_holes(mask), ndi.binary_fill_holes(mask))
    opened = mods.morphology.binary_opening(mask, omorph.disk(2))
    assert np.array_equal(opened, omorph.binary_opening(mask, omorph.disk(2)))
    # the refine_boundaries chain on the probability map
    got = mods.rb.refine_boundaries(prob)
    binary_mask = prob < 0.5
    assert np.array_equal(got["binary_mask"], binary_mask)
    assert np.array_equal(got["distance"], ndi.distance_transform_edt(binary_mask))
    lab, n = ndi.label(binary_mask, structure=np.ones((3, 3)))
    assert n > 5000  # ~10k particles separated by their boundaries
    # markers: labelled local maxima; properties at full size + the oracle on a 1024x1024 corner
    assert got["markers"].max() >= n * 0.5 and not (got["local_max"] & ~binary_mask).any()
    assert np.array_equal(got["markers"] > 0, got["local_max"])
    want = orefine.refine_boundaries(prob[:1024, :1024])
    crop = mods.rb.refine_boundaries(np.ascontiguousarray(prob[:1024, :1024]))
    for k in ("binary_mask", "distance", "local_max", "markers"):
        assert np.array_equal(crop[k], want[k]), k
    # the watershed (refine_boundaries.py:73) at full size: a valid flood of the whole mask from the markers
    from particle_col_image_segmentation_b200 import segmentation

    labels, sweeps = segmentation.watershed(prob, got["markers"], mask=binary_mask, return_sweeps=True)
    assert labels.dtype == np.int32 and sweeps >= 1
    assert np.array_equal(labels[got["markers"] > 0], got["markers"][got["markers"] > 0])
    assert (labels[~binary_mask] == 0).all()
    comp_has_marker = np.zeros(n + 1, bool)
    lab4, n4 = ndi.label(binary_mask)  # 4-connected: the flood's connectivity
    comp_has_marker = np.zeros(n4 + 1, bool)
    comp_has_marker[np.unique(lab4[got["markers"] > 0])] = True
    assert np.array_equal(labels != 0, comp_has_marker[lab4] & binary_mask)  # flooded = mask components holding a marker
    same = np.zeros(labels.shape, bool)
    same[1:] |= labels[1:] == labels[:-1]
    same[:-1] |= labels[:-1] == labels[1:]
    same[:, 1:] |= labels[:, 1:] == labels[:, :-1]
    same[:, :-1] |= labels[:, :-1] == labels[:, 1:]
    assert (same | (got["markers"] > 0) | (labels == 0)).all()  # every flooded pixel hangs on a neighbour of its label


def test_class_image_2048_single_file_path(mods):
    """The reference's native size (tiff_analysis.py:734): denoise + measure + recreate on a 2048^2 class image."""
    from oracle import l2 as ol2
    from particle_col_image_segmentation_b200 import tiff_analysis as ta

    from helpers import assert_summary_equal

    raw = synth.class_image(2048, 2048, seed=1234)
    types = {1: "3D05", 2: "Particle", 3: "Background"}
    out = ta.process_single_array(raw, types)
    den = ndi.median_filter(raw, size=5)
    assert np.array_equal(out["denoised"], den)
    want = ol2.get_cell_positions_and_areas(den, types, merged=True)
    got = ta.get_cell_positions_and_areas(den, types, merged=True)
    assert_summary_equal(ol2.summarize_positions(got), ol2.summarize_positions(want))
    assert out["cell_count"] == ol2.get_cell_counts_and_densities(want[0], want[1], want[2])[0]
    rec, area = ol2.recreate_particle_area(den, types, want[2])
    assert np.array_equal(out["recreated"], rec) and float(out["particle_area"]) == float(area)
