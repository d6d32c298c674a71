"""The device-backed mirrors of the reference's L2 functions against (a) the outputs
of the reference's own functions frozen in tests/golden and (b) the oracle on fresh
seeded inputs.  Areas, centroids, bboxes, counts and images are bit-exact."""

import numpy as np
import pytest
import torch
from scipy import ndimage as ndi

from oracle import l2 as ol2
from oracle import nanosims as onano
from oracle import refine as orefine
from particle_col_image_segmentation_b200 import synth

from helpers import assert_summary_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ta():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from particle_col_image_segmentation_b200 import tiff_analysis

    return tiff_analysis


def test_single_file_path_matches_reference_golden(ta, golden):
    arrays, meta = golden
    types = {1: "3D05", 2: "Particle", 3: "Background"}
    den = ta.median_filter(arrays["A_raw"], size=ta.DENOISE_SIZE)
    assert np.array_equal(den, arrays["A_denoised"])
    res = ta.get_cell_positions_and_areas(den, types, merged=True)
    assert_summary_equal(ol2.summarize_positions(res), meta["A_positions"])
    cnt, dens, ratio = ta.get_cell_counts_and_densities(res[0], res[1], res[2])
    assert {k: int(v) for k, v in cnt.items()} == meta["A_counts"]["count"]
    assert dens == meta["A_counts"]["density"] and ratio == meta["A_counts"]["ratio"]
    rec, area = ta.recreate_particle_area(den, types, res[2])
    assert np.array_equal(rec, arrays["A_recreated"]) and float(area) == meta["A_particle_area"]
    _, images = ta.get_cell_clusters_from_distances(den, res[0], res[1], types)
    for k, im in images.items():
        assert im.dtype == bool and np.array_equal(im, arrays[f"A_merged_image_{k}"]), k
    out = ta.process_single_array(arrays["A_raw"][None], types)
    assert np.array_equal(out["recreated"], arrays["A_recreated"]) and float(out["particle_area"]) == meta["A_particle_area"]


def test_multiclass_matches_reference_golden(ta, golden):
    arrays, meta = golden
    res = ta.get_cell_positions_and_areas(arrays["B_image"], ta.BASE_TYPE_MAP, merged=True)
    assert_summary_equal(ol2.summarize_positions(res), meta["B_positions"])
    rec, area = ta.recreate_particle_area(arrays["B_image"], ta.BASE_TYPE_MAP, res[2])
    assert np.array_equal(rec, arrays["B_recreated"]) and float(area) == meta["B_particle_area"]


def test_channel_ops_match_reference_golden(ta, golden):
    arrays, meta = golden
    dapi, rfp = arrays["C_dapi"], arrays["C_rfp"]
    assert np.array_equal(ta.combine_cell_positions_and_clusters(dapi, rfp), arrays["C_dapi_updated"])
    base = ta.get_rfp_base_arr(rfp.copy(), ["3D05", "6B07"])
    assert np.array_equal(base, arrays["C_rfp_base"])
    comb = ta.combine_channels(base.copy(), {"RFP": rfp, "DAPI": dapi}, ["3D05", "6B07"])
    assert np.array_equal(comb, arrays["C_combined"])
    assert np.array_equal(ta.relabel_other_channel(rfp, "RFP"), arrays["C_other_updated"])
    up, n = ta.fill_particle_area(dapi, 2, 1, 2)
    assert np.array_equal(up, arrays["C_fill"]) and int(n) == meta["C_fill_count"]


@pytest.mark.parametrize("seed,size", [(21, 256), (22, 300), (23, 512)])
def test_l2_matches_oracle_fresh_inputs(ta, seed, size):
    raw = synth.class_image(size, size + 37, seed=seed, noise=0.03)
    den = ta.median_filter(raw, size=5)
    assert np.array_equal(den, ndi.median_filter(raw, size=5))
    types = {1: "C3M10", 2: "Particle", 3: "Background"}
    got = ta.get_cell_positions_and_areas(den, types, merged=True)
    want = ol2.get_cell_positions_and_areas(den, types, merged=True)
    assert_summary_equal(ol2.summarize_positions(got), ol2.summarize_positions(want))
    g, ng = ta.recreate_particle_area(den, types, got[2])
    w, nw = ol2.recreate_particle_area(den, types, want[2])
    assert np.array_equal(g, w) and ng == nw
    other = synth.class_image(size, size + 37, seed=seed + 50, noise=0.0)
    assert np.array_equal(ta.combine_cell_positions_and_clusters(den, other), ol2.combine_cell_positions_and_clusters(den, other))
    # get_merged_regions at the public boundary (numpy mask + region list)
    regs = got[0].get("C3M10", []) + got[1].get("C3M10", [])
    mg, img = ta.get_merged_regions(den == 1, regs)
    oregs = want[0].get("C3M10", []) + want[1].get("C3M10", [])
    mo, imo = ol2.get_merged_regions(den == 1, oregs)
    assert np.array_equal(img, imo) and len(mg) == len(mo)
    for a, b in zip(mg, mo):
        assert a["area"] == b["area"] and a["bbox"] == b["bbox"] and np.array_equal(a["centroid"], b["centroid"])


def test_fill_particle_area_edge_cases(ta):
    img = np.full((40, 50), 3, np.uint8)  # no particle at all: scipy's EDT quirk decides
    img[0:3, 0:4] = 1
    img[20:25, 20:25] = 1
    g, ng = ta.fill_particle_area(img, 2, 1, 2)
    w, nw = ol2.fill_particle_area(img, 2, 1, 2)
    assert np.array_equal(g, w) and ng == nw
    img[10:12, 30:33] = 2
    g, ng = ta.fill_particle_area(img, 2, 1, 2)
    w, nw = ol2.fill_particle_area(img, 2, 1, 2)
    assert np.array_equal(g, w) and ng == nw


def test_normalize_ds_arr(ta):
    a = np.zeros((8, 9), np.uint8)
    assert ta.normalize_ds_arr(a[None]).shape == (8, 9) and ta.normalize_ds_arr(a[..., None]).shape == (8, 9)
    assert ta.normalize_ds_arr(a) is a
    with pytest.raises(ValueError):
        ta.normalize_ds_arr(np.zeros((2, 3, 4)))


@pytest.mark.parametrize("size", [160, 400])
def test_refine_boundaries(ta, size):
    from particle_col_image_segmentation_b200 import refine_boundaries as rb

    _, prob = synth.touching_particles(size, size + 11, seed=size)
    got = rb.refine_boundaries(prob)
    want = orefine.refine_boundaries(prob)
    for k in ("binary_mask", "distance", "local_max", "markers"):
        assert got[k].dtype == want[k].dtype, k
        assert np.array_equal(got[k], want[k]), k


@pytest.mark.parametrize("k", [5, 7])
def test_nanosims(ta, k):
    from particle_col_image_segmentation_b200 import nanosims

    planes, roi, set_id, agg = synth.nanosims_stack(256, k, 120, seed=1004)
    red = np.isin(roi, np.nonzero(set_id == 1)[0] + 1)
    green = np.isin(roi, np.nonzero(set_id == 2)[0] + 1)
    got = nanosims.analyse(planes, red, green, agg)
    want = onano.analyse(planes, red, green, agg)
    assert got.shape == want.shape
    assert np.array_equal(got, want), np.argwhere(got != want)[:5]
    c, t, m = nanosims.activity_vs_distance(got[:, 2 + k], got[:, -1], np.linspace(0, 5, 11))
    c2, t2, m2 = onano.activity_vs_distance(want[:, 2 + k], want[:, -1], np.linspace(0, 5, 11))
    assert np.array_equal(c, c2) and np.array_equal(t, t2)
    assert np.array_equal(m, m2, equal_nan=True)  # bin means (empty bins are NaN on both sides); north_star tolerance 1e-5, met bit-exactly


def test_cell_cell_distances_match_brute_force(ta):
    """SURVEY 8f row 4 (refine_boundaries.py:8-12 goal 3): nearest same-strain / other-strain neighbour."""
    img = synth.multi_class_image(384, 384, seed=21)
    cell_pos, _, _, _ = ta.get_cell_positions_and_areas(img, ta.BASE_TYPE_MAP)
    cell_pos = dict(cell_pos)
    cell_pos["ghost"] = []  # a strain without cells
    got = ta.get_cell_cell_distances(cell_pos)
    want = ol2.get_cell_cell_distances(cell_pos)
    assert set(got) == set(want) and any(len(v[0]) > 3 for v in got.values())
    for k in want:
        assert np.array_equal(got[k][0], want[k][0]), k  # bit-exact distances
        assert np.array_equal(got[k][1], want[k][1]), k
    # a strain with a single cell has no same-strain neighbour
    from particle_col_image_segmentation_b200 import ops

    one = torch.tensor([[3.0, 4.0]], dtype=torch.float64, device="cuda")
    d, j = ops.nearest(one, one, exclude_self=True)
    assert np.isinf(d.item()) and j.item() == -1
    pts = torch.rand((1000, 2), dtype=torch.float64, device="cuda") * 500
    d, j = ops.nearest(pts, pts, exclude_self=True)
    p = pts.cpu().numpy()
    dx = p[:, None, 0] - p[None, :, 0]
    dy = p[:, None, 1] - p[None, :, 1]
    d2 = dx * dx + dy * dy
    np.fill_diagonal(d2, np.inf)
    assert np.array_equal(j.cpu().numpy(), d2.argmin(1)) and np.array_equal(d.cpu().numpy(), np.sqrt(d2.min(1)))


def test_nanosims_ratio_images():
    """SURVEY 8f row 3 (.m:17-69): imgaussfilt + isotope ratio + uint8 scaling.  The device kernels and
    oracle/nanosims.py do the same IEEE operations in the same order: bit-exact.  Against an independent
    Gaussian (scipy correlate1d, mode='nearest') the filter agrees to 1e-13 relative (different summation
    order); MATLAB itself cannot be run here (parity with MATLAB unpinned, north_star tolerance 1e-5)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from particle_col_image_segmentation_b200 import nanosims

    planes, _, _, _ = synth.nanosims_stack(n=130, k=7, n_rois=30, seed=8)
    rng = np.random.default_rng(2)
    ions = {name: planes[i] for i, name in enumerate(onano.PLANES_7)}
    ions["Esi"] = rng.poisson(40, planes[0].shape).astype(np.float64)
    ions["12C"][5, 7] = 0.0
    ions["13C"][5, 7] = 0.0  # 0 / 0 -> NaN in the raw ratio image, uint8(NaN) = 0
    for sigma in (1, 1.5, 0.6, 3.2):
        got = nanosims.imgaussfilt(ions["17O"], sigma)
        want = onano.imgaussfilt(ions["17O"], sigma)
        assert got.dtype == np.float64 and np.array_equal(got, want), sigma
        r = int(np.ceil(2 * sigma))
        x = np.arange(-r, r + 1.0)
        w = np.exp(-x * x / (2 * sigma * sigma))
        w /= w.sum()
        ref = ndi.correlate1d(ndi.correlate1d(ions["17O"], w, axis=0, mode="nearest"), w, axis=1, mode="nearest")
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=0)
    got = nanosims.ratio_images(ions)
    want = onano.ratio_images(ions)
    assert set(got) == set(want) and len(got) == 17
    for k in want:
        assert got[k].dtype == np.uint8 and got[k].shape == (128, 128)
        assert np.array_equal(got[k], want[k]), (k, np.argwhere(got[k] != want[k])[:4])
        assert got[k].max() == 255 or k == "N14C12ESIratio"
    assert got["C13ratimg"][4, 6] == 0  # the 0/0 pixel (after the one-pixel crop)
    # rounding is half away from zero and saturating
    x = np.array([[0.0, 0.5, 1.5, 2.5, 254.5, 255.0, 127.49999999999999]])
    assert nanosims.scaled_uint8(x).tolist() == [[0, 1, 2, 3, 255, 255, 127]]
    assert np.array_equal(nanosims.scaled_uint8(x), onano.scaled_uint8(x))


def test_nanosims_imresize_and_resized_roi_sums():
    """.m:125, :189: ``imresize(holder, [n n])`` (bicubic, antialiased) and the sums under the resized ROI masks.  The
    device resize runs the restatement's tap tables in the same order without FMA: bit-exact.  The per-ROI sums use
    the adjoint resize of the ion planes (one resize per plane instead of one per ROI): same numbers up to summation
    order, compared at 1e-12 (north_star tolerance for activities: 1e-5).  MATLAB itself cannot be run here."""
    from particle_col_image_segmentation_b200 import nanosims

    rng = np.random.default_rng(11)
    a = rng.random((57, 83))
    for shape in ((57, 83), (100, 120), (31, 40), (90, 30), (20, 200)):
        got, want = nanosims.imresize(a, shape), onano.imresize(a, shape)
        assert got.shape == tuple(shape) and np.array_equal(got, want), shape
    # ROI image 96 x 96 painted on a 64 x 64 acquisition (and the other way round)
    for n_roi_img, n_acq in ((96, 64), (48, 64)):
        planes, _, _, _ = synth.nanosims_stack(n_acq, 5, 20, seed=7)
        _, roi, set_id, agg = synth.nanosims_stack(n_roi_img, 5, 20, seed=9)
        red = np.isin(roi, np.nonzero(set_id == 1)[0] + 1)
        green = np.isin(roi, np.nonzero(set_id == 2)[0] + 1)
        got = nanosims.analyse(planes, red, green, agg)
        want = onano.analyse(planes, red, green, agg)
        assert got.shape == want.shape and got.shape[0] > 4
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-9, equal_nan=True)


def test_h5_entry_points(tmp_path, ta):
    """File -> result through the ilastik export reader: same answers as the array entry points."""
    from particle_col_image_segmentation_b200 import h5_io, refine_boundaries as rb

    img = synth.class_image(256, 256, seed=12)
    p = h5_io.write_dataset(str(tmp_path / "seg.h5"), img[:, :, None], chunks=(64, 64, 1), compression="gzip")
    types = {1: "C3M10", 2: "Particle", 3: "Background"}
    got = ta.process_single_h5_file(p, types)
    want = ta.process_single_array(img, types)
    assert np.array_equal(got["recreated"], want["recreated"]) and got["cell_count"] == want["cell_count"] and got["particle_area"] == want["particle_area"]
    _, prob = synth.touching_particles(160, 192, seed=4, pitch=32.0)
    stack = np.stack([prob * 0, prob * 0, 1 - prob, prob]).astype(np.float32)
    q = h5_io.write_dataset(str(tmp_path / "prob.h5"), stack, chunks=(1, 64, 64), compression="gzip", shuffle=True)
    a, b = rb.refine_boundaries_h5(q), rb.refine_boundaries(prob)
    for k in ("binary_mask", "distance", "local_max", "markers"):
        assert np.array_equal(a[k], b[k]), k
