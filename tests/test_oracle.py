"""The oracle against the reference's frozen outputs, the live reference (when
present) and brute-force definitions.  CPU only."""

import numpy as np
import pytest
from scipy import ndimage as ndi

from oracle import l2, nanosims, pipeline, ref_loader, refine
from oracle.skimage_shim import filters, measure, morphology
from particle_col_image_segmentation_b200 import synth

from helpers import assert_summary_equal


# ---------------------------------------------------------------- golden (reference outputs)
def test_l2_positions_match_golden(golden):
    arrays, meta = golden
    den = ndi.median_filter(arrays["A_raw"], size=5)
    assert np.array_equal(den, arrays["A_denoised"])
    types = {1: "3D05", 2: "Particle", 3: "Background"}
    res = l2.get_cell_positions_and_areas(den, types, merged=True)
    assert_summary_equal(l2.summarize_positions(res), meta["A_positions"])
    cnt, dens, ratio = l2.get_cell_counts_and_densities(res[0], res[1], res[2])
    assert {k: int(v) for k, v in cnt.items()} == meta["A_counts"]["count"]
    assert dens == meta["A_counts"]["density"] and ratio == meta["A_counts"]["ratio"]
    rec, area = l2.recreate_particle_area(den, types, res[2])
    assert np.array_equal(rec, arrays["A_recreated"]) and float(area) == meta["A_particle_area"]
    _, images = l2.get_cell_clusters_from_distances(den, res[0], res[1], types)
    for k, im in images.items():
        assert np.array_equal(im, arrays[f"A_merged_image_{k}"]), k


def test_l2_multiclass_match_golden(golden):
    arrays, meta = golden
    res = l2.get_cell_positions_and_areas(arrays["B_image"], l2.BASE_TYPE_MAP, merged=True)
    assert_summary_equal(l2.summarize_positions(res), meta["B_positions"])
    rec, area = l2.recreate_particle_area(arrays["B_image"], l2.BASE_TYPE_MAP, res[2])
    assert np.array_equal(rec, arrays["B_recreated"]) and float(area) == meta["B_particle_area"]


def test_l2_channel_ops_match_golden(golden):
    arrays, meta = golden
    dapi, rfp = arrays["C_dapi"], arrays["C_rfp"]
    assert np.array_equal(l2.combine_cell_positions_and_clusters(dapi, rfp), arrays["C_dapi_updated"])
    base = l2.get_rfp_base_arr(rfp.copy(), ["3D05", "6B07"])
    assert np.array_equal(base, arrays["C_rfp_base"])
    comb = l2.combine_channels(base.copy(), {"RFP": rfp, "DAPI": dapi}, ["3D05", "6B07"])
    assert np.array_equal(comb, arrays["C_combined"])
    assert np.array_equal(l2.relabel_other_channel(rfp, "RFP"), arrays["C_other_updated"])
    up, n = l2.fill_particle_area(dapi, 2, 1, 2)
    assert np.array_equal(up, arrays["C_fill"]) and int(n) == meta["C_fill_count"]


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_l2_matches_live_reference():
    ta = ref_loader.load_tiff_analysis()
    for seed in (11, 12):
        img = ndi.median_filter(synth.class_image(256, 256, seed=seed, noise=0.03), size=5)
        types = {1: "6B07", 2: "Particle", 3: "Background"}
        a = l2.summarize_positions(ta.get_cell_positions_and_areas(img, types, merged=True))
        b = l2.summarize_positions(l2.get_cell_positions_and_areas(img, types, merged=True))
        assert_summary_equal(b, a)
        ra, na = ta.fill_particle_area(img, 2, 1, 2)
        rb, nb = l2.fill_particle_area(img, 2, 1, 2)
        assert np.array_equal(ra, rb) and na == nb
        other = synth.class_image(256, 256, seed=seed + 100, noise=0.0)
        assert np.array_equal(ta.combine_cell_positions_and_clusters(img, other), l2.combine_cell_positions_and_clusters(img, other))


# ---------------------------------------------------------------- shim primitives vs brute force
def _brute_label(img, conn8=True):
    h, w = img.shape
    out = np.zeros((h, w), dtype=np.int64)
    n = 0
    nb = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)] if conn8 else [(-1, 0), (0, -1), (0, 1), (1, 0)]
    for y in range(h):
        for x in range(w):
            if img[y, x] == 0 or out[y, x]:
                continue
            n += 1
            out[y, x] = n
            stack = [(y, x)]
            while stack:
                cy, cx = stack.pop()
                for dy, dx in nb:
                    yy, xx = cy + dy, cx + dx
                    if 0 <= yy < h and 0 <= xx < w and not out[yy, xx] and img[yy, xx] == img[y, x]:
                        out[yy, xx] = n
                        stack.append((yy, xx))
    return out


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_label_multivalued_vs_brute(seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 4, (23, 31)).astype(np.uint8)
    got = measure.label(img)
    assert got.dtype == np.int64
    assert np.array_equal(got, _brute_label(img))
    b = img > 1
    lb = measure.label(b)
    assert lb.dtype == np.int32 and np.array_equal(lb, _brute_label(b))
    assert np.array_equal(measure.label(b, connectivity=1), _brute_label(b, conn8=False))


def test_label_u_shape_raster_order():
    img = np.zeros((6, 9), dtype=np.uint8)
    img[1:5, 1] = 1
    img[1:5, 7] = 1
    img[4, 1:8] = 1  # arms join late in raster order
    img[0, 4] = 1
    lab = measure.label(img)
    assert lab[0, 4] == 1 and lab[1, 1] == 2 and lab[1, 7] == 2 and lab.max() == 2


def test_regionprops_fields():
    rng = np.random.default_rng(3)
    img = (rng.random((40, 50)) < 0.3).astype(np.uint8) * rng.integers(1, 3, (40, 50)).astype(np.uint8)
    lab = measure.label(img)
    inten = rng.integers(0, 65535, img.shape).astype(np.uint16)
    regs = measure.regionprops(lab, intensity_image=inten)
    assert [r.label for r in regs] == list(range(1, lab.max() + 1))
    tab = pipeline.region_table(lab, inten)
    for r, row in zip(regs, tab):
        ys, xs = np.nonzero(lab == r.label)
        assert r.area == len(ys) == row[2] and isinstance(r.area, float)
        assert r.centroid == (ys.astype(np.float64).mean(), xs.astype(np.float64).mean()) == (row[3], row[4])
        assert r.bbox == (ys.min(), xs.min(), ys.max() + 1, xs.max() + 1) == tuple(int(v) for v in row[5:9])
        assert tuple(r.coords[0]) == (ys[0], xs[0]) == (row[9], row[10])
        assert r.intensity_mean == inten[ys, xs].mean() == row[12]
        assert r["area"] == r.area
    regs[0].cells = 3
    assert regs[0].cells == 3


def test_disk_and_dilation_vs_edt():
    assert morphology.disk(2).sum() == 13 and morphology.disk(20).sum() == 1257
    rng = np.random.default_rng(4)
    m = rng.random((70, 90)) < 0.01
    m[0, 0] = m[-1, -1] = True
    for r in (1, 2, 5, 20):
        d2 = np.rint(ndi.distance_transform_edt(~m) ** 2).astype(np.int64)
        assert np.array_equal(morphology.binary_dilation(m, morphology.disk(r)), d2 <= r * r)
        # erosion with outside-True is the dual
        e = morphology.binary_erosion(~m, morphology.disk(r))
        assert np.array_equal(e, ~(d2 <= r * r))


def _brute_local_maxima(img):
    h, w = img.shape
    out = np.zeros((h, w), bool)
    if h < 3 or w < 3:
        return out
    lab = _brute_label(np.ones_like(img, dtype=np.uint8))  # placeholder to reuse flood code
    # plateau labelling by value
    vals, inv = np.unique(img, return_inverse=True)
    lab = _brute_label((inv.reshape(h, w) + 1).astype(np.int64))
    lo = img.min()
    for k in range(1, lab.max() + 1):
        ys, xs = np.nonzero(lab == k)
        v = img[ys[0], xs[0]]
        ok = True
        for y, x in zip(ys, xs):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    yy, xx = y + dy, x + dx
                    nv = img[yy, xx] if (0 <= yy < h and 0 <= xx < w) else lo
                    if nv > v or ((not (0 <= yy < h and 0 <= xx < w)) and nv >= v):
                        ok = False
        out[ys, xs] = ok
    return out


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_local_maxima_vs_brute(seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 4, (12, 15)).astype(np.float64)
    assert np.array_equal(morphology.local_maxima(img), _brute_local_maxima(img))
    assert not morphology.local_maxima(np.full((7, 6), 42.0)).any()
    with pytest.warns(UserWarning):
        assert not morphology.local_maxima(np.ones((2, 9))).any()


def test_remove_small_objects_and_fill_holes():
    rng = np.random.default_rng(6)
    m = rng.random((50, 60)) < 0.35
    for conn in (1, 2):
        got = morphology.remove_small_objects(m, 6, connectivity=conn)
        lab = _brute_label(m, conn8=(conn == 2))
        sizes = np.bincount(lab.ravel())
        assert np.array_equal(got, m & (sizes[lab] >= 6))
    filled = ndi.binary_fill_holes(m)
    bg = _brute_label(~m, conn8=False)
    border = set(bg[0]) | set(bg[-1]) | set(bg[:, 0]) | set(bg[:, -1])
    assert np.array_equal(filled, m | ~np.isin(bg, list(border)))


def _brute_otsu(img):
    flat = img.ravel().astype(np.int64)
    lo, hi = flat.min(), flat.max()
    if lo == hi:
        return lo
    best, arg = -1.0, lo
    for t in range(lo, hi):
        a, b = flat[flat <= t], flat[flat > t]
        w1, w2 = np.float32(len(a)), np.float32(len(b))
        v = np.float64(w1 * w2) * (a.sum() / np.float64(w1) - b.sum() / np.float64(w2)) ** 2
        if v > best:
            best, arg = v, t
    return arg


def test_otsu_vs_brute():
    rng = np.random.default_rng(8)
    for _ in range(4):
        img = np.concatenate([rng.integers(10, 60, 700), rng.integers(90, 200, 300)]).astype(np.uint16).reshape(25, 40)
        assert filters.threshold_otsu(img) == _brute_otsu(img)
    assert filters.threshold_otsu(np.full((5, 5), 7, np.uint16)) == 7


def test_pipeline_oracle_runs():
    img = synth.slice_u16(128, 160, seed=5)
    r = pipeline.segment_slice(img)
    assert r["labels"].dtype == np.int32 and r["edt"].dtype == np.float64
    assert r["table"].shape == (r["labels"].max(), len(pipeline.TABLE_COLUMNS))
    assert not (r["refined"] & ~ndi.binary_fill_holes(r["mask"])).any()


def test_refine_oracle_runs():
    _, prob = synth.touching_particles(160, 160, seed=2)
    r = refine.refine_boundaries(prob)
    assert r["markers"].max() > 0 and r["local_max"].dtype == bool
    assert not (r["local_max"] & ~r["binary_mask"]).any()


def test_nanosims_vs_loops():
    planes, roi, set_id, agg = synth.nanosims_stack(96, 7, 20, seed=3)
    red = np.isin(roi, np.nonzero(set_id == 1)[0] + 1)
    green = np.isin(roi, np.nonzero(set_id == 2)[0] + 1)
    tab = nanosims.analyse(planes, red, green, agg)
    lab, n = nanosims.matlab_label(red)
    # column-major numbering: first pixel of ROI i precedes that of ROI i+1 in column-major order
    firsts = [np.flatnonzero((lab == i).T.ravel())[0] for i in range(1, n + 1)]
    assert firsts == sorted(firsts)
    for i in range(n):
        m = lab == i + 1
        for j in range(7):
            assert tab[i, 2 + j] == (planes[j] * m).sum()
        ys, xs = np.nonzero(m)
        assert tab[i, -4] == xs.mean() + 1 and tab[i, -3] == ys.mean() + 1
    c13 = tab[:, 2 + 1] / (tab[:, 2 + 1] + tab[:, 2 + 0])
    assert np.array_equal(tab[:, 9], c13)
    cnt, tot, mean = nanosims.activity_vs_distance(tab[:, 9], tab[:, -1], np.linspace(0, 4, 9))
    assert cnt.sum() == len(tab)


def test_imresize_restatement_properties():
    """oracle/nanosims.py::imresize (MATLAB bicubic with antialiasing, .m:125): equal sizes are the identity bit for
    bit, constants and -- away from the mirrored border -- linear ramps are reproduced, the tap tables sum to one,
    and the separable evaluation equals the dense matrices built from the same tables."""
    from oracle import nanosims as onano

    rng = np.random.default_rng(3)
    a = rng.random((40, 52))
    assert np.array_equal(onano.imresize(a, (40, 52)), a)
    assert np.allclose(onano.imresize(np.full((30, 30), 3.5), (47, 19)), 3.5, rtol=0, atol=1e-14)
    for n_in, n_out in ((52, 30), (40, 64), (7, 7), (5, 50), (50, 5)):
        idx, w = onano.resize_contributions(n_in, n_out)
        assert idx.min() >= 0 and idx.max() < n_in and np.allclose(w.sum(1), 1.0, rtol=0, atol=1e-14)
        if n_out < n_in:  # antialiasing stretches the kernel by n_in / n_out
            assert idx.shape[1] >= int(np.ceil(4 * n_in / n_out))

    def dense(n_in, n_out):
        idx, w = onano.resize_contributions(n_in, n_out)
        m = np.zeros((n_out, n_in))
        np.add.at(m, (np.repeat(np.arange(n_out), idx.shape[1]), idx.ravel()), w.ravel())
        return m

    np.testing.assert_allclose(onano.imresize(a, (64, 30)), dense(40, 64) @ (a @ dense(52, 30).T), rtol=0, atol=1e-14)
    ramp = np.tile(np.arange(20.0), (20, 1))
    assert np.allclose(np.diff(onano.imresize(ramp, (40, 40))[10, 6:34]), 0.5)
    # per-ROI sums under resized masks: with equal sizes they are the plain masked sums
    planes, roi, _, _ = synth.nanosims_stack(64, 5, 12, seed=5)
    n = int(roi.max())
    np.testing.assert_allclose(onano.roi_sums_resized(planes, roi, n), onano.roi_sums(planes, roi, n), rtol=1e-13)
