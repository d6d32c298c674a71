"""Parity of every CUDA primitive with the oracle (scipy + scikit-image restatement).

Bit-exact for masks, labels, counts and -- because the EDT is an exact integer
squared distance followed by an IEEE sqrt -- for distances too.  All calls go
through the drop-in entry points, i.e. through the C ABI (libpcs.so).
"""

import numpy as np
import pytest
import torch
from scipy import ndimage as ndi

from oracle import pipeline as opipe
from oracle.skimage_shim import filters as ofilters
from oracle.skimage_shim import measure as omeasure
from oracle.skimage_shim import morphology as omorph
from particle_col_image_segmentation_b200 import synth

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1), (3, 5), (7, 31), (16, 32), (9, 33), (40, 64), (33, 65), (64, 100), (130, 257), (256, 512)]


@pytest.fixture(scope="module")
def pcs():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from particle_col_image_segmentation_b200 import filters, measure, morphology, ndimage, ops

    class NS:
        pass

    ns = NS()
    ns.filters, ns.measure, ns.morphology, ns.ndimage, ns.ops = filters, measure, morphology, ndimage, ops
    return ns


def assert_same(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert got.dtype == want.dtype, f"{what}: dtype {got.dtype} != {want.dtype}"
    if not np.array_equal(got, want):
        bad = np.argwhere(got != want)
        raise AssertionError(f"{what}: {len(bad)} of {got.size} differ; first at {bad[0].tolist()}: got {got[tuple(bad[0])]} want {want[tuple(bad[0])]}")


def masks(shape, seed):
    rng = np.random.default_rng(seed)
    h, w = shape
    yield "empty", np.zeros(shape, bool)
    yield "full", np.ones(shape, bool)
    for p in (0.02, 0.3, 0.5, 0.62, 0.9):
        yield f"rand{p}", rng.random(shape) < p
    yy, xx = np.mgrid[0:h, 0:w]
    yield "checker", ((yy + xx) % 2 == 0)
    yield "diag", (yy == xx) | (yy + xx == w - 1)
    yield "hstripes", (yy % 2 == 0)
    yield "vstripes", (xx % 3 == 0)
    blob = ndi.binary_dilation(rng.random(shape) < 0.01, iterations=3)
    yield "blobs", blob
    # comb: teeth joined at the bottom (components that merge late in raster order)
    comb = (xx % 4 == 0) | (yy == h - 1)
    yield "comb", comb
    spiral = np.zeros(shape, bool)
    spiral[::2, :] = True
    spiral[1::4, -1:] = True
    spiral[3::4, :1] = True
    yield "serpentine", spiral


# ---------------------------------------------------------------- K2 / packing
@pytest.mark.parametrize("shape", SHAPES)
def test_pack_unpack_compare(pcs, shape):
    rng = np.random.default_rng(1)
    dev = torch.device("cuda")
    for dt, hi in ((np.uint8, 6), (np.uint16, 60000), (np.int32, 1000), (np.float32, 1.0), (np.float64, 1.0)):
        a = (rng.random(shape) * hi).astype(dt)
        t = torch.from_numpy(a).to(dev).unsqueeze(0)
        thr = hi / 2
        for op in (">", ">=", "<", "<=", "==", "!="):
            th = a.flat[0] if op in ("==", "!=") else (int(thr) if np.issubdtype(dt, np.integer) else float(dt(thr)))
            bits, mask = pcs.ops.compare(t, op, th, want_mask=True)
            want = eval(f"a {op} th")
            assert_same(mask[0].cpu().numpy().astype(bool), want, f"compare {dt.__name__} {op}")
            assert_same(pcs.ops.unpack(bits, shape[1], torch.bool)[0].cpu().numpy(), want, f"bits {dt.__name__} {op}")
            assert int(pcs.ops.count(bits, shape[1])[0]) == int(want.sum())


def test_lut_member_assign(pcs):
    rng = np.random.default_rng(2)
    a = rng.integers(0, 6, (70, 131)).astype(np.uint8)
    t = torch.from_numpy(a).cuda().unsqueeze(0)
    bits, mask = pcs.ops.member_u8(t, [1, 4], want_mask=True)
    assert_same(mask[0].cpu().numpy().astype(bool), np.isin(a, [1, 4]))
    lut = np.arange(256, dtype=np.uint8)
    lut[3], lut[2] = 5, 4
    t2 = t.clone()
    pcs.ops.lut_u8_(t2, lut)
    assert_same(t2[0].cpu().numpy(), lut[a])
    pcs.ops.assign_where_u8_(t2, bits, 9)
    want = lut[a]
    want[np.isin(a, [1, 4])] = 9
    assert_same(t2[0].cpu().numpy(), want)
    nb = pcs.ops.logic(bits, None, "not", a.shape[1])
    assert_same(pcs.ops.unpack(nb, a.shape[1], torch.bool)[0].cpu().numpy(), ~np.isin(a, [1, 4]))


# ---------------------------------------------------------------- K1
def test_otsu(pcs):
    rng = np.random.default_rng(3)
    cases = [synth.slice_u16(128, 160, seed=s) for s in (1, 2)]
    cases.append(synth.slice_u16(512, 512, seed=1001))
    cases.append(rng.integers(0, 65536, (200, 300)).astype(np.uint16))
    cases.append(rng.integers(100, 110, (64, 64)).astype(np.uint16))
    cases.append(np.full((33, 47), 1234, np.uint16))
    two = np.zeros((50, 50), np.uint16)
    two[:10] = 65535
    cases.append(two)
    for i, img in enumerate(cases):
        got = pcs.filters.threshold_otsu(img)
        want = ofilters.threshold_otsu(img)
        assert int(got) == int(want), f"case {i}: got {got} want {want}"
    # batched, per-slice thresholds
    st = synth.zstack_u16(3, 96, 128, seed=9)
    thr = pcs.ops.otsu_u16(torch.from_numpy(st).cuda())
    assert thr.cpu().tolist() == [int(ofilters.threshold_otsu(s)) for s in st]


def test_histogram_exact_when_counters_saturate(pcs):
    """The histogram CTA walks several tiles and flushes its packed 16-bit counters only when one could
    overflow during the next tile: images dominated by one or two values must force those flushes and stay
    exact; odd sizes put later slices off 16-byte alignment (scalar path)."""
    rng = np.random.default_rng(12)
    for shape, hot in (((2, 1024, 1024), 0.7), ((1, 2048, 2048), 0.995), ((3, 333, 1001), 0.5), ((2, 2048, 2048), 0.0)):
        st = rng.integers(0, 65536, shape).astype(np.uint16)
        m = rng.random(shape) < hot
        st[m] = 1000
        st[m & (rng.random(shape) < 0.3)] = 1001  # the neighbour counter shares the 32-bit word with bin 1000
        t = torch.from_numpy(st).cuda()
        thr, hist = pcs.ops.otsu_u16(t, return_hist=True)
        for i in range(shape[0]):
            want = np.bincount(st[i].ravel(), minlength=65536)
            assert np.array_equal(hist[i].cpu().numpy().astype(np.int64), want), (shape, hot, i)
            assert int(thr[i]) == int(ofilters.threshold_otsu(st[i])), (shape, hot, i)


# ---------------------------------------------------------------- K3
@pytest.mark.parametrize("size", [3, 5, 7])
def test_median_u8(pcs, size):
    rng = np.random.default_rng(4)
    # (1, 1) .. (3, 40): a side shorter than the window, where scipy's 'reflect' folds an index more than once
    for shape in [(1, 1), (2, 5), (1, 70), (3, 40), (6, 2), (7, 7), (16, 32), (33, 65), (130, 257)]:
        for hi in (4, 256):
            a = rng.integers(0, hi, shape).astype(np.uint8)
            assert_same(pcs.ndimage.median_filter(a, size=size), ndi.median_filter(a, size=size), f"median {shape} hi={hi}")
        m = rng.random(shape) < 0.5
        bits = pcs.ops.pack(torch.from_numpy(m).cuda().unsqueeze(0))
        got = pcs.ops.unpack(pcs.ops.majority(bits, shape[1], size), shape[1], torch.bool)[0].cpu().numpy()
        assert_same(got, ndi.median_filter(m.astype(np.uint8), size=size).astype(bool), f"majority {shape}")
    cls = synth.class_image(256, 256, seed=5, noise=0.05)
    assert_same(pcs.ndimage.median_filter(cls, size=5), ndi.median_filter(cls, size=5), "class image")


# ---------------------------------------------------------------- K4
@pytest.mark.parametrize("shape", SHAPES)
def test_label_binary(pcs, shape):
    for name, m in masks(shape, 5):
        for conn, st in ((2, np.ones((3, 3))), (1, None)):
            want, n = ndi.label(m, structure=st)
            got, gn = pcs.measure.label(m, connectivity=conn, return_num=True)
            assert gn == n, f"{name} conn{conn}: count {gn} != {n}"
            assert_same(got, want, f"label {name} conn{conn} {shape}")
    got, n = pcs.ndimage.label(m)
    assert_same(got, ndi.label(m)[0], "ndimage.label default")


@pytest.mark.parametrize("shape", [(3, 5), (9, 33), (64, 100), (130, 257)])
def test_label_multivalued(pcs, shape):
    rng = np.random.default_rng(6)
    for nval in (2, 4, 7):
        a = rng.integers(0, nval, shape).astype(np.uint8)
        want = omeasure.label(a)
        got = pcs.measure.label(a)
        assert_same(got, want, f"multi label nval={nval}")
        assert_same(pcs.measure.label(a, connectivity=1), omeasure.label(a, connectivity=1), "multi conn1")
    cls = synth.class_image(256, 256, seed=6, noise=0.02)
    assert_same(pcs.measure.label(cls), omeasure.label(cls), "class image")


def test_label_batched(pcs):
    rng = np.random.default_rng(7)
    m = rng.random((5, 70, 97)) < 0.45
    bits = pcs.ops.pack(torch.from_numpy(m).cuda())
    labels, counts, offsets = pcs.ops.label_bits(bits, 97, connectivity=8)
    for i in range(5):
        want, n = ndi.label(m[i], structure=np.ones((3, 3)))
        assert_same(labels[i].cpu().numpy(), want, f"slice {i}")
        assert int(counts[i]) == n
    assert offsets.cpu().tolist() == [0] + np.cumsum(counts.cpu().numpy()).tolist()


# ---------------------------------------------------------------- K8
def test_regionprops(pcs):
    rng = np.random.default_rng(8)
    img = synth.slice_u16(200, 230, seed=8)
    m = img > 2000
    lab = ndi.label(m, structure=np.ones((3, 3)))[0]
    want = opipe.region_table(lab, img)
    regs = pcs.measure.regionprops(lab, intensity_image=img)
    assert [r.label for r in regs] == list(range(1, lab.max() + 1))
    for r, w in zip(regs, want):
        assert r.area == w[2] and r.centroid == (w[3], w[4]), r.label
        assert r.bbox == tuple(int(v) for v in w[5:9]) and r.first_pixel == (int(w[9]), int(w[10]))
        assert r.intensity_sum == w[11] and r.intensity_mean == w[12]
        assert tuple(r.coords[0]) == r.first_pixel
    # labels with gaps and a multi-valued int64 label image
    a = rng.integers(0, 4, (90, 131)).astype(np.uint8)
    lab2 = omeasure.label(a)
    lab2[lab2 == 3] = 0
    regs2 = pcs.measure.regionprops(lab2)
    oregs = omeasure.regionprops(lab2)
    assert [r.label for r in regs2] == [r.label for r in oregs]
    for r, o in zip(regs2, oregs):
        assert r.area == o.area and r.centroid == o.centroid and r.bbox == o.bbox and tuple(o.coords[0]) == r.first_pixel
    # overlap counts (tiff_analysis.py:268-279)
    other = rng.random(lab.shape) < 0.3
    regs3 = pcs.measure.regionprops(lab, overlap_mask=other)
    for r in regs3:
        assert r.overlap == int(((lab == r.label) & other).sum())


# ---------------------------------------------------------------- K6 / small objects / selection
@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (9, 33), (64, 100), (130, 257)])
def test_fill_holes_remove_small(pcs, shape):
    for name, m in masks(shape, 9):
        assert_same(pcs.ndimage.binary_fill_holes(m), ndi.binary_fill_holes(m), f"fill holes {name} {shape}")
        for conn in (1, 2):
            assert_same(pcs.morphology.remove_small_objects(m, 5, connectivity=conn), omorph.remove_small_objects(m, 5, connectivity=conn), f"remove small {name} conn{conn}")


def test_select_components(pcs):
    rng = np.random.default_rng(10)
    m = rng.random((80, 140)) < 0.4
    seeds = (rng.random(m.shape) < 0.01) & m
    lab = ndi.label(m, structure=np.ones((3, 3)))[0]
    want = np.isin(lab, np.unique(lab[seeds])) & m
    bm = pcs.ops.pack(torch.from_numpy(m).cuda().unsqueeze(0))
    bs = pcs.ops.pack(torch.from_numpy(seeds).cuda().unsqueeze(0))
    got = pcs.ops.unpack(pcs.ops.select_components(bm, bs, 140), 140, torch.bool)[0].cpu().numpy()
    assert_same(got, want)


# ---------------------------------------------------------------- K5
@pytest.mark.parametrize("shape", [(3, 5), (9, 33), (64, 100), (130, 257)])
def test_morphology(pcs, shape):
    rng = np.random.default_rng(11)
    fps = {f"disk{r}": omorph.disk(r) for r in (1, 2, 3, 5, 20)}
    fps["sq3"] = np.ones((3, 3), np.uint8)
    fps["rect"] = np.ones((3, 7), np.uint8)
    fps["wide"] = np.ones((1, 41), np.uint8)
    asym = np.zeros((5, 5), np.uint8)
    asym[0, 0] = asym[2, 2] = asym[3, 4] = asym[4, 1] = 1
    fps["asym"] = asym
    fps["even"] = np.ones((2, 4), np.uint8)
    for p in (0.02, 0.5, 0.97):
        m = rng.random(shape) < p
        for name, fp in fps.items():
            for bv in (0, 1):
                assert_same(pcs.ndimage.binary_dilation(m, fp, border_value=bv), ndi.binary_dilation(m, fp, border_value=bv), f"dilate {name} bv{bv} p{p}")
                assert_same(pcs.ndimage.binary_erosion(m, fp, border_value=bv), ndi.binary_erosion(m, fp, border_value=bv), f"erode {name} bv{bv} p{p}")
            assert_same(pcs.morphology.binary_dilation(m, fp), omorph.binary_dilation(m, fp), f"sk dilate {name}")
            assert_same(pcs.morphology.binary_erosion(m, fp), omorph.binary_erosion(m, fp), f"sk erode {name}")
        for name in ("disk2", "sq3"):
            assert_same(pcs.morphology.binary_opening(m, fps[name]), omorph.binary_opening(m, fps[name]), f"open {name}")
            assert_same(pcs.morphology.binary_closing(m, fps[name]), omorph.binary_closing(m, fps[name]), f"close {name}")
            assert_same(pcs.ndimage.binary_opening(m, fps[name]), ndi.binary_opening(m, fps[name]), f"ndi open {name}")
            assert_same(pcs.ndimage.binary_closing(m, fps[name]), ndi.binary_closing(m, fps[name]), f"ndi close {name}")


# ---------------------------------------------------------------- K7
@pytest.mark.parametrize("shape", SHAPES + [(300, 1000), (1000, 300)])
def test_edt(pcs, shape):
    for name, m in masks(shape, 12):
        want = ndi.distance_transform_edt(m)
        got = pcs.ndimage.distance_transform_edt(m)
        assert_same(got, want, f"edt {name} {shape}")
    rng = np.random.default_rng(13)
    far = np.ones(shape, bool)
    far[rng.integers(0, shape[0]), rng.integers(0, shape[1])] = False
    assert_same(pcs.ndimage.distance_transform_edt(far), ndi.distance_transform_edt(far), "single background pixel")
    sq = pcs.ndimage.distance_transform_edt(far, return_squared=True)
    assert_same(sq, np.rint(ndi.distance_transform_edt(far) ** 2).astype(np.int32), "squared")


# ---------------------------------------------------------------- K9
@pytest.mark.parametrize("shape", [(3, 3), (3, 5), (9, 33), (64, 100), (130, 257)])
def test_local_maxima(pcs, shape):
    rng = np.random.default_rng(14)
    for nval in (2, 3, 6, 50):
        img = rng.integers(0, nval, shape).astype(np.float64)
        assert_same(pcs.morphology.local_maxima(img), omorph.local_maxima(img), f"local maxima nval={nval}")
    assert not pcs.morphology.local_maxima(np.full(shape, 3.0)).any()
    m = rng.random(shape) < 0.7
    d = ndi.distance_transform_edt(m)
    assert_same(pcs.morphology.local_maxima(d), omorph.local_maxima(d), "maxima of an EDT")
    i32 = rng.integers(0, 5, shape).astype(np.int32)
    assert_same(pcs.morphology.local_maxima(i32, connectivity=1), omorph.local_maxima(i32, connectivity=1), "int32 conn1")


# ---------------------------------------------------------------- K11
def test_roi_sums_and_min_dist(pcs):
    from oracle import nanosims as on

    planes, roi, set_id, agg = synth.nanosims_stack(96, 7, 20, seed=3)
    n = int(roi.max())
    want = on.roi_sums(planes, roi, n)
    got = pcs.ops.roi_sums(torch.from_numpy(roi).cuda(), torch.from_numpy(planes).cuda(), n).cpu().numpy()
    assert_same(got, want)  # Poisson counts: integer-valued doubles, exact in any order
    rng = np.random.default_rng(15)
    a, b = rng.random((37, 2)) * 100, rng.random((53, 2)) * 100
    want_a, want_b = on.nearest_between(a, b)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    assert_same(pcs.ops.min_dist(ta, tb).cpu().numpy(), want_a)
    assert_same(pcs.ops.min_dist(tb, ta).cpu().numpy(), want_b)
