"""Seeded differential cases shared by tests/test_gpu_fuzz.py (fixed seed lists, run by the driver's GPU tier) and
the open-ended fuzzers in scratch/ (python scratch/fuzz_*.py [seconds] [first seed]).  Case i of a family is a pure
function of its seed, so any failure is reproducible from the seed alone.

Each `check_*` runs one case on the device and compares it with the oracle (scipy for the scipy-backed calls,
oracle/skimage_shim for the scikit-image ones, oracle/l2.py -- pinned to the unmodified reference by tests/golden --
for the tiff_analysis mirrors).  They raise AssertionError with the seed in the message."""
import numpy as np
from scipy import ndimage as ndi


def pipeline_case(seed):
    rng = np.random.default_rng(seed)
    from particle_col_image_segmentation_b200 import synth

    kind = int(rng.integers(0, 4))
    Z, H, W = int(rng.integers(1, 4)), int(rng.integers(1, 300)), int(rng.integers(1, 700))
    if kind == 0:
        st = synth.zstack_u16(Z, max(H, 8), max(W, 8), seed=int(rng.integers(1 << 30)))
    elif kind == 1:
        st = rng.integers(0, int(rng.choice([2, 300, 65536])), (Z, H, W)).astype(np.uint16)
    elif kind == 2:
        st = (rng.random((Z, H, W)) < rng.uniform(0.05, 0.95)).astype(np.uint16) * int(rng.integers(1, 60000))
        st += rng.integers(0, 3, st.shape).astype(np.uint16)
    else:
        base = ndi.gaussian_filter(rng.random((Z, H, W)), (0, rng.uniform(0.5, 6), rng.uniform(0.5, 6)))
        st = (base * 60000).astype(np.uint16)
    return st, int(rng.choice([0, 3, 5, 7])), int(rng.choice([1, 2, 20, 200])), int(rng.integers(1, 4))


def check_pipeline(seed):
    from oracle import pipeline as opipe
    from particle_col_image_segmentation_b200 import split_zstack

    st, dn, ms, ch = pipeline_case(seed)
    got = split_zstack.segment_zstack(st, denoise_size=dn, min_size=ms, chunk=ch)
    want = opipe.segment_zstack(st, denoise_size=dn, min_size=ms)
    for k in ("threshold", "mask", "labels", "refined", "edt", "table", "counts"):
        assert np.array_equal(got[k], want[k]), f"pipeline seed {seed}: {k} differs (shape {st.shape}, denoise {dn}, min_size {ms}, chunk {ch})"


def check_primitives(seed):
    from oracle import skimage_shim as sk
    from oracle.skimage_shim import morphology as om
    from particle_col_image_segmentation_b200 import filters as pf
    from particle_col_image_segmentation_b200 import measure as pm
    from particle_col_image_segmentation_b200 import morphology as pmo
    from particle_col_image_segmentation_b200 import ndimage as pnd

    rng = np.random.default_rng(seed)
    H, W = int(rng.integers(1, 200)), int(rng.integers(1, 400))
    p = rng.uniform(0.02, 0.98)
    m = rng.random((H, W)) < p
    if rng.random() < 0.3:
        m = ndi.binary_opening(m, iterations=int(rng.integers(1, 3)))
    info = f"primitives seed {seed} ({H}x{W}, p={p:.2f})"

    def chk(name, got, want):
        assert np.array_equal(got, want) and got.dtype == want.dtype, f"{info}: {name}"

    for conn, st in ((2, np.ones((3, 3))), (1, None)):
        chk(f"label{conn}", pm.label(m, connectivity=conn), ndi.label(m, structure=st)[0].astype(np.int32))
    cls = rng.integers(0, int(rng.integers(2, 6)), (H, W)).astype(np.uint8)
    chk("label_multi", pm.label(cls), sk.measure.label(cls))
    chk("fill_holes", pnd.binary_fill_holes(m), ndi.binary_fill_holes(m))
    if not m.all():
        chk("edt", pnd.distance_transform_edt(m), ndi.distance_transform_edt(m))
    r = int(rng.choice([1, 2, 3, 5, 20]))
    chk(f"dilate{r}", pmo.binary_dilation(m, om.disk(r)), om.binary_dilation(m, om.disk(r)))
    chk(f"erode{r}", pmo.binary_erosion(m, om.disk(min(r, 3))), om.binary_erosion(m, om.disk(min(r, 3))))
    ms = int(rng.choice([1, 3, 20, 100]))
    chk("remove_small", pmo.remove_small_objects(m, ms, connectivity=2), om.remove_small_objects(m, ms, connectivity=2))
    if H >= 3 and W >= 3:
        f = ndi.gaussian_filter(rng.random((H, W)), rng.uniform(0.3, 3)).astype(np.float64)
        f = np.round(f * rng.choice([5, 50, 1e6])) if rng.random() < 0.5 else f
        chk("local_maxima", pmo.local_maxima(f), om.local_maxima(f))
    sz = int(rng.choice([3, 5, 7]))
    a = rng.integers(0, int(rng.choice([3, 256])), (H, W)).astype(np.uint8)
    chk(f"median{sz}", pnd.median_filter(a, size=sz), ndi.median_filter(a, size=sz))
    u = rng.integers(0, int(rng.choice([2, 4000, 65536])), (H, W)).astype(np.uint16)
    assert int(pf.threshold_otsu(u)) == int(sk.filters.threshold_otsu(u)), f"{info}: otsu"


L2_TYPES = {1: "C3M10", 2: "Particle", 3: "Background"}


def check_l2(seed):
    """Returns "raises" when the reference's own ValueError path was taken (an image with clusters but no cells,
    tiff_analysis.py:781), else "ok"."""
    from helpers import assert_summary_equal
    from oracle import l2 as ol2
    from particle_col_image_segmentation_b200 import synth
    from particle_col_image_segmentation_b200 import tiff_analysis as ta

    rng = np.random.default_rng(seed)
    H, W, iseed = int(rng.integers(40, 400)), int(rng.integers(40, 500)), int(rng.integers(1 << 30))
    noise = float(rng.choice([0.0, 0.01, 0.05, 0.15]))
    info = f"l2 seed {seed} ({H}x{W}, image seed {iseed}, noise {noise})"
    raw = synth.class_image(H, W, seed=iseed, noise=noise)
    den = ta.median_filter(raw, size=5)
    assert np.array_equal(den, ndi.median_filter(raw, size=5)), f"{info}: median"
    try:
        want = ol2.get_cell_positions_and_areas(den, L2_TYPES, merged=True)
    except ValueError as e_ref:
        try:
            ta.get_cell_positions_and_areas(den, L2_TYPES, merged=True)
        except ValueError as e_dev:
            assert str(e_dev) == str(e_ref), f"{info}: different error text"
            return "raises"
        raise AssertionError(f"{info}: the reference raises ValueError, the device path does not")
    got = ta.get_cell_positions_and_areas(den, L2_TYPES, merged=True)
    assert_summary_equal(ol2.summarize_positions(got), ol2.summarize_positions(want))
    g, ng = ta.recreate_particle_area(den, L2_TYPES, got[2])
    w, nw = ol2.recreate_particle_area(den, L2_TYPES, want[2])
    assert np.array_equal(g, w) and ng == nw, f"{info}: recreate_particle_area"
    other = synth.class_image(H, W, seed=iseed + 50, noise=0.0)
    assert np.array_equal(ta.combine_cell_positions_and_clusters(den, other), ol2.combine_cell_positions_and_clusters(den, other)), f"{info}: combine"
    up, cnt = ta.fill_particle_area(den, 2, 1, 2)
    up2, cnt2 = ol2.fill_particle_area(den, 2, 1, 2)
    assert np.array_equal(up, up2) and int(cnt) == int(cnt2), f"{info}: fill_particle_area"
    c1 = ta.get_cell_counts_and_densities(got[0], got[1], got[2])
    c2 = ol2.get_cell_counts_and_densities(want[0], want[1], want[2])
    assert c1 == c2, f"{info}: counts"
    return "ok"
