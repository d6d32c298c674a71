"""Randomised differential run of the drop-in primitives against scipy / the scikit-image restatement
(not part of the test suite; run on a GPU box: python scratch/fuzz_primitives.py [seconds])."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scipy import ndimage as ndi
from oracle import skimage_shim as sk
from oracle.skimage_shim import morphology as om
from particle_col_image_segmentation_b200 import measure as pm, morphology as pmo, ndimage as pnd, filters as pf

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(time.time()) % 100000)
t0 = time.time(); n = 0; bad = 0
def chk(name, got, want, info):
    global bad
    if isinstance(want, tuple):
        ok = all(np.array_equal(g, w) for g, w in zip(got, want))
    else:
        ok = np.array_equal(got, want) and got.dtype == want.dtype
    if not ok:
        bad += 1
        print("MISMATCH", name, info, flush=True)
while time.time() - t0 < budget:
    n += 1
    H = int(rng.integers(1, 200)); W = int(rng.integers(1, 400))
    p = rng.uniform(0.02, 0.98)
    m = rng.random((H, W)) < p
    if rng.random() < 0.3:
        m = ndi.binary_opening(m, iterations=int(rng.integers(1, 3)))
    info = (H, W, round(p, 2))
    try:
        for conn, st in ((2, np.ones((3, 3))), (1, None)):
            chk(f"label{conn}", pm.label(m, connectivity=conn), ndi.label(m, structure=st)[0].astype(np.int32), info)
        cls = rng.integers(0, int(rng.integers(2, 6)), (H, W)).astype(np.uint8)
        chk("label_multi", pm.label(cls), sk.measure.label(cls), info)
        chk("fill_holes", pnd.binary_fill_holes(m), ndi.binary_fill_holes(m), info)
        if not m.all():
            chk("edt", pnd.distance_transform_edt(m), ndi.distance_transform_edt(m), info)
        r = int(rng.choice([1, 2, 3, 5, 20]))
        chk(f"dilate{r}", pmo.binary_dilation(m, om.disk(r)), om.binary_dilation(m, om.disk(r)), info)
        chk(f"erode{r}", pmo.binary_erosion(m, om.disk(min(r, 3))), om.binary_erosion(m, om.disk(min(r, 3))), info)
        ms = int(rng.choice([1, 3, 20, 100]))
        chk("remove_small", pmo.remove_small_objects(m, ms, connectivity=2), om.remove_small_objects(m, ms, connectivity=2), info)
        if H >= 3 and W >= 3:
            f = ndi.gaussian_filter(rng.random((H, W)), rng.uniform(0.3, 3)).astype(np.float64)
            f = np.round(f * rng.choice([5, 50, 1e6])) if rng.random() < 0.5 else f
            chk("local_maxima", pmo.local_maxima(f), om.local_maxima(f), info)
        sz = int(rng.choice([3, 5, 7]))
        a = rng.integers(0, int(rng.choice([3, 256])), (H, W)).astype(np.uint8)
        chk(f"median{sz}", pnd.median_filter(a, size=sz), ndi.median_filter(a, size=sz), info)
        u = rng.integers(0, int(rng.choice([2, 4000, 65536])), (H, W)).astype(np.uint16)
        if int(pf.threshold_otsu(u)) != int(sk.filters.threshold_otsu(u)):
            bad += 1; print("MISMATCH otsu", info, flush=True)
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("ERROR", type(e).__name__, str(e)[:160], info, flush=True)
print(f"fuzz primitives: {n} cases, {bad} bad, {time.time() - t0:.0f} s")
