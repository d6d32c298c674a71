"""CPU oracle of the full z-stack segment pipeline (the benchmarked "step").

TEST INFRASTRUCTURE ONLY.  The reference joins ``split_zstack.py`` to
``tiff_analysis.py`` through files and ilastik (an external GUI classifier), so
there is no single reference function for "threshold -> label -> refine -> EDT
-> regionprops" on a uint16 slice.  The pipeline below is the composition the
north_star names, built only from library calls whose semantics the reference
uses (or, for Otsu / small-object removal, the library function the north_star
implies -- SURVEY.md section 0.1):

  1. ``t = threshold_otsu(slice)``; ``mask = slice > t``        (ilastik's role)
  2. ``mask = median_filter(mask.astype(uint8), size=5)``        tiff_analysis.py:122, :643
  3. ``labels = label(mask)``  (8-connected, int32)              tiff_analysis.py:260, :743
  4. ``regionprops(labels, intensity_image=slice)``              tiff_analysis.py:746
  5. ``refined = binary_fill_holes(remove_small_objects(mask, min_size, connectivity=2))``
                                                                 tiff_analysis.py:769 (area filter), :880
  6. ``edt = distance_transform_edt(refined)``                   refine_boundaries.py:60
"""

import numpy as np
from scipy import ndimage as ndi

from .skimage_shim.filters import threshold_otsu
from .skimage_shim.measure import label
from .skimage_shim.morphology import remove_small_objects

TABLE_COLUMNS = ("z", "label", "area", "centroid_y", "centroid_x", "min_row", "min_col", "max_row", "max_col", "first_row", "first_col", "intensity_sum", "intensity_mean")


def region_table(labels, intensity=None, z=0):
    """Per-label table with the columns the reference consumes (area, centroid,
    bbox, first pixel: tiff_analysis.py:754-773, :843-863, :1041-1044) plus the
    integrated and mean intensity.  float64, one row per label, label order."""
    n = int(labels.max())
    flat = labels.ravel()
    h, w = labels.shape
    area = np.bincount(flat, minlength=n + 1)[1:].astype(np.float64)
    yy, xx = np.divmod(np.arange(flat.size), w)
    sy = np.bincount(flat, weights=yy, minlength=n + 1)[1:]
    sx = np.bincount(flat, weights=xx, minlength=n + 1)[1:]
    objs = ndi.find_objects(labels)
    bbox = np.array([[s[0].start, s[1].start, s[0].stop, s[1].stop] for s in objs], dtype=np.float64).reshape(n, 4)
    first = np.full(n + 1, flat.size, dtype=np.int64)
    np.minimum.at(first, flat, np.arange(flat.size))
    fy, fx = np.divmod(first[1:], w)
    if intensity is not None:
        si = np.bincount(flat, weights=intensity.ravel().astype(np.float64), minlength=n + 1)[1:]
    else:
        si = np.zeros(n)
    with np.errstate(invalid="ignore", divide="ignore"):
        tab = np.column_stack([np.full(n, float(z)), np.arange(1, n + 1, dtype=np.float64), area, sy / area, sx / area, bbox, fy, fx, si, si / area])
    return tab.reshape(n, len(TABLE_COLUMNS))


def segment_slice(img, denoise_size=5, min_size=20, z=0):
    """One slice through the full pipeline; returns a dict of the five outputs."""
    img = np.asarray(img)
    t = threshold_otsu(img)
    mask = img > t
    if denoise_size and denoise_size > 1:
        mask = ndi.median_filter(mask.astype(np.uint8), size=denoise_size).astype(bool)
    labels = label(mask)  # bool -> scipy.ndimage.label, 8-connected, int32
    table = region_table(labels, img, z=z)
    refined = remove_small_objects(mask, min_size=min_size, connectivity=2)
    refined = ndi.binary_fill_holes(refined)
    edt = ndi.distance_transform_edt(refined)
    return {"threshold": int(t), "mask": mask, "labels": labels, "refined": refined, "edt": edt, "table": table}


def segment_zstack(stack, denoise_size=5, min_size=20, z0=0):
    """``(Z, Y, X)`` stack, slice by slice (split_zstack.py:52 iterates slices; every
    slice is an independent 2-D problem)."""
    outs = [segment_slice(s, denoise_size, min_size, z=z0 + i) for i, s in enumerate(stack)]
    res = {k: np.stack([o[k] for o in outs]) for k in ("mask", "labels", "refined", "edt")}
    res["threshold"] = np.array([o["threshold"] for o in outs])
    res["table"] = np.concatenate([o["table"] for o in outs], axis=0)
    res["counts"] = np.array([len(o["table"]) for o in outs])
    return res
