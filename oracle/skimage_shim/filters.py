"""``skimage.filters`` restatement (threshold_otsu only).  TEST INFRASTRUCTURE ONLY.

No reference call site (``refine_boundaries.py:22`` imports ``filters`` but never
uses it); the north_star names Otsu, so the oracle is the library function.
"""

import numpy as np

__all__ = ["threshold_otsu"]


def _histogram_integer(image):
    """``skimage.exposure.histogram`` for integer images with
    ``source_range='image'``: one bin per integer from min to max."""
    flat = image.reshape(-1)
    lo, hi = int(flat.min()), int(flat.max())
    counts = np.bincount((flat.astype(np.int64) - lo), minlength=hi - lo + 1)
    centers = np.arange(lo, hi + 1)
    return counts, centers


def threshold_otsu(image=None, nbins=256, *, hist=None):
    """Otsu threshold, restating scikit-image 0.25.2 ``thresholding.threshold_otsu``.

    Integer images use one bin per integer between min and max (``nbins`` is
    ignored); counts are cast to float32 (``_validate_image_histogram``), the
    cumulative weights therefore run in float32 and the class means in float64;
    the between-class variance is ``w1[:-1] * w2[1:] * (m1[:-1] - m2[1:])**2`` and
    the first arg-max wins.  A single-valued image returns that value.  The mask
    is ``image > threshold``.
    """
    if hist is not None:
        if isinstance(hist, (tuple, list)):
            counts, centers = hist
        else:
            counts, centers = hist, np.arange(len(hist))
        counts = np.asarray(counts)
        centers = np.asarray(centers)
    else:
        image = np.asarray(image)
        first = image.reshape(-1)[0]
        if np.all(image == first):
            return first
        if np.issubdtype(image.dtype, np.integer):
            counts, centers = _histogram_integer(image)
        else:
            counts, edges = np.histogram(image.reshape(-1), bins=nbins)
            centers = (edges[:-1] + edges[1:]) / 2.0
    counts = counts.astype("float32", copy=False)
    weight1 = np.cumsum(counts)
    weight2 = np.cumsum(counts[::-1])[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        mean1 = np.cumsum(counts * centers) / weight1
        mean2 = (np.cumsum((counts * centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    idx = np.argmax(variance12)
    return centers[idx]
