"""``skimage.morphology`` restatement.  TEST INFRASTRUCTURE ONLY."""

import warnings

import numpy as np
from scipy import ndimage as ndi
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

__all__ = [
    "disk",
    "square",
    "binary_dilation",
    "binary_erosion",
    "binary_opening",
    "binary_closing",
    "local_maxima",
    "remove_small_objects",
    "remove_small_holes",
]


def disk(radius, dtype=np.uint8, *, strict_radius=True, decomposition=None):
    """``(2r+1, 2r+1)`` footprint with ``x**2 + y**2 <= r**2``
    (scikit-image ``footprints.disk``; tiff_analysis.py:827, :990)."""
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    if not strict_radius:
        radius += 0.5
    return np.array((X**2 + Y**2) <= radius**2, dtype=dtype)


def square(width, dtype=np.uint8):
    return np.ones((width, width), dtype=dtype)


def _default_fp(image):
    return ndi.generate_binary_structure(np.ndim(image), 1)


def binary_dilation(image, footprint=None, out=None, *, mode="ignore"):
    """``ndi.binary_dilation(image, structure=footprint)``; outside the image is
    False unless ``mode='max'`` (scikit-image ``binary.py``; tiff_analysis.py:828, :990)."""
    if footprint is None:
        footprint = _default_fp(image)
    border = mode == "max"
    res = ndi.binary_dilation(np.asarray(image) != 0, structure=np.asarray(footprint) != 0, border_value=border)
    if out is not None:
        out[...] = res
        return out
    return res


def binary_erosion(image, footprint=None, out=None, *, mode="ignore"):
    """``ndi.binary_erosion``; outside the image is True unless ``mode='min'``
    (scikit-image 0.25 default ``mode='ignore'``).  No reference call site."""
    if footprint is None:
        footprint = _default_fp(image)
    border = mode != "min"
    res = ndi.binary_erosion(np.asarray(image) != 0, structure=np.asarray(footprint) != 0, border_value=border)
    if out is not None:
        out[...] = res
        return out
    return res


def binary_opening(image, footprint=None, out=None, *, mode="ignore"):
    """erosion then dilation (scikit-image ``binary_opening``).  No reference call site."""
    return binary_dilation(binary_erosion(image, footprint, mode=mode), footprint, out=out, mode=mode)


def binary_closing(image, footprint=None, out=None, *, mode="ignore"):
    """dilation then erosion (scikit-image ``binary_closing``).  No reference call site."""
    return binary_erosion(binary_dilation(image, footprint, mode=mode), footprint, out=out, mode=mode)


def _plateau_labels(image, connectivity):
    """Label maximal connected sets of equal value (every pixel gets a label)."""
    h, w = image.shape
    idx = np.arange(h * w).reshape(h, w)
    pairs_a, pairs_b = [], []
    shifts = [(0, 1), (1, 0)]
    if connectivity >= 2:
        shifts += [(1, 1), (1, -1)]
    for dy, dx in shifts:
        a = image[: h - dy, max(0, -dx) : w - max(0, dx)]
        b = image[dy:, max(0, dx) : w - max(0, -dx)]
        ia = idx[: h - dy, max(0, -dx) : w - max(0, dx)]
        ib = idx[dy:, max(0, dx) : w - max(0, -dx)]
        eq = a == b
        pairs_a.append(ia[eq])
        pairs_b.append(ib[eq])
    pa = np.concatenate(pairs_a)
    pb = np.concatenate(pairs_b)
    g = coo_matrix((np.ones(len(pa), dtype=np.int8), (pa, pb)), shape=(h * w, h * w))
    n, lab = connected_components(g, directed=False)
    return n, lab.reshape(h, w)


def local_maxima(image, footprint=None, connectivity=None, indices=False, allow_borders=True):
    """Plateau local maxima (scikit-image ``extrema.local_maxima``; refine_boundaries.py:63).

    A maximum is a connected set of equal-valued pixels all of whose neighbours
    (full connectivity by default) are strictly lower.  With ``allow_borders`` the
    image is padded with its own minimum, so a plateau touching the border must
    also be strictly above the global minimum -- a constant image has no maxima.
    Images with a side shorter than 3 return no maxima.
    """
    image = np.asarray(image)
    if image.ndim != 2:
        raise NotImplementedError("oracle restates the 2-D case only")
    if connectivity is None:
        connectivity = image.ndim
    out = np.zeros(image.shape, dtype=bool)
    if any(s < 3 for s in image.shape):
        warnings.warn("maxima can't exist for an image with any dimension smaller 3", stacklevel=2)
        return np.nonzero(out) if indices else out
    h, w = image.shape
    fill = image.min()
    pad = np.pad(image, 1, mode="constant", constant_values=fill)
    higher = np.zeros(image.shape, dtype=bool)  # some neighbour is strictly higher
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if (dy, dx) == (0, 0) or (connectivity < 2 and dy and dx):
                continue
            higher |= pad[1 + dy : 1 + dy + h, 1 + dx : 1 + dx + w] > image
    n, lab = _plateau_labels(image, connectivity)
    bad = np.bincount(lab.ravel(), weights=higher.ravel().astype(np.float64), minlength=n) > 0
    border = np.zeros(image.shape, dtype=bool)
    border[0, :] = border[-1, :] = border[:, 0] = border[:, -1] = True
    if allow_borders:
        # the padding equals the global minimum: a border plateau survives only above it
        bad |= np.bincount(lab.ravel(), weights=(border & (image <= fill)).ravel().astype(np.float64), minlength=n) > 0
    else:
        bad |= np.bincount(lab.ravel(), weights=border.ravel().astype(np.float64), minlength=n) > 0
    out = ~bad[lab]
    return np.nonzero(out) if indices else out


def remove_small_objects(ar, min_size=64, connectivity=1, *, out=None):
    """Drop components with fewer than ``min_size`` pixels (scikit-image
    ``misc.remove_small_objects``; bool input is labelled with ``ndi.label`` at the
    given connectivity, integer input is treated as already labelled).
    The reference applies the same rule to regions (tiff_analysis.py:769-773)."""
    ar = np.asarray(ar)
    if out is None:
        out = ar.copy()
    else:
        out[...] = ar
    if min_size == 0:
        return out
    if out.dtype == bool:
        ccs, _ = ndi.label(ar, ndi.generate_binary_structure(ar.ndim, connectivity))
    else:
        ccs = out
    sizes = np.bincount(ccs.ravel())
    too_small = sizes < min_size
    out[too_small[ccs]] = 0
    return out


def remove_small_holes(ar, area_threshold=64, connectivity=1, *, out=None):
    ar = np.asarray(ar).astype(bool)
    res = ~remove_small_objects(~ar, area_threshold, connectivity)
    if out is not None:
        out[...] = res
        return out
    return res
