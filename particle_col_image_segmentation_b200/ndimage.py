"""Drop-in replacements for the ``scipy.ndimage`` calls on the hot path.

Same names, argument meaning and return dtypes as the calls the reference makes:

* ``median_filter(ds_arr, size=5)``            tiff_analysis.py:122, :643
* ``binary_fill_holes(merged_image)``          tiff_analysis.py:880
* ``distance_transform_edt(~particle_mask)``   tiff_analysis.py:996; refine_boundaries.py:60
* ``label(mask, structure)``                   what ``skimage.measure.label`` delegates to for bool input
* ``binary_dilation / erosion / opening / closing``  north_star morphology rows

numpy in -> numpy out; a CUDA tensor in -> a CUDA tensor out.  2-D images only
(the reference never filters anything else, tiff_analysis.py:727-737).
"""

import numpy as np
import torch

from . import _io, ops


def median_filter(input, size=3, mode="reflect"):
    """``scipy.ndimage.median_filter`` for uint8 / bool 2-D images, odd ``size`` <= 7."""
    if mode != "reflect":
        raise NotImplementedError("only mode='reflect' (scipy's default, the one the reference uses)")
    if isinstance(size, (tuple, list)):
        if len(set(size)) != 1:
            raise NotImplementedError("square windows only")
        size = size[0]
    np_in = _io.is_numpy(input)
    is_bool = (input.dtype == np.bool_) if np_in else (input.dtype == torch.bool)
    t = _io.image_2d(input)
    if is_bool:
        t = t.view(torch.uint8)
    if t.dtype != torch.uint8:
        raise NotImplementedError(f"median_filter: uint8 / bool images only, got {t.dtype}")
    if size == 1:
        out = t.clone()
    else:
        out = ops.median_u8(t, size)
    out = out[0]
    if is_bool:
        out = out.view(torch.bool)
    return _io.back(out, np_in)


def _structure_connectivity(structure, default):
    if structure is None:
        return default
    s = np.asarray(structure) != 0
    if s.shape != (3, 3):
        raise NotImplementedError("label: 3x3 structures only")
    if np.array_equal(s, np.ones((3, 3), bool)):
        return 8
    if np.array_equal(s, np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], bool)):
        return 4
    raise NotImplementedError("label: structure must be the 4- or 8-neighbourhood")


def label(input, structure=None):
    """``scipy.ndimage.label``: ``(int32 labels, num_features)``; default 4-connectivity."""
    np_in = _io.is_numpy(input)
    conn = _structure_connectivity(structure, 4)
    bits, H, W = _io.mask_bits(input)
    labels, counts, _ = ops.label_bits(bits, W, connectivity=conn)
    return _io.back(labels[0], np_in), int(counts[0].item())


def binary_fill_holes(input, structure=None):
    """``scipy.ndimage.binary_fill_holes`` (default structure: 4-connected background)."""
    if structure is not None and _structure_connectivity(structure, 4) != 4:
        raise NotImplementedError("binary_fill_holes: default (4-connected) structure only")
    np_in = _io.is_numpy(input)
    bits, H, W = _io.mask_bits(input)
    return _io.bits_to_bool(ops.fill_holes(bits, W), W, np_in)


def distance_transform_edt(input, return_squared=False):
    """``scipy.ndimage.distance_transform_edt``: float64 distance of every non-zero pixel
    to the nearest zero pixel (unit sampling)."""
    np_in = _io.is_numpy(input)
    bits, H, W = _io.mask_bits(input)
    dist, sq, _ = ops.edt(bits, W, want_dist=not return_squared, want_sq=return_squared)
    return _io.back((sq if return_squared else dist)[0], np_in)


def _default_structure():
    return np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=bool)


def _is_disk(fp):
    fp = np.asarray(fp) != 0
    if fp.ndim != 2 or fp.shape[0] != fp.shape[1] or fp.shape[0] % 2 == 0:
        return None
    r = fp.shape[0] // 2
    L = np.arange(-r, r + 1)
    X, Y = np.meshgrid(L, L)
    return r if np.array_equal(fp, (X**2 + Y**2) <= r * r) else None


def _dilate(bits, W, structure, border_value):
    r = _is_disk(structure)
    if r is not None and r >= 3 and not border_value:
        return ops.dilate_disk(bits, W, r)  # EDT^2 <= r^2, bit-exact
    return ops.dilate(bits, W, structure, border_value)


def _erode(bits, W, structure, border_value):
    return ops.erode(bits, W, structure, border_value)


def binary_dilation(input, structure=None, iterations=1, border_value=0):
    if iterations != 1:
        raise NotImplementedError("iterations != 1")
    np_in = _io.is_numpy(input)
    bits, H, W = _io.mask_bits(input)
    structure = _default_structure() if structure is None else structure
    return _io.bits_to_bool(_dilate(bits, W, structure, border_value), W, np_in)


def binary_erosion(input, structure=None, iterations=1, border_value=0):
    if iterations != 1:
        raise NotImplementedError("iterations != 1")
    np_in = _io.is_numpy(input)
    bits, H, W = _io.mask_bits(input)
    structure = _default_structure() if structure is None else structure
    return _io.bits_to_bool(_erode(bits, W, structure, border_value), W, np_in)


def binary_opening(input, structure=None, iterations=1, border_value=0):
    if iterations != 1:
        raise NotImplementedError("iterations != 1")
    np_in = _io.is_numpy(input)
    bits, H, W = _io.mask_bits(input)
    structure = _default_structure() if structure is None else structure
    tmp = _erode(bits, W, structure, border_value)
    return _io.bits_to_bool(_dilate(tmp, W, structure, border_value), W, np_in)


def binary_closing(input, structure=None, iterations=1, border_value=0):
    if iterations != 1:
        raise NotImplementedError("iterations != 1")
    np_in = _io.is_numpy(input)
    bits, H, W = _io.mask_bits(input)
    structure = _default_structure() if structure is None else structure
    tmp = _dilate(bits, W, structure, border_value)
    return _io.bits_to_bool(_erode(tmp, W, structure, border_value), W, np_in)
