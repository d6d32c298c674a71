"""Drop-in replacements for the ``skimage.morphology`` calls on the hot path.

* ``disk(r)``, ``binary_dilation(mask, footprint)``   tiff_analysis.py:827-828, :990
* ``local_maxima(distance)``                          refine_boundaries.py:63
* ``binary_erosion / opening / closing``, ``remove_small_objects`` -- north_star rows
  without a reference call site; semantics of scikit-image 0.25.2.
"""

import warnings

import numpy as np
import torch

from . import _io, ndimage, ops


def disk(radius, dtype=np.uint8, *, strict_radius=True, decomposition=None):
    """``(2r+1)^2`` footprint with ``x^2 + y^2 <= r^2`` (host-side; a kernel parameter)."""
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    if not strict_radius:
        radius += 0.5
    return np.array((X**2 + Y**2) <= radius**2, dtype=dtype)


def _fp(image, footprint):
    if footprint is None:
        return np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=bool)
    return np.asarray(footprint)


def _finish(res, out):
    if out is not None:
        out[...] = res
        return out
    return res


def binary_dilation(image, footprint=None, out=None, *, mode="ignore"):
    """Outside the image is False unless ``mode='max'``.  A ``disk(r >= 3)`` footprint is
    evaluated as ``EDT(~image)^2 <= r^2`` (bit-exact and independent of r)."""
    return _finish(ndimage.binary_dilation(image, _fp(image, footprint), border_value=int(mode == "max")), out)


def binary_erosion(image, footprint=None, out=None, *, mode="ignore"):
    """Outside the image is True unless ``mode='min'`` (scikit-image 0.25 default)."""
    return _finish(ndimage.binary_erosion(image, _fp(image, footprint), border_value=int(mode != "min")), out)


def binary_opening(image, footprint=None, out=None, *, mode="ignore"):
    np_in = _io.is_numpy(image)
    bits, H, W = _io.mask_bits(image)
    fp = _fp(image, footprint)
    tmp = ndimage._erode(bits, W, fp, int(mode != "min"))
    return _finish(_io.bits_to_bool(ndimage._dilate(tmp, W, fp, int(mode == "max")), W, np_in), out)


def binary_closing(image, footprint=None, out=None, *, mode="ignore"):
    np_in = _io.is_numpy(image)
    bits, H, W = _io.mask_bits(image)
    fp = _fp(image, footprint)
    tmp = ndimage._dilate(bits, W, fp, int(mode == "max"))
    return _finish(_io.bits_to_bool(ndimage._erode(tmp, W, fp, int(mode != "min")), W, np_in), out)


def local_maxima(image, footprint=None, connectivity=None, indices=False, allow_borders=True):
    """Plateau maxima with full connectivity by default; a constant image and images with
    a side shorter than 3 have none (scikit-image ``extrema.local_maxima``)."""
    if footprint is not None or not allow_borders:
        raise NotImplementedError("default footprint and allow_borders=True only")
    if image.ndim != 2:
        raise NotImplementedError("2-D images only")
    if connectivity is None:
        connectivity = 2
    np_in = _io.is_numpy(image)
    H, W = image.shape
    if H < 3 or W < 3:
        warnings.warn("maxima can't exist for an image with any dimension smaller 3", stacklevel=2)
        res = np.zeros((H, W), dtype=bool)
        res = res if np_in else torch.from_numpy(res).to(image.device)
    else:
        t = _io.image_2d(image)
        if t.dtype == torch.bool:
            t = t.view(torch.uint8)
        if t.dtype not in (torch.uint8, torch.uint16, torch.int32, torch.float32, torch.float64):
            t = t.to(torch.float64) if t.dtype.is_floating_point else t.to(torch.int32)
        bits = ops.local_maxima(t, connectivity=8 if connectivity == 2 else 4)
        res = _io.bits_to_bool(bits, W, np_in)
    if indices:
        return np.nonzero(res) if np_in else torch.nonzero(res, as_tuple=True)
    return res


def remove_small_objects(ar, min_size=64, connectivity=1, *, out=None):
    """Drop components with fewer than ``min_size`` pixels (bool input)."""
    np_in = _io.is_numpy(ar)
    is_bool = (ar.dtype == np.bool_) if np_in else (ar.dtype == torch.bool)
    if not is_bool:
        raise NotImplementedError("bool masks only (the pipeline filters labelled regions through the table)")
    bits, H, W = _io.mask_bits(ar)
    if min_size == 0:
        return _finish(_io.bits_to_bool(bits, W, np_in), out)
    res = ops.remove_small(bits, W, min_size, connectivity=8 if connectivity == 2 else 4)
    return _finish(_io.bits_to_bool(res, W, np_in), out)
