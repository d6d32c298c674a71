"""Drop-in ``skimage.filters.threshold_otsu`` for uint16 / uint8 images (north_star;
the reference imports ``filters`` at refine_boundaries.py:22 but never calls it)."""

import numpy as np
import torch

from . import _io, ops


def threshold_otsu(image=None, nbins=256, *, hist=None):
    """One bin per integer between min and max; returns the threshold as a numpy
    scalar of the image dtype (the mask is ``image > threshold``)."""
    if hist is not None:
        raise NotImplementedError("precomputed histograms")
    np_in = _io.is_numpy(image)
    t = _io.to_device(image)
    if t.dtype == torch.uint8:
        t = t.to(torch.int32).to(torch.uint16)
    if t.dtype != torch.uint16:
        raise NotImplementedError(f"threshold_otsu: uint8 / uint16 images only, got {t.dtype}")
    thr = ops.otsu_u16(t.reshape(1, 1, -1) if t.dim() != 2 else t.unsqueeze(0))
    v = int(thr[0].item())
    if np_in:
        return np.asarray(image).dtype.type(v)
    return v
