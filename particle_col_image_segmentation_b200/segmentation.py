"""``skimage.segmentation`` drop-in for the one call the path makes: ``watershed``.

``refine_boundaries.py:73``: ``labels = watershed(boundary_map, markers, mask=binary_mask)``.
The flood runs on the device (``pcs_watershed_f64``): every pixel joins the 4-neighbour of smallest
bottleneck cost, which is what scikit-image's sequential priority flood produces whenever no two
competing pixels share a value; with ties the result is still a valid flood but the tie-break (hops, then
label) is not scikit-image's insertion age (see ``csrc/pcs_watershed.cu``).
"""

import ctypes

import numpy as np
import torch

from . import _io, _lib, ops

__all__ = ["watershed"]


def watershed(image, markers=None, connectivity=1, offset=None, mask=None, compactness=0, watershed_line=False, return_sweeps=False):
    """Marker-controlled watershed of a 2-D image; int32 labels, 0 where the flood does not reach.
    Only what the reference uses is implemented: explicit ``markers``, ``connectivity=1``, no compactness,
    no watershed line.  numpy in -> numpy out, CUDA tensors in -> CUDA tensor out."""
    if markers is None or isinstance(markers, (int, np.integer)):
        raise NotImplementedError("watershed needs an explicit marker image (refine_boundaries.py:64)")
    if connectivity != 1 or offset is not None or compactness or watershed_line:
        raise NotImplementedError("only connectivity=1, compactness=0, watershed_line=False (refine_boundaries.py:73)")
    np_in = _io.is_numpy(image)
    img = _io.to_device(image)
    if img.dim() != 2:
        raise ValueError(f"expected a 2-D image, got shape {tuple(img.shape)}")
    if img.dtype != torch.float64:
        img = img.to(torch.float64)  # exact for every integer and float32 input: the order of values is kept
    mk = _io.to_device(markers)
    if tuple(mk.shape) != tuple(img.shape):
        raise ValueError(f"markers {tuple(mk.shape)} and image {tuple(img.shape)} differ in shape")
    mk = mk.to(torch.int32)
    H, W = (int(v) for v in img.shape)
    bits = None
    if mask is not None:
        bits, mh, mw = _io.mask_bits(mask)
        if (mh, mw) != (H, W):
            raise ValueError("mask and image differ in shape")
    lib = _lib.load()
    n = lib.pcs_watershed_workspace_bytes(1, H, W)
    ws = torch.empty(n, dtype=torch.uint8, device=img.device)
    out = torch.empty((H, W), dtype=torch.int32, device=img.device)
    sweeps = ctypes.c_int(0)
    _lib.call("pcs_watershed_f64", ops._p(img), ops._p(mk), ops._p(bits), ops._p(out), 1, H, W, 0, ctypes.addressof(sweeps), ops._p(ws), n, ops._stream())
    res = _io.back(out, np_in)
    return (res, sweeps.value) if return_sweeps else res
