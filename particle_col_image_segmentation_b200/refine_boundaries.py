"""Device-backed ``refine_boundaries.py`` as a function.

The reference file is top-level script code bound to a hard-coded HDF5 path
(refine_boundaries.py:28-31); its pixel work (:44-64) is: threshold the ilastik
boundary-probability map, EDT of the foreground, plateau local maxima of the
distance, label the maxima as watershed markers, and flood the boundary map from them (:73, in code
the author marks as not working yet, :54; ``segmentation.watershed``).
"""

import numpy as np
import torch

from . import _io, ops


def refine_boundaries(boundary_map, threshold=0.5, run_watershed=False, return_sweeps=False):
    """-> dict(binary_mask bool, distance float64, local_max bool, markers int32[, labels int32]).

    ``binary_mask = boundary_map < threshold``            refine_boundaries.py:44-45
    ``distance = distance_transform_edt(binary_mask)``    refine_boundaries.py:60
    ``local_max = local_maxima(distance)``                refine_boundaries.py:63
    ``markers = label(local_max)``                        refine_boundaries.py:64
    ``labels = watershed(boundary_map, markers, mask=binary_mask)``   refine_boundaries.py:73 (``run_watershed``)

    Maxima are found on the exact integer squared distance: sqrt is strictly
    monotone, so plateaus and strict inequalities are the same as on ``distance``.
    """
    np_in = _io.is_numpy(boundary_map)
    t = _io.image_2d(boundary_map)
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    H, W = int(t.shape[1]), int(t.shape[2])
    bits = ops.compare(t, "<", float(threshold))[0]
    dist, sq, _ = ops.edt(bits, W, want_dist=True, want_sq=True)
    if H < 3 or W < 3:
        maxima = torch.zeros_like(bits)
    else:
        maxima = ops.local_maxima(sq, connectivity=8)
    markers, counts, _ = ops.label_bits(maxima, W, connectivity=8, dtype=torch.int32)
    out = {
        "binary_mask": _io.bits_to_bool(bits, W, np_in),
        "distance": _io.back(dist[0], np_in),
        "local_max": _io.bits_to_bool(maxima, W, np_in),
        "markers": _io.back(markers[0], np_in),
    }
    if run_watershed:
        from . import segmentation

        lab, sweeps = segmentation.watershed(t[0], markers[0], mask=ops.unpack(bits, W, torch.bool)[0], return_sweeps=True)
        out["labels"] = _io.back(lab, np_in)
        if return_sweeps:
            out["sweeps"] = sweeps  # relaxation sweeps the flood took (diagnostic; not a reference output)
    return out


def refine_boundaries_h5(file_path, channel=3, dataset="exported_data", **kwargs):
    """refine_boundaries.py:28-34: load ``exported_data`` from the ilastik probability export, take the boundary channel
    (channel-first, index 3 in the reference) and run ``refine_boundaries`` on it."""
    from . import h5_io

    with h5_io.File(file_path, "r") as f:
        probabilities = np.array(f[dataset])
    return refine_boundaries(probabilities[channel], **kwargs)
