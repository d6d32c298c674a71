"""CPU restatement of ``refine_boundaries.py`` as a function.  TEST INFRASTRUCTURE ONLY.

The reference file is top-level script code that opens a hard-coded HDF5 file on
import (refine_boundaries.py:28-31), so it cannot be imported; lines :44-73 are
restated here.
"""

import numpy as np
from scipy import ndimage as ndi

from .skimage_shim import measure, morphology, segmentation


def refine_boundaries(boundary_map, threshold=0.5, run_watershed=False):
    """refine_boundaries.py:44-73.

    ``binary_mask = boundary_map < threshold`` (:44-45); ``distance =
    distance_transform_edt(binary_mask)`` (:60); ``local_max =
    local_maxima(distance)`` (:63); ``markers = label(local_max)`` (:64);
    optionally ``watershed(boundary_map, markers, mask=binary_mask)`` (:73, a
    SURVEY 8(f) "next" row).
    """
    boundary_map = np.asarray(boundary_map)
    binary_mask = boundary_map < threshold
    distance = ndi.distance_transform_edt(binary_mask)
    local_max = morphology.local_maxima(distance)
    markers = measure.label(local_max)
    out = {"binary_mask": binary_mask, "distance": distance, "local_max": local_max, "markers": markers}
    if run_watershed:
        out["labels"] = segmentation.watershed(boundary_map, markers, mask=binary_mask)
    return out
