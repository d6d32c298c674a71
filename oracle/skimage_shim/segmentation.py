"""``skimage.segmentation`` restatement (watershed).  TEST INFRASTRUCTURE ONLY.

``refine_boundaries.py:73`` calls ``watershed(boundary_map, markers,
mask=binary_mask)`` in code its author marks as not yet working
(``refine_boundaries.py:54``).  SURVEY.md section 8(f) lists it as a "next" row;
this restatement exists so the script's tail can be exercised on CPU.
"""

import heapq

import numpy as np

__all__ = ["watershed"]


def watershed(image, markers=None, connectivity=1, offset=None, mask=None, compactness=0, watershed_line=False):
    """Priority-flood watershed: pop the lowest (value, age) pixel, give each
    unlabeled in-mask neighbour its label, push it with the next age
    (scikit-image ``_watershed_cy.watershed_raveled`` with compactness 0).
    Ties between equal-valued seeds are resolved here by raster order, which
    scikit-image's binary heap does not guarantee: parity of this function is
    unpinned."""
    image = np.asarray(image)
    if markers is None or compactness or watershed_line:
        raise NotImplementedError
    h, w = image.shape
    out = np.array(markers, dtype=np.int32)
    if mask is None:
        mask = np.ones(image.shape, dtype=bool)
    out[~mask] = 0
    if connectivity == 1:
        nb = [(-1, 0), (0, -1), (0, 1), (1, 0)]
    else:
        nb = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dy, dx) != (0, 0)]
    heap = []
    age = 0
    for y, x in np.argwhere(out > 0):
        heapq.heappush(heap, (image[y, x], 0, int(y), int(x)))
    while heap:
        _, _, y, x = heapq.heappop(heap)
        for dy, dx in nb:
            yy, xx = y + dy, x + dx
            if not (0 <= yy < h and 0 <= xx < w) or not mask[yy, xx] or out[yy, xx]:
                continue
            age += 1
            out[yy, xx] = out[y, x]
            heapq.heappush(heap, (image[yy, xx], age, yy, xx))
    return out
