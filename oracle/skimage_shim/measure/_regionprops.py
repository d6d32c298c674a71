"""``skimage.measure._regionprops`` restatement.  TEST INFRASTRUCTURE ONLY.

Only the properties the reference reads are provided: ``label``, ``area``,
``centroid``, ``coords``, ``bbox``, ``slice``, ``image`` (tiff_analysis.py:263-275,
:754-781, :843-863, :1033, :1041-1044) plus ``intensity_mean`` /
``image_intensity`` and an integrated intensity for the north_star's
"mean / integrated intensity" rows.
"""

import numpy as np
from scipy import ndimage as ndi

__all__ = ["RegionProperties", "regionprops"]


class RegionProperties:
    """Lazy per-label view, as in scikit-image 0.25.2 ``RegionProperties``.

    ``area`` is the pixel count times the (unit) pixel area, i.e. a float64;
    ``centroid`` is the float64 mean of the integer pixel coordinates (row, col);
    ``coords`` lists pixels in raster order; ``bbox`` is half-open
    ``(min_row, min_col, max_row, max_col)``.  ``obj["area"]`` and attribute
    assignment (``obj.cells = n``, tiff_analysis.py:781) are supported.
    """

    def __init__(self, slice, label, label_image, intensity_image=None, cache_active=True):
        self.label = int(label)
        self.slice = slice
        self._label_image = label_image
        self._intensity_image = intensity_image
        self._ndim = label_image.ndim

    @property
    def image(self):
        return self._label_image[self.slice] == self.label

    @property
    def area(self):
        return np.sum(self.image) * 1.0

    num_pixels = property(lambda self: int(np.sum(self.image)))

    @property
    def coords(self):
        idx = np.argwhere(self.image)
        off = np.array([self.slice[i].start for i in range(self._ndim)])
        return idx + off

    @property
    def centroid(self):
        return tuple(self.coords.astype(np.float64).mean(axis=0))

    @property
    def bbox(self):
        return tuple([self.slice[i].start for i in range(self._ndim)] + [self.slice[i].stop for i in range(self._ndim)])

    @property
    def image_intensity(self):
        if self._intensity_image is None:
            raise AttributeError("No intensity image specified.")
        return self._intensity_image[self.slice] * self.image

    @property
    def intensity_mean(self):
        return np.mean(self._intensity_image[self.slice][self.image], axis=0)

    mean_intensity = intensity_mean

    @property
    def intensity_sum(self):
        """Integrated intensity (not a scikit-image property; the MATLAB script's
        ``sum(sum(plane.*roimask))``, .m:126-132)."""
        return np.sum(self._intensity_image[self.slice][self.image], dtype=np.float64)

    def __getitem__(self, key):
        return getattr(self, key)

    def __eq__(self, other):
        return self is other

    __hash__ = object.__hash__


def regionprops(label_image, intensity_image=None, cache=True, **kwargs):
    """``scipy.ndimage.find_objects`` per label; absent labels are skipped and the
    list is ordered by label (scikit-image 0.25.2 ``regionprops``)."""
    label_image = np.asarray(label_image)
    if label_image.ndim not in (2, 3):
        raise TypeError("Only 2-D and 3-D images supported.")
    if not np.issubdtype(label_image.dtype, np.integer):
        raise TypeError("Non-integer label_image types are ambiguous")
    regions = []
    for i, sl in enumerate(ndi.find_objects(label_image)):
        if sl is None:
            continue
        regions.append(RegionProperties(sl, i + 1, label_image, intensity_image, cache))
    return regions
