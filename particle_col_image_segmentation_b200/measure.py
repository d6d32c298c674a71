"""Drop-in replacements for ``skimage.measure.label`` / ``regionprops``.

* ``label(z_slice)``          tiff_analysis.py:260 (bool), :743 (multi-valued uint8), :829; refine_boundaries.py:64
* ``regionprops(label_im)``   tiff_analysis.py:263, :746 (+ ``.area .centroid .coords[0] .bbox .label``,
                              ``obj["area"]`` :1033 and attribute assignment ``obj.cells = n`` :781)

The per-label numbers come from one device pass (``pcs_region_table``): integer
area / coordinate sums / bounding box / first raster pixel / intensity sum.
Centroids and means are a single float64 division of exact integer sums, which
is what numpy's ``mean`` of integer coordinates evaluates to, bit for bit.
"""

import numpy as np
import torch

from . import _io, ops


def label(label_image, background=None, return_num=False, connectivity=None):
    """``skimage.measure.label`` for 2-D images.

    bool input -> int32 labels (as ``scipy.ndimage.label``); any other dtype ->
    int64 labels of equal-valued, non-``background`` (0) components.  Labels run
    1..N in raster order of each component's first pixel.
    """
    np_in = _io.is_numpy(label_image)
    ndim = label_image.ndim
    if ndim != 2:
        raise NotImplementedError("2-D images only")
    if connectivity is None:
        connectivity = 2
    if connectivity not in (1, 2):
        raise ValueError(f"Connectivity for 2D image should be in [1, ..., 2]. Got {connectivity}.")
    conn = 8 if connectivity == 2 else 4
    is_bool = (label_image.dtype == np.bool_) if np_in else (label_image.dtype == torch.bool)
    if is_bool:
        bits, H, W = _io.mask_bits(label_image)
        labels, counts, _ = ops.label_bits(bits, W, connectivity=conn, dtype=torch.int32)
    else:
        if background not in (None, 0):
            raise NotImplementedError("background other than 0")
        t = _io.image_2d(label_image)
        if t.dtype not in (torch.uint8, torch.uint16, torch.int32, torch.float32, torch.float64):
            t = t.to(torch.int32) if not t.dtype.is_floating_point else t.to(torch.float64)
        labels, counts, _ = ops.label_values(t, connectivity=conn, dtype=torch.int64)
    out = _io.back(labels[0], np_in)
    return (out, int(counts[0].item())) if return_num else out


class LabelHolder:
    """The label image a batch of ``RegionProperties`` was measured on.  Regions made by one ``regionprops`` /
    ``get_cell_positions_and_areas`` call share one holder, and the holder fetches a device image to the host at
    most once.  (A cache keyed by ``data_ptr`` served pixels of a freed image whenever the allocator handed its
    address to the next one.)"""

    __slots__ = ("image", "_host")

    def __init__(self, image):
        self.image, self._host = image, None

    def host(self):
        if self._host is None:
            li = self.image
            self._host = li.cpu().numpy() if isinstance(li, torch.Tensor) else np.asarray(li)
        return self._host


class RegionProperties:
    """Table-backed stand-in for ``skimage.measure._regionprops.RegionProperties``."""

    def __init__(self, label, row, label_image, intensity_image, shape):
        self.label = int(label)
        self._row = row
        self._label_image = label_image if isinstance(label_image, LabelHolder) else LabelHolder(label_image)
        self._intensity_image = intensity_image
        self._shape = shape

    @property
    def area(self):
        return np.float64(self._row[ops.T_AREA]) * 1.0

    @property
    def num_pixels(self):
        return int(self._row[ops.T_AREA])

    @property
    def centroid(self):
        a = np.float64(self._row[ops.T_AREA])
        return (np.float64(self._row[ops.T_SUMY]) / a, np.float64(self._row[ops.T_SUMX]) / a)

    @property
    def bbox(self):
        r = self._row
        return (int(r[ops.T_MINY]), int(r[ops.T_MINX]), int(r[ops.T_MAXY]) + 1, int(r[ops.T_MAXX]) + 1)

    @property
    def slice(self):
        b = self.bbox
        return (slice(b[0], b[2]), slice(b[1], b[3]))

    @property
    def first_pixel(self):
        """``coords[0]`` without materialising the coordinate list (tiff_analysis.py:1042)."""
        return divmod(int(self._row[ops.T_FIRST]), self._shape[1])

    def _host_labels(self):
        return self._label_image.host()  # device labels: fetched once, shared between the regions of one call

    @property
    def image(self):
        return self._host_labels()[self.slice] == self.label

    @property
    def coords(self):
        sl = self.slice
        idx = np.argwhere(self.image)
        return idx + np.array([sl[0].start, sl[1].start])

    @property
    def intensity_sum(self):
        if self._intensity_image is None:
            raise AttributeError("No intensity image specified.")
        return np.float64(self._row[ops.T_SUMI])

    @property
    def intensity_mean(self):
        return self.intensity_sum / np.float64(self._row[ops.T_AREA])

    mean_intensity = intensity_mean

    @property
    def overlap(self):
        return int(self._row[ops.T_OVERLAP])

    def __getitem__(self, key):
        return getattr(self, key)

    def __eq__(self, other):
        return self is other

    __hash__ = object.__hash__


def region_table_host(label_image, intensity_image=None, overlap_mask=None, n_labels=None):
    """One device pass -> ``(int64 table[PCS_TABLE_COLS, n], n)`` on the host."""
    lab = _io.image_2d(label_image)
    if lab.dtype not in (torch.int32, torch.int64):
        if lab.dtype.is_floating_point or lab.dtype == torch.bool:
            raise TypeError("Non-integer label_image types are ambiguous")
        lab = lab.to(torch.int32)
    if n_labels is None:
        n_labels = ops.max_label(lab) if lab.numel() else 0
    cap = max(1, n_labels)
    table = ops.new_table(cap, lab.device)
    inten = None
    if intensity_image is not None:
        inten = _io.image_2d(intensity_image)
        if inten.dtype == torch.bool:
            inten = inten.view(torch.uint8)
        if inten.dtype not in (torch.uint8, torch.uint16):
            raise NotImplementedError("intensity images must be uint8 or uint16")
    ov = None
    if overlap_mask is not None:
        ov = _io.mask_bits(overlap_mask)[0]
    ops.region_table(lab, None, table, intensity=inten, ov_bits=ov)
    return table.cpu().numpy(), n_labels


def regionprops(label_image, intensity_image=None, cache=True, *, overlap_mask=None, n_labels=None, **kwargs):
    """``skimage.measure.regionprops``: one object per present label, in label order."""
    if label_image.ndim != 2:
        raise TypeError("Only 2-D images supported.")
    tab, n = region_table_host(label_image, intensity_image, overlap_mask, n_labels)
    shape = tuple(label_image.shape)
    holder = LabelHolder(label_image)
    return [RegionProperties(i + 1, tab[:, i], holder, intensity_image, shape) for i in np.flatnonzero(tab[ops.T_AREA, :n] > 0).tolist()]
