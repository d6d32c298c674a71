"""CPU restatement of the NanoSIMS MATLAB script's per-ROI reductions.

TEST INFRASTRUCTURE ONLY.  Follows
``HCN_nanosims_rois_activity_distance_5iso_YG.m``; MATLAB is not installed, so
this restatement is pinned only against brute-force loops (parity unpinned by
the reference).  Decisions the restatement had to take (SURVEY.md 8a caveats):

* ROI numbering: MATLAB ``regionprops`` on a logical numbers 8-connected
  components in COLUMN-major order of their first pixel (.m:104, :173) --
  restated as raster-order labelling of the transposed mask.
* ``imresize`` (.m:125, :189) of the per-ROI ``holder`` image to the acquisition
  size is MATLAB's default: bicubic (a = -0.5), antialiased when shrinking,
  one dimension at a time, smaller scale first (``imresize`` below, restated from
  the algorithm of imresize.m's ``contributions``).  With equal sizes every tap
  table is the identity and the sums reduce to the masked sums.
* Centroids are ``(x, y)`` = (column, row), 1-based (.m:164-165, :228-229).
* Boundary pixels (``bwboundaries``, .m:290-291) are taken as the mask pixels with
  a 4-neighbour outside the mask; they are ``(row, col)`` 1-based and are compared
  with ``(x, y)`` centroids exactly as the script does (.m:301) -- a latent axis
  swap in the reference that is reproduced, not fixed.
"""

import numpy as np
from scipy import ndimage as ndi

# .m:154 column order of the seven ion planes
PLANES_7 = ("12C", "13C", "14N12C", "15N12C", "16O", "17O", "18O")
# activity = plane[num] / sum(plane[den])   (.m:136-139)
ACTIVITIES_7 = (("13C", (1, (1, 0))), ("15N", (3, (2, 3))), ("17O", (5, (6, 5, 4))), ("18O", (6, (6, 5, 4))))
ACTIVITIES_5 = (("13C", (1, (1, 0))), ("15N", (3, (2, 3))))


def matlab_label(mask):
    """8-connected components numbered in column-major order (MATLAB ``bwconncomp``)."""
    lab, n = ndi.label(np.ascontiguousarray(mask.T), structure=np.ones((3, 3)))
    return np.ascontiguousarray(lab.T), n


def roi_sums(planes, roi_labels, n_rois):
    """.m:126-132, :190-196 -- ``sum(sum(plane .* roimask))`` for every plane and ROI."""
    k = planes.shape[0]
    out = np.zeros((n_rois, k), dtype=np.float64)
    flat = roi_labels.ravel()
    for j in range(k):
        out[:, j] = np.bincount(flat, weights=planes[j].ravel(), minlength=n_rois + 1)[1 : n_rois + 1]
    return out


def activities(sums, spec):
    """.m:136-139 -- isotope fractions from the per-ROI sums."""
    cols = []
    for _, (num, den) in spec:
        d = np.zeros(len(sums))
        for j in den:  # the script adds the planes left to right
            d = d + sums[:, j]
        with np.errstate(divide="ignore", invalid="ignore"):
            cols.append(sums[:, num] / d)
    return np.column_stack(cols) if cols else np.zeros((len(sums), 0))


def roi_centroids_xy(roi_labels, n_rois):
    """.m:164-165 -- ``regionprops(roimask, 'Centroid')``: (x, y), 1-based."""
    flat = roi_labels.ravel()
    h, w = roi_labels.shape
    yy, xx = np.divmod(np.arange(flat.size), w)
    cnt = np.bincount(flat, minlength=n_rois + 1)[1 : n_rois + 1].astype(np.float64)
    sy = np.bincount(flat, weights=yy, minlength=n_rois + 1)[1 : n_rois + 1]
    sx = np.bincount(flat, weights=xx, minlength=n_rois + 1)[1 : n_rois + 1]
    return np.column_stack([sx / cnt + 1.0, sy / cnt + 1.0])


def nearest_between(a_xy, b_xy):
    """.m:260-263 -- ``pdist2`` then the row / column minima."""
    d = np.sqrt(((a_xy[:, None, :] - b_xy[None, :, :]) ** 2).sum(-1))
    return d.min(axis=1), d.min(axis=0)


def boundary_pixels(mask):
    """.m:290-291 -- pixels of ``mask`` with a 4-neighbour outside it, (row, col) 1-based,
    raster order."""
    m = np.pad(mask.astype(bool), 1)
    inner = m[1:-1, 1:-1]
    all4 = m[:-2, 1:-1] & m[2:, 1:-1] & m[1:-1, :-2] & m[1:-1, 2:]
    return np.argwhere(inner & ~all4).astype(np.float64) + 1.0


def min_dist_to_points(xy, pts):
    """.m:301-304 -- ``min(pdist2(positions, bd_position)')``."""
    d2 = ((xy[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    return np.sqrt(d2.min(axis=1))


def _cubic(x):
    """imresize.m ``cubic``: Keys kernel, a = -0.5."""
    ax = np.abs(x)
    ax2, ax3 = ax * ax, ax * ax * ax
    return (1.5 * ax3 - 2.5 * ax2 + 1) * (ax <= 1) + (-0.5 * ax3 + 2.5 * ax2 - 4 * ax + 2) * ((1 < ax) & (ax <= 2))


def resize_contributions(in_length, out_length, antialiasing=True):
    """imresize.m ``contributions`` for the bicubic kernel (width 4): for every output index the 0-based source
    indices (mirrored at the ends) and the weights (normalised to sum 1); all-zero tap columns are dropped."""
    scale = out_length / in_length
    kernel_width = 4.0
    if scale < 1 and antialiasing:
        h = lambda x: scale * _cubic(scale * x)  # noqa: E731
        kernel_width = kernel_width / scale
    else:
        h = _cubic
    x = np.arange(1, out_length + 1, dtype=np.float64)[:, None]
    u = x / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(u - kernel_width / 2)
    P = int(np.ceil(kernel_width)) + 2
    indices = left + np.arange(P)[None, :]
    weights = h(u - indices)
    weights = weights / weights.sum(axis=1, keepdims=True)
    aux = np.concatenate([np.arange(1, in_length + 1), np.arange(in_length, 0, -1)])
    indices = aux[np.mod(indices.astype(np.int64) - 1, aux.size)]
    keep = np.any(weights != 0, axis=0)
    return (indices[:, keep] - 1).astype(np.int32), np.ascontiguousarray(weights[:, keep])


def _apply_taps(a, idx, w, axis):
    """out[i] = sum_p w[i, p] * a[idx[i, p]] along ``axis``, taps in order, one multiply and one add each."""
    a = np.moveaxis(np.asarray(a, dtype=np.float64), axis, 0)
    out = np.zeros((idx.shape[0],) + a.shape[1:])
    for p in range(idx.shape[1]):
        out = out + w[:, p].reshape((-1,) + (1,) * (a.ndim - 1)) * a[idx[:, p]]
    return np.moveaxis(out, 0, axis)


def imresize(a, out_shape, antialiasing=True):
    """MATLAB ``imresize(A, [rows cols])`` with its defaults for a double image (.m:125): bicubic, antialiasing,
    the dimension with the smaller scale first (rows on a tie)."""
    a = np.asarray(a, dtype=np.float64)
    scales = [out_shape[0] / a.shape[0], out_shape[1] / a.shape[1]]
    for dim in sorted((0, 1), key=lambda d: scales[d]):
        idx, w = resize_contributions(a.shape[dim], out_shape[dim], antialiasing)
        a = _apply_taps(a, idx, w, dim)
    return a


def roi_sums_resized(planes, roi_labels, n_rois):
    """.m:122-132 when the ROI image and the acquisition differ in size: per ROI, ``holder`` (1 on the ROI's pixels)
    is resized to the acquisition size and every plane is summed under the resulting fractional mask."""
    out = np.zeros((n_rois, planes.shape[0]))
    for i in range(n_rois):
        roimask = imresize((roi_labels == i + 1).astype(np.float64), planes.shape[1:])
        for k in range(planes.shape[0]):
            out[i, k] = (planes[k] * roimask).sum()
    return out


def analyse(planes, red_mask, green_mask, agg_mask, raster=19.0, acq=512.0):
    """Whole-script restatement: rows ``[set, i, sums..., act..., act*100..., x, y,
    nearest_um, boundary_um]`` for the red then the green ROIs
    (.m:154, :216, :249-252, :265-268, :306-309)."""
    k = planes.shape[0]
    spec = ACTIVITIES_7 if k >= 7 else ACTIVITIES_5
    rows, pos = [], []
    for set_id, mask in ((1, red_mask), (2, green_mask)):
        lab, n = matlab_label(mask)
        s = roi_sums(planes, lab, n) if lab.shape == planes.shape[1:] else roi_sums_resized(planes, lab, n)
        act = activities(s, spec)
        xy = roi_centroids_xy(lab, n)
        rows.append(np.column_stack([np.full(n, float(set_id)), np.arange(1, n + 1, dtype=np.float64), s, act, act * 100.0]))
        pos.append(xy)
    a_near, b_near = nearest_between(pos[0], pos[1])
    bd = boundary_pixels(agg_mask)
    scale = raster / acq  # ./(512/raster), .m:267, :308
    near = np.concatenate([a_near, b_near]) * scale
    bdist = np.concatenate([min_dist_to_points(pos[0], bd), min_dist_to_points(pos[1], bd)]) * scale
    return np.column_stack([np.concatenate(rows), np.concatenate(pos), near, bdist])


def activity_vs_distance(activity, distance, edges):
    """north_star "activity-vs-distance binning" (no reference call site: the script
    writes one CSV row per ROI).  ``np.digitize`` + ``np.bincount``: per bin the ROI
    count, the activity sum and the mean."""
    idx = np.digitize(distance, edges)
    nb = len(edges) + 1
    cnt = np.bincount(idx, minlength=nb).astype(np.float64)
    tot = np.bincount(idx, weights=activity, minlength=nb)
    with np.errstate(divide="ignore", invalid="ignore"):
        return cnt, tot, tot / cnt


# ---------------------------------------------------------------- ratio images (.m:17-69)
def imgaussfilt(a, sigma):
    """``imgaussfilt(A, sigma)`` (.m:43): replicate border, columns then rows, products summed in tap order
    (one rounding per operation, as the device kernel does).  MATLAB's own summation order is not documented."""
    a = np.asarray(a, dtype=np.float64)
    r = int(np.ceil(2.0 * sigma))
    import math

    w = [math.exp(-float(t * t) / (2.0 * sigma * sigma)) for t in range(-r, r + 1)]  # libm exp, as the C side
    s = 0.0
    for v in w:
        s += v
    w = [v / s for v in w]
    h, wd = a.shape
    pad = np.pad(a, ((r, r), (0, 0)), mode="edge")
    tmp = np.zeros_like(a)
    for t in range(2 * r + 1):
        tmp = tmp + w[t] * pad[t : t + h]
    pad = np.pad(tmp, ((0, 0), (r, r)), mode="edge")
    out = np.zeros_like(a)
    for t in range(2 * r + 1):
        out = out + w[t] * pad[:, t : t + wd]
    return out


def scaled_uint8(num, dens=()):
    """``uint8(R .* (255 / max(R(:))))``, MATLAB conversion: round half away from zero, saturate, NaN -> 0."""
    r = np.asarray(num, dtype=np.float64)
    if dens:
        den = np.asarray(dens[0], dtype=np.float64)
        for d in dens[1:]:
            den = den + np.asarray(d, dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = r / den
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        v = r * (255.0 / np.nanmax(r))
        fl = np.floor(np.abs(v))
        q = np.sign(v) * (fl + (np.abs(v) - fl >= 0.5))  # round half away from zero, exactly
    q = np.where(np.isnan(q), 0.0, np.clip(q, 0.0, 255.0))
    return q.astype(np.uint8)


def ratio_images(ions):
    """.m:17-69 with the script's variable names (see the device mirror for the line map)."""
    raw = {k: np.asarray(v, dtype=np.float64)[1:-1, 1:-1] for k, v in ions.items()}
    g1 = {k: imgaussfilt(raw[k], 1) for k in ("15N12C", "14N12C", "16O", "17O", "18O") if k in raw}
    g15 = {k: imgaussfilt(raw[k], 1.5) for k in ("12C", "13C", "Esi") if k in raw}
    out = {}
    for k, name in (("12C", "C12img"), ("13C", "C13img"), ("14N12C", "N14C12img"), ("15N12C", "N15C12img"), ("16O", "O16img"), ("17O", "O17img"), ("18O", "O18img")):
        if k in raw:
            out[name] = scaled_uint8(raw[k])
    if "15N12C" in raw and "14N12C" in raw:
        out["N15ratioimg"] = scaled_uint8(g1["15N12C"], (g1["15N12C"], g1["14N12C"]))
        out["N15ratimg"] = scaled_uint8(raw["15N12C"], (raw["15N12C"], raw["14N12C"]))
    if "12C" in raw and "13C" in raw:
        out["C13ratioimg"] = scaled_uint8(g15["13C"], (g15["13C"], g15["12C"]))
        out["C13ratimg"] = scaled_uint8(raw["13C"], (raw["13C"], raw["12C"]))
        if "14N12C" in raw:
            out["N14C12C12ratio"] = scaled_uint8(g1["14N12C"], (g15["12C"],))
    if all(k in raw for k in ("16O", "17O", "18O")):
        dens_g = (g1["18O"], g1["17O"], g1["16O"])
        dens_r = (raw["18O"], raw["17O"], raw["16O"])
        out["O17ratioimg"] = scaled_uint8(g1["17O"], dens_g)
        out["O18ratioimg"] = scaled_uint8(g1["18O"], dens_g)
        out["O17ratimg"] = scaled_uint8(raw["17O"], dens_r)
        out["O18ratimg"] = scaled_uint8(raw["18O"], dens_r)
    if "Esi" in raw and "14N12C" in raw:
        out["N14C12ESIratio"] = scaled_uint8(raw["14N12C"], (raw["Esi"],))
    return out
